#!/bin/bash
# per-launch durations of a short bench run (cold-cache, serialised: use the kernels' SHARES, not the absolutes)
mkdir -p gpurun_out
VBG_CUDA_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-roofline --no-cpu-baseline --no-train --no-library-bar --no-input-pipeline --no-serving > gpurun_out/ncu_launches.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv | head -60
