#!/bin/bash
# one ncu --set full capture of the kernels matching the regex in $1 (default: the FFN-up pair GEMM) during a short bench run
mkdir -p gpurun_out
RX=${1:-gemm_ps2_kernel}
VBG_CUDA_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -c ${2:-4} -o gpurun_out/ncu_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-library-bar --no-input-pipeline --no-serving > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?"
tail -5 gpurun_out/ncu_full.log
