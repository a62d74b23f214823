#!/bin/bash
# GPU job 9: re-verify HEAD after container re-creation: all gpu tests, smoke, bench, launch list, full ncu of scatter/ROI/attention/GEMM.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[cfg|^\[|passed|failed|Error|assert |mismatch" | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python scripts/fixed_cost.py > gpurun_out/fixed_cost.log 2>&1
timeout 300 python scripts/bn_sweep.py > gpurun_out/bn_sweep.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; echo "bench exit $?" >> gpurun_out/bench_r1_g.err
VBG_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_g.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench9.log 2>&1
VBG_CUDA_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention|roi_align|grid_scatter" -s 6 -c 6 -o gpurun_out/prof_hbm_attn \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full9a.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/fixed_cost.log; cat gpurun_out/bn_sweep.log; head -c 1600 gpurun_out/bench_r1_g.json; tail -3 gpurun_out/bench_r1_g.err
