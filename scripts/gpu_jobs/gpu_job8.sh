#!/bin/bash
# GPU job 8: attention v3 (two softmax groups), fused seg CE, fixed-cost probe, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_ops.log
echo "pytest ops exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_ops.log
timeout 1500 python -m pytest tests/test_gpu_forward.py -m gpu -q -s 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch" | tail -40 > gpurun_out/pytest_fwd.log
timeout 300 python scripts/fixed_cost.py > gpurun_out/fixed_cost.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; echo "bench exit $?" >> gpurun_out/bench_r1_f.err
VBG_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_f.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench8.log 2>&1
tail -8 gpurun_out/pytest_ops.log; tail -8 gpurun_out/pytest_fwd.log; cat gpurun_out/fixed_cost.log; head -c 1400 gpurun_out/bench_r1_f.json; tail -3 gpurun_out/bench_r1_f.err
