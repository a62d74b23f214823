#!/bin/bash
# First GPU job: parity tests, smoke, short bench, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench exit $?" >> gpurun_out/bench_fp32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_fp32.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -q -k "index_map or scatter or roi or segment or transform" > gpurun_out/sanitizer.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_fp32.json | head -c 3000; tail -3 gpurun_out/bench_fp32.err; tail -5 gpurun_out/sanitizer.log
