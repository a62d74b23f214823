#!/bin/bash
# GPU job 2: validate the tcgen05 TF32 path, bench it, launch list + full ncu capture of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "probe exit $?" >> gpurun_out/tc_probe.log
timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?" >> gpurun_out/bench_tf32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_tf32.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 6 -o gpurun_out/prof_gemm_tc \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1
tail -40 gpurun_out/tc_probe.log; tail -8 gpurun_out/pytest_gpu.log; head -c 2500 gpurun_out/bench_tf32.json; tail -3 gpurun_out/bench_tf32.err
