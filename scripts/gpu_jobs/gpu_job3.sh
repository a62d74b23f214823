#!/bin/bash
# GPU job 3: validate the bf16x3 tensor-core mode (GEMM / conv / strided conv / stem), bench it, launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -60 > gpurun_out/pytest_ops.log
echo "pytest ops exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert " | tail -60 > gpurun_out/pytest_fwd.log
timeout 600 python scripts/tc_probe.py > gpurun_out/tc_probe3.log 2>&1; echo "probe exit $?" >> gpurun_out/tc_probe3.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench exit $?" >> gpurun_out/bench_bf16x3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_bf16x3.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench3.log 2>&1
tail -15 gpurun_out/pytest_ops.log; cat gpurun_out/pytest_fwd.log; tail -40 gpurun_out/tc_probe3.log; head -c 1500 gpurun_out/bench_bf16x3.json; tail -3 gpurun_out/bench_bf16x3.err
