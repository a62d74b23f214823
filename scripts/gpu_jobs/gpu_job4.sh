#!/bin/bash
# GPU job 4: in-place bf16x3 conversion, tcgen05 attention, full ncu capture of the bf16x3 GEMM and attention kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_ops.log
echo "pytest ops exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert " | tail -60 > gpurun_out/pytest_fwd.log
timeout 600 python scripts/tc_probe.py 2>&1 | grep -E "^\[gemm |^\[conv B8|^\[conv B1024|stem|exit" | tail -30 > gpurun_out/tc_probe4.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16x3_b.json 2> gpurun_out/bench_bf16x3_b.err; echo "bench exit $?" >> gpurun_out/bench_bf16x3_b.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_bf16x3_b.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc3_kernel|attention_tc" -s 60 -c 8 -o gpurun_out/prof_tc3 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full4.log 2>&1
tail -12 gpurun_out/pytest_ops.log; cat gpurun_out/pytest_fwd.log; cat gpurun_out/tc_probe4.log; head -c 1200 gpurun_out/bench_bf16x3_b.json; tail -3 gpurun_out/bench_bf16x3_b.err
