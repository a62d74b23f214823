#!/bin/bash
# GPU job 5: persistent bf16x3 GEMM + CUDA-graph replay; full ncu captures of the attention and BERT-shaped GEMM kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_ops.log
echo "pytest ops exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert |mismatch" | tail -40 > gpurun_out/pytest_fwd.log
timeout 600 python scripts/tc_probe.py 2>&1 | grep -E "^\[gemm |^\[conv B8|^\[conv B1024|stem|exit|Error|error" | tail -30 > gpurun_out/tc_probe5.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; echo "bench exit $?" >> gpurun_out/bench_r1_c.err
VBG_CUDA_GRAPHS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_c_nograph.json 2> gpurun_out/bench_r1_c_nograph.err
VBG_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_c.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench5.log 2>&1
VBG_CUDA_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_tc" -s 2 -c 2 -o gpurun_out/prof_attn \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full5a.log 2>&1
VBG_CUDA_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc3_kernel<\(int\)(192|256)" -s 8 -c 4 -o gpurun_out/prof_tc3p \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full5b.log 2>&1
tail -6 gpurun_out/pytest_ops.log; cat gpurun_out/pytest_fwd.log | tail -8; cat gpurun_out/tc_probe5.log; head -c 1500 gpurun_out/bench_r1_c.json; tail -3 gpurun_out/bench_r1_c.err; head -c 600 gpurun_out/bench_r1_c_nograph.json
