#!/bin/bash
# GPU job 6: split-QKV + TMA-fed attention (MN-major V), graph test, full ncu of the BERT GEMMs and the new attention.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_ops.log
echo "pytest ops exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert |mismatch" | tail -40 > gpurun_out/pytest_fwd.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.err; echo "bench exit $?" >> gpurun_out/bench_r1_d.err
VBG_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_d.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench6.log 2>&1
VBG_CUDA_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_split" -s 2 -c 2 -o gpurun_out/prof_attn2 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full6a.log 2>&1
VBG_CUDA_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc3_kernel" -s 4 -c 4 -o gpurun_out/prof_tc3p \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full6b.log 2>&1
tail -8 gpurun_out/pytest_ops.log; tail -6 gpurun_out/pytest_fwd.log; head -c 1300 gpurun_out/bench_r1_d.json; tail -3 gpurun_out/bench_r1_d.err
