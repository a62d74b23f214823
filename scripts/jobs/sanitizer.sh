#!/bin/bash
# compute-sanitizer (TOOL=memcheck|racecheck|synccheck, default memcheck) over the kernel-level GPU tests; ${@} = test files / -k expression
mkdir -p gpurun_out
TESTS=${@:-tests/test_gpu_shards.py tests/test_gpu_train_ops.py}
timeout 1500 /usr/local/cuda/bin/compute-sanitizer --tool ${TOOL:-memcheck} --report-api-errors no --error-exitcode 7 --print-limit 20 python -m pytest $TESTS -m gpu -q -x --timeout 1400 > gpurun_out/sanitizer_${TOOL:-memcheck}.log 2>&1
echo "${TOOL:-memcheck} exit $?" >> gpurun_out/sanitizer_${TOOL:-memcheck}.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|check exit|Race|Hazard|at vbg::|Saved host backtrace" gpurun_out/sanitizer_${TOOL:-memcheck}.log | head -40
