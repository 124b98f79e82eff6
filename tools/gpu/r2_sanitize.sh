#!/bin/bash
# compute-sanitizer memcheck over parity tests (small shapes): $1 = test file(s), $2 = -k filter
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 0 python -m pytest ${1:-tests/test_kernels_gpu.py} -q -m gpu -p no:cacheprovider -x ${2:+-k "$2"} > gpurun_out/sanitize.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/sanitize.log | head -20
