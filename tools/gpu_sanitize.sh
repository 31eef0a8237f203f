#!/bin/bash
# compute-sanitizer memcheck over the kernel-level parity tests (bounded; the suite is small-shaped)
TAG=${1:-san}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py tests/test_node_sharded.py tests/test_gpu_builder.py -x -q -m gpu -k "not full_shape" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${TAG}_memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|exit" gpurun_out/${TAG}_memcheck.log | tail -12
