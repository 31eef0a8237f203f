#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -40 gpurun_out/r2i_pytest.log
