#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_models.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
tail -8 gpurun_out/r2o_pytest.log
timeout 600 python tools/bench_train.py --batch 16 --steps 3 > gpurun_out/r2o_train.json 2>&1
tail -1 gpurun_out/r2o_train.json
