#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2p_pytest.log
tail -12 gpurun_out/r2p_pytest.log
timeout 600 python tools/bench_train.py --batch 16 --steps 3 > gpurun_out/r2p_train.json 2>&1
tail -1 gpurun_out/r2p_train.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2p_launches_train.csv \
    python tools/bench_train.py --batch 8 --steps 1 --warmup 1 > gpurun_out/r2p_ncu_train.log 2>&1
