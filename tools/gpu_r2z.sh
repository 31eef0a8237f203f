#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -5
python tools/bench_train.py --batch 16 --steps 5 --warmup 2 2>&1 | tail -1
