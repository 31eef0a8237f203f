#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py tests/test_gpu_golden.py tests/test_gpu_ops.py -m gpu -q -x -k "hgt or HGT or segment or rel_transform or config3" 2>&1 | tail -8
python tools/bench_hgt.py --precision bf16 2>&1 | tail -1
python tools/bench_hgt.py --precision fp16 --check 2 2>&1 | tail -1
