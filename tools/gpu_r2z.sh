#!/bin/bash
python -m pytest tests/test_gpu_models.py tests/test_gpu_ops.py -m gpu -q -x -k "hgt or HGT or seg or config3" 2>&1 | tail -4
python tools/bench_hgt.py --precision bf16 2>&1 | tail -1
python tools/bench_hgt.py --precision fp16 --check 2 2>&1 | tail -1
