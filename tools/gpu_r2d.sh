#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r2d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -25 gpurun_out/r2d_pytest.log
timeout 300 python tools/forward_breakdown.py > gpurun_out/r2d_breakdown.jsonl 2>&1
cat gpurun_out/r2d_breakdown.jsonl | tail -20
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -5 gpurun_out/r2d_bench.err
tail -c 3500 gpurun_out/r2d_bench.json
