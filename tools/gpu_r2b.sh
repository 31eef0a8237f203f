#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -25 gpurun_out/r2b_pytest.log
timeout 600 python tools/sweep_dev.py --gemm-dbg --precisions fp16 > gpurun_out/r2b_sweep.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:typed_linear_tc -c 6 -f -o gpurun_out/r2b_gemm \
    python tools/prof_gemm.py fp16 > gpurun_out/r2b_ncu_gemm.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 600 gpurun_out/r2b_bench.json
