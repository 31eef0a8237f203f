#!/bin/bash
# Round-2 GPU visit A: parity under the new operand formats, GEMM attribution sweep, per-kernel bench, contract bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 600 python tools/sweep_dev.py --gemm-dbg > gpurun_out/r2a_sweep.jsonl 2>&1
tail -5 gpurun_out/r2a_sweep.jsonl
timeout 600 python tools/bench_kernels.py > gpurun_out/r2a_kernels.jsonl 2>&1
tail -3 gpurun_out/r2a_kernels.jsonl
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 1500 gpurun_out/r2a_bench.json
