#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_builder.py tests/test_node_sharded.py -m gpu -q > gpurun_out/r2n_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
tail -15 gpurun_out/r2n_pytest.log
timeout 600 python tools/bench_node_sharded.py --check > gpurun_out/r2n_c4_1gpu.json 2>&1
tail -2 gpurun_out/r2n_c4_1gpu.json
