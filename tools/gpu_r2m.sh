#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r2m_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
timeout 300 python tools/forward_breakdown.py > gpurun_out/r2m_breakdown.jsonl 2>&1
cat gpurun_out/r2m_breakdown.jsonl | tail -14
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline_dense']['kernel_ms'], d['roofline']['kernel_ms'])"
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1; lscpu | head -25 >> gpurun_out/r2m_topo.txt; numactl -H >> gpurun_out/r2m_topo.txt 2>&1; nproc >> gpurun_out/r2m_topo.txt
