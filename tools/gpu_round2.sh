#!/bin/bash
# One GPU-box visit (round 2): full parity suite, smoke, contract bench (both arms), per-kernel bench, forward breakdown.
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 400 python tools/bench_kernels.py > gpurun_out/${TAG}_kernels.jsonl 2>&1
timeout 300 python tools/bench_hgt.py --precision bf16 > gpurun_out/${TAG}_hgt_bf16.json 2>&1
timeout 300 python tools/bench_hgt.py --precision fp16 --check 2 > gpurun_out/${TAG}_hgt_fp16.json 2>&1
timeout 300 python tools/bench_hgt.py --precision bf16x3 > gpurun_out/${TAG}_hgt_bf16x3.json 2>&1
timeout 300 python tools/hgt_breakdown.py bf16 > gpurun_out/${TAG}_hgt_breakdown.json 2>&1
timeout 300 python tools/bench_train.py --batch 16 --steps 5 --warmup 2 > gpurun_out/${TAG}_train16.json 2>&1
timeout 300 python tools/bench_node_sharded.py --check > gpurun_out/${TAG}_config4_1gpu.json 2>&1
timeout 300 python tools/forward_breakdown.py > gpurun_out/${TAG}_breakdown.jsonl 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'dense', d['roofline_dense']['frac'], 'attn', d['roofline']['frac'])
print('train', d.get('train_step',{}).get('ms_per_step'), 'cpu', d.get('cpu_baseline'))
r=json.load(open('gpurun_out/${TAG}_bench_ref.json')); print('ref', r['value'])
PY
