#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/forward_breakdown.py > gpurun_out/r2f_breakdown.jsonl 2>&1
cat gpurun_out/r2f_breakdown.jsonl | tail -16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err
tail -3 gpurun_out/r2f_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
t=d['train_step']; print(t['ms_per_step'], t['graphs_per_s']); print(t['per_rank'])
PY
