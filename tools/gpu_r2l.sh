#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench8.json 2> gpurun_out/r2l_bench8.err
tail -3 gpurun_out/r2l_bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench8.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
t=d['train_step']; print(t['ms_per_step'], t['graphs_per_s'])
for r in t['per_rank']: print(r)
PY
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r2l_pytest_multi.log 2>&1
tail -5 gpurun_out/r2l_pytest_multi.log
