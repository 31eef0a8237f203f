#!/bin/bash
# A/B/A of an environment knob on bench.py's e2e leg within one box:  bash tools/gpu_ab2.sh "ENV=1"
for v in "" "$1" "" "$1"; do
env $v python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$v]', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], '%.3f ms/slide' % d['e2e']['ms_per_step'], 'sync %.3f' % d['e2e']['sync_ms_per_step'])"
done
