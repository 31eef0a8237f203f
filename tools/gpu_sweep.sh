#!/bin/bash
# Development GPU visit: parity tests, knob sweep, ncu capture of the edge-attention kernel.
TAG=${1:-sweep}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/sweep_dev.py ${SWEEP_ARGS} > gpurun_out/${TAG}_sweep.jsonl 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json
timeout 300 python tools/prof_e2e.py > gpurun_out/${TAG}_prof_e2e.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 2 -f -o gpurun_out/${TAG}_attn_fwd \
    python tools/prof_attn.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
