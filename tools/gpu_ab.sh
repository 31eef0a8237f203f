#!/bin/bash
# A/B of a development knob on the contract bench:  bash tools/gpu_ab.sh TAG "ENV=1"
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/sweep_dev.py > gpurun_out/${TAG}_sweep.jsonl 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_a.json 2> gpurun_out/${TAG}_bench_a.err
env $2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_b.json 2> gpurun_out/${TAG}_bench_b.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_a2.json 2>> gpurun_out/${TAG}_bench_a.err
for f in a b a2; do python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$f.json"))
print("$f", "value %.4g  ms %.4f  eager %.4f  e2e %.4g (%.3f ms)  attn %.1f us  gemm %.1f us" % (d["value"], d["ms_per_step"], d["config"]["eager_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"]*1e3, d["roofline_dense"]["kernel_ms"]*1e3))
PY
done
tail -3 gpurun_out/${TAG}_bench_a.err
