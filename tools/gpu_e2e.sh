#!/bin/bash
TAG=${1:-e2e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slide_io.py tests/test_gpu_models.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g  ms %.4f  eager %.4f  e2e %.4g (%.3f ms; sync %.3f ms)" % (d["value"], d["ms_per_step"], d["config"]["eager_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["sync_ms_per_step"]))
PY
tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python tools/sweep_dev.py --gemm-dbg > gpurun_out/${TAG}_sweep.jsonl 2>&1
