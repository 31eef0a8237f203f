#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_slide_io.py tests/test_data.py tests/test_explainer.py -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2s_pytest.log
tail -12 gpurun_out/r2s_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
tail -3 gpurun_out/r2s_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2s_bench.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e'])[:900])"
