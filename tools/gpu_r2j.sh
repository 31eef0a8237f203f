#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_multi.py > gpurun_out/r2j_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
grep -E "config3|passed|failed|Error|error" gpurun_out/r2j_pytest.log | tail -30
tail -30 gpurun_out/r2j_pytest.log
for prec in bf16 bf16x3 fp16; do
WSI_HGT_PRECISION=$prec timeout 600 python tools/bench_hgt.py --precision $prec >> gpurun_out/r2j_hgt.jsonl 2>&1
done
cat gpurun_out/r2j_hgt.jsonl | tail -5
