#!/bin/bash
# One GPU-box visit: parity tests, contract bench, per-kernel bench, train bench, ncu launch list + full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python tools/bench_kernels.py > gpurun_out/${TAG}_kernels.jsonl 2>&1
cat gpurun_out/${TAG}_kernels.jsonl
timeout 600 python tools/bench_train.py --batch 16 --steps 3 > gpurun_out/${TAG}_train.json 2>&1
tail -3 gpurun_out/${TAG}_train.json
timeout 600 python tools/bench_hgt.py --check 1 > gpurun_out/${TAG}_hgt.json 2>&1
tail -2 gpurun_out/${TAG}_hgt.json
timeout 300 python tools/bench_hgt.py --hidden 200 > gpurun_out/${TAG}_hgt200.json 2>&1
tail -1 gpurun_out/${TAG}_hgt200.json
timeout 600 python tools/bench_node_sharded.py --check > gpurun_out/${TAG}_c4_1gpu.json 2>&1
tail -1 gpurun_out/${TAG}_c4_1gpu.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_train.csv \
    python tools/bench_train.py --batch 4 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_train.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 2 -f -o gpurun_out/${TAG}_attn_fwd \
    python tools/prof_attn.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:typed_linear_tc -s 3 -c 2 -f -o gpurun_out/${TAG}_typed_linear_tc \
    python tools/prof_gemm.py > gpurun_out/${TAG}_ncu_gemm.log 2>&1
ls -la gpurun_out
