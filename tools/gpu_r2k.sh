#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2k_launches_train.csv \
    python tools/bench_train.py --batch 8 --steps 1 --warmup 1 > gpurun_out/r2k_ncu_train.log 2>&1
tail -2 gpurun_out/r2k_ncu_train.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 2 -f -o gpurun_out/r2k_attn_hbm \
    python tools/prof_attn.py > gpurun_out/r2k_ncu_attn.log 2>&1
tail -2 gpurun_out/r2k_ncu_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:typed_linear_tc -c 6 -f -o gpurun_out/r2k_gemm \
    python tools/prof_gemm.py fp16 > gpurun_out/r2k_ncu_gemm.log 2>&1
tail -2 gpurun_out/r2k_ncu_gemm.log
ls -la gpurun_out | tail -5
