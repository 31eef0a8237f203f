#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep_dev.py --attn-only > gpurun_out/r2g_attn_k5.jsonl 2>&1
cat gpurun_out/r2g_attn_k5.jsonl
timeout 600 python tools/sweep_dev.py --attn-only --k 8 --nodes 100000 > gpurun_out/r2g_attn_k8_100k.jsonl 2>&1
cat gpurun_out/r2g_attn_k8_100k.jsonl
