#!/bin/bash
mkdir -p gpurun_out
for n in 12288 16384 24576 32768 49152 65536; do
  for k in 5 8; do
    echo "nodes $n k $k" >> gpurun_out/r2h_attn_cross.jsonl
    timeout 600 python tools/sweep_dev.py --attn-only --k $k --nodes $n 2>&1 | grep hetero >> gpurun_out/r2h_attn_cross.jsonl
  done
done
cat gpurun_out/r2h_attn_cross.jsonl
