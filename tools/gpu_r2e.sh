#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/forward_breakdown.py --reps 1 > gpurun_out/r2e_memcheck.log 2>&1
grep -v "^\[W" gpurun_out/r2e_memcheck.log | head -60
