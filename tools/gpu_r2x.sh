#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --warp-sampling-interval 2 --clock-control none --import-source on \
    -k regex:typed_linear_tc -s 1 -c 1 -f -o gpurun_out/r2x_gemm python tools/prof_gemm.py fp16 > gpurun_out/r2x_ncu.log 2>&1
tail -3 gpurun_out/r2x_ncu.log
