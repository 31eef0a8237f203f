#!/bin/bash
# opcode mix / stall samples of the two forward attention kernels: register path on the L2-resident config-2 graph,
# TMA ring on the DRAM-resident config-4 graph
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --clock-control none --import-source on \
   -k regex:"attn_fwd_vec" -s 2 -c 1 -f -o gpurun_out/prof_attn_vec python tools/prof_attn.py --nodes 8192 --k 5 > gpurun_out/prof_attn_ncu_vec.log 2>&1
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --clock-control none --import-source on \
   -k regex:"attn_fwd_tma" -s 2 -c 1 -f -o gpurun_out/prof_attn_ring python tools/prof_attn.py > gpurun_out/prof_attn_ncu_ring.log 2>&1
tail -1 gpurun_out/prof_attn_ncu_vec.log; tail -1 gpurun_out/prof_attn_ncu_ring.log
