#!/bin/bash
# ncu --set full of the attention backward kernels inside a 16-slide train step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_bwd_kernel|attn_bwd_src" -c 2 -f -o gpurun_out/prof_attn_bwd_attn_bwd \
   python tools/bench_train.py --batch 16 --steps 1 --warmup 2 --cuda-profile > gpurun_out/prof_attn_bwd_ncu.log 2>&1
tail -2 gpurun_out/prof_attn_bwd_ncu.log
ncu -i gpurun_out/prof_attn_bwd_attn_bwd.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/prof_attn_bwd_raw.csv
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/prof_attn_bwd_raw.csv')))
hdr = rows[0]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__average_warp_latency_issue_stalled_long_scoreboard.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__occupancy_limit_registers','sm__maximum_warps_per_active_cycle_pct','smsp__inst_executed.sum','lts__t_bytes.sum','l1tex__t_bytes.sum']
idx = {h:i for i,h in enumerate(hdr)}
for r in rows[2:]:
    print('----')
    for w in want:
        if w in idx: print(w, '=', r[idx[w]])
PY
