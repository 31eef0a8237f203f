#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2r_launches_hgt.csv \
    python tools/bench_hgt.py --precision bf16 --steps 1 > gpurun_out/r2r_ncu_hgt.log 2>&1
tail -1 gpurun_out/r2r_ncu_hgt.log | cut -c1-300
