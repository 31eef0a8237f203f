#!/bin/bash
# 16-slide train step, then its STEADY-STATE launch list (profiles/r2_launches_train.txt)
mkdir -p gpurun_out
python tools/bench_train.py --batch 16 --steps 5 --warmup 2 > gpurun_out/prof_train_train16.json 2>gpurun_out/prof_train_train16.err
cat gpurun_out/prof_train_train16.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/prof_train_launches_train.csv \
   python tools/bench_train.py --batch 16 --steps 2 --warmup 2 --cuda-profile > gpurun_out/prof_train_ncu.log 2>&1
tail -2 gpurun_out/prof_train_ncu.log
python tools/launch_summary.py gpurun_out/prof_train_launches_train.csv gpurun_out/prof_train_launches_train.txt "python tools/bench_train.py --batch 16 --steps 2 --warmup 2 --cuda-profile (steady state: 2 timed steps only)" | head -70
