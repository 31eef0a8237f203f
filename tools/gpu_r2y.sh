#!/bin/bash
mkdir -p gpurun_out
python tools/bench_train.py --batch 16 --steps 5 --warmup 2 > gpurun_out/r2y_train16.json 2>gpurun_out/r2y_train16.err
cat gpurun_out/r2y_train16.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2y_launches_train.csv \
   python tools/bench_train.py --batch 16 --steps 2 --warmup 2 --cuda-profile > gpurun_out/r2y_ncu.log 2>&1
tail -2 gpurun_out/r2y_ncu.log
python tools/launch_summary.py gpurun_out/r2y_launches_train.csv gpurun_out/r2y_launches_train.txt "python tools/bench_train.py --batch 16 --steps 2 --warmup 2 --cuda-profile (steady state: 2 timed steps only)" | head -70
