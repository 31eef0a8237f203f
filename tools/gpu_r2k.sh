#!/bin/bash
mkdir -p gpurun_out
python tools/prof_knn.py | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2k_knn.csv python tools/prof_knn.py > gpurun_out/r2k_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2k_knn.csv gpurun_out/r2k_launches_knn.txt "python tools/prof_knn.py (config-4 edge builder, 100k nodes, second call)" | head -30
