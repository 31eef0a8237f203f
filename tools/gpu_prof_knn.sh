#!/bin/bash
# config-4 edge builder: time and launch list (profiles/r2_launches_knn_builder.txt)
mkdir -p gpurun_out
python tools/prof_knn.py | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/prof_knn_knn.csv python tools/prof_knn.py > gpurun_out/prof_knn_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/prof_knn_knn.csv gpurun_out/prof_knn_launches_knn.txt "python tools/prof_knn.py (config-4 edge builder, 100k nodes, second call)" | head -30
