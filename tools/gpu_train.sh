#!/bin/bash
# GPU visit for the training path + HGT config-3 bench + builder: tests first.
TAG=${1:-tr}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/bench_train.py --batch 16 --steps 3 > gpurun_out/${TAG}_train.json 2>&1
tail -2 gpurun_out/${TAG}_train.json
timeout 600 python tools/bench_hgt.py --check 1 > gpurun_out/${TAG}_hgt.json 2>&1
tail -2 gpurun_out/${TAG}_hgt.json
timeout 300 python tools/bench_hgt.py --hidden 200 > gpurun_out/${TAG}_hgt200.json 2>&1
tail -1 gpurun_out/${TAG}_hgt200.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches_train.csv \
    python tools/bench_train.py --batch 4 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_train.log 2>&1
