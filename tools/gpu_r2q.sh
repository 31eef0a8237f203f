#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_node_sharded.py tests/test_gpu_multi.py tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2q_pytest.log
tail -12 gpurun_out/r2q_pytest.log
for wire in fp32 auto; do
for n in 1 2 4; do
  if [ $n = 1 ]; then
    timeout 600 python tools/bench_node_sharded.py --check --kv-wire $wire > gpurun_out/r2q_c4_${n}gpu_$wire.json 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/bench_node_sharded.py --check --kv-wire $wire > gpurun_out/r2q_c4_${n}gpu_$wire.json 2>&1
  fi
  tail -1 gpurun_out/r2q_c4_${n}gpu_$wire.json | cut -c1-900
done
done
