#!/bin/bash
# tensor-pipe utilisation / DRAM bytes of the tcgen05 GEMM and weight-gradient kernels at training batch scale (profiles/r2_batch_scale_tensor_pipe.txt)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed \
   --clock-control none -k regex:"typed_wgrad|typed_linear_tc|wgrad_reduce" --csv --log-file gpurun_out/prof_gemm_wgrad_metrics.csv python tools/prof_wgrad.py > gpurun_out/prof_gemm_ncu.log 2>&1
tail -2 gpurun_out/prof_gemm_ncu.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/prof_gemm_wgrad_metrics.csv')))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == 'ID': hdr = r; continue
    if not hdr or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    key = (d['ID'], d['Kernel Name'][:60])
    agg.setdefault(key, {})[d['Metric Name']] = d['Metric Value'] + ' ' + d['Metric Unit']
for k, v in agg.items():
    print(k[0], k[1], '|', v.get('gpu__time_duration.sum'), '| tensor', v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'), '| dram rd', v.get('dram__bytes_read.sum'), 'wr', v.get('dram__bytes_write.sum'))
PY
