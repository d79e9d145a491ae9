#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prune.py tests/test_gpu_parity.py tests/test_gpu_multi.py -q -x > gpurun_out/r4a_pytest.log 2>&1; tail -5 gpurun_out/r4a_pytest.log
timeout 300 python tools/probe_prune.py
SKM_PRUNE_F32=1 timeout 300 python tools/probe_prune.py
for PAIRS in 3 5 6; do SKM_PRUNE_PAIRS=$PAIRS timeout 300 python tools/probe_prune.py; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_prefix16|k_assign_bounded|k_exact_assign|k_build_table" -s 12 -c 10 --csv --log-file gpurun_out/r4a_prune_ncu.csv python tools/probe_prune.py 2000000 > gpurun_out/r4a_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r4a_prune_ncu.csv', errors='replace')))
st=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr=rows[st]; ki,mi,vi=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value")
cur={}
for r in rows[st+1:]:
    if len(r)<=vi: continue
    cur.setdefault((r[0], r[ki].split('(')[0][-40:]),{})[r[mi].split('.')[0][-12:]+'.'+r[mi].split('.')[-1][-8:]]=r[vi]
for k,v in cur.items(): print(k, v)
PY
