#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prune.py -q > gpurun_out/r3i_pytest.log 2>&1; tail -2 gpurun_out/r3i_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_assign_fast|k_assign_bounded|k_exact_assign" -s 20 -c 12 --csv --log-file gpurun_out/r3i_prune_ncu.csv python tools/probe_prune.py 2000000 > gpurun_out/r3i_ncu.log 2>&1
tail -2 gpurun_out/r3i_ncu.log
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r3i_prune_ncu.csv', errors='replace')))
st=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr=rows[st]; ki,mi,vi=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value")
cur={}
for r in rows[st+1:]:
    if len(r)<=vi: continue
    key=(r[0], r[ki].split('(')[0][-40:])
    cur.setdefault(key,{})[r[mi]]=r[vi]
for k,v in cur.items(): print(k, v)
PY
