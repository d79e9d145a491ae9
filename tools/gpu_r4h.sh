#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_list.py
SKM_LIST_MIN=2000000000 timeout 300 python tools/probe_list.py
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum,sm__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_assign_list|k_exact_assign" -c 30 --csv --log-file gpurun_out/r4h_list_ncu.csv python tools/probe_list.py 12500000 12 > gpurun_out/r4h_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r4h_list_ncu.csv', errors='replace')))
st=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr=rows[st]; ki,mi,vi=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value")
gi=hdr.index("Grid Size") if "Grid Size" in hdr else None
cur={}
for r in rows[st+1:]:
    if len(r)<=vi: continue
    cur.setdefault((r[0], r[ki].split('(')[0][-32:], r[gi] if gi is not None else ''),{})[r[mi].split('.')[0][-14:]]=r[vi]
for k,v in cur.items(): print(k, v)
PY
