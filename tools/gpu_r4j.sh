#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prune.py -q -x 2>&1 | tail -2
SKM_REEVAL=list timeout 900 python -m pytest tests/test_gpu_prune.py -q -x -k list_pass 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q -x -k "bounded or incremental or modes" 2>&1 | tail -2
echo cols; timeout 300 python tools/probe_list.py
echo list; SKM_REEVAL=list timeout 300 python tools/probe_list.py
timeout 600 python tools/probe_traj.py config3 100 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['speedup'], d['default_total'], d['modes_total'], d['identical']); print(d['default_ms'][:40]); print(d['modes_ms'][:40])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum,sm__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"k_assign_cols" -c 12 --csv --log-file gpurun_out/r4j_cols_ncu.csv python tools/probe_list.py 12500000 14 > gpurun_out/r4j_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r4j_cols_ncu.csv', errors='replace')))
st=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr=rows[st]; ki,mi,vi=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value")
cur={}
for r in rows[st+1:]:
    if len(r)<=vi: continue
    cur.setdefault((r[0], r[ki].split('(')[0][-24:]),{})[r[mi].split('.')[0][-14:]]=r[vi]
for k,v in cur.items(): print(k, v)
PY
