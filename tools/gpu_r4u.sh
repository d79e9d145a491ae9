#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_prefix16|k_assign_cols" -s 4 -c 2 -o gpurun_out/r4u_full python tools/probe_prune.py 2000000 > gpurun_out/r4u_ncu.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py full gpurun_out/r4u_full.ncu-rep
ncu -i gpurun_out/r4u_full.ncu-rep --page raw --csv > gpurun_out/r4u_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r4u_raw.csv')))
hdr,units=rows[0],rows[1]
for vals in rows[2:]:
    print(vals[hdr.index("Kernel Name")][:60])
    st=[(float(vals[i].replace(',','')),h) for i,h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('.ratio') and vals[i]]
    for v,h in sorted(st,reverse=True)[:6]: print(f"   {v:8.2f} {h}")
PY
