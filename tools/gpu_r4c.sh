#!/bin/bash
mkdir -p gpurun_out
for BW in 8 16 32; do for W in 4 8 16; do echo "BW=$BW W=$W"; SKM_K2_BW=$BW SKM_K2_WARPS=$W timeout 300 python tools/probe_prune.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['accumulate_ms'], d['assign_ms'])"; done; done
timeout 600 ncu --set full --clock-control none -k regex:"k_accumulate_csr" -s 3 -c 1 -o gpurun_out/r4c_k2 python tools/probe_prune.py 2000000 > gpurun_out/r4c_ncu.log 2>&1
python tools/ncu_summary.py full gpurun_out/r4c_k2.ncu-rep
ncu -i gpurun_out/r4c_k2.ncu-rep --page raw --csv > gpurun_out/r4c_k2_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r4c_k2_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
st=[(float(vals[i].replace(',','')),h) for i,h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('.ratio') and vals[i]]
for v,h in sorted(st,reverse=True)[:8]: print(f"{v:8.2f} {h}")
PY
