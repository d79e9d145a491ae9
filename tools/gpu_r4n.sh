#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r4n_full.log 2>&1; echo "full rc=$?"; tail -6 gpurun_out/r4n_full.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -k "devices_option or bit_reproducible" 2>&1 | tail -1; done
timeout 600 python bench.py > gpurun_out/r4n_bench.json 2> gpurun_out/r4n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4n_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d.get('parity',{}).get('assignment_mismatches'), 'objective', d.get('objective'))
c=d['config3']; print('config3', c['ms_per_step'], c['roofline']['frac'], c['roofline'].get('traffic'), c['parity']['assignment_mismatches'], c['roofline']['step_breakdown_ms'])
PY
