#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2v_clocks.csv &
SMI=$!
timeout 1200 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
kill $SMI
tail -5 gpurun_out/r2v_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d.get('parity'))
print('e2e', d['e2e']['value'], d['e2e'].get('whole_job_variant'))
print('cpu', d['cpu_baseline'])
for k in ('config3','unstructured','strong_scaling'):
    v=d.get(k,{})
    print(k, {x:v.get(x) for x in ('ms_per_step','value','parity','rechecked_last_step','error','k1_ms_per_pass')}, v.get('roofline',{}).get('frac'))
for k in ('trajectory_config2','trajectory_config3','trajectory_unstructured'):
    v=d.get(k,{})
    print(k, v.get('speedup'), v.get('error'), v.get('default',{}).get('iterations'), v.get('default',{}).get('total_ms'), v.get('bounded_incremental',{}).get('total_ms'), v.get('iterations_with_identical_assignments'))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_ref.json 2> gpurun_out/r2v_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2v_ref.json
