#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_prune.py -q > gpurun_out/r4p_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4p_pytest_multi.log
tail -3 gpurun_out/r4p_pytest_multi.log
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r4p_bench$N.json 2> gpurun_out/r4p_bench$N.err; echo "bench$N rc=$?"
tail -2 gpurun_out/r4p_bench$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4p_bench$N.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity']['assignment_mismatches'], d['parity']['ranks'])
print('e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'))
for k in ('config3','unstructured','strong_scaling'):
    v=d.get(k,{})
    print(k, {x:v.get(x) for x in ('ms_per_step','value','rechecked_last_step','error')}, (v.get('parity') or {}).get('assignment_mismatches'))
for k in ('trajectory_config2','trajectory_config3','trajectory_unstructured'):
    v=d.get(k,{})
    print(k, v.get('speedup'), v.get('error'), v.get('default',{}).get('total_ms'), v.get('bounded_incremental',{}).get('total_ms'), v.get('iterations_with_identical_assignments'), v.get('identical_on_all_ranks'))
PY
