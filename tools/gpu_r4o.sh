#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r4o_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4o_pytest_multi.log
tail -3 gpurun_out/r4o_pytest_multi.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r4o_bench$N.json 2> gpurun_out/r4o_bench$N.err; echo "bench$N rc=$?"
tail -3 gpurun_out/r4o_bench$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4o_bench$N.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'])
print('e2e', d['e2e']['value'], d['e2e'].get('host_cpus_bound_to_gpu'), d['e2e'].get('ms_per_step'))
for k in ('config3','unstructured','strong_scaling'):
    v=d.get(k,{})
    print(k, {x:v.get(x) for x in ('ms_per_step','value','parity','rechecked_last_step','error')})
for k in ('trajectory_config2','trajectory_config3','trajectory_unstructured'):
    v=d.get(k,{})
    print(k, v.get('speedup'), v.get('error'), v.get('default',{}).get('iterations'), v.get('default',{}).get('total_ms'), v.get('bounded_incremental',{}).get('total_ms'), v.get('iterations_with_identical_assignments'), v.get('identical_on_all_ranks'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/bench_kpp.py > gpurun_out/r4o_kpp2.json 2> gpurun_out/r4o_kpp2.err; cat gpurun_out/r4o_kpp2.json
