#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -30 gpurun_out/r2c_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --traj-iters 40 > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r2c_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench2.json'))
for k in ('value','ms_per_step','parity','strong_scaling'):
    print(k, d.get(k))
for k in ('config3','unstructured'):
    v=d.get(k,{})
    print(k, {x:v.get(x) for x in ('ms_per_step','value','parity','rechecked_last_step','error')})
for k in ('trajectory_config2','trajectory_config3','trajectory_unstructured'):
    v=d.get(k,{})
    print(k, v.get('speedup'), v.get('error'), v.get('default',{}).get('iterations'), v.get('default',{}).get('total_ms'), v.get('bounded_incremental',{}).get('total_ms'), v.get('iterations_with_identical_assignments'), v.get('bounded_incremental',{}).get('ms_per_iteration'))
print(d['e2e'])
PY
