#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -k "multi or kpp" > gpurun_out/r2x_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest_multi.log
tail -5 gpurun_out/r2x_pytest_multi.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --cpu-seconds 4 --traj-iters 40 > gpurun_out/r2x_bench$N.json 2> gpurun_out/r2x_bench$N.err; echo "bench$N rc=$?"
tail -3 gpurun_out/r2x_bench$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2x_bench$N.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'parity', d['parity'])
print('e2e', d['e2e']['value'])
for k in ('config3','unstructured','strong_scaling'):
    v=d.get(k,{})
    print(k, {x:v.get(x) for x in ('ms_per_step','value','parity','rechecked_last_step','error')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/bench_kpp.py > gpurun_out/r2x_kpp2.json 2> gpurun_out/r2x_kpp2.err; cat gpurun_out/r2x_kpp2.json
