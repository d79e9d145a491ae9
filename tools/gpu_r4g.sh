#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prune.py tests/test_gpu_parity.py tests/test_gpu_pipeline.py -q -x > gpurun_out/r4g_pytest.log 2>&1; tail -4 gpurun_out/r4g_pytest.log
timeout 600 python -m pytest tests -m gpu -q -x -k "bounded or incremental or modes" 2>&1 | tail -2
echo "--- list pass on"; timeout 600 python tools/probe_traj.py config3 100 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['speedup'], d['default_total'], d['modes_total'], d['identical']); print(d['default_ms'][:60]); print(d['modes_ms'][:60])"
echo "--- list pass off"; SKM_LIST_MIN=2000000000 timeout 600 python tools/probe_traj.py config3 100 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['speedup'], d['default_total'], d['modes_total'], d['identical']); print(d['default_ms'][:60]); print(d['modes_ms'][:60])"
