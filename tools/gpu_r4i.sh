#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prune.py -q -x 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q -x -k "bounded or incremental or modes" 2>&1 | tail -2
timeout 300 python tools/probe_list.py
timeout 600 python tools/probe_traj.py config3 100 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['speedup'], d['default_total'], d['modes_total'], d['identical']); print(d['default_ms'][:40]); print(d['modes_ms'][:40])"
