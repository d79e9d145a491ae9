#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_second_pass.py -q -x 2>&1 | tail -1
for UC in 8 16 32; do echo "UC=$UC"; SKM_DENSE_SUMS_UC=$UC timeout 300 python tools/bench_stages.py --only-dense 2>&1 | grep "k_dense_sums" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['K'], d['kernel_ms'], round(d['frac_of_hbm_peak'],3))"; done
