#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_second_pass.py -q -x 2>&1 | tail -3
echo "--- register bins"; timeout 300 python tools/bench_stages.py --only-dense 2>&1 | grep "k_dense_sums"
echo "--- shared-memory bins"; SKM_DENSE_SUMS_SMEM=1 timeout 300 python tools/bench_stages.py --only-dense 2>&1 | grep "k_dense_sums"
