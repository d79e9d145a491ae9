#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dct_datafile.py tests/test_gpu_second_pass.py -x -q 2>&1 | tail -30 | tee gpurun_out/r2g_pytest.log
timeout 600 python tools/bench_stages.py --only-dense --no-cpu 2>&1 | tee gpurun_out/r2g_stages_dense.txt
SKM_NO_TC=1 timeout 600 python tools/bench_stages.py --only-dense --no-cpu 2>&1 | grep "second pass" | tee gpurun_out/r2g_stages_dense_notc.txt
