#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_multi.py -q -k "devices_option" 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_dct_datafile.py tests/test_gpu_multi.py -q -x -k "v73 or devices_option" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_bounded.py tests/test_gpu_dct_datafile.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -q -x 2>&1 | tail -1
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_multi.py -q -x -k "devices_option" > gpurun_out/r4l_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -c "Uninitialized" gpurun_out/r4l_initcheck.log; grep -A12 "Uninitialized" gpurun_out/r4l_initcheck.log | head -60
