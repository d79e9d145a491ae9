#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -q -k "hadamard or fwht or mix" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -4 gpurun_out/r2y_pytest.log
SKM_FWHT_X=0 timeout 300 python tools/probe_fwht.py > gpurun_out/r2y_fwht_shfl.txt 2> gpurun_out/r2y_fwht_shfl.err; tail -2 gpurun_out/r2y_fwht_shfl.err; cat gpurun_out/r2y_fwht_shfl.txt
SKM_FWHT_X=1 timeout 300 python tools/probe_fwht.py > gpurun_out/r2y_fwht_x.txt 2> gpurun_out/r2y_fwht_x.err; tail -2 gpurun_out/r2y_fwht_x.err; cat gpurun_out/r2y_fwht_x.txt
