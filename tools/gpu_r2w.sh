#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "kpp" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
tail -8 gpurun_out/r2w_pytest.log
timeout 600 python tools/bench_kpp.py > gpurun_out/r2w_kpp1.json 2> gpurun_out/r2w_kpp1.err; echo "kpp rc=$?"; tail -2 gpurun_out/r2w_kpp1.err; cat gpurun_out/r2w_kpp1.json
SKM_NO_KPP_FILTER=1 timeout 600 python tools/bench_kpp.py > gpurun_out/r2w_kpp1_nofilter.json 2> gpurun_out/r2w_kpp1_nofilter.err; cat gpurun_out/r2w_kpp1_nofilter.json
