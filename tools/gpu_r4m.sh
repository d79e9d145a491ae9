#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r4m_full.log 2>&1; echo "full rc=$?"; tail -8 gpurun_out/r4m_full.log
timeout 600 python -m pytest tests/test_gpu_bounded.py tests/test_gpu_dct_datafile.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -q -x > gpurun_out/r4m_four.log 2>&1; echo "four rc=$?"; tail -5 gpurun_out/r4m_four.log
