#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tcsparse.py -q -x -k "bound and (64-1024 or 40-200 or 2-64)" > gpurun_out/r3b_memcheck_tcs.log 2>&1; echo "memcheck tcs rc=$?"
tail -4 gpurun_out/r3b_memcheck_tcs.log
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipelined_create.py tests/test_gpu_parity.py -q -x -k "pipelined or kpp" > gpurun_out/r3b_memcheck_pipe.log 2>&1; echo "memcheck pipe/kpp rc=$?"
tail -4 gpurun_out/r3b_memcheck_pipe.log
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tcsparse.py -q -x -k "bound and (40-200 or 2-64)" > gpurun_out/r3b_racecheck_tcs.log 2>&1; echo "racecheck tcs rc=$?"
tail -6 gpurun_out/r3b_racecheck_tcs.log
grep -c "ERROR SUMMARY" gpurun_out/r3b_*.log; grep "ERROR SUMMARY" gpurun_out/r3b_*.log | tail -5
