#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_prune.py tests/test_gpu_parity.py -q -x -k "list_pass and (64-1024 or 100-512) or bit_reproducible" > gpurun_out/r4q_memcheck_cols.log 2>&1; echo "memcheck cols rc=$?"; tail -2 gpurun_out/r4q_memcheck_cols.log
timeout 600 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_prune.py tests/test_gpu_parity.py -q -x -k "list_pass and 33-256 or bit_reproducible" > gpurun_out/r4q_racecheck_cols.log 2>&1; echo "racecheck cols rc=$?"; tail -2 gpurun_out/r4q_racecheck_cols.log
timeout 900 $CS --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_prune.py tests/test_gpu_multi.py -q -x -k "(list_pass and 33-256) or (matches_reference and 64-1024 and half) or devices_option or modes_enabled" > gpurun_out/r4q_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -2 gpurun_out/r4q_initcheck.log
grep -c "Uninitialized" gpurun_out/r4q_initcheck.log
grep -B1 -A6 "Uninitialized __global__" gpurun_out/r4q_initcheck.log | grep -v "Host Frame" | head -50
grep "Host Frame:.*libskm\|access by cudaMemcpy" gpurun_out/r4q_initcheck.log | sort | uniq -c | sort -rn | head
