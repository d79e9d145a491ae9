#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool initcheck --print-limit 400 --error-exitcode 9 python -m pytest tests/test_gpu_prune.py tests/test_gpu_multi.py -q -x -k "(list_pass and 33-256) or (matches_reference and 64-1024 and half) or devices_option or modes_enabled" > gpurun_out/r4r_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -2 gpurun_out/r4r_initcheck.log
grep -o "at .*in [a-z_0-9]*\.cu:[0-9]*\|access by cudaMemcpy source" gpurun_out/r4r_initcheck.log | sort | uniq -c | sort -rn | head
grep -A4 "access by cudaMemcpy" gpurun_out/r4r_initcheck.log | grep "libskm" | sort | uniq -c | head
