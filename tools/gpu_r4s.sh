#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for T in test_gpu_pipeline test_gpu_second_pass test_gpu_dct_datafile test_gpu_pipelined_create; do
  timeout 500 $CS --tool initcheck --print-limit 300 --error-exitcode 9 python -m pytest tests/$T.py -q -x > gpurun_out/r4s_init_$T.log 2>&1; echo "$T rc=$?"; tail -2 gpurun_out/r4s_init_$T.log | head -1
  grep -o "at .*in [a-z_0-9]*\.cu:[0-9]*\|access by cudaMemcpy source\|access by cudaMemset" gpurun_out/r4s_init_$T.log | sort | uniq -c | sort -rn | head -6
  grep -A5 "access by cudaMemcpy" gpurun_out/r4s_init_$T.log | grep "libskm" | sort | uniq -c | head -4
done
