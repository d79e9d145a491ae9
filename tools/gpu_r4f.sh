#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_prune.py -q -x -k "half and (64-1024 or 130-300 or 33-256)" > gpurun_out/r4f_memcheck_prefix16.log 2>&1; echo "memcheck prefix16 rc=$?"
tail -3 gpurun_out/r4f_memcheck_prefix16.log
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_prune.py -q -x -k "half and mixture and (64-1024 or 130-300)" > gpurun_out/r4f_racecheck_prefix16.log 2>&1; echo "racecheck prefix16 rc=$?"
tail -3 gpurun_out/r4f_racecheck_prefix16.log
timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -q -x -k "fwht_inplace_involution or fused_fwht" > gpurun_out/r4f_memcheck_fwht.log 2>&1; echo "memcheck fwht rc=$?"
tail -3 gpurun_out/r4f_memcheck_fwht.log
timeout 600 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -q -x -k "fwht_inplace_involution" > gpurun_out/r4f_racecheck_fwht.log 2>&1; echo "racecheck fwht rc=$?"
tail -3 gpurun_out/r4f_racecheck_fwht.log
grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/r4f_*.log | tail -8
timeout 600 python -m pytest tests/test_gpu_dct_datafile.py -q -x 2>&1 | tail -2
