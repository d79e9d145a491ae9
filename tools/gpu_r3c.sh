#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tcsparse.py -q -x -k "bound and (40-200 or 2-64)" > gpurun_out/r3c_racecheck_tcs.log 2>&1; echo "racecheck tcs rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r3c_racecheck_tcs.log | tail -3
grep -E "Race reported|and (Write|Read) access" gpurun_out/r3c_racecheck_tcs.log | sed 's/.*tcsparse.cu:/line /' | sort | uniq -c | sort -rn | head -8
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "kpp_filtered" > gpurun_out/r3c_racecheck_kpp.log 2>&1; echo "racecheck kpp rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r3c_racecheck_kpp.log | tail -3
timeout 300 python tools/probe_tc.py config3 2000000 > gpurun_out/r3c_probe.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r3c_probe.json')); print('tc assign_ms', d['tc']['assign_ms'], 'gather', d['gather']['assign_ms'], d['tc']['same_as_gather'])"
