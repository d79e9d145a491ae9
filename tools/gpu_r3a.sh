#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcsparse.py -q -x > gpurun_out/r3a_pytest.log 2>&1; tail -2 gpurun_out/r3a_pytest.log
timeout 300 python tools/probe_tc.py config3 2000000 > gpurun_out/r3a_base.json 2> gpurun_out/r3a_base.err; python -c "
import json; d=json.load(open('gpurun_out/r3a_base.json')); print('base tc assign_ms', d['tc']['assign_ms'], 'gather', d['gather']['assign_ms'])"
SKM_TC_HAMMER=1 timeout 300 python tools/probe_tc.py config3 2000000 > gpurun_out/r3a_hammer.json 2> gpurun_out/r3a_hammer.err; python -c "
import json; d=json.load(open('gpurun_out/r3a_hammer.json')); print('hammer tc assign_ms', d['tc']['assign_ms'], 'same', d['tc']['same_as_gather'])"
grep "tc hammer" gpurun_out/r3a_hammer.err | tail -3
