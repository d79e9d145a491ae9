#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcsparse.py -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -30 gpurun_out/r2l_pytest.log
timeout 300 python tools/probe_tc.py config3 2000000 > gpurun_out/r2l_probe_2e6.json 2> gpurun_out/r2l_probe_2e6.err; echo "probe rc=$?"
tail -3 gpurun_out/r2l_probe_2e6.err; cat gpurun_out/r2l_probe_2e6.json
timeout 300 python tools/probe_tc.py config3 > gpurun_out/r2l_probe.json 2> gpurun_out/r2l_probe.err; echo "probe rc=$?"
tail -3 gpurun_out/r2l_probe.err; cat gpurun_out/r2l_probe.json
