#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcsparse.py -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -15 gpurun_out/r2q_pytest.log
timeout 300 python tools/probe_tc.py config3 > gpurun_out/r2q_probe.json 2> gpurun_out/r2q_probe.err; echo "probe rc=$?"
tail -3 gpurun_out/r2q_probe.err; cat gpurun_out/r2q_probe.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tcs_filter -s 2 -c 1 -o gpurun_out/r2q_tcs python tools/probe_tc.py config3 2000000 > gpurun_out/r2q_ncu.log 2>&1
tail -2 gpurun_out/r2q_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python tools/probe_tc.py config3 2000000 > gpurun_out/r2q_l.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2q_launches.csv | head -8
