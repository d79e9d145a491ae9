#!/bin/bash
mkdir -p gpurun_out
# launch list of the headline command (config 2) and of the config-3 shape, one pass each: gpu__time_duration only
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|::k_" -c 300 --csv --log-file gpurun_out/r4t_launches_config2.csv python bench.py --steps 4 --warmup 3 --no-extra --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r4t_b2.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|::k_" -c 300 --csv --log-file gpurun_out/r4t_launches_config3.csv python bench.py --config config3 --steps 4 --warmup 3 --no-extra --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r4t_b3.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py launches gpurun_out/r4t_launches_config2.csv
python tools/ncu_summary.py launches gpurun_out/r4t_launches_config3.csv
