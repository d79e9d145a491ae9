#!/bin/bash
# round 2, call A: regression tests, new bench legs (tiny then full), K=64 micro-benchmarks
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --config tiny --steps 5 --cpu-seconds 2 > gpurun_out/r2a_bench_tiny.json 2> gpurun_out/r2a_bench_tiny.err; echo "tiny rc=$?"
tail -5 gpurun_out/r2a_bench_tiny.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2a_bench.err
cd tools/microbench
timeout 300 ./k64_probe 12500000 0 > ../../gpurun_out/r2a_probe_random.txt 2>&1
timeout 300 ./k64_probe 12500000 1 > ../../gpurun_out/r2a_probe_cfree.txt 2>&1
cat ../../gpurun_out/r2a_probe_random.txt ../../gpurun_out/r2a_probe_cfree.txt
timeout 600 ncu --set full --clock-control none -k regex:probe_ -o ../../gpurun_out/r2a_probe ./k64_probe 2000000 1 > ../../gpurun_out/r2a_probe_ncu.log 2>&1
cd ../..
ls -la gpurun_out | tail -12
