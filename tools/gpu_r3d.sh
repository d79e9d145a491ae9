#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_multi.py > gpurun_out/r3d_multi.json 2> gpurun_out/r3d_multi.err; echo "multi rc=$?"; tail -3 gpurun_out/r3d_multi.err; cat gpurun_out/r3d_multi.json
