#!/bin/bash
mkdir -p gpurun_out
SKM_PRUNE=0 timeout 600 python tools/probe_traj.py config3 40 2> gpurun_out/r3h_a.err | tee gpurun_out/r3h_noprune.json | cut -c1-900
timeout 600 python tools/probe_traj.py config3 40 2> gpurun_out/r3h_b.err | tee gpurun_out/r3h_prune.json | cut -c1-900
