#!/bin/bash
mkdir -p gpurun_out
for k in 0 52 48 44 40 32; do SKM_BOUNDED_KSM=$k python tools/probe_bounded.py config3; done 2>&1 | tee gpurun_out/r2e_bounded_ksm.txt
SKM_TRACE=1 python tools/probe_upload.py 2>&1 | tee gpurun_out/r2e_upload_trace.txt
