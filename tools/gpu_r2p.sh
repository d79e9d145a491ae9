#!/bin/bash
mkdir -p gpurun_out
for K in 64 128 32; do
timeout 300 python tools/probe_tc.py config3 2000000 mixture $K > gpurun_out/r2p_probe_K$K.json 2> gpurun_out/r2p_probe_K$K.err; echo "probe K=$K rc=$?"
tail -2 gpurun_out/r2p_probe_K$K.err; cat gpurun_out/r2p_probe_K$K.json
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_tcs_filter -c 6 --csv --log-file gpurun_out/r2p_k128.csv python tools/probe_tc.py config3 2000000 mixture 128 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2p_k128.csv | cut -d, -f5,12- | head -20
