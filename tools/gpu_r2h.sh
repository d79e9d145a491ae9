#!/bin/bash
# profiles for round 2: launch lists (gpu__time_duration) and --set full captures of the new kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (a) where the dataset build spends its device time
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2h_upload_launches.csv -c 400 python tools/probe_upload.py > gpurun_out/r2h_upload.log 2>&1
# (b) launch list of the headline bench command (2 timed iterations)
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2h_bench_launches.csv -c 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --e2e-steps 0 > gpurun_out/r2h_bench.log 2>&1
# (c) --set full of the tensor-core kernels and the verify kernel (dense second pass + DCT), one launch each
timeout 900 $NCU --set full --import-source on -k regex:"k_tc_scores|k_dense_verify|k_tc_dct|k_cast_split" -c 6 -o gpurun_out/r2h_tc python tools/bench_stages.py --only-dense --no-cpu > gpurun_out/r2h_tc.log 2>&1
# (d) --set full of the bounded pass at config 3 (fixed point)
timeout 900 $NCU --set full --import-source on -k regex:"k_assign_bounded" -s 4 -c 1 -o gpurun_out/r2h_bounded python tools/probe_bounded.py config3 > gpurun_out/r2h_bounded.log 2>&1
for f in r2h_tc r2h_bounded; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_tcgen05_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,launch__registers_per_thread,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio > gpurun_out/${f}_metrics.csv 2>&1
done
ls -la gpurun_out/r2h_*
du -sh gpurun_out
