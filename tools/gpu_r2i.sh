#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -x -q -k "hadamard or fwht or mix or sample" 2>&1 | tail -5 | tee gpurun_out/r2i_pytest.log
echo "--- TMA prefetch (default)"; timeout 600 python tools/bench_stages.py --no-cpu 2>&1 | grep "K4" | tee gpurun_out/r2i_fwht_tma.txt
echo "--- no TMA"; SKM_FWHT_NO_TMA=1 timeout 600 python tools/bench_stages.py --no-cpu 2>&1 | grep "K4 fwht_f32_inplace" | tee gpurun_out/r2i_fwht_notma.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:"k_fwht_cta|k_fwht_sample_cta" -c 8 -o gpurun_out/r2i_fwht python tools/bench_stages.py --no-cpu --quick > gpurun_out/r2i_fwht_ncu.log 2>&1
ncu -i gpurun_out/r2i_fwht.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum > gpurun_out/r2i_fwht_metrics.csv 2>&1
cat gpurun_out/r2i_fwht_metrics.csv | cut -c1-400 | head -30
