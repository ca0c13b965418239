#!/bin/bash
# ncu evidence of ONE forward (NVTX range 'fwd' of tools/one_forward.py); only CSV read-outs leave the box (<64 MiB)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 900 ncu --metrics $M --clock-control none --nvtx --nvtx-include "fwd/" --csv --log-file gpurun_out/fwd_metrics.csv python tools/one_forward.py > gpurun_out/ncu_metrics.log 2>&1; echo "metrics rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "fwd/" -k regex:"k_spconv_tc4|k_spconv_umma_ts" -s 8 -c 6 -o /tmp/prof_final python tools/one_forward.py > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/final_raw.csv 2>/dev/null
for i in 0 1 2 3 4 5; do ncu -i /tmp/prof_final.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > gpurun_out/final_src_$i.csv 2>/dev/null; done
ls -la gpurun_out /tmp/prof_final.ncu-rep | tail -14
