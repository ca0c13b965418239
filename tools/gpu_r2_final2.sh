#!/bin/bash
# closing evidence pass of round 2 on ONE box (the CPU reference arm and the c4 sweep were taken by tools/gpu_r2_final.sh earlier in the round)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_final_pytest_gpu.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_final_pytest_gpu.log | head -10
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 600 ncu --metrics $M --clock-control none --nvtx --nvtx-include "fwd/" --csv --log-file gpurun_out/r02_final_fwd_metrics.csv python tools/one_forward.py > gpurun_out/r02_final_ncu_metrics.log 2>&1; echo "ncu metrics rc=$? $(wc -l < gpurun_out/r02_final_fwd_metrics.csv) lines"
python tools/ncu_fwd_summary.py gpurun_out/r02_final_fwd_metrics.csv gpurun_out/r02_final_fwd > /dev/null 2>&1 && cp gpurun_out/r02_final_fwd_traffic.json profiles/traffic.json
timeout 600 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/r02_final_calls.jsonl > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_final_bench.json'));print(d['value'],d['ms_per_step'],d['timing']['ms_per_step_min'],'e2e',d['e2e']['value'],d['roofline']['frac'],d['roofline']['traffic'],d['cpu_baseline']['value'],d['gpu_launches'])" || tail -5 gpurun_out/r02_final_bench.err
timeout 300 python bench.py --workload train --steps 12 --warmup 12 --dump-launches gpurun_out/r02_final_train_calls.jsonl > gpurun_out/r02_final_bench_train.json 2> gpurun_out/r02_final_bench_train.err; echo "train rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_final_bench_train.json'));print(d['value'],d['ms_per_step'],d['kernels']['insmos_sparse_conv_wgrad'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_final_bench_launches.csv python bench.py --steps 2 --warmup 1 --streams 0 --no-cpu-baseline --min-timed-s 0 > gpurun_out/r02_final_bench_under_ncu.log 2>&1; echo "ncu launch list rc=$? $(wc -l < gpurun_out/r02_final_bench_launches.csv) lines"
