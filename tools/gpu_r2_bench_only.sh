#!/bin/bash
mkdir -p gpurun_out
cp gpurun_out_traffic.json profiles/traffic.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/r02_final_calls.jsonl > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_final_bench.json'));print(d['value'],d['ms_per_step'],d['timing'],'e2e',d['e2e']['value'],d['roofline']['frac'],d['roofline']['traffic'],d['cpu_baseline']['value'],d['gpu_launches'])" || tail -5 gpurun_out/r02_final_bench.err
