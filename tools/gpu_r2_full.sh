#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_full.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_pytest_gpu_full.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_pytest_gpu_full.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --dump-launches gpurun_out/r02_calls_v4.jsonl > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r02_bench_v4.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value']);[print(k,v) for k,v in list(d['kernels'].items())[:14]]" || tail -5 gpurun_out/r02_bench_v4.err
