#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bev.py tests/test_gpu_c2_golden.py tests/test_gpu_model.py -q -x > gpurun_out/r02_pytest_epi.log 2>&1; echo "conv/bev/c2/model rc=$? $(tail -1 gpurun_out/r02_pytest_epi.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_pytest_epi.log | head -20
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --dump-launches gpurun_out/r02_calls_v6.jsonl > gpurun_out/r02_bench_v6.json 2> gpurun_out/r02_bench_v6.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_v6.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value']);[print(k,v) for k,v in list(d['kernels'].items())[:8]]" || tail -5 gpurun_out/r02_bench_v6.err
