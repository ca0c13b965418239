#!/bin/bash
# quick validation: GPU suite + one bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_quick_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_quick_pytest.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_quick_pytest.log | head -10
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --dump-launches gpurun_out/r02_quick_calls.jsonl > gpurun_out/r02_quick_bench.json 2> gpurun_out/r02_quick_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_quick_bench.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value']);[print(k,v) for k,v in list(d['kernels'].items())[:16]]" || tail -5 gpurun_out/r02_quick_bench.err
