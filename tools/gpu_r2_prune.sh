#!/bin/bash
# dead-row elimination: tests + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deadrows.py tests/test_gpu_model.py tests/test_gpu_c2_golden.py -q --tb=short > gpurun_out/r02_prune_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_prune_pytest.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_prune_pytest.log | head -20
for v in 1 0; do
INSMOS_TPRUNE=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --dump-launches gpurun_out/r02_prune${v}_calls.jsonl > gpurun_out/r02_prune${v}_bench.json 2> gpurun_out/r02_prune${v}_bench.err; echo "bench prune=$v rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_prune${v}_bench.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['roofline']['frac']);[print(k,v) for k,v in list(d['kernels'].items())[:8]]" || tail -5 gpurun_out/r02_prune${v}_bench.err
done
