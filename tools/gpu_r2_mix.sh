#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_eval_mode.py tests/test_gpu_deadrows.py tests/test_gpu_conv.py -q --tb=short > gpurun_out/r02_mix_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_mix_pytest.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_mix_pytest.log | head -20
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --dump-launches gpurun_out/r02_mix_calls.jsonl > gpurun_out/r02_mix_bench.json 2> gpurun_out/r02_mix_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_mix_bench.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['roofline']['frac'],d['roofline']['ms_per_step']);[print(k,v) for k,v in list(d['kernels'].items())[:4]]" || tail -5 gpurun_out/r02_mix_bench.err
timeout 600 python bench.py --workload train --steps 12 --warmup 12 --dump-launches gpurun_out/r02_train2_calls.jsonl > gpurun_out/r02_bench_train2.json 2> gpurun_out/r02_bench_train2.err; echo "bench train rc=$?"; tail -3 gpurun_out/r02_bench_train2.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_train2.json'))
print(d['value'], d['ms_per_step'], d['step_wall_ms'])
for k,v in list(d['kernels'].items())[:6]: print(k, v)"
