#!/bin/bash
# N3: training tests + config-5 bench line on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q --tb=short > gpurun_out/r02_train_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_train_pytest.log)"
timeout 600 python bench.py --workload train --steps 12 --warmup 6 --dump-launches gpurun_out/r02_train_calls.jsonl > gpurun_out/r02_bench_train.json 2> gpurun_out/r02_bench_train.err; echo "bench train rc=$?"; tail -3 gpurun_out/r02_bench_train.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_train.json'))
print(d["step_wall_ms"]); print(d["value"], d["ms_per_step"], d['losses_first_last'], d['gpu_launches'], d['profiled_step_ms'])
for k,v in list(d['kernels'].items())[:14]: print(k, v)
P
