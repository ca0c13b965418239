#!/bin/bash
# quick validation: GPU test suite + one bench line (no CPU baseline leg)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; python -c "import json;d=json.load(open('gpurun_out/bench_last.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['roofline']['frac'])" || tail -3 gpurun_out/bench_last.err
