#!/bin/bash
# round-end style validation: build check, smoke, the GPU test suite, the default bench line (with cpu_baseline)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_default.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline'],d['cpu_baseline']['value'],d['clocks'],d['gpu_launches'])"; tail -3 gpurun_out/bench_default.err
