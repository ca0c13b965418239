#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not fullsize" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" > gpurun_out/pytest_line.txt
run() { env $(echo "$2" | tr ',' ' ') timeout 300 python bench.py --steps 30 --warmup 4 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "import json;d=json.load(open('gpurun_out/bench_$1.json'));k=d['kernels'];print('$1 [$2]',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'sum',round(sum(v['ms_per_step'] for v in k.values()),3))" || tail -3 gpurun_out/bench_$1.err; }
run A0 "X=0"
run A1 "X=0"
cat gpurun_out/pytest_line.txt
