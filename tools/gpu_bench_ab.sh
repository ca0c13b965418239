#!/bin/bash
# A/B of an environment switch inside ONE box (boxes differ by a few percent): $1 = "VAR=value" of the B arm
mkdir -p gpurun_out
for rep in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 4 --no-cpu-baseline > gpurun_out/bench_a$rep.json 2> gpurun_out/bench_a.err; python -c "import json;d=json.load(open('gpurun_out/bench_a$rep.json'));print('A',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])"
  env $1 timeout 300 python bench.py --steps 30 --warmup 4 --no-cpu-baseline > gpurun_out/bench_b$rep.json 2> gpurun_out/bench_b.err; python -c "import json;d=json.load(open('gpurun_out/bench_b$rep.json'));print('B',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])"
done
