#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" > gpurun_out/pytest_line.txt
for rep in 1 2 3; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_dump.jsonl > gpurun_out/bench$rep.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench$rep.json'));k=d['kernels'];print(d['value'],d['ms_per_step'],d['e2e']['value'],'sum',round(sum(v['ms_per_step'] for v in k.values()),3),'tc',k['insmos_sparse_conv_fwd_tc']['ms_per_step'],'rb',round(sum(v['ms_per_step'] for n,v in k.items() if 'rulebook' in n),3))"; tail -3 gpurun_out/bench.err
done
cat gpurun_out/pytest_line.txt
