#!/bin/bash
# A/B of environment switches inside ONE box: each argument is "VAR=value[,VAR2=value2]" for one arm; arm A = defaults
mkdir -p gpurun_out
run() { env $(echo "$2" | tr ',' ' ') timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --dump-launches gpurun_out/dump_$1.jsonl > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "import json;d=json.load(open('gpurun_out/bench_$1.json'));k=d['kernels'];print('$1 [$2]',d['value'],d['ms_per_step'],'tc',k.get('insmos_sparse_conv_fwd_tc',{}).get('ms_per_step'),'umma',k.get('insmos_sparse_conv_fwd_umma',{}).get('ms_per_step'),'rb',sum(v['ms_per_step'] for n,v in k.items() if 'rulebook' in n))" || tail -3 gpurun_out/bench_$1.err; }
run A0 "X=0"
i=1
for arm in "$@"; do run B$i "$arm"; i=$((i+1)); done
run A1 "X=0"
