#!/bin/bash
mkdir -p gpurun_out
INSMOS_FFMA2=1 timeout 200 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "matches_oracle or tile_sizes_and_epilogue or strided" > gpurun_out/pytest_ff2.log 2>&1; echo "pytest(ffma2) rc=$? $(tail -1 gpurun_out/pytest_ff2.log)"; grep -E "^E  |FAILED" gpurun_out/pytest_ff2.log | head -6
run() { env $(echo "$2" | tr ',' ' ') timeout 200 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --dump-launches gpurun_out/dump_$1.jsonl > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "import json;d=json.load(open('gpurun_out/bench_$1.json'));k=d['kernels'];print('$1 [$2]',d['value'],d['ms_per_step'],'sum',round(sum(v['ms_per_step'] for v in k.values()),3),'tc',k.get('insmos_sparse_conv_fwd_tc',{}).get('ms_per_step'),'ffma',k.get('insmos_sparse_conv_fwd_ffma',{}).get('ms_per_step'))" || tail -3 gpurun_out/bench_$1.err; }
run A0 "X=0"
run B1 "INSMOS_FFMA2=1"
