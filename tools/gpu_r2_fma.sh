#!/bin/bash
mkdir -p gpurun_out
(cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate) > gpurun_out/r02_ffma_rate.txt 2>&1; cat gpurun_out/r02_ffma_rate.txt
timeout 900 python -m pytest tests/test_gpu_conv.py -q -x > gpurun_out/r02_pytest_conv.log 2>&1; echo "conv tests rc=$? $(tail -1 gpurun_out/r02_pytest_conv.log)"; grep -E "^E  |FAILED" gpurun_out/r02_pytest_conv.log | head -10
timeout 900 python tools/bench_convs.py 10 > gpurun_out/r02_bench_convs_v1.jsonl 2> gpurun_out/r02_bench_convs_v1.err; echo "bench_convs rc=$?"; tail -3 gpurun_out/r02_bench_convs_v1.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r02_bench_convs_v1.jsonl')]
keys=sorted({(r['ts'],r['TM'],r['Cin'],r['Cout']) for r in rows})
for k in keys:
    print(k, ' | '.join('%s %.3f' % (r['variant'].replace('fma ','').replace(' (mma.sync 3xTF32)',''), r.get('ms',-1)) for r in rows if (r['ts'],r['TM'],r['Cin'],r['Cout'])==k))
print('max diff', max(r.get('max_diff_vs_first',0) for r in rows))
PY
timeout 600 python -m pytest tests/test_gpu_c2_golden.py -q -x > gpurun_out/r02_c2_golden_fma.log 2>&1; echo "c2 golden (fma default) rc=$? $(tail -1 gpurun_out/r02_c2_golden_fma.log)"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1 --dump-launches gpurun_out/r02_calls_v1.jsonl > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_v1.json'));print(d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['roofline']);[print(k,v) for k,v in list(d['kernels'].items())[:8]]" || tail -5 gpurun_out/r02_bench_v1.err
