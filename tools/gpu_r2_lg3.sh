#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rulebook.py tests/test_gpu_c2_golden.py -q -x > gpurun_out/r02_pytest_lg3.log 2>&1; echo "rulebook + c2 golden rc=$? $(tail -1 gpurun_out/r02_pytest_lg3.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_pytest_lg3.log | head -20
for lg in 0 1; do
  INSMOS_LEAFGRID=$lg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --dump-launches gpurun_out/r02_calls_lg3_$lg.jsonl > gpurun_out/r02_bench_lg3_$lg.json 2> gpurun_out/r02_bench_lg3_$lg.err; echo "lg=$lg rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02_bench_lg3_$lg.json'));k=d['kernels'];print('lg $lg value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])
rows=[json.loads(l) for l in open('gpurun_out/r02_calls_lg3_$lg.jsonl')]
print('   rulebook family ms:', round(sum(r['ms'] for r in rows if 'rulebook' in r['call'] or 'leafgrid' in r['call'] or 'xblock' in r['call']),4), [ (r['call'][16:], r['ms'], r.get('K')) for r in rows if ('rulebook' in r['call'] and r.get('K',0)>=27 and r.get('n_out',0)>50000) or 'leafgrid' in r['call']])
" || tail -5 gpurun_out/r02_bench_lg3_$lg.err
done
