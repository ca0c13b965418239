#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rulebook.py -q -x > gpurun_out/r02_pytest_lg.log 2>&1; echo "rulebook tests rc=$? $(tail -1 gpurun_out/r02_pytest_lg.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_pytest_lg.log | head -20
timeout 600 python -m pytest tests/test_gpu_c2_golden.py tests/test_gpu_model.py -q -x > gpurun_out/r02_c2_golden_lg.log 2>&1; echo "c2 golden + model rc=$? $(tail -1 gpurun_out/r02_c2_golden_lg.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_c2_golden_lg.log | head
for lg in 0 1; do
  INSMOS_LEAFGRID=$lg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --dump-launches gpurun_out/r02_calls_lg$lg.jsonl > gpurun_out/r02_bench_lg$lg.json 2> gpurun_out/r02_bench_lg$lg.err; echo "lg=$lg rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r02_bench_lg$lg.json'));k=d['kernels'];print('lg $lg value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value']);[print('   ',n,k[n]) for n in k if 'rulebook' in n or 'leafgrid' in n or 'xblock' in n or 'table_clear' in n]" || tail -5 gpurun_out/r02_bench_lg$lg.err
done
timeout 300 python bench.py --workload c4 --steps 10 > gpurun_out/r02_bench_c4_lg.json 2> gpurun_out/r02_bench_c4_lg.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_c4_lg.json'))
for r in d['rows']: print(r)"
