#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
INSMOS_HOLD_GIL=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 > gpurun_out/r02_gil${v}_bench.json 2> gpurun_out/r02_gil${v}_bench.err; echo "bench hold_gil=$v rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_gil${v}_bench.json'));print(d['value'],d['ms_per_step'],d['timing']['ms_per_step_min'],'e2e',d['e2e']['value'])" || tail -5 gpurun_out/r02_gil${v}_bench.err
done
INSMOS_HOLD_GIL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --streams 3 > gpurun_out/r02_gil1_s3_bench.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_gil1_s3_bench.json'));print('streams3',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])"
INSMOS_HOLD_GIL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --streams 0 > gpurun_out/r02_gil1_s0_bench.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02_gil1_s0_bench.json'));print('streams0',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])"
