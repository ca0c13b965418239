#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv
timeout 300 python -m pytest tests/test_staging.py -m gpu -q 2>&1 | tail -3
for rep in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench$rep.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench$rep.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['clocks'])"; tail -3 gpurun_out/bench.err
done
