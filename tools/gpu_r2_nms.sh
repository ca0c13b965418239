#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_model.py -q -x > gpurun_out/r02_pytest_nms.log 2>&1; echo "detect/model tests rc=$? $(tail -1 gpurun_out/r02_pytest_nms.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_pytest_nms.log | head -20
timeout 600 python -m pytest tests/test_gpu_c2_golden.py -q -x > gpurun_out/r02_c2_golden_nms.log 2>&1; echo "c2 golden rc=$? $(tail -1 gpurun_out/r02_c2_golden_nms.log)"
for d in 1 0; do
  INSMOS_NMS_DENSE=$d timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 --dump-launches gpurun_out/r02_calls_nms$d.jsonl > gpurun_out/r02_bench_nms$d.json 2> gpurun_out/r02_bench_nms$d.err; echo "dense=$d rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r02_bench_nms$d.json'));print('dense $d value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'nms',d['kernels'].get('insmos_nms_rotated'),d['kernels'].get('insmos_nms_rotated_pairs'))" || tail -5 gpurun_out/r02_bench_nms$d.err
done
