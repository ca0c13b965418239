#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m pytest tests/test_refine.py tests/test_gpu_engine.py -q -x -m gpu > gpurun_out/r02_pytest_refine.log 2>&1; echo "refine/engine tests rc=$? $(tail -1 gpurun_out/r02_pytest_refine.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_pytest_refine.log | head -10
for n in 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 --min-timed-s 1.5 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; echo "n=$n rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r02_bench_n$n.json'));print('n=$n value',d['value'],d['ms_per_step'],d['timing'],'e2e',d['e2e'])" || tail -20 gpurun_out/r02_bench_n$n.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-s 1.5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; python -c "import json;d=json.load(open('gpurun_out/r02_bench_n1.json'));print('n=1 value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])"
