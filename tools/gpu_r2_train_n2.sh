#!/bin/bash
# N3 at 2 GPUs: one sample per rank, ONE all-reduce of the flat gradient buffer per step
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --workload train --gpus 2 --steps 12 --warmup 12 > gpurun_out/r02_bench_train_n2.json 2> gpurun_out/r02_bench_train_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_train_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_train_n2.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['step_wall_ms'], d['losses_first_last'][1])"
