#!/bin/bash
# forward bench at 8 GPUs (one sample per rank, fixed-size NCCL gather of logits)
mkdir -p gpurun_out
N=${1:-8}
nproc; nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['timing'])"
