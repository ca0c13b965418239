#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(gpu) rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_dump.jsonl > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
head -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
