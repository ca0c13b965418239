#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_rulebook.py tests/test_gpu_bev.py -m gpu -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest(conv,rulebook,bev) rc=$?" | tee -a gpurun_out/summary.txt
tail -8 gpurun_out/pytest_conv.log
timeout 200 python tools/umma_accuracy.py > gpurun_out/umma_accuracy.log 2>&1; cat gpurun_out/umma_accuracy.log
timeout 120 python tools/time_bev.py > gpurun_out/time_bev.log 2>&1; cat gpurun_out/time_bev.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_conv.py --deselect tests/test_gpu_rulebook.py --deselect tests/test_gpu_bev.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(rest) rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_dump.jsonl > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
head -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
