#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_rulebook.py -m gpu -x -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest(conv,rulebook) rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_conv.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_conv.py --deselect tests/test_gpu_rulebook.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest(rest) rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/launch_dump.jsonl > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
head -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spconv_umma -c 8 -o gpurun_out/prof_umma python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_umma.log 2>&1; echo "ncu umma rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spconv_tc3|k_conv_nhwc_tcgen05" -c 26 -o gpurun_out/prof_tc3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc3.log 2>&1; echo "ncu tc3 rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out
