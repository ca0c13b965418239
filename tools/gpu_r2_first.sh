#!/bin/bash
# round 2, first GPU call: C2 golden parity, full GPU suite, bench (both arms), compute-sanitizer on a subset
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_c2_golden.py -q -s > gpurun_out/r02_c2_golden.log 2>&1; echo "c2 golden rc=$? $(tail -1 gpurun_out/r02_c2_golden.log)"
grep -E "free-running|FAILED|Error|assert" gpurun_out/r02_c2_golden.log | head -20
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_c2_golden.py > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_pytest_gpu.log)"
grep -E "FAILED|Error" gpurun_out/r02_pytest_gpu.log | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/r02_calls_v0.jsonl > gpurun_out/r02_bench_v0.json 2> gpurun_out/r02_bench_v0.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_v0.json'));print(d['value'],d['ms_per_step'],d['timing'],'e2e',d['e2e']['value'],d['roofline']['frac'],d.get('cpu_baseline',{}).get('sample'))" || tail -5 gpurun_out/r02_bench_v0.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_v0.json 2> gpurun_out/r02_bench_ref_v0.err; echo "ref rc=$?"; head -c 600 gpurun_out/r02_bench_ref_v0.json
timeout 300 python bench.py --workload c4 --steps 10 > gpurun_out/r02_bench_c4_v0.json 2> gpurun_out/r02_bench_c4_v0.err; echo "c4 rc=$?"; head -c 1500 gpurun_out/r02_bench_c4_v0.json
# sanitizer: memcheck over the kernel suites (small inputs)
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_$tool.log \
    python -m pytest tests/test_gpu_rulebook.py tests/test_gpu_coords.py "tests/test_gpu_conv.py::test_sparse_conv_umma_matches_oracle" "tests/test_gpu_conv.py::test_sparse_conv_matches_oracle" tests/test_gpu_detect.py -q -x > gpurun_out/r02_sanitizer_${tool}_pytest.log 2>&1
  echo "sanitizer $tool rc=$? $(tail -1 gpurun_out/r02_sanitizer_${tool}_pytest.log) | $(grep -c 'ERROR SUMMARY' gpurun_out/r02_sanitizer_$tool.log) $(grep 'ERROR SUMMARY' gpurun_out/r02_sanitizer_$tool.log | tail -1)"
done
