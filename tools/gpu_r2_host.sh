#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_c2_golden.py -q --tb=short -x > gpurun_out/r02_c2_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_c2_pytest.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_c2_pytest.log | head
timeout 300 python tools/host_profile.py > gpurun_out/r02_host_profile.txt 2>&1; head -60 gpurun_out/r02_host_profile.txt
