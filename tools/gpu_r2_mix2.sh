#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_eval_mode.py tests/test_gpu_deadrows.py -q --tb=short > gpurun_out/r02_mix_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_mix_pytest.log)"; grep -E "^E  |FAILED|Error" gpurun_out/r02_mix_pytest.log | head -20
