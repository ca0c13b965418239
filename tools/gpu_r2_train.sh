#!/bin/bash
# training-step tests (N3) on the GPU box
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -q --tb=short -s > gpurun_out/r02_train_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_train_pytest.log)"
grep -E "^E  |FAILED|Error|worst" gpurun_out/r02_train_pytest.log | head -60
