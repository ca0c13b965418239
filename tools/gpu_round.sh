#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_full.log)"; grep -E "Error|assert|FAILED" gpurun_out/pytest_full.log | head -12
