#!/bin/bash
# ncu launch list (gpu__time_duration only) of ONE forward of the final build
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "fwd/" --csv --log-file gpurun_out/launches_final.csv python tools/one_forward.py > gpurun_out/ncu_launches.log 2>&1; echo "rc=$? $(wc -l < gpurun_out/launches_final.csv) lines"
