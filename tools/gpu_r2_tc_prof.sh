#!/bin/bash
mkdir -p gpurun_out
for L in 9 16; do
  timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "fwd/" -k regex:"k_spconv_tc4" -s $L -c 1 -o /tmp/prof_tc_$L python tools/one_forward.py > gpurun_out/r02_ncu_tc_$L.log 2>&1; echo "full $L rc=$?"
  ncu -i /tmp/prof_tc_$L.ncu-rep --page raw --csv > gpurun_out/r02_tc_${L}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_tc_$L.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_tc_${L}_src.csv 2>/dev/null
  ncu -i /tmp/prof_tc_$L.ncu-rep --page details > gpurun_out/r02_tc_${L}_details.txt 2>/dev/null
done
