#!/bin/bash
# ncu --set full of two k_spconv_umma_ts launches of one forward: sparse 128->128 (7th launch) and dense BEV 128->128 (12th)
mkdir -p gpurun_out
for L in 7 12; do
  timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "fwd/" -k regex:"k_spconv_umma_ts" -s $L -c 1 -o /tmp/prof_umma_$L python tools/one_forward.py > gpurun_out/r02_ncu_umma_$L.log 2>&1; echo "full $L rc=$?"
  ncu -i /tmp/prof_umma_$L.ncu-rep --page raw --csv > gpurun_out/r02_umma_${L}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_umma_$L.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_umma_${L}_src.csv 2>/dev/null
  ncu -i /tmp/prof_umma_$L.ncu-rep --page details > gpurun_out/r02_umma_${L}_details.txt 2>/dev/null
done
ls -la gpurun_out | grep umma
