#!/bin/bash
# ncu --set full of the 3^4 ts1 rule-book build (8th k_rulebook_tiles launch of the forward), voxel-table and leaf-grid variants
mkdir -p gpurun_out
for LG in 0 1; do
  INSMOS_LEAFGRID=$LG timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "fwd/" -k regex:"k_rulebook_tiles" -s 7 -c 1 -o /tmp/prof_rb_$LG python tools/one_forward.py > gpurun_out/r02_ncu_rb_$LG.log 2>&1; echo "full lg=$LG rc=$?"
  ncu -i /tmp/prof_rb_$LG.ncu-rep --page raw --csv > gpurun_out/r02_rb_${LG}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_rb_$LG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_rb_${LG}_src.csv 2>/dev/null
  ncu -i /tmp/prof_rb_$LG.ncu-rep --page details > gpurun_out/r02_rb_${LG}_details.txt 2>/dev/null
done
ls -la gpurun_out | grep r02_rb
