#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "fwd/" -k regex:"k_rulebook_tiles" -s 7 -c 1 -o /tmp/prof_rb_v2 python tools/one_forward.py > gpurun_out/r02_ncu_rb_v2.log 2>&1; echo "full rc=$?"
ncu -i /tmp/prof_rb_v2.ncu-rep --page raw --csv > gpurun_out/r02_rb_v2_raw.csv 2>/dev/null
ncu -i /tmp/prof_rb_v2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_rb_v2_src.csv 2>/dev/null
ncu -i /tmp/prof_rb_v2.ncu-rep --page details > gpurun_out/r02_rb_v2_details.txt 2>/dev/null
