#!/bin/bash
# ncu --set full capture of the GENERAL pair kernels (two-fluid box) with per-source-line export -> gpurun_out/prof_gen_*.csv
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'^k_force$|^k_density$' -s 2 -c 2 -o /tmp/prof_gen -f \
    python tools/run_config.py dustybox ${1:-64} 2 > gpurun_out/prof_gen.log 2>&1
ncu -i /tmp/prof_gen.ncu-rep --page raw --csv > gpurun_out/prof_gen.raw.csv 2>/dev/null
ncu -i /tmp/prof_gen.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_density > gpurun_out/prof_gen_dens.lines.csv 2>/dev/null
ncu -i /tmp/prof_gen.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_force > gpurun_out/prof_gen_force.lines.csv 2>/dev/null
ls -la gpurun_out/prof_gen*
