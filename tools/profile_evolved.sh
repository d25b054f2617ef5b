#!/bin/bash
# ncu --set full capture of the density kernel on the EVOLVED state (big rounds, iterating set) with per-source-line export
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_density.*1, .bool.1, .bool.1>' -s 3 -c 1 -o /tmp/prof_evo -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --evolve 20 > gpurun_out/prof_evo.log 2>&1
ncu -i /tmp/prof_evo.ncu-rep --page raw --csv > gpurun_out/prof_evo.raw.csv 2>/dev/null
ncu -i /tmp/prof_evo.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_evo_dens.lines.csv 2>/dev/null
ls -la gpurun_out/prof_evo*; tail -3 gpurun_out/prof_evo.log | cut -c1-300
