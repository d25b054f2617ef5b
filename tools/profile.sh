#!/bin/bash
# ncu captures for profiles/: [launches] launch list of one bench run; full capture of the two pair kernels
set -x
mkdir -p gpurun_out
if [ "$1" == "launches" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
fi
ncu --set full --clock-control none --import-source on -k regex:'k_density|k_force' -s 2 -c 2 -o gpurun_out/prof_pair -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof.log 2>&1
ls -la gpurun_out
