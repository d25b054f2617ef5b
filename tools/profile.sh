#!/bin/bash
# ncu captures for profiles/ (run on the GPU box through gpurun):
#   launches        : launch list (gpu__time_duration) of one bench.py run            -> gpurun_out/launches.csv
#   pair kernels    : --set full capture of the density / force pair kernels           -> gpurun_out/prof_pair.ncu-rep
#   gravity kernels : launch list + --set full capture of k_g_p2p / k_g_walk (C5, 1M)  -> gpurun_out/launches_grav.csv, prof_grav.ncu-rep
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_density|k_force' -s 2 -c 2 -o gpurun_out/prof_pair -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof.log 2>&1
if [ "$1" == "gravity" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_grav.csv \
    python tools/bench_gravity.py 1e6 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:'k_g_p2p|k_g_walk|k_force_fast|k_density' -s 30 -c 24 -o gpurun_out/prof_grav -f \
    python tools/bench_gravity.py 1e6 2 > /dev/null 2>&1
fi
# gpurun_out/ is capped at 64 MiB: keep the raw-metric CSVs (read by tools/make_profile_summaries.py), drop the big reports
for r in prof_pair prof_grav; do [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; done
rm -f gpurun_out/prof_grav.ncu-rep
[ "$KEEP_REP" == "1" ] || rm -f gpurun_out/prof_pair.ncu-rep
ls -la gpurun_out
