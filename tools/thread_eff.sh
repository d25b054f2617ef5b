#!/bin/bash
# thread efficiency (active lanes per executed warp instruction) of the pair kernels on several configurations: ncu, two metrics
mkdir -p gpurun_out
for c in "orstang 128" "shock 256" "sphere 1e6" "mhdblast 100"; do
  set -- $c
  ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --kernel-name-base demangled \
      -k regex:'k_force_fast|k_density' -s 2 -c 2 --csv python tools/run_config.py $1 $2 2 2>/dev/null | grep -v "^==" | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=[r for r in rows if 'Kernel Name' in r]
if hdr:
    h=hdr[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value')
    out={}
    for r in rows:
        if len(r)>iv and r is not h and r[ik]!='Kernel Name':
            out.setdefault(r[ik][:60],{})[r[im].split('.')[0][-30:]]=r[iv]
    for k,v in out.items(): print('$c', k, v)
"
done
