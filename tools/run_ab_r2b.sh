#!/bin/bash
# disordered-state A/B (round 2): leaf size and group packing on random positions / evolved turbulence
run() { echo -n "$1 | $2 : "; env $1 python bench.py $2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); p=d['roofline']['passes']; print('ms %.3f tree %.3f dens %.3f (k %.3f its %.2f) force %.3f (k %.3f) trial %.0f' % (d['ms_per_step'], d['phases_ms']['tree'], d['phases_ms']['dens'], p['density']['ms'], p['density']['its_mean'], d['phases_ms']['force'], p['force']['ms'], d['neighbours']['trial_mean']))"; }
for st in "--nx 128 --positions random" "--nx 128 --evolve 20"; do
  run "X=1" "$st"
  run "X=1" "$st --max-leaf 4"
  run "X=1" "$st --max-leaf 6"
  run "SPHGPU_GROUP_PACK=128" "$st"
  run "SPHGPU_GROUP_PACK=256" "$st --max-leaf 4"
  run "X=1" "$st --max-cell 24"
done
