#!/bin/bash
# A/B on the GPU box: round-1 pair kernels (one warp per target group) against the round-2 CTA-per-group kernels
mkdir -p gpurun_out
for mode in 0 1; do
  for nx in 128 100; do
    SPHGPU_PAIR_CTA=$mode python bench.py --nx $nx --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_cta_${mode}_${nx}.err | tail -1 > gpurun_out/ab_cta_${mode}_${nx}.json
    python - <<PY
import json
d=json.load(open("gpurun_out/ab_cta_${mode}_${nx}.json")); p=d["roofline"]["passes"]
print("pair_cta=$mode nx=$nx ms/step %.3f dens %.3f (%.3f) force %.3f (%.3f) phases %s" % (d["ms_per_step"], p["density"]["ms"], p["density"]["frac_fp64"], p["force"]["ms"], p["force"]["frac_fp64"], {k: round(v,3) for k,v in d["phases_ms"].items()}))
PY
  done
done
