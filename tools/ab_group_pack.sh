#!/bin/bash
mkdir -p gpurun_out
( SPHGPU_GROUP_PACK=256 timeout 40 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/pack_tests.log 2>&1
for v in 256 0 1024; do
  SPHGPU_GROUP_PACK=$v timeout 25 python bench.py --nx 100 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['roofline']['passes']; print('pack=$v', round(d['ms_per_step'],3), d['phases_ms'], 'kd', round(p['density']['ms'],3), 'kf', round(p['force']['ms'],3), d['neighbours'])" >> gpurun_out/pack_bench.log 2>&1
done
cat gpurun_out/pack_tests.log gpurun_out/pack_bench.log
