for tag in a049b22 d05b572 ebdab15 head a049b22 head; do
  SPHGPU_LIB=$PWD/build/variants/libsphgpu_$tag.so python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['roofline']['passes']; print('$tag', round(d['ms_per_step'],3), d['phases_ms'], 'dens', round(p['density']['ms'],3), 'force', round(p['force']['ms'],3))"
done
