#!/bin/bash
# A/B builds of the pair kernels with different launch bounds / pairs per trip: builds libsphgpu_<tag>.so variants (CPU side)
# usage: tools/variants.sh build   |   tools/variants.sh run   (run = on the GPU box: times each variant with bench.py)
cd "$(dirname "$0")/.."
VARIANTS=("base:")   # add "tag:-Dflags" entries to A/B a build.  Round 2 sets, all measured and rejected: 512/576-candidate rounds (profiles/r02_disordered_states_ab.txt); next-trip prefetch of the force record head (three-sector records: 1.38 -> 1.68 / 1.52 ms at 4 / 3 CTAs per SM; two-sector records: 1.297 -> 1.302 ms); XTRA_MINB=2 (two-fluid box 50.3 -> 55.2 ms); FORCE_MINB=5 with the two-sector records (1.298 -> 1.355 ms)
if [ "$1" == "build" ]; then
  mkdir -p build/variants
  for v in "${VARIANTS[@]}"; do
    tag=${v%%:*}; flags=${v#*:}
    ( cd phantom_b200/csrc && for f in ${VARIANT_FILES:-force density neigh}; do /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr $flags -c $f.cu -o ../../build/variants/${f}_$tag.o & done; wait
      /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../../build/variants/libsphgpu_$tag.so sphgpu.o tree.o ../../build/variants/density_$tag.o ../../build/variants/force_$tag.o cons2prim.o ../../build/variants/neigh_$tag.o halo.o gravity.o step.o dist.o -ldl )
    echo built $tag
  done
else
  for v in "${VARIANTS[@]}"; do
    tag=${v%%:*}
    if [ -n "$VARIANT_CMD" ]; then
      echo -n "$tag "; SPHGPU_LIB=$PWD/build/variants/libsphgpu_$tag.so $VARIANT_CMD 2>&1 | tail -1 | cut -c1-400
    else
    SPHGPU_LIB=$PWD/build/variants/libsphgpu_$tag.so python bench.py ${BENCH_ARGS:---nx 100} --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['roofline']['passes']; print('$tag', round(d['ms_per_step'],3), 'dens', round(p['density']['ms'],3), 'force', round(p['force']['ms'],3))"
    fi
  done
fi
