for tag in base x2; do
  for c in "dustybox 128" "dustydisc 1e6" "shock 256"; do
    echo -n "$tag $c: "; SPHGPU_LIB=$PWD/build/variants/libsphgpu_$tag.so python tools/run_config.py $c 2 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['wall_ms'], d['kernels'])"
  done
done
