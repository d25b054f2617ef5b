python -m pytest tests/test_gpu_reference_answers.py -x -q 2>&1 | tail -4
for args in "--nx 100" "--nx 128 --positions random" "--nx 128"; do
  echo "== $args"
  BENCH_ARGS="$args --no-extras" tools/variants.sh run
done
echo "== evolve"
python bench.py --nx 128 --evolve 20 --steps 10 --no-cpu-baseline --no-extras 2>gpurun_out/evolve.err | tail -1 > gpurun_out/bench_evolve20.json
python bench.py --nx 128 --positions random --evolve 20 --steps 10 --no-cpu-baseline --no-extras 2>gpurun_out/evolve_r.err | tail -1 > gpurun_out/bench_random_evolve20.json
python - <<'PY'
import json
for f in ("gpurun_out/bench_evolve20.json","gpurun_out/bench_random_evolve20.json"):
    try:
        d=json.load(open(f)); p=d["roofline"]["passes"]
        print(f, "ms/step %.3f value %.3e dens %.3f (its %.3f, frac %.3f) force %.3f (%.3f) trial %.1f" % (d["ms_per_step"], d["value"], p["density"]["ms"], p["density"]["its_mean"], p["density"]["frac_fp64"], p["force"]["ms"], p["force"]["frac_fp64"], d["neighbours"]["trial_mean"]))
    except Exception as e: print(f, "failed", e)
PY
tail -3 gpurun_out/evolve.err gpurun_out/evolve_r.err
