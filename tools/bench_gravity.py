#!/usr/bin/env python
"""Timing of the self-gravity pass (C5: uniform random sphere, tree_accuracy 0.5) on one GPU: tree build, FMM walk, P2P."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from phantom_b200 import setups
from phantom_b200.api import SphGpu

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.RandomState(1234)
# numpy generator for large N (the ran2 sequence is used by the parity tests; here only the distribution matters)
u = rng.uniform(-1, 1, size=(int(n * 2.2), 3)); u = u[np.sum(u * u, axis=1) < 1.0][:n]
part = setups.setup_random_sphere(n=1000)
p = part.params
p.massoftype[1] = 1.0 / n
xyzh = np.zeros((n, 4)); xyzh[:, :3] = u
xyzh[:, 3] = p.hfact * (p.massoftype[1] / (1.0 / (4. / 3. * np.pi))) ** (1. / 3.)
part = setups.Particles(p, xyzh)
part.vxyzu[:, 3] = 0.05
part.alphaind[:, 0] = 1.0
g = SphGpu(p.copy())
g.upload(part)
for r in range(reps):
    t = time.time(); sc = g.derivs_resident(1); wall = (time.time() - t) * 1e3
    ph, kt, gt = g.timings_ms(), g.kernel_timings_ms(), g.gravity_timings_ms()
    print(json.dumps(dict(n=n, wall_ms=round(wall, 2), phases=ph, pair_kernels=kt, gravity=gt, npairs_gravity=sc.npairs_gravity, nm2l=sc.nm2l,
                          p2p_per_particle=sc.npairs_gravity / n, m2l_per_particle=sc.nm2l / n, npairs_force=sc.npairs_force,
                          p2p_tflops=25 * sc.npairs_gravity / (gt["p2p"] * 1e-3) / 1e12 if gt["p2p"] > 0 else None)))
