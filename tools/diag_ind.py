"""per-substep phase / kernel times of individual-timestep stepping on a perturbed turbulent box"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from phantom_b200 import setups
from phantom_b200.api import SphGpu
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
part = setups.setup_turb(nx=nx, ind_timesteps=True)
part.alphaind[:, 0] = 1.0
rng = np.random.RandomState(3)
part.xyzh[:, :3] += 0.3 / nx * (rng.rand(part.npart, 3) - 0.5)
for rc in (0, 1):
    g = SphGpu(part.params.copy()); g.set_option("refcompat_hmax", rc)
    g.upload(part); g.set_timestep_bins(0, 0, 0)
    sc = g.derivs_resident(1)
    dtmax = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
    p = part.params.copy(); p.dtmax = dtmax; g.set_params(p)
    sc = g.derivs_resident(1)
    nb = int(sc.nbinmaxnew)
    g.init_step_resident(0., dtmax, nb)
    isf, t = 0, 0.
    for k in range(12):
        isf += 1
        na, nal = g.set_active_particles_resident(nb, isf)
        t0 = time.perf_counter(); out = g.step_ind_resident(t, dtmax / 2 ** nb, dtmax); w = (time.perf_counter() - t0) * 1e3
        print("refcompat", rc, "substep", k, "nbinmax", nb, "active", na, "wall_ms", round(w, 2), {a: round(b, 3) for a, b in g.timings_ms().items()}, {a: round(b, 3) for a, b in g.kernel_timings_ms().items()})
        t = isf / 2. ** nb * dtmax
        nbn = int(out.scalars.nbinmaxnew)
        if nbn != nb:
            isf = isf // 2 ** (nb - nbn) if nbn < nb else isf * 2 ** (nbn - nb)
            nb = nbn
        if isf == 2 ** nb:
            break
