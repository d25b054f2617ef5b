#!/usr/bin/env python
"""Full-size runs of the BASELINE.json configurations on one GPU: device-resident derivs, per-phase times, pair counts.
usage: tools/run_config.py {shock|turb|mhdblast|orstang|dustydisc|sphere} SIZE [reps]"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from phantom_b200 import setups
from phantom_b200.api import SphGpu

name, size = sys.argv[1], int(float(sys.argv[2]))
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t0 = time.time()
if name == "shock":
    part = setups.setup_shock(nx=size, width=int(os.environ.get("SHOCK_WIDTH", "1")))
elif name == "turb":
    part = setups.setup_turb(nx=size)
elif name == "mhdblast":
    part = setups.setup_mhdblast(nx=size)
elif name == "orstang":
    part = setups.setup_orstang(nx=size)
elif name == "dustydisc":
    part = setups.setup_dustydisc(ngas=size, ndust=size // 4)
    part.params.dtmax = 1.0
elif name == "dustybox":
    part, _ = setups.setup_dustybox(nx=size, idrag=2, lattice="cubic")
elif name == "sphere":
    rng = np.random.RandomState(1234)
    u = rng.uniform(-1, 1, size=(int(size * 2.2), 3)); u = u[np.sum(u * u, axis=1) < 1.0][:size]
    p = setups.setup_random_sphere(n=1000).params
    p.massoftype[1] = 1.0 / size
    xyzh = np.zeros((size, 4)); xyzh[:, :3] = u
    xyzh[:, 3] = p.hfact * (p.massoftype[1] / (1.0 / (4. / 3. * np.pi))) ** (1. / 3.)
    part = setups.Particles(p, xyzh); part.vxyzu[:, 3] = 0.05
else:
    raise SystemExit(__doc__)
part.alphaind[:, 0] = 1.0
tsetup = time.time() - t0
g = SphGpu(part.params.copy())
if part.params.ind_timesteps:
    g.set_timestep_bins(0, 0, 0)
peak = g.measure_fp64_peak()
import re
fc = {k: int(v) for k, v in re.findall(r"#define\s+(\w+)\s+(\d+)", open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "phantom_b200", "csrc",
                                                                                      "roofline_constants.h")).read())}
p = part.params
g.upload(part)
for r in range(reps):
    t = time.time(); sc = g.derivs_resident(1); wall = (time.time() - t) * 1e3
    nact = int(np.sum(part.iphase > 0))
    kt, gt = g.kernel_timings_ms(), (g.gravity_timings_ms() if p.gravity else None)
    # algorithmic flop of the pair kernels (roofline_constants.h x measured pair counts) against the FP64 peak measured on this device
    fd = (fc["FLOP_DENS_PAIR_HYDRO"] + (fc["FLOP_DENS_PAIR_MHD_EXTRA"] if p.mhd else 0)) * sc.npairs_density + fc["FLOP_DENS_EPILOGUE"] * sc.nrhocalc
    ff = ((fc["FLOP_FORCE_PAIR_ISOTHERMAL"] if p.isothermal else fc["FLOP_FORCE_PAIR_ADIABATIC"]) + (fc["FLOP_FORCE_PAIR_MHD_EXTRA"] if p.mhd else 0)
          + (fc["FLOP_FORCE_PAIR_GRAV_EXTRA"] if p.gravity else 0)) * sc.npairs_force + fc["FLOP_FORCE_EPILOGUE"] * nact
    roof = {"fp64_peak_tflops": round(peak, 2), "density_frac": round(fd / (kt["density"] * 1e-3) / 1e12 / peak, 4), "force_frac": round(ff / (kt["force"] * 1e-3) / 1e12 / peak, 4)}
    if p.gravity and gt["p2p"] > 0:
        roof["p2p_frac"] = round(fc["FLOP_GRAV_P2P_PAIR"] * sc.npairs_gravity / (gt["p2p"] * 1e-3) / 1e12 / peak, 4)
        roof["tree_and_walk_ms"] = round(gt["total"] - gt["p2p"], 3)
        roof["m2l_tflops_over_tree_and_walk"] = round(fc["FLOP_GRAV_M2L"] * sc.nm2l / ((gt["total"] - gt["p2p"]) * 1e-3) / 1e12, 3)
    print(json.dumps(dict(config=name, npart=part.npart, setup_s=round(tsetup, 1), wall_ms=round(wall, 2), updates_per_s=round(nact / (wall * 1e-3)),
                          phases={k: round(v, 3) for k, v in g.timings_ms().items()}, kernels={k: round(v, 3) for k, v in kt.items()},
                          gravity={k: round(v, 3) for k, v in gt.items()} if gt else None, roofline=roof,
                          neigh_mean=round(sc.actualmean, 2), neigh_max=sc.maxactual, trial_mean=round(sc.trialmean, 1), trial_max=sc.maxtrial,
                          its_mean=round(sc.nrhocalc / max(sc.np, 1), 3), npairs_force=sc.npairs_force,
                          npairs_gravity=sc.npairs_gravity, nm2l=sc.nm2l)), flush=True)
