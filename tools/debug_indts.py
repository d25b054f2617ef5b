import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
from phantom_b200 import setups
from phantom_b200.params import IGAS
from phantom_b200.api import SphGpu
from oraclelib import Oracle

def run(istepfrac, byphase, icall):
    part, _ = setups.setup_test_derivs(nx=18, lattice="random", ind_timesteps=1)
    n = part.npart
    rng = setups.Ran2(-1357)
    nbinmax = 3
    part.ibin_old[:] = np.minimum((rng.draw(n) * (nbinmax + 1)).astype(np.int8), nbinmax)
    part.ibin[:] = part.ibin_old
    part.params.dtmax = 0.02
    active = (istepfrac % (2 ** (nbinmax - part.ibin.astype(np.int64)))) == 0
    part.iphase[~active] = -IGAS
    part.alphaind[:, 0] = 0.3
    part.gradh[:, 0] = 1.0
    ibinnow = 2
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    o.build_tree(po); o.densityiterate(po); po.params.set_boundaries_to_active = 0; o.set_params(po.params); o.cons2prim(po)
    so = o.force(po, icall, 0.0, nbinmax=nbinmax, ibinnow=ibinnow, istepfrac=istepfrac)
    g = SphGpu(pg.params.copy())
    g.set_timestep_bins(nbinmax, ibinnow, istepfrac)
    if not byphase:
        g.derivs(pg, icall=1)
    else:
        g.build_tree(pg); g.densityiterate(pg); pg.params.set_boundaries_to_active = 0; g.set_params(pg.params); g.cons2prim_everything(pg)
        g.force(pg, icall)
    fs = np.sqrt(np.mean(po.fxyzu[:, :3] ** 2))
    err = np.max(np.abs(pg.fxyzu[:, :3] - po.fxyzu[:, :3]), axis=1) / fs
    bad = np.where(err > 1e-8)[0]
    print(f"istepfrac={istepfrac} byphase={byphase} icall={icall}: nactive={active.sum()} nbad={len(bad)} maxerr={err.max():.3e}",
          "dh", np.max(np.abs(pg.xyzh[:, 3] - po.xyzh[:, 3]) / po.xyzh[:, 3]),
          "deos", np.max(np.abs(pg.eos_vars - po.eos_vars)), "dalpha", np.max(np.abs(pg.alphaind - po.alphaind)),
          "ddvdx", np.max(np.abs(pg.dvdx - po.dvdx)), "dgradh", np.max(np.abs(pg.gradh - po.gradh)), "ddivv", np.max(np.abs(pg.divcurlv - po.divcurlv)))
    if len(bad):
        print("  bad active?", active[bad][:10], "du err", np.max(np.abs(pg.fxyzu[:, 3] - po.fxyzu[:, 3])))

for args in [(2, False, 1), (2, True, 1), (2, True, 2), (4, True, 1), (4, False, 1)]:
    run(*args)
