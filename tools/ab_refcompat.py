import sys, os, json, time
sys.path.insert(0, "/root/repo")
import numpy as np
from phantom_b200 import setups
from phantom_b200.api import SphGpu
part = setups.setup_dustydisc(ngas=1000000, ndust=250000); part.params.dtmax = 1.0
part.alphaind[:, 0] = 1.0
for rc in (1, 0):
    g = SphGpu(part.params.copy()); g.set_option("refcompat_hmax", rc); g.set_timestep_bins(0, 0, 0)
    g.upload(part)
    for r in range(2):
        t = time.time(); sc = g.derivs_resident(1); wall = (time.time() - t) * 1e3
    print("refcompat", rc, round(wall, 2), g.kernel_timings_ms(), sc.npairs_force)
