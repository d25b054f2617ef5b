#!/usr/bin/env python
"""Writes tests/golden/*.npz: outputs of the CPU oracle on small seeded inputs, committed so that (a) later edits of the oracle cannot
drift silently and (b) the CUDA path is checked against fixed numbers as well as against the live oracle.
The reference itself cannot produce these (Fortran, no compiler in the image or on the GPU box): the vectors are the oracle's, and the
oracle is pinned against the reference's own known answers in tests/test_oracle_known_answers.py.
Regenerate with:  python tests/golden/make_golden.py   (only after a deliberate, reviewed change of the restatement)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from phantom_b200 import setups  # noqa: E402
from oraclelib import Oracle  # noqa: E402


def cases():
    """name -> Particles (inputs are regenerated from seeds by the tests, only outputs are stored)"""
    out = {}
    part, _ = setups.setup_test_derivs(nx=8, lattice="random")
    part.alphaind[:, 0] = 0.5
    out["hydro_random_512"] = part
    part, _ = setups.setup_test_derivs(nx=8, lattice="random", mhd=True)
    part.alphaind[:, 0] = 0.4
    out["mhd_random_512"] = part
    part = setups.setup_random_sphere(n=400)
    part.alphaind[:, 0] = 0.5
    out["gravity_sphere_400"] = part
    part, _ = setups.setup_dustybox(nx=8, idrag=1)
    out["dustybox_epstein_512"] = part
    return out


def outputs(part):
    sd, sf = Oracle(part.params).derivs(part)
    return dict(h=part.xyzh[:, 3], fxyzu=part.fxyzu, gradh=part.gradh, divcurlv=part.divcurlv, dBevol=part.dBevol, poten=part.poten,
                tstop=part.tstop, dustfrac=part.dustfrac, alphaloc=part.alphaind[:, 1],
                scalars=np.array([sf.dtcourant, sf.dtforce, sd.rhomax, float(sd.nactualtot), float(sf.npairs_force)]))


if __name__ == "__main__":
    for name, part in cases().items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **outputs(part))
        print("wrote", name)
