"""Committed golden vectors (tests/golden/, written by tests/golden/make_golden.py): the oracle must keep reproducing them, and the CUDA
path must match them within the north_star tolerances."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

CASES = sorted(make_golden.cases().keys())


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def close(a, b, tol):
    s = np.sqrt(np.mean(np.asarray(b, dtype=np.float64) ** 2)) + 1e-300
    return np.max(np.abs(np.asarray(a, dtype=np.float64) - b) / (np.abs(b) + s)) < tol


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    part = make_golden.cases()[name]
    got, ref = make_golden.outputs(part), load(name)
    for k in ("h", "fxyzu", "gradh", "divcurlv", "dBevol", "poten", "dustfrac", "alphaloc"):
        assert close(got[k], ref[k], 1e-12), k
    fin = ref["tstop"] < 1e28
    assert close(got["tstop"][fin], ref["tstop"][fin], 1e-12)
    assert np.array_equal(got["scalars"][3:], ref["scalars"][3:]) and close(got["scalars"][:3], ref["scalars"][:3], 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_matches_golden(name):
    from phantom_b200.api import SphGpu
    part = make_golden.cases()[name]
    sc = SphGpu(part.params.copy()).derivs(part)
    ref = load(name)
    assert np.max(np.abs(part.xyzh[:, 3] - ref["h"]) / ref["h"]) < 1e-10
    assert close(part.fxyzu, ref["fxyzu"], 1e-8)
    assert close(part.dBevol, ref["dBevol"], 1e-8)
    assert np.max(np.abs(part.gradh - ref["gradh"])) <= 3e-7 * np.max(np.abs(ref["gradh"]))
    assert np.max(np.abs(part.poten - ref["poten"])) <= 3e-7 * (np.max(np.abs(ref["poten"])) + 1e-300)
    assert int(ref["scalars"][3]) == sc.nactualtot and int(ref["scalars"][4]) == sc.npairs_force
    assert abs(sc.dtcourant - ref["scalars"][0]) <= 1e-10 * ref["scalars"][0] and abs(sc.dtforce - ref["scalars"][1]) <= 1e-8 * ref["scalars"][1]
