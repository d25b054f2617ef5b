"""Stepping parity (BASELINE.json north_star: "total momentum and energy track the reference to 1e-10 over 100 steps"):
the device-resident leapfrog (sphgpu_step_resident, step_leapfrog.f90:95) + device energies (sphgpu_energies_resident,
energies.f90:64) against the numpy restatement of the same step driving the CPU oracle, on the same initial conditions."""
import numpy as np
import pytest

from phantom_b200 import setups
from oraclelib import Oracle
import steplib

pytestmark = pytest.mark.gpu


def run_pair(part, nsteps, dtfac=1.0):
    from phantom_b200.api import SphGpu
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    sc_o = steplib.oracle_derivs(o, po, 1)
    g = SphGpu(pg.params.copy())
    g.upload(pg)
    sc_g = g.derivs_resident(1)
    hist = []
    dto = dtfac * min(sc_o.dtcourant, sc_o.dtforce)
    dtg = dtfac * min(sc_g.dtcourant, sc_g.dtforce)
    for it in range(nsteps):
        assert abs(dtg - dto) <= 1e-9 * dto, (it, dtg, dto)
        sc_o, dterr_o, errmax_o, its_o = steplib.step_leapfrog(o, po, dto)
        out = g.step_resident(dtg)
        assert out.its == its_o, (it, out.its, its_o)
        eo, eg = steplib.energies(po), g.energies_resident()
        hist.append((eo, eg))
        dto = dtfac * min(sc_o.dtcourant, sc_o.dtforce, dterr_o)
        dtg = dtfac * min(out.dtcourant, out.dtforce, out.dterr)
    g.download(pg)
    return po, pg, hist


def check_hist(hist, tol=1e-10):
    e0 = hist[0][0]
    escale = abs(e0["ekin"]) + abs(e0["etherm"]) + abs(e0["emag"]) + abs(e0["epot"])
    pscale = np.sqrt(2. * max(e0["ekin"], 1e-300) * e0["mtot"])       # |p| scale: sqrt(2 m E_kin)
    for it, (eo, eg) in enumerate(hist):
        assert abs(eg.etot - eo["etot"]) <= tol * escale, (it, eg.etot, eo["etot"])
        for k in ("ekin", "etherm", "emag", "epot"):
            assert abs(getattr(eg, k) - eo[k]) <= tol * escale, (it, k)
        dmom = np.array([eg.xmom, eg.ymom, eg.zmom]) - eo["mom"]
        assert np.max(np.abs(dmom)) <= tol * pscale, (it, dmom)
        assert abs(eg.mtot - eo["mtot"]) <= 1e-13 * eo["mtot"]


def test_sod_shock_100_steps():
    # C1: SETUP=shock, 100 global steps
    part = setups.setup_shock(nx=16)
    part.alphaind[:, 0] = 1.0
    po, pg, hist = run_pair(part, 100)
    check_hist(hist)
    assert np.max(np.abs(pg.xyzh[:, 3] - po.xyzh[:, 3]) / po.xyzh[:, 3]) < 1e-9
    assert np.max(np.abs(pg.xyzh[:, :3] - po.xyzh[:, :3])) < 1e-10


def test_orszag_tang_30_steps():
    part = setups.setup_orstang(nx=24)
    part.alphaind[:, 0] = 1.0
    po, pg, hist = run_pair(part, 30)
    check_hist(hist)
    assert hist[-1][1].emag > 0.


def test_turbulent_box_30_steps():
    part = setups.setup_turb(nx=16)
    part.alphaind[:, 0] = 1.0
    po, pg, hist = run_pair(part, 30)
    check_hist(hist)


def test_selfgravitating_sphere_20_steps():
    part = setups.setup_random_sphere(n=2000)
    part.alphaind[:, 0] = 1.0
    po, pg, hist = run_pair(part, 20)
    check_hist(hist, tol=1e-9)
    assert hist[-1][1].epot < 0.
