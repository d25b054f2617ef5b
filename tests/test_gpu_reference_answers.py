"""The CUDA path against what the REFERENCE holds, and against the oracle at BASELINE sizes:
* the density-contrast blob of test_derivs.f90:651-713 must give the reference's own integers (total 37263216, max 988,
  mean 57.466651861721814) on the GPU too, and agree with the oracle on all 648432 particles at the north_star tolerances;
* the analytic MHD / AV known answers of test_derivs.f90 at the reference's tolerances, through the C ABI;
* turb 128^3 (BASELINE configs[1], the bench workload) and a 100^3 MHD lattice: GPU == oracle on every particle."""
import numpy as np
import pytest

from phantom_b200 import setups
from phantom_b200.params import IGAS
from oraclelib import Oracle
from derivs_functions import Fields, nfailed_f, nfailed_v

pytestmark = pytest.mark.gpu
TOL_H, TOL_F = 1e-10, 1e-8


def gpu(params):
    from phantom_b200.api import SphGpu
    return SphGpu(params.copy())


def relmax(a, b):
    s = np.sqrt(np.mean(b.astype(np.float64) ** 2)) + 1e-300
    return np.max(np.abs(a - b) / (np.abs(b) + s))


def parity(po, pg, sdo, sfo, sg):
    assert np.max(np.abs(pg.xyzh[:, 3] - po.xyzh[:, 3]) / po.xyzh[:, 3]) < TOL_H
    assert sg.nactualtot == sdo.nactualtot and sg.maxactual == sdo.maxactual and sg.npairs_force == sfo.npairs_force
    assert relmax(pg.fxyzu[:, :3], po.fxyzu[:, :3]) < TOL_F
    if po.params.maxvxyzu == 4:
        assert relmax(pg.fxyzu[:, 3], po.fxyzu[:, 3]) < TOL_F
    if po.params.mhd:
        assert relmax(pg.dBevol, po.dBevol) < TOL_F
    assert abs(sg.dtcourant - sfo.dtcourant) <= 1e-10 * sfo.dtcourant


def test_density_contrast_blob_reference_integers():
    part, n, hblob = setups.setup_density_contrast()
    po, pg = part.copy(), part.copy()
    sdo, sfo = Oracle(po.params).derivs(po)
    sg = gpu(pg.params).derivs(pg)
    # the reference's own numbers (test_derivs.f90:698-707)
    assert sg.nactualtot == 37263216 and sg.maxactual == 988
    assert abs(sg.actualmean - 57.466651861721814) <= 2.e-16 * 57.466651861721814
    parity(po, pg, sdo, sfo, sg)
    f = Fields(pg.xyzh[:n], pg.params)
    assert nfailed_v(pg.xyzh[:n, 3], hblob, 3.6e-4)[0] == 0
    assert nfailed_f(pg.divcurlv[:n, 0], f.divv(), 1.e-3)[0] == 0
    assert nfailed_v(pg.gradh[:n, 0], 1.01948, 1.e-5)[0] == 0


def test_mhd_lattice_100_reference_tolerances_and_parity():
    # test_derivs.f90:465-512 on the GPU: analytic dB/dt, div B, curl B, MHD forces at the reference's tolerances; parity on 10^6 particles
    part, hzero = setups.setup_test_derivs(nx=100, dissipation=False, mhd=True, polyk=0.)
    Bext = (2.0e-1, 3.0e-1, 0.5)
    f = Fields(part.xyzh, part.params, Bext)
    part.vxyzu[:] = 0.
    part.vxyzu[:, 0], part.vxyzu[:, 1], part.vxyzu[:, 2] = f.vx(), f.vy(), f.vz()
    part.xyzh[:, 3] = hzero
    rho1 = 1. / (part.params.massoftype[IGAS] * (part.params.hfact / part.xyzh[:, 3]) ** 3)
    part.Bevol[:, 0], part.Bevol[:, 1], part.Bevol[:, 2], part.Bevol[:, 3] = f.Bx() * rho1, f.By() * rho1, f.Bz() * rho1, 0.
    po, pg = part.copy(), part.copy()
    sdo, sfo = Oracle(po.params).derivs(po)
    sg = gpu(pg.params).derivs(pg)
    parity(po, pg, sdo, sfo, sg)
    f = Fields(pg.xyzh, pg.params, Bext)
    rho1 = 1. / (pg.params.massoftype[IGAS] * (pg.params.hfact / pg.xyzh[:, 3]) ** 3)
    Bx, By, Bz = f.Bx(), f.By(), f.Bz()
    cx, cy, cz = f.curlB()
    for x, val, tol in ((pg.divBsymm, f.divB(), 2.e-3), (pg.dBevol[:, 0], rho1 * Bx * f.dvxdx(), 2.e-3),
                        (pg.dBevol[:, 1], rho1 * (Bx * f.dvydx() + Bz * f.dvydz()), 2.e-3), (pg.dBevol[:, 2], rho1 * By * f.dvzdy(), 2.e-2),
                        (pg.fxyzu[:, 0], -(By * f.dBydx() - Bz * f.dBxdz()) / 5.0, 2.5e-2), (pg.fxyzu[:, 1], -(Bz * f.dBzdy() - Bx * f.dBydx()) / 5.0, 2.5e-2),
                        (pg.fxyzu[:, 2], -(Bx * f.dBxdz() - By * f.dBzdy()) / 5.0, 2.5e-2), (pg.divcurlB[:, 0], f.divB(), 1.e-3),
                        (pg.divcurlB[:, 1], cx, 1.e-3), (pg.divcurlB[:, 2], cy, 1.e-3), (pg.divcurlB[:, 3], cz, 1.e-3)):
        assert nfailed_f(x, val, tol)[0] == 0


def test_turb_128_full_size_parity():
    # BASELINE configs[1], the bench workload: 2 097 152 particles, GPU == oracle on every particle
    part = setups.setup_turb(nx=128)
    po, pg = part.copy(), part.copy()
    sdo, sfo = Oracle(po.params).derivs(po)
    sg = gpu(pg.params).derivs(pg)
    parity(po, pg, sdo, sfo, sg)
    assert np.array_equal(pg.alphaind[:, 1], po.alphaind[:, 1]) or np.max(np.abs(pg.alphaind[:, 1] - po.alphaind[:, 1])) < 1e-6


def test_dustybox_on_the_device_matches_the_analytic_decay():
    """DUSTYBOX (test_dust.f90:132-311) through sphgpu_step_resident: 100 steps, the reference's cubic-kernel tolerances on v and f of both
    phases and on E_kin at every step"""
    import math
    from phantom_b200.params import default_params, IDUST
    from phantom_b200.api import SphGpu
    nx = 24
    dz = 2. * math.sqrt(6.) / nx
    p = default_params(dust=1, idrag=2, K_code=0.35, ieos=1, polyk=1., gamma=1., alpha=0., alphamax=0., alphau=0., alphaB=0., tolh=1.e-4,
                       xmin=-0.5, xmax=0.5, ymin=-0.25, ymax=0.25, zmin=-dz, zmax=dz)
    lat = setups.unifdis_closepacked(p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax, 1. / nx, p.hfact, periodic=True)
    n1 = len(lat)
    totmass = (p.xmax - p.xmin) * (p.ymax - p.ymin) * (p.zmax - p.zmin)
    p.massoftype[IGAS] = totmass / n1
    p.massoftype[IDUST] = totmass / n1
    iphase = np.concatenate([np.full(n1, IGAS, dtype=np.int8), np.full(n1, IDUST, dtype=np.int8)])
    part = setups.Particles(p, np.concatenate([lat, lat]), iphase)
    dust = iphase == IDUST
    part.vxyzu[dust, 0] = 1.
    g = SphGpu(part.params.copy())
    g.upload(part)
    g.derivs_resident(1)
    K, dt, t = 0.35, 1.e-3, 0.
    for it in range(100):
        t += dt
        g.step_resident(dt)
        dv = math.exp(-2. * K * t)
        vg, vd = 0.5 * (1. - dv), 0.5 * (1. + dv)
        fd = K * (vg - vd)
        if it % 10 == 9 or it < 3:
            g.download(part)
            for x, val, tol in ((part.vxyzu[dust, 0], vd, 1.e-4), (part.fxyzu[dust, 0], fd, 3.e-3), (part.vxyzu[~dust, 0], vg, 1.e-4),
                                (part.fxyzu[~dust, 0], -fd, 3.e-3)):
                assert nfailed_v(x, val, tol)[0] == 0, it
        ekin = g.energies_resident().ekin
        assert nfailed_v(np.array([ekin]), 0.5 * totmass * (vd ** 2 + vg ** 2), 1.e-4)[0] == 0, it
