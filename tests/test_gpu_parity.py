"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's north_star: neighbour sets bit-exact; h and rho to 1e-10 relative;
accelerations, du/dt and dB/dt to 1e-8 relative.  real*4 outputs (gradh, divcurlv, dvdx, alphaind) are
compared at float precision because the reference stores them rounded (part.F90:50-52,114).
"""
import math
import numpy as np
import pytest

from phantom_b200 import setups
from phantom_b200.params import IGAS, IBOUNDARY
from oraclelib import Oracle

pytestmark = pytest.mark.gpu

TOL_H = 1e-10
TOL_F = 1e-8
TOL_R4 = 3e-7     # two float ulps: value differences of 1e-15 can flip the real*4 rounding


def gpu(params, **opts):
    from phantom_b200.api import SphGpu
    g = SphGpu(params.copy())
    for k, v in opts.items():
        g.set_option(k, v)
    return g


def rel_err(a, b, floor):
    return np.max(np.abs(a - b) / (np.abs(b) + floor))


def run_both(part, by_phase=False, **opts):
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    sdo, sfo = o.derivs(po)
    g = gpu(pg.params, **opts)
    if by_phase:
        sdg, sfg = g.derivs_by_phase(pg)
    else:
        sfg = g.derivs(pg)
        sdg = sfg
    return po, pg, (sdo, sfo), (sdg, sfg), g


def check_hydro(po, pg, mhd=False):
    m = po.params.massoftype[IGAS]
    ho, hg = po.xyzh[:, 3], pg.xyzh[:, 3]
    assert np.array_equal(po.xyzh[:, :3], pg.xyzh[:, :3])
    assert rel_err(hg, ho, 0.) < TOL_H
    rho_o, rho_g = m * (po.params.hfact / ho) ** 3, m * (pg.params.hfact / hg) ** 3
    assert rel_err(rho_g, rho_o, 0.) < 3 * TOL_H
    assert rel_err(pg.gradh[:, 0], po.gradh[:, 0], 0.) < TOL_R4
    sc = np.sqrt(np.mean(po.dvdx.astype(np.float64) ** 2)) + 1e-30
    assert np.max(np.abs(pg.dvdx - po.dvdx)) < TOL_R4 * 30 * sc + 1e-6 * sc
    assert np.max(np.abs(pg.alphaind[:, 2] - po.alphaind[:, 2])) <= 1e-5 * (np.max(np.abs(po.alphaind[:, 2])) + 1e-30)
    assert np.max(np.abs(pg.alphaind[:, 1] - po.alphaind[:, 1])) <= 1e-5
    assert rel_err(pg.eos_vars[:, 0], po.eos_vars[:, 0], 1e-300) < 1e-9
    nvu = po.params.maxvxyzu
    fs = np.sqrt(np.mean(po.fxyzu[:, :3] ** 2))
    assert np.max(np.abs(pg.fxyzu[:, :3] - po.fxyzu[:, :3])) < TOL_F * fs * 10
    assert np.max(np.abs(pg.fxyzu[:, :3] - po.fxyzu[:, :3]) / (np.abs(po.fxyzu[:, :3]) + fs)) < TOL_F
    if nvu == 4:
        us = np.sqrt(np.mean(po.fxyzu[:, 3] ** 2)) + 1e-300
        assert np.max(np.abs(pg.fxyzu[:, 3] - po.fxyzu[:, 3]) / (np.abs(po.fxyzu[:, 3]) + us)) < TOL_F
    dsc = np.sqrt(np.mean(po.divcurlv.astype(np.float64) ** 2)) + 1e-30
    assert np.max(np.abs(pg.divcurlv - po.divcurlv)) < 1e-5 * dsc
    if mhd:
        bs = np.sqrt(np.mean(po.dBevol ** 2)) + 1e-300
        assert np.max(np.abs(pg.dBevol - po.dBevol) / (np.abs(po.dBevol) + bs)) < TOL_F


def test_lattice_known_answers_on_gpu():
    # test_derivs.f90:203-212 exact integers + :1116,:1123 constants, now produced by the CUDA path
    part, hzero = setups.setup_test_derivs(nx=32, dissipation=False)
    g = gpu(part.params)
    sc = g.derivs(part)
    n = part.npart
    assert sc.np == n and sc.actualmean == 57.0 and sc.maxactual == 57 and sc.nactualtot == 57 * n
    assert sc.nrhocalc == 2 * n
    assert np.max(np.abs(part.xyzh[:, 3] - hzero) / hzero) < 3.6e-4
    assert np.max(np.abs(part.gradh[:, 0] - 1.01948)) / 1.01948 < 1.e-5


@pytest.mark.parametrize("lattice,nx", [("cubic", 24), ("random", 20), ("closepacked", 20)])
def test_derivs_parity_adiabatic(lattice, nx):
    part, _ = setups.setup_test_derivs(nx=nx, lattice=lattice)
    part.alphaind[:, 0] = 0.5
    po, pg, so, sg, g = run_both(part)
    check_hydro(po, pg)
    assert abs(sg[1].dtcourant - so[1].dtcourant) <= 1e-10 * so[1].dtcourant
    assert abs(sg[1].dtforce - so[1].dtforce) <= 1e-8 * so[1].dtforce
    assert sg[0].nactualtot == so[0].nactualtot and sg[0].maxactual == so[0].maxactual
    assert sg[1].npairs_force == so[1].npairs_force


def test_derivs_parity_packed_target_groups():
    # option group_pack: target groups are runs of whole leaf cells instead of subtrees (tree.cu k_groups_packed); grouping must not
    # change neighbour sets or results beyond summation order
    part, _ = setups.setup_test_derivs(nx=20, lattice="random")
    part.alphaind[:, 0] = 0.5
    po, pg, so, sg, g = run_both(part, group_pack=256)
    check_hydro(po, pg)
    assert sg[0].nactualtot == so[0].nactualtot and sg[0].maxactual == so[0].maxactual
    assert sg[1].npairs_force == so[1].npairs_force


def test_derivs_parity_isothermal_random_h():
    part, _ = setups.setup_test_derivs(nx=20, lattice="random", isothermal=True)
    rng = setups.Ran2(-24358)
    part.xyzh[:, 3] *= (0.8 + 0.4 * rng.draw(part.npart))
    part.alphaind[:, 0] = 1.0
    po, pg, so, sg, g = run_both(part, by_phase=True)
    check_hydro(po, pg)
    assert sg[0].nactualtot == so[0].nactualtot


@pytest.mark.parametrize("mhd,gravity", [(False, False), (True, False)])
def test_derivs_parity_iterating_set_second_call(mhd, gravity):
    """A set whose h-rho iteration did real work in the previous call takes the density instantiation that keeps the staged round and
    its hit masks over the iterations (density.cu: REUSE; masks built 0.5 % wide, the exact FP64 test decides).  Same state handed
    over twice: the second call must give the oracle's h, forces and neighbour totals and the iteration counts of the plain
    instantiation, whether a particle needs one, two or more iterations."""
    part, _ = setups.setup_test_derivs(nx=22, lattice="random", mhd=mhd)
    rng = setups.Ran2(-9753)
    fac = 0.9 + 0.25 * rng.draw(part.npart)
    fac[::7] = 1.0                                   # some particles start converged
    part.alphaind[:, 0] = 0.5
    po = part.copy()
    o = Oracle(po.params)
    o.derivs(po)                                     # converged h
    start = part.copy()
    start.xyzh[:, 3] = po.xyzh[:, 3] * fac
    po2 = start.copy()
    o2 = Oracle(po2.params)
    sdo, sfo = o2.derivs(po2)
    assert sdo.nrhocalc > 1.5 * sdo.np               # the set iterates
    g = gpu(start.params)
    warm = start.copy()
    sw = g.derivs(warm)                              # first call: plain instantiation, sets the hint
    pg = start.copy()
    sg = g.derivs(pg)
    check_hydro(po2, pg, mhd=mhd)
    assert rel_err(pg.xyzh[:, 3], warm.xyzh[:, 3], 0.) < 1e-13                               # and so does the plain instantiation
    assert sg.nactualtot == sdo.nactualtot and sg.maxactual == sdo.maxactual and sg.npairs_force == sfo.npairs_force
    # iteration counts: per particle here, per leaf cell of its own tree in the oracle -- the two instantiations must agree with each other
    assert sg.nrhocalc == sw.nrhocalc and sg.nrhocalc > 1.5 * sg.np, (sg.nrhocalc, sw.nrhocalc, sdo.nrhocalc, sg.np)
    assert sg.npairs_density == sw.npairs_density


@pytest.mark.parametrize("const_av", [False, True])
def test_derivs_parity_isothermal_two_sector_records(const_av):
    """ieos = 1 without energy, MHD or self-gravity: full derivs on the device take the force instantiation whose neighbour records
    are two sectors wide, P/rho^2, rho and the gradW factor derived from h per pair (force.cu: iso1_derive).  Against the oracle, and
    against the three-sector instantiation (option no_iso1) of the same library."""
    part, _ = setups.setup_test_derivs(nx=22, lattice="random", isothermal=True)
    rng = setups.Ran2(-1357)
    part.xyzh[:, 3] *= (0.9 + 0.2 * rng.draw(part.npart))
    part.alphaind[:, 0] = 0.1 + 0.9 * rng.draw(part.npart)
    if const_av:
        part.params.const_av = 1
        part.params.alpha = 0.7
    po, pg, so, sg, g = run_both(part)
    if const_av:      # (the reference does not fill dvdx / alpha_loc with constant alpha; h and forces are what this test is about)
        fso = np.sqrt(np.mean(po.fxyzu[:, :3] ** 2))
        assert rel_err(pg.xyzh[:, 3], po.xyzh[:, 3], 0.) < TOL_H
        assert np.max(np.abs(pg.fxyzu[:, :3] - po.fxyzu[:, :3]) / (np.abs(po.fxyzu[:, :3]) + fso)) < TOL_F
    else:
        check_hydro(po, pg)
    assert sg[1].npairs_force == so[1].npairs_force
    p3 = part.copy()
    s3 = gpu(p3.params, no_iso1=1).derivs(p3)
    fs = np.sqrt(np.mean(p3.fxyzu[:, :3] ** 2))
    assert np.array_equal(p3.xyzh, pg.xyzh)
    d = np.max(np.abs(p3.fxyzu[:, :3] - pg.fxyzu[:, :3]))
    assert 0. < d < 1e-12 * fs                                   # different instantiations (not bitwise), same forces
    assert abs(s3.dtcourant - sg[1].dtcourant) <= 1e-13 * s3.dtcourant and s3.npairs_force == sg[1].npairs_force


def test_derivs_parity_quintic():
    part, _ = setups.setup_test_derivs(nx=18, lattice="random", kernel=1, hfact=1.0)
    part.alphaind[:, 0] = 0.3
    po, pg, so, sg, g = run_both(part)
    check_hydro(po, pg)


def test_derivs_parity_mhd():
    part, _ = setups.setup_test_derivs(nx=20, lattice="random", mhd=True)
    part.alphaind[:, 0] = 0.4
    po, pg, so, sg, g = run_both(part)
    check_hydro(po, pg, mhd=True)


def test_neighbour_sets_bit_exact():
    # test_neigh.f90:264-367 restated against the CUDA walk: sets (not only counts) equal the oracle's
    part, _ = setups.setup_test_derivs(nx=16, lattice="random")
    rng = setups.Ran2(-24358)
    part.xyzh[:, 3] *= (0.6 + 1.2 * rng.draw(part.npart))
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    o.build_tree(po)
    g = gpu(pg.params)
    g.build_tree(pg)
    for sym in (False, True):
        offo, lsto = o.neighbour_sets(po, symmetric=sym)
        offg, lstg = g.neighbour_sets(pg.npart, symmetric=sym)
        assert np.array_equal(offo, offg)
        rows = np.repeat(np.arange(pg.npart), np.diff(offg))
        order = np.lexsort((lstg, rows))
        assert np.array_equal(lstg[order], lsto)
        tot, cnt = o.neighbour_counts_bruteforce(po, symmetric=sym)
        assert np.array_equal(np.diff(offg), cnt)


def test_boundary_and_inactive_particles():
    # boundary particles contribute as neighbours but receive no update (dens.F90:1329, force.F90:2255)
    part, _ = setups.setup_test_derivs(nx=16, lattice="random")
    part.iphase[::7] = IBOUNDARY
    part.params.massoftype[IBOUNDARY] = part.params.massoftype[IGAS]
    part.params.set_boundaries_to_active = 0
    part.gradh[:, 0] = 1.0
    part.alphaind[:, 0] = 0.2
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    o.build_tree(po); o.densityiterate(po); o.cons2prim(po); o.force(po)
    g = gpu(pg.params)
    g.derivs(pg)
    check_hydro(po, pg)
    b = part.iphase == IBOUNDARY
    assert np.array_equal(pg.xyzh[b, 3], part.xyzh[b, 3]) and np.all(pg.fxyzu[b] == 0.)


@pytest.mark.parametrize("max_leaf", [1, 3])
def test_tree_cell_capacity_retry(max_leaf):
    """The tree is built with ONE host round trip on cell arrays sized from n/3 (or the previous build); a set that makes more leaf
    cells than that -- one particle per cell here -- raises the overflow flag on the device and the build is repeated at the true
    size (tree.cu).  Results must not depend on it."""
    part, _ = setups.setup_test_derivs(nx=16, lattice="random")
    part.alphaind[:, 0] = 0.5
    po, pg, so, sg, g = run_both(part, max_leaf=max_leaf)
    check_hydro(po, pg)
    assert sg[0].nactualtot == so[0].nactualtot and sg[1].npairs_force == so[1].npairs_force
    pg2 = part.copy()
    sg2 = g.derivs(pg2)                              # second build on the same context: capacity remembered from the first
    assert np.array_equal(pg2.xyzh, pg.xyzh) and np.array_equal(pg2.fxyzu, pg.fxyzu)


def test_three_sort_classes_mixed_types():
    """Gas, dust and a third kind (star-type particles) in one set: the sort key carries the class, so leaf cells, target groups and
    staged rounds are single-class and the general kernels choose the pair body per class pair.  Every class gets the density of its
    own kind, gas alone gets hydro forces (dens.F90:717-741, force.F90:1539)."""
    from phantom_b200.params import IDUST
    part, _ = setups.setup_dustybox(nx=14, idrag=2, lattice="random")
    ISTAR = 4
    rng = setups.Ran2(-4242)
    star = (rng.draw(part.npart) < 0.1) & (part.iphase == IGAS)
    part.iphase[star] = ISTAR
    part.params.massoftype[ISTAR] = 0.3 * part.params.massoftype[IGAS]
    part.xyzh[star, 3] *= 2.0
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    sdo, sfo = o.derivs(po)
    g = gpu(pg.params)
    sg = g.derivs(pg)
    assert rel_err(pg.xyzh[:, 3], po.xyzh[:, 3], 0.) < TOL_H
    fs = np.sqrt(np.mean(po.fxyzu[:, :3] ** 2))
    assert np.max(np.abs(pg.fxyzu[:, :3] - po.fxyzu[:, :3]) / (np.abs(po.fxyzu[:, :3]) + fs)) < TOL_F
    assert np.all(pg.fxyzu[star, :3] == po.fxyzu[star, :3])
    assert sg.nactualtot == sdo.nactualtot and sg.npairs_force == sfo.npairs_force
    assert np.max(np.abs(pg.dustfrac - po.dustfrac)) <= 1e-12 * (np.max(np.abs(po.dustfrac)) + 1e-300)


def test_periodic_wrap_and_errors():
    from phantom_b200.api import SphGpuError
    part, _ = setups.setup_test_derivs(nx=10, lattice="random")
    part.xyzh[0, 0] += 1.0      # outside the box: build_tree wraps it in place (kdtree.F90:387)
    ref = part.xyzh[0, 0] - 1.0
    g = gpu(part.params)
    g.build_tree(part)
    assert abs(part.xyzh[0, 0] - ref) < 1e-15
    part.xyzh[3, 1] = np.nan
    with pytest.raises(SphGpuError) as e:
        g.build_tree(part)
    assert e.value.code == 3
    part.xyzh[:, 3] = -1.0      # all dead
    with pytest.raises(SphGpuError) as e:
        g.build_tree(part)
    assert e.value.code == 4


def test_resident_equals_literal():
    part, _ = setups.setup_test_derivs(nx=16, lattice="random")
    pa, pb = part.copy(), part.copy()
    g = gpu(pa.params)
    g.derivs(pa)
    g2 = gpu(pb.params)
    g2.upload(pb)
    g2.derivs_resident(1)
    g2.download(pb)
    for k in ("xyzh", "fxyzu", "gradh", "divcurlv", "dvdx", "alphaind", "eos_vars"):
        assert np.array_equal(getattr(pa, k), getattr(pb, k)), k
    assert g2.launch_count() > 10
