"""GPU parity tests for the remaining physics rows of SURVEY.md section 8: individual timestep bins
(force.F90:1346-1358, :3272-3310), two-fluid gas-dust drag (force.F90:1852-1989, dust.f90:161-276) and tree
self-gravity (kdtree.F90:531-929 build with moments, :1357-1840 FMM, force.F90:1303-1339, :1992-2053, :2909-2927).
The CUDA path runs through the C ABI; the checker is the CPU oracle on the same seeded inputs.
Tolerances: BASELINE.json north_star (h, rho 1e-10; accelerations, du/dt 1e-8 relative; integer outputs exact)."""
import numpy as np
import pytest

from phantom_b200 import setups
from phantom_b200.params import IGAS, IBOUNDARY, IDUST
from oraclelib import Oracle

pytestmark = pytest.mark.gpu

TOL_H = 1e-10
TOL_F = 1e-8


def gpu(params):
    from phantom_b200.api import SphGpu
    return SphGpu(params.copy())


def relmax(a, b):
    s = np.sqrt(np.mean(b.astype(np.float64) ** 2)) + 1e-300
    return np.max(np.abs(a - b) / (np.abs(b) + s))


# ------------------------------------------------------------------------------------------------
#  individual timesteps
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("refcompat", [True, False])
@pytest.mark.parametrize("istepfrac,icall", [(0, 1), (4, 1), (2, 2)])
def test_individual_timestep_bins(istepfrac, icall, refcompat):
    part, _ = setups.setup_test_derivs(nx=18, lattice="random", ind_timesteps=1)
    n = part.npart
    rng = setups.Ran2(-1357)
    nbinmax = 3
    part.ibin_old[:] = np.minimum((rng.draw(n) * (nbinmax + 1)).astype(np.int8), nbinmax)
    part.ibin[:] = part.ibin_old
    part.params.dtmax = 0.02
    # activity as set by the step routine: active iff mod(istepfrac, 2**(nbinmax-ibin)) == 0 (utils_indtimesteps.f90:114-178)
    active = (istepfrac % (2 ** (nbinmax - part.ibin.astype(np.int64)))) == 0
    part.iphase[~active] = -IGAS
    part.alphaind[:, 0] = 0.3
    part.gradh[:, 0] = 1.0
    ibinnow = nbinmax if istepfrac % 2 else int(nbinmax - np.log2(np.gcd(istepfrac, 2 ** nbinmax)) if istepfrac else 0)
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    o.build_tree(po); o.densityiterate(po); po.params.set_boundaries_to_active = 0; o.set_params(po.params); o.cons2prim(po)
    so = o.force(po, icall, 0.0, nbinmax=nbinmax, ibinnow=ibinnow, istepfrac=istepfrac)
    g = gpu(pg.params)
    if not refcompat:
        g.set_option("refcompat_hmax", 0)
    g.set_timestep_bins(nbinmax, ibinnow, istepfrac)
    sg = g.derivs(pg, icall=1) if icall == 1 else None
    if icall == 2:
        g.build_tree(pg); g.densityiterate(pg); pg.params.set_boundaries_to_active = 0; g.set_params(pg.params); g.cons2prim_everything(pg)
        sg = g.force(pg, 2)
    assert active.sum() > 0 and (istepfrac == 0 or (~active).sum() > 0)
    assert np.max(np.abs(pg.xyzh[:, 3] - po.xyzh[:, 3]) / po.xyzh[:, 3]) < TOL_H
    tot, cnt = o.neighbour_counts_bruteforce(po, symmetric=True)
    exact_pairs = int(np.sum(cnt[active]))
    fs = np.sqrt(np.mean(po.fxyzu[active, :3] ** 2))
    err = np.max(np.abs(pg.fxyzu[:, :3] - po.fxyzu[:, :3]), axis=1) / fs
    if refcompat:
        # Default with individual timesteps: the reference's own neighbour sets.  The reference's force walk misses pairs that only an
        # inactive j reaches when j's leaf re-walked during the h-rho iteration: set_hmaxcell then stores 1.01*max(h) over the ACTIVE
        # members (dens.F90:343-345, :1275-1289; neigh_kdtree.f90:115-131) and the walk prunes on it (kdtree.F90:1288-1293).  The CUDA
        # path rebuilds the reference's tree, replays that hmax history and drops exactly those pairs: same pair count, same forces,
        # same bins and wake flags on EVERY particle.
        assert sg.npairs_force == so.npairs_force
        assert np.max(err[active]) < TOL_F
        assert np.array_equal(pg.ibin, po.ibin) and np.array_equal(pg.ibin_wake, po.ibin_wake) and sg.nbinmaxnew == so.nbinmaxnew
        if istepfrac:
            assert so.npairs_force <= exact_pairs
    else:
        # option refcompat_hmax = 0: the pair criterion q2i < R^2 .or. q2j < R^2 evaluated exactly = the O(N^2) count
        # (test_neigh.f90:264-367); every particle whose force differs from the reference's is accounted for by a missed pair
        assert sg.npairs_force == exact_pairs and so.npairs_force <= exact_pairs
        assert np.sum(err[active] > TOL_F) <= exact_pairs - so.npairs_force
    # inactive particles keep what they had (force.F90:2255)
    assert np.array_equal(pg.fxyzu[~active], part.fxyzu[~active])


# ------------------------------------------------------------------------------------------------
#  two-fluid dust
# ------------------------------------------------------------------------------------------------
def dusty_box(nx=16, idrag=2, isothermal=False, seed=-2468):
    return setups.setup_dustybox(nx=nx, idrag=idrag, isothermal=isothermal, seed=seed)


@pytest.mark.parametrize("idrag,isothermal", [(2, False), (3, True), (1, False)])
def test_two_fluid_dust_drag(idrag, isothermal):
    part, isdust = dusty_box(idrag=idrag, isothermal=isothermal)
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    sdo, sfo = o.derivs(po)
    g = gpu(pg.params)
    sg = g.derivs(pg)
    assert np.max(np.abs(pg.xyzh[:, 3] - po.xyzh[:, 3]) / po.xyzh[:, 3]) < TOL_H
    assert sg.nactualtot == sdo.nactualtot
    gas = ~isdust
    assert np.all(po.dustfrac[gas] > 0.) and np.all(po.dustfrac[isdust] == 0.)
    assert np.max(np.abs(pg.dustfrac - po.dustfrac)) < 1e-10 * np.max(po.dustfrac)
    assert relmax(pg.fxyzu[:, :3], po.fxyzu[:, :3]) < TOL_F
    if not isothermal:
        assert relmax(pg.fxyzu[gas, 3], po.fxyzu[gas, 3]) < TOL_F
    assert np.all(np.isfinite(po.tstop)) and np.min(po.tstop) > 0.
    assert np.max(np.abs(pg.tstop - po.tstop) / po.tstop) < 1e-9
    assert abs(sg.dtforce - sfo.dtforce) <= 1e-8 * sfo.dtforce
    assert abs(sg.dtcourant - sfo.dtcourant) <= 1e-10 * sfo.dtcourant
    assert sg.npairs_force == sfo.npairs_force
    if idrag == 1:   # both drag regimes are exercised
        assert np.min(po.tstop) < 0.5 * np.max(po.tstop)


# ------------------------------------------------------------------------------------------------
#  self-gravity
# ------------------------------------------------------------------------------------------------
def node_table_oracle(o, npart):
    ids = np.abs(o.inodeparts(npart))
    out = {}
    for n in range(1, o.ncells() + 1):
        rec, irec = o.node(n)
        i1, i2 = irec[4], irec[5]
        if i1 <= 0 or i2 < i1:
            continue
        out[tuple(sorted(ids[i1 - 1:i2]))] = (rec, irec[0] == 0)
    return out


@pytest.mark.parametrize("n,kernel", [(3000, 0), (2000, 1)])
def test_selfgravity_fmm_parity(n, kernel):
    part = setups.setup_random_sphere(n=n)
    if kernel == 1:
        part.params.kernel = 1
        part.params.hfact = 1.0
        part.xyzh[:, 3] *= 1.0 / 1.2
    rng = setups.Ran2(-97531)
    part.vxyzu[:, :3] = 0.1 * (rng.draw(3 * part.npart).reshape(-1, 3) - 0.5)
    part.alphaind[:, 0] = 0.5
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    sdo, sfo = o.derivs(po)
    g = gpu(pg.params)
    sg = g.derivs(pg)
    assert np.max(np.abs(pg.xyzh[:, 3] - po.xyzh[:, 3]) / po.xyzh[:, 3]) < TOL_H
    assert np.max(np.abs(pg.gradh[:, 1] - po.gradh[:, 1])) <= 3e-7 * np.max(np.abs(po.gradh[:, 1]))
    # ---- the tree: same nodes (as particle sets) with the same kdnode records
    rec, irec, ids = g.gravity_tree(pg.npart)
    tab = node_table_oracle(o, po.npart)
    assert len(rec) == len(tab)
    worst = 0.
    for k in range(len(rec)):
        key = tuple(sorted(ids[irec[k, 3]:irec[k, 3] + irec[k, 4]]))
        assert key in tab
        ro, leaf = tab[key]
        assert leaf == (irec[k, 0] < 0)
        sc = np.array([1., 1., 1., 1., 1., ro[5]] + [ro[5]] * 6)
        worst = max(worst, np.max(np.abs(rec[k] - ro) / sc))
        assert abs(rec[k, 4] - ro[4]) <= TOL_H * ro[4], (k, rec[k, 4], ro[4])   # node hmax: the replay of set_hmaxcell matches the tree's history
    assert worst < 1e-13
    # ---- forces and potential
    assert relmax(pg.fxyzu[:, :3], po.fxyzu[:, :3]) < TOL_F
    assert relmax(pg.fxyzu[:, 3], po.fxyzu[:, 3]) < TOL_F
    assert np.max(np.abs(pg.poten - po.poten)) <= 3e-7 * np.max(np.abs(po.poten))
    assert abs(sg.dtforce - sfo.dtforce) <= 1e-8 * sfo.dtforce
    assert sg.npairs_force == sfo.npairs_force
    assert sg.npairs_gravity > 0 and sg.nm2l > 0
    # momentum conservation of the gravitational + pressure force (test_gravity.f90:390-392 scale)
    m = pg.params.massoftype[IGAS]
    assert np.max(np.abs(np.sum(m * pg.fxyzu[:, :3], axis=0))) < 1e-3 * np.sum(m * np.linalg.norm(pg.fxyzu[:, :3], axis=1))


def test_selfgravity_against_direct_sum():
    # test_gravity.f90:300-400: tree force vs the direct sum on the uniform random sphere (tolerances 7.2e-3 / 6e-3 / 9.4e-3 on
    # the force components, 5.2e-4 on the total potential, potential = -3/5 GMM/R to 3.6e-2).  The direct sum is the oracle with
    # tree_accuracy = 0 (no node pair is ever accepted: every pair is summed particle by particle with softening).
    part = setups.setup_random_sphere(n=3000)
    part.params.alpha = 0.
    pd, pg = part.copy(), part.copy()
    pd.params.tree_accuracy = 0.0
    od = Oracle(pd.params)
    od.derivs(pd)
    g = gpu(pg.params)
    sg = g.derivs(pg)
    assert sg.nm2l > 0
    scale = np.max(np.abs(pd.fxyzu[:, :3]))
    for k, tol in enumerate((7.2e-3, 6.e-3, 9.4e-3)):
        assert np.max(np.abs(pg.fxyzu[:, k] - pd.fxyzu[:, k])) / scale < tol
    epot, phitot = float(np.sum(pg.poten.astype(np.float64))), float(np.sum(pd.poten.astype(np.float64)))
    assert abs(epot - phitot) / abs(phitot) < 1.e-3      # 5.2e-4 in the reference at its own (larger) particle number
    assert abs(epot - (-3. / 5.)) / (3. / 5.) < 3.6e-2


# ------------------------------------------------------------------------------------------------
#  turbulent driving
# ------------------------------------------------------------------------------------------------
def stirring_modes(seed=7, kmax=3):
    """a mode set shaped like init_stir's (forcing.f90:160-200): wave vectors 2 pi (i,j,k), paraboloid amplitudes, random aka/akb"""
    rng = np.random.RandomState(seed)
    modes = []
    for i in range(0, kmax + 1):
        for j in range(0, kmax + 1):
            for k in range(0, kmax + 1):
                k2 = i * i + j * j + k * k
                if 1 <= k2 <= kmax * kmax:
                    for sj in ((1, -1) if j else (1,)):
                        for sk in ((1, -1) if k else (1,)):
                            modes.append((i, sj * j, sk * k))
    mode = 2. * np.pi * np.array(modes, dtype=np.float64)
    kk = np.linalg.norm(mode, axis=1) / (2. * np.pi)
    ampl = 4. * (0. - 1.) / ((kmax - 1.) ** 2) * (kk - 0.5 * (1. + kmax)) ** 2 + 1.
    aka, akb = rng.normal(size=mode.shape), rng.normal(size=mode.shape)
    return mode, ampl, aka, akb


@pytest.mark.parametrize("correct_mean,ind_ts", [(False, False), (True, False), (True, True)])
def test_turbulent_driving(correct_mean, ind_ts):
    part = setups.setup_turb(nx=16, ind_timesteps=ind_ts)
    part.params.driving = 1
    part.alphaind[:, 0] = 1.0
    if ind_ts:
        part.iphase[::3] = -IGAS
        part.gradh[:, 0] = 1.0
        part.fxyzu[:, :3] = 0.123
    mode, ampl, aka, akb = stirring_modes()
    assert len(ampl) > 64                      # more than one shared-memory tile of modes
    args = dict(amplfac=0.7, solweightnorm=1.3, correct_mean_force=correct_mean)
    po, pg = part.copy(), part.copy()
    o = Oracle(po.params)
    o.build_tree(po); o.densityiterate(po); po.params.set_boundaries_to_active = 0; o.set_params(po.params); o.cons2prim(po)
    o.forcing(po, mode, ampl, aka, akb, **args)
    fdrive = po.fxyzu[:, :3].copy()
    so = o.force(po, 1, 0.0)
    g = gpu(pg.params)
    g.set_forcing_modes(mode, ampl, aka, akb, **args)
    sg = g.derivs(pg)
    active = part.iphase > 0
    assert np.sqrt(np.mean(fdrive[active] ** 2)) > 0.1 * np.sqrt(np.mean((po.fxyzu[active, :3] - fdrive[active]) ** 2))   # driving is a visible part of the force
    assert relmax(pg.fxyzu[active, :3], po.fxyzu[active, :3]) < TOL_F
    if ind_ts:
        assert np.array_equal(pg.fxyzu[~active], part.fxyzu[~active])
    # the driving acceleration alone (st_calcAccel)
    g2 = gpu(part.params)
    p2 = part.copy()
    g2.set_forcing_modes(mode, ampl, aka, akb, **args)
    g2.upload(p2); g2.forcing_resident(); g2.download(p2)
    scale = np.sqrt(np.mean(fdrive[active] ** 2))
    assert np.max(np.abs(p2.fxyzu[active, :3] - fdrive[active])) < 1e-12 * scale


def _canonical_tree(rec, irec):
    """node records relabelled in breadth-first order (left before right): the library numbers the two children of a split node from an
    atomic counter, so the labels inside one level vary from run to run while the tree itself does not"""
    order, q = [], [0]
    while q:
        nxt = []
        for d in q:
            order.append(d)
            if irec[d, 0] >= 0:
                nxt += [irec[d, 0], irec[d, 1]]
        q = nxt
    order = np.array(order)
    new = np.full(len(irec), -1, dtype=np.int64)
    new[order] = np.arange(len(order))
    ir = irec[order].astype(np.int64)
    for col in (0, 1, 2):
        m = ir[:, col] >= 0
        ir[m, col] = new[ir[m, col]]
    return rec[order], ir


def test_gravity_pass_is_deterministic_and_tree_bitwise_repeatable():
    """the level-synchronous gravity build accumulates node masses and centres of mass with double atomics, whose
    order varies from run to run, and numbers child nodes from an atomic counter.  Fresh contexts on the same sphere must still give
    the same tree (same splits, same particle order in the leaves, compared after relabelling the nodes breadth-first), the same
    P2P / M2L counts and forces equal to round-off (the pivots move by at most an ulp, which does not flip any particle across a
    split here)."""
    part = setups.setup_random_sphere(n=6000)
    part.alphaind[:, 0] = 0.5
    runs = []
    for _ in range(3):
        pg = part.copy()
        g = gpu(pg.params)
        sc = g.derivs(pg)
        rec, irec, ids = g.gravity_tree(pg.npart)
        rec, irec = _canonical_tree(rec, irec)
        runs.append((pg, sc, rec, irec, ids))
    p0, s0, r0, i0, d0 = runs[0]
    assert len(i0) > 500 and np.all(i0[1:, 2] >= 0)
    fs = np.sqrt(np.mean(p0.fxyzu[:, :3] ** 2))
    for pg, sc, rec, irec, ids in runs[1:]:
        assert np.array_equal(ids, d0)                                              # same particle order in the leaves
        assert np.array_equal(irec, i0)                                             # same topology, slots, counts and levels
        assert sc.npairs_gravity == s0.npairs_gravity and sc.nm2l == s0.nm2l and sc.npairs_force == s0.npairs_force
        assert np.max(np.abs(rec - r0)) <= 1e-13 * np.max(np.abs(r0))               # node moments to round-off of the atomic sums
        assert np.max(np.abs(pg.fxyzu[:, :3] - p0.fxyzu[:, :3])) <= 1e-12 * fs
        assert np.array_equal(pg.xyzh, p0.xyzh)                                     # the SPH side has no atomics in its sums: bitwise
