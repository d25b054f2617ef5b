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


# ---- individual timesteps -------------------------------------------------------------------------------------------------------------
class _OracleInd:
    """the evolve.f90 / step_leapfrog.f90 sequence with -DIND_TIMESTEPS on the CPU oracle (tests/steplib.py)"""

    def __init__(self, part):
        self.part = part
        self.o = Oracle(part.params)

    def first_derivs(self):
        o, part = self.o, self.part
        o.build_tree(part); o.densityiterate(part, 1); part.params.set_boundaries_to_active = 0; o.set_params(part.params); o.cons2prim(part)
        return o.force(part, 1, 0.0, nbinmax=0, ibinnow=0, istepfrac=0)

    def init_step(self, time, dtmax, nbinmax):
        self.twas = steplib.init_step_ind(self.part, time, dtmax, nbinmax)

    def substep(self, t, dt, dtmax, nbinmax, istepfrac):
        nactive, ibinnow = steplib.set_active_particles(self.part, nbinmax, istepfrac)
        sc, nbnew = steplib.step_leapfrog_ind(self.o, self.part, self.twas, t, dt, dtmax, nbinmax, ibinnow, istepfrac)
        return nactive, nbnew, sc

    def state(self):
        return self.part


class _GpuInd:
    def __init__(self, part):
        from phantom_b200.api import SphGpu
        self.part = part
        self.g = SphGpu(part.params.copy())
        self.g.upload(part)

    def first_derivs(self):
        self.g.set_timestep_bins(0, 0, 0)
        return self.g.derivs_resident(1)

    def init_step(self, time, dtmax, nbinmax):
        self.g.init_step_resident(time, dtmax, nbinmax)

    def substep(self, t, dt, dtmax, nbinmax, istepfrac):
        nactive, nalive = self.g.set_active_particles_resident(nbinmax, istepfrac)
        out = self.g.step_ind_resident(t, dt, dtmax)
        return nactive, int(out.scalars.nbinmaxnew), out.scalars

    def state(self):
        self.g.download(self.part)
        return self.part


def _evolve_ind(b, dtmax, nsub):
    """evolve.f90:196-203 + evolve_utils.F90:65-87: istepfrac, time and nbinmax bookkeeping around step()"""
    sc = b.first_derivs()
    nbinmax = int(sc.nbinmaxnew)
    b.init_step(0., dtmax, nbinmax)
    istepfrac, t, log = 0, 0., []
    for _ in range(nsub):
        dt = dtmax / 2 ** nbinmax
        istepfrac += 1
        nactive, nbnew, sc = b.substep(t, dt, dtmax, nbinmax, istepfrac)
        t = istepfrac / 2. ** nbinmax * dtmax
        log.append((istepfrac, nbinmax, nactive, nbnew, int(sc.npairs_force)))
        if nbnew != nbinmax:                                           # change_nbinmax (utils_indtimesteps.f90:186-222)
            if nbnew < nbinmax:
                assert istepfrac % 2 ** (nbinmax - nbnew) == 0
                istepfrac //= 2 ** (nbinmax - nbnew)
            else:
                istepfrac *= 2 ** (nbnew - nbinmax)
            nbinmax = nbnew
        if istepfrac == 2 ** nbinmax:
            break
    return log, t


def test_individual_timestep_stepping_matches_oracle():
    """turbulent box with -DIND_TIMESTEPS: set_active_particles + step() per smallest timestep until the bins synchronise (or 24
    substeps), on the device (sphgpu_set_active_particles_resident, sphgpu_step_ind_resident) and on the oracle: the same active counts,
    bin populations and pair counts every substep, positions / velocities / h to 1e-10"""
    part = setups.setup_turb(nx=14, ind_timesteps=True)
    part.alphaind[:, 0] = 1.0
    rng = setups.Ran2(-8642)
    part.xyzh[:, :3] += 0.15 / 14 * (rng.draw(3 * part.npart).reshape(-1, 3) - 0.5)      # break the lattice: a spread of timesteps
    dtmax = 0.02
    part.params.dtmax = dtmax
    po, pg = part.copy(), part.copy()
    bo, bg = _OracleInd(po), _GpuInd(pg)
    log_o, t_o = _evolve_ind(bo, dtmax, 24)
    log_g, t_g = _evolve_ind(bg, dtmax, 24)
    assert log_g == log_o, (log_g[:6], log_o[:6])
    assert t_g == t_o and len(log_o) >= 4
    assert len({l[2] for l in log_o}) > 1                                # the number of active particles does change between substeps
    so, sg = bo.state(), bg.state()
    assert np.array_equal(sg.ibin, so.ibin) and np.array_equal(sg.iphase, so.iphase)
    dx = sg.xyzh[:, :3] - so.xyzh[:, :3]
    assert np.max(np.abs(dx - np.round(dx))) < 1e-10
    assert np.max(np.abs(sg.xyzh[:, 3] - so.xyzh[:, 3]) / so.xyzh[:, 3]) < 1e-9
    vs = np.sqrt(np.mean(so.vxyzu[:, :3] ** 2))
    assert np.max(np.abs(sg.vxyzu[:, :3] - so.vxyzu[:, :3])) < 1e-10 * vs
