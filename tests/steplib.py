"""Host-side (numpy) restatement of the reference's leapfrog step and energy diagnostics, driving the CPU oracle
(TEST INFRASTRUCTURE ONLY).  step: src/main/step_leapfrog.f90:95-760 with global timesteps and substep_sph
(substepping.F90:241-264); energies: src/main/energies.f90:205-706."""
import math
import numpy as np

from phantom_b200.params import IGAS, IBOUNDARY

TINY = np.finfo(np.float64).tiny


def oracle_derivs(o, part, icall, dt=0.0):
    """derivs(icall) (deriv.f90:113-192): icall = 1 tree + density + cons2prim + force; icall = 2 cons2prim + force"""
    if icall == 1:
        o.build_tree(part)
        o.densityiterate(part, 1)
        part.params.set_boundaries_to_active = 0
        o.set_params(part.params)
    o.cons2prim(part)
    return o.force(part, icall, dt)


def step_leapfrog(o, part, dt, tolv=1.e-2, maxits=30):
    p = part.params
    nvu = p.maxvxyzu
    hdt = 0.5 * dt
    itype = np.abs(part.iphase.astype(np.int64))
    live = ~(part.xyzh[:, 3] < TINY)
    nb = live & (itype != IBOUNDARY)
    gas = nb & (itype == IGAS)
    pm = np.array([p.massoftype[t] for t in range(8)])[itype]
    v, f, B, dB = part.vxyzu, part.fxyzu, part.Bevol, part.dBevol
    # predictor (:183-235)
    v[nb] += hdt * f[nb]
    if p.mhd:
        B[gas] += hdt * dB[gas]
    # substep_sph
    part.xyzh[live, :3] += dt * v[live, :3]
    # predict_sph (:307-400)
    vtrue, Btrue = v.copy(), B.copy()
    h = part.xyzh[:, 3]
    rho = pm * (p.hfact / np.abs(h)) ** 3
    dhdrho = -h / (3. * rho)
    hnew = h - dt * dhdrho * rho * part.divcurlv[:, 0].astype(np.float64)
    part.xyzh[nb, 3] = hnew[nb]
    vpred = vtrue.copy()
    vpred[nb] = vtrue[nb] + hdt * f[nb]
    Bpred = Btrue.copy()
    if p.mhd:
        Bpred[gas] = Btrue[gas] + hdt * dB[gas]
    if not p.const_av:
        cs = part.eos_vars[:, 1]
        tdecay1 = 0.1 * cs / part.xyzh[:, 3]
        ddenom = 1. / (1. + dt * tdecay1)
        aloc = part.alphaind[:, 1].astype(np.float64)
        a1 = part.alphaind[:, 0].astype(np.float64)
        new = np.where(a1 < aloc, aloc, (a1 + dt * aloc * tdecay1) * ddenom).astype(np.float32)
        part.alphaind[nb, 0] = new[nb]
    part.vxyzu[:], part.Bevol[:] = vpred, Bpred
    sc = oracle_derivs(o, part, 1, dt)
    its, converged, dterr, errmax = 0, False, 1.e29, 0.
    while its < maxits and not converged:
        its += 1
        vnew = vtrue.copy()
        vnew[nb] = vtrue[nb] + hdt * part.fxyzu[nb]
        err = np.sum((vnew[nb, :3] - part.vxyzu[nb, :3]) ** 2, axis=1)
        emax = float(err.max()) if err.size else 0.
        v2mean = float(np.mean(np.sum(vnew[nb, :3] ** 2, axis=1))) if err.size else 0.
        vtrue = vnew
        if p.mhd:
            Btrue[gas] = Btrue[gas] + hdt * part.dBevol[gas]
        errmax = emax / math.sqrt(v2mean) if v2mean > TINY else 0.
        errtol = tolv
        dtf = min(sc.dtcourant, sc.dtforce)
        if dtf > dt and dtf < 1.e29:
            errtol = errtol * (dt / dtf) ** 2
        if its == 1 and errtol > TINY and errmax > np.finfo(np.float64).eps:
            dterr = dt * math.sqrt(errtol / errmax)
        converged = errmax < tolv
        if not converged:
            part.vxyzu[nb] = vtrue[nb]
            vtrue[nb] = vtrue[nb] - hdt * part.fxyzu[nb]
            if p.mhd:
                part.Bevol[nb] = Btrue[nb]
                Btrue[gas] = Btrue[gas] - hdt * part.dBevol[gas]
            sc = oracle_derivs(o, part, 2, dt)
    part.vxyzu[:], part.Bevol[:] = vtrue, Btrue
    return sc, dterr, errmax, its


def energies(part):
    p = part.params
    itype = np.abs(part.iphase.astype(np.int64))
    live = ~(part.xyzh[:, 3] < TINY)
    pm = np.array([p.massoftype[t] for t in range(8)])[itype]
    x, v = part.xyzh[live], part.vxyzu[live]
    m = pm[live]
    rho = m * (p.hfact / np.abs(x[:, 3])) ** 3
    gas = itype[live] == IGAS
    out = {}
    out["ekin"] = 0.5 * float(np.sum(m * np.sum(v[:, :3] ** 2, axis=1)))
    if p.maxvxyzu >= 4:
        out["etherm"] = float(np.sum((m * v[:, 3])[gas]))
    elif p.ieos == 2 and p.gamma > 1.001:
        out["etherm"] = float(np.sum((m * (part.eos_vars[live, 0] / rho) / (p.gamma - 1.))[gas]))
    else:
        out["etherm"] = 0.
    if p.mhd:
        Bx = part.Bevol[live, :3] * rho[:, None]
        out["emag"] = 0.5 * float(np.sum((m * np.sum(Bx * Bx, axis=1) / rho)[gas]))
    else:
        out["emag"] = 0.
    out["epot"] = float(np.sum(part.poten[live].astype(np.float64))) if p.gravity else 0.
    out["etot"] = out["ekin"] + out["etherm"] + out["emag"] + out["epot"]
    mom = np.sum(m[:, None] * v[:, :3], axis=0)
    ang = np.sum(m[:, None] * np.cross(x[:, :3], v[:, :3]), axis=0)
    out["mom"], out["totmom"] = mom, float(np.linalg.norm(mom))
    out["ang"], out["angtot"] = ang, float(np.linalg.norm(ang))
    out["mtot"] = float(np.sum(m))
    out["com"] = np.sum(m[:, None] * x[:, :3], axis=0) / out["mtot"]
    return out


# ---- individual timesteps (-DIND_TIMESTEPS): step_leapfrog.f90:57-80, :157-164, :183-235, :307-400, :470-560 + utils_indtimesteps.f90 ----
def get_dt(dtmax, ibin):
    return dtmax / 2. ** np.asarray(ibin, dtype=np.float64)          # utils_indtimesteps.f90:45-51


def init_step_ind(part, time, dtmax, nbinmax):
    """init_step (step_leapfrog.f90:57-80): at t = 0 every particle starts in the finest bin (boundary particles in bin 0); twas =
    the half step of the particle's own bin.  Returns twas."""
    itype = np.abs(part.iphase.astype(np.int64))
    if time < TINY:
        part.ibin[:] = nbinmax
        part.ibin[itype == IBOUNDARY] = 0
    return time + 0.5 * get_dt(dtmax, part.ibin)


def set_active_particles(part, nbinmax, istepfrac):
    """utils_indtimesteps.f90:114-178: active iff mod(istepfrac, 2**(nbinmax - ibin)) == 0; returns (nactive, ibinnow)"""
    itype = np.abs(part.iphase.astype(np.int64))
    live = ~(part.xyzh[:, 3] < TINY)
    part.ibin[live & (itype == IBOUNDARY)] = 0
    act = (istepfrac % (2 ** (nbinmax - part.ibin.astype(np.int64)))) == 0
    part.iphase[live] = np.where(act[live], itype[live], -itype[live]).astype(np.int8)
    ibinnow, i = nbinmax, 0
    while ibinnow == nbinmax and i < nbinmax:
        if istepfrac % (2 ** (nbinmax - i)) == 0:
            ibinnow = i
        i += 1
    return int(np.sum(act & live)), ibinnow


def step_leapfrog_ind(o, part, twas, t, dtsph, dtmax, nbinmax, ibinnow, istepfrac):
    """one call of step() with individual timesteps; returns (force scalars, new nbinmax).  part.iphase carries the activity flags
    set_active_particles gave it; twas is updated in place."""
    p = part.params
    itype = np.abs(part.iphase.astype(np.int64))
    live = ~(part.xyzh[:, 3] < TINY)
    nb = live & (itype != IBOUNDARY)
    gas = nb & (itype == IGAS)
    active = part.iphase > 0
    pm = np.array([p.massoftype[k] for k in range(8)])[itype]
    v, f, B, dB = part.vxyzu, part.fxyzu, part.Bevol, part.dBevol
    timei = t
    # twas of every bin at the end of this step, for particles that are woken up (:157-164)
    time_now = timei + dtsph
    bins = np.arange(31)
    tdt = get_dt(dtmax, bins)
    ttwas = (np.floor(time_now * (1. / tdt)).astype(np.int64) + 0.5) * tdt
    # predictor (:183-235): everybody goes to its own half step
    part.ibin_old[live & active] = part.ibin[live & active]
    hdti = twas - timei
    v[nb] += hdti[nb, None] * f[nb]
    if p.mhd:
        B[gas] += hdti[gas, None] * dB[gas]
    part.xyzh[live, :3] += dtsph * v[live, :3]                     # substep_sph
    timei += dtsph
    # predict_sph (:307-400)
    vtrue, Btrue = v.copy(), B.copy()
    h = part.xyzh[:, 3]
    rho = pm * (p.hfact / np.abs(h)) ** 3
    dhdrho = -h / (3. * rho)
    hnew = h - dtsph * dhdrho * rho * part.divcurlv[:, 0].astype(np.float64)
    part.xyzh[nb, 3] = hnew[nb]
    hdti = timei - twas
    vpred, Bpred = vtrue.copy(), Btrue.copy()
    vpred[nb] = vtrue[nb] + hdti[nb, None] * f[nb]
    if p.mhd:
        Bpred[gas] = Btrue[gas] + hdti[gas, None] * dB[gas]
    if not p.const_av:
        cs = part.eos_vars[:, 1]
        tdecay1 = 0.1 * cs / part.xyzh[:, 3]
        ddenom = 1. / (1. + dtsph * tdecay1)
        aloc = part.alphaind[:, 1].astype(np.float64)
        a1 = part.alphaind[:, 0].astype(np.float64)
        new = np.where(a1 < aloc, aloc, (a1 + dtsph * aloc * tdecay1) * ddenom).astype(np.float32)
        part.alphaind[nb, 0] = new[nb]
    part.vxyzu[:], part.Bevol[:] = vpred, Bpred
    # derivs(1): the force pass moves the active particles between bins and flags the neighbours to wake (force.F90:1346-1358, :3272-3310)
    o.build_tree(part)
    o.densityiterate(part, 1)
    part.params.set_boundaries_to_active = 0
    o.set_params(part.params)
    o.cons2prim(part)
    sc = o.force(part, 1, dtsph, nbinmax=nbinmax, ibinnow=ibinnow, istepfrac=istepfrac)
    nbinmax_new = int(sc.nbinmaxnew)
    # corrector (:470-560)
    f, dB = part.fxyzu, part.dBevol
    act = nb & active
    part.ibin_wake[act] = 0
    hd = timei - twas
    dti = hd + 0.5 * get_dt(dtmax, part.ibin)
    vtrue[act] += dti[act, None] * f[act]
    if p.mhd:
        ga = act & (itype == IGAS)
        Btrue[ga] += dti[ga, None] * dB[ga]
    twas[act] += dti[act]
    hd = timei - twas                                              # synchronise all particles
    vtrue[nb] += hd[nb, None] * f[nb]
    if p.mhd:
        Btrue[gas] += hd[gas, None] * dB[gas]
    wake = nb & (part.ibin_wake > part.ibin)
    if np.any(wake):
        w = np.minimum(part.ibin_wake[wake].astype(np.int64), nbinmax_new)
        twas[wake] = ttwas[w]
        part.ibin[wake] = w.astype(np.int8)
        part.ibin_wake[wake] = 0
    part.vxyzu[:], part.Bevol[:] = vtrue, Btrue
    return sc, nbinmax_new
