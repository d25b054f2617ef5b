"""Synthetic initial conditions restating the reference's IC generators.

The reference builds its ICs with `phantomsetup` (src/setup/); it cannot be
compiled here, so the generators used by BASELINE.json's configs and by the
reference's own unit tests are restated in numpy:

  set_unifdis  cubic / closepacked / random   src/setup/set_unifdis.f90:181-233, :354-518, :520-560
  ran2 / get_random (L'Ecuyer 1988)            src/main/random.f90:38-114
  test_derivs velocity / energy fields          src/tests/test_derivs.f90:1216-1240, :1604-1613
  Sod shock tube                                src/setup/set_shock.f90:35-140, setup_shock.f90:497-499
  MHD blast / Orszag-Tang                       src/setup/setup_mhdblast.f90:70-123, setup_orstang.f90:88-125

Everything returns plain numpy arrays in the Fortran layout the hot-path
interfaces take: xyzh(4,N) is a C-contiguous (N,4) array, etc.
"""
import math
import numpy as np

from .params import default_params, IGAS, IBOUNDARY, IDUST, KERNEL_CUBIC, KERNEL_QUINTIC

_M1, _M2 = 2147483563, 2147483399
_A1, _A2 = 40014, 40692
EPS = np.finfo(np.float64).eps


class Ran2:
    """ran2(iseed) of src/main/random.f90: two int32 LCG states; a negative seed resets the second."""

    def __init__(self, iseed=-43587):
        self.s1 = int(iseed)
        self.s2 = 123456789

    def draw(self, n):
        """n successive deviates, vectorised: s_k = s_0 * a^k mod m (Schrage's trick in the
        reference only avoids int32 overflow; the recurrence is a plain modular product)."""
        if self.s1 < 0:
            self.s2 = 123456789
        out = np.empty(n, dtype=np.float64)
        B = 1 << 15
        p1 = np.empty(B, dtype=np.uint64)
        p2 = np.empty(B, dtype=np.uint64)
        a1 = a2 = 1
        for k in range(B):
            a1 = (a1 * _A1) % _M1
            a2 = (a2 * _A2) % _M2
            p1[k] = a1
            p2[k] = a2
        s1 = self.s1 % _M1
        s2 = self.s2 % _M2
        done = 0
        while done < n:
            m = min(B, n - done)
            v1 = (np.uint64(s1) * p1[:m]) % np.uint64(_M1)
            v2 = (np.uint64(s2) * p2[:m]) % np.uint64(_M2)
            z = v1.astype(np.int64) - v2.astype(np.int64)
            z[z < 1] += 2147483562
            out[done:done + m] = z / 2147483563.0
            s1, s2 = int(v1[-1]), int(v2[-1])
            done += m
        self.s1, self.s2 = s1, s2
        return out


def _nint(x):
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def unifdis_cubic(xmin, xmax, ymin, ymax, zmin, zmax, delta, hfact):
    """set_unifdis 'cubic' (set_unifdis.f90:181-233): ids run x fastest, then y, then z."""
    dxb, dyb, dzb = xmax - xmin, ymax - ymin, zmax - zmin
    nx, ny, nz = _nint(dxb / delta), _nint(dyb / delta), _nint(dzb / delta)
    dx, dy, dz = dxb / nx, dyb / ny, dzb / nz
    x = xmin + (np.arange(1, nx + 1) - 0.5) * dx
    y = ymin + (np.arange(1, ny + 1) - 0.5) * dy
    z = zmin + (np.arange(1, nz + 1) - 0.5) * dz
    xyzh = np.empty((nz, ny, nx, 4))
    xyzh[..., 0] = x[None, None, :]
    xyzh[..., 1] = y[None, :, None]
    xyzh[..., 2] = z[:, None, None]
    xyzh[..., 3] = hfact * dx
    return np.ascontiguousarray(xyzh.reshape(-1, 4))


def closepacked_ny_nz(delta, ymin, ymax, zmin, zmax, periodic=True):
    dyb, dzb = ymax - ymin, zmax - zmin
    deltay = delta * math.sqrt(3. / 4.)
    deltaz = delta * math.sqrt(6.) / 3.
    ny = int((1. - EPS) * dyb / deltay) + 1
    nz = int((1. - EPS) * dzb / deltaz) + 1
    if periodic:
        ny = 2 * (ny // 2)
        nz = 3 * (nz // 3)
    return ny, nz


def unifdis_closepacked(xmin, xmax, ymin, ymax, zmin, zmax, delta, hfact, periodic=True, npy=0, npz=0):
    """set_unifdis 'closepacked' (set_unifdis.f90:354-518), ABC stacking."""
    dxb, dyb, dzb = xmax - xmin, ymax - ymin, zmax - zmin
    deltax = delta
    deltay = delta * math.sqrt(3. / 4.)
    deltaz = delta * math.sqrt(6.) / 3.
    delx = 0.5 * delta
    dely = 1. / 3. * deltay
    nx = int((1. - EPS) * dxb / deltax) + 1
    ny = int((1. - EPS) * dyb / deltay) + 1
    nz = int((1. - EPS) * dzb / deltaz) + 1
    if npy > 0:
        ny = npy
    if npz > 0:
        nz = npz
    if periodic:
        ny = 2 * (ny // 2)
        nz = 3 * (nz // 3)
        if npy <= 0:
            deltax = dxb / float(nx)
        deltay = dyb / float(ny)
        deltaz = dzb / float(nz)
        dely = 1. / 3. * deltay
    k = np.arange(1, nx + 1)[None, None, :]
    l = np.arange(1, ny + 1)[None, :, None]
    m = np.arange(1, nz + 1)[:, None, None]
    jy = l % 2
    jz = m % 3
    xstart = np.full((nz, ny, 1), xmin + 0.5 * delx)
    ystart = np.full((nz, ny, 1), ymin + 0.5 * dely)
    zstart = zmin + 0.5 * deltaz
    jyb = np.broadcast_to(jy, (nz, ny, 1))
    jzb = np.broadcast_to(jz, (nz, ny, 1))
    third = (jzb == 0)
    second = (jzb == 2)
    first = (jzb == 1)
    ystart = ystart + np.where(third, 2. * dely, 0.) + np.where(second, dely, 0.)
    xstart = xstart + np.where(third & (jyb == 0), delx, 0.) + np.where(second & (jyb == 1), delx, 0.) \
        + np.where(first & (jyb == 0), delx, 0.)
    xyzh = np.empty((nz, ny, nx, 4))
    xyzh[..., 0] = xstart + (k - 1.0) * deltax
    xyzh[..., 1] = ystart + (l - 1.0) * deltay
    xyzh[..., 2] = zstart + (m - 1.0) * deltaz
    xyzh[..., 3] = hfact * deltax
    return np.ascontiguousarray(xyzh.reshape(-1, 4))


def unifdis_random(xmin, xmax, ymin, ymax, zmin, zmax, delta, hfact, iseed=-43587, npnew=None, rmax=None):
    """set_unifdis 'random' (set_unifdis.f90:520-560); rmax trims to a sphere (in_range on rr2)."""
    dxb, dyb, dzb = xmax - xmin, ymax - ymin, zmax - zmin
    if npnew is None:
        npnew = _nint(dxb / delta) * _nint(dyb / delta) * _nint(dzb / delta)
    rng = Ran2(iseed)
    got = []
    ngot = 0
    while ngot < npnew:
        ntry = max(1024, int((npnew - ngot) * (2.0 if rmax else 1.0)))
        r = rng.draw(3 * ntry).reshape(ntry, 3)
        pts = np.empty((ntry, 3))
        pts[:, 0] = xmin + r[:, 0] * dxb
        pts[:, 1] = ymin + r[:, 1] * dyb
        pts[:, 2] = zmin + r[:, 2] * dzb
        if rmax is not None:
            rr2 = pts[:, 0] * pts[:, 0] + pts[:, 1] * pts[:, 1] + pts[:, 2] * pts[:, 2]
            pts = pts[rr2 < rmax * rmax]
        got.append(pts)
        ngot += len(pts)
    pts = np.concatenate(got)[:npnew]
    xyzh = np.empty((npnew, 4))
    xyzh[:, :3] = pts
    xyzh[:, 3] = hfact * delta
    return xyzh


# ---------------------------------------------------------------------------------------------
#  a bundle of particle arrays in the layouts the hot-path interfaces take
# ---------------------------------------------------------------------------------------------
class Particles:
    """Host mirror of the slice of part.F90 (src/main/part.F90:47-416) that the hot path touches."""

    def __init__(self, params, xyzh, iphase=None):
        n = len(xyzh)
        p = params
        nvu = p.maxvxyzu
        self.params = p
        self.npart = n
        self.xyzh = np.ascontiguousarray(xyzh, dtype=np.float64)
        self.vxyzu = np.zeros((n, nvu))
        self.fxyzu = np.zeros((n, nvu))
        self.fext = np.zeros((n, 3))
        self.Bevol = np.zeros((n, 4))
        self.dBevol = np.zeros((n, 4))
        self.iphase = np.full(n, IGAS, dtype=np.int8) if iphase is None else np.ascontiguousarray(iphase, dtype=np.int8)
        self.divcurlv = np.zeros((n, 1), dtype=np.float32)
        self.divcurlB = np.zeros((n, 4), dtype=np.float32)
        self.alphaind = np.zeros((n, 3), dtype=np.float32)
        self.gradh = np.zeros((n, p.ngradh), dtype=np.float32)
        self.dvdx = np.zeros((n, 9), dtype=np.float32)
        self.eos_vars = np.zeros((n, 7))
        self.Bxyz = np.zeros((n, 3))
        self.poten = np.zeros(n, dtype=np.float32)
        self.divBsymm = np.zeros(n, dtype=np.float32)
        self.dustfrac = np.zeros(n)
        self.tstop = np.zeros(n)
        self.ibin = np.zeros(n, dtype=np.int8)
        self.ibin_old = np.zeros(n, dtype=np.int8)
        self.ibin_wake = np.zeros(n, dtype=np.int8)

    def copy(self):
        q = Particles.__new__(Particles)
        q.params = self.params.copy()
        q.npart = self.npart
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                setattr(q, k, v.copy())
        return q


# ---------------------------------------------------------------------------------------------
#  the reference's test_derivs configuration (src/tests/test_derivs.f90:128-163)
# ---------------------------------------------------------------------------------------------
def test_derivs_fields(xyzh, p):
    """vx,vy,vz (test_derivs.f90:1216-1240) and utherm (:1604-1613)."""
    dxb, dyb, dzb = p.xmax - p.xmin, p.ymax - p.ymin, p.zmax - p.zmin
    x, y, z = xyzh[:, 0], xyzh[:, 1], xyzh[:, 2]
    pi = math.pi
    vx = 0.5 / pi * dxb * np.sin(2. * pi * (x - p.xmin) / dxb)
    vy = 0.5 / pi * dxb * np.sin(2. * pi * (x - p.xmin) / dxb) - 0.5 / pi * dzb * np.sin(2. * pi * (z - p.zmin) / dzb)
    vz = 0.05 / pi * dyb * np.cos(4. * pi * (y - p.ymin) / dyb)
    u = 0.5 / pi * (3. + np.sin(2. * pi * (x - p.xmin) / dxb) + np.cos(2. * pi * (y - p.ymin) / dyb)
                    + np.sin(2. * pi * (z - p.zmin) / dzb))
    return vx, vy, vz, u


def setup_test_derivs(nx=100, rhozero=5.0, tolh=1.e-5, lattice="cubic", isothermal=False, mhd=False, iseed=-43587,
                      dissipation=True, **kw):
    """test_derivs.f90:128-163; dissipation=False restates reset_dissipation_to_zero (:989-1003)."""
    if not dissipation:
        kw = {**dict(alpha=0., alphau=0., alphaB=0., beta=0.), **kw}
    p = default_params(**{**dict(tolh=tolh, isothermal=int(isothermal), mhd=int(mhd), ieos=1 if isothermal else 2), **kw})
    if isothermal:
        p.polyk, p.gamma = 3.0, 1.0
    dxb = p.xmax - p.xmin
    psep = dxb / nx
    if lattice == "cubic":
        xyzh = unifdis_cubic(p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax, psep, p.hfact)
    elif lattice == "closepacked":
        xyzh = unifdis_closepacked(p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax, psep, p.hfact, periodic=bool(p.periodic))
    else:
        xyzh = unifdis_random(p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax, psep, p.hfact, iseed=iseed)
    n = len(xyzh)
    totmass = rhozero * dxb * (p.ymax - p.ymin) * (p.zmax - p.zmin)
    p.massoftype[IGAS] = totmass / n
    part = Particles(p, xyzh)
    vx, vy, vz, u = test_derivs_fields(xyzh, p)
    part.vxyzu[:, 0], part.vxyzu[:, 1], part.vxyzu[:, 2] = vx, vy, vz
    if not isothermal:
        part.vxyzu[:, 3] = u
    if mhd:
        # a smooth divergence-free test field, stored as B/rho
        x, y, z = xyzh[:, 0], xyzh[:, 1], xyzh[:, 2]
        pi = math.pi
        Bx = 0.5 * np.sin(2 * pi * (y - p.ymin)) + 0.3
        By = 0.4 * np.sin(2 * pi * (z - p.zmin))
        Bz = 0.6 * np.cos(2 * pi * (x - p.xmin)) - 0.2
        part.Bevol[:, 0], part.Bevol[:, 1], part.Bevol[:, 2] = Bx / rhozero, By / rhozero, Bz / rhozero
        part.Bevol[:, 3] = 0.05 * np.sin(2 * pi * (x - p.xmin)) * np.cos(2 * pi * (y - p.ymin))
    hzero = p.hfact * (p.massoftype[IGAS] / rhozero) ** (1. / 3.)
    return part, hzero


def unifdis_cubic_shell(xmin, xmax, ymin, ymax, zmin, zmax, delta, hfact, rmin=None, rmax=None):
    """set_unifdis 'cubic' with the rmin/rmax mask (set_unifdis.f90:124-133, :181-233, in_range :615-620: both ends inclusive):
    same loop order (x fastest), only the index range that can pass the mask is scanned."""
    dxb, dyb, dzb = xmax - xmin, ymax - ymin, zmax - zmin
    nx, ny, nz = _nint(dxb / delta), _nint(dyb / delta), _nint(dzb / delta)
    dx, dy, dz = dxb / nx, dyb / ny, dzb / nz
    rmin2 = rmin * rmin if rmin is not None else 0.
    rmax2 = rmax * rmax if rmax is not None else np.finfo(np.float64).max

    def axis(lo, d, n):
        c = lo + (np.arange(1, n + 1) - 0.5) * d
        if rmax is not None:
            c = c[np.abs(c) <= rmax * (1. + 1e-12) + 1e-300]
        return c
    x, y, z = axis(xmin, dx, nx), axis(ymin, dy, ny), axis(zmin, dz, nz)
    out = []
    for zi in z:                                   # slabs keep the memory bounded (the blob lattice is 500^3)
        rcyl2 = x[None, :] * x[None, :] + y[:, None] * y[:, None]
        rr2 = rcyl2 + zi * zi
        keep = (rmin2 <= rr2) & (rr2 <= rmax2)
        jj, ii = np.nonzero(keep)                  # row-major: y slow, x fast = the reference's loop order
        if len(ii):
            slab = np.empty((len(ii), 4))
            slab[:, 0], slab[:, 1], slab[:, 2], slab[:, 3] = x[ii], y[jj], zi, hfact * dx
            out.append(slab)
    return np.concatenate(out) if out else np.empty((0, 4))


def setup_density_contrast(rhozero=5.0, tolh=1.e-5, isothermal=False):
    """test_derivs.f90:651-713 (derivscontrast): a blob of 1000 x the density of the surrounding 50^3 medium, blob lattice spacing
    psep/10, no dissipation.  Returns the particles, the number of test particles (r <= rblob - 2 hfact psep) and hblob.
    Known answers held by the reference for the cubic kernel (:698-707): mean neighbours 57.466651861721814, max 988, total 37263216."""
    p = default_params(tolh=tolh, isothermal=int(isothermal), ieos=1 if isothermal else 2, alpha=0., alphau=0., alphaB=0., beta=0.)
    if isothermal:
        p.polyk, p.gamma = 3.0, 1.0
    dxb, dyb, dzb = p.xmax - p.xmin, p.ymax - p.ymin, p.zmax - p.zmin
    psep = dxb / 50.
    rblob = 0.1
    rhoblob = 1000. * rhozero
    psepblob = psep * (rhozero / rhoblob) ** (1. / 3.)
    rtest = rblob - 2. * p.hfact * psep
    box = (p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax)
    inner = unifdis_cubic_shell(*box, psepblob, p.hfact, rmax=rtest)
    shell = unifdis_cubic_shell(*box, psepblob, p.hfact, rmin=rtest, rmax=rblob)
    medium = unifdis_cubic_shell(*box, psep, p.hfact, rmin=rblob)
    xyzh = np.concatenate([inner, shell, medium])
    nparttest, npartblob, npart = len(inner), len(inner) + len(shell), len(xyzh)
    totvol = dxb * dyb * dzb - 4. / 3. * math.pi * rblob ** 3
    p.massoftype[IGAS] = rhozero * totvol / (npart - npartblob)
    part = Particles(p, xyzh)
    vx, vy, vz, u = test_derivs_fields(xyzh, p)
    part.vxyzu[:, 0], part.vxyzu[:, 1], part.vxyzu[:, 2] = vx, vy, vz
    if not isothermal:
        part.vxyzu[:, 3] = u
    hblob = p.hfact * (p.massoftype[IGAS] / rhoblob) ** (1. / 3.)
    return part, nparttest, hblob


# ---------------------------------------------------------------------------------------------
#  BASELINE.json configs
# ---------------------------------------------------------------------------------------------
def _solenoidal_field(xyzh, box, seed, mach, cs, kmin=1, kmax=3):
    """Seeded solenoidal Gaussian velocity field, k in [kmin,kmax]*2pi/L, rms Mach `mach`
    (substitute for the OU-driven state; data/forcing/forcing.dat is not shipped)."""
    rng = np.random.RandomState(seed)
    xmin, L = box
    x = (xyzh[:, :3] - xmin) / L
    v = np.zeros((len(xyzh), 3))
    modes = []
    for kx in range(-kmax, kmax + 1):
        for ky in range(-kmax, kmax + 1):
            for kz in range(0, kmax + 1):
                k2 = kx * kx + ky * ky + kz * kz
                if k2 < kmin * kmin or k2 > kmax * kmax:
                    continue
                kv = np.array([kx, ky, kz], dtype=float)
                a = rng.normal(size=3) * k2 ** (-1.0)
                b = rng.normal(size=3) * k2 ** (-1.0)
                a -= kv * (a @ kv) / k2          # project out the compressive part
                b -= kv * (b @ kv) / k2
                modes.append((kv, a, b))

    def fill(lo, hi):                            # the same operations per particle, whatever the chunking: results do not depend on it
        xc, vc = x[lo:hi], v[lo:hi]
        for kv, a, b in modes:
            ph = 2. * math.pi * (xc @ kv)
            vc += np.outer(np.cos(ph), a) + np.outer(np.sin(ph), b)
    n = len(xyzh)
    chunk = 1 << 17
    if n <= chunk:
        fill(0, n)
    else:                                        # numpy releases the GIL inside the ufuncs: threads over particle chunks
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 8)) as ex:
            list(ex.map(lambda lo: fill(lo, min(lo + chunk, n)), range(0, n, chunk)))
    vrms = math.sqrt(np.mean(np.sum(v * v, axis=1)))
    return v * (mach * cs / vrms)


def setup_turb(nx=128, mach=5.0, seed=1234, ind_timesteps=False, positions="lattice"):
    """C2: SETUP=turb -- isothermal periodic box [0,1]^3, cubic lattice nx^3, cs=1, rho0=1
    (setup_turb.f90:101-176).  positions="random": the same number of particles at uniformly random positions (set_unifdis 'random'),
    a disordered state for benchmarks (target groups no longer align with the lattice)."""
    p = default_params(isothermal=1, ieos=1, polyk=1.0, gamma=1.0, ind_timesteps=int(ind_timesteps),
                       xmin=0., xmax=1., ymin=0., ymax=1., zmin=0., zmax=1., dtmax=0.025 if ind_timesteps else 1e29)
    if positions == "random":
        xyzh = unifdis_random(0., 1., 0., 1., 0., 1., 1.0 / nx, p.hfact)
    else:
        xyzh = unifdis_cubic(0., 1., 0., 1., 0., 1., 1.0 / nx, p.hfact)
    p.massoftype[IGAS] = 1.0 / len(xyzh)
    part = Particles(p, xyzh)
    part.vxyzu[:, :3] = _solenoidal_field(xyzh, (0.0, 1.0), seed, mach, 1.0)
    return part


def block_dims(world):
    """(bx,by,bz) blocks of the weak-scaling box: 1 -> (1,1,1), 2 -> (2,1,1), 4 -> (2,2,1), 8 -> (2,2,2)"""
    dims = [1, 1, 1]
    k = 0
    n = world
    while n > 1:
        dims[k % 3] *= 2
        n //= 2
        k += 1
    return tuple(dims)


def setup_turb_block(nx, world, rank, mach=5.0, seed=1234):
    """Weak-scaling version of C2 for `world` GPUs: a periodic box of bx x by x bz unit blocks, each block an nx^3 lattice
    (same particle spacing, mass and h as setup_turb); returns the particles of block `rank` and all block boxes."""
    bx, by, bz = block_dims(world)
    p = default_params(isothermal=1, ieos=1, polyk=1.0, gamma=1.0, xmin=0., xmax=float(bx), ymin=0., ymax=float(by), zmin=0., zmax=float(bz))
    ix, iy, iz = rank % bx, (rank // bx) % by, rank // (bx * by)
    xyzh = unifdis_cubic(float(ix), ix + 1., float(iy), iy + 1., float(iz), iz + 1., 1.0 / nx, p.hfact)
    p.massoftype[IGAS] = 1.0 / (nx ** 3)
    part = Particles(p, xyzh)
    scaled = xyzh.copy()
    scaled[:, 0] /= bx
    scaled[:, 1] /= by
    scaled[:, 2] /= bz
    part.vxyzu[:, :3] = _solenoidal_field(scaled, (0.0, 1.0), seed, mach, 1.0)
    boxes = np.array([[r % bx, (r // bx) % by, r // (bx * by), r % bx + 1., (r // bx) % by + 1., r // (bx * by) + 1.] for r in range(world)], dtype=float)
    return part, boxes


def setup_shock(nx=256, gamma=5. / 3., width=1):
    """C1: SETUP=shock -- 3D Sod tube, quintic kernel, adiabatic, closepacked, periodic in y,z
    (setup_shock.f90:497-499 states; set_shock.f90:35-140; adjust_shock_boundaries :180-209)."""
    rhoL, rhoR, prL, prR = 1.0, 0.125, 1.0, 0.1
    radkern, hfact = 3.0, 1.0
    xleft, xright, xshock = -0.5, 0.5, 0.0
    dxleft = (xshock - xleft) / nx          # nx particles across the left half (nx=256 -> 1/512)
    dxright = dxleft * (rhoL / rhoR) ** (1. / 3.)
    fac = -6. * width * (int(1.99 * radkern / 6.) + 1) * max(dxleft, dxright)      # width > 1: a thicker tube than adjust_shock_boundaries makes (measurements)
    ymin, zmin = fac * math.sqrt(0.75), fac * math.sqrt(6.) / 3.
    ymax, zmax = -ymin, -zmin
    p = default_params(kernel=KERNEL_QUINTIC, hfact=hfact, gamma=gamma, ieos=2, isothermal=0,
                       xmin=xleft - 1000. * dxleft, xmax=xright + 1000. * dxright, ymin=ymin, ymax=ymax, zmin=zmin, zmax=zmax)
    left = unifdis_closepacked(xleft, xshock, ymin, ymax, zmin, zmax, dxleft, hfact, periodic=True)
    massgas = (xshock - xleft) * (ymax - ymin) * (zmax - zmin) * rhoL / len(left)
    ny, nz = closepacked_ny_nz(dxright, ymin, ymax, zmin, zmax)
    totmassR = (xright - xshock) * (ymax - ymin) * (zmax - zmin) * rhoR
    dxright = (xright - xshock) / ((totmassR / massgas) / (ny * nz))
    right = unifdis_closepacked(xshock, xright, ymin, ymax, zmin, zmax, dxright, hfact, periodic=True, npy=ny, npz=nz)
    xyzh = np.concatenate([left, right])
    n = len(xyzh)
    p.massoftype[IGAS] = massgas
    p.massoftype[IBOUNDARY] = massgas
    isleft = xyzh[:, 0] < xshock
    rho = np.where(isleft, rhoL, rhoR)
    pr = np.where(isleft, prL, prR)
    xyzh[:, 3] = hfact * (massgas / rho) ** (1. / 3.)
    # boundary particles within nbpts spacings of the x ends (setup_shock.f90:233-250)
    nbpts = _nint(2.01 * radkern * hfact)
    iphase = np.full(n, IGAS, dtype=np.int8)
    iphase[(xyzh[:, 0] < xleft + nbpts * dxleft) | (xyzh[:, 0] > xright - nbpts * dxright)] = IBOUNDARY
    part = Particles(p, xyzh, iphase)
    part.vxyzu[:, 3] = pr / ((gamma - 1.) * rho)
    return part


def setup_mhdblast(nx=64):
    """C3(i): SETUP=mhdblast (setup_mhdblast.f90:70-123): closepacked in [-0.5,0.5]^3, gamma=1.4,
    rho=1, B=(10/sqrt2, 0, 10/sqrt2), P=100 inside r<0.125 else 1."""
    gamma = 1.4
    p = default_params(mhd=1, gamma=gamma, ieos=2)
    xyzh = unifdis_closepacked(p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax, 1.0 / nx, p.hfact, periodic=True)
    n = len(xyzh)
    p.massoftype[IGAS] = 1.0 / n
    part = Particles(p, xyzh)
    r = np.sqrt(np.sum(xyzh[:, :3] ** 2, axis=1))
    pr = np.where(r < 0.125, 100.0, 1.0)
    part.vxyzu[:, 3] = pr / ((gamma - 1.) * 1.0)
    B0 = 10. / math.sqrt(2.)
    part.Bevol[:, 0] = B0 / 1.0
    part.Bevol[:, 2] = B0 / 1.0
    return part


def setup_orstang(nx=64, nlayers=12):
    """C3(ii): SETUP=orstang (setup_orstang.f90:88-125): thin slab, v=(-sin2piy, sin2pix, 0),
    B=B0(-sin2piy, sin4pix, 0), B0=1/sqrt(4pi), beta0=10/3, M0=1, gamma=5/3."""
    gamma = 5. / 3.
    betazero, machzero = 10. / 3., 1.0
    const = 4. * math.pi
    bzero = 1.0 / math.sqrt(const)
    przero = 0.5 * bzero ** 2 * betazero
    rhozero = gamma * przero * machzero
    deltax = 1.0 / nx
    dz = 4. * math.sqrt(6.) / nx * (nlayers / 12.0)
    p = default_params(mhd=1, gamma=gamma, ieos=2, xmin=-0.5, xmax=0.5, ymin=-0.5, ymax=0.5, zmin=-dz, zmax=dz)
    xyzh = unifdis_closepacked(p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax, deltax, p.hfact, periodic=True)
    n = len(xyzh)
    totmass = rhozero * 1.0 * 1.0 * (2 * dz)
    p.massoftype[IGAS] = totmass / n
    part = Particles(p, xyzh)
    x, y = xyzh[:, 0], xyzh[:, 1]
    part.xyzh[:, 3] = p.hfact * (p.massoftype[IGAS] / rhozero) ** (1. / 3.)
    part.vxyzu[:, 0] = -np.sin(2. * math.pi * (y - p.ymin))
    part.vxyzu[:, 1] = np.sin(2. * math.pi * (x - p.xmin))
    part.vxyzu[:, 3] = przero / ((gamma - 1.) * rhozero)
    part.Bevol[:, 0] = -bzero * np.sin(2. * math.pi * (y - p.ymin)) / rhozero
    part.Bevol[:, 1] = bzero * np.sin(4. * math.pi * (x - p.xmin)) / rhozero
    return part


def setup_random_sphere(n=1000, iseed=-43587, gravity=True, u=0.05):
    """C5 / test_gravity.f90:300-330: uniform random sphere R=1, M=1."""
    p = default_params(periodic=0, gravity=int(gravity), ieos=2, xmin=-1., xmax=1., ymin=-1., ymax=1., zmin=-1., zmax=1.)
    psep = 2.0 / (n * 6.0 / math.pi) ** (1. / 3.)
    xyzh = unifdis_random(-1., 1., -1., 1., -1., 1., psep, p.hfact, iseed=iseed, npnew=n, rmax=1.0)
    p.massoftype[IGAS] = 1.0 / n
    rho0 = 1.0 / (4. / 3. * math.pi)
    xyzh[:, 3] = p.hfact * (p.massoftype[IGAS] / rho0) ** (1. / 3.)
    part = Particles(p, xyzh)
    part.vxyzu[:, 3] = u
    return part


def setup_dustydisc(ngas=8000, ndust=2000, iseed=-43587, r_in=1.0, r_out=150.0, pindex=1.0, qindex=0.25, hor_in=0.05,
                    alpha_ss=0.005, disc_mass=0.05, dust_to_gas=0.01, ind_timesteps=True):
    """C4: SETUP=dustydisc -- gas + dust (two-fluid, dust as particles) accretion disc around a central star that is handled
    outside the hot path (its force enters through fext).  Reference defaults: R_in=1, R_out=150, Sigma ~ R^-pindex,
    c_s ~ R^-qindex, H/R = 0.05 at R_in, alphaSS = 0.005 (setup_disc.f90:427-443); locally isothermal ieos = 3 (:709);
    disc_viscosity because alphaSS > 0 (set_disc.f90:488-489) with alpha_AV from alphaSS (setup_disc.f90:1166);
    dust_method = 2, dust-to-gas 0.01, 1 cm grains of 3 g/cm^3, Epstein/Stokes drag idrag = 1 (set_dust_options.f90:96-103);
    quintic kernel (DUST default, build/Makefile:213-218); IND_TIMESTEPS (Makefile_setups:541-549).
    Positions are Monte-Carlo sampled with ran2 (seed -43587): R from Sigma(R) R dR, z Gaussian with H(R), uniform azimuth.
    Code units: au, solar mass, G = 1."""
    udist, umass = 1.496e13, 1.989e33
    unit_density = umass / udist ** 3
    p = default_params(kernel=KERNEL_QUINTIC, hfact=1.0, periodic=0, isothermal=1, ieos=3, dust=1, idrag=1, disc_viscosity=1,
                       ind_timesteps=int(ind_timesteps), qfacdisc=qindex, gamma=1.0,
                       xmin=-r_out, xmax=r_out, ymin=-r_out, ymax=r_out, zmin=-r_out, zmax=r_out)
    cs_in = hor_in * math.sqrt(1.0 / r_in)                 # H = c_s / Omega with M_star = 1
    p.polyk = cs_in ** 2 * r_in ** (2. * qindex)           # c_s^2 = polyk R^-2q (eos.f90:226-234)
    # alpha_AV from alphaSS: alpha_SS ~ alpha_AV/10 <h>/H (setup_disc.f90:1166, Lodato & Price 2010); <h>/H ~ 0.5 at this resolution
    p.alpha = min(max(10. * alpha_ss / 0.5, 0.01), 1.0)
    p.const_av = 0
    p.grainsize = 1.0 / udist
    p.graindens = 3.0 / unit_density
    mass_mol_gas = 2. * 1.67262158e-24 / umass             # dust.f90:96-99 (init_drag)
    cross_section_gas = 2.367e-15 / udist ** 2
    p.seff = math.pi / math.sqrt(2.) * 5. / 64. * mass_mol_gas / cross_section_gas
    p.massoftype[IGAS] = disc_mass / ngas
    p.massoftype[IDUST] = disc_mass * dust_to_gas / ndust
    rng = Ran2(iseed)

    def sample(n, hfac):
        u = rng.draw(3 * n).reshape(n, 3)
        if abs(pindex - 2.) < 1e-12:
            R = r_in * (r_out / r_in) ** u[:, 0]
        else:                                              # P(R) ~ R^(1-pindex)
            e = 2. - pindex
            R = (r_in ** e + u[:, 0] * (r_out ** e - r_in ** e)) ** (1. / e)
        phi = 2. * math.pi * u[:, 1]
        H = hfac * hor_in * r_in * (R / r_in) ** (1.5 - qindex)
        g = rng.draw(2 * n).reshape(n, 2)
        z = H * np.sqrt(-2. * np.log(np.maximum(g[:, 0], 1e-300))) * np.cos(2. * math.pi * g[:, 1])
        return R, phi, z, H

    Rg, phig, zg, Hg = sample(ngas, 1.0)
    Rd, phid, zd, Hd = sample(ndust, 0.5)                  # settled dust layer
    R = np.concatenate([Rg, Rd]); phi = np.concatenate([phig, phid]); z = np.concatenate([zg, zd]); H = np.concatenate([Hg, Hd])
    n = ngas + ndust
    iphase = np.full(n, IGAS, dtype=np.int8)
    iphase[ngas:] = IDUST
    pm = np.where(iphase == IGAS, p.massoftype[IGAS], p.massoftype[IDUST])
    ntype = np.where(iphase == IGAS, ngas, ndust)
    sig0 = (2. - pindex) / (2. * math.pi * (r_out ** (2. - pindex) - r_in ** (2. - pindex))) if abs(pindex - 2.) > 1e-12 else 1.
    rho_mid = pm * ntype * sig0 * R ** (-pindex) / (math.sqrt(2. * math.pi) * H)
    rho = np.maximum(rho_mid * np.exp(-0.5 * (z / H) ** 2), 1e-3 * rho_mid)
    xyzh = np.zeros((n, 4))
    xyzh[:, 0], xyzh[:, 1], xyzh[:, 2] = R * np.cos(phi), R * np.sin(phi), z
    xyzh[:, 3] = p.hfact * (pm / rho) ** (1. / 3.)
    part = Particles(p, xyzh, iphase)
    vk = np.sqrt(1.0 / R)                                   # Keplerian; gas slightly sub-Keplerian through the pressure gradient
    cs2 = p.polyk * R ** (-2. * qindex)
    vphi = np.where(iphase == IGAS, vk * np.sqrt(np.maximum(1. - (pindex + qindex + 1.5) * cs2 / vk ** 2, 0.)), vk)
    part.vxyzu[:, 0], part.vxyzu[:, 1] = -vphi * np.sin(phi), vphi * np.cos(phi)
    r3 = (R * R + z * z) ** 1.5
    part.fext[:, 0], part.fext[:, 1], part.fext[:, 2] = -xyzh[:, 0] / r3, -xyzh[:, 1] / r3, -z / r3      # central star, M = 1
    part.alphaind[:, 0] = p.alpha
    return part


def setup_dustybox(nx=16, idrag=2, isothermal=False, seed=-2468, lattice="random"):
    """two-fluid dusty box (test_dust.f90-like): 30 % of the particles are dust (type idust) drifting through the gas"""
    part, _ = setup_test_derivs(nx=nx, lattice=lattice, isothermal=isothermal, dust=1, idrag=idrag)
    n = part.npart
    rng = Ran2(seed)
    isdust = rng.draw(n) < 0.3
    part.iphase[isdust] = IDUST
    p = part.params
    p.massoftype[IDUST] = 0.05 * p.massoftype[IGAS]
    part.xyzh[isdust, 3] *= (1.0 / 0.3) ** (1. / 3.)        # each phase gets its own smoothing length from its own number density
    part.xyzh[~isdust, 3] *= (1.0 / 0.7) ** (1. / 3.)
    part.vxyzu[isdust, :3] *= 0.5                             # relative drift between the phases
    if idrag == 2:
        p.K_code = 3.0
    elif idrag == 3:
        p.K_code = 0.04
    else:                                                     # Epstein/Stokes: put the box across the kn = 1 transition
        p.grainsize, p.graindens, p.seff = 0.02, 30.0, 0.01 * 5.0 * 4. / 9.
    part.alphaind[:, 0] = 0.2
    return part, isdust
