"""Analytic fields of the reference's derivs test suite, restated in numpy (test infrastructure).

Every function cites the lines of /root/reference/src/tests/test_derivs.f90 it follows; `checkvalf` / `checkval` restate the
pass criteria of src/tests/utils_testsuite.f90 so that the known-answer tests assert exactly what the reference's own suite asserts
(number of particles outside the tolerance == 0)."""
import math
import numpy as np

PI = math.pi
SMALLVAL = 1.e-6                     # utils_testsuite.f90:56


def nfailed_f(x, val, tol):
    """checkvalfuncr8/r4 (utils_testsuite.f90:195-278): err = |x - val|, divided by |val| when |val| > smallval and err > tol."""
    x = np.asarray(x, dtype=np.float64)
    val = np.broadcast_to(np.asarray(val, dtype=np.float64), x.shape)
    err = np.abs(x - val)
    rel = (np.abs(val) > SMALLVAL) & (err > tol)
    err = np.where(rel, err / np.where(rel, np.abs(val), 1.), err)
    bad = (err > tol) | np.isnan(err)
    return int(np.count_nonzero(bad)), float(np.nanmax(err)) if err.size else 0.


def nfailed_v(x, val, tol):
    """checkval on an array against a scalar (utils_testsuite.f90:470-520): the same error measure as nfailed_f."""
    return nfailed_f(x, np.full(np.shape(x), float(val)), tol)


class Fields:
    """vx..dvzdzdz (test_derivs.f90:1215-1448), divv/curlv (:1455-1492), B and derivatives (:1947-2100), psi (:2380-2389)."""

    def __init__(self, xyzh, p, Bext=(0., 0., 0.)):
        self.p = p
        self.x, self.y, self.z, self.h = xyzh[:, 0], xyzh[:, 1], xyzh[:, 2], xyzh[:, 3]
        self.dxb, self.dyb, self.dzb = p.xmax - p.xmin, p.ymax - p.ymin, p.zmax - p.zmin
        self.ax = 2. * PI * (self.x - p.xmin) / self.dxb
        self.ay = 2. * PI * (self.y - p.ymin) / self.dyb
        self.az = 2. * PI * (self.z - p.zmin) / self.dzb
        self.Bext = Bext
        self.zero = np.zeros_like(self.x)

    # velocity (:1215-1240)
    def vx(self): return 0.5 / PI * self.dxb * np.sin(self.ax)
    def vy(self): return 0.5 / PI * self.dxb * np.sin(self.ax) - 0.5 / PI * self.dzb * np.sin(self.az)
    def vz(self): return 0.05 / PI * self.dyb * np.cos(2. * self.ay)
    # first derivatives (:1242-1312)
    def dvxdx(self): return np.cos(self.ax)
    def dvydx(self): return np.cos(self.ax)
    def dvydz(self): return -np.cos(self.az)
    def dvzdy(self): return -0.2 * np.sin(2. * self.ay)
    # second derivatives that do not vanish (:1317-1448)
    def dvxdxdx(self): return -2. * PI / self.dxb * np.sin(self.ax)
    def dvydxdx(self): return -2. * PI / self.dxb * np.sin(self.ax)
    def dvydzdz(self): return 2. * PI / self.dzb * np.sin(self.az)
    def dvzdydy(self): return -0.8 * PI / self.dyb * np.cos(2. * self.ay)
    def divv(self): return self.dvxdx()                                   # dvydy = dvzdz = 0 (:1455-1460)
    def curlv(self):                                                       # (:1467-1492)
        return self.dvzdy() - self.dvydz(), self.zero, self.dvydx()
    # thermal energy (:1604-1613), constant variant (:1621-1626)
    def utherm(self): return 0.5 / PI * (3. + np.sin(self.ax) + np.cos(self.ay) + np.sin(self.az))
    # magnetic field (:1947-1978) and first derivatives (:1980-2062)
    def Bx(self): return -5. / PI * self.dzb * np.cos(self.az) + self.Bext[0] + 0.5 / PI * self.dxb * np.sin(self.ax)
    def By(self): return 5. / PI * self.dxb * np.sin(self.ax) + self.Bext[1]
    def Bz(self): return 15. / PI * self.dyb * np.cos(self.ay) + self.Bext[2]
    def dBxdx(self): return np.cos(self.ax)
    def dBxdz(self): return 10. * np.sin(self.az)
    def dBydx(self): return 10. * np.cos(self.ax)
    def dBzdy(self): return -30. * np.sin(self.ay)
    def divB(self): return self.dBxdx()                                    # (:2175-2180)
    def curlB(self):                                                       # (:2187-2207)
        return self.dBzdy(), self.dBxdz(), self.dBydx()
    def psi(self):                                                         # (:2380-2389)
        return 0.5 / PI * self.dxb * np.sin(self.ax) - 0.5 / PI * self.dzb * np.cos(self.az) + 0.5 / PI * self.dyb * np.sin(self.ay)


def ddivvdt_full(f):
    """ddivvdtfunc (test_derivs.f90:1536-1544) written out with every term: div a - (dvxdx^2 + dvydy^2 + dvzdz^2 +
    2 (dvxdy dvydx + dvxdz dvzdx + dvydz dvzdy))."""
    return f.divv() - (f.dvxdx() ** 2 + 2. * (f.dvydz() * f.dvzdy()))
