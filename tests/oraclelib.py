"""ctypes front-end of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess
import numpy as np

from phantom_b200.params import SphParams, SphScalars

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("sph_oracle.cpp", "sph_oracle_force.inc", "sph_oracle.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src if os.path.exists(s)):
            build()
        L = C.CDLL(so)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(SphParams)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_params.argtypes = [C.c_void_p, C.POINTER(SphParams)]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_last_error.argtypes = [C.c_void_p]
        L.oracle_tree_ncells.restype = C.c_int64
        L.oracle_tree_ncells.argtypes = [C.c_void_p]
        L.oracle_get_neighbour_list.restype = C.c_int64
        L.oracle_neighbour_sets.restype = C.c_int64
        L.oracle_neighbour_counts_bruteforce.restype = C.c_int64
        L.oracle_ran2.restype = C.c_double
        L.oracle_kernel.argtypes = [C.c_int, C.c_double, C.c_double] + [C.POINTER(C.c_double)] * 6
        L.oracle_kernel_constants.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 7
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Mirror of the reference call sequence in derivs (src/main/deriv.f90:113-192)."""

    def __init__(self, params):
        self.L = lib()
        self.params = params
        self.h = C.c_void_p(self.L.oracle_create(C.byref(params)))

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    def set_threads(self, n):
        self.L.oracle_set_threads(C.c_int(n))

    def max_threads(self):
        return self.L.oracle_get_max_threads()

    def _check(self, ierr):
        if ierr != 0:
            raise RuntimeError(self.L.oracle_last_error(self.h).decode())

    def set_params(self, params):
        self.params = params
        self.L.oracle_set_params(self.h, C.byref(params))

    def build_tree(self, part):
        self._check(self.L.oracle_build_tree(self.h, C.c_int64(part.npart), _p(part.xyzh), _p(part.iphase)))

    def ncells(self):
        return self.L.oracle_tree_ncells(self.h)

    def node(self, n):
        rec = np.zeros(12)
        irec = np.zeros(6, dtype=np.int32)
        self.L.oracle_tree_get_node(self.h, C.c_int64(n), _p(rec), _p(irec))
        return rec, irec

    def inodeparts(self, npart):
        out = np.zeros(npart, dtype=np.int32)
        self.L.oracle_tree_get_inodeparts(self.h, _p(out))
        return out

    def neighbour_list(self, icell, getj=False, maxlist=1 << 20):
        out = np.zeros(maxlist, dtype=np.int32)
        n = self.L.oracle_get_neighbour_list(self.h, C.c_int64(icell), C.c_int(int(getj)), _p(out), C.c_int64(maxlist))
        return out[:n]

    def densityiterate(self, part, icall=1):
        sc = SphScalars()
        self._check(self.L.oracle_densityiterate(
            self.h, C.c_int(icall), C.c_int64(part.npart), _p(part.xyzh), _p(part.vxyzu), _p(part.fxyzu), _p(part.fext),
            _p(part.Bevol), _p(part.iphase), _p(part.divcurlv), _p(part.divcurlB), _p(part.alphaind), _p(part.gradh),
            _p(part.dvdx), _p(part.dustfrac), C.byref(sc)))
        return sc

    def cons2prim(self, part):
        self._check(self.L.oracle_cons2prim(
            self.h, C.c_int64(part.npart), _p(part.xyzh), _p(part.vxyzu), _p(part.dvdx), _p(part.Bevol), _p(part.iphase),
            _p(part.eos_vars), _p(part.alphaind), _p(part.Bxyz)))

    def force(self, part, icall=1, dt=0.0, nbinmax=0, ibinnow=0, istepfrac=0):
        sc = SphScalars()
        self._check(self.L.oracle_force(
            self.h, C.c_int(icall), C.c_int64(part.npart), _p(part.xyzh), _p(part.vxyzu), _p(part.fxyzu), _p(part.divcurlv),
            _p(part.divcurlB), _p(part.Bevol), _p(part.dBevol), _p(part.fext), _p(part.eos_vars), _p(part.alphaind),
            _p(part.gradh), _p(part.dvdx), _p(part.iphase), _p(part.dustfrac), C.c_double(dt), _p(part.poten),
            _p(part.divBsymm), _p(part.tstop), _p(part.ibin), _p(part.ibin_wake), _p(part.ibin_old),
            C.c_int(nbinmax), C.c_int(ibinnow), C.c_int(istepfrac), C.byref(sc)))
        return sc

    def forcing(self, part, mode, ampl, aka, akb, amplfac=1.0, solweightnorm=1.0, correct_mean_force=False):
        mode, ampl, aka, akb = [np.ascontiguousarray(a, dtype=np.float64) for a in (mode, ampl, aka, akb)]
        self.L.oracle_forcing(self.h, C.c_int64(part.npart), _p(part.xyzh), _p(part.iphase), _p(part.fxyzu), C.c_int(len(ampl)), _p(mode), _p(ampl),
                              _p(aka), _p(akb), C.c_double(amplfac), C.c_double(solweightnorm), C.c_int(int(correct_mean_force)))

    def derivs(self, part, icall=1, dt=0.0):
        """derivs(icall=1): tree -> density -> cons2prim -> force (deriv.f90:113-192)."""
        self.build_tree(part)
        sd = self.densityiterate(part, 1)
        self.params.set_boundaries_to_active = 0      # deriv.f90:146
        self.set_params(self.params)
        self.cons2prim(part)
        sf = self.force(part, icall, dt)
        return sd, sf

    def neighbour_sets(self, part, symmetric=False):
        n = part.npart
        off = np.zeros(n + 1, dtype=np.int64)
        cap = 400 * n
        lst = np.zeros(cap, dtype=np.int32)
        tot = self.L.oracle_neighbour_sets(self.h, C.c_int64(n), _p(part.xyzh), _p(part.iphase), C.c_int(int(symmetric)),
                                           _p(off), _p(lst), C.c_int64(cap))
        if tot < 0:
            cap = -tot
            lst = np.zeros(cap, dtype=np.int32)
            tot = self.L.oracle_neighbour_sets(self.h, C.c_int64(n), _p(part.xyzh), _p(part.iphase), C.c_int(int(symmetric)),
                                               _p(off), _p(lst), C.c_int64(cap))
        return off, lst[:tot]

    def neighbour_counts_bruteforce(self, part, symmetric=False):
        cnt = np.zeros(part.npart, dtype=np.int32)
        tot = self.L.oracle_neighbour_counts_bruteforce(self.h, C.c_int64(part.npart), _p(part.xyzh), C.c_int(int(symmetric)), _p(cnt))
        return tot, cnt


def kernel(kid, q):
    L = lib()
    out = [C.c_double() for _ in range(6)]
    L.oracle_kernel(kid, q * q, q, *[C.byref(o) for o in out])
    return tuple(o.value for o in out)   # w, grw, dphidh, potensoft, fsoft, wdrag


def kernel_constants(kid):
    L = lib()
    out = [C.c_double() for _ in range(7)]
    L.oracle_kernel_constants(kid, *[C.byref(o) for o in out])
    return tuple(o.value for o in out)   # radkern, cnormk, wab0, gradh0, dphidh0, cnormk_drag, hfact_default


def ran2(seed_ref):
    """seed_ref: one-element int32 numpy array, updated in place."""
    return lib().oracle_ran2(_p(seed_ref))
