"""Host-side mirror of the reference's hot-path interface on top of the C ABI (include/sphgpu.h).

`SphGpu` exposes the same call sequence `derivs` makes in the reference
(src/main/deriv.f90:113-192):

    build_tree(npart,nactive,xyzh,vxyzu)        neigh_kdtree.f90:161
    densityiterate(icall,...)                   dens.F90:117
    cons2prim_everything(...)                   cons2prim.f90:274
    force(icall,...)                            force.F90:193

with the same argument meaning (arrays in Fortran layout, updated in place)
and the reference's error behaviour (`fatal` -> `SphGpuError`).  There is no
CPU fallback: if the CUDA library is missing or no device is present the
constructor raises.
"""
import ctypes as C
import os
import numpy as np

from .params import SphParams, SphScalars, SphStepOut, SphEnergies

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPHGPU_LIB", os.path.join(_HERE, "libsphgpu.so"))   # SPHGPU_LIB: A/B builds of the same library (tools/variants.sh)

F_XYZH, F_VXYZU, F_FXYZU, F_FEXT, F_BEVOL, F_DBEVOL, F_EOSVARS, F_DIVCURLV, F_DIVCURLB, F_ALPHAIND, F_GRADH, F_DVDX, \
    F_POTEN, F_DIVBSYMM, F_IPHASE, F_IBIN, F_DUSTFRAC, F_TSTOP = [1 << k for k in range(18)]
F_ALL = (1 << 64) - 1

ERRORS = {1: "CUDA", 2: "ARG", 3: "NAN", 4: "NOPART", 5: "NOCONVERGE", 6: "NEGH", 7: "OVERFLOW", 8: "STATE"}

EXPORTS = [
    "sphgpu_create", "sphgpu_destroy", "sphgpu_set_params", "sphgpu_last_error", "sphgpu_set_option", "sphgpu_get_timings",
    "sphgpu_launch_count", "sphgpu_get_kernel_timings", "sphgpu_upload", "sphgpu_download", "sphgpu_build_tree_resident", "sphgpu_densityiterate_resident",
    "sphgpu_cons2prim_resident", "sphgpu_force_resident", "sphgpu_derivs_resident", "sphgpu_build_tree", "sphgpu_densityiterate",
    "sphgpu_cons2prim_everything", "sphgpu_force", "sphgpu_derivs", "sphgpu_get_neighbour_stats", "sphgpu_neighbour_sets",
    "sphgpu_measure_fp64_peak", "sphgpu_measure_copy_bw", "sphgpu_local_hmax", "sphgpu_halo_select", "sphgpu_halo_pack",
    "sphgpu_halo_recvbuf", "sphgpu_halo_unpack", "sphgpu_nghost", "sphgpu_set_timestep_bins", "sphgpu_get_gravity_timings", "sphgpu_gravity_tree", "sphgpu_step_resident", "sphgpu_energies_resident", "sphgpu_gravity_gather_pack", "sphgpu_gravity_gather_recvbuf",
    "sphgpu_gravity_gather_unpack", "sphgpu_density_hmax_used", "sphgpu_halo_restore_h", "sphgpu_set_forcing_modes", "sphgpu_forcing_resident", "sphgpu_get_copy_bytes", "sphgpu_density_hgrow",
    "sphgpu_dist_get_unique_id", "sphgpu_dist_init", "sphgpu_dist_finalize", "sphgpu_dist_set_boxes", "sphgpu_dist_get_boxes", "sphgpu_dist_set_ids",
    "sphgpu_dist_get_ids", "sphgpu_dist_nlocal", "sphgpu_dist_derivs", "sphgpu_dist_migrate", "sphgpu_dist_rebalance", "sphgpu_dist_step",
    "sphgpu_dist_energies", "sphgpu_dist_stats", "sphgpu_init_step_resident", "sphgpu_set_active_particles_resident", "sphgpu_step_ind_resident",
]


class SphGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sphgpu error {code} ({ERRORS.get(code, '?')}): {msg}")
        self.code = code


class HostArrays(C.Structure):
    _fields_ = [("npart", C.c_int64)] + [(k, C.c_void_p) for k in (
        "xyzh", "vxyzu", "fxyzu", "fext", "Bevol", "dBevol", "eos_vars", "divcurlv", "divcurlB", "alphaind", "gradh", "dvdx",
        "poten", "divBsymm", "iphase", "ibin", "ibin_old", "ibin_wake", "dustfrac", "tstop")]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises (never falls back) if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C phantom_b200/csrc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.sphgpu_create.argtypes = [C.POINTER(SphParams), i32, C.POINTER(vp)]
        L.sphgpu_destroy.argtypes = [vp]
        L.sphgpu_destroy.restype = None
        L.sphgpu_set_params.argtypes = [vp, C.POINTER(SphParams)]
        L.sphgpu_last_error.argtypes = [vp]
        L.sphgpu_last_error.restype = C.c_char_p
        L.sphgpu_set_option.argtypes = [vp, C.c_char_p, dbl]
        L.sphgpu_get_timings.argtypes = [vp, C.POINTER(dbl)]
        L.sphgpu_get_kernel_timings.argtypes = [vp, C.POINTER(dbl)]
        L.sphgpu_launch_count.argtypes = [vp]
        L.sphgpu_launch_count.restype = i64
        L.sphgpu_upload.argtypes = [vp, C.POINTER(HostArrays), C.c_uint64]
        L.sphgpu_download.argtypes = [vp, C.POINTER(HostArrays), C.c_uint64]
        L.sphgpu_build_tree_resident.argtypes = [vp]
        L.sphgpu_densityiterate_resident.argtypes = [vp, i32, C.POINTER(SphScalars)]
        L.sphgpu_cons2prim_resident.argtypes = [vp]
        L.sphgpu_force_resident.argtypes = [vp, i32, dbl, C.POINTER(SphScalars)]
        L.sphgpu_derivs_resident.argtypes = [vp, i32, dbl, C.POINTER(SphScalars)]
        L.sphgpu_build_tree.argtypes = [vp, i64, i64, vp, vp, vp]
        L.sphgpu_densityiterate.argtypes = [vp, i32, i64, i64, vp, vp, vp, vp, vp, C.POINTER(dbl), vp, vp, vp, vp, vp, vp, C.POINTER(SphScalars)]
        L.sphgpu_cons2prim_everything.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
        L.sphgpu_force.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, vp, vp, vp, vp, vp, vp, vp, C.POINTER(SphScalars)]
        L.sphgpu_derivs.argtypes = [vp, i32, C.POINTER(HostArrays), dbl, C.POINTER(SphScalars)]
        L.sphgpu_get_neighbour_stats.argtypes = [vp, C.POINTER(SphScalars)]
        L.sphgpu_neighbour_sets.argtypes = [vp, i32, vp, vp, i64]
        L.sphgpu_neighbour_sets.restype = i64
        L.sphgpu_measure_fp64_peak.argtypes = [vp]
        L.sphgpu_measure_fp64_peak.restype = dbl
        L.sphgpu_measure_copy_bw.argtypes = [vp]
        L.sphgpu_measure_copy_bw.restype = dbl
        L.sphgpu_local_hmax.argtypes = [vp, C.POINTER(dbl)]
        L.sphgpu_halo_select.argtypes = [vp, i32, i32, vp, dbl, vp]
        L.sphgpu_halo_pack.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i32)]
        L.sphgpu_halo_recvbuf.argtypes = [vp, i64, i32, C.POINTER(vp)]
        L.sphgpu_halo_unpack.argtypes = [vp, i32, i64]
        L.sphgpu_nghost.argtypes = [vp]
        L.sphgpu_nghost.restype = i64
        L.sphgpu_set_timestep_bins.argtypes = [vp, i32, i32, i32]
        L.sphgpu_get_gravity_timings.argtypes = [vp, C.POINTER(dbl)]
        L.sphgpu_gravity_tree.argtypes = [vp, i64, vp, vp, vp]
        L.sphgpu_gravity_tree.restype = i64
        L.sphgpu_step_resident.argtypes = [vp, dbl, dbl, C.POINTER(SphStepOut)]
        L.sphgpu_energies_resident.argtypes = [vp, C.POINTER(SphEnergies)]
        L.sphgpu_set_forcing_modes.argtypes = [vp, i32, vp, vp, vp, vp, dbl, dbl, i32]
        L.sphgpu_forcing_resident.argtypes = [vp]
        L.sphgpu_get_copy_bytes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.sphgpu_density_hgrow.argtypes = [vp, C.POINTER(dbl)]
        L.sphgpu_density_hmax_used.argtypes = [vp, C.POINTER(dbl)]
        L.sphgpu_halo_restore_h.argtypes = [vp]
        L.sphgpu_gravity_gather_pack.argtypes = [vp, C.POINTER(vp), C.POINTER(i32)]
        L.sphgpu_gravity_gather_recvbuf.argtypes = [vp, i32, i64, C.POINTER(vp)]
        L.sphgpu_gravity_gather_unpack.argtypes = [vp, i32, i32, i64, vp]
        L.sphgpu_dist_get_unique_id.argtypes = [vp, i32]
        L.sphgpu_dist_init.argtypes = [vp, vp, i32, i32]
        L.sphgpu_dist_finalize.argtypes = [vp]
        L.sphgpu_dist_set_boxes.argtypes = [vp, vp]
        L.sphgpu_dist_get_boxes.argtypes = [vp, vp]
        L.sphgpu_dist_set_ids.argtypes = [vp, vp, i64]
        L.sphgpu_dist_get_ids.argtypes = [vp, vp, i64]
        L.sphgpu_dist_nlocal.argtypes = [vp]
        L.sphgpu_dist_nlocal.restype = i64
        L.sphgpu_dist_derivs.argtypes = [vp, i32, dbl, C.POINTER(SphScalars)]
        L.sphgpu_dist_migrate.argtypes = [vp, i32, C.POINTER(i64)]
        L.sphgpu_dist_rebalance.argtypes = [vp, vp, C.POINTER(i64)]
        L.sphgpu_dist_step.argtypes = [vp, dbl, dbl, C.POINTER(SphStepOut)]
        L.sphgpu_dist_energies.argtypes = [vp, C.POINTER(SphEnergies)]
        L.sphgpu_dist_stats.argtypes = [vp, C.POINTER(dbl)]
        L.sphgpu_init_step_resident.argtypes = [vp, dbl, dbl, i32]
        L.sphgpu_set_active_particles_resident.argtypes = [vp, i32, i32, C.POINTER(i64), C.POINTER(i64)]
        L.sphgpu_step_ind_resident.argtypes = [vp, dbl, dbl, dbl, C.POINTER(SphStepOut)]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_arrays(part):
    h = HostArrays()
    h.npart = part.npart
    for name, _ in HostArrays._fields_[1:]:
        arr = getattr(part, name, None)
        setattr(h, name, None if arr is None else arr.ctypes.data)
    return h


class SphGpu:
    """One context = one GPU (one process per GPU under torchrun)."""

    def __init__(self, params, device=0):
        self.L = load_library()
        self.params = params
        h = C.c_void_p()
        rc = self.L.sphgpu_create(C.byref(params), device, C.byref(h))
        if rc != 0:
            raise SphGpuError(rc, "sphgpu_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.sphgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SphGpuError(rc, self.L.sphgpu_last_error(self.h).decode())

    def set_params(self, params):
        self.params = params
        self._check(self.L.sphgpu_set_params(self.h, C.byref(params)))

    def set_option(self, name, value):
        self._check(self.L.sphgpu_set_option(self.h, name.encode(), float(value)))

    # ---- literal mode: the reference's argument lists ------------------------------------------------
    def build_tree(self, part):
        self._check(self.L.sphgpu_build_tree(self.h, part.npart, part.npart, _p(part.xyzh), _p(part.vxyzu), _p(part.iphase)))

    def densityiterate(self, part, icall=1):
        sc = SphScalars()
        stressmax = C.c_double(0.)
        self._check(self.L.sphgpu_densityiterate(
            self.h, icall, part.npart, part.npart, _p(part.xyzh), _p(part.vxyzu), _p(part.divcurlv), _p(part.divcurlB), _p(part.Bevol),
            C.byref(stressmax), _p(part.fxyzu), _p(part.fext), _p(part.alphaind), _p(part.gradh), _p(part.dvdx), _p(part.iphase), C.byref(sc)))
        if self.params.dust:               # hidden module output dustfrac (dens.F90:1612-1624)
            self.download(part, F_DUSTFRAC)
        return sc

    def cons2prim_everything(self, part):
        self._check(self.L.sphgpu_cons2prim_everything(
            self.h, part.npart, _p(part.xyzh), _p(part.vxyzu), _p(part.dvdx), _p(part.eos_vars), _p(part.Bevol), _p(part.Bxyz),
            _p(part.alphaind), _p(part.iphase)))

    def set_timestep_bins(self, nbinmax, ibinnow, istepfrac):
        """timestep_ind module inputs of the force pass (utils_indtimesteps.f90)"""
        self._check(self.L.sphgpu_set_timestep_bins(self.h, int(nbinmax), int(ibinnow), int(istepfrac)))

    def force(self, part, icall=1, dt=0.0):
        sc = SphScalars()
        if self.params.ind_timesteps:      # hidden module inputs ibin, ibin_old, ibin_wake (part.F90)
            self.upload(part, F_IBIN)
        self._force_literal(part, icall, dt, sc)
        extra = (F_IBIN if self.params.ind_timesteps else 0) | (F_TSTOP if self.params.dust else 0)
        if extra:                          # hidden module outputs ibin, ibin_wake, tstop
            self.download(part, extra)
        return sc

    def _force_literal(self, part, icall, dt, sc):
        self._check(self.L.sphgpu_force(
            self.h, icall, part.npart, _p(part.xyzh), _p(part.vxyzu), _p(part.fxyzu), _p(part.divcurlv), _p(part.divcurlB), _p(part.Bevol),
            _p(part.dBevol), _p(part.fext), dt, 0.0, _p(part.eos_vars), _p(part.alphaind), _p(part.gradh), _p(part.dvdx), _p(part.iphase),
            _p(part.poten), _p(part.divBsymm), C.byref(sc)))

    def derivs(self, part, icall=1, dt=0.0):
        """derivs(icall,...) with host arrays in and out (deriv.f90:37)."""
        sc = SphScalars()
        h = host_arrays(part)
        self._check(self.L.sphgpu_derivs(self.h, icall, C.byref(h), dt, C.byref(sc)))
        return sc

    def derivs_by_phase(self, part):
        """the four literal calls in the order of deriv.f90:113-192; returns (density scalars, force scalars)"""
        self.build_tree(part)
        sd = self.densityiterate(part, 1)
        self.params.set_boundaries_to_active = 0
        self.set_params(self.params)
        self.cons2prim_everything(part)
        sf = self.force(part, 1)
        return sd, sf

    # ---- resident mode ----------------------------------------------------------------------------------
    def upload(self, part, mask=F_ALL):
        self.npart_uploaded = part.npart        # owned particles of this context (ghosts are appended by the halo exchange)
        h = host_arrays(part)
        self._check(self.L.sphgpu_upload(self.h, C.byref(h), mask))

    def download(self, part, mask=F_ALL):
        h = host_arrays(part)
        self._check(self.L.sphgpu_download(self.h, C.byref(h), mask))

    def derivs_resident(self, icall=1, dt=0.0):
        sc = SphScalars()
        self._check(self.L.sphgpu_derivs_resident(self.h, icall, dt, C.byref(sc)))
        return sc

    def build_tree_resident(self):
        self._check(self.L.sphgpu_build_tree_resident(self.h))

    def densityiterate_resident(self, icall=1):
        sc = SphScalars()
        self._check(self.L.sphgpu_densityiterate_resident(self.h, icall, C.byref(sc)))
        return sc

    def cons2prim_resident(self):
        self._check(self.L.sphgpu_cons2prim_resident(self.h))

    def force_resident(self, icall=1, dt=0.0):
        sc = SphScalars()
        self._check(self.L.sphgpu_force_resident(self.h, icall, dt, C.byref(sc)))
        return sc

    def set_forcing_modes(self, mode, ampl, aka, akb, amplfac=1.0, solweightnorm=1.0, correct_mean_force=False):
        """current stirring mode set of forcing.f90 (st_mode, st_ampl, st_aka, st_akb, st_amplfac, st_solweightnorm)"""
        mode, ampl, aka, akb = [np.ascontiguousarray(a, dtype=np.float64) for a in (mode, ampl, aka, akb)]
        self._check(self.L.sphgpu_set_forcing_modes(self.h, len(ampl), _p(mode), _p(ampl), _p(aka), _p(akb), float(amplfac), float(solweightnorm),
                                                    int(correct_mean_force)))

    def forcing_resident(self):
        self._check(self.L.sphgpu_forcing_resident(self.h))

    def step_resident(self, dtsph, tolv=1.e-2):
        """one leapfrog step of the resident state (step_leapfrog.f90:95, global timesteps)"""
        out = SphStepOut()
        self._check(self.L.sphgpu_step_resident(self.h, float(dtsph), float(tolv), C.byref(out)))
        return out

    def init_step_resident(self, time, dtmax, nbinmax):
        """init_step with individual timesteps (step_leapfrog.f90:57-80)"""
        self._check(self.L.sphgpu_init_step_resident(self.h, float(time), float(dtmax), int(nbinmax)))

    def set_active_particles_resident(self, nbinmax, istepfrac):
        """set_active_particles (utils_indtimesteps.f90:114-178) -> (nactive, nalive)"""
        na, nl = C.c_int64(), C.c_int64()
        self._check(self.L.sphgpu_set_active_particles_resident(self.h, int(nbinmax), int(istepfrac), C.byref(na), C.byref(nl)))
        return na.value, nl.value

    def step_ind_resident(self, t, dtsph, dtmax):
        """one call of step() with individual timesteps (dtsph = dtmax / 2^nbinmax)"""
        out = SphStepOut()
        self._check(self.L.sphgpu_step_ind_resident(self.h, float(t), float(dtsph), float(dtmax), C.byref(out)))
        return out

    def energies_resident(self):
        """compute_energies (energies.f90:64) on the resident state"""
        out = SphEnergies()
        self._check(self.L.sphgpu_energies_resident(self.h, C.byref(out)))
        return out

    def timings_ms(self):
        t = (C.c_double * 4)()
        self.L.sphgpu_get_timings(self.h, t)
        return dict(tree=t[0], dens=t[1], cons2prim=t[2], force=t[3])

    def kernel_timings_ms(self):
        t = (C.c_double * 2)()
        self.L.sphgpu_get_kernel_timings(self.h, t)
        return dict(density=t[0], force=t[1])

    def gravity_timings_ms(self):
        t = (C.c_double * 2)()
        self.L.sphgpu_get_gravity_timings(self.h, t)
        return dict(total=t[0], p2p=t[1])

    def gravity_tree(self, npart):
        """kdnode records of the self-gravity tree after the last force call (see include/sphgpu.h)"""
        nn = self.L.sphgpu_gravity_tree(self.h, 0, None, None, None)
        if nn < 0:
            raise SphGpuError(8, "no gravity tree: call force with gravity first")
        rec = np.zeros((nn, 12))
        irec = np.zeros((nn, 6), dtype=np.int32)
        ids = np.zeros(npart, dtype=np.int32)
        self.L.sphgpu_gravity_tree(self.h, nn, _p(rec), _p(irec), _p(ids))
        return rec, irec, ids

    def copy_bytes(self):
        """(host->device, device->host) bytes of the last derivs() call"""
        a, b = C.c_int64(), C.c_int64()
        self._check(self.L.sphgpu_get_copy_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return self.L.sphgpu_launch_count(self.h)

    def neighbour_sets(self, npart, symmetric=False):
        off = np.zeros(npart + 1, dtype=np.int64)
        cap = 200 * npart
        lst = np.zeros(cap, dtype=np.int32)
        tot = self.L.sphgpu_neighbour_sets(self.h, int(symmetric), _p(off), _p(lst), cap)
        if tot < -1:
            cap = -tot
            lst = np.zeros(cap, dtype=np.int32)
            tot = self.L.sphgpu_neighbour_sets(self.h, int(symmetric), _p(off), _p(lst), cap)
        if tot < 0:
            raise SphGpuError(7, self.L.sphgpu_last_error(self.h).decode())
        return off, lst[:tot]

    # ---- multi-GPU halo (see halo.py for the orchestration) ---------------------------------------------
    def local_hmax(self):
        v = C.c_double()
        self._check(self.L.sphgpu_local_hmax(self.h, C.byref(v)))
        return v.value

    def density_hmax_used(self):
        v = C.c_double()
        self._check(self.L.sphgpu_density_hmax_used(self.h, C.byref(v)))
        return v.value

    def density_hgrow(self):
        v = C.c_double()
        self._check(self.L.sphgpu_density_hgrow(self.h, C.byref(v)))
        return v.value

    def halo_restore_h(self):
        self._check(self.L.sphgpu_halo_restore_h(self.h))

    def halo_select(self, nranks, myrank, boxes, dhalo):
        boxes = np.ascontiguousarray(boxes, dtype=np.float64)
        counts = np.zeros(nranks, dtype=np.int64)
        self._check(self.L.sphgpu_halo_select(self.h, nranks, myrank, _p(boxes), float(dhalo), _p(counts)))
        return counts

    def halo_pack(self, stage):
        ptr, rd = C.c_void_p(), C.c_int32()
        self._check(self.L.sphgpu_halo_pack(self.h, stage, C.byref(ptr), C.byref(rd)))
        return ptr.value, rd.value

    def halo_recvbuf(self, nrecords, record_doubles):
        ptr = C.c_void_p()
        self._check(self.L.sphgpu_halo_recvbuf(self.h, int(nrecords), int(record_doubles), C.byref(ptr)))
        return ptr.value

    def halo_unpack(self, stage, nghost):
        self._check(self.L.sphgpu_halo_unpack(self.h, stage, int(nghost)))

    def gravity_gather_pack(self):
        ptr, rd = C.c_void_p(), C.c_int32()
        self._check(self.L.sphgpu_gravity_gather_pack(self.h, C.byref(ptr), C.byref(rd)))
        return ptr.value, rd.value

    def gravity_gather_recvbuf(self, nranks, stride):
        ptr = C.c_void_p()
        self._check(self.L.sphgpu_gravity_gather_recvbuf(self.h, int(nranks), int(stride), C.byref(ptr)))
        return ptr.value

    def gravity_gather_unpack(self, nranks, myrank, stride, counts):
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        self._check(self.L.sphgpu_gravity_gather_unpack(self.h, int(nranks), int(myrank), int(stride), _p(counts)))

    def nghost(self):
        return self.L.sphgpu_nghost(self.h)

    # ---- multi-GPU driver behind the C ABI (csrc/dist.cu; halo.DistSph wraps these) ---------------------
    @staticmethod
    def dist_unique_id():
        """128-byte NCCL id (rank 0 calls this and broadcasts it)"""
        buf = (C.c_ubyte * 128)()
        rc = load_library().sphgpu_dist_get_unique_id(C.cast(buf, C.c_void_p), 128)
        if rc != 0:
            raise SphGpuError(rc, "sphgpu_dist_get_unique_id failed (NCCL not loadable?)")
        return bytes(buf)

    def dist_init(self, uid, nranks, rank):
        buf = (C.c_ubyte * 128).from_buffer_copy(uid)
        self._check(self.L.sphgpu_dist_init(self.h, C.cast(buf, C.c_void_p), int(nranks), int(rank)))

    def dist_finalize(self):
        self._check(self.L.sphgpu_dist_finalize(self.h))

    def dist_set_boxes(self, boxes):
        boxes = np.ascontiguousarray(boxes, dtype=np.float64)
        self._check(self.L.sphgpu_dist_set_boxes(self.h, _p(boxes)))

    def dist_get_boxes(self, nranks):
        boxes = np.zeros((nranks, 6))
        self._check(self.L.sphgpu_dist_get_boxes(self.h, _p(boxes)))
        return boxes

    def dist_set_ids(self, ids=None, base=0):
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
        self._check(self.L.sphgpu_dist_set_ids(self.h, _p(ids), int(base)))

    def dist_get_ids(self):
        n = self.dist_nlocal()
        ids = np.zeros(n, dtype=np.int64)
        self._check(self.L.sphgpu_dist_get_ids(self.h, _p(ids), n))
        return ids

    def dist_nlocal(self):
        return int(self.L.sphgpu_dist_nlocal(self.h))

    def dist_derivs(self, icall=1, dt=0.0):
        sc = SphScalars()
        self._check(self.L.sphgpu_dist_derivs(self.h, int(icall), float(dt), C.byref(sc)))
        return sc

    def dist_migrate(self, in_step=False):
        n = C.c_int64()
        self._check(self.L.sphgpu_dist_migrate(self.h, int(in_step), C.byref(n)))
        return n.value

    def dist_rebalance(self, domain):
        domain = np.ascontiguousarray(domain, dtype=np.float64)
        n = C.c_int64()
        self._check(self.L.sphgpu_dist_rebalance(self.h, _p(domain), C.byref(n)))
        return n.value

    def dist_step(self, dtsph, tolv=1.e-2):
        out = SphStepOut()
        self._check(self.L.sphgpu_dist_step(self.h, float(dtsph), float(tolv), C.byref(out)))
        return out

    def dist_energies(self):
        out = SphEnergies()
        self._check(self.L.sphgpu_dist_energies(self.h, C.byref(out)))
        return out

    def dist_stats(self):
        t = (C.c_double * 8)()
        self._check(self.L.sphgpu_dist_stats(self.h, t))
        return dict(ghost_capacity=int(t[0]), nghost=int(t[1]), halo_bytes=int(t[2]), rounds=int(t[3]), migrated=int(t[4]), hmax_used=t[5], ms=t[6],
                    nlocal=int(t[7]))

    def measure_fp64_peak(self):
        return self.L.sphgpu_measure_fp64_peak(self.h)

    def measure_copy_bw(self):
        return self.L.sphgpu_measure_copy_bw(self.h)
