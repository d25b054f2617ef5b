"""ctypes mirrors of the structs in include/sphgpu.h.

`SphParams` carries what the reference keeps in compile-time cpp flags
(build/Makefile:189-300 -> module dim, src/main/config.F90) and in module
variables (options.f90, timestep.f90, shock_capturing.f90, eos.f90,
boundary.f90, kdtree.F90:46) and that the hot path reads implicitly
(SURVEY.md section 8b "hidden inputs").
"""
import ctypes as C

MAXTYPES = 8
# particle types, src/main/part.F90:428-438
IGAS, IBOUNDARY, ISTAR, IDARKMATTER, IBULGE, IDUST = 1, 3, 4, 5, 6, 7
KERNEL_CUBIC, KERNEL_QUINTIC = 0, 1


class SphParams(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("periodic", C.c_int32), ("isothermal", C.c_int32), ("mhd", C.c_int32),
        ("gravity", C.c_int32), ("dust", C.c_int32), ("const_av", C.c_int32), ("ind_timesteps", C.c_int32),
        ("disc_viscosity", C.c_int32), ("ieos", C.c_int32),
        ("ipdv_heating", C.c_int32), ("ishock_heating", C.c_int32), ("iresistive_heating", C.c_int32),
        ("set_boundaries_to_active", C.c_int32), ("idrag", C.c_int32), ("driving", C.c_int32), ("reserved_i", C.c_int32 * 4),
        ("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
        ("zmin", C.c_double), ("zmax", C.c_double),
        ("hfact", C.c_double), ("tolh", C.c_double),
        ("massoftype", C.c_double * MAXTYPES),
        ("alpha", C.c_double), ("alphamax", C.c_double), ("alphau", C.c_double), ("alphaB", C.c_double),
        ("beta", C.c_double),
        ("polyk", C.c_double), ("gamma", C.c_double), ("qfacdisc", C.c_double), ("cs_min", C.c_double),
        ("C_cour", C.c_double), ("C_force", C.c_double), ("dtmax", C.c_double),
        ("psidecayfac", C.c_double), ("overcleanfac", C.c_double),
        ("tree_accuracy", C.c_double),
        ("grainsize", C.c_double), ("graindens", C.c_double), ("K_code", C.c_double),
        ("seff", C.c_double), ("temp_coef_mu", C.c_double),
        ("reserved_d", C.c_double * 6),
    ]

    @property
    def maxvxyzu(self):
        return 3 if self.isothermal else 4

    @property
    def ngradh(self):
        return 2 if self.gravity else 1

    def copy(self):
        q = SphParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(SphParams))
        return q


class SphScalars(C.Structure):
    _fields_ = [
        ("dtcourant", C.c_double), ("dtforce", C.c_double), ("dtmini", C.c_double), ("dtmaxi", C.c_double),
        ("rhomax", C.c_double), ("trialmean", C.c_double), ("actualmean", C.c_double),
        ("maxtrial", C.c_int64), ("maxactual", C.c_int64), ("nrhocalc", C.c_int64), ("nactualtot", C.c_int64),
        ("np", C.c_int64), ("ncalls_neigh", C.c_int64),
        ("npairs_density", C.c_int64), ("npairs_force", C.c_int64), ("nbinmaxnew", C.c_int64),
        ("npairs_gravity", C.c_int64), ("nm2l", C.c_int64), ("reserved", C.c_int64 * 1),
    ]


def default_params(**kw):
    """Reference defaults: options.f90:92-93 (tolh), timestep.f90:52-62 (C_cour, C_force,
    psidecayfac, overcleanfac), shock_capturing.f90:47-64 (alpha..beta), kdtree.F90:46,
    eos.f90:1896-1898, boundary.f90:70-75, kernel_cubic.f90:30 (hfact_default)."""
    p = SphParams()
    p.kernel = KERNEL_CUBIC
    p.periodic = 1
    p.isothermal = 0
    p.ieos = 2
    p.ipdv_heating = p.ishock_heating = p.iresistive_heating = 1
    p.set_boundaries_to_active = 1
    p.xmin = p.ymin = p.zmin = -0.5
    p.xmax = p.ymax = p.zmax = 0.5
    p.hfact = 1.2
    p.tolh = 1.0e-4
    p.alpha, p.alphamax, p.alphau, p.alphaB, p.beta = 0.0, 1.0, 1.0, 1.0, 2.0
    p.polyk, p.gamma, p.qfacdisc, p.cs_min = 1.0, 5.0 / 3.0, 0.75, 0.0
    p.C_cour, p.C_force, p.dtmax = 0.3, 0.25, 1.0e29
    p.psidecayfac, p.overcleanfac = 1.0, 1.0
    p.tree_accuracy = 0.5
    for k, v in kw.items():
        if k == "massoftype":
            for i, m in enumerate(v):
                p.massoftype[i] = m
        else:
            setattr(p, k, v)
    return p


class SphStepOut(C.Structure):
    """sphgpu_step_out (include/sphgpu.h)"""
    _fields_ = [("dtcourant", C.c_double), ("dtforce", C.c_double), ("dterr", C.c_double), ("errmax", C.c_double),
                ("its", C.c_int64), ("scalars", SphScalars)]


class SphEnergies(C.Structure):
    """sphgpu_energies (include/sphgpu.h): energies.f90 module variables ekin, etherm, emag, epot, etot, totmom, angtot, mtot, xyzcom"""
    _fields_ = [(k, C.c_double) for k in ("ekin", "etherm", "emag", "epot", "etot", "totmom", "xmom", "ymom", "zmom", "angtot",
                                          "angx", "angy", "angz", "mtot", "xcom", "ycom", "zcom", "rhomax")] + [("np", C.c_int64)]
