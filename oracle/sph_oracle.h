/*
 * sph_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (C++/OpenMP, IEEE double, no FMA contraction) of the hot
 * path of danieljprice/phantom v2026.0.1:
 *   build_tree      src/main/neigh_kdtree.f90:161  -> maketree src/main/kdtree.F90:117
 *   densityiterate  src/main/dens.F90:117
 *   cons2prim_everything src/main/cons2prim.f90:274
 *   force           src/main/force.F90:193
 *
 * PARITY STATUS: the reference is pure Fortran and no Fortran compiler exists
 * in the build container nor on the GPU box, so the reference binary cannot
 * be run.  The oracle is pinned against every known-answer value the
 * reference's own test-suite holds for this path (tests/test_oracle_*.py cite
 * them), but parity against an actual reference run is "unpinned".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (phantom_b200/) never links or imports it.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAXTYPES 8 /* massoftype(1:7) used: igas=1 iboundary=3 istar=4 idarkmatter=5 ibulge=6 idust=7 (part.F90:428-438) */

/* Field-for-field the same layout as sphgpu_params in include/sphgpu.h
 * (kept as a separate declaration on purpose: the oracle shares no code
 * with the product). */
typedef struct oracle_params {
    /* compile-time flag tuple of the reference (build/Makefile:189-300) */
    int32_t kernel;          /* 0 = M4 cubic (kernel_cubic.f90), 1 = M6 quintic (kernel_quintic.f90) */
    int32_t periodic;        /* -DPERIODIC */
    int32_t isothermal;      /* -DISOTHERMAL -> maxvxyzu = 3 */
    int32_t mhd;             /* -DMHD */
    int32_t gravity;         /* -DGRAVITY */
    int32_t dust;            /* -DDUST (two-fluid, dust as particles) */
    int32_t const_av;        /* -DCONST_AV -> nalpha = 0 */
    int32_t ind_timesteps;   /* -DIND_TIMESTEPS */
    int32_t disc_viscosity;  /* dim:disc_viscosity */
    int32_t ieos;            /* eos.f90: 1 isothermal, 2 adiabatic/polytropic, 3 locally isothermal disc */
    int32_t ipdv_heating, ishock_heating, iresistive_heating; /* eos.f90:1896-1898 */
    int32_t set_boundaries_to_active; /* part.F90:439 */
    int32_t idrag;           /* dust.f90: 1 Epstein/Stokes, 2 const K, 3 const ts */
    int32_t driving;         /* -DDRIVING: turbulent stirring (forcing.f90); force() then ADDS to fxyzu (force.F90:2969-2973) */
    int32_t reserved_i[4];
    /* boundary.f90 */
    double xmin, xmax, ymin, ymax, zmin, zmax;
    /* part / options */
    double hfact, tolh;
    double massoftype[ORACLE_MAXTYPES]; /* index = itype (0 unused) */
    /* shock_capturing.f90:47-64 */
    double alpha, alphamax, alphau, alphaB, beta;
    /* eos */
    double polyk, gamma, qfacdisc, cs_min;
    /* timestep.f90 */
    double C_cour, C_force, dtmax, psidecayfac, overcleanfac;
    /* kdtree.F90:46 */
    double tree_accuracy;
    /* dust.f90 (two-fluid): grain size/density in code units, K_code */
    double grainsize, graindens, K_code;
    double seff;             /* dust.f90:96-99 (init_drag): effective surface density, code units */
    double temp_coef_mu;     /* eos.f90:194: temperature_coef*gmw -> eos_vars(itemp) */
    double reserved_d[6];
} oracle_params;

/* scalars the reference returns through module variables
 * (timestep: dtcourant, dtforce, rhomaxnow; dens.F90:1137 neighbour stats) */
typedef struct oracle_scalars {
    double dtcourant, dtforce, dtmini, dtmaxi, rhomax;
    double trialmean, actualmean;
    int64_t maxtrial, maxactual, nrhocalc, nactualtot, np, ncalls_neigh;
    int64_t npairs_density, npairs_force;  /* real interacting pairs (for roofline accounting) */
    int64_t nbinmaxnew;
    int64_t npairs_gravity, nm2l;
    int64_t reserved[1];
} oracle_scalars;

typedef struct oracle_ctx oracle_ctx;

oracle_ctx *oracle_create(const oracle_params *p);
void oracle_destroy(oracle_ctx *c);
void oracle_set_params(oracle_ctx *c, const oracle_params *p);
void oracle_set_threads(int nthreads);
int  oracle_get_max_threads(void);
const char *oracle_last_error(oracle_ctx *c);

/* build_tree(npart,nactive,xyzh,vxyzu): xyzh(4,npart) is inout (periodic wrap) */
int oracle_build_tree(oracle_ctx *c, int64_t npart, double *xyzh, const int8_t *iphase);

/* tree inspection (for tests) */
int64_t oracle_tree_ncells(oracle_ctx *c);
/* node record: xcen[3], size, hmax, mass, quads[6] -> 12 doubles ; ints: leftchild,rightchild,parent,leaf_is_active,i1,i2 */
int oracle_tree_get_node(oracle_ctx *c, int64_t inode, double *rec12, int32_t *irec6);
int oracle_tree_get_inodeparts(oracle_ctx *c, int32_t *out /* npart */);
/* trial neighbour list of a leaf (get_neighbour_list, neigh_kdtree.f90:218); returns nneigh, fills list (1-based ids) */
int64_t oracle_get_neighbour_list(oracle_ctx *c, int64_t icell, int getj, int32_t *list, int64_t maxlist);

/* densityiterate(icall, ...) dens.F90:117.  Arrays in Fortran layout:
 * xyzh(4,n) vxyzu(maxvxyzu,n) fxyzu(maxvxyzu,n) fext(3,n) Bevol(4,n)
 * divcurlv(1,n) divcurlB(4,n) alphaind(3,n) gradh(ngradh,n) dvdx(9,n) : real*4 */
int oracle_densityiterate(oracle_ctx *c, int icall, int64_t npart, double *xyzh,
                          const double *vxyzu, const double *fxyzu, const double *fext,
                          const double *Bevol, const int8_t *iphase,
                          float *divcurlv, float *divcurlB, float *alphaind, float *gradh,
                          float *dvdx, double *dustfrac, oracle_scalars *out);

/* cons2prim_everything cons2prim.f90:274 ; eos_vars(7,n) */
int oracle_cons2prim(oracle_ctx *c, int64_t npart, const double *xyzh, const double *vxyzu,
                     const float *dvdx, const double *Bevol, const int8_t *iphase,
                     double *eos_vars, float *alphaind, double *Bxyz);

/* force(icall, ...) force.F90:193 */
int oracle_force(oracle_ctx *c, int icall, int64_t npart, const double *xyzh, const double *vxyzu,
                 double *fxyzu, float *divcurlv, const float *divcurlB, const double *Bevol,
                 double *dBevol, const double *fext, const double *eos_vars,
                 const float *alphaind, const float *gradh, const float *dvdx,
                 const int8_t *iphase, const double *dustfrac, double dt,
                 float *poten, float *divBsymm, double *tstop,
                 int8_t *ibin, int8_t *ibin_wake, const int8_t *ibin_old, int nbinmax, int ibinnow, int istepfrac,
                 oracle_scalars *out);

/* exact neighbour sets {j != i : r_ij^2/h_i^2 < radkern^2} (density) or
 * {q2i < R2 or q2j < R2} (force, symmetric=1) found through the tree, CSR; returns total or -needed */
int64_t oracle_neighbour_sets(oracle_ctx *c, int64_t npart, const double *xyzh, const int8_t *iphase,
                              int symmetric, int64_t *offsets /* npart+1 */, int32_t *list, int64_t maxlist);
/* O(N^2) brute force version of the same (test_neigh.f90:264-367) */
int64_t oracle_neighbour_counts_bruteforce(oracle_ctx *c, int64_t npart, const double *xyzh,
                                           int symmetric, int32_t *counts);

/* SPH kernel functions (kernel_cubic.f90 / kernel_quintic.f90) */
void oracle_kernel(int kernel, double q2, double q, double *wkern, double *grkern, double *dphidh,
                   double *potensoft, double *fsoft, double *wdrag);
void oracle_kernel_constants(int kernel, double *radkern, double *cnormk, double *wab0, double *gradh0,
                             double *dphidh0, double *cnormk_drag, double *hfact_default);

/* gravity pieces (kdtree.F90:1702, :1527, :1799) */
void oracle_compute_M2L(double dx, double dy, double dz, double dr, double totmass, const double *quads, double *fnode20);
void oracle_expand_fgrav(const double *fnode20, double dx, double dy, double dz, double *fxyzpot4);
void oracle_propagate_fnode(double *fnode_dst20, const double *fnode_src20, double dx, double dy, double dz);

/* st_calcAccel (forcing.f90:728-830): turbulent stirring acceleration into fxyzu(1:3,:) */
void oracle_forcing(oracle_ctx *c, int64_t npart, const double *xyzh, const int8_t *iphase, double *fxyzu, int nmodes, const double *mode,
                    const double *ampl, const double *aka, const double *akb, double amplfac, double solweightnorm, int correct_mean_force);

/* L'Ecuyer ran2 (random.f90) */
double oracle_ran2(int32_t *iseed);

#ifdef __cplusplus
}
#endif
#endif
