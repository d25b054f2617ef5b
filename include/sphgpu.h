/*
 * sphgpu.h -- C ABI of the B200-native SPH derivative engine.
 *
 * The entry points are what a Fortran `bind(C)` shim binds in place of the
 * reference's hot-path subroutines (danieljprice/phantom v2026.0.1):
 *
 *   sphgpu_build_tree              <-> build_tree            src/main/neigh_kdtree.f90:161-208
 *   sphgpu_densityiterate          <-> densityiterate        src/main/dens.F90:117-152
 *   sphgpu_cons2prim_everything    <-> cons2prim_everything  src/main/cons2prim.f90:274-291
 *   sphgpu_force                   <-> force                 src/main/force.F90:193-263
 *   sphgpu_derivs                  <-> derivs                src/main/deriv.f90:37-232  (the four above in sequence)
 *   sphgpu_get_neighbour_stats     <-> get_neighbour_stats   src/main/dens.F90:1137-1155
 *
 * All array arguments are HOST pointers in the reference's Fortran
 * column-major layout (xyzh(4,n) = n records of 4 doubles, ...), 1-based
 * particle identity = position in the array.  "Literal" calls copy in,
 * compute on the device and copy out, exactly like the Fortran argument
 * lists; between sphgpu_upload and sphgpu_download the state is resident
 * on the device and the *_resident calls touch no host memory.
 * Everything the reference reads from module variables / cpp flags is passed
 * explicitly in sphgpu_params.
 *
 * Return value: 0 on success, non-zero error code otherwise; the message is
 * returned by sphgpu_last_error (the shim maps it to `call fatal('gpu',msg)`,
 * src/main/io.F90:525-560).  One host thread per context; not re-entrant.
 */
#ifndef SPHGPU_H
#define SPHGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHGPU_MAXTYPES 8

/* error codes */
enum {
    SPHGPU_OK = 0,
    SPHGPU_ERR_CUDA = 1,          /* CUDA runtime failure (also: no device) */
    SPHGPU_ERR_ARG = 2,           /* bad argument */
    SPHGPU_ERR_NAN = 3,           /* NaN in particle position           (kdtree.F90:391) */
    SPHGPU_ERR_NOPART = 4,        /* no live particles                  (kdtree.F90:162-164) */
    SPHGPU_ERR_NOCONVERGE = 5,    /* density iteration failed           (dens.F90:1443-1455) */
    SPHGPU_ERR_NEGH = 6,          /* h < 0 in force                     (force.F90:2271) */
    SPHGPU_ERR_OVERFLOW = 7,      /* internal neighbour scratch overflow */
    SPHGPU_ERR_STATE = 8          /* call sequence violated (e.g. density before tree) */
};

/* cpp-flag tuple + module variables of the reference the path reads implicitly */
typedef struct sphgpu_params {
    int32_t kernel;          /* 0 cubic (kernel_cubic.f90) 1 quintic (kernel_quintic.f90); build/Makefile:283-290 */
    int32_t periodic;        /* -DPERIODIC */
    int32_t isothermal;      /* -DISOTHERMAL: vxyzu/fxyzu have 3 rows instead of 4 */
    int32_t mhd;             /* -DMHD */
    int32_t gravity;         /* -DGRAVITY */
    int32_t dust;            /* -DDUST */
    int32_t const_av;        /* -DCONST_AV */
    int32_t ind_timesteps;   /* -DIND_TIMESTEPS */
    int32_t disc_viscosity;  /* dim: disc_viscosity */
    int32_t ieos;            /* eos.f90:183-256: 1, 2, 3 */
    int32_t ipdv_heating, ishock_heating, iresistive_heating; /* eos.f90:1896-1898 */
    int32_t set_boundaries_to_active;                          /* part.F90:439 */
    int32_t idrag;
    int32_t driving;         /* -DDRIVING: turbulent stirring (forcing.f90); force() then ADDS to fxyzu (force.F90:2969-2973) */
    int32_t reserved_i[4];
    double xmin, xmax, ymin, ymax, zmin, zmax;                 /* boundary.f90 */
    double hfact, tolh;                                        /* part.F90 hfact, options.f90:92 tolh */
    double massoftype[SPHGPU_MAXTYPES];                        /* part.F90 massoftype(itype), index 0 unused */
    double alpha, alphamax, alphau, alphaB, beta;              /* shock_capturing.f90:47-64 */
    double polyk, gamma, qfacdisc, cs_min;                     /* eos.f90 */
    double C_cour, C_force, dtmax, psidecayfac, overcleanfac;  /* timestep.f90:52-62 */
    double tree_accuracy;                                      /* kdtree.F90:46 */
    double grainsize, graindens, K_code;                       /* dust.f90: grain size / intrinsic density (code units), K_code(1) */
    double seff;                                               /* dust.f90:96-99 init_drag: effective surface density for the mean free path (code units) */
    double temp_coef_mu;                                       /* eos.f90:194,239,255: temperature_coef*gmw, eos_vars(itemp) = temp_coef_mu * P/rho ; 0 stores 0 */
    double reserved_d[6];
} sphgpu_params;

/* module-variable outputs: timestep:dtcourant,dtforce,rhomaxnow ; dens.F90:104-106 statistics */
typedef struct sphgpu_scalars {
    double dtcourant, dtforce, dtmini, dtmaxi, rhomax;
    double trialmean, actualmean;
    int64_t maxtrial, maxactual, nrhocalc, nactualtot, np, ncalls_neigh;
    int64_t npairs_density, npairs_force;   /* real interacting pairs evaluated (roofline accounting) */
    int64_t nbinmaxnew;
    int64_t npairs_gravity, nm2l;           /* Newtonian P2P pairs outside both kernels; accepted node-node M2L evaluations */
    int64_t reserved[1];
} sphgpu_scalars;

/* host array bundle for upload/download (any pointer may be NULL = skip) */
typedef struct sphgpu_host_arrays {
    int64_t npart;
    double *xyzh;        /* (4,n)          */
    double *vxyzu;       /* (maxvxyzu,n)   */
    double *fxyzu;       /* (maxvxyzu,n)   */
    double *fext;        /* (3,n)          */
    double *Bevol;       /* (4,n)          */
    double *dBevol;      /* (4,n)          */
    double *eos_vars;    /* (7,n)          */
    float *divcurlv;     /* (1,n)          */
    float *divcurlB;     /* (4,n)          */
    float *alphaind;     /* (3,n)          */
    float *gradh;        /* (ngradh,n)     */
    float *dvdx;         /* (9,n)          */
    float *poten;        /* (n)            */
    float *divBsymm;     /* (n)            */
    int8_t *iphase;      /* (n)            */
    int8_t *ibin, *ibin_old, *ibin_wake;   /* (n) timestep_ind / part.F90 */
    double *dustfrac;    /* (n)  two-fluid dust-to-gas ratio seen by gas particles (dens.F90:1612-1624) */
    double *tstop;       /* (n)  stopping time (force.F90:3240) */
} sphgpu_host_arrays;

/* field mask bits for upload/download */
#define SPHGPU_F_XYZH      (1ull << 0)
#define SPHGPU_F_VXYZU     (1ull << 1)
#define SPHGPU_F_FXYZU     (1ull << 2)
#define SPHGPU_F_FEXT      (1ull << 3)
#define SPHGPU_F_BEVOL     (1ull << 4)
#define SPHGPU_F_DBEVOL    (1ull << 5)
#define SPHGPU_F_EOSVARS   (1ull << 6)
#define SPHGPU_F_DIVCURLV  (1ull << 7)
#define SPHGPU_F_DIVCURLB  (1ull << 8)
#define SPHGPU_F_ALPHAIND  (1ull << 9)
#define SPHGPU_F_GRADH     (1ull << 10)
#define SPHGPU_F_DVDX      (1ull << 11)
#define SPHGPU_F_POTEN     (1ull << 12)
#define SPHGPU_F_DIVBSYMM  (1ull << 13)
#define SPHGPU_F_IPHASE    (1ull << 14)
#define SPHGPU_F_IBIN      (1ull << 15)
#define SPHGPU_F_DUSTFRAC  (1ull << 16)
#define SPHGPU_F_TSTOP     (1ull << 17)
#define SPHGPU_F_ALL       (~0ull)

typedef struct sphgpu_ctx sphgpu_ctx;

int  sphgpu_create(const sphgpu_params *params, int device, sphgpu_ctx **out);
void sphgpu_destroy(sphgpu_ctx *ctx);
int  sphgpu_set_params(sphgpu_ctx *ctx, const sphgpu_params *params);
const char *sphgpu_last_error(sphgpu_ctx *ctx);
/* Options by name; returns 0 if known.  Tuning: "max_cell" (targets per group, <= 32), "max_leaf" (particles per leaf cell),
 * "group_pack", "list_margin", "hilbert", "scratch_per_warp", "always_refit", "grav_p2p_per_particle".
 * Behaviour: "refcompat_hmax" (1/0; default = on with individual timesteps: neighbour sets pruned exactly as the reference's walk
 * prunes them while some particle is inactive), "force_general" (1 = general pair kernels even for an all-gas set), "no_iso1"
 * (1 = keep the three-sector force records for an isothermal set instead of deriving P, rho and c_s from h per pair).
 * "legacy_stream" = 1 replaces the context's non-blocking compute stream by a BLOCKING one, i.e. one implicitly ordered with the
 * legacy default stream: a host that issues its NCCL collectives relative to the default stream (torch.distributed, or a Fortran
 * MPI+NCCL driver doing the same) can then chain sphgpu_halo_select -> _pack -> all-to-all -> _unpack without any host
 * synchronisation; the halo entry points skip their cudaStreamSynchronize in that mode. */
int  sphgpu_set_option(sphgpu_ctx *ctx, const char *name, double value);
/* per-phase device times of the last derivs (ms): tree, dens, cons2prim, force ; utils_timing.f90 labels */
int  sphgpu_get_timings(sphgpu_ctx *ctx, double *ms4);
/* device time (ms) of the two dominant kernels of the last call: [0] density pair kernel, [1] force pair kernel */
int  sphgpu_get_kernel_timings(sphgpu_ctx *ctx, double *ms2);
/* device time (ms) of the last self-gravity pass: [0] whole pass (tree + FMM walk + P2P), [1] the P2P kernel alone */
int  sphgpu_get_gravity_timings(sphgpu_ctx *ctx, double *ms2);
/* node records of the self-gravity tree (maketree with -DGRAVITY, kdtree.F90:531-929; kdnode, dtype_kdtree.F90:53-68) after the last force call:
 * rec12 = {xcen[3], size, hmax, mass, quads[6]}, irec6 = {leftchild, rightchild, parent, first slot, count, level} (0-based, -1 = none),
 * ids = inodeparts (1-based particle ids by slot).  rec12 == NULL: returns only the number of nodes; -1 when no tree exists */
int64_t sphgpu_gravity_tree(sphgpu_ctx *ctx, int64_t maxnodes, double *rec12, int32_t *irec6, int32_t *ids);
/* number of kernel launches issued by this context since creation */
int64_t sphgpu_launch_count(sphgpu_ctx *ctx);

/* ---- resident mode ------------------------------------------------------------------- */
int sphgpu_upload(sphgpu_ctx *ctx, const sphgpu_host_arrays *h, uint64_t mask);
int sphgpu_download(sphgpu_ctx *ctx, sphgpu_host_arrays *h, uint64_t mask);
int sphgpu_build_tree_resident(sphgpu_ctx *ctx);
int sphgpu_densityiterate_resident(sphgpu_ctx *ctx, int icall, sphgpu_scalars *out);
int sphgpu_cons2prim_resident(sphgpu_ctx *ctx);
int sphgpu_force_resident(sphgpu_ctx *ctx, int icall, double dt, sphgpu_scalars *out);
int sphgpu_derivs_resident(sphgpu_ctx *ctx, int icall, double dt, sphgpu_scalars *out);

/* ---- literal mode: the reference's argument lists, host pointers ------------------------ */
/* build_tree(npart,nactive,xyzh,vxyzu)  [+ hidden input iphase]; xyzh is inout (periodic wrap) */
int sphgpu_build_tree(sphgpu_ctx *ctx, int64_t npart, int64_t nactive, double *xyzh, const double *vxyzu,
                      const int8_t *iphase);
/* densityiterate(icall,npart,nactive,xyzh,vxyzu,divcurlv,divcurlB,Bevol,stressmax,fxyzu,fext,alphaind,gradh,rad,radprop,dvdx,apr_level) */
int sphgpu_densityiterate(sphgpu_ctx *ctx, int icall, int64_t npart, int64_t nactive, double *xyzh, const double *vxyzu,
                          float *divcurlv, float *divcurlB, const double *Bevol, double *stressmax,
                          const double *fxyzu, const double *fext, float *alphaind, float *gradh, float *dvdx,
                          const int8_t *iphase, sphgpu_scalars *out);
/* cons2prim_everything(npart,xyzh,vxyzu,dvdx,rad,eos_vars,radprop,Bevol,Bxyz,dustevol,dustfrac,alphaind) */
int sphgpu_cons2prim_everything(sphgpu_ctx *ctx, int64_t npart, const double *xyzh, const double *vxyzu,
                                const float *dvdx, double *eos_vars, const double *Bevol, double *Bxyz,
                                float *alphaind, const int8_t *iphase);
/* force(icall,npart,xyzh,vxyzu,fxyzu,divcurlv,divcurlB,Bevol,dBevol,...,fext,...,dt,stressmax,eos_vars,...)  (force.F90:193-196).
 * Not carried: dustfrac (device-resident between the passes; host copy through sphgpu_host_arrays), fxyz_drag (implicit drag only),
 * ddustevol/dustprop/dustgasprop/Vrel_disp (one-fluid dust, growth), ipart_rhomax (sink creation), rad/drad/radprop/dens/metrics/apr_level:
 * all outside the scope of this path (INTEGRATION.md section 4 lists each). */
int sphgpu_force(sphgpu_ctx *ctx, int icall, int64_t npart, const double *xyzh, const double *vxyzu, double *fxyzu,
                 float *divcurlv, const float *divcurlB, const double *Bevol, double *dBevol, const double *fext,
                 double dt, double stressmax, const double *eos_vars, const float *alphaind, const float *gradh,
                 const float *dvdx, const int8_t *iphase, float *poten, float *divBsymm, sphgpu_scalars *out);
/* derivs(icall,...): upload -> tree -> density -> cons2prim -> force -> download, all arrays of the bundle */
int sphgpu_derivs(sphgpu_ctx *ctx, int icall, sphgpu_host_arrays *h, double dt, sphgpu_scalars *out);

/* bytes the last sphgpu_derivs copied host->device and device->host (arrays every particle overwrites are not uploaded when all are active) */
int sphgpu_get_copy_bytes(sphgpu_ctx *ctx, int64_t *h2d, int64_t *d2h);
/* get_neighbour_stats(trialmean,actualmean,maxtrial,maxactual,nrhocalc,nactualtot) of the last density call */
int sphgpu_get_neighbour_stats(sphgpu_ctx *ctx, sphgpu_scalars *out);
/* exact neighbour sets of the last tree/density state in CSR form (1-based ids, sorted), for parity tests;
 * symmetric=0: {j!=i : q2i < radkern2}; symmetric=1: {q2i < radkern2 or q2j < radkern2}.
 * returns total count, or -(needed) when maxlist is too small */
int64_t sphgpu_neighbour_sets(sphgpu_ctx *ctx, int symmetric, int64_t *offsets, int32_t *list, int64_t maxlist);

/* individual timesteps (-DIND_TIMESTEPS): module timestep_ind inputs nbinmax, ibinnow, istepfrac (utils_indtimesteps.f90);
 * the force pass then updates ibin (get_newbin + Saitoh-Makino limiter, force.F90:3272-3310), flags ibin_wake of every
 * neighbour of an active particle (force.F90:1346-1358) and returns nbinmaxnew in sphgpu_scalars */
int sphgpu_set_timestep_bins(sphgpu_ctx *ctx, int nbinmax, int ibinnow, int istepfrac);

/* ---- multi-GPU halo (one context per rank / GPU): replaces the MPI cell export of mpi_dens.F90 / mpi_force.F90 /
 * mpi_derivs.F90:197-522.  Ghost particles are appended after the owned ones as inactive (neighbour-only) particles.
 * The host side moves the packed buffers with NCCL (all-to-all-v) directly between the returned DEVICE pointers. */
int sphgpu_local_hmax(sphgpu_ctx *ctx, double *hmax);
/* halo sufficiency (the reference re-exports cells when h grows, mpi_dens.F90 / dens.F90:343-365): largest trial h of the last
 * density pass; when radkern * (global max of it) exceeds the halo width used, restore the pre-density h and repeat with a wider halo */
int sphgpu_density_hmax_used(sphgpu_ctx *ctx, double *hmax);
int sphgpu_halo_restore_h(sphgpu_ctx *ctx);
/* largest h_new/h_old of the last density pass on this rank (reduced over ranks by the driver, handed back as option "halo_hgrow") */
int sphgpu_density_hgrow(sphgpu_ctx *ctx, double *hgrow);
/* boxes = 6 doubles per rank {lo xyz, hi xyz}; counts[r] = owned particles within dhalo of rank r's box (minimum image) */
int sphgpu_halo_select(sphgpu_ctx *ctx, int nranks, int myrank, const double *boxes, double dhalo, int64_t *counts);
/* stage 1 (before build_tree): 16 doubles/ghost; stage 2 (after densityiterate): 4 doubles/ghost {h, gradh, alpha, gradsoft} */
int sphgpu_halo_pack(sphgpu_ctx *ctx, int stage, void **sendptr_device, int *record_doubles);
int sphgpu_halo_recvbuf(sphgpu_ctx *ctx, int64_t nrecords, int record_doubles, void **recvptr_device);
int sphgpu_halo_unpack(sphgpu_ctx *ctx, int stage, int64_t nghost);
int64_t sphgpu_nghost(sphgpu_ctx *ctx);
/* self-gravity across GPUs (replaces maketreeglobal + the remote cell export for the gravity terms, kdtree.F90:2044-2300, mpi_force.F90):
 * after the density pass every rank packs 13 doubles per OWNED particle {x,y,z,h, iphase, h at build_tree, iterations, h history(6)},
 * the ranks all-gather them (NCCL, padded to `stride` records per rank) into the buffer returned by _recvbuf, and _unpack makes the
 * gathered set the input of the gravity pass of the next force call: every rank builds the same tree of the whole set and evaluates
 * walk / M2L / P2P only for nodes holding its own particles, so the result equals the single-GPU result to round-off. */
int sphgpu_gravity_gather_pack(sphgpu_ctx *ctx, void **sendptr_device, int *record_doubles);
int sphgpu_gravity_gather_recvbuf(sphgpu_ctx *ctx, int nranks, int64_t stride, void **recvptr_device);
int sphgpu_gravity_gather_unpack(sphgpu_ctx *ctx, int nranks, int myrank, int64_t stride, const int64_t *counts);

/* ---- the multi-GPU path behind the C ABI (dist.cu): one context per rank / GPU, NCCL over NVLink.  A Fortran host replaces
 * balancedomains (mpi_balance.F90:82), the MPI cell export (mpi_derivs.F90:197-522) and reduceall_mpi with these calls; it only has
 * to broadcast the 128-byte NCCL id that rank 0 obtains from sphgpu_dist_get_unique_id (MPI_Bcast, or a file).  NCCL is loaded with
 * dlopen at sphgpu_dist_init, so a single-GPU user needs none.  Sequence: upload the OWNED particles -> sphgpu_dist_init ->
 * sphgpu_dist_set_ids -> sphgpu_dist_set_boxes (or sphgpu_dist_rebalance) -> sphgpu_dist_derivs / sphgpu_dist_step ... -> download. */
int sphgpu_dist_get_unique_id(void *id, int nbytes);                       /* nbytes >= 128 */
int sphgpu_dist_init(sphgpu_ctx *ctx, const void *id, int nranks, int rank);
int sphgpu_dist_finalize(sphgpu_ctx *ctx);
/* boxes = 6 doubles per rank {lo xyz, hi xyz}, tiling the periodic box (or the bounding box of the set) */
int sphgpu_dist_set_boxes(sphgpu_ctx *ctx, const double *boxes);
int sphgpu_dist_get_boxes(sphgpu_ctx *ctx, double *boxes);
/* global identities of the owned particles (they travel with a particle when it migrates); ids == NULL numbers them base, base+1, ... */
int sphgpu_dist_set_ids(sphgpu_ctx *ctx, const int64_t *ids, int64_t base);
int sphgpu_dist_get_ids(sphgpu_ctx *ctx, int64_t *ids, int64_t maxn);
int64_t sphgpu_dist_nlocal(sphgpu_ctx *ctx);
/* derivs(icall = 1 | 2) on the decomposed set; the scalars are reduced over the ranks (dtcourant, dtforce: min; rhomax: max; counts: sum) */
int sphgpu_dist_derivs(sphgpu_ctx *ctx, int icall, double dt, sphgpu_scalars *out);
/* balancedomains: hand every owned particle that left this rank's box to its new owner (whole record); returns the new owned count.
 * in_step = 1 inside a leapfrog step (the evolved v, B travel too); a caller between steps passes 0 */
int sphgpu_dist_migrate(sphgpu_ctx *ctx, int in_step, int64_t *nlocal_new);
/* the reference's domain split (kdtree.F90:2098-2160, applied globally): bisection at the centre of mass along the longest axis,
 * log2(nranks) levels, moments all-reduced; followed by a migration.  domain = {lo xyz, hi xyz} */
int sphgpu_dist_rebalance(sphgpu_ctx *ctx, const double *domain, int64_t *nlocal_new);
/* out8: [0] ghost capacity, [1] ghosts received, [2] bytes sent + received, [3] exchange rounds of the last dist_derivs,
 * [4] particles handed over by the last migration, [5] largest trial h (halo width = radkern x this x 1.15 + overhang),
 * [6] device time of the last dist_derivs in ms (CUDA events on the stream that also carries the NCCL transfers), [7] owned particles */
int sphgpu_dist_stats(sphgpu_ctx *ctx, double *out8);

/* ---- the callers either side of the path, resident on the device -------------------------------- */
/* step (src/main/step_leapfrog.f90:95-760) with global timesteps and substep_sph (substepping.F90:241-264):
 * predictor, drift, predict_sph (h prediction, alpha decay), derivs(1), corrector iterated with derivs(2) until the velocity
 * error is below tolv (timestep.f90 tolv = 1e-2).  dterr as check_velocity_error returns it (:769-843). */
typedef struct sphgpu_step_out {
    double dtcourant, dtforce, dterr, errmax;
    int64_t its;
    sphgpu_scalars scalars;     /* of the last derivs call of the step */
} sphgpu_step_out;
int sphgpu_step_resident(sphgpu_ctx *ctx, double dtsph, double tolv, sphgpu_step_out *out);
/* the same step with -DIND_TIMESTEPS (step_leapfrog.f90:57-80, :157-164, :183-235, :470-560; utils_indtimesteps.f90): the host keeps
 * time, istepfrac and nbinmax as evolve.f90 does and calls, per smallest timestep dtsph = dtmax / 2^nbinmax,
 *   sphgpu_set_active_particles_resident (set_active_particles: the activity flags of this substep, ibinnow) and
 *   sphgpu_step_ind_resident (predictor to every particle's own half step twas, drift, predict_sph, derivs, corrector of the active
 *   particles into their new bins, synchronisation, wake-up of flagged neighbours); out->scalars.nbinmaxnew is the new nbinmax.
 * sphgpu_init_step_resident is init_step: at time 0 every particle starts in bin nbinmax, twas = time + dt(ibin)/2. */
int sphgpu_init_step_resident(sphgpu_ctx *ctx, double time, double dtmax, int nbinmax);
int sphgpu_set_active_particles_resident(sphgpu_ctx *ctx, int nbinmax, int istepfrac, int64_t *nactive, int64_t *nalive);
int sphgpu_step_ind_resident(sphgpu_ctx *ctx, double t, double dtsph, double dtmax, sphgpu_step_out *out);

/* compute_energies (src/main/energies.f90:64-760): the conserved-quantity sums of the resident state */
typedef struct sphgpu_energies {
    double ekin, etherm, emag, epot, etot;
    double totmom, xmom, ymom, zmom, angtot, angx, angy, angz;
    double mtot, xcom, ycom, zcom, rhomax;
    int64_t np;
} sphgpu_energies;
int sphgpu_energies_resident(sphgpu_ctx *ctx, sphgpu_energies *out);
/* sphgpu_step_resident on the decomposed set: migration, then the leapfrog step with ghost exchanges inside every derivs and the
 * velocity-error norm reduced over the ranks */
int sphgpu_dist_step(sphgpu_ctx *ctx, double dtsph, double tolv, sphgpu_step_out *out);
/* compute_energies reduced over the ranks */
int sphgpu_dist_energies(sphgpu_ctx *ctx, sphgpu_energies *out);

/* turbulent driving: st_calcAccel (src/main/forcing.f90:728-830), called by derivs before force when
 * -DDRIVING (deriv.f90:178-182).  The host keeps the Ornstein-Uhlenbeck phases (st_ounoiseupdate / st_calcPhases use the Fortran
 * random_number stream) and hands the current mode set over whenever it changes: mode(3,n) wave vectors, ampl(n), aka(3,n), akb(3,n),
 * st_amplfac, st_solweightnorm, correct_mean_force.  With params.driving = 1 every derivs call evaluates
 * f_i = 2 amplfac solweightnorm sum_m ampl_m (aka_m cos(k_m.x_i) - akb_m sin(k_m.x_i)) into fxyzu before the force pass adds to it. */
int sphgpu_set_forcing_modes(sphgpu_ctx *ctx, int nmodes, const double *mode, const double *ampl, const double *aka, const double *akb,
                             double amplfac, double solweightnorm, int correct_mean_force);
/* st_calcAccel alone on the resident state (fxyzu(1:3,:) of the active particles is overwritten) */
int sphgpu_forcing_resident(sphgpu_ctx *ctx);

/* register-resident DFMA microbenchmark: returns measured FP64 TFLOP/s of this device (roofline denominator) */
double sphgpu_measure_fp64_peak(sphgpu_ctx *ctx);
/* device copy bandwidth GB/s (read+write) */
double sphgpu_measure_copy_bw(sphgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
