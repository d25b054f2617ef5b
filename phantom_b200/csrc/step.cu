// step.cu -- the callers either side of the hot path, kept on the device so that a run never moves particle data over PCIe
// between steps:
//   * leapfrog step with global timesteps: step (src/main/step_leapfrog.f90:95-760) -- velocity predictor (:183-235), drift of
//     substep_sph (substepping.F90:241-264), predict_sph with the h prediction and the Cullen-Dehnen alpha decay (:307-400),
//     derivs, corrector with the velocity-error iteration (:482-623, check_velocity_error :769-843);
//   * conserved-quantity diagnostics: compute_energies (src/main/energies.f90:64-760; ekin, etherm, emag, epot, linear and
//     angular momentum, centre of mass).
// All of it is O(N) streaming over the caller-ordered arrays (HBM-bound, ~250 B per particle and step against the ~20 kflop of
// the derivative evaluation); the kernels are plain grid-stride AXPYs, the reductions one block sum + one atomic per block.
// External forces and sink particles (substep with fext, step_leapfrog.f90:280-284) and individual timesteps are outside this
// routine: the host integrates those and calls sphgpu_derivs.
#include "common.cuh"
#include <float.h>
#include <string.h>
#include <vector>

namespace {

struct StepArgs {
    int64_t n; int nvu, mhd, nalpha, multitype;
    double *xyzh, *v, *vpred, *f, *B, *Bpred, *dB, *eos_vars; float *divcurlv, *alphaind; const int8_t *iphase;
    double hfact, dt, hdt; double massoftype[SPHGPU_MAXTYPES];
    double *red;       // [0] errmax (as ordered bits), [1] v2mean sum, [2] np
    // individual timesteps (step_leapfrog.f90 with -DIND_TIMESTEPS): every particle sits at its own half step twas(i)
    int ind; double timei; double *twas; int8_t *ibin, *ibin_old, *ibin_wake; int8_t *iphase_w;
    int nbinmax; double thdt[32], ttwas[32];      // ibin_dts(ithdt,:) and ibin_dts(ittwas,:) (:74-79, :157-164)
};

__device__ __forceinline__ bool dead(double h) { return h < DBL_MIN; }     // isdead_or_accreted (part.F90:931)

// velocity predictor (step_leapfrog.f90:183-235): v, u, B/rho, psi to the half step with the "slow" forces
__global__ void k_predict(const StepArgs a)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        if (dead(a.xyzh[4 * i + 3])) continue;
        const int itype = abs((int)a.iphase[i]);
        if (a.ind && a.iphase[i] > 0) a.ibin_old[i] = a.ibin[i];                    // only required for ibin_neigh in force (:195)
        if (itype == IBOUNDARY) continue;
        const double hdt = a.ind ? a.twas[i] - a.timei : a.hdt;                      // :199 synchronise to the particle's own half step
        for (int k = 0; k < a.nvu; k++) a.v[a.nvu * i + k] += hdt * a.f[a.nvu * i + k];
        if (a.mhd && itype == IGAS) for (int k = 0; k < 4; k++) a.B[4 * i + k] += hdt * a.dB[4 * i + k];
    }
}

// substep_sph (substepping.F90:241-264): main position update
__global__ void k_drift(const StepArgs a)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        if (dead(a.xyzh[4 * i + 3])) continue;
        for (int k = 0; k < 3; k++) a.xyzh[4 * i + k] += a.dt * a.v[a.nvu * i + k];
    }
}

// predict_sph (step_leapfrog.f90:307-400): h prediction, v/u/B to the full step for the force evaluation, alpha decay
__global__ void k_predict_sph(const StepArgs a)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double h = a.xyzh[4 * i + 3];
        if (dead(h)) continue;
        const int itype = abs((int)a.iphase[i]);
        if (itype == IBOUNDARY) {
            for (int k = 0; k < a.nvu; k++) a.vpred[a.nvu * i + k] = a.v[a.nvu * i + k];
            if (a.mhd) for (int k = 0; k < 4; k++) a.Bpred[4 * i + k] = a.B[4 * i + k];
            continue;
        }
        const double pmassi = a.massoftype[a.multitype ? itype : IGAS];
        const double rhoi = rhoh_d(h, pmassi, a.hfact);
        const double dhdrhoi = -h / (3. * rhoi);                                        // part.F90:791
        const double hnew = h - a.dt * dhdrhoi * rhoi * (double)a.divcurlv[i];           // :332
        a.xyzh[4 * i + 3] = hnew;
        const double hdt = a.ind ? a.timei - a.twas[i] : a.hdt;                      // :341 interpolate to the end time
        for (int k = 0; k < a.nvu; k++) a.vpred[a.nvu * i + k] = a.v[a.nvu * i + k] + hdt * a.f[a.nvu * i + k];
        if (a.mhd) {
            if (itype == IGAS) for (int k = 0; k < 4; k++) a.Bpred[4 * i + k] = a.B[4 * i + k] + hdt * a.dB[4 * i + k];
            else for (int k = 0; k < 4; k++) a.Bpred[4 * i + k] = a.B[4 * i + k];
        }
        if (a.nalpha >= 2) {                                                             // Cullen & Dehnen (2010) switch, :378-389
            const double spsoundi = a.eos_vars[7 * i + 1];
            const double tdecay1 = 0.1 * spsoundi / hnew;                                // avdecayconst (shock_capturing.f90:32)
            const double ddenom = 1. / (1. + a.dt * tdecay1);
            const double alphaloci = (double)a.alphaind[3 * i + 1];
            const float a1 = a.alphaind[3 * i];
            if ((double)a1 < alphaloci) a.alphaind[3 * i] = (float)alphaloci;
            else a.alphaind[3 * i] = (float)(((double)a1 + a.dt * alphaloci * tdecay1) * ddenom);
        }
    }
}

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.;
    if (threadIdx.x < 32) { r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.; r = warp_sum(r); }
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_max(double v, double *sh)
{
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.;
    if (threadIdx.x < 32) { r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.; r = warp_max(r); }
    __syncthreads();
    return r;
}

// corrector (step_leapfrog.f90:482-623, global timesteps): v to the full step, error against the predicted v
__global__ void k_correct(const StepArgs a)
{
    __shared__ double sh[32];
    double errmax = 0., v2sum = 0., np = 0.;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        if (dead(a.xyzh[4 * i + 3])) continue;
        const int itype = abs((int)a.iphase[i]);
        if (itype == IBOUNDARY) continue;
        const double vx = a.v[a.nvu * i] + a.hdt * a.f[a.nvu * i], vy = a.v[a.nvu * i + 1] + a.hdt * a.f[a.nvu * i + 1],
                     vz = a.v[a.nvu * i + 2] + a.hdt * a.f[a.nvu * i + 2];
        const double ex = vx - a.vpred[a.nvu * i], ey = vy - a.vpred[a.nvu * i + 1], ez = vz - a.vpred[a.nvu * i + 2];
        errmax = fmax(errmax, ex * ex + ey * ey + ez * ez);
        v2sum += vx * vx + vy * vy + vz * vz; np += 1.;
        a.v[a.nvu * i] = vx; a.v[a.nvu * i + 1] = vy; a.v[a.nvu * i + 2] = vz;
        if (a.nvu >= 4) a.v[a.nvu * i + 3] += a.hdt * a.f[a.nvu * i + 3];
        if (a.mhd && itype == IGAS) for (int k = 0; k < 4; k++) a.B[4 * i + k] += a.hdt * a.dB[4 * i + k];
    }
    errmax = block_max(errmax, sh); v2sum = block_sum(v2sum, sh); np = block_sum(np, sh);
    if (threadIdx.x == 0) { atomic_max_pos(&a.red[0], errmax); atomicAdd(&a.red[1], v2sum); atomicAdd(&a.red[2], np); }
}

// corrector with individual timesteps (step_leapfrog.f90:470-560): active particles finish their step and move to the half step of
// their NEW bin, everybody is synchronised to the current time, flagged neighbours are woken into the bin of the particle that woke them
__global__ void k_correct_ind(const StepArgs a)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        if (dead(a.xyzh[4 * i + 3])) continue;
        const int itype = abs((int)a.iphase[i]);
        if (itype == IBOUNDARY) continue;
        double tw = a.twas[i];
        if (a.iphase[i] > 0) {
            a.ibin_wake[i] = 0;                                                      // cannot wake active particles
            const double dti = (a.timei - tw) + a.thdt[a.ibin[i]];
            for (int k = 0; k < a.nvu; k++) a.v[a.nvu * i + k] += dti * a.f[a.nvu * i + k];
            if (a.mhd && itype == IGAS) for (int k = 0; k < 4; k++) a.B[4 * i + k] += dti * a.dB[4 * i + k];
            tw += dti;
        }
        const double hdti = a.timei - tw;                                            // synchronise all particles
        for (int k = 0; k < a.nvu; k++) a.v[a.nvu * i + k] += hdti * a.f[a.nvu * i + k];
        if (a.mhd && itype == IGAS) for (int k = 0; k < 4; k++) a.B[4 * i + k] += hdti * a.dB[4 * i + k];
        if (a.ibin_wake[i] > a.ibin[i]) {                                            // wake inactive particles for the next step
            const int w = min(a.nbinmax, (int)a.ibin_wake[i]);
            tw = a.ttwas[w];
            a.ibin[i] = (int8_t)w;
            a.ibin_wake[i] = 0;
        }
        a.twas[i] = tw;
    }
}

// set_active_particles (utils_indtimesteps.f90:114-178): iphase carries the activity flag of this (sub)step
__global__ void k_set_active(int64_t n, const double *__restrict__ xyzh, int8_t *__restrict__ iphase, int8_t *__restrict__ ibin, int nbinmax, int istepfrac,
                             unsigned long long *cnt)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool alive = false, act = false;
    if (i < n && !dead(xyzh[4 * i + 3])) {
        alive = true;
        const int itype = abs((int)iphase[i]);
        if (itype == IBOUNDARY) ibin[i] = 0;                                        // boundary particles are never active
        const int b = ibin[i];
        if (b > nbinmax) atomicMax(&cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_STATE);
        act = b <= nbinmax && (istepfrac % (1 << (nbinmax - b))) == 0;
        iphase[i] = (int8_t)(act ? itype : -itype);
    }
    const unsigned ma = __ballot_sync(FULLMASK, act), ml = __ballot_sync(FULLMASK, alive);
    if (lane_id() == 0) { if (ma) atomicAdd(&cnt[2], (unsigned long long)__popc(ma)); if (ml) atomicAdd(&cnt[3], (unsigned long long)__popc(ml)); }
}
__global__ void k_init_step(int64_t n, const int8_t *__restrict__ iphase, int8_t *__restrict__ ibin, double *__restrict__ twas, double time, double dtmax,
                            int nbinmax, int reset)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (reset) ibin[i] = (abs((int)iphase[i]) == IBOUNDARY) ? (int8_t)0 : (int8_t)nbinmax;       // step_leapfrog.f90:58-64
    twas[i] = time + 0.5 * dtmax / (double)(1 << ibin[i]);                                        // :69-72
}

// not converged (step_leapfrog.f90:651-701): the new v becomes the prediction, v goes back to the half step
__global__ void k_unconverged(const StepArgs a)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const int itype = abs((int)a.iphase[i]);
        if (itype == IBOUNDARY) continue;
        for (int k = 0; k < a.nvu; k++) { const double v = a.v[a.nvu * i + k]; a.vpred[a.nvu * i + k] = v; a.v[a.nvu * i + k] = v - a.hdt * a.f[a.nvu * i + k]; }
        if (a.mhd) for (int k = 0; k < 4; k++) {
            const double b = a.B[4 * i + k]; a.Bpred[4 * i + k] = b;
            if (itype == IGAS) a.B[4 * i + k] = b - a.hdt * a.dB[4 * i + k];
        }
    }
}

// compute_energies (energies.f90:205-706), the sums this path defines
struct EnArgs {
    int64_t n; int nvu, mhd, gravity, ieos, multitype;
    const double *xyzh, *v, *B, *eos_vars; const float *poten; const int8_t *iphase;
    double hfact, gamma; double massoftype[SPHGPU_MAXTYPES];
    double *out;       // 16 sums
};
enum { E_KIN = 0, E_THERM, E_MAG, E_POT, E_XMOM, E_YMOM, E_ZMOM, E_ANGX, E_ANGY, E_ANGZ, E_MTOT, E_XCOM, E_YCOM, E_ZCOM, E_NP, E_RHOMAX, E_COUNT };

__global__ void k_energies(const EnArgs a)
{
    __shared__ double sh[32];
    double s[E_COUNT];
    for (int k = 0; k < E_COUNT; k++) s[k] = 0.;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double4 x = reinterpret_cast<const double4 *>(a.xyzh)[i];
        if (dead(x.w)) continue;                                   // accreted particles are outside this path
        const int itype = abs((int)a.iphase[i]);
        const double pm = a.massoftype[a.multitype ? itype : IGAS];
        const double rho = rhoh_d(x.w, pm, a.hfact);
        const double vx = a.v[a.nvu * i], vy = a.v[a.nvu * i + 1], vz = a.v[a.nvu * i + 2];
        s[E_XCOM] += pm * x.x; s[E_YCOM] += pm * x.y; s[E_ZCOM] += pm * x.z;
        s[E_MTOT] += pm; s[E_NP] += 1.; s[E_RHOMAX] = fmax(s[E_RHOMAX], rho);
        s[E_XMOM] += pm * vx; s[E_YMOM] += pm * vy; s[E_ZMOM] += pm * vz;
        s[E_ANGX] += pm * (x.y * vz - x.z * vy); s[E_ANGY] += pm * (x.z * vx - x.x * vz); s[E_ANGZ] += pm * (x.x * vy - x.y * vx);
        s[E_KIN] += pm * (vx * vx + vy * vy + vz * vz);
        if (a.gravity) s[E_POT] += (double)a.poten[i];
        if (itype == IGAS) {
            if (a.nvu >= 4) s[E_THERM] += pm * a.v[a.nvu * i + 3];
            else if (a.ieos == 2 && a.gamma > 1.001) s[E_THERM] += pm * (a.eos_vars[7 * i] / rho) / (a.gamma - 1.);     // :413-416
            if (a.mhd) {
                const double bx = a.B[4 * i] * rho, by = a.B[4 * i + 1] * rho, bz = a.B[4 * i + 2] * rho;
                s[E_MAG] += pm * (bx * bx + by * by + bz * bz) * (1. / rho);
            }
        }
    }
    for (int k = 0; k < E_COUNT; k++) {
        const double r = (k == E_RHOMAX) ? block_max(s[k], sh) : block_sum(s[k], sh);
        if (threadIdx.x == 0) { if (k == E_RHOMAX) atomic_max_pos(&a.out[k], r); else atomicAdd(&a.out[k], r); }
    }
}

// st_calcAccel (forcing.f90:728-830): thread = particle, the mode table streams through shared memory in tiles
struct ForcingArgs {
    int64_t n; int nvu, nmodes, ind_ts, correct_mean;
    const double *xyzh; const int8_t *iphase; double *f; const double *tab;     // tab: 10 doubles per mode {k(3), ampl, aka(3), akb(3)}
    double fac; double *fmean;                                                     // fac = 2 amplfac solweightnorm
};
#define FORC_TILE 64

__global__ void __launch_bounds__(256) k_forcing(const ForcingArgs a)
{
    __shared__ double tile[FORC_TILE * 10];
    __shared__ double sh[32];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool live = false, active = false;
    double x = 0., y = 0., z = 0.;
    if (i < a.n) {
        const double4 p = reinterpret_cast<const double4 *>(a.xyzh)[i];
        live = true;                                        // the reference loops over 1..npart without a dead-particle test
        active = a.ind_ts ? (a.iphase[i] > 0) : true;
        x = p.x; y = p.y; z = p.z;
    }
    const bool work = live && (active || a.correct_mean);
    double fx = 0., fy = 0., fz = 0.;
    for (int m0 = 0; m0 < a.nmodes; m0 += FORC_TILE) {
        const int nt = min(FORC_TILE, a.nmodes - m0);
        __syncthreads();
        for (int k = threadIdx.x; k < nt * 10; k += blockDim.x) tile[k] = a.tab[(size_t)m0 * 10 + k];
        __syncthreads();
        if (work) {
            for (int m = 0; m < nt; m++) {
                const double *t = tile + 10 * m;
                const double kdotx = t[0] * x + t[1] * y + t[2] * z;
                double im, re;
                sincos(kdotx, &im, &re);
                fx += t[3] * (t[4] * re - t[7] * im);
                fy += t[3] * (t[5] * re - t[8] * im);
                fz += t[3] * (t[6] * re - t[9] * im);
            }
        }
    }
    fx *= a.fac; fy *= a.fac; fz *= a.fac;
    if (live && active) { a.f[a.nvu * i] = fx; a.f[a.nvu * i + 1] = fy; a.f[a.nvu * i + 2] = fz; }
    if (a.correct_mean) {
        if (!work) { fx = fy = fz = 0.; }
        const double sx = block_sum(fx, sh), sy = block_sum(fy, sh), sz = block_sum(fz, sh);
        if (threadIdx.x == 0) { atomicAdd(&a.fmean[0], sx); atomicAdd(&a.fmean[1], sy); atomicAdd(&a.fmean[2], sz); }
    }
}

__global__ void k_forcing_mean(const ForcingArgs a)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const bool active = a.ind_ts ? (a.iphase[i] > 0) : true;
    if (!active) return;
    const double inv = 1. / (double)a.n;
    a.f[a.nvu * i] -= a.fmean[0] * inv; a.f[a.nvu * i + 1] -= a.fmean[1] * inv; a.f[a.nvu * i + 2] -= a.fmean[2] * inv;
}

}  // namespace

#define SL(c, kern, ...) do { kern<<<(c)->numSMs * 8, 256, 0, (c)->stream>>>(__VA_ARGS__); (c)->launches++; } while (0)

extern "C" {

int sphgpu_step_resident(sphgpu_ctx *c, double dtsph, double tolv, sphgpu_step_out *out)
{
    if (!c || !(dtsph > 0.)) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const sphgpu_params &p = c->hp.p;
    if (p.ind_timesteps) { c->err = "step: this is the global-timestep leapfrog; with ind_timesteps use sphgpu_step_ind_resident"; return SPHGPU_ERR_ARG; }
    // a decomposed set (dist.cu) integrates its owned particles; the ghosts behind them are refreshed inside every derivs
    if (c->nghost > 0 && !c->dist) { c->err = "step: the context holds ghost particles of a host-driven halo exchange; use sphgpu_dist_step"; return SPHGPU_ERR_STATE; }
    const int64_t n = c->dist ? c->nlocal : c->npart;
    const int nvu = c->hp.nvu;
    if (n <= 0) return SPHGPU_ERR_STATE;
    cudaStream_t st = c->stream;
    CUDA_TRY(c, c->v_true.ensure((size_t)nvu * n)); if (p.mhd) CUDA_TRY(c, c->B_true.ensure(4 * (size_t)n));
    CUDA_TRY(c, c->dscal.ensure(DS_COUNT));
    // the evolved v, B live in v_true/B_true during the step; c->vxyzu / c->Bevol hold the predicted values derivs reads
    CUDA_TRY(c, cudaMemcpyAsync(c->v_true.p, c->vxyzu.p, sizeof(double) * nvu * n, cudaMemcpyDeviceToDevice, st));
    if (p.mhd) CUDA_TRY(c, cudaMemcpyAsync(c->B_true.p, c->Bevol.p, sizeof(double) * 4 * n, cudaMemcpyDeviceToDevice, st));
    StepArgs a; memset(&a, 0, sizeof a);
    a.n = n; a.nvu = nvu; a.mhd = p.mhd; a.nalpha = c->hp.nalpha; a.multitype = 1;
    a.hfact = p.hfact; a.dt = dtsph; a.hdt = 0.5 * dtsph;
    for (int k = 0; k < SPHGPU_MAXTYPES; k++) a.massoftype[k] = p.massoftype[k];
    auto bind = [&]() {        // the canonical arrays may be reallocated when a derivs call appends more ghosts
        a.xyzh = c->xyzh.p; a.v = c->v_true.p; a.vpred = c->vxyzu.p; a.f = c->fxyzu.p; a.B = c->B_true.p; a.Bpred = c->Bevol.p; a.dB = c->dBevol.p;
        a.eos_vars = c->eos_vars.p; a.divcurlv = c->divcurlv.p; a.alphaind = c->alphaind.p; a.iphase = c->iphase.p; a.red = c->dscal.p + 16;
    };
    auto derivs = [&](int icall, sphgpu_scalars *sc) -> int {
        const int r = c->dist ? sphgpu_dist_hook_derivs(c, icall, dtsph, sc) : sphgpu_derivs_resident(c, icall, dtsph, sc);
        bind();
        return r;
    };
    bind();
    SL(c, k_predict, a);
    SL(c, k_drift, a);
    SL(c, k_predict_sph, a);
    c->tree_valid = false;
    sphgpu_scalars sc;
    TRY(derivs(1, &sc));
    int its = 0; bool converged = false;
    double errmax = 0., dterr = 1.e29;
    while (its < 30 && !converged) {                                  // step_leapfrog.f90:441-759
        its++;
        CUDA_TRY(c, cudaMemsetAsync(a.red, 0, 3 * sizeof(double), st));
        SL(c, k_correct, a);
        double red[3];
        CUDA_TRY(c, cudaMemcpyAsync(red, a.red, sizeof red, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
        if (c->dist) TRY(sphgpu_dist_hook_reduce_err(c, red));        // reduceall_mpi of errmax, v2mean, np (step_leapfrog.f90:790-792)
        // check_velocity_error (step_leapfrog.f90:769-843)
        const double v2mean = red[2] > 0. ? red[1] / red[2] : 0.;
        errmax = v2mean > DBL_MIN ? red[0] / sqrt(v2mean) : 0.;
        double errtol = tolv;
        if (tolv < 1.e2) {
            const double dtf = fmin(sc.dtcourant, sc.dtforce);
            if (dtf > dtsph && dtf < 1.e29) errtol = errtol * (dtsph / dtf) * (dtsph / dtf);
            if (its == 1 && errtol > DBL_MIN && errmax > DBL_EPSILON) dterr = dtsph * sqrt(errtol / errmax);
            converged = errmax < tolv;
        } else converged = true;
        if (!converged) {
            SL(c, k_unconverged, a);
            TRY(derivs(2, &sc));
        }
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->vxyzu.p, c->v_true.p, sizeof(double) * nvu * n, cudaMemcpyDeviceToDevice, st));
    if (p.mhd) CUDA_TRY(c, cudaMemcpyAsync(c->Bevol.p, c->B_true.p, sizeof(double) * 4 * n, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    CUDA_TRY(c, cudaGetLastError());
    if (out) { memset(out, 0, sizeof *out); out->dtcourant = sc.dtcourant; out->dtforce = sc.dtforce; out->dterr = dterr; out->errmax = errmax; out->its = its; out->scalars = sc; }
    return SPHGPU_OK;
}

// ---- individual timesteps ---------------------------------------------------------------------------------------------------------
// init_step (step_leapfrog.f90:57-80): at time 0 every particle starts in the finest bin; twas = the half step of each particle's bin
int sphgpu_init_step_resident(sphgpu_ctx *c, double time, double dtmax, int nbinmax)
{
    if (!c || !(dtmax > 0.) || nbinmax < 0 || nbinmax > 30) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int64_t n = c->dist ? c->nlocal : c->npart;
    if (n <= 0) return SPHGPU_ERR_STATE;
    CUDA_TRY(c, c->twas.ensure(n));
    k_init_step<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, c->iphase.p, c->ibin.p, c->twas.p, time, dtmax, nbinmax, time < DBL_MIN ? 1 : 0);
    c->launches++;
    c->nbinmax = nbinmax;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPHGPU_OK;
}

// set_active_particles (utils_indtimesteps.f90:114-178) on the resident state; also hands nbinmax, istepfrac and the resulting ibinnow to
// the force pass (what sphgpu_set_timestep_bins does for a host-driven run)
int sphgpu_set_active_particles_resident(sphgpu_ctx *c, int nbinmax, int istepfrac, int64_t *nactive, int64_t *nalive)
{
    if (!c || nbinmax < 0 || nbinmax > 30 || istepfrac < 0) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int64_t n = c->dist ? c->nlocal : c->npart;
    if (n <= 0) return SPHGPU_ERR_STATE;
    CUDA_TRY(c, c->counters.ensure(CNT_COUNT));
    CUDA_TRY(c, cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(unsigned long long), c->stream));
    k_set_active<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, c->xyzh.p, c->iphase.p, c->ibin.p, nbinmax, istepfrac, c->counters.p);
    c->launches++;
    unsigned long long h[2];
    CUDA_TRY(c, cudaMemcpyAsync(h, c->counters.p + 2, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long err = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&err, c->counters.p + CNT_ERR, sizeof err, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (err) { c->err = "set_active_particles: timestep bin exceeds max bins"; return SPHGPU_ERR_STATE; }
    int ibinnow = nbinmax;
    for (int i = 0; ibinnow == nbinmax && i < nbinmax; i++) if (istepfrac % (1 << (nbinmax - i)) == 0) ibinnow = i;
    c->nbinmax = nbinmax; c->ibinnow = ibinnow; c->istepfrac = istepfrac;
    c->tree_valid = false;                                           // the active counts of the cells changed
    if (nactive) *nactive = (int64_t)h[0];
    if (nalive) *nalive = (int64_t)h[1];
    return SPHGPU_OK;
}

// step (step_leapfrog.f90:95-760) with -DIND_TIMESTEPS: dtsph is the smallest timestep dtmax / 2^nbinmax, t the time at the start of it.
// The force pass moves active particles between bins; out->scalars.nbinmaxnew is the new nbinmax (timestep_ind module variable).
int sphgpu_step_ind_resident(sphgpu_ctx *c, double t, double dtsph, double dtmax, sphgpu_step_out *out)
{
    if (!c || !(dtsph > 0.) || !(dtmax > 0.)) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const sphgpu_params &p = c->hp.p;
    if (!p.ind_timesteps) { c->err = "step_ind: the context was not created with ind_timesteps"; return SPHGPU_ERR_ARG; }
    if (c->nghost > 0 && !c->dist) { c->err = "step_ind: the context holds ghost particles of a host-driven halo exchange"; return SPHGPU_ERR_STATE; }
    const int64_t n = c->dist ? c->nlocal : c->npart;
    const int nvu = c->hp.nvu;
    if (n <= 0 || c->twas.cap < (size_t)n) { c->err = "step_ind: sphgpu_init_step_resident has not been called"; return SPHGPU_ERR_STATE; }
    cudaStream_t st = c->stream;
    CUDA_TRY(c, c->v_true.ensure((size_t)nvu * n)); if (p.mhd) CUDA_TRY(c, c->B_true.ensure(4 * (size_t)n));
    CUDA_TRY(c, cudaMemcpyAsync(c->v_true.p, c->vxyzu.p, sizeof(double) * nvu * n, cudaMemcpyDeviceToDevice, st));
    if (p.mhd) CUDA_TRY(c, cudaMemcpyAsync(c->B_true.p, c->Bevol.p, sizeof(double) * 4 * n, cudaMemcpyDeviceToDevice, st));
    StepArgs a; memset(&a, 0, sizeof a);
    a.n = n; a.nvu = nvu; a.mhd = p.mhd; a.nalpha = c->hp.nalpha; a.multitype = 1;
    a.hfact = p.hfact; a.dt = dtsph; a.hdt = 0.5 * dtsph; a.ind = 1; a.timei = t;
    for (int k = 0; k < SPHGPU_MAXTYPES; k++) a.massoftype[k] = p.massoftype[k];
    const double time_now = t + dtsph;
    for (int b = 0; b <= 30; b++) {                                  // ibin_dts (:74-79, :157-164)
        const double dtb = dtmax / (double)(1ull << b);
        a.thdt[b] = 0.5 * dtb;
        a.ttwas[b] = ((double)(long long)(time_now * (1.0 / dtb)) + 0.5) * dtb;
    }
    auto bind = [&]() {
        a.xyzh = c->xyzh.p; a.v = c->v_true.p; a.vpred = c->vxyzu.p; a.f = c->fxyzu.p; a.B = c->B_true.p; a.Bpred = c->Bevol.p; a.dB = c->dBevol.p;
        a.eos_vars = c->eos_vars.p; a.divcurlv = c->divcurlv.p; a.alphaind = c->alphaind.p; a.iphase = c->iphase.p; a.red = c->dscal.p + 16;
        a.twas = c->twas.p; a.ibin = c->ibin.p; a.ibin_old = c->ibin_old.p; a.ibin_wake = c->ibin_wake.p;
    };
    bind();
    SL(c, k_predict, a);
    SL(c, k_drift, a);
    a.timei = time_now;
    SL(c, k_predict_sph, a);
    c->tree_valid = false;
    sphgpu_scalars sc;
    TRY(c->dist ? sphgpu_dist_hook_derivs(c, 1, dtsph, &sc) : sphgpu_derivs_resident(c, 1, dtsph, &sc));
    bind();
    c->nbinmax = (int)sc.nbinmaxnew;                                 // force.F90:792
    a.nbinmax = c->nbinmax;
    SL(c, k_correct_ind, a);
    CUDA_TRY(c, cudaMemcpyAsync(c->vxyzu.p, c->v_true.p, sizeof(double) * nvu * n, cudaMemcpyDeviceToDevice, st));
    if (p.mhd) CUDA_TRY(c, cudaMemcpyAsync(c->Bevol.p, c->B_true.p, sizeof(double) * 4 * n, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    CUDA_TRY(c, cudaGetLastError());
    if (out) { memset(out, 0, sizeof *out); out->dtcourant = sc.dtcourant; out->dtforce = sc.dtforce; out->dterr = 1.e29; out->errmax = 0.; out->its = 1; out->scalars = sc; }
    return SPHGPU_OK;
}

int sphgpu_set_forcing_modes(sphgpu_ctx *c, int nmodes, const double *mode, const double *ampl, const double *aka, const double *akb, double amplfac,
                             double solweightnorm, int correct_mean_force)
{
    if (!c || nmodes < 0 || (nmodes > 0 && (!mode || !ampl || !aka || !akb))) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    std::vector<double> tab((size_t)10 * (nmodes > 0 ? nmodes : 1));
    for (int m = 0; m < nmodes; m++) {
        double *t = tab.data() + 10 * (size_t)m;
        t[0] = mode[3 * m]; t[1] = mode[3 * m + 1]; t[2] = mode[3 * m + 2]; t[3] = ampl[m];
        for (int k = 0; k < 3; k++) { t[4 + k] = aka[3 * m + k]; t[7 + k] = akb[3 * m + k]; }
    }
    CUDA_TRY(c, c->forc_tab.ensure(tab.size()));
    CUDA_TRY(c, cudaMemcpyAsync(c->forc_tab.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->forc_nmodes = nmodes; c->forc_fac = 2. * amplfac * solweightnorm; c->forc_correct_mean = correct_mean_force;
    return SPHGPU_OK;
}

int sphgpu_forcing_resident(sphgpu_ctx *c)
{
    if (!c) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int64_t n = c->nlocal > 0 ? c->nlocal : c->npart;
    if (n <= 0) return SPHGPU_ERR_STATE;
    CUDA_TRY(c, c->dscal.ensure(DS_COUNT));
    CUDA_TRY(c, c->forc_tab.ensure(10));
    ForcingArgs a; memset(&a, 0, sizeof a);
    a.n = n; a.nvu = c->hp.nvu; a.nmodes = c->forc_nmodes; a.ind_ts = c->hp.p.ind_timesteps; a.correct_mean = c->forc_correct_mean;
    a.xyzh = c->xyzh.p; a.iphase = c->iphase.p; a.f = c->fxyzu.p; a.tab = c->forc_tab.p; a.fac = c->forc_fac; a.fmean = c->dscal.p + 16;
    CUDA_TRY(c, cudaMemsetAsync(a.fmean, 0, 3 * sizeof(double), c->stream));
    k_forcing<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a);
    c->launches++;
    if (a.correct_mean) { k_forcing_mean<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a); c->launches++; }
    CUDA_TRY(c, cudaGetLastError());
    return SPHGPU_OK;
}

int sphgpu_energies_resident(sphgpu_ctx *c, sphgpu_energies *out)
{
    if (!c || !out) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const sphgpu_params &p = c->hp.p;
    const int64_t n = c->nlocal > 0 ? c->nlocal : c->npart;          // ghosts are not summed
    if (n <= 0) return SPHGPU_ERR_STATE;
    CUDA_TRY(c, c->dscal.ensure(DS_COUNT));
    double *dout = c->dscal.p + 16;
    CUDA_TRY(c, cudaMemsetAsync(dout, 0, E_COUNT * sizeof(double), c->stream));
    EnArgs a; memset(&a, 0, sizeof a);
    a.n = n; a.nvu = c->hp.nvu; a.mhd = p.mhd; a.gravity = p.gravity; a.ieos = p.ieos; a.multitype = 1;
    a.xyzh = c->xyzh.p; a.v = c->vxyzu.p; a.B = c->Bevol.p; a.eos_vars = c->eos_vars.p; a.poten = c->poten.p; a.iphase = c->iphase.p;
    a.hfact = p.hfact; a.gamma = p.gamma; a.out = dout;
    for (int k = 0; k < SPHGPU_MAXTYPES; k++) a.massoftype[k] = p.massoftype[k];
    SL(c, k_energies, a);
    double h[E_COUNT];
    CUDA_TRY(c, cudaMemcpyAsync(h, dout, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaGetLastError());
    memset(out, 0, sizeof *out);
    out->ekin = 0.5 * h[E_KIN]; out->etherm = h[E_THERM]; out->emag = 0.5 * h[E_MAG]; out->epot = h[E_POT];        // energies.f90:676-691
    out->etot = out->ekin + out->etherm + out->emag + out->epot;
    out->xmom = h[E_XMOM]; out->ymom = h[E_YMOM]; out->zmom = h[E_ZMOM];
    out->totmom = sqrt(h[E_XMOM] * h[E_XMOM] + h[E_YMOM] * h[E_YMOM] + h[E_ZMOM] * h[E_ZMOM]);
    out->angx = h[E_ANGX]; out->angy = h[E_ANGY]; out->angz = h[E_ANGZ];
    out->angtot = sqrt(h[E_ANGX] * h[E_ANGX] + h[E_ANGY] * h[E_ANGY] + h[E_ANGZ] * h[E_ANGZ]);
    out->mtot = h[E_MTOT];
    const double dm = h[E_MTOT] > 0. ? 1. / h[E_MTOT] : 0.;
    out->xcom = h[E_XCOM] * dm; out->ycom = h[E_YCOM] * dm; out->zcom = h[E_ZCOM] * dm;
    out->np = (int64_t)h[E_NP]; out->rhomax = h[E_RHOMAX];
    return SPHGPU_OK;
}

}  // extern "C"
