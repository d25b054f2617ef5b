// cons2prim.cu -- per-particle EOS + Cullen-Dehnen alpha_loc between the density and force passes:
// replaces cons2prim_everything (src/main/cons2prim.f90:274-456) for ieos = 1,2,3
// (equationofstate, src/main/eos.f90:183-256; get_alphaloc / xi_limiter, src/main/shock_capturing.f90:131-178).
// One thread per particle on the canonical (caller-order) arrays; HBM-bound, ~120 B/particle.
#include "common.cuh"
#include <float.h>

namespace {

__global__ void k_cons2prim(int64_t n, const double *__restrict__ xyzh, const double *__restrict__ vxyzu, const float *__restrict__ dvdx,
                            const double *__restrict__ Bevol, const int8_t *__restrict__ iphase, double *__restrict__ eos_vars,
                            float *__restrict__ alphaind, double *__restrict__ Bxyz, const __grid_constant__ DevParams dp, unsigned long long *cnt)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const sphgpu_params &p = dp.p;
    const double4 x = reinterpret_cast<const double4 *>(xyzh)[i];
    if (x.w < DBL_MIN) return;                                   // isdead_or_accreted
    bool act, gas, dust; int itype;
    get_partinfo_d(iphase[i], p.set_boundaries_to_active, p.dust, act, gas, dust, itype);
    const double hi = x.w, pmassi = p.massoftype[itype];
    const double rhoi = rhoh_d(hi, pmassi, p.hfact);
    const double rhogas = rhoi;
    if (!gas) return;
    double ponrhoi, spsoundi;
    if (p.ieos == 1) { ponrhoi = p.polyk; spsoundi = sqrt(ponrhoi); }
    else if (p.ieos == 2) {
        if (dp.nvu >= 4) {
            const double eni = vxyzu[(size_t)dp.nvu * i + 3];
            if (eni < 0.) { atomicMax(&cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_ARG); atomicMax(&cnt[CNT_ERRID], (unsigned long long)(i + 1)); }
            if (p.gamma > 1.0001) ponrhoi = (p.gamma - 1.) * eni; else ponrhoi = 2. / 3. * eni;
        } else ponrhoi = p.polyk * pow(rhogas, p.gamma - 1.);
        spsoundi = sqrt(p.gamma * ponrhoi);
    } else {
        ponrhoi = p.polyk * pow(x.x * x.x + x.y * x.y + x.z * x.z, -p.qfacdisc);
        ponrhoi = fmax(ponrhoi, p.cs_min * p.cs_min);
        spsoundi = sqrt(ponrhoi);
    }
    double *ev = eos_vars + 7 * (size_t)i;
    ev[0] = ponrhoi * rhogas; ev[1] = spsoundi; ev[2] = p.temp_coef_mu * ponrhoi;   // igasP, ics, itemp (cons2prim.f90:390-392); imu, iX, iZ, igamma are not touched
    if (dp.nalpha >= 2) {
        const float *d = dvdx + 9 * (size_t)i;
        const double dvxdx = d[0], dvxdy = d[1], dvxdz = d[2], dvydx = d[3], dvydy = d[4], dvydz = d[5], dvzdx = d[6], dvzdy = d[7], dvzdz = d[8];
        const double divv = dvxdx + dvydy + dvzdz;
        const double curlvx = dvzdy - dvydz, curlvy = dvxdz - dvzdx, curlvz = dvydx - dvxdy;
        const double m = fmax(-divv, 0.);
        const double fac = m * m;
        const double traceS = curlvx * curlvx + curlvy * curlvy + curlvz * curlvz;
        const double xi_lim = (fac + traceS > DBL_EPSILON) ? fac / (fac + traceS) : 1.;
        const double divvdti = (double)alphaind[3 * (size_t)i + 2];
        const double source = 10. * hi * hi * xi_lim * fmax(-divvdti, 0.);
        const double temp = spsoundi * spsoundi;
        double alphaloc;
        if (temp > DBL_EPSILON) alphaloc = fmax(fmin(source / temp, p.alphamax), p.alpha);
        else alphaloc = p.alpha;
        alphaind[3 * (size_t)i + 1] = (float)alphaloc;
    }
    if (p.mhd && Bxyz) {
        const double4 B = reinterpret_cast<const double4 *>(Bevol)[i];
        Bxyz[3 * (size_t)i] = B.x * rhoi; Bxyz[3 * (size_t)i + 1] = B.y * rhoi; Bxyz[3 * (size_t)i + 2] = B.z * rhoi;
    }
}

}  // namespace

int cons2prim_run(sphgpu_ctx *c)
{
    const int64_t n = c->npart;
    CUDA_TRY(c, c->counters.ensure(CNT_COUNT));
    CUDA_TRY(c, cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 4, c->stream));
    if (c->hp.p.mhd) CUDA_TRY(c, c->Bxyz.ensure(3 * n));
    k_cons2prim<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(n, c->xyzh.p, c->vxyzu.p, c->dvdx.p, c->Bevol.p, c->iphase.p, c->eos_vars.p, c->alphaind.p,
                                                                 c->hp.p.mhd ? c->Bxyz.p : nullptr, c->hp, c->counters.p);
    c->launches++;
    c->eos_on_device = true;
    CUDA_TRY(c, cudaGetLastError());
    return SPHGPU_OK;
}
