// halo.cu -- ghost-particle halo for the multi-GPU path (one process per GPU): replaces the reference's MPI cell export
// (src/main/mpi_dens.F90, mpi_force.F90, mpi_derivs.F90:197-522) with two packed exchanges per derivs.
//
// Each rank owns the particles of one box of a spatial decomposition (the same mass-centre / longest-axis bisection the
// reference uses globally, kdtree.F90:2098-2160, done on the host).  Ghosts = remote particles within
// dhalo = radkern * hmax_global * margin of this rank's box (minimum image).  They are appended after the nlocal owned
// particles as INACTIVE particles (iphase < 0): the tree / density / force kernels already treat inactive particles as
// neighbour-only (individual-timestep semantics, dens.F90:1329, force.F90:2255), so no kernel changes are needed and the
// owner alone computes each particle's sums (no reverse reduction, unlike mpi_derivs.F90:462-522).
//   stage 1 (before the tree):  {x,y,z,h, v(3),u, f+fext(3), B/rho(3),psi, iphase}   16 doubles per ghost
//   stage 2 (after the density): {h, gradh, alpha, gradsoft}                           4 doubles per ghost
// The transfers themselves are NCCL all-to-all-v over NVLink issued by the host side (torch.distributed) directly on the
// device buffers returned by sphgpu_halo_pack / sphgpu_halo_recvbuf.
#include "common.cuh"
#include <float.h>

namespace {

__device__ __forceinline__ double axis_gap(double x, double lo, double hi, double L, bool periodic)
{
    double g = fmax(0., fmax(lo - x, x - hi));
    if (periodic) {
        // nearest image of the interval
        const double g2 = fmax(0., fmax(lo - (x - L), (x - L) - hi));
        const double g3 = fmax(0., fmax(lo - (x + L), (x + L) - hi));
        g = fmin(g, fmin(g2, g3));
    }
    return g;
}

template <bool FILL>
__global__ void k_halo_select(int64_t nlocal, const double *__restrict__ xyzh, int nranks, int myrank, const double *__restrict__ boxes, double dhalo,
                              double Lx, double Ly, double Lz, int periodic, unsigned long long *cnt, const long long *off, int *sendidx)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    const double4 x = reinterpret_cast<const double4 *>(xyzh)[i];
    if (x.w < DBL_MIN) return;
    for (int r = 0; r < nranks; r++) {
        if (r == myrank) continue;
        const double *b = boxes + 6 * r;
        const double gx = axis_gap(x.x, b[0], b[3], Lx, periodic), gy = axis_gap(x.y, b[1], b[4], Ly, periodic), gz = axis_gap(x.z, b[2], b[5], Lz, periodic);
        if (gx * gx + gy * gy + gz * gz < dhalo * dhalo) {
            const unsigned long long k = atomicAdd(&cnt[r], 1ull);
            if (FILL) sendidx[off[r] + (long long)k] = (int)i;
        }
    }
}

__global__ void k_halo_pack1(int64_t nsend, const int *__restrict__ sendidx, const double *__restrict__ xyzh, const double *__restrict__ vxyzu,
                             const double *__restrict__ fxyzu, const double *__restrict__ fext, const double *__restrict__ Bevol,
                             const int8_t *__restrict__ iphase, int nvu, int mhd, double *__restrict__ out)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nsend) return;
    const int i = sendidx[k];
    double *o = out + 16 * k;
    const double4 x = reinterpret_cast<const double4 *>(xyzh)[i];
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
    const double *v = vxyzu + (size_t)nvu * i, *f = fxyzu + (size_t)nvu * i, *fe = fext + 3 * (size_t)i;
    o[4] = v[0]; o[5] = v[1]; o[6] = v[2]; o[7] = nvu >= 4 ? v[3] : 0.;
    o[8] = f[0] + fe[0]; o[9] = f[1] + fe[1]; o[10] = f[2] + fe[2];
    if (mhd) { const double4 B = reinterpret_cast<const double4 *>(Bevol)[i]; o[11] = B.x; o[12] = B.y; o[13] = B.z; o[14] = B.w; }
    else { o[11] = o[12] = o[13] = o[14] = 0.; }
    o[15] = (double)iphase[i];
}

__global__ void k_halo_unpack1(int64_t nghost, int64_t nlocal, const double *__restrict__ in, double *__restrict__ xyzh, double *__restrict__ vxyzu,
                               double *__restrict__ fxyzu, double *__restrict__ fext, double *__restrict__ Bevol, int8_t *__restrict__ iphase, int nvu,
                               int mhd)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nghost) return;
    const int64_t i = nlocal + k;
    const double *o = in + 16 * k;
    reinterpret_cast<double4 *>(xyzh)[i] = make_double4(o[0], o[1], o[2], o[3]);
    double *v = vxyzu + (size_t)nvu * i, *f = fxyzu + (size_t)nvu * i, *fe = fext + 3 * (size_t)i;
    v[0] = o[4]; v[1] = o[5]; v[2] = o[6]; if (nvu >= 4) { v[3] = o[7]; f[3] = 0.; }
    f[0] = o[8]; f[1] = o[9]; f[2] = o[10];
    fe[0] = fe[1] = fe[2] = 0.;
    if (mhd) reinterpret_cast<double4 *>(Bevol)[i] = make_double4(o[11], o[12], o[13], o[14]);
    const int t = abs((int)o[15]);
    iphase[i] = (int8_t)(-t);                    // inactive: neighbour only
}

__global__ void k_halo_pack2(int64_t nsend, const int *__restrict__ sendidx, const double *__restrict__ xyzh, const float *__restrict__ gradh,
                             const float *__restrict__ alphaind, int ngradh, double *__restrict__ out)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nsend) return;
    const int i = sendidx[k];
    double *o = out + 4 * k;
    o[0] = xyzh[4 * (size_t)i + 3];
    o[1] = (double)gradh[(size_t)ngradh * i];
    o[2] = (double)alphaind[3 * (size_t)i];
    o[3] = ngradh > 1 ? (double)gradh[(size_t)ngradh * i + 1] : 0.;
}

__global__ void k_halo_unpack2(int64_t nghost, int64_t nlocal, const double *__restrict__ in, double *__restrict__ xyzh, float *__restrict__ gradh,
                               float *__restrict__ alphaind, int ngradh)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nghost) return;
    const int64_t i = nlocal + k;
    const double *o = in + 4 * k;
    xyzh[4 * (size_t)i + 3] = o[0];
    gradh[(size_t)ngradh * i] = (float)o[1];
    alphaind[3 * (size_t)i] = (float)o[2];
    if (ngradh > 1) gradh[(size_t)ngradh * i + 1] = (float)o[3];
}

__global__ void k_refresh_h(int64_t nlive, const int *__restrict__ perm, const double *__restrict__ xyzh, double4 *__restrict__ pos4)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    pos4[s].w = xyzh[4 * (size_t)perm[s] + 3];
}

__global__ void k_hmax(int64_t n, const double *__restrict__ xyzh, double *out)
{
    double m = 0.;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmax(m, xyzh[4 * i + 3]);
    m = warp_max(m);
    if (lane_id() == 0) atomic_max_pos(out, m);
}

}  // namespace

static inline int nblk(int64_t n, int b) { return (int)((n + b - 1) / b); }

extern "C" {

// max smoothing length of the owned particles (the host reduces it over ranks to size the halo)
int sphgpu_local_hmax(sphgpu_ctx *c, double *hmax)
{
    if (!c || !hmax) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, c->dscal.ensure(DS_COUNT));
    CUDA_TRY(c, cudaMemsetAsync(c->dscal.p + DS_COUNT - 1, 0, sizeof(double), c->stream));
    k_hmax<<<c->numSMs * 4, 256, 0, c->stream>>>(c->nlocal, c->xyzh.p, c->dscal.p + DS_COUNT - 1);
    c->launches++;
    CUDA_TRY(c, cudaMemcpyAsync(hmax, c->dscal.p + DS_COUNT - 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPHGPU_OK;
}

// largest trial smoothing length of the last density pass: if radkern * this exceeds the halo width the ghosts were selected with,
// some particle iterated without all its neighbours -> the caller restores h and repeats with a wider halo
int sphgpu_density_hmax_used(sphgpu_ctx *c, double *hmax)
{
    if (!c || !hmax) return SPHGPU_ERR_ARG;
    *hmax = c->dens_hmax_used;
    return SPHGPU_OK;
}

// largest h_new / h_old of the last density pass on this rank; the driver reduces it over the ranks and hands the global value back
// (sphgpu_set_option "halo_hgrow") so that stage 2 can inflate the tree's hmax instead of refitting it
int sphgpu_density_hgrow(sphgpu_ctx *c, double *hgrow)
{
    if (!c || !hgrow) return SPHGPU_ERR_ARG;
    *hgrow = c->dens_hgrow;
    return SPHGPU_OK;
}

__global__ void k_restore_h(int64_t n, double *__restrict__ xyzh, const double *__restrict__ h_build)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) xyzh[4 * i + 3] = h_build[i];
}

// put back the smoothing lengths the last build_tree saw (owned particles), for a halo-widening retry of the density pass
int sphgpu_halo_restore_h(sphgpu_ctx *c)
{
    if (!c) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->h_build.cap < (size_t)c->nlocal) { c->err = "halo_restore_h: no build_tree has run"; return SPHGPU_ERR_STATE; }
    k_restore_h<<<(unsigned)((c->nlocal + 255) / 256), 256, 0, c->stream>>>(c->nlocal, c->xyzh.p, c->h_build.p);
    c->launches++;
    c->tree_valid = false;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPHGPU_OK;
}

// boxes: 6 doubles per rank {xlo,ylo,zlo,xhi,yhi,zhi}; counts[nranks] receives the number of owned particles each rank needs
int sphgpu_halo_select(sphgpu_ctx *c, int nranks, int myrank, const double *boxes, double dhalo, int64_t *counts)
{
    if (!c || !boxes || !counts || nranks < 1 || myrank < 0 || myrank >= nranks) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->halo_nranks = nranks; c->halo_rank = myrank;
    c->npart = c->nlocal; c->nghost = 0; c->tree_valid = false;      // drop the ghosts of the previous step
    const int64_t n = c->nlocal;
    CUDA_TRY(c, c->halo_boxes.ensure(6 * nranks)); CUDA_TRY(c, c->halo_cnt.ensure(2 * nranks + 2));
    CUDA_TRY(c, cudaMemcpyAsync(c->halo_boxes.p, boxes, sizeof(double) * 6 * nranks, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->halo_cnt.p, 0, sizeof(unsigned long long) * nranks, c->stream));
    const sphgpu_params &p = c->hp.p;
    k_halo_select<false><<<nblk(n, 256), 256, 0, c->stream>>>(n, c->xyzh.p, nranks, myrank, c->halo_boxes.p, dhalo, c->hp.dxbound, c->hp.dybound, c->hp.dzbound,
                                                              p.periodic, c->halo_cnt.p, nullptr, nullptr);
    c->launches++;
    std::vector<unsigned long long> hc(nranks);
    CUDA_TRY(c, cudaMemcpyAsync(hc.data(), c->halo_cnt.p, sizeof(unsigned long long) * nranks, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->halo_sendcnt.assign(nranks, 0); c->halo_sendoff.assign(nranks + 1, 0);
    for (int r = 0; r < nranks; r++) { c->halo_sendcnt[r] = (long long)hc[r]; c->halo_sendoff[r + 1] = c->halo_sendoff[r] + c->halo_sendcnt[r]; counts[r] = (int64_t)hc[r]; }
    const long long tot = c->halo_sendoff[nranks];
    CUDA_TRY(c, c->halo_sendidx.ensure(tot + 1));
    long long *doff = reinterpret_cast<long long *>(c->halo_cnt.p + nranks);
    CUDA_TRY(c, cudaMemcpyAsync(doff, c->halo_sendoff.data(), sizeof(long long) * (nranks + 1), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->halo_cnt.p, 0, sizeof(unsigned long long) * nranks, c->stream));
    k_halo_select<true><<<nblk(n, 256), 256, 0, c->stream>>>(n, c->xyzh.p, nranks, myrank, c->halo_boxes.p, dhalo, c->hp.dxbound, c->hp.dybound, c->hp.dzbound,
                                                             p.periodic, c->halo_cnt.p, doff, c->halo_sendidx.p);
    c->launches++;
    if (!c->stream_blocking) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaGetLastError());
    return SPHGPU_OK;
}

// packs the selected particles (destination-major) and returns the device pointer of the send buffer
int sphgpu_halo_pack(sphgpu_ctx *c, int stage, void **sendptr, int *record_doubles)
{
    if (!c || !sendptr || !record_doubles || (stage != 1 && stage != 2)) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const long long tot = c->halo_sendoff.empty() ? 0 : c->halo_sendoff[c->halo_nranks];
    const int rd = stage == 1 ? 16 : 4;
    CUDA_TRY(c, c->halo_sendbuf.ensure((size_t)rd * (tot + 1)));
    if (tot > 0) {
        if (stage == 1) k_halo_pack1<<<nblk(tot, 256), 256, 0, c->stream>>>(tot, c->halo_sendidx.p, c->xyzh.p, c->vxyzu.p, c->fxyzu.p, c->fext.p, c->Bevol.p,
                                                                            c->iphase.p, c->hp.nvu, c->hp.p.mhd, c->halo_sendbuf.p);
        else k_halo_pack2<<<nblk(tot, 256), 256, 0, c->stream>>>(tot, c->halo_sendidx.p, c->xyzh.p, c->gradh.p, c->alphaind.p, c->hp.ngradh, c->halo_sendbuf.p);
        c->launches++;
    }
    if (!c->stream_blocking) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaGetLastError());
    *sendptr = c->halo_sendbuf.p; *record_doubles = rd;
    return SPHGPU_OK;
}

int sphgpu_halo_recvbuf(sphgpu_ctx *c, int64_t nrecords, int record_doubles, void **recvptr)
{
    if (!c || !recvptr || nrecords < 0) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, c->halo_recvbuf.ensure((size_t)record_doubles * (nrecords + 1)));
    *recvptr = c->halo_recvbuf.p;
    return SPHGPU_OK;
}

// stage 1: append nghost inactive particles after the owned ones; stage 2: refresh their h, gradh, alpha and the tree's hmax
int sphgpu_halo_unpack(sphgpu_ctx *c, int stage, int64_t nghost)
{
    if (!c || nghost < 0 || (stage != 1 && stage != 2)) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (stage == 1) {
        const int64_t ntot = c->nlocal + nghost;
        TRY(ensure_all_keep(c, ntot, c->nlocal));
        if (nghost > 0) {
            k_halo_unpack1<<<nblk(nghost, 256), 256, 0, c->stream>>>(nghost, c->nlocal, c->halo_recvbuf.p, c->xyzh.p, c->vxyzu.p, c->fxyzu.p, c->fext.p,
                                                                     c->Bevol.p, c->iphase.p, c->hp.nvu, c->hp.p.mhd);
            c->launches++;
        }
        c->nghost = nghost; c->npart = ntot; c->tree_valid = false;
    } else {
        if (nghost != c->nghost) { c->err = "halo_unpack: stage 2 ghost count differs from stage 1"; return SPHGPU_ERR_ARG; }
        if (nghost > 0) {
            k_halo_unpack2<<<nblk(nghost, 256), 256, 0, c->stream>>>(nghost, c->nlocal, c->halo_recvbuf.p, c->xyzh.p, c->gradh.p, c->alphaind.p, c->hp.ngradh);
            c->launches++;
            if (c->tree_valid) {
                k_refresh_h<<<nblk(c->nlive, 256), 256, 0, c->stream>>>(c->nlive, c->perm.p, c->xyzh.p, c->pos4.p);
                c->launches++;
                // the ghosts' h grew by at most the global growth factor: inflate the tree's hmax by it when it is small, else refit
                if (c->halo_hgrow > 0. && c->halo_hgrow <= 1.02 && !c->always_refit) c->hscale = fmax(c->hscale, fmax(c->halo_hgrow, 1.) * (1. + 1e-12));
                else TRY(tree_refit_hmax(c));
                c->halo_hgrow = 0.;
            }
        }
    }
    if (!c->stream_blocking) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaGetLastError());
    return SPHGPU_OK;
}

int64_t sphgpu_nghost(sphgpu_ctx *c) { return c ? c->nghost : 0; }

}  // extern "C"
