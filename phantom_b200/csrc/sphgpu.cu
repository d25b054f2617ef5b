// sphgpu.cu -- the C ABI of include/sphgpu.h: context management, resident state transfers, and the literal-mode
// wrappers mirroring build_tree / densityiterate / cons2prim_everything / force / derivs of the reference.
#include <cstdlib>
#include "common.cuh"
#include <string.h>
#include <new>

KernConsts make_kern_consts(int kernel)
{
    KernConsts k;
    const double pi = 3.14159265358979323846264338327950288419716939937510582097494459;
    if (kernel == 0) { k.radkern = 2.; k.radkern2 = 4.; k.cnormk = 1. / pi; k.wab0 = 1.; k.gradh0 = -3.; k.dphidh0 = 1.4; k.cnormk_drag = 10. / (9. * pi); }
    else { k.radkern = 3.; k.radkern2 = 9.; k.cnormk = 1. / (120. * pi); k.wab0 = 66.; k.gradh0 = -198.; k.dphidh0 = 239. / 210.; k.cnormk_drag = 1. / (168. * pi); }
    return k;
}

static void set_params_internal(sphgpu_ctx *c, const sphgpu_params *p)
{
    c->hp.p = *p;
    c->hp.kc = make_kern_consts(p->kernel);
    c->hp.dxbound = p->xmax - p->xmin; c->hp.dybound = p->ymax - p->ymin; c->hp.dzbound = p->zmax - p->zmin;
    c->hp.nvu = p->isothermal ? 3 : 4;
    c->hp.ngradh = p->gravity ? 2 : 1;
    c->hp.nalpha = p->const_av ? 0 : 3;
}

template <typename T>
static int upload_arr(sphgpu_ctx *c, DevBuf<T> &buf, const T *h, size_t count)
{
    CUDA_TRY(c, buf.ensure(count));
    if (h) CUDA_TRY(c, cudaMemcpyAsync(buf.p, h, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return SPHGPU_OK;
}
template <typename T>
static int download_arr(sphgpu_ctx *c, const DevBuf<T> &buf, T *h, size_t count)
{
    if (h && buf.p) CUDA_TRY(c, cudaMemcpyAsync(h, buf.p, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    return SPHGPU_OK;
}
template <typename T>
static int ensure_zero(sphgpu_ctx *c, DevBuf<T> &buf, size_t count)
{
    if (buf.cap >= count) return SPHGPU_OK;
    CUDA_TRY(c, buf.ensure(count));
    CUDA_TRY(c, cudaMemsetAsync(buf.p, 0, buf.cap * sizeof(T), c->stream));
    return SPHGPU_OK;
}

// every canonical array exists (zero-filled) once npart is known, so kernels never see a null pointer
static int ensure_all(sphgpu_ctx *c, int64_t n)
{
    const int nvu = c->hp.nvu, ng = c->hp.ngradh;
    TRY(ensure_zero(c, c->xyzh, 4 * n)); TRY(ensure_zero(c, c->vxyzu, (size_t)nvu * n)); TRY(ensure_zero(c, c->fxyzu, (size_t)nvu * n));
    TRY(ensure_zero(c, c->fext, 3 * n)); TRY(ensure_zero(c, c->Bevol, 4 * n)); TRY(ensure_zero(c, c->dBevol, 4 * n));
    TRY(ensure_zero(c, c->eos_vars, 7 * n)); TRY(ensure_zero(c, c->divcurlv, n)); TRY(ensure_zero(c, c->divcurlB, 4 * n));
    TRY(ensure_zero(c, c->alphaind, 3 * n)); TRY(ensure_zero(c, c->gradh, (size_t)ng * n)); TRY(ensure_zero(c, c->dvdx, 9 * n));
    TRY(ensure_zero(c, c->poten, n)); TRY(ensure_zero(c, c->divBsymm, n)); TRY(ensure_zero(c, c->iphase, n));
    TRY(ensure_zero(c, c->ibin, n)); TRY(ensure_zero(c, c->ibin_old, n)); TRY(ensure_zero(c, c->ibin_wake, n));
    TRY(ensure_zero(c, c->dustfrac, n)); TRY(ensure_zero(c, c->tstop, n));
    TRY(ensure_zero(c, c->counters, CNT_COUNT)); TRY(ensure_zero(c, c->dscal, DS_COUNT));
    return SPHGPU_OK;
}

// grow every canonical array to n particles keeping the first `keep` (ghost append)
int ensure_all_keep(sphgpu_ctx *c, int64_t n, int64_t keep)
{
    const int nvu = c->hp.nvu, ng = c->hp.ngradh;
    cudaStream_t st = c->stream;
#define GROW(buf, w) CUDA_TRY(c, (buf).ensure_keep((size_t)(w) * n, (size_t)(w) * keep, st))
    GROW(c->xyzh, 4); GROW(c->vxyzu, nvu); GROW(c->fxyzu, nvu); GROW(c->fext, 3); GROW(c->Bevol, 4); GROW(c->dBevol, 4); GROW(c->eos_vars, 7);
    GROW(c->divcurlv, 1); GROW(c->divcurlB, 4); GROW(c->alphaind, 3); GROW(c->gradh, ng); GROW(c->dvdx, 9); GROW(c->poten, 1); GROW(c->divBsymm, 1);
    GROW(c->iphase, 1); GROW(c->ibin, 1); GROW(c->ibin_old, 1); GROW(c->ibin_wake, 1); GROW(c->dustfrac, 1); GROW(c->tstop, 1);
#undef GROW
    return SPHGPU_OK;
}

// ---- microbenchmarks for the roofline denominators ----------------------------------------------
__global__ void k_dfma_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double b = 1.0000001, cc = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
        a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_copy(const double4 *__restrict__ a, double4 *__restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

extern "C" {

int sphgpu_create(const sphgpu_params *params, int device, sphgpu_ctx **out)
{
    if (!params || !out) return SPHGPU_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return SPHGPU_ERR_CUDA;   // no CPU fallback: fail loudly
    if (device < 0 || device >= ndev) return SPHGPU_ERR_ARG;
    sphgpu_ctx *c = new (std::nothrow) sphgpu_ctx();
    if (!c) return SPHGPU_ERR_ARG;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return SPHGPU_ERR_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->numSMs = prop.multiProcessorCount;
    cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int k = 0; k < 16; k++) cudaEventCreate(&c->ev[k]);
    if (const char *e = getenv("SPHGPU_GROUP_PACK")) c->group_pack = atoi(e) < 0 ? 0 : (atoi(e) > 4096 ? 4096 : atoi(e));   // A/B switch for whole test runs
    set_params_internal(c, params);
    *out = c;
    return SPHGPU_OK;
}

void sphgpu_destroy(sphgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    sphgpu_dist_finalize(c);
    c->gid.release();
    c->xyzh.release(); c->vxyzu.release(); c->fxyzu.release(); c->fext.release(); c->Bevol.release(); c->dBevol.release(); c->eos_vars.release(); c->Bxyz.release();
    c->divcurlv.release(); c->divcurlB.release(); c->alphaind.release(); c->gradh.release(); c->dvdx.release(); c->poten.release(); c->divBsymm.release();
    c->iphase.release(); c->ibin.release(); c->ibin_old.release(); c->ibin_wake.release();
    c->keys.release(); c->keys_alt.release(); c->perm.release(); c->perm_alt.release(); c->pos4.release(); c->vel4.release(); c->acc4.release(); c->bev4.release();
    c->stype.release(); c->hnew.release(); c->frecC.release(); c->frecD.release(); c->frecE.release(); c->frec.release(); c->drec.release(); c->v_true.release(); c->B_true.release(); c->twas.release(); c->forc_tab.release();
    c->s_gradh.release(); c->s_divv.release(); c->s_dvdx.release(); c->s_alpha3.release(); c->s_divcurlB.release(); c->s_fxyzu.release(); c->s_dB.release();
    c->s_ibin.release(); c->s_ibinold.release(); c->s_wake.release(); c->s_ibinnew.release(); c->s_gsoft.release(); c->s_tstop.release(); c->s_dustfrac.release(); c->dustfrac.release(); c->tstop.release(); c->gacc.release(); c->s_divvf.release(); c->s_poten.release(); c->s_divBsymm.release(); c->s_nneigh.release();
    c->cpl.release(); c->cellflag.release(); c->cellid_scan.release(); c->cells.release(); c->groups.release(); c->cellkeys.release(); c->nodes.release(); c->nodeflag.release();
    c->halo_sendidx.release(); c->halo_cnt.release(); c->halo_boxes.release(); c->halo_sendbuf.release(); c->halo_recvbuf.release();
    c->cubtemp.release(); c->scratch.release(); c->nodesf.release(); c->stage_idx.release(); c->counters.release(); c->dscal.release();
    for (int k = 0; k < 16; k++) cudaEventDestroy(c->ev[k]);
    if (c->copy_in) { cudaStreamDestroy(c->copy_in); cudaStreamDestroy(c->copy_out); for (int k = 0; k < 6; k++) cudaEventDestroy(c->cev[k]); }
    gravity_release(c); c->h_build.release(); c->h_hist.release(); c->h_its.release(); c->ref_nodes.release(); c->ref_leaf.release(); c->ref_leaf_sorted.release();
    cudaStreamDestroy(c->stream);
    delete c;
}

int sphgpu_set_params(sphgpu_ctx *c, const sphgpu_params *p)
{
    if (!c || !p) return SPHGPU_ERR_ARG;
    const bool layout_change = (p->isothermal != c->hp.p.isothermal) || (p->gravity != c->hp.p.gravity);
    set_params_internal(c, p);
    if (layout_change) { c->vxyzu.release(); c->fxyzu.release(); c->gradh.release(); c->tree_valid = false; }
    return SPHGPU_OK;
}

const char *sphgpu_last_error(sphgpu_ctx *c) { return c ? c->err.c_str() : "null context"; }

int sphgpu_set_option(sphgpu_ctx *c, const char *name, double value)
{
    if (!c || !name) return SPHGPU_ERR_ARG;
    if (!strcmp(name, "max_cell")) { int v = (int)value; c->max_cell = v < 1 ? 1 : (v > 32 ? 32 : v); c->tree_valid = false; return 0; }
    if (!strcmp(name, "group_pack")) { int v = (int)value; c->group_pack = v < 0 ? 0 : (v > 4096 ? 4096 : v); c->tree_valid = false; return 0; }
    if (!strcmp(name, "max_leaf")) { int v = (int)value; c->max_leaf = v < 1 ? 1 : (v > 32 ? 32 : v); c->tree_valid = false; return 0; }
    if (!strcmp(name, "no_iso1")) { c->no_iso1 = value != 0.; return 0; }
    if (!strcmp(name, "list_margin")) { c->list_margin = value < 1. ? 1. : value; return 0; }
    if (!strcmp(name, "hilbert")) { c->hilbert = value != 0.; c->tree_valid = false; return 0; }
    if (!strcmp(name, "halo_hgrow")) { c->halo_hgrow = value; return 0; }
    if (!strcmp(name, "always_refit")) { c->always_refit = value != 0.; return 0; }
    if (!strcmp(name, "refcompat_hmax")) { c->refcompat = value < 0. ? -1 : (value != 0. ? 1 : 0); c->ref_valid = false; return 0; }
    if (!strcmp(name, "force_general")) { c->force_general = value != 0.; return 0; }
    if (!strcmp(name, "grav_p2p_per_particle")) { c->grav_p2p_per_particle = (int)value < 8 ? 8 : (int)value; return 0; }
    if (!strcmp(name, "legacy_stream")) {
        // Multi-GPU plumbing (torch.distributed) issues its collectives relative to the legacy default stream.  A BLOCKING compute stream
        // is implicitly ordered with it, so the halo pack / exchange / unpack sequence needs no host synchronisation in between.
        if (value != 0. && !c->stream_blocking) {
            cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream);
            if (cudaStreamCreate(&c->stream) != cudaSuccess) return SPHGPU_ERR_CUDA;
            c->stream_blocking = true;
        }
        return 0;
    }
    if (!strcmp(name, "scratch_per_warp")) { c->scratch_per_warp = (int)value; c->stage_idx.release(); return 0; }
    return SPHGPU_ERR_ARG;
}

int sphgpu_set_timestep_bins(sphgpu_ctx *c, int nbinmax, int ibinnow, int istepfrac)
{
    if (!c) return SPHGPU_ERR_ARG;
    c->nbinmax = nbinmax; c->ibinnow = ibinnow; c->istepfrac = istepfrac;
    return SPHGPU_OK;
}

int sphgpu_get_timings(sphgpu_ctx *c, double *ms4)
{
    if (!c || !ms4) return SPHGPU_ERR_ARG;
    for (int k = 0; k < 4; k++) ms4[k] = c->ms_phase[k];
    return SPHGPU_OK;
}
int64_t sphgpu_launch_count(sphgpu_ctx *c) { return c ? c->launches : 0; }
int sphgpu_get_gravity_timings(sphgpu_ctx *c, double *ms2)
{
    if (!c || !ms2) return SPHGPU_ERR_ARG;
    ms2[0] = c->ms_gravity[0]; ms2[1] = c->ms_gravity[1];
    return SPHGPU_OK;
}
int sphgpu_gravity_gather_pack(sphgpu_ctx *c, void **sendptr_device, int *record_doubles)
{
    if (!c || !sendptr_device || !record_doubles) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    return gravity_gather_pack(c, sendptr_device, record_doubles);
}
int sphgpu_gravity_gather_recvbuf(sphgpu_ctx *c, int nranks, int64_t stride, void **recvptr_device)
{
    if (!c || !recvptr_device) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    return gravity_gather_recvbuf(c, nranks, stride, recvptr_device);
}
int sphgpu_gravity_gather_unpack(sphgpu_ctx *c, int nranks, int myrank, int64_t stride, const int64_t *counts)
{
    if (!c || !counts) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    return gravity_gather_unpack(c, nranks, myrank, stride, counts);
}
int64_t sphgpu_gravity_tree(sphgpu_ctx *c, int64_t maxnodes, double *rec12, int32_t *irec6, int32_t *ids)
{
    if (!c) return -1;
    cudaSetDevice(c->device);
    return gravity_tree_dump(c, maxnodes, rec12, irec6, ids);
}
int sphgpu_get_kernel_timings(sphgpu_ctx *c, double *ms2)
{
    if (!c || !ms2) return SPHGPU_ERR_ARG;
    ms2[0] = c->ms_kernel[0]; ms2[1] = c->ms_kernel[1];
    return SPHGPU_OK;
}

int sphgpu_upload(sphgpu_ctx *c, const sphgpu_host_arrays *h, uint64_t mask)
{
    sphgpu_dist_mark_dirty(c);
    if (!c || !h || h->npart <= 0) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int64_t n = h->npart;
    if (n != c->npart) { c->tree_valid = false; c->dens_valid = false; }
    c->npart = n; c->nlocal = n; c->nghost = 0;
    TRY(ensure_all(c, n));
    const int nvu = c->hp.nvu, ng = c->hp.ngradh;
    if (mask & SPHGPU_F_XYZH) { TRY(upload_arr(c, c->xyzh, h->xyzh, 4 * n)); if (h->xyzh) c->tree_valid = false; }
    if (mask & SPHGPU_F_VXYZU) TRY(upload_arr(c, c->vxyzu, h->vxyzu, (size_t)nvu * n));
    if (mask & SPHGPU_F_FXYZU) TRY(upload_arr(c, c->fxyzu, h->fxyzu, (size_t)nvu * n));
    if (mask & SPHGPU_F_FEXT) TRY(upload_arr(c, c->fext, h->fext, 3 * n));
    if (mask & SPHGPU_F_BEVOL) TRY(upload_arr(c, c->Bevol, h->Bevol, 4 * n));
    if (mask & SPHGPU_F_DBEVOL) TRY(upload_arr(c, c->dBevol, h->dBevol, 4 * n));
    if (mask & SPHGPU_F_EOSVARS) { TRY(upload_arr(c, c->eos_vars, h->eos_vars, 7 * n)); c->eos_on_device = false; }
    if (mask & SPHGPU_F_DIVCURLV) TRY(upload_arr(c, c->divcurlv, h->divcurlv, n));
    if (mask & SPHGPU_F_DIVCURLB) TRY(upload_arr(c, c->divcurlB, h->divcurlB, 4 * n));
    if (mask & SPHGPU_F_ALPHAIND) TRY(upload_arr(c, c->alphaind, h->alphaind, 3 * n));
    if (mask & SPHGPU_F_GRADH) TRY(upload_arr(c, c->gradh, h->gradh, (size_t)ng * n));
    if (mask & SPHGPU_F_DVDX) TRY(upload_arr(c, c->dvdx, h->dvdx, 9 * n));
    if (mask & SPHGPU_F_POTEN) TRY(upload_arr(c, c->poten, h->poten, n));
    if (mask & SPHGPU_F_DIVBSYMM) TRY(upload_arr(c, c->divBsymm, h->divBsymm, n));
    if (mask & SPHGPU_F_IPHASE) { TRY(upload_arr(c, c->iphase, h->iphase, n)); if (h->iphase) c->tree_valid = false; }
    if (mask & SPHGPU_F_IBIN) { TRY(upload_arr(c, c->ibin, h->ibin, n)); TRY(upload_arr(c, c->ibin_old, h->ibin_old, n)); TRY(upload_arr(c, c->ibin_wake, h->ibin_wake, n)); }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPHGPU_OK;
}

int sphgpu_download(sphgpu_ctx *c, sphgpu_host_arrays *h, uint64_t mask)
{
    if (!c || !h) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int64_t n = (h->npart > 0 && h->npart < c->npart) ? h->npart : c->npart;    // ghosts (beyond nlocal) are never returned
    const int nvu = c->hp.nvu, ng = c->hp.ngradh;
    if (mask & SPHGPU_F_XYZH) TRY(download_arr(c, c->xyzh, h->xyzh, 4 * n));
    if (mask & SPHGPU_F_VXYZU) TRY(download_arr(c, c->vxyzu, h->vxyzu, (size_t)nvu * n));
    if (mask & SPHGPU_F_FXYZU) TRY(download_arr(c, c->fxyzu, h->fxyzu, (size_t)nvu * n));
    if (mask & SPHGPU_F_FEXT) TRY(download_arr(c, c->fext, h->fext, 3 * n));
    if (mask & SPHGPU_F_BEVOL) TRY(download_arr(c, c->Bevol, h->Bevol, 4 * n));
    if (mask & SPHGPU_F_DBEVOL) TRY(download_arr(c, c->dBevol, h->dBevol, 4 * n));
    if (mask & SPHGPU_F_EOSVARS) TRY(download_arr(c, c->eos_vars, h->eos_vars, 7 * n));
    if (mask & SPHGPU_F_DIVCURLV) TRY(download_arr(c, c->divcurlv, h->divcurlv, n));
    if (mask & SPHGPU_F_DIVCURLB) TRY(download_arr(c, c->divcurlB, h->divcurlB, 4 * n));
    if (mask & SPHGPU_F_ALPHAIND) TRY(download_arr(c, c->alphaind, h->alphaind, 3 * n));
    if (mask & SPHGPU_F_GRADH) TRY(download_arr(c, c->gradh, h->gradh, (size_t)ng * n));
    if (mask & SPHGPU_F_DVDX) TRY(download_arr(c, c->dvdx, h->dvdx, 9 * n));
    if (mask & SPHGPU_F_POTEN) TRY(download_arr(c, c->poten, h->poten, n));
    if (mask & SPHGPU_F_DIVBSYMM) TRY(download_arr(c, c->divBsymm, h->divBsymm, n));
    if (mask & SPHGPU_F_IPHASE) TRY(download_arr(c, c->iphase, h->iphase, n));
    if (mask & SPHGPU_F_IBIN) { TRY(download_arr(c, c->ibin, h->ibin, n)); TRY(download_arr(c, c->ibin_wake, h->ibin_wake, n)); }
    if (mask & SPHGPU_F_DUSTFRAC) TRY(download_arr(c, c->dustfrac, h->dustfrac, n));
    if (mask & SPHGPU_F_TSTOP) TRY(download_arr(c, c->tstop, h->tstop, n));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPHGPU_OK;
}

// ---- resident mode --------------------------------------------------------------------------------
int sphgpu_build_tree_resident(sphgpu_ctx *c)
{
    if (!c) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    return tree_build(c);
}
int sphgpu_densityiterate_resident(sphgpu_ctx *c, int icall, sphgpu_scalars *out)
{
    if (!c) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    return density_run(c, icall, out);
}
int sphgpu_cons2prim_resident(sphgpu_ctx *c)
{
    if (!c) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    return cons2prim_run(c);
}
int sphgpu_force_resident(sphgpu_ctx *c, int icall, double dt, sphgpu_scalars *out)
{
    if (!c) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    return force_run(c, icall, dt, out);
}

// derivs (deriv.f90:113-232): icall = 1 tree + density + cons2prim + force ; icall = 2 cons2prim + force only
int sphgpu_derivs_resident(sphgpu_ctx *c, int icall, double dt, sphgpu_scalars *out)
{
    if (!c || icall < 0 || icall > 2) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaEventRecord(c->ev[0], c->stream);
    if (icall == 1 || icall == 0) TRY(tree_build(c));
    cudaEventRecord(c->ev[1], c->stream);
    if (icall == 1) {
        TRY(density_run(c, 1, nullptr));
        c->hp.p.set_boundaries_to_active = 0;                     // deriv.f90:146
    }
    cudaEventRecord(c->ev[2], c->stream);
    TRY(cons2prim_run(c));
    cudaEventRecord(c->ev[3], c->stream);
    if (c->hp.p.driving) TRY(sphgpu_forcing_resident(c));              // forceit (deriv.f90:178-182)
    TRY(force_run(c, icall, dt, out));
    cudaEventRecord(c->ev[4], c->stream);
    cudaEventSynchronize(c->ev[4]);
    for (int k = 0; k < 4; k++) { float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[k], c->ev[k + 1]); c->ms_phase[k] = ms; }
    return SPHGPU_OK;
}

// ---- literal mode -----------------------------------------------------------------------------------
int sphgpu_build_tree(sphgpu_ctx *c, int64_t npart, int64_t nactive, double *xyzh, const double *vxyzu, const int8_t *iphase)
{
    (void)nactive; (void)vxyzu;
    if (!c || !xyzh || !iphase) return SPHGPU_ERR_ARG;
    sphgpu_host_arrays h; memset(&h, 0, sizeof h);
    h.npart = npart; h.xyzh = xyzh; h.iphase = const_cast<int8_t *>(iphase);
    TRY(sphgpu_upload(c, &h, SPHGPU_F_XYZH | SPHGPU_F_IPHASE));
    TRY(tree_build(c));
    return sphgpu_download(c, &h, SPHGPU_F_XYZH);                 // xyzh is inout: periodic wrap (kdtree.F90:387)
}

int sphgpu_densityiterate(sphgpu_ctx *c, int icall, int64_t npart, int64_t nactive, double *xyzh, const double *vxyzu, float *divcurlv,
                          float *divcurlB, const double *Bevol, double *stressmax, const double *fxyzu, const double *fext, float *alphaind,
                          float *gradh, float *dvdx, const int8_t *iphase, sphgpu_scalars *out)
{
    (void)nactive; (void)iphase;
    if (!c || npart != c->npart) { if (c) c->err = "densityiterate: npart differs from the tree"; return SPHGPU_ERR_ARG; }
    sphgpu_host_arrays h; memset(&h, 0, sizeof h);
    h.npart = npart; h.vxyzu = const_cast<double *>(vxyzu); h.fxyzu = const_cast<double *>(fxyzu); h.fext = const_cast<double *>(fext);
    h.Bevol = const_cast<double *>(Bevol); h.alphaind = alphaind; h.gradh = gradh; h.divcurlv = divcurlv; h.divcurlB = divcurlB; h.dvdx = dvdx;
    h.xyzh = xyzh;
    // xyzh itself stays as uploaded by build_tree (the tree owns the wrapped positions); everything else goes in
    uint64_t in = SPHGPU_F_VXYZU | SPHGPU_F_FXYZU | SPHGPU_F_FEXT | SPHGPU_F_ALPHAIND | SPHGPU_F_GRADH | SPHGPU_F_DIVCURLV | SPHGPU_F_DVDX;
    if (c->hp.p.mhd) in |= SPHGPU_F_BEVOL | SPHGPU_F_DIVCURLB;
    h.xyzh = nullptr;
    TRY(sphgpu_upload(c, &h, in));
    TRY(density_run(c, icall, out));
    if (stressmax) *stressmax = 0.;                               // get_max_stress zeroes it unconditionally (dens.F90:1075-1076)
    h.xyzh = xyzh;
    uint64_t outm = SPHGPU_F_XYZH | SPHGPU_F_GRADH | SPHGPU_F_DIVCURLV | SPHGPU_F_DVDX | SPHGPU_F_ALPHAIND;
    if (c->hp.p.mhd) outm |= SPHGPU_F_DIVCURLB;
    return sphgpu_download(c, &h, outm);
}

int sphgpu_cons2prim_everything(sphgpu_ctx *c, int64_t npart, const double *xyzh, const double *vxyzu, const float *dvdx, double *eos_vars,
                                const double *Bevol, double *Bxyz, float *alphaind, const int8_t *iphase)
{
    (void)xyzh; (void)iphase;
    if (!c || npart != c->npart) { if (c) c->err = "cons2prim: npart differs from the resident state"; return SPHGPU_ERR_ARG; }
    sphgpu_host_arrays h; memset(&h, 0, sizeof h);
    h.npart = npart; h.vxyzu = const_cast<double *>(vxyzu); h.dvdx = const_cast<float *>(dvdx); h.alphaind = alphaind; h.Bevol = const_cast<double *>(Bevol);
    h.eos_vars = eos_vars;
    const bool tv = c->tree_valid;
    TRY(sphgpu_upload(c, &h, SPHGPU_F_VXYZU | SPHGPU_F_DVDX | SPHGPU_F_ALPHAIND | (c->hp.p.mhd ? SPHGPU_F_BEVOL : 0)));
    c->tree_valid = tv;
    TRY(cons2prim_run(c));
    TRY(sphgpu_download(c, &h, SPHGPU_F_EOSVARS | SPHGPU_F_ALPHAIND));
    if (Bxyz && c->hp.p.mhd) { CUDA_TRY(c, cudaMemcpy(Bxyz, c->Bxyz.p, sizeof(double) * 3 * npart, cudaMemcpyDeviceToHost)); }
    return SPHGPU_OK;
}

int sphgpu_force(sphgpu_ctx *c, int icall, int64_t npart, const double *xyzh, const double *vxyzu, double *fxyzu, float *divcurlv,
                 const float *divcurlB, const double *Bevol, double *dBevol, const double *fext, double dt, double stressmax, const double *eos_vars,
                 const float *alphaind, const float *gradh, const float *dvdx, const int8_t *iphase, float *poten, float *divBsymm, sphgpu_scalars *out)
{
    (void)xyzh; (void)stressmax; (void)iphase; (void)fext;
    if (!c || npart != c->npart) { if (c) c->err = "force: npart differs from the tree"; return SPHGPU_ERR_ARG; }
    sphgpu_host_arrays h; memset(&h, 0, sizeof h);
    h.npart = npart; h.vxyzu = const_cast<double *>(vxyzu); h.fxyzu = fxyzu; h.divcurlv = divcurlv; h.divcurlB = const_cast<float *>(divcurlB);
    h.Bevol = const_cast<double *>(Bevol); h.dBevol = dBevol; h.eos_vars = const_cast<double *>(eos_vars); h.alphaind = const_cast<float *>(alphaind);
    h.gradh = const_cast<float *>(gradh); h.dvdx = const_cast<float *>(dvdx); h.poten = poten; h.divBsymm = divBsymm;
    const bool tv = c->tree_valid;
    uint64_t in = SPHGPU_F_VXYZU | SPHGPU_F_FXYZU | SPHGPU_F_DIVCURLV | SPHGPU_F_EOSVARS | SPHGPU_F_ALPHAIND | SPHGPU_F_GRADH | SPHGPU_F_DVDX;
    if (c->hp.p.mhd) in |= SPHGPU_F_BEVOL | SPHGPU_F_DIVCURLB;
    TRY(sphgpu_upload(c, &h, in));
    c->tree_valid = tv;
    TRY(force_run(c, icall, dt, out));
    uint64_t outm = SPHGPU_F_FXYZU | SPHGPU_F_DIVCURLV;
    if (c->hp.p.mhd) outm |= SPHGPU_F_DBEVOL | SPHGPU_F_DIVBSYMM;
    if (c->hp.p.gravity) outm |= SPHGPU_F_POTEN;
    return sphgpu_download(c, &h, outm);
}

// derivs with host arrays in and out.  The copies are PIPELINED against the passes on two copy streams (pinned host buffers make
// them truly asynchronous): positions go first and the tree builds while v, f, fext, alpha follow; the density outputs go back
// while cons2prim and the force pass run, eos_vars/alpha during the force pass, and only fxyzu, divv (+ dB/dt, poten, bins)
// after it.  Arrays that the passes overwrite for EVERY particle (gradh, divcurlv, dvdx, eos_vars) are uploaded only when some
// particle is inactive or a boundary particle and therefore keeps its stored values (test_derivs.f90:233-238).
int sphgpu_derivs(sphgpu_ctx *c, int icall, sphgpu_host_arrays *h, double dt, sphgpu_scalars *out)
{
    if (!c || !h || h->npart <= 0 || icall < 0 || icall > 2) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const sphgpu_params &p = c->hp.p;
    const int64_t n = h->npart;
    const int nvu = c->hp.nvu, ng = c->hp.ngradh;
    if (!c->copy_in) { CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking)); CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
                       for (int k = 0; k < 6; k++) CUDA_TRY(c, cudaEventCreateWithFlags(&c->cev[k], cudaEventDisableTiming)); }
    if (n != c->npart) { c->tree_valid = false; c->dens_valid = false; }
    c->npart = n; c->nlocal = n; c->nghost = 0;
    TRY(ensure_all(c, n));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));           // ensure_all may have zero-filled new buffers on the compute stream
    bool all_active = (h->iphase != nullptr) && !p.set_boundaries_to_active && icall == 1;
    if (all_active) { const int8_t *ip = h->iphase; for (int64_t i = 0; i < n; i++) if (ip[i] != IGAS) { all_active = false; break; } }
    cudaStream_t si = c->copy_in, so = c->copy_out;
    c->bytes_h2d = c->bytes_d2h = 0;
#define H2D(buf, ptr, count) do { if (ptr) { CUDA_TRY(c, cudaMemcpyAsync((buf).p, ptr, sizeof(*(buf).p) * (size_t)(count), cudaMemcpyHostToDevice, si)); c->bytes_h2d += (int64_t)(sizeof(*(buf).p) * (size_t)(count)); } } while (0)
#define D2H(buf, ptr, count) do { if ((ptr) && (buf).p) { CUDA_TRY(c, cudaMemcpyAsync(ptr, (buf).p, sizeof(*(buf).p) * (size_t)(count), cudaMemcpyDeviceToHost, so)); c->bytes_d2h += (int64_t)(sizeof(*(buf).p) * (size_t)(count)); } } while (0)
    // ---- wave 1: what the tree needs
    if (icall != 2) { H2D(c->xyzh, h->xyzh, 4 * n); H2D(c->iphase, h->iphase, n); c->tree_valid = false; }
    CUDA_TRY(c, cudaEventRecord(c->cev[0], si));
    // ---- wave 2: what the density pass reads or writes (alphaind: it stores div a in the third column)
    H2D(c->vxyzu, h->vxyzu, (size_t)nvu * n); H2D(c->fxyzu, h->fxyzu, (size_t)nvu * n); H2D(c->fext, h->fext, 3 * n); H2D(c->alphaind, h->alphaind, 3 * n);
    if (p.mhd) { H2D(c->Bevol, h->Bevol, 4 * n); }
    auto upload_stored = [&]() -> int {      // arrays that keep their stored values for particles no pass writes
        H2D(c->gradh, h->gradh, (size_t)ng * n); H2D(c->divcurlv, h->divcurlv, n); H2D(c->dvdx, h->dvdx, 9 * n);
        if (p.mhd) H2D(c->divcurlB, h->divcurlB, 4 * n);
        return SPHGPU_OK;
    };
    if (!all_active) TRY(upload_stored());
    CUDA_TRY(c, cudaEventRecord(c->cev[1], si));
    // ---- wave 3: first read by cons2prim / force; travels while the density pass runs.
    // eos_vars always goes: cons2prim writes the rows igasP, ics, itemp only, the host's imu, iX, iZ, igamma must survive the download
    H2D(c->eos_vars, h->eos_vars, 7 * n);
    if (p.ind_timesteps) { H2D(c->ibin, h->ibin, n); H2D(c->ibin_old, h->ibin_old, n); H2D(c->ibin_wake, h->ibin_wake, n); }
    CUDA_TRY(c, cudaEventRecord(c->cev[5], si));
    // ---- passes
    cudaEventRecord(c->ev[0], c->stream);
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->cev[0], 0));
    if (icall == 1 || icall == 0) TRY(tree_build(c));
    if (all_active && c->nlive < n) {        // dead / accreted particles (h <= 0) among "all gas": their stored values must survive too
        TRY(upload_stored());
        CUDA_TRY(c, cudaEventRecord(c->cev[1], si));
        CUDA_TRY(c, cudaEventRecord(c->cev[5], si));
    }
    cudaEventRecord(c->ev[1], c->stream);
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->cev[1], 0));
    if (icall == 1) {
        TRY(density_run(c, 1, nullptr));
        c->hp.p.set_boundaries_to_active = 0;                     // deriv.f90:146
    }
    cudaEventRecord(c->ev[2], c->stream);
    CUDA_TRY(c, cudaEventRecord(c->cev[2], c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(so, c->cev[2], 0));          // density outputs go home while cons2prim and force run
    D2H(c->xyzh, h->xyzh, 4 * n); D2H(c->gradh, h->gradh, (size_t)ng * n); D2H(c->dvdx, h->dvdx, 9 * n);
    if (p.mhd) D2H(c->divcurlB, h->divcurlB, 4 * n);
    if (p.dust) D2H(c->dustfrac, h->dustfrac, n);
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->cev[5], 0));
    TRY(cons2prim_run(c));
    cudaEventRecord(c->ev[3], c->stream);
    CUDA_TRY(c, cudaEventRecord(c->cev[3], c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(so, c->cev[3], 0));
    D2H(c->eos_vars, h->eos_vars, 7 * n); D2H(c->alphaind, h->alphaind, 3 * n);
    if (p.driving) TRY(sphgpu_forcing_resident(c));              // forceit (deriv.f90:178-182)
    TRY(force_run(c, icall, dt, out));
    cudaEventRecord(c->ev[4], c->stream);
    CUDA_TRY(c, cudaEventRecord(c->cev[4], c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(so, c->cev[4], 0));
    D2H(c->fxyzu, h->fxyzu, (size_t)nvu * n); D2H(c->divcurlv, h->divcurlv, n);
    if (p.mhd) { D2H(c->dBevol, h->dBevol, 4 * n); D2H(c->divBsymm, h->divBsymm, n); }
    if (p.gravity) D2H(c->poten, h->poten, n);
    if (p.dust) D2H(c->tstop, h->tstop, n);
    if (p.ind_timesteps) { D2H(c->ibin, h->ibin, n); D2H(c->ibin_wake, h->ibin_wake, n); }
#undef H2D
#undef D2H
    CUDA_TRY(c, cudaStreamSynchronize(so));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; k++) { float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[k], c->ev[k + 1]); c->ms_phase[k] = ms; }
    return SPHGPU_OK;
}

int sphgpu_get_copy_bytes(sphgpu_ctx *c, int64_t *h2d, int64_t *d2h)
{
    if (!c || !h2d || !d2h) return SPHGPU_ERR_ARG;
    *h2d = c->bytes_h2d; *d2h = c->bytes_d2h;
    return SPHGPU_OK;
}

int sphgpu_get_neighbour_stats(sphgpu_ctx *c, sphgpu_scalars *out)
{
    if (!c || !out) return SPHGPU_ERR_ARG;
    *out = c->last_dens;
    return SPHGPU_OK;
}

int64_t sphgpu_neighbour_sets(sphgpu_ctx *c, int symmetric, int64_t *offsets, int32_t *list, int64_t maxlist)
{
    if (!c) return -1;
    cudaSetDevice(c->device);
    return neighbour_sets_run(c, symmetric, offsets, list, maxlist);
}

double sphgpu_measure_fp64_peak(sphgpu_ctx *c)
{
    if (!c) return -1.;
    cudaSetDevice(c->device);
    const int blocks = c->numSMs * 8, threads = 256, iters = 20000;
    double *d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return -1.;
    double best = 0.;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(c->ev[5], c->stream);
        k_dfma_peak<<<blocks, threads, 0, c->stream>>>(d, iters);
        cudaEventRecord(c->ev[6], c->stream);
        cudaEventSynchronize(c->ev[6]);
        float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
        const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(d);
    return best;
}

double sphgpu_measure_copy_bw(sphgpu_ctx *c)
{
    if (!c) return -1.;
    cudaSetDevice(c->device);
    const size_t n = (size_t)1 << 25;   // 32 Mi double4 = 1 GiB per buffer
    double4 *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, n * sizeof(double4)) != cudaSuccess) return -1.;
    if (cudaMalloc(&b, n * sizeof(double4)) != cudaSuccess) { cudaFree(a); return -1.; }
    cudaMemsetAsync(a, 0, n * sizeof(double4), c->stream);
    double best = 0.;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(c->ev[5], c->stream);
        k_copy<<<c->numSMs * 16, 512, 0, c->stream>>>(a, b, n);
        cudaEventRecord(c->ev[6], c->stream);
        cudaEventSynchronize(c->ev[6]);
        float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
        const double gbs = 2.0 * n * sizeof(double4) / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbs > best) best = gbs;
    }
    cudaFree(a); cudaFree(b);
    return best;
}

}  // extern "C"
