// dist.cu -- the multi-GPU path behind the C ABI: one process (or host thread) per GPU, NCCL over NVLink / NVSwitch.
//
// Replaces, for the hot path, the reference's MPI layer:
//   balancedomains (src/main/mpi_balance.F90:82-173)            -> sphgpu_dist_migrate: particles that left their rank's box are handed
//                                                                   to the new owner with their whole record, holes filled from the tail
//   maketreeglobal's domain split (src/main/kdtree.F90:2044-2300) -> sphgpu_dist_rebalance: recursive bisection at the centre of mass along
//                                                                   the longest axis, moments all-reduced level by level
//   send_cell / recv_cells / combine_cells (mpi_derivs.F90:197-522, mpi_dens.F90, mpi_force.F90)
//                                                                -> sphgpu_dist_derivs: ghost-particle halo, two exchanges per derivs(1)
//   reduceall_mpi (force.F90:848-852, dens.F90:546-549)          -> one all-reduce for the scalars of a derivs call
//   step (step_leapfrog.f90:95-760) on a decomposed set          -> sphgpu_dist_step
//
// Exchange protocol.  Every pair of ranks (s -> r) owns a FIXED-CAPACITY block {count, capacity, records[capacity]} whose capacity both
// sides derive from the count of the previous successful exchange (x1.3 + 64), so a step needs neither a count exchange nor a host read
// of the counts: selection, packing, the grouped ncclSend/ncclRecv and unpacking are queued on the context's stream back to back.  The
// receiver sizes its arrays for the sum of the capacities; slots beyond the received counts are dead particles (h = 0) that the tree
// drops like accreted ones (part.F90:931).  A block that overflowed is detected from its header in the same all-reduce that checks the
// halo width; both sides then learn the true counts and the exchange is repeated (first call, or a sudden change of the halo).
// NCCL is loaded with dlopen("libnccl.so.2") when sphgpu_dist_init is first called -- a single-GPU user needs no NCCL at all, and a host
// that already carries an NCCL (torch.distributed, an MPI+NCCL Fortran driver) shares that copy.
#include "common.cuh"
#include <dlfcn.h>
#include <float.h>
#include <nccl.h>
#include <string.h>
#include <algorithm>

int sphgpu_dist_hook_derivs(sphgpu_ctx *c, int icall, double dt, sphgpu_scalars *out);

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api(std::string &err)
{
    static NcclApi api;
    if (api.handle) return &api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (!api.handle) { err = std::string("dist: cannot load NCCL: ") + dlerror(); return nullptr; }
#define LOAD(field, sym) do { *(void **)(&api.field) = dlsym(api.handle, sym); if (!api.field) { err = std::string("dist: NCCL symbol missing: ") + sym; api.handle = nullptr; return nullptr; } } while (0)
    LOAD(GetUniqueId, "ncclGetUniqueId"); LOAD(CommInitRank, "ncclCommInitRank"); LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(GroupStart, "ncclGroupStart"); LOAD(GroupEnd, "ncclGroupEnd"); LOAD(Send, "ncclSend"); LOAD(Recv, "ncclRecv");
    LOAD(AllReduce, "ncclAllReduce"); LOAD(AllGather, "ncclAllGather"); LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    return &api;
}

}  // namespace

#define NREC 66            // doubles per migrating particle (k_mig_pack)
#define HDR 2              // doubles in front of every block: {count, capacity}

struct DistState {
    NcclApi *api = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    std::vector<double> boxes;                 // 6 per rank {lo, hi}
    double glo[3] = {0, 0, 0}, ghi[3] = {0, 0, 0};   // the decomposed domain (periodic box, or the bounding box handed to set_boxes)
    DevBuf<double> d_boxes;
    // ghost blocks
    std::vector<long long> cap_send, cap_recv, off_send, off_recv;     // records per peer; record offsets of the blocks
    bool caps_valid = false;
    DevBuf<long long> d_caps;                  // [0..P) cap_send, [P..2P) off_send, [2P..3P) cap_recv, [3P..4P) off_recv
    DevBuf<unsigned long long> cnt;            // [0..P) selected per peer (may exceed the capacity), [P..2P) received counts of the last stage 1
    DevBuf<int> sendidx;                       // [sum cap_send]
    DevBuf<double> sendbuf, recvbuf;
    DevBuf<double> red, red_out;               // all-reduce staging
    double hu_prev = -1., overhang = 0., margin = 1.15;
    bool geom_dirty = true;                    // owned positions / h / boxes changed since the halo width inputs were last reduced
    long long ncap_ghost = 0;
    int64_t halo_bytes = 0;
    int rounds = 0;
    // migration
    long long mig_cap = 1024;                  // records per migration block: the SAME on every rank (block sizes must match pairwise)
    long long mig_want = 0;                    // largest number of leavers this rank had for one peer since the capacity was last agreed
    DevBuf<unsigned long long> mcnt;           // [0..P) leaving per peer, [P] holes, [P+1] sources
    DevBuf<int> midx, mflag, mhole, msrc;
    DevBuf<double> msend, mrecv, mtmp;
    int64_t nmigrated_last = 0;
    cudaEvent_t ev[2];
    double ms_exchange = 0.;
};

namespace {

#define NCCL_TRY(c, d, call)                                                                                          \
    do {                                                                                                              \
        ncclResult_t r__ = (call);                                                                                    \
        if (r__ != ncclSuccess) { (c)->err = std::string(#call) + ": " + (d)->api->GetErrorString(r__); return SPHGPU_ERR_CUDA; } \
    } while (0)

static inline int nblk(int64_t n, int b) { return (int)((n + b - 1) / b); }
static inline long long grow_cap(long long count) { return std::max<long long>(256, (long long)(1.3 * (double)count) + 64); }

__device__ __forceinline__ double axis_gap(double x, double lo, double hi, double L, bool periodic)
{
    double g = fmax(0., fmax(lo - x, x - hi));
    if (periodic) {
        const double g2 = fmax(0., fmax(lo - (x - L), (x - L) - hi));
        const double g3 = fmax(0., fmax(lo - (x + L), (x + L) - hi));
        g = fmin(g, fmin(g2, g3));
    }
    return g;
}

// owned particles within dhalo of each foreign box: slot k of peer r's list, as long as it fits the block
__global__ void k_dist_select(int64_t nlocal, const double *__restrict__ xyzh, int nranks, int myrank, const double *__restrict__ boxes, double dhalo,
                              double Lx, double Ly, double Lz, int periodic, unsigned long long *cnt, const long long *caps, int *sendidx)
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
            const long long k = (long long)atomicAdd(&cnt[r], 1ull);
            if (k < caps[r]) sendidx[caps[nranks + r] + k] = (int)i;
        }
    }
}

// stage 1: {x,y,z,h, v(3),u, f+fext(3), B/rho(3),psi, iphase}; stage 2: {h, gradh, alpha, gradsoft}; stage 3 (derivs(2)): {v(3),u, B/rho(3),psi}
struct PackArgs {
    const double *xyzh, *vxyzu, *fxyzu, *fext, *Bevol; const float *gradh, *alphaind; const int8_t *iphase;
    int nvu, mhd, ngradh, nranks, stage, rd;
    const unsigned long long *cnt; const long long *caps; const int *sendidx; double *out;
};
__global__ void k_dist_pack(const PackArgs a)
{
    const int r = blockIdx.y;
    const long long cap = a.caps[r], off = a.caps[a.nranks + r];
    const long long n = min((long long)a.cnt[r], cap);
    double *blk = a.out + (size_t)HDR * r + (size_t)a.rd * off;
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k == 0) { blk[0] = (double)a.cnt[r]; blk[1] = (double)cap; }
    if (k >= n) return;
    const int i = a.sendidx[off + k];
    double *o = blk + HDR + (size_t)a.rd * k;
    if (a.stage == 1) {
        const double4 x = reinterpret_cast<const double4 *>(a.xyzh)[i];
        o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
        const double *v = a.vxyzu + (size_t)a.nvu * i, *f = a.fxyzu + (size_t)a.nvu * i, *fe = a.fext + 3 * (size_t)i;
        o[4] = v[0]; o[5] = v[1]; o[6] = v[2]; o[7] = a.nvu >= 4 ? v[3] : 0.;
        o[8] = f[0] + fe[0]; o[9] = f[1] + fe[1]; o[10] = f[2] + fe[2];
        if (a.mhd) { const double4 B = reinterpret_cast<const double4 *>(a.Bevol)[i]; o[11] = B.x; o[12] = B.y; o[13] = B.z; o[14] = B.w; }
        else { o[11] = o[12] = o[13] = o[14] = 0.; }
        o[15] = (double)a.iphase[i];
    } else if (a.stage == 2) {
        o[0] = a.xyzh[4 * (size_t)i + 3];
        o[1] = (double)a.gradh[(size_t)a.ngradh * i];
        o[2] = (double)a.alphaind[3 * (size_t)i];
        o[3] = a.ngradh > 1 ? (double)a.gradh[(size_t)a.ngradh * i + 1] : 0.;
    } else {
        const double *v = a.vxyzu + (size_t)a.nvu * i;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = a.nvu >= 4 ? v[3] : 0.;
        if (a.mhd) { const double4 B = reinterpret_cast<const double4 *>(a.Bevol)[i]; o[4] = B.x; o[5] = B.y; o[6] = B.z; o[7] = B.w; }
        else { o[4] = o[5] = o[6] = o[7] = 0.; }
    }
}

struct UnpackArgs {
    double *xyzh, *vxyzu, *fxyzu, *fext, *Bevol; float *gradh, *alphaind; int8_t *iphase;
    int nvu, mhd, ngradh, nranks, myrank, stage, rd;
    int64_t nlocal; long long ncap;
    const long long *caps;         // [2P..3P) cap_recv, [3P..4P) off_recv
    unsigned long long *rcnt;      // received counts (written by stage 1, read by the later stages)
    const double *in;
};
// ghosts are compacted behind the owned particles in rank order; the slots up to nlocal + ncap that stay unused are dead particles
__global__ void k_dist_unpack(const UnpackArgs a)
{
    const int r = blockIdx.y;
    const long long *cap_recv = a.caps + 2 * a.nranks, *off_recv = a.caps + 3 * a.nranks;
    long long goff = 0, total = 0;                                  // ghosts received from the ranks before r ; from all ranks
    for (int q = 0; q < a.nranks; q++) {
        long long cq = 0;
        if (q != a.myrank) {
            if (a.stage == 1) cq = min((long long)a.in[(size_t)HDR * q + (size_t)a.rd * off_recv[q]], cap_recv[q]);
            else cq = (long long)a.rcnt[q];
        }
        if (q < r) goff += cq;
        total += cq;
    }
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (a.stage == 1) {
        if (r == a.myrank) {                                          // this block of threads kills the unused tail instead
            const long long j = total + k;
            if (j < a.ncap) {
                const int64_t i = a.nlocal + j;
                reinterpret_cast<double4 *>(a.xyzh)[i] = make_double4(0., 0., 0., 0.);
                a.iphase[i] = (int8_t)(-IGAS);
            }
            if (k == 0) a.rcnt[a.nranks] = (unsigned long long)total;
            return;
        }
        if (k == 0) a.rcnt[r] = (unsigned long long)min((long long)a.in[(size_t)HDR * r + (size_t)a.rd * off_recv[r]], cap_recv[r]);
    }
    if (r == a.myrank) return;
    const long long n = (a.stage == 1) ? min((long long)a.in[(size_t)HDR * r + (size_t)a.rd * off_recv[r]], cap_recv[r]) : (long long)a.rcnt[r];
    if (k >= n) return;
    const double *o = a.in + (size_t)HDR * r + (size_t)a.rd * off_recv[r] + HDR + (size_t)a.rd * k;
    const int64_t i = a.nlocal + goff + k;
    if (a.stage == 1) {
        reinterpret_cast<double4 *>(a.xyzh)[i] = make_double4(o[0], o[1], o[2], o[3]);
        double *v = a.vxyzu + (size_t)a.nvu * i, *f = a.fxyzu + (size_t)a.nvu * i, *fe = a.fext + 3 * (size_t)i;
        v[0] = o[4]; v[1] = o[5]; v[2] = o[6]; if (a.nvu >= 4) { v[3] = o[7]; f[3] = 0.; }
        f[0] = o[8]; f[1] = o[9]; f[2] = o[10];
        fe[0] = fe[1] = fe[2] = 0.;
        if (a.mhd) reinterpret_cast<double4 *>(a.Bevol)[i] = make_double4(o[11], o[12], o[13], o[14]);
        a.iphase[i] = (int8_t)(-abs((int)o[15]));                    // inactive: neighbour only
    } else if (a.stage == 2) {
        a.xyzh[4 * (size_t)i + 3] = o[0];
        a.gradh[(size_t)a.ngradh * i] = (float)o[1];
        a.alphaind[3 * (size_t)i] = (float)o[2];
        if (a.ngradh > 1) a.gradh[(size_t)a.ngradh * i + 1] = (float)o[3];
    } else {
        double *v = a.vxyzu + (size_t)a.nvu * i;
        v[0] = o[0]; v[1] = o[1]; v[2] = o[2]; if (a.nvu >= 4) v[3] = o[3];
        if (a.mhd) reinterpret_cast<double4 *>(a.Bevol)[i] = make_double4(o[4], o[5], o[6], o[7]);
    }
}

__global__ void k_refresh_h2(int64_t nlive, const int *__restrict__ perm, const double *__restrict__ xyzh, double4 *__restrict__ pos4)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    pos4[s].w = xyzh[4 * (size_t)perm[s] + 3];
}

__global__ void k_restore_h2(int64_t n, double *__restrict__ xyzh, const double *__restrict__ h_build)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) xyzh[4 * i + 3] = h_build[i];
}

// largest h and largest distance outside the own box over the owned particles -> red[0], red[1] (ordered bits of non-negative doubles)
__global__ void k_dist_hmax_overhang(int64_t n, const double *__restrict__ xyzh, const double *__restrict__ box, double Lx, double Ly, double Lz, int periodic,
                                     double *out)
{
    double hm = 0., ov = 0.;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double4 x = reinterpret_cast<const double4 *>(xyzh)[i];
        if (x.w < DBL_MIN) continue;
        hm = fmax(hm, x.w);
        const double gx = axis_gap(x.x, box[0], box[3], Lx, periodic), gy = axis_gap(x.y, box[1], box[4], Ly, periodic), gz = axis_gap(x.z, box[2], box[5], Lz, periodic);
        ov = fmax(ov, sqrt(gx * gx + gy * gy + gz * gz));
    }
    hm = warp_max(hm); ov = warp_max(ov);
    if (lane_id() == 0) { atomic_max_pos(out, hm); atomic_max_pos(out + 1, ov); }
}

// any block whose count exceeds its capacity -> flag (as a double, for the max all-reduce)
__global__ void k_dist_overflow(int nranks, int myrank, const unsigned long long *cnt, const long long *caps, const double *recv, int rd, double *flag)
{
    const int r = threadIdx.x;
    if (r >= nranks || r == myrank) return;
    const long long *cap_recv = caps + 2 * nranks, *off_recv = caps + 3 * nranks;
    bool over = (long long)cnt[r] > caps[r];
    if (recv) over = over || (long long)recv[(size_t)HDR * r + (size_t)rd * off_recv[r]] > cap_recv[r];
    if (over) *flag = 1.;
}

// ---- migration -------------------------------------------------------------------------------------------------------------------
struct MigArrays {
    double *xyzh, *vxyzu, *fxyzu, *fext, *Bevol, *dBevol, *eos_vars, *dustfrac, *tstop, *v_true, *B_true;
    float *divcurlv, *divcurlB, *alphaind, *gradh, *dvdx, *poten, *divBsymm;
    int8_t *iphase, *ibin, *ibin_old, *ibin_wake;
    long long *gid;
    int nvu, ngradh;
};
__device__ __forceinline__ void mig_pack(const MigArrays &a, int64_t i, double *o)
{
    for (int k = 0; k < 4; k++) o[k] = a.xyzh[4 * i + k];
    for (int k = 0; k < 4; k++) { o[4 + k] = k < a.nvu ? a.vxyzu[a.nvu * i + k] : 0.; o[8 + k] = k < a.nvu ? a.fxyzu[a.nvu * i + k] : 0.; }
    for (int k = 0; k < 3; k++) o[12 + k] = a.fext[3 * i + k];
    for (int k = 0; k < 4; k++) { o[15 + k] = a.Bevol[4 * i + k]; o[19 + k] = a.dBevol[4 * i + k]; }
    for (int k = 0; k < 7; k++) o[23 + k] = a.eos_vars[7 * i + k];
    o[30] = a.divcurlv[i];
    for (int k = 0; k < 4; k++) o[31 + k] = a.divcurlB[4 * i + k];
    for (int k = 0; k < 3; k++) o[35 + k] = a.alphaind[3 * i + k];
    for (int k = 0; k < 2; k++) o[38 + k] = k < a.ngradh ? a.gradh[a.ngradh * i + k] : 0.;
    for (int k = 0; k < 9; k++) o[40 + k] = a.dvdx[9 * i + k];
    o[49] = a.poten[i]; o[50] = a.divBsymm[i]; o[51] = a.iphase[i]; o[52] = a.ibin[i]; o[53] = a.ibin_old[i]; o[54] = a.ibin_wake[i];
    o[55] = a.dustfrac[i]; o[56] = a.tstop[i]; o[57] = (double)a.gid[i];
    for (int k = 0; k < 4; k++) { o[58 + k] = (a.v_true && k < a.nvu) ? a.v_true[a.nvu * i + k] : 0.; o[62 + k] = a.B_true ? a.B_true[4 * i + k] : 0.; }
}
__device__ __forceinline__ void mig_unpack(const MigArrays &a, int64_t i, const double *o)
{
    for (int k = 0; k < 4; k++) a.xyzh[4 * i + k] = o[k];
    for (int k = 0; k < a.nvu; k++) { a.vxyzu[a.nvu * i + k] = o[4 + k]; a.fxyzu[a.nvu * i + k] = o[8 + k]; }
    for (int k = 0; k < 3; k++) a.fext[3 * i + k] = o[12 + k];
    for (int k = 0; k < 4; k++) { a.Bevol[4 * i + k] = o[15 + k]; a.dBevol[4 * i + k] = o[19 + k]; }
    for (int k = 0; k < 7; k++) a.eos_vars[7 * i + k] = o[23 + k];
    a.divcurlv[i] = (float)o[30];
    for (int k = 0; k < 4; k++) a.divcurlB[4 * i + k] = (float)o[31 + k];
    for (int k = 0; k < 3; k++) a.alphaind[3 * i + k] = (float)o[35 + k];
    for (int k = 0; k < a.ngradh; k++) a.gradh[a.ngradh * i + k] = (float)o[38 + k];
    for (int k = 0; k < 9; k++) a.dvdx[9 * i + k] = (float)o[40 + k];
    a.poten[i] = (float)o[49]; a.divBsymm[i] = (float)o[50]; a.iphase[i] = (int8_t)o[51]; a.ibin[i] = (int8_t)o[52]; a.ibin_old[i] = (int8_t)o[53];
    a.ibin_wake[i] = (int8_t)o[54]; a.dustfrac[i] = o[55]; a.tstop[i] = o[56]; a.gid[i] = (long long)o[57];
    if (a.v_true) for (int k = 0; k < a.nvu; k++) a.v_true[a.nvu * i + k] = o[58 + k];
    if (a.B_true) for (int k = 0; k < 4; k++) a.B_true[4 * i + k] = o[62 + k];
}

__device__ __forceinline__ double wrap_into(double x, double lo, double hi)
{
    const double L = hi - lo;
    if (x < lo) x += L; else if (x >= hi) x -= L;
    return x;
}

// owner of every owned particle (boxes tile the domain; positions are wrapped / clamped into it first); leavers are listed per new owner
__global__ void k_mig_owner(int64_t nlocal, const double *__restrict__ xyzh, int nranks, int myrank, const double *__restrict__ boxes, double glox, double gloy,
                            double gloz, double ghix, double ghiy, double ghiz, int periodic, unsigned long long *cnt, long long cap, int *idx, int *flag)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    flag[i] = 0;
    const double4 p = reinterpret_cast<const double4 *>(xyzh)[i];
    if (p.w < DBL_MIN) return;
    double x = p.x, y = p.y, z = p.z;
    if (periodic) { x = wrap_into(x, glox, ghix); y = wrap_into(y, gloy, ghiy); z = wrap_into(z, gloz, ghiz); }
    else { x = fmin(fmax(x, glox), ghix); y = fmin(fmax(y, gloy), ghiy); z = fmin(fmax(z, gloz), ghiz); }
    const double *b = boxes + 6 * myrank;
    if (x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5]) return;     // still at home (faces count as home)
    for (int r = 0; r < nranks; r++) {
        if (r == myrank) continue;
        b = boxes + 6 * r;
        if (x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5]) {
            const long long k = (long long)atomicAdd(&cnt[r], 1ull);
            if (k < cap) { idx[(long long)r * cap + k] = (int)i; flag[i] = 1; }       // a full block: the particle stays one more step
            return;
        }
    }
}
__global__ void k_mig_pack(const MigArrays a, int nranks, long long cap, const unsigned long long *cnt, const int *idx, double *out)
{
    const int r = blockIdx.y;
    const long long n = min((long long)cnt[r], cap);
    double *blk = out + (size_t)r * (HDR + (size_t)NREC * cap);
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k == 0) { blk[0] = (double)n; blk[1] = (double)cap; }
    if (k >= n) return;
    mig_pack(a, idx[(long long)r * cap + k], blk + HDR + (size_t)NREC * k);
}
// holes = leavers below the new end n1; sources = stayers at or above it
__global__ void k_mig_lists(int64_t nlocal, int64_t n1, const int *flag, unsigned long long *cnt2, int *hole, int *src)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    if (i < n1 && flag[i]) hole[atomicAdd(&cnt2[0], 1ull)] = (int)i;
    if (i >= n1 && !flag[i]) src[atomicAdd(&cnt2[1], 1ull)] = (int)i;
}
__global__ void k_mig_move_out(const MigArrays a, long long n, const int *src, double *tmp)
{
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k < n) mig_pack(a, src[k], tmp + (size_t)NREC * k);
}
__global__ void k_mig_move_in(const MigArrays a, long long n, const int *hole, const double *tmp)
{
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k < n) mig_unpack(a, hole[k], tmp + (size_t)NREC * k);
}
__global__ void k_mig_arrive(const MigArrays a, int nranks, int myrank, long long cap, int64_t n1, const double *in)
{
    const int r = blockIdx.y;
    if (r == myrank) return;
    long long off = 0;
    for (int q = 0; q < r; q++) if (q != myrank) off += (long long)in[(size_t)q * (HDR + (size_t)NREC * cap)];
    const double *blk = in + (size_t)r * (HDR + (size_t)NREC * cap);
    const long long n = (long long)blk[0];
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    mig_unpack(a, n1 + off + k, blk + HDR + (size_t)NREC * k);
}
__global__ void k_gid_init(int64_t n, long long base, long long *gid)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) gid[i] = base + i;
}

// moments of the owned particles inside each box of the current bisection level: {sum m, sum m x_axis} per box -> red[2 b], red[2 b + 1]
__global__ void k_orb_moments(int64_t n, const double *__restrict__ xyzh, const int8_t *__restrict__ iphase, int nbox, const double *__restrict__ boxes,
                              const int *__restrict__ axis, double glox, double gloy, double gloz, double ghix, double ghiy, double ghiz, int periodic,
                              const double *massoftype, double *red)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 p = reinterpret_cast<const double4 *>(xyzh)[i];
    if (p.w < DBL_MIN) return;
    double x[3] = {p.x, p.y, p.z};
    if (periodic) { x[0] = wrap_into(x[0], glox, ghix); x[1] = wrap_into(x[1], gloy, ghiy); x[2] = wrap_into(x[2], gloz, ghiz); }
    else { x[0] = fmin(fmax(x[0], glox), ghix); x[1] = fmin(fmax(x[1], gloy), ghiy); x[2] = fmin(fmax(x[2], gloz), ghiz); }
    const double m = massoftype[abs((int)iphase[i])];
    for (int b = 0; b < nbox; b++) {
        const double *bb = boxes + 6 * b;
        if (x[0] >= bb[0] && x[0] <= bb[3] && x[1] >= bb[1] && x[1] <= bb[4] && x[2] >= bb[2] && x[2] <= bb[5]) {
            atomicAdd(&red[2 * b], m); atomicAdd(&red[2 * b + 1], m * x[axis[b]]);
            return;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------------------
int upload_caps(sphgpu_ctx *c, DistState *d)
{
    const int P = d->nranks;
    d->off_send.assign(P, 0); d->off_recv.assign(P, 0);
    long long os = 0, orr = 0;
    for (int r = 0; r < P; r++) { d->off_send[r] = os; os += d->cap_send[r]; d->off_recv[r] = orr; orr += d->cap_recv[r]; }
    d->ncap_ghost = orr;
    std::vector<long long> h(4 * P);
    for (int r = 0; r < P; r++) { h[r] = d->cap_send[r]; h[P + r] = d->off_send[r]; h[2 * P + r] = d->cap_recv[r]; h[3 * P + r] = d->off_recv[r]; }
    CUDA_TRY(c, d->d_caps.ensure(4 * P));
    CUDA_TRY(c, cudaMemcpyAsync(d->d_caps.p, h.data(), sizeof(long long) * 4 * P, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));                 // h is a stack vector
    CUDA_TRY(c, d->sendidx.ensure((size_t)os + 1));
    return SPHGPU_OK;
}

// grouped point-to-point exchange of the per-peer blocks (rd doubles per record)
int exchange_blocks(sphgpu_ctx *c, DistState *d, int rd)
{
    const int P = d->nranks;
    NCCL_TRY(c, d, d->api->GroupStart());
    for (int r = 0; r < P; r++) {
        if (r == d->rank) continue;
        const size_t ns = HDR + (size_t)rd * d->cap_send[r], nr = HDR + (size_t)rd * d->cap_recv[r];
        NCCL_TRY(c, d, d->api->Send(d->sendbuf.p + (size_t)HDR * r + (size_t)rd * d->off_send[r], ns, ncclDouble, r, d->comm, c->stream));
        NCCL_TRY(c, d, d->api->Recv(d->recvbuf.p + (size_t)HDR * r + (size_t)rd * d->off_recv[r], nr, ncclDouble, r, d->comm, c->stream));
        d->halo_bytes += (int64_t)sizeof(double) * (int64_t)(ns + nr);
    }
    NCCL_TRY(c, d, d->api->GroupEnd());
    return SPHGPU_OK;
}

// select + pack + exchange + unpack of one stage.  stage 1 also (re)defines which particles are ghosts where.
int halo_stage(sphgpu_ctx *c, DistState *d, int stage, double dhalo)
{
    const int P = d->nranks;
    const sphgpu_params &p = c->hp.p;
    const int rd = stage == 1 ? 16 : (stage == 2 ? 4 : 8);
    long long maxs = 1, maxr = 1, sums = 0;
    for (int r = 0; r < P; r++) { maxs = std::max(maxs, d->cap_send[r]); maxr = std::max(maxr, d->cap_recv[r]); sums += d->cap_send[r]; }
    CUDA_TRY(c, d->sendbuf.ensure((size_t)HDR * P + (size_t)16 * sums + 16)); CUDA_TRY(c, d->recvbuf.ensure((size_t)HDR * P + (size_t)16 * d->ncap_ghost + 16));
    if (stage == 1) {
        c->npart = c->nlocal; c->nghost = 0; c->tree_valid = false;
        CUDA_TRY(c, cudaMemsetAsync(d->cnt.p, 0, sizeof(unsigned long long) * P, c->stream));
        k_dist_select<<<nblk(c->nlocal, 256), 256, 0, c->stream>>>(c->nlocal, c->xyzh.p, P, d->rank, d->d_boxes.p, dhalo, c->hp.dxbound, c->hp.dybound, c->hp.dzbound,
                                                                   p.periodic, d->cnt.p, d->d_caps.p, d->sendidx.p);
        c->launches++;
    }
    PackArgs pa; memset(&pa, 0, sizeof pa);
    pa.xyzh = c->xyzh.p; pa.vxyzu = c->vxyzu.p; pa.fxyzu = c->fxyzu.p; pa.fext = c->fext.p; pa.Bevol = c->Bevol.p; pa.gradh = c->gradh.p; pa.alphaind = c->alphaind.p;
    pa.iphase = c->iphase.p; pa.nvu = c->hp.nvu; pa.mhd = p.mhd; pa.ngradh = c->hp.ngradh; pa.nranks = P; pa.stage = stage; pa.rd = rd;
    pa.cnt = d->cnt.p; pa.caps = d->d_caps.p; pa.sendidx = d->sendidx.p; pa.out = d->sendbuf.p;
    k_dist_pack<<<dim3(nblk(maxs, 256), P), 256, 0, c->stream>>>(pa);
    c->launches++;
    TRY(exchange_blocks(c, d, rd));
    if (stage == 1) {
        const int64_t ntot = c->nlocal + d->ncap_ghost;
        TRY(ensure_all_keep(c, ntot, c->nlocal));
        c->nghost = d->ncap_ghost; c->npart = ntot;
    }
    UnpackArgs ua; memset(&ua, 0, sizeof ua);
    ua.xyzh = c->xyzh.p; ua.vxyzu = c->vxyzu.p; ua.fxyzu = c->fxyzu.p; ua.fext = c->fext.p; ua.Bevol = c->Bevol.p; ua.gradh = c->gradh.p; ua.alphaind = c->alphaind.p;
    ua.iphase = c->iphase.p; ua.nvu = c->hp.nvu; ua.mhd = p.mhd; ua.ngradh = c->hp.ngradh; ua.nranks = P; ua.myrank = d->rank; ua.stage = stage; ua.rd = rd;
    ua.nlocal = c->nlocal; ua.ncap = d->ncap_ghost; ua.caps = d->d_caps.p; ua.rcnt = d->cnt.p + P; ua.in = d->recvbuf.p;
    k_dist_unpack<<<dim3(nblk(std::max(maxr, d->ncap_ghost), 256), P), 256, 0, c->stream>>>(ua);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return SPHGPU_OK;
}

// the true counts of the last selection go to the peers, both sides size the blocks from them (first call, and after an overflow)
int renegotiate_caps(sphgpu_ctx *c, DistState *d)
{
    const int P = d->nranks;
    std::vector<unsigned long long> hs(P), hr(P);
    CUDA_TRY(c, d->red.ensure(4 * P + 64));
    unsigned long long *dsend = reinterpret_cast<unsigned long long *>(d->red.p), *drecv = dsend + P;
    CUDA_TRY(c, cudaMemcpyAsync(dsend, d->cnt.p, sizeof(unsigned long long) * P, cudaMemcpyDeviceToDevice, c->stream));
    NCCL_TRY(c, d, d->api->GroupStart());
    for (int r = 0; r < P; r++) {
        if (r == d->rank) continue;
        NCCL_TRY(c, d, d->api->Send(dsend + r, 1, ncclUint64, r, d->comm, c->stream));
        NCCL_TRY(c, d, d->api->Recv(drecv + r, 1, ncclUint64, r, d->comm, c->stream));
    }
    NCCL_TRY(c, d, d->api->GroupEnd());
    CUDA_TRY(c, cudaMemcpyAsync(hs.data(), dsend, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(hr.data(), drecv, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < P; r++) {
        if (r == d->rank) { d->cap_send[r] = d->cap_recv[r] = 0; continue; }
        d->cap_send[r] = std::max(d->cap_send[r], grow_cap((long long)hs[r]));
        d->cap_recv[r] = std::max(d->cap_recv[r], grow_cap((long long)hr[r]));
    }
    TRY(upload_caps(c, d));
    d->caps_valid = true;
    return SPHGPU_OK;
}

// count-only selection (nothing is listed): the input of the very first negotiation
int count_selection(sphgpu_ctx *c, DistState *d, double dhalo)
{
    const int P = d->nranks;
    std::vector<long long> zero(4 * P, 0);
    CUDA_TRY(c, d->d_caps.ensure(4 * P)); CUDA_TRY(c, d->sendidx.ensure(1));
    CUDA_TRY(c, cudaMemcpyAsync(d->d_caps.p, zero.data(), sizeof(long long) * 4 * P, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(d->cnt.p, 0, sizeof(unsigned long long) * P, c->stream));
    k_dist_select<<<nblk(c->nlocal, 256), 256, 0, c->stream>>>(c->nlocal, c->xyzh.p, P, d->rank, d->d_boxes.p, dhalo, c->hp.dxbound, c->hp.dybound, c->hp.dzbound,
                                                               c->hp.p.periodic, d->cnt.p, d->d_caps.p, d->sendidx.p);
    c->launches++;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPHGPU_OK;
}

int allreduce(sphgpu_ctx *c, DistState *d, const double *h_in, double *h_out, int n, ncclRedOp_t op)
{
    CUDA_TRY(c, d->red.ensure(256)); CUDA_TRY(c, d->red_out.ensure(256));
    CUDA_TRY(c, cudaMemcpyAsync(d->red.p, h_in, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(c, d, d->api->AllReduce(d->red.p, d->red_out.p, n, ncclDouble, op, d->comm, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(h_out, d->red_out.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPHGPU_OK;
}

// self-gravity: every rank builds the tree of the whole set from an all-gather of {x, h, type, h history} (see include/sphgpu.h)
int gather_gravity_set(sphgpu_ctx *c, DistState *d)
{
    const int P = d->nranks;
    double mine[64] = {0}, all[64] = {0};
    if (P > 64) { c->err = "dist: more than 64 ranks"; return SPHGPU_ERR_ARG; }
    mine[d->rank] = (double)c->nlocal;
    TRY(allreduce(c, d, mine, all, P, ncclSum));
    std::vector<int64_t> counts(P);
    int64_t stride = 0;
    for (int r = 0; r < P; r++) { counts[r] = (int64_t)all[r]; stride = std::max(stride, counts[r]); }
    void *sendptr = nullptr, *recvptr = nullptr; int rd = 0;
    TRY(gravity_gather_pack(c, &sendptr, &rd));
    TRY(gravity_gather_recvbuf(c, P, stride, &recvptr));
    // ranks with fewer particles send a padded block: the pack buffer is at least nlocal records, the gather needs `stride`
    CUDA_TRY(c, d->msend.ensure((size_t)rd * stride + 1));
    CUDA_TRY(c, cudaMemsetAsync(d->msend.p, 0, sizeof(double) * (size_t)rd * stride, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(d->msend.p, sendptr, sizeof(double) * (size_t)rd * c->nlocal, cudaMemcpyDeviceToDevice, c->stream));
    NCCL_TRY(c, d, d->api->AllGather(d->msend.p, recvptr, (size_t)rd * stride, ncclDouble, d->comm, c->stream));
    d->halo_bytes += (int64_t)sizeof(double) * rd * stride * P;
    TRY(gravity_gather_unpack(c, P, d->rank, stride, counts.data()));
    return SPHGPU_OK;
}

MigArrays mig_arrays(sphgpu_ctx *c, bool in_step)
{
    MigArrays a; memset(&a, 0, sizeof a);
    a.xyzh = c->xyzh.p; a.vxyzu = c->vxyzu.p; a.fxyzu = c->fxyzu.p; a.fext = c->fext.p; a.Bevol = c->Bevol.p; a.dBevol = c->dBevol.p; a.eos_vars = c->eos_vars.p;
    a.dustfrac = c->dustfrac.p; a.tstop = c->tstop.p; a.divcurlv = c->divcurlv.p; a.divcurlB = c->divcurlB.p; a.alphaind = c->alphaind.p; a.gradh = c->gradh.p;
    a.dvdx = c->dvdx.p; a.poten = c->poten.p; a.divBsymm = c->divBsymm.p; a.iphase = c->iphase.p; a.ibin = c->ibin.p; a.ibin_old = c->ibin_old.p;
    a.ibin_wake = c->ibin_wake.p; a.gid = c->gid.p; a.nvu = c->hp.nvu; a.ngradh = c->hp.ngradh;
    a.v_true = in_step ? c->v_true.p : nullptr; a.B_true = (in_step && c->hp.p.mhd) ? c->B_true.p : nullptr;
    return a;
}

}  // namespace

extern "C" {

int sphgpu_dist_get_unique_id(void *id, int nbytes)
{
    std::string err;
    NcclApi *api = nccl_api(err);
    if (!api || !id || nbytes < (int)sizeof(ncclUniqueId)) return SPHGPU_ERR_ARG;
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != ncclSuccess) return SPHGPU_ERR_CUDA;
    memcpy(id, &u, sizeof u);
    return SPHGPU_OK;
}

int sphgpu_dist_init(sphgpu_ctx *c, const void *id, int nranks, int rank)
{
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    NcclApi *api = nccl_api(c->err);
    if (!api) return SPHGPU_ERR_CUDA;
    if (c->dist) sphgpu_dist_finalize(c);
    DistState *d = new DistState();
    d->api = api; d->nranks = nranks; d->rank = rank;
    ncclUniqueId u; memcpy(&u, id, sizeof u);
    ncclResult_t r = api->CommInitRank(&d->comm, nranks, u, rank);
    if (r != ncclSuccess) { c->err = std::string("ncclCommInitRank: ") + api->GetErrorString(r); delete d; return SPHGPU_ERR_CUDA; }
    d->cap_send.assign(nranks, 0); d->cap_recv.assign(nranks, 0);
    cudaEventCreate(&d->ev[0]); cudaEventCreate(&d->ev[1]);
    if (d->cnt.ensure(2 * nranks + 2) != cudaSuccess || d->mcnt.ensure(nranks + 2) != cudaSuccess || d->red.ensure(256) != cudaSuccess ||
        d->red_out.ensure(256) != cudaSuccess) { c->err = "dist: out of memory"; delete d; return SPHGPU_ERR_CUDA; }
    c->dist = d;
    c->halo_nranks = nranks; c->halo_rank = rank;
    return SPHGPU_OK;
}

int sphgpu_dist_finalize(sphgpu_ctx *c)
{
    if (!c || !c->dist) return SPHGPU_OK;
    DistState *d = c->dist;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (d->comm) d->api->CommDestroy(d->comm);
    d->d_boxes.release(); d->d_caps.release(); d->cnt.release(); d->sendidx.release(); d->sendbuf.release(); d->recvbuf.release(); d->red.release(); d->red_out.release();
    d->mcnt.release(); d->midx.release(); d->mflag.release(); d->mhole.release(); d->msrc.release(); d->msend.release(); d->mrecv.release(); d->mtmp.release();
    cudaEventDestroy(d->ev[0]); cudaEventDestroy(d->ev[1]);
    delete d;
    c->dist = nullptr;
    return SPHGPU_OK;
}

// boxes: 6 doubles per rank {lo xyz, hi xyz} tiling the domain [glo, ghi] (the periodic box, or the bounding box of the set)
int sphgpu_dist_set_boxes(sphgpu_ctx *c, const double *boxes)
{
    sphgpu_dist_mark_dirty(c);
    if (!c || !c->dist || !boxes) return SPHGPU_ERR_ARG;
    DistState *d = c->dist;
    CUDA_TRY(c, cudaSetDevice(c->device));
    d->boxes.assign(boxes, boxes + 6 * d->nranks);
    for (int k = 0; k < 3; k++) { d->glo[k] = DBL_MAX; d->ghi[k] = -DBL_MAX; }
    for (int r = 0; r < d->nranks; r++) for (int k = 0; k < 3; k++) { d->glo[k] = std::min(d->glo[k], boxes[6 * r + k]); d->ghi[k] = std::max(d->ghi[k], boxes[6 * r + 3 + k]); }
    CUDA_TRY(c, d->d_boxes.ensure(6 * d->nranks));
    CUDA_TRY(c, cudaMemcpyAsync(d->d_boxes.p, d->boxes.data(), sizeof(double) * 6 * d->nranks, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    d->hu_prev = -1.;
    return SPHGPU_OK;
}

int sphgpu_dist_get_boxes(sphgpu_ctx *c, double *boxes)
{
    if (!c || !c->dist || !boxes || c->dist->boxes.empty()) return SPHGPU_ERR_ARG;
    memcpy(boxes, c->dist->boxes.data(), sizeof(double) * 6 * c->dist->nranks);
    return SPHGPU_OK;
}

// global particle identities of the owned particles (they travel with the particles when these migrate); ids == NULL: base + position
int sphgpu_dist_set_ids(sphgpu_ctx *c, const int64_t *ids, int64_t base)
{
    sphgpu_dist_mark_dirty(c);
    if (!c || c->nlocal <= 0) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, c->gid.ensure(c->nlocal));
    if (ids) CUDA_TRY(c, cudaMemcpyAsync(c->gid.p, ids, sizeof(long long) * c->nlocal, cudaMemcpyHostToDevice, c->stream));
    else { k_gid_init<<<nblk(c->nlocal, 256), 256, 0, c->stream>>>(c->nlocal, (long long)base, c->gid.p); c->launches++; }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SPHGPU_OK;
}
int sphgpu_dist_get_ids(sphgpu_ctx *c, int64_t *ids, int64_t maxn)
{
    if (!c || !ids || maxn < c->nlocal || c->gid.cap < (size_t)c->nlocal) return SPHGPU_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaMemcpy(ids, c->gid.p, sizeof(long long) * c->nlocal, cudaMemcpyDeviceToHost));
    return SPHGPU_OK;
}
int64_t sphgpu_dist_nlocal(sphgpu_ctx *c) { return c ? c->nlocal : 0; }

/* halo statistics of the last dist_derivs: [0] ghost capacity, [1] ghosts received, [2] bytes moved (send + receive), [3] exchange rounds,
 * [4] particles migrated by the last sphgpu_dist_migrate, [5] largest trial h, [6] device time of the whole call (ms, CUDA events on the
 * context's stream, which also carries the NCCL transfers), [7] owned particles */
int sphgpu_dist_stats(sphgpu_ctx *c, double *out6)
{
    if (!c || !c->dist || !out6) return SPHGPU_ERR_ARG;
    DistState *d = c->dist;
    unsigned long long got = 0;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaMemcpy(&got, d->cnt.p + 2 * d->nranks, sizeof got, cudaMemcpyDeviceToHost));
    out6[0] = (double)d->ncap_ghost; out6[1] = (double)got; out6[2] = (double)d->halo_bytes; out6[3] = (double)d->rounds; out6[4] = (double)d->nmigrated_last;
    out6[5] = d->hu_prev; out6[6] = d->ms_exchange; out6[7] = (double)c->nlocal;
    return SPHGPU_OK;
}

// derivs on the decomposed set: tree + density + cons2prim + force with the ghost exchanges; the scalars come back reduced over the ranks
int sphgpu_dist_derivs(sphgpu_ctx *c, int icall, double dt, sphgpu_scalars *out)
{
    if (!c || !c->dist || icall < 1 || icall > 2) return SPHGPU_ERR_ARG;
    DistState *d = c->dist;
    if (d->boxes.empty()) { c->err = "dist_derivs: no domain boxes (sphgpu_dist_set_boxes / sphgpu_dist_rebalance)"; return SPHGPU_ERR_STATE; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int P = d->nranks;
    const sphgpu_params &p = c->hp.p;
    const double radkern = c->hp.kc.radkern;
    d->halo_bytes = 0; d->rounds = 0;
    cudaEventRecord(d->ev[0], c->stream);
    sphgpu_scalars sd; memset(&sd, 0, sizeof sd);
    if (icall == 1) {
        // halo width: radkern x (largest trial h of the previous density pass, all ranks) x margin + how far owned particles stick out of
        // their boxes; both come out of ONE all-reduce (the first call measures h directly)
        if (d->geom_dirty || d->hu_prev < 0.) {   // (a repeated derivs on an unchanged state has nothing new to reduce)
            const double want = (double)d->mig_want;
            CUDA_TRY(c, cudaMemsetAsync(d->red.p, 0, 2 * sizeof(double), c->stream));
            CUDA_TRY(c, cudaMemcpyAsync(d->red.p + 2, &want, sizeof want, cudaMemcpyHostToDevice, c->stream));
            k_dist_hmax_overhang<<<c->numSMs * 4, 256, 0, c->stream>>>(c->nlocal, c->xyzh.p, d->d_boxes.p + 6 * d->rank, c->hp.dxbound, c->hp.dybound, c->hp.dzbound,
                                                                       p.periodic, d->red.p);
            c->launches++;
            double glob[3];
            NCCL_TRY(c, d, d->api->AllReduce(d->red.p, d->red_out.p, 3, ncclDouble, ncclMax, d->comm, c->stream));
            CUDA_TRY(c, cudaMemcpyAsync(glob, d->red_out.p, sizeof glob, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            if (d->hu_prev < 0.) d->hu_prev = glob[0];
            d->hu_prev = std::max(d->hu_prev, glob[0]);
            d->overhang = glob[1];
            if ((long long)glob[2] > d->mig_cap / 2) d->mig_cap = 4 * (long long)glob[2];      // the same decision on every rank
            d->mig_want = 0;
            d->geom_dirty = false;
        }
        double dhalo = radkern * d->hu_prev * d->margin + 2. * d->overhang;
        while (true) {
            d->rounds++;
            if (!d->caps_valid) { TRY(count_selection(c, d, dhalo)); TRY(renegotiate_caps(c, d)); }
            TRY(halo_stage(c, d, 1, dhalo));
            TRY(tree_build(c));
            TRY(density_run(c, 1, &sd));
            // one all-reduce: widest trial h, largest growth of any h, any overflowed block
            double loc[3] = {c->dens_hmax_used, c->dens_hgrow, 0.}, glob[3];
            CUDA_TRY(c, cudaMemcpyAsync(d->red.p, loc, sizeof loc, cudaMemcpyHostToDevice, c->stream));
            k_dist_overflow<<<1, 64, 0, c->stream>>>(P, d->rank, d->cnt.p, d->d_caps.p, d->recvbuf.p, 16, d->red.p + 2);
            c->launches++;
            NCCL_TRY(c, d, d->api->AllReduce(d->red.p, d->red_out.p, 3, ncclDouble, ncclMax, d->comm, c->stream));
            CUDA_TRY(c, cudaMemcpyAsync(glob, d->red_out.p, sizeof glob, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            d->hu_prev = glob[0];
            c->halo_hgrow = glob[1];
            const bool too_narrow = radkern * glob[0] + 2. * d->overhang > dhalo;
            if (glob[2] == 0. && !too_narrow) break;
            if (d->rounds >= 8) { c->err = "dist_derivs: the ghost halo did not converge in 8 rounds (h grows faster than the halo)"; return SPHGPU_ERR_STATE; }
            // the reference re-exports a cell whenever its h outgrows the search radius (dens.F90:343-365): restore h, widen, repeat
            if (too_narrow) dhalo = radkern * glob[0] * d->margin + 2. * d->overhang;
            k_restore_h2<<<nblk(c->nlocal, 256), 256, 0, c->stream>>>(c->nlocal, c->xyzh.p, c->h_build.p);
            c->launches++;
            c->tree_valid = false;
            TRY(count_selection(c, d, dhalo)); TRY(renegotiate_caps(c, d));
        }
        // stage 2: the ghosts' new h, gradh, alpha; the tree's hmax follow (inflate by the global growth when small, else refit)
        TRY(halo_stage(c, d, 2, dhalo));
        k_refresh_h2<<<nblk(c->nlive, 256), 256, 0, c->stream>>>(c->nlive, c->perm.p, c->xyzh.p, c->pos4.p);
        c->launches++;
        if (c->halo_hgrow > 0. && c->halo_hgrow <= 1.02 && !c->always_refit) c->hscale = fmax(c->hscale, fmax(c->halo_hgrow, 1.) * (1. + 1e-12));
        else TRY(tree_refit_hmax(c));
        c->halo_hgrow = 0.;
        c->hp.p.set_boundaries_to_active = 0;                         // deriv.f90:146
    } else {
        if (!c->tree_valid || !d->caps_valid) { c->err = "dist_derivs(2): no derivs(1) before it"; return SPHGPU_ERR_STATE; }
        sd = c->last_dens;
        TRY(halo_stage(c, d, 3, 0.));                                 // the ghosts' predicted v, u, B of this corrector iteration
    }
    TRY(cons2prim_run(c));
    if (p.gravity && icall == 1) TRY(gather_gravity_set(c, d));
    if (p.driving) TRY(sphgpu_forcing_resident(c));
    sphgpu_scalars sf; memset(&sf, 0, sizeof sf);
    TRY(force_run(c, icall, dt, &sf));
    // reduceall_mpi of the step's scalars: one max (with the minima negated) and one sum
    {
        double mx[4] = {-sf.dtcourant, -sf.dtforce, sd.rhomax, (double)sd.maxactual}, gmx[4];
        double sm[8] = {(double)sd.np, (double)sd.nrhocalc, (double)sd.nactualtot, (double)sd.npairs_density, (double)sf.npairs_force,
                        (double)sf.npairs_gravity, (double)sf.nm2l, (double)sd.ncalls_neigh}, gsm[8];
        CUDA_TRY(c, cudaMemcpyAsync(d->red.p, mx, sizeof mx, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(d->red.p + 16, sm, sizeof sm, cudaMemcpyHostToDevice, c->stream));
        NCCL_TRY(c, d, d->api->GroupStart());
        NCCL_TRY(c, d, d->api->AllReduce(d->red.p, d->red_out.p, 4, ncclDouble, ncclMax, d->comm, c->stream));
        NCCL_TRY(c, d, d->api->AllReduce(d->red.p + 16, d->red_out.p + 16, 8, ncclDouble, ncclSum, d->comm, c->stream));
        NCCL_TRY(c, d, d->api->GroupEnd());
        CUDA_TRY(c, cudaMemcpyAsync(gmx, d->red_out.p, sizeof gmx, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(gsm, d->red_out.p + 16, sizeof gsm, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        sf.dtcourant = -gmx[0]; sf.dtforce = -gmx[1]; sf.rhomax = gmx[2]; sf.maxactual = (int64_t)gmx[3];
        sf.np = (int64_t)gsm[0]; sf.nrhocalc = (int64_t)gsm[1]; sf.nactualtot = (int64_t)gsm[2]; sf.npairs_density = (int64_t)gsm[3];
        sf.npairs_force = (int64_t)gsm[4]; sf.npairs_gravity = (int64_t)gsm[5]; sf.nm2l = (int64_t)gsm[6]; sf.ncalls_neigh = (int64_t)gsm[7];
        sf.actualmean = sf.np ? (double)sf.nactualtot / (double)sf.np : -1.;
        sf.trialmean = sd.trialmean; sf.maxtrial = sd.maxtrial;
    }
    cudaEventRecord(d->ev[1], c->stream);
    cudaEventSynchronize(d->ev[1]);
    { float ms = 0.f; cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]); d->ms_exchange = ms; }
    if (out) *out = sf;
    return SPHGPU_OK;
}

// balancedomains (mpi_balance.F90:82-173): owned particles whose (wrapped) position lies in another rank's box move there with their
// whole record; the holes are filled from the tail of the local arrays, arrivals are appended.  in_step: the leapfrog's v_true / B_true
// travel too.  Returns the new number of owned particles.
int sphgpu_dist_migrate(sphgpu_ctx *c, int in_step, int64_t *nlocal_new)
{
    sphgpu_dist_mark_dirty(c);
    if (!c || !c->dist) return SPHGPU_ERR_ARG;
    DistState *d = c->dist;
    if (d->boxes.empty()) { c->err = "dist_migrate: no domain boxes"; return SPHGPU_ERR_STATE; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int P = d->nranks;
    const int64_t n0 = c->nlocal;
    if (c->gid.cap < (size_t)n0) { c->err = "dist_migrate: no particle ids (sphgpu_dist_set_ids)"; return SPHGPU_ERR_STATE; }
    c->npart = n0; c->nghost = 0; c->tree_valid = false;            // ghosts are dropped: the set changes
    const long long cap = d->mig_cap;
    const size_t blk = HDR + (size_t)NREC * cap;
    CUDA_TRY(c, d->midx.ensure((size_t)P * cap + 1)); CUDA_TRY(c, d->mflag.ensure(n0 + 1)); CUDA_TRY(c, d->msend.ensure(blk * P)); CUDA_TRY(c, d->mrecv.ensure(blk * P));
    CUDA_TRY(c, cudaMemsetAsync(d->mcnt.p, 0, sizeof(unsigned long long) * (P + 2), c->stream));
    k_mig_owner<<<nblk(n0, 256), 256, 0, c->stream>>>(n0, c->xyzh.p, P, d->rank, d->d_boxes.p, d->glo[0], d->glo[1], d->glo[2], d->ghi[0], d->ghi[1], d->ghi[2],
                                                     c->hp.p.periodic, d->mcnt.p, cap, d->midx.p, d->mflag.p);
    MigArrays a = mig_arrays(c, in_step != 0);
    k_mig_pack<<<dim3(nblk(cap, 128), P), 128, 0, c->stream>>>(a, P, cap, d->mcnt.p, d->midx.p, d->msend.p);
    c->launches += 2;
    NCCL_TRY(c, d, d->api->GroupStart());
    for (int r = 0; r < P; r++) {
        if (r == d->rank) continue;
        NCCL_TRY(c, d, d->api->Send(d->msend.p + blk * r, blk, ncclDouble, r, d->comm, c->stream));
        NCCL_TRY(c, d, d->api->Recv(d->mrecv.p + blk * r, blk, ncclDouble, r, d->comm, c->stream));
    }
    NCCL_TRY(c, d, d->api->GroupEnd());
    // counts: what left (clipped to the blocks) and what arrived
    std::vector<unsigned long long> hc(P + 2);
    std::vector<double> hdr(P, 0.);
    CUDA_TRY(c, cudaMemcpyAsync(hc.data(), d->mcnt.p, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, c->stream));
    for (int r = 0; r < P; r++) if (r != d->rank) CUDA_TRY(c, cudaMemcpyAsync(&hdr[r], d->mrecv.p + blk * r, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    long long nsent = 0, narr = 0, want_max = 0;
    for (int r = 0; r < P; r++) { if (r == d->rank) continue; nsent += std::min<long long>((long long)hc[r], cap); narr += (long long)hdr[r]; want_max = std::max(want_max, (long long)hc[r]); }
    const int64_t n1 = n0 - nsent, n2 = n1 + narr;
    if (nsent > 0) {                                                  // fill the holes below n1 with the stayers above it
        CUDA_TRY(c, d->mhole.ensure(nsent + 1)); CUDA_TRY(c, d->msrc.ensure(nsent + 1)); CUDA_TRY(c, d->mtmp.ensure((size_t)NREC * nsent + 1));
        k_mig_lists<<<nblk(n0, 256), 256, 0, c->stream>>>(n0, n1, d->mflag.p, d->mcnt.p + P, d->mhole.p, d->msrc.p);
        unsigned long long nh[2];
        CUDA_TRY(c, cudaMemcpyAsync(nh, d->mcnt.p + P, sizeof nh, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (nh[0] != nh[1]) { c->err = "dist_migrate: hole / source lists differ in length"; return SPHGPU_ERR_STATE; }
        if (nh[0] > 0) {
            k_mig_move_out<<<nblk((long long)nh[0], 128), 128, 0, c->stream>>>(a, (long long)nh[0], d->msrc.p, d->mtmp.p);
            k_mig_move_in<<<nblk((long long)nh[0], 128), 128, 0, c->stream>>>(a, (long long)nh[0], d->mhole.p, d->mtmp.p);
            c->launches += 2;
        }
        c->launches++;
    }
    if (narr > 0) {
        TRY(ensure_all_keep(c, n2, n1));
        if (c->gid.ensure_keep(n2, n1, c->stream) != cudaSuccess) { c->err = "dist_migrate: out of memory"; return SPHGPU_ERR_CUDA; }
        if (in_step) {
            if (c->v_true.ensure_keep((size_t)c->hp.nvu * n2, (size_t)c->hp.nvu * n1, c->stream) != cudaSuccess) return SPHGPU_ERR_CUDA;
            if (c->hp.p.mhd && c->B_true.ensure_keep(4 * (size_t)n2, 4 * (size_t)n1, c->stream) != cudaSuccess) return SPHGPU_ERR_CUDA;
        }
        a = mig_arrays(c, in_step != 0);                             // the buffers may have moved
        k_mig_arrive<<<dim3(nblk(cap, 128), P), 128, 0, c->stream>>>(a, P, d->rank, cap, n1, d->mrecv.p);
        c->launches++;
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaGetLastError());
    c->nlocal = n2; c->npart = n2;
    d->nmigrated_last = nsent;
    d->mig_want = std::max(d->mig_want, want_max);                 // agreed over the ranks in the next dist_derivs (a full block only delays a particle)
    if (nlocal_new) *nlocal_new = n2;
    return SPHGPU_OK;
}

// the reference's domain split (kdtree.F90:2098-2160 applied globally): log2(P) levels of bisection at the centre of mass along the
// longest axis, the moments of every level all-reduced; followed by a migration.  domain = {lo xyz, hi xyz} to be tiled.
int sphgpu_dist_rebalance(sphgpu_ctx *c, const double *domain, int64_t *nlocal_new)
{
    sphgpu_dist_mark_dirty(c);
    if (!c || !c->dist || !domain) return SPHGPU_ERR_ARG;
    DistState *d = c->dist;
    const int P = d->nranks;
    if (P & (P - 1)) { c->err = "dist_rebalance: the number of ranks must be a power of two"; return SPHGPU_ERR_ARG; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    std::vector<double> boxes(domain, domain + 6);
    DevBuf<double> dbox, dmass; DevBuf<int> daxis;
    CUDA_TRY(c, dbox.ensure(6 * P)); CUDA_TRY(c, daxis.ensure(P)); CUDA_TRY(c, dmass.ensure(SPHGPU_MAXTYPES));
    CUDA_TRY(c, cudaMemcpyAsync(dmass.p, c->hp.p.massoftype, sizeof(double) * SPHGPU_MAXTYPES, cudaMemcpyHostToDevice, c->stream));
    int rc = SPHGPU_OK;
    for (int nb = 1; nb < P && rc == SPHGPU_OK; nb *= 2) {
        std::vector<int> axis(nb);
        for (int b = 0; b < nb; b++) {
            const double *bb = &boxes[6 * b];
            const double e[3] = {bb[3] - bb[0], bb[4] - bb[1], bb[5] - bb[2]};
            axis[b] = (e[0] >= e[1] && e[0] >= e[2]) ? 0 : (e[1] >= e[2] ? 1 : 2);
        }
        std::vector<double> loc(2 * nb, 0.), glob(2 * nb, 0.);
        if (cudaMemcpyAsync(dbox.p, boxes.data(), sizeof(double) * 6 * nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(daxis.p, axis.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemsetAsync(d->red.p, 0, sizeof(double) * 2 * nb, c->stream) != cudaSuccess) { rc = SPHGPU_ERR_CUDA; break; }
        k_orb_moments<<<nblk(c->nlocal, 256), 256, 0, c->stream>>>(c->nlocal, c->xyzh.p, c->iphase.p, nb, dbox.p, daxis.p, domain[0], domain[1], domain[2], domain[3],
                                                                   domain[4], domain[5], c->hp.p.periodic, dmass.p, d->red.p);
        c->launches++;
        if (cudaMemcpyAsync(loc.data(), d->red.p, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = SPHGPU_ERR_CUDA; break; }
        cudaStreamSynchronize(c->stream);
        rc = allreduce(c, d, loc.data(), glob.data(), 2 * nb, ncclSum);
        if (rc != SPHGPU_OK) break;
        std::vector<double> next(12 * nb);
        for (int b = 0; b < nb; b++) {
            const double *bb = &boxes[6 * b];
            const int ax = axis[b];
            double pivot = glob[2 * b] > 0. ? glob[2 * b + 1] / glob[2 * b] : 0.5 * (bb[ax] + bb[3 + ax]);
            pivot = std::min(std::max(pivot, bb[ax]), bb[3 + ax]);
            double *l = &next[12 * b], *r = &next[12 * b + 6];
            memcpy(l, bb, 6 * sizeof(double)); memcpy(r, bb, 6 * sizeof(double));
            l[3 + ax] = pivot; r[ax] = pivot;
        }
        boxes.swap(next);
    }
    dbox.release(); daxis.release(); dmass.release();
    if (rc != SPHGPU_OK) return rc;
    TRY(sphgpu_dist_set_boxes(c, boxes.data()));
    c->dist->caps_valid = false;                                     // the neighbours of every box changed
    std::fill(c->dist->cap_send.begin(), c->dist->cap_send.end(), 0); std::fill(c->dist->cap_recv.begin(), c->dist->cap_recv.end(), 0);
    // a rebalance may move a large fraction of the set: count the leavers per peer first and agree on blocks that hold them all
    {
        d = c->dist;
        CUDA_TRY(c, d->mflag.ensure(c->nlocal + 1)); CUDA_TRY(c, d->midx.ensure(1));
        CUDA_TRY(c, cudaMemsetAsync(d->mcnt.p, 0, sizeof(unsigned long long) * (P + 2), c->stream));
        k_mig_owner<<<nblk(c->nlocal, 256), 256, 0, c->stream>>>(c->nlocal, c->xyzh.p, P, d->rank, d->d_boxes.p, d->glo[0], d->glo[1], d->glo[2], d->ghi[0], d->ghi[1],
                                                                 d->ghi[2], c->hp.p.periodic, d->mcnt.p, 0, d->midx.p, d->mflag.p);
        c->launches++;
        std::vector<unsigned long long> hc(P);
        CUDA_TRY(c, cudaMemcpyAsync(hc.data(), d->mcnt.p, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        double loc[1] = {0.}, glob[1];
        for (int r = 0; r < P; r++) if (r != d->rank) loc[0] = std::max(loc[0], (double)hc[r]);
        TRY(allreduce(c, d, loc, glob, 1, ncclMax));
        d->mig_cap = std::max<long long>(d->mig_cap, (long long)glob[0] + 1024);
    }
    return sphgpu_dist_migrate(c, 0, nlocal_new);
}

// the leapfrog step of sphgpu_step_resident on the decomposed set: migration at the start of the step, ghost exchanges inside every derivs,
// the velocity-error norm reduced over the ranks (check_velocity_error reduces errmax, v2mean and np with reduceall_mpi, :790-792)
int sphgpu_dist_step(sphgpu_ctx *c, double dtsph, double tolv, sphgpu_step_out *out)
{
    if (!c || !c->dist) return SPHGPU_ERR_ARG;
    TRY(sphgpu_dist_migrate(c, 0, nullptr));
    return sphgpu_step_resident(c, dtsph, tolv, out);                 // step.cu routes its derivs and reductions through the hooks below
}

// compute_energies with the sums reduced over the ranks (energies.f90:648-674)
int sphgpu_dist_energies(sphgpu_ctx *c, sphgpu_energies *out)
{
    if (!c || !c->dist || !out) return SPHGPU_ERR_ARG;
    DistState *d = c->dist;
    sphgpu_energies e;
    TRY(sphgpu_energies_resident(c, &e));
    double sm[16] = {e.ekin, e.etherm, e.emag, e.epot, e.xmom, e.ymom, e.zmom, e.angx, e.angy, e.angz, e.mtot, e.xcom * e.mtot, e.ycom * e.mtot, e.zcom * e.mtot,
                     (double)e.np, 0.}, g[16];
    double mx[1] = {e.rhomax}, gm[1];
    TRY(allreduce(c, d, sm, g, 15, ncclSum));
    TRY(allreduce(c, d, mx, gm, 1, ncclMax));
    memset(out, 0, sizeof *out);
    out->ekin = g[0]; out->etherm = g[1]; out->emag = g[2]; out->epot = g[3]; out->etot = g[0] + g[1] + g[2] + g[3];
    out->xmom = g[4]; out->ymom = g[5]; out->zmom = g[6]; out->totmom = sqrt(g[4] * g[4] + g[5] * g[5] + g[6] * g[6]);
    out->angx = g[7]; out->angy = g[8]; out->angz = g[9]; out->angtot = sqrt(g[7] * g[7] + g[8] * g[8] + g[9] * g[9]);
    out->mtot = g[10];
    const double dm = g[10] > 0. ? 1. / g[10] : 0.;
    out->xcom = g[11] * dm; out->ycom = g[12] * dm; out->zcom = g[13] * dm;
    out->np = (int64_t)g[14]; out->rhomax = gm[0];
    return SPHGPU_OK;
}

}  // extern "C"

// ---- hooks used by step.cu -----------------------------------------------------------------------------------------------------------
int sphgpu_dist_hook_derivs(sphgpu_ctx *c, int icall, double dt, sphgpu_scalars *out) { if (c && c->dist) c->dist->geom_dirty = true; return sphgpu_dist_derivs(c, icall, dt, out); }
void sphgpu_dist_mark_dirty(sphgpu_ctx *c) { if (c && c->dist) c->dist->geom_dirty = true; }
int sphgpu_dist_hook_reduce_err(sphgpu_ctx *c, double *red3)
{
    DistState *d = c->dist;
    double mx[1] = {red3[0]}, gm[1], sm[2] = {red3[1], red3[2]}, gs[2];
    TRY(allreduce(c, d, mx, gm, 1, ncclMax));
    TRY(allreduce(c, d, sm, gs, 2, ncclSum));
    red3[0] = gm[0]; red3[1] = gs[0]; red3[2] = gs[1];
    return SPHGPU_OK;
}
