// common.cuh -- context, device buffers and small device helpers shared by the kernels of the
// SPH hot path.  Hand-written for sm_100a; no fallback path exists.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/sphgpu.h"

#define FULLMASK 0xffffffffu

// particle types (src/main/part.F90:428-438)
enum { IGAS = 1, IBOUNDARY = 3, IDUST = 7 };

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grow while preserving the first `keep` elements (used when ghost particles are appended to the local set)
    cudaError_t ensure_keep(size_t n, size_t keep, cudaStream_t st)
    {
        if (n <= cap) return cudaSuccess;
        T *q = nullptr;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc((void **)&q, want * sizeof(T));
        if (e != cudaSuccess) return e;
        if (p && keep) { e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return e; }
        if (want > keep) cudaMemsetAsync(q + keep, 0, (want - keep) * sizeof(T), st);
        cudaStreamSynchronize(st);
        if (p) cudaFree(p);
        p = q; cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// kernel constants (kernel_cubic.f90:23-30, kernel_quintic.f90:23-30)
struct KernConsts { double radkern, radkern2, cnormk, wab0, gradh0, dphidh0, cnormk_drag; };

// parameters as seen by device code
struct DevParams {
    sphgpu_params p;
    KernConsts kc;
    double dxbound, dybound, dzbound;
    int nvu, ngradh, nalpha;
};

// tree node of the cell hierarchy (binary radix tree over leaf cells).  Children boxes are stored in the parent
// so that one 128-byte read tests both children.
struct __align__(16) TreeNode {
    double lo[2][3];   // child bbox min
    double hi[2][3];   // child bbox max
    double hmax[2];    // child max smoothing length
    int child[2];      // >= 0: internal node id ; < 0: ~cell id
    int parent;
    int pad;
    int cnt[2];        // particles below each child
    int start[2];      // first sorted slot of each child's range
    int act[2];        // active particles below each child
};

// FP32 copy used by the walk: boxes rounded OUTWARD, hmax rounded up (conservative tests), 64 bytes = 2 sectors
struct __align__(16) TreeNodeF {
    float lo[2][3];
    float hi[2][3];
    float hmax[2];
    int child[2];      // >= 0: internal node id ; < 0: leaf, bits 0-30 = (first sorted slot << 5) | (count - 1)
};

// node of the reference-topology kd-tree as the force pass of the reference-compatible neighbour mode needs it (gravity.cu builds it)
struct __align__(16) RefNode { double xcen[3], size, hmax; int parent, pad; };

struct Cell {          // leaf cell = run of <= max_cell Morton-consecutive particles
    double lo[3], hi[3];
    double hmax;
    int start, count;
    int active;        // number of active particles
    int parent;        // internal node owning this leaf
};

struct sphgpu_ctx {
    int device = 0;
    std::string err;
    DevParams hp;              // host copy
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[16];
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // sphgpu_derivs: host<->device copies pipelined against the passes
    cudaEvent_t cev[6];
    int64_t bytes_h2d = 0, bytes_d2h = 0;   // bytes the last sphgpu_derivs moved over PCIe
    int numSMs = 148;
    int64_t launches = 0;
    double ms_phase[4] = {0, 0, 0, 0};
    double ms_kernel[2] = {0, 0};   // k_density, k_force alone (CUDA events on the launching stream)
    double ms_gravity[2] = {0, 0};  // whole gravity pass, k_g_p2p alone
    int64_t npairs_gravity = 0, nm2l = 0;
    int grav_p2p_per_particle = 48; // capacity of the leaf P2P lists (source leaves per particle)
    struct GravState *grav = nullptr;
    bool grav_tree_valid = false;
    double hscale = 1.;             // largest growth factor of any h since the tree's hmax were refitted (force walk inflates hmax by it)
    double dens_hgrow = 0., halo_hgrow = 0.;   // largest h_new/h_old of the last density pass (local) ; global value handed in by the halo driver
    double dens_hmax_used = 0.;     // largest trial h any active particle took during the last density pass (halo sufficiency check)
    DevBuf<double> h_build, h_hist; // h at build_tree and after each h-rho iteration (replayed by k_g_hmax_leaf)
    DevBuf<int> h_its;
    // tuning
    int max_cell = 32;       // target group: <= 32 particles (one lane per target)
    int max_leaf = 8;        // tree leaf (source granularity of the walk)
    int group_pack = 0;      // > 0: target groups packed from whole leaf cells inside subtrees of <= group_pack particles (tree.cu k_groups_packed)
    double list_margin = 1.02;
    int scratch_per_warp = 4096;    // capacity (cells) of a warp's cell list
    // ---- canonical (original particle order) device arrays = device mirror of part.F90 ----
    int64_t npart = 0;
    DevBuf<double> xyzh, vxyzu, fxyzu, fext, Bevol, dBevol, eos_vars, Bxyz;
    DevBuf<float> divcurlv, divcurlB, alphaind, gradh, dvdx, poten, divBsymm;
    DevBuf<int8_t> iphase, ibin, ibin_old, ibin_wake;
    DevBuf<double> dustfrac, tstop;
    DevBuf<double> forc_tab; int forc_nmodes = 0, forc_correct_mean = 0; double forc_fac = 0.;   // turbulent driving mode table (forcing.f90)
    DevBuf<double> v_true, B_true;          // step.cu: the evolved v, B/rho while vxyzu/Bevol hold the predicted values
    DevBuf<double> twas;                    // step.cu, individual timesteps: the time each particle's v sits at (part.F90 twas)
    DevBuf<double4> gacc;                   // far-field gravity {fx,fy,fz,pot} per particle (gravity.cu -> force epilogue)
    // ---- sorted working set ----
    int64_t nlive = 0;
    bool tree_valid = false, dens_valid = false;
    DevBuf<unsigned long long> keys, keys_alt;
    DevBuf<int> perm, perm_alt;            // sorted slot -> original index
    DevBuf<double4> pos4;                   // x,y,z,h   (sorted)
    DevBuf<double4> vel4, acc4, bev4;       // density inputs (sorted)
    DevBuf<int8_t> stype;                   // iphase (sorted)
    DevBuf<double> hnew;                    // density output h (sorted)
    DevBuf<double4> frecC, frecD, frecE;    // force j-side records (sorted)
    DevBuf<double4> drec;                   // packed records of the single-type fast density path (4 x 32 B per particle)
    DevBuf<double4> frec;                   // packed records of the all-gas fast path (5 x 32 B per particle)
    bool hilbert = false;                   // space-filling curve of the particle order: Morton (default) or Hilbert (option "hilbert" 1; measured equal on B200)
    bool always_refit = false;              // option: refit the tree's hmax after every density pass (A/B testing)
    bool force_general = false;             // option: route everything through the general force kernel (A/B testing)
    // Individual timesteps: the reference's force walk can MISS a pair that only an inactive neighbour j reaches, because a leaf's hmax
    // is overwritten by 1.01 max(h) over its ACTIVE members when the leaf re-walks during the h-rho iteration (dens.F90:343-345,
    // :1275-1289; neigh_kdtree.f90:115-131) and the walk prunes on that (kdtree.F90:1288-1293).  refcompat = 1 (the default whenever
    // ind_timesteps is set; option "refcompat_hmax") drops exactly those pairs: the reference's own tree is built (gravity.cu), its node
    // hmax history replayed, and a pair with q2i >= R^2, q2j < R^2 is kept only if the reference's walk from i's leaf reaches j's leaf.
    // refcompat = 0 evaluates the pair criterion q2i < R^2 .or. q2j < R^2 exactly (every pair, as an O(N^2) search would).
    int refcompat = -1;                     // -1: follow ind_timesteps
    DevBuf<RefNode> ref_nodes; DevBuf<int> ref_leaf, ref_leaf_sorted;   // reference tree nodes ; leaf of every particle (caller order / sorted slot)
    bool ref_valid = false;
    DevBuf<float> s_gradh, s_divv, s_dvdx, s_alpha3, s_divcurlB;   // sorted density outputs
    DevBuf<double4> s_fxyzu, s_dB;          // sorted force outputs
    DevBuf<float> s_divvf, s_poten, s_divBsymm;
    DevBuf<int> s_nneigh;
    // tree
    int64_t ncells = 0;
    DevBuf<unsigned char> cpl;
    DevBuf<int> cellflag, cellid_scan;
    DevBuf<Cell> cells;
    DevBuf<Cell> groups;                    // target groups = maximal subtrees with <= max_cell particles
    int64_t ngroups = 0;
    DevBuf<unsigned long long> cellkeys;
    DevBuf<TreeNode> nodes;
    DevBuf<TreeNodeF> nodesf;
    bool multitype = false;                 // any particle that is not plain gas (boundary, dust, ...)
    bool any_inactive = true;               // some live particle carries the inactive flag (tree build); false: every particle is a target
    bool wl_ordered = false;                // wl_order holds the groups of the current lists by falling length
    bool eos_on_device = false;             // eos_vars (P, c_s) were written by cons2prim_run after the last upload of that array
    bool no_iso1 = false;                   // option: keep the three-sector force records for isothermal sets too (A/B, tests)
    bool dens_reuse = false;                // the last density pass iterated: keep staged rounds and masks over the h-rho iterations
    int class_mask = 1;                     // sort classes present (bit 0 gas/boundary, 1 dust, 2 other): the class rounds of the general pair kernels
    DevBuf<int> wl_list, wl_ncl; DevBuf<float> wl_reach;   // cell lists prepared by k_walk_lists (walk.cuh)
    DevBuf<int> wl_key, wl_key2, wl_iota, wl_order;        // target groups by falling list length (walk.cuh: walk_order_run)
    bool stream_blocking = false;   // the compute stream synchronises implicitly with the legacy default stream (option "legacy_stream")
    double dens_trial_hint = 0., dens_trial_max = 0.;   // mean / max candidates per target group in the last density pass (choose the round size of the next)
    bool wl_force_ok = false;       // the prepared lists are symmetric lists of the current groups ...
    double wl_cover = 0.;           // ... and cover every h up to wl_cover x the tree's hmax (force pass reuses them while hscale <= wl_cover)
    int walk_cap = 192;             // cells per prepared list (longer lists are walked inside the pair kernel)
    DevBuf<int> stage_idx;                  // per-warp cell lists of the pair kernels (candidates themselves are staged in shared memory)
    DevBuf<int> nodeflag;
    DevBuf<char> cubtemp;
    DevBuf<int> scratch;                    // per-warp candidate lists
    DevBuf<unsigned long long> counters;    // device scalars block
    DevBuf<double> dscal;                   // device double scalars (bbox, dt minima, ...)
    sphgpu_scalars last_dens{}, last_force{};
    int nbinmax = 0, ibinnow = 0, istepfrac = 0;   // timestep_ind module state
    DevBuf<int8_t> s_ibin, s_ibinold, s_ibinnew;   // sorted copies for the force pass
    DevBuf<int> s_wake;
    DevBuf<double> s_gsoft, s_tstop, s_dustfrac;
    // ---- multi-GPU halo state (halo.cu): ghosts are appended after the nlocal owned particles as inactive particles ----
    int64_t nlocal = 0, nghost = 0;
    int halo_nranks = 1, halo_rank = 0;
    std::vector<long long> halo_sendcnt, halo_sendoff;
    DevBuf<int> halo_sendidx;
    DevBuf<unsigned long long> halo_cnt;
    DevBuf<double> halo_boxes, halo_sendbuf, halo_recvbuf;
    // ---- multi-GPU driver behind the C ABI (dist.cu): NCCL communicator, domain boxes, exchange blocks; global particle ids
    struct DistState *dist = nullptr;
    DevBuf<long long> gid;
};

#define CUDA_TRY(ctx, call)                                                                           \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                         \
            return SPHGPU_ERR_CUDA;                                                                   \
        }                                                                                             \
    } while (0)

#define TRY(call)                       \
    do {                                \
        int r__ = (call);               \
        if (r__ != SPHGPU_OK) return r__; \
    } while (0)

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// rho(h) = m (hfact/h)^3 (part.F90:779)
__device__ __forceinline__ double rhoh_d(double hi, double pmassi, double hfact)
{
    double r = hfact / fabs(hi);
    return pmassi * (r * r * r);
}

// Sort class of a particle type: 0 = gas and boundary, 1 = dust, 2 = everything else.  The class is the top of the sort key, so every
// leaf cell, every target group and every staged round of candidates holds ONE class and the pair kernels pick their pair body per
// (class of the targets, class of the candidates) for the whole warp instead of branching per lane.  An all-gas set is class 0
// throughout: same order as a plain Morton sort.
__device__ __forceinline__ int sort_class(int8_t iphase)
{
    const int ta = iphase < 0 ? -iphase : iphase;
    return (ta == IGAS || ta == IBOUNDARY) ? 0 : (ta == IDUST ? 1 : 2);
}

// decode iphase (part.F90:1026-1067), gas + boundary + one dust type
__device__ __forceinline__ void get_partinfo_d(int8_t iphasei, int set_boundaries_to_active, int use_dust, bool &isactive, bool &isgas,
                                               bool &isdust, int &itype)
{
    if (iphasei >= 0) { isactive = true; itype = iphasei; }
    else { isactive = false; itype = -iphasei; }
    isgas = (itype == IGAS || itype == IBOUNDARY);
    isdust = use_dust ? (itype == IDUST) : false;
    if (itype == IBOUNDARY) {
        // part.F90:1052-1060 activates every boundary particle while set_boundaries_to_active holds (the first density pass of a run, when
        // all particles are active anyway).  One handed over with the inactive flag set stays inactive here: that is how the ghost copies
        // of the multi-GPU halo arrive, and a ghost must never become a target.
        if (set_boundaries_to_active && iphasei > 0) { isactive = true; itype = IGAS; }
        else isactive = false;
    }
}

// 1/sqrt(x) for a normal x > 0 (a third-order step on the hardware seed, as the CUDA library does, without its special-case branch);
// returns 0 for x == 0 or subnormal: coincident particles, for which the reference takes rij1 = 1/(0 + epsilon) times zero separation
__device__ __forceinline__ double rsqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(e, 0.375, 0.5);
    y = fma(p * e, y, y);
    return (__double2hiint(x) >= 0x00100000) ? y : 0.;
}

// max/min without the NaN canonicalisation of fmax/fmin (3 instructions instead of ~7); operands here are never NaN
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(FULLMASK, v, s);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v = fmax(v, __shfl_xor_sync(FULLMASK, v, s));
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v = fmin(v, __shfl_xor_sync(FULLMASK, v, s));
    return v;
}

// butterfly "transpose" reduction of 32 per-lane partial sums: on return lane L holds sum over lanes of v[L].
// 31 shuffles instead of 32*5.
__device__ __forceinline__ double warp_transpose_reduce32(double (&v)[32])
{
    const int lane = lane_id();
#pragma unroll
    for (int s = 16, cnt = 32; s >= 1; s >>= 1, cnt >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < cnt / 2; k++) {
            const double send = upper ? v[k] : v[k + cnt / 2];
            const double keep = upper ? v[k + cnt / 2] : v[k];
            v[k] = keep + __shfl_xor_sync(FULLMASK, send, s);
        }
    }
    return v[0];
}

// atomic min/max on non-negative doubles through their ordered bit patterns
__device__ __forceinline__ void atomic_min_pos(double *addr, double v) { atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v)); }
__device__ __forceinline__ void atomic_max_pos(double *addr, double v) { atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v)); }

// indices into ctx->counters (unsigned long long)
enum { CNT_WORK = 0, CNT_ERR, CNT_ERRID, CNT_NPAIRS, CNT_NTRIAL, CNT_NCALC, CNT_NACT, CNT_MAXACT, CNT_MAXTRIAL, CNT_NP, CNT_NWALK, CNT_NLIVE, CNT_NBINMAX, CNT_NCHECKBIN, CNT_MULTITYPE, CNT_NSURV,
       CNT_NGRAVPAIRS = 24, CNT_NM2L = 25, CNT_NCELLS = 26, CNT_CELLOVER = 27, CNT_CLASS1 = 28, CNT_CLASS2 = 29, CNT_ANYINACTIVE = 30, CNT_COUNT = 32 };
// indices into ctx->dscal (double)
enum { DS_XMIN = 0, DS_YMIN, DS_ZMIN, DS_XMAX, DS_YMAX, DS_ZMAX, DS_DTCOURANT, DS_DTFORCE, DS_DTMINI, DS_DTMAXI, DS_RHOMAX, DS_HUSED, DS_HGROW, DS_COUNT = 32 };

// internal API between translation units
int tree_build(sphgpu_ctx *c);
int tree_refit_hmax(sphgpu_ctx *c);
int density_run(sphgpu_ctx *c, int icall, sphgpu_scalars *out);
int cons2prim_run(sphgpu_ctx *c);
int force_run(sphgpu_ctx *c, int icall, double dt, sphgpu_scalars *out);
int gravity_run(sphgpu_ctx *c);
int refcompat_prepare(sphgpu_ctx *c);      // reference tree + node hmax history + leaf of every particle (gravity.cu)
// (with every particle active the reference's leaf hmax cover every member's kernel, as with global timesteps: nothing to replicate)
static inline bool refcompat_on(const sphgpu_ctx *c) { return (c->refcompat < 0 ? c->hp.p.ind_timesteps != 0 : c->refcompat != 0) && c->any_inactive; }
void gravity_release(sphgpu_ctx *c);
int gravity_gather_pack(sphgpu_ctx *c, void **sendptr, int *record_doubles);
int gravity_gather_recvbuf(sphgpu_ctx *c, int nranks, int64_t stride, void **recvptr);
int gravity_gather_unpack(sphgpu_ctx *c, int nranks, int myrank, int64_t stride, const int64_t *counts);
int64_t gravity_tree_dump(sphgpu_ctx *c, int64_t maxnodes, double *rec12, int32_t *irec6, int32_t *ids);
#define SPHGPU_HHIST 6   // h-rho iterations logged per particle for the node-hmax replay
int64_t neighbour_sets_run(sphgpu_ctx *c, int symmetric, int64_t *offsets, int32_t *list, int64_t maxlist);
KernConsts make_kern_consts(int kernel);
int ensure_all_keep(sphgpu_ctx *c, int64_t n, int64_t keep);
int sphgpu_dist_hook_derivs(sphgpu_ctx *c, int icall, double dt, sphgpu_scalars *out);
void sphgpu_dist_mark_dirty(sphgpu_ctx *c);     // positions, h or boxes of a distributed context changed (upload, step, migration)
int sphgpu_dist_hook_reduce_err(sphgpu_ctx *c, double *red3);
