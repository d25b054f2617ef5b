// walk.cuh -- warp-cooperative tree walk, shared-memory candidate staging and prefiltered pair scan shared by the density and
// force passes.  Replaces getneigh + cache_neighbours (src/main/kdtree.F90:1221-1347, :1175-1213).
//
// One WARP owns one target group (<= 32 particles, lane = target in the pair loop):
//   walk   : pops up to 16 tree nodes per step; lane L tests child (L&1) of node (L>>1) with FP32 boxes that were
//            rounded OUTWARD (conservative), periodic minimum-image gaps; internal hits go to a shared-memory stack,
//            leaf hits are appended to the group's CELL LIST (one packed int per 8-particle leaf; a few hundred bytes per
//            group in a per-warp global slice -- the only global scratch the pair kernels touch).
//   rounds : the candidates are consumed in rounds of <= ROUND particles.  Each round copies the particles of the next
//            cells ONCE into SHARED memory as packed FP16 {x,y,z relative to the target-group centre (nearest periodic image)
//            in units of the staged extent, limit on r^2} + slot index -- the reference's xyzcache (kdtree.F90:1175), but on
//            chip, two candidates per 16-byte word, and only a filter.
//   masks  : lane = TARGET (its own scaled position in registers), loop over candidate PAIRS broadcast from shared memory:
//            a conservative half2 distance test (error bound below) decides two candidates in 9 instructions and ORs the
//            two result bits into the lane's 32-bit hit mask of the chunk; nothing on the FP64 pipe, no ballots.
//   pairs  : lane = TARGET.  Every lane walks its own hit masks and evaluates its own neighbours; the pair body
//            re-evaluates the EXACT reference test in FP64 (non-contracted mul/add in the reference's association
//            order, dens.F90:671-679 / force.F90:1271-1287) so set membership is bit-identical.  The per-particle
//            sums stay in the lane's registers: no cross-lane reduction, no queue, no synchronisation in the pair loop.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>
#include <cub/cub.cuh>

#define WALK_STACK 256
#ifndef ROUND_DEFAULT
#define ROUND_DEFAULT 384      // staged candidates per round
#endif

// Per-warp staging block.  ROUND_ candidates per round; the fast kernels also keep P2_ double2 parts + P1_ double parts of each
// candidate (its FP64 position) here: the L1 data pipe was the measured limiter of the pair loops (82-93 % busy with wavefronts,
// one per lane and 16-byte load of a gathered record), and a shared-memory gather of 32 different addresses costs a few wavefronts.
// OWNSTACK = false: the walk stack (only used by the rare in-kernel walk, before anything is staged) lives in the rec2 area.
template <int ROUND_, int P2_, int P1_, bool OWNSTACK = true>
struct __align__(16) WarpSharedT {
    static constexpr int ROUND = ROUND_, NCHUNK = ROUND_ / 32, P2 = P2_, P1 = P1_;
    static_assert(OWNSTACK || (size_t)P2_ * ROUND_ * 16 >= WALK_STACK * sizeof(int), "rec2 too small to hold the walk stack");
    int stack[OWNSTACK ? WALK_STACK : 4];
    __device__ __forceinline__ int *walk_stack() { return OWNSTACK ? stack : reinterpret_cast<int *>(&rec2[0][0]); }
    uint4 hp[NCHUNK][16];               // staged candidates, two per word: half2 {x, y, z, limit}; slot c*32+b sits in word b&15, half b>>4
    int sidx[ROUND_];                   // their sorted particle slots
    unsigned hm[NCHUNK][32];            // hm[chunk][t] = candidates of the chunk inside target t's (FP16, conservative) radius
    int selfslot[32];                   // slot at which target t itself is staged in this round, -1 if it is not
    unsigned nzsave[32];                // density pass: the lane's non-empty mask words and the radius its masks were built for, kept
    float rmask[32];                    // over the h-rho iterations of a group that was staged in one round (out of the registers)
    double2 rec2[P2_ > 0 ? P2_ : 1][P2_ > 0 ? ROUND_ : 1];     // structure of arrays: conflict-free staging stores
    double rec1[P1_ > 0 ? P1_ : 1][P1_ > 0 ? ROUND_ : 1];
};
typedef WarpSharedT<ROUND_DEFAULT, 0, 0> WarpShared;     // neighbour-list kernel (static shared memory)
#ifndef ROUND_GENERAL
#define ROUND_GENERAL 384
#endif
typedef WarpSharedT<ROUND_GENERAL, 0, 0> WarpSharedGeneral;   // general pair kernels (dynamic shared memory).  768 measured: two-fluid box 48.9 -> 45.9 ms but dusty disc 541 -> 583 ms (less L1 left for its five gathers per pair)

// Scale of the FP16 filter of one target group: coordinates relative to the group centre times `scale` lie in [-1, 1].
struct FilterScale { float scale, slack; };
// The lane's own target: scaled position and limit on the scaled r^2, each duplicated in both halves
struct FilterTarget { __half2 x, y, z, lim; };

// squared minimum-image gap between two boxes, FP32 (inputs already rounded outward)
template <bool PERIODIC>
__device__ __forceinline__ float box_gap2f(const float *tlo, const float *thi, float slo0, float slo1, float slo2, float shi0, float shi1, float shi2,
                                           float Lx, float Ly, float Lz)
{
    float g2 = 0.f;
    {
        const float d1 = slo0 - thi[0], d2 = tlo[0] - shi0;
        float g = fmaxf(0.f, fmaxf(d1, d2));
        if (PERIODIC) g = fmaxf(0.f, fminf(g, fminf(d1, d2) + Lx));
        g2 = fmaf(g, g, g2);
    }
    {
        const float d1 = slo1 - thi[1], d2 = tlo[1] - shi1;
        float g = fmaxf(0.f, fmaxf(d1, d2));
        if (PERIODIC) g = fmaxf(0.f, fminf(g, fminf(d1, d2) + Ly));
        g2 = fmaf(g, g, g2);
    }
    {
        const float d1 = slo2 - thi[2], d2 = tlo[2] - shi2;
        float g = fmaxf(0.f, fmaxf(d1, d2));
        if (PERIODIC) g = fmaxf(0.f, fminf(g, fminf(d1, d2) + Lz));
        g2 = fmaf(g, g, g2);
    }
    return g2;
}

// Walk.  tlo/thi: target-group box (FP32, outward rounded); rcut_t: search radius of the group (incl. margin);
// SYM: also open nodes whose own radkern*hmax reaches the target box (force pass, get_hj of kdtree.F90:1288-1291).
// Fills clist[0..ncl) with the packed hit cells ((start << 5) | (count - 1)) and returns ncl, or -1 when the list (cap) or the
// stack overflows.  reach = max over hit cells of (search radius used + cell extent): every staged coordinate relative to the
// group centre is bounded by halfext + reach (input of the FP32 error bound).
template <bool SYM, bool PERIODIC>
__device__ int warp_walk(const TreeNodeF *__restrict__ nodes, const Cell *__restrict__ cells, int ncells, const float *tlo, const float *thi, float rcut_t,
                         float radkern, float fLx, float fLy, float fLz, int *stack, int *__restrict__ clist, int cap, float &reach)
{
    const int lane = lane_id();
    int ncl = 0;
    float rmax = 0.f;
    if (ncells == 1) {
        const Cell c0 = cells[0];
        if (lane == 0) clist[0] = (c0.start << 5) | (c0.count - 1);
        rmax = fmaxf(rcut_t, SYM ? radkern * (float)c0.hmax * 1.0001f : 0.f) + (float)fmax(c0.hi[0] - c0.lo[0], fmax(c0.hi[1] - c0.lo[1], c0.hi[2] - c0.lo[2])) * 1.0001f;
        ncl = 1;
    } else {
        int sp = 1;
        if (lane == 0) stack[0] = 0;
        __syncwarp();
        while (sp > 0) {
            int npop = min(16, sp);
            if (sp + npop > WALK_STACK - 2) npop = 1;
            const int slot = lane & 1, which = lane >> 1;
            int node = -1;
            if (which < npop) node = stack[sp - 1 - which];
            sp -= npop;
            __syncwarp();
            bool hit = false;
            int child = 0;
            float ext = 0.f;
            if (node >= 0) {
                const TreeNodeF *nd = &nodes[node];
                child = nd->child[slot];
                float rc = rcut_t;
                if (SYM) rc = fmaxf(rc, radkern * nd->hmax[slot]);
                rc *= 1.00001f;
                const float l0 = nd->lo[slot][0], l1 = nd->lo[slot][1], l2 = nd->lo[slot][2], h0 = nd->hi[slot][0], h1 = nd->hi[slot][1], h2 = nd->hi[slot][2];
                const float g2 = box_gap2f<PERIODIC>(tlo, thi, l0, l1, l2, h0, h1, h2, fLx, fLy, fLz);
                hit = g2 <= rc * rc;
                ext = rc + fmaxf(h0 - l0, fmaxf(h1 - l1, h2 - l2)) * 1.0001f;
            }
            const bool leafhit = hit && child < 0;
            const unsigned mint = __ballot_sync(FULLMASK, hit && child >= 0);
            const unsigned mleaf = __ballot_sync(FULLMASK, leafhit);
            if (sp + __popc(mint) > WALK_STACK) return -1;
            if (hit && child >= 0) stack[sp + __popc(mint & ((1u << lane) - 1))] = child;
            sp += __popc(mint);
            if (mleaf) {
                const int nl = __popc(mleaf);
                if (ncl + nl > cap) return -1;
                if (leafhit) {
                    clist[ncl + __popc(mleaf & ((1u << lane) - 1))] = child & 0x7fffffff;
                    rmax = fmaxf(rmax, ext);
                }
                ncl += nl;
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(FULLMASK, rmax, s));
    reach = rmax;
    __syncwarp();
    return ncl;
}

// ---- cell lists ahead of the pair kernels ----------------------------------------------------------------------------------------
// The walk is a chain of dependent node reads (one L2 round trip per level) with almost no arithmetic: inside a pair kernel, at 12-16
// warps per SM, it costs ~15 % of the time in scoreboard stalls.  k_walk_lists does it for every target group in a kernel of its own
// (one warp per group, a 1 KB stack per warp, 48+ warps per SM hide the latency) and leaves the packed cell lists in global memory;
// the pair kernels read them back coalesced.  A group whose list exceeds `cap` gets ncl = -1 and is walked by the pair kernel itself
// into its per-warp slice, as is every re-walk of the density iteration.
struct WalkLists {
    int *list;          // [ngroups][cap] packed cells
    int *ncl;           // [ngroups] cells in the list, -1: not prepared
    float *reach;       // [ngroups] see warp_walk
    int cap;
    const int *order;   // [ngroups] groups by falling list length, or NULL: the order in which the persistent warps take them
};

template <bool SYM, bool PERIODIC>
__global__ void __launch_bounds__(256) k_walk_lists(const TreeNodeF *__restrict__ nodes, const Cell *__restrict__ cells, int ncells,
                                                    const Cell *__restrict__ groups, int ngroups, double rfac, float radkern, float fLx, float fLy, float fLz,
                                                    WalkLists wl)
{
    __shared__ int stacks[8][WALK_STACK];
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= ngroups) return;
    const int lane = lane_id();
    const Cell &cell = groups[g];
    if (cell.active == 0) { if (lane == 0) wl.ncl[g] = 0; return; }
    float tlo[3], thi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { tlo[k] = __double2float_rd(cell.lo[k]); thi[k] = __double2float_ru(cell.hi[k]); }
    float reach = 0.f;
    const int ncl = warp_walk<SYM, PERIODIC>(nodes, cells, ncells, tlo, thi, __double2float_ru(rfac * cell.hmax), radkern, fLx, fLy, fLz, stacks[threadIdx.x >> 5],
                                             wl.list + (size_t)g * wl.cap, wl.cap, reach);
    if (lane == 0) { wl.ncl[g] = ncl; wl.reach[g] = reach; }
}

// host side: prepare the lists of all target groups for search radius rfac * hmax(group) (SYM: or the candidates' own radius)
static inline int walk_lists_run(sphgpu_ctx *c, bool sym, double rfac, double radkern_eff, WalkLists &wl)
{
    const int ng = (int)c->ngroups;
    CUDA_TRY(c, c->wl_list.ensure((size_t)ng * c->walk_cap)); CUDA_TRY(c, c->wl_ncl.ensure(ng)); CUDA_TRY(c, c->wl_reach.ensure(ng));
    wl.list = c->wl_list.p; wl.ncl = c->wl_ncl.p; wl.reach = c->wl_reach.p; wl.cap = c->walk_cap; wl.order = nullptr;
    const float rk = nextafterf((float)radkern_eff, 3.0e38f);
    const float fLx = (float)c->hp.dxbound, fLy = (float)c->hp.dybound, fLz = (float)c->hp.dzbound;
    const int grid = (ng + 7) / 8;
    const bool per = c->hp.p.periodic;
#define WALK_LAUNCH(S, P) k_walk_lists<S, P><<<grid, 256, 0, c->stream>>>(c->nodesf.p, c->cells.p, (int)c->ncells, c->groups.p, ng, rfac, rk, fLx, fLy, fLz, wl)
    if (sym) { if (per) WALK_LAUNCH(true, true); else WALK_LAUNCH(true, false); }
    else { if (per) WALK_LAUNCH(false, true); else WALK_LAUNCH(false, false); }
#undef WALK_LAUNCH
    c->launches++;
    return SPHGPU_OK;
}

// Heavy groups first.  A set with a few groups whose lists are many times the mean (the contact of the shock tube, the inner rim of a
// disc) otherwise ends on whichever warp picked one of them up last; the persistent kernels take the groups in the order of falling
// list length instead.  Only when the previous pass saw such a tail (max > 2 x mean candidates): a uniform set gains nothing.
static // (key = list length in units of twice the mean, capped at 7; the sort is stable, so the bulk -- key 0 -- keeps its Morton order and
// with it the cache locality between consecutive groups: a full sort by length made the dusty disc 10 % slower)
__global__ void k_walk_order_keys(int ng, const int *__restrict__ ncl, int unit, int *__restrict__ key, int *__restrict__ iota)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ng) return;
    const int n = ncl[g];
    key[g] = n < 0 ? 7 : min(n / unit, 7);                        // -1: list longer than the cap, walked by the pair kernel itself
    iota[g] = g;
}
static inline int walk_order_run(sphgpu_ctx *c, WalkLists &wl)
{
    wl.order = nullptr;
    if (!(c->dens_trial_hint > 0. && c->dens_trial_max > 2. * c->dens_trial_hint)) return SPHGPU_OK;
    const int ng = (int)c->ngroups;
    CUDA_TRY(c, c->wl_key.ensure(ng)); CUDA_TRY(c, c->wl_key2.ensure(ng)); CUDA_TRY(c, c->wl_iota.ensure(ng)); CUDA_TRY(c, c->wl_order.ensure(ng));
    const int unit = std::max(1, (int)(2. * c->dens_trial_hint / 6.));          // twice the mean list length in cells (a leaf cell holds ~6 particles)
    k_walk_order_keys<<<(ng + 255) / 256, 256, 0, c->stream>>>(ng, wl.ncl, unit, c->wl_key.p, c->wl_iota.p);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, c->wl_key.p, c->wl_key2.p, c->wl_iota.p, c->wl_order.p, ng, 0, 3, c->stream);
    CUDA_TRY(c, c->cubtemp.ensure(tb));
    size_t tbb = c->cubtemp.cap;
    CUDA_TRY(c, cub::DeviceRadixSort::SortPairsDescending(c->cubtemp.p, tbb, c->wl_key.p, c->wl_key2.p, c->wl_iota.p, c->wl_order.p, ng, 0, 3, c->stream));
    c->launches += 5;
    wl.order = c->wl_order.p;
    return SPHGPU_OK;
}

// ---- FP16 prefilter -------------------------------------------------------------------------------------------------------
// Every staged coordinate u = scale*(x - centre) has |u| <= 1 (scale = 1/(halfext + reach), see warp_walk).  Rounding u to FP16 errs
// by <= 2^-11 |u|, so a coordinate DIFFERENCE of a target (|u_t| <= tau = scale*halfext) and a candidate, rounded once more,
// errs by <= 2^-11 (|u_t| + |u_c| + |a|) <= 2^-11 (2 tau + 2 rho) whenever the true |a| <= rho, the scaled kernel radius of the
// pair (rho <= 1).  Over three axes: |r_f - r| <= sqrt(3) 2^-10 (tau + 1) =: slack.  The three half2 roundings of the sum of squares
// add <= 1.5e-3 relatively.  Hence a true pair (r < rho) always satisfies r_f^2 < (rho + slack)^2 * 1.002, with the limit rounded
// UP to FP16; flushed subnormals only lower r_f^2.  The filter may pass non-neighbours (about 2 % here); the FP64 test decides.
__device__ __forceinline__ FilterScale filter_scale(float halfext, float reach)
{
    FilterScale fs;
    const float ext = halfext * 1.0001f + reach + 1e-30f;
    fs.scale = __frcp_rd(ext);
    fs.slack = 1.7321f * 9.77e-4f * (halfext * 1.0001f * __frcp_rd(ext) * 1.0001f + 1.f) + 1e-6f;
    return fs;
}
// Thin periodic boxes (the shock tube's y and z): a target group whose search sphere reaches beyond half a box length cannot use the
// nearest image relative to the group CENTRE for every target.  While the kernel radii stay well below half the box (0.45 L) the
// filter takes the minimum image per pair itself, a' = a - L rint(a / L) in the scaled FP16 coordinates (three more half2 operations
// per axis), instead of being switched off and leaving every candidate to the FP64 test.  Error: the rounding of L to FP16 and of the
// fused multiply-add add <= 2^-11 (L_s + |a'|) per axis with L_s = L x scale <= ~2.2 here; filter_scale_wrap widens the slack for it.
struct FilterWrap { __half2 L[3], iL[3]; };
__device__ __forceinline__ FilterWrap filter_wrap(const FilterScale &fs, float Lx, float Ly, float Lz)
{
    FilterWrap w;
    const float L[3] = {Lx, Ly, Lz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const __half l = __float2half_rn(fminf(L[k] * fs.scale, 3.0e4f));       // a direction without a (near) image: n = rint(a / 3e4) = 0
        w.L[k] = __half2half2(l);
        w.iL[k] = __float2half2_rn(1.f / __half2float(l));
    }
    return w;
}
__device__ __forceinline__ void filter_scale_wrap(FilterScale &fs) { fs.slack = 3.f * fs.slack + 2.0e-3f; }

// limit on the scaled, FP16-evaluated r^2 for a kernel radius rc (>= the exact one); rc < 0: pass everything
__device__ __forceinline__ __half filter_limit(const FilterScale &fs, float rc)
{
    if (rc < 0.f) return __ushort_as_half((unsigned short)0x7c00);           // +inf
    const float r = __fmul_ru(rc, fs.scale * 1.000001f) + fs.slack;
    return __float2half_ru(r * r * 1.0021f + 1e-7f);
}
// rc: kernel radius of the target (rounded up), 0 for a target that takes no part, < 0 to switch the filter off (wide periodic search)
__device__ __forceinline__ FilterTarget filter_target(const FilterScale &fs, float rx, float ry, float rz, float rc)
{
    FilterTarget t;
    t.x = __float2half2_rn(rx * fs.scale); t.y = __float2half2_rn(ry * fs.scale); t.z = __float2half2_rn(rz * fs.scale);
    t.lim = __half2half2(rc == 0.f ? __ushort_as_half((unsigned short)0) : filter_limit(fs, rc));
    return t;
}

// Copy the particles of the next cells of the list into the shared-memory round buffer.  Two passes so that no load waits on another:
// (1) lane = cell, 32 cells per step: one coalesced read of the packed list, a warp scan of the counts, slot -> particle index;
// (2) lane = slot: the position records of all slots are fetched independently, scaled to FP16 and stored.
// w = h (WINV = false) or 1/h (WINV = true).  Returns the number staged; cellpos advances.  cls >= 0: only the cells of that sort
// class are staged (general kernels: one pair body per class pair for the whole warp); a round may then come back empty.
struct NoRecord { __device__ __forceinline__ void operator()(int, int, const double2 &, const double2 &) const {} };
// posrec2[j * stride2] = {x, y}, [j * stride2 + 1] = {z, w}.  gstart: first sorted slot of the target group (selfslot bookkeeping);
// stage_rec(slot, j, xy, zw): whatever else the caller wants staged.
template <bool PERIODIC, bool WINV, class WS, class F = NoRecord>
__device__ __forceinline__ int stage_round(WS &ws, const int *__restrict__ clist, int ncl, int &cellpos, const double2 *__restrict__ posrec2, int stride2,
                                           double cx, double cy, double cz, double Lx, double Ly, double Lz, float radkern, int maxleaf,
                                           const FilterScale &fs, bool interior = false, int gstart = 0, F stage_rec = F(), int cls = -1,
                                           const int8_t *__restrict__ stype = nullptr)
{
    constexpr int ROUND = WS::ROUND;
    ws.selfslot[lane_id()] = -1;
    const int lane = lane_id();
    int n = 0;
    while (cellpos < ncl) {
        const int nb = min(32, min(ncl - cellpos, (ROUND - n) / maxleaf));      // cells (<= maxleaf particles each) that are sure to fit
        if (nb <= 0) break;
        int start = 0, cnt = 0;
        if (lane < nb) { const int pk = clist[cellpos + lane]; start = pk >> 5; cnt = (pk & 31) + 1; }
        if (cls >= 0 && cnt > 0 && sort_class(stype[start]) != cls) cnt = 0;     // rounds of one sort class (cells never mix classes)
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += t; }
        const int base = n + incl - cnt;
        for (int k = 0; k < cnt; k++) ws.sidx[base + k] = start + k;
        n += __shfl_sync(FULLMASK, incl, 31);
        cellpos += nb;
    }
    __syncwarp();
#pragma unroll 4
    for (int slot = lane; slot < n; slot += 32) {
        const int j = ws.sidx[slot];
        const double2 pxy = posrec2[(size_t)j * stride2], pzw = posrec2[(size_t)j * stride2 + 1];
        const double4 p = make_double4(pxy.x, pxy.y, pzw.x, pzw.y);
        double rx = p.x - cx, ry = p.y - cy, rz = p.z - cz;
        if (PERIODIC && !interior) {                  // interior: no candidate of this group lies across the periodic boundary
            if (rx > 0.5 * Lx) rx -= Lx; else if (rx < -0.5 * Lx) rx += Lx;
            if (ry > 0.5 * Ly) ry -= Ly; else if (ry < -0.5 * Ly) ry += Ly;
            if (rz > 0.5 * Lz) rz -= Lz; else if (rz < -0.5 * Lz) rz += Lz;
        }
        float rkh;
        if (WINV) rkh = __fmul_ru(radkern, __frcp_ru(__double2float_rd(p.w)));   // >= radkern * h_j
        else rkh = __double2float_ru((double)radkern * p.w);
        // |u| <= 1 by construction; the clamp only keeps a stray value finite (it could not be a neighbour: limits are <= ~1)
        const float ux = fminf(fmaxf((float)rx * fs.scale, -8.f), 8.f), uy = fminf(fmaxf((float)ry * fs.scale, -8.f), 8.f),
                    uz = fminf(fmaxf((float)rz * fs.scale, -8.f), 8.f);
        __half *w = reinterpret_cast<__half *>(&ws.hp[slot >> 5][slot & 15]) + ((slot >> 4) & 1);
        w[0] = __float2half_rn(ux); w[2] = __float2half_rn(uy); w[4] = __float2half_rn(uz); w[6] = filter_limit(fs, rkh);
        const unsigned t = (unsigned)(j - gstart);
        if (t < 32u) ws.selfslot[t] = slot;
        stage_rec(slot, j, pxy, pzw);
    }
    __syncwarp();
    return n;
}

// hit masks for the n staged candidates of the round: lane = target, loop over candidate pairs.  SYM: a pair passes when it is
// inside the target's OR the candidate's radius (force pass); targets with limit 0 (inactive, converged) get empty masks.
// Returns the lane's non-empty chunks as a bit mask (NCHUNK <= 32).
template <bool SYM, bool WRAP = false, class WS>
__device__ __forceinline__ unsigned build_masks(WS &ws, int n, const FilterTarget &t, const FilterWrap *fw = nullptr)
{
    unsigned nz = 0u;
    const int lane = lane_id();
    const int nchunk = (n + 31) >> 5;
    const bool takes_part = __low2float(t.lim) > 0.f;
    for (int c = 0; c < nchunk; c++) {
        unsigned mine = 0u;
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const uint4 q = ws.hp[c][p];
            __half2 ax = __hsub2(t.x, *reinterpret_cast<const __half2 *>(&q.x)), ay = __hsub2(t.y, *reinterpret_cast<const __half2 *>(&q.y)),
                    az = __hsub2(t.z, *reinterpret_cast<const __half2 *>(&q.z));
            if (WRAP) {                                       // minimum image per pair (thin periodic box)
                ax = __hfma2(__hneg2(h2rint(__hmul2(ax, fw->iL[0]))), fw->L[0], ax);
                ay = __hfma2(__hneg2(h2rint(__hmul2(ay, fw->iL[1]))), fw->L[1], ay);
                az = __hfma2(__hneg2(h2rint(__hmul2(az, fw->iL[2]))), fw->L[2], az);
            }
            const __half2 r2 = __hfma2(az, az, __hfma2(ay, ay, __hmul2(ax, ax)));
            const __half2 lim = SYM ? __hmax2(t.lim, *reinterpret_cast<const __half2 *>(&q.w)) : t.lim;
            mine |= __hlt2_mask(r2, lim) & (0x00010001u << p);
        }
        const int left = n - c * 32;                          // slots beyond n hold stale data
        if (left < 32) mine &= (1u << left) - 1u;
        if (!takes_part) mine = 0u;
        ws.hm[c][lane] = mine;
        nz |= (mine != 0u) ? (1u << c) : 0u;
    }
    __syncwarp();
    return nz;
}

// advance this lane to its next hit; returns the staged slot or -1 when the lane has consumed all its hits of the round
template <class WS>
__device__ __forceinline__ int next_hit(const WS &ws, int lane, int nchunk, int &c, unsigned &m)
{
    while (m == 0u) {
        if (++c >= nchunk) return -1;
        m = ws.hm[c][lane];
    }
    const int bit = __ffs(m) - 1;
    m &= m - 1;
    return c * 32 + bit;
}

// The same through an explicit 32-bit shared-window address.  The address is made opaque once per kernel so that the compiler keeps
// it in a register; otherwise it re-derives it (S2R + LEA + IMAD) at every use inside the pair loop.
template <class WS>
__device__ __forceinline__ unsigned ws_shared_addr(const WS &ws)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(&ws);
    asm volatile("mov.u32 %0, %0;" : "+r"(a));
    return a;
}
// One 256-bit load (sm_100: LDG.E.256) of a 32-byte-aligned double4.  A gathered record costs the L1 data pipe one wavefront per lane
// and load INSTRUCTION, so halving the instructions halves the load on the pipe that limits the pair loops.
__device__ __forceinline__ double4 ldg256(const double4 *p)
{
    double4 v;
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 lds_d2(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_d(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// hm_lane = shared address of ws.hm[0][lane]
__device__ __forceinline__ int next_hit_s(unsigned hm_lane, int nchunk, int &c, unsigned &m)
{
    while (m == 0u) {
        if (++c >= nchunk) return -1;
        m = lds_u32(hm_lane + 128u * (unsigned)c);
    }
    const int bit = __ffs(m) - 1;
    m &= m - 1;
    return c * 32 + bit;
}

// The next TWO hits of the lane in one convergent instruction sequence (the while-loop form above makes the warp replay its body once
// per distinct path of its lanes: measured 2.3 passes per call).  nz = the lane's not yet visited non-empty mask words (a bit per
// chunk, from build_masks): an exhausted word is replaced by the next non-empty one in a single step.  slot0 < 0: the lane is done.
__device__ __forceinline__ void next_hits2(unsigned hm_lane, unsigned &nz, int &c, unsigned &m, int &slot0, int &slot1)
{
    if (m == 0u && nz) { c = __ffs(nz) - 1; nz &= nz - 1u; m = lds_u32(hm_lane + 128u * (unsigned)c); }
    slot0 = m ? c * 32 + (__ffs(m) - 1) : -1;
    m &= m - 1u;
    if (m == 0u && nz) { c = __ffs(nz) - 1; nz &= nz - 1u; m = lds_u32(hm_lane + 128u * (unsigned)c); }
    slot1 = m ? c * 32 + (__ffs(m) - 1) : -1;
    m &= m - 1u;
}

// exact reference separation: dx = xi - xj, minimum image (dens.F90:666-670), rij2 = dx*dx + dy*dy + dz*dz evaluated
// left to right without FMA contraction so that set membership is bit-identical to the gfortran build
template <bool PERIODIC>
__device__ __forceinline__ double pair_r2(double xi, double yi, double zi, const double4 &pj, double Lx, double Ly, double Lz, double &dx, double &dy,
                                          double &dz)
{
    dx = xi - pj.x; dy = yi - pj.y; dz = zi - pj.z;
    if (PERIODIC) {
        if (fabs(dx) > 0.5 * Lx) dx = dx - copysign(Lx, dx);
        if (fabs(dy) > 0.5 * Ly) dy = dy - copysign(Ly, dy);
        if (fabs(dz) > 0.5 * Lz) dz = dz - copysign(Lz, dz);
    }
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

