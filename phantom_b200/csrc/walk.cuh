// walk.cuh -- warp-cooperative tree walk, shared-memory candidate staging and prefiltered pair scan shared by the density and
// force passes.  Replaces getneigh + cache_neighbours (src/main/kdtree.F90:1221-1347, :1175-1213).
//
// One WARP owns one target group (<= 32 particles, lane = target in the pair loop):
//   walk   : pops up to 16 tree nodes per step; lane L tests child (L&1) of node (L>>1) with FP32 boxes that were
//            rounded OUTWARD (conservative), periodic minimum-image gaps; internal hits go to a shared-memory stack,
//            leaf hits are appended to the group's CELL LIST (one packed int per 8-particle leaf; a few hundred bytes per
//            group in a per-warp global slice -- the only global scratch the pair kernels touch).
//   rounds : the candidates are consumed in rounds of <= ROUND particles.  Each round copies the particles of the next
//            cells ONCE into SHARED memory as float4 {x,y,z relative to the target-group centre (nearest periodic image),
//            radkern*h_j} + slot index -- the reference's xyzcache (kdtree.F90:1175), but FP32, on chip and only a filter.
//   masks  : lane = staged candidate, loop over the group's targets (broadcast from shared memory): a conservative
//            FP32 distance test (error bound derived from the staged extent) + one ballot per target gives, per chunk of
//            32 candidates, a 32-bit hit mask per target; ~10 instructions per (chunk, target), none on the FP64 pipe.
//   pairs  : lane = TARGET.  Every lane walks its own hit masks and evaluates its own neighbours; the pair body
//            re-evaluates the EXACT reference test in FP64 (non-contracted mul/add in the reference's association
//            order, dens.F90:671-679 / force.F90:1271-1287) so set membership is bit-identical.  The per-particle
//            sums stay in the lane's registers: no cross-lane reduction, no queue, no synchronisation in the pair loop.
#pragma once
#include "common.cuh"

#define WALK_STACK 256
#ifndef ROUND
#define ROUND 384              // staged candidates per round
#endif
#define NCHUNK (ROUND / 32)    // hit-mask chunks per round

struct WarpShared {
    int stack[WALK_STACK];
    float4 spos[ROUND];                 // staged candidates: relative position + radkern*h_j (rounded up)
    int sidx[ROUND];                    // their sorted particle slots
    unsigned hm[NCHUNK][32];            // hm[chunk][t] = candidates of the chunk inside target t's (FP32, conservative) radius
    float4 tgt[32];                     // per target: group-relative position + FP32 limit on r^2
};

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
                         float radkern, float fLx, float fLy, float fLz, WarpShared &ws, int *__restrict__ clist, int cap, float &reach)
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
        if (lane == 0) ws.stack[0] = 0;
        __syncwarp();
        while (sp > 0) {
            int npop = min(16, sp);
            if (sp + npop > WALK_STACK - 2) npop = 1;
            const int slot = lane & 1, which = lane >> 1;
            int node = -1;
            if (which < npop) node = ws.stack[sp - 1 - which];
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
            if (hit && child >= 0) ws.stack[sp + __popc(mint & ((1u << lane) - 1))] = child;
            sp += __popc(mint);
            if (mleaf) {
                const int nl = __popc(mleaf);
                if (ncl + nl > cap) return -1;
                if (leafhit) {
                    const Cell *cl = &cells[~child];
                    clist[ncl + __popc(mleaf & ((1u << lane) - 1))] = (cl->start << 5) | (cl->count - 1);
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

// Copy the particles of the next cells of the list into the shared-memory round buffer: 4 cells per step, 8 lanes per cell.
// posrec[j * stride] = {x, y, z, w} with w = h (WINV = false) or 1/h (WINV = true).  Returns the number staged; cellpos advances.
template <bool PERIODIC, bool WINV>
__device__ __forceinline__ int stage_round(WarpShared &ws, const int *__restrict__ clist, int ncl, int &cellpos, const double4 *__restrict__ posrec, int stride,
                                           double cx, double cy, double cz, double Lx, double Ly, double Lz, float radkern, int maxleaf)
{
    const int lane = lane_id();
    const int sub = lane >> 3, l8 = lane & 7;
    int n = 0;
    while (cellpos < ncl && n + 4 * maxleaf <= ROUND) {  // the next 4 cells (<= maxleaf particles each) always fit
        int start = 0, cnt = 0;
        if (cellpos + sub < ncl) { const int pk = clist[cellpos + sub]; start = pk >> 5; cnt = (pk & 31) + 1; }
        // exclusive offsets of the (up to) 4 cells of this step
        const int c1 = __shfl_sync(FULLMASK, cnt, 0), c2 = __shfl_sync(FULLMASK, cnt, 8), c3 = __shfl_sync(FULLMASK, cnt, 16), c4 = __shfl_sync(FULLMASK, cnt, 24);
        const int off = n + (sub > 0 ? c1 : 0) + (sub > 1 ? c2 : 0) + (sub > 2 ? c3 : 0);
        for (int k = l8; k < cnt; k += 8) {
            const int j = start + k;
            const double4 p = posrec[(size_t)j * stride];
            double rx = p.x - cx, ry = p.y - cy, rz = p.z - cz;
            if (PERIODIC) {
                if (rx > 0.5 * Lx) rx -= Lx; else if (rx < -0.5 * Lx) rx += Lx;
                if (ry > 0.5 * Ly) ry -= Ly; else if (ry < -0.5 * Ly) ry += Ly;
                if (rz > 0.5 * Lz) rz -= Lz; else if (rz < -0.5 * Lz) rz += Lz;
            }
            float rkh;
            if (WINV) rkh = __fmul_ru(radkern, __frcp_ru(__double2float_rd(p.w)));   // >= radkern * h_j
            else rkh = __double2float_ru((double)radkern * p.w);
            ws.spos[off + k] = make_float4((float)rx, (float)ry, (float)rz, rkh);
            ws.sidx[off + k] = j;
        }
        n += c1 + c2 + c3 + c4;
        cellpos += 4;
    }
    __syncwarp();
    return n;
}

__device__ __forceinline__ float prefilter_slack(float maxrel);
__device__ __forceinline__ float prefilter_limit(float rc, float slack);

// hit masks for the n staged candidates of the round: lane = candidate, loop over targets
template <bool SYM>
__device__ __forceinline__ void build_masks(WarpShared &ws, int n, int ntargets, float slack)
{
    const int lane = lane_id();
    const int nchunk = (n + 31) >> 5;
    for (int c = 0; c < nchunk; c++) {
        const int i = c * 32 + lane;
        const bool valid = i < n;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) r = ws.spos[i];
        float limj = 0.f;
        if (SYM) limj = prefilter_limit(r.w, slack);
        unsigned mine = 0u;
#pragma unroll 4
        for (int t = 0; t < ntargets; t++) {
            const float4 tg = ws.tgt[t];
            const float ax = tg.x - r.x, ay = tg.y - r.y, az = tg.z - r.z;
            const float r2 = fmaf(az, az, fmaf(ay, ay, ax * ax));
            const float lim = SYM ? ((tg.w > 0.f) ? fmaxf(tg.w, limj) : 0.f) : tg.w;
            const unsigned m = __ballot_sync(FULLMASK, valid && (r2 < lim));
            if (lane == t) mine = m;
        }
        ws.hm[c][lane] = mine;
    }
    __syncwarp();
}

// advance this lane to its next hit; returns the staged slot or -1 when the lane has consumed all its hits of the round
__device__ __forceinline__ int next_hit(const WarpShared &ws, int lane, int nchunk, int &c, unsigned &m)
{
    while (m == 0u) {
        if (++c >= nchunk) return -1;
        m = ws.hm[c][lane];
    }
    const int bit = __ffs(m) - 1;
    m &= m - 1;
    return c * 32 + bit;
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

// FP32 prefilter limits.  A staged coordinate differs from the exact relative coordinate by at most 2^-24*maxrel; so does the
// target's.  |r_f - r| <= sqrt(3)*2*2^-24*maxrel =: e.  For a true pair r < rc  =>  r_f^2 < (rc + e)^2 (1 + 4 ulp).
__device__ __forceinline__ float prefilter_slack(float maxrel) { return 2.1e-7f * maxrel + 1e-30f; }
__device__ __forceinline__ float prefilter_limit(float rc, float slack)
{
    const float r = rc + slack;
    return r * r * 1.000002f;
}
