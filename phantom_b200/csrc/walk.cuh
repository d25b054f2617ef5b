// walk.cuh -- warp-cooperative tree walk, candidate staging and prefiltered pair scan shared by the density and
// force passes.  Replaces getneigh + cache_neighbours (src/main/kdtree.F90:1221-1347, :1175-1213).
//
// One WARP owns one leaf cell:
//   walk   : pops up to 16 tree nodes per step; lane L tests child (L&1) of node (L>>1) with FP32 boxes that were
//            rounded OUTWARD (conservative), periodic minimum-image gaps; internal hits go to a shared-memory stack,
//            leaf hits to a small shared-memory cell list.
//   stage  : the particles of the hit cells are copied ONCE per cell into the warp's scratch slice as
//            float4 {x,y,z relative to the target-cell centre (nearest periodic image), radkern*h_j} + int index, so the
//            per-target scan reads contiguous 16-byte records (like the reference's xyzcache, but FP32 and only a filter).
//   masks  : lane = staged candidate, loop over the cell's <= 32 targets (broadcast from shared memory): a conservative
//            FP32 distance test (error bound derived from the staged extent) + one ballot per target gives, per chunk of
//            32 candidates, a 32-bit hit mask per target; ~10 instructions per (chunk, target), none on the FP64 pipe.
//   pairs  : lane = TARGET.  Every lane walks its own hit masks and evaluates its own neighbours; the pair body
//            re-evaluates the EXACT reference test in FP64 (non-contracted mul/add in the reference's association
//            order, dens.F90:671-679 / force.F90:1271-1287) so set membership is bit-identical.  The per-particle
//            sums stay in the lane's registers: no cross-lane reduction, no queue, no synchronisation in the pair loop.
#pragma once
#include "common.cuh"

#define WALK_STACK 256
#define CELLLIST 128
#define MAXCHUNK 64            // hit masks cover 64 chunks x 32 = 2048 staged candidates per round

struct WarpShared {
    int stack[WALK_STACK];
    int celllist[CELLLIST];             // packed (start << 5) | (count - 1)
    unsigned hm[MAXCHUNK][32];          // hm[chunk][t] = candidates of the chunk inside target t's (FP32, conservative) radius
    float4 tgt[32];                     // per target: cell-relative position + FP32 limit on r^2
};

struct Staged {
    float4 *pos;     // relative position + radkern*h_j (rounded up)
    int *idx;        // sorted particle slot
    int n;           // number of staged candidates
    float maxrel;    // max |relative coordinate| staged (for the FP32 error bound)
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

// copy the particles of the cells in ws.celllist[0..ncl) into the staging slice; 4 cells per step, 8 lanes per cell
template <bool PERIODIC>
__device__ __forceinline__ bool stage_cells(const WarpShared &ws, int ncl, const double4 *__restrict__ pos4, double cx, double cy, double cz, double Lx,
                                            double Ly, double Lz, float radkern, Staged &st, int cap)
{
    const int lane = lane_id();
    const int sub = lane >> 3, l8 = lane & 7;
    for (int c0 = 0; c0 < ncl; c0 += 4) {
        int start = 0, cnt = 0;
        if (c0 + sub < ncl) { const int pk = ws.celllist[c0 + sub]; start = pk >> 5; cnt = (pk & 31) + 1; }
        // exclusive offsets of the (up to) 4 cells of this step
        const int c1 = __shfl_sync(FULLMASK, cnt, 0), c2 = __shfl_sync(FULLMASK, cnt, 8), c3 = __shfl_sync(FULLMASK, cnt, 16), c4 = __shfl_sync(FULLMASK, cnt, 24);
        const int off = (sub > 0 ? c1 : 0) + (sub > 1 ? c2 : 0) + (sub > 2 ? c3 : 0);
        const int total = c1 + c2 + c3 + c4;
        if (st.n + total > cap) return false;
        for (int k = l8; k < cnt; k += 8) {
            const int j = start + k;
            const double4 p = pos4[j];
            double rx = p.x - cx, ry = p.y - cy, rz = p.z - cz;
            if (PERIODIC) {
                if (rx > 0.5 * Lx) rx -= Lx; else if (rx < -0.5 * Lx) rx += Lx;
                if (ry > 0.5 * Ly) ry -= Ly; else if (ry < -0.5 * Ly) ry += Ly;
                if (rz > 0.5 * Lz) rz -= Lz; else if (rz < -0.5 * Lz) rz += Lz;
            }
            const float fx = (float)rx, fy = (float)ry, fz = (float)rz;
            st.pos[st.n + off + k] = make_float4(fx, fy, fz, __double2float_ru((double)radkern * p.w));
            st.idx[st.n + off + k] = j;
            st.maxrel = fmaxf(st.maxrel, fmaxf(fabsf(fx), fmaxf(fabsf(fy), fabsf(fz))));
        }
        st.n += total;
    }
    return true;
}

// Walk + stage.  tlo/thi: target-cell box (FP32, outward rounded); rcut_t: search radius of the target cell (incl. margin);
// SYM: also open nodes whose own radkern*hmax reaches the target box (force pass, get_hj of kdtree.F90:1288-1291).
// Returns false when the scratch slice (cap) or the stack overflows.
template <bool SYM, bool PERIODIC>
__device__ bool warp_walk_stage(const TreeNodeF *__restrict__ nodes, const Cell *__restrict__ cells, int ncells, const double4 *__restrict__ pos4,
                                const float *tlo, const float *thi, float rcut_t, float radkern, double cx, double cy, double cz, double Lx, double Ly,
                                double Lz, WarpShared &ws, Staged &st, int cap)
{
    const int lane = lane_id();
    st.n = 0; st.maxrel = 0.f;
    const float fLx = (float)Lx, fLy = (float)Ly, fLz = (float)Lz;
    int ncl = 0;
    if (ncells == 1) {
        if (lane == 0) ws.celllist[0] = (cells[0].start << 5) | (cells[0].count - 1);
        __syncwarp();
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
            if (node >= 0) {
                const TreeNodeF *nd = &nodes[node];
                child = nd->child[slot];
                float rc = rcut_t;
                if (SYM) rc = fmaxf(rc, radkern * nd->hmax[slot]);
                rc *= 1.00001f;
                const float g2 = box_gap2f<PERIODIC>(tlo, thi, nd->lo[slot][0], nd->lo[slot][1], nd->lo[slot][2], nd->hi[slot][0], nd->hi[slot][1],
                                                     nd->hi[slot][2], fLx, fLy, fLz);
                hit = g2 <= rc * rc;
            }
            const unsigned mint = __ballot_sync(FULLMASK, hit && child >= 0);
            const unsigned mleaf = __ballot_sync(FULLMASK, hit && child < 0);
            if (sp + __popc(mint) > WALK_STACK) return false;
            if (hit && child >= 0) ws.stack[sp + __popc(mint & ((1u << lane) - 1))] = child;
            sp += __popc(mint);
            if (mleaf) {
                const int nl = __popc(mleaf);
                if (ncl + nl > CELLLIST) {           // flush the cell list into the staging slice
                    __syncwarp();
                    if (!stage_cells<PERIODIC>(ws, ncl, pos4, cx, cy, cz, Lx, Ly, Lz, radkern, st, cap)) return false;
                    ncl = 0;
                    __syncwarp();
                }
                if (hit && child < 0) {
                    const Cell *cl = &cells[~child];
                    ws.celllist[ncl + __popc(mleaf & ((1u << lane) - 1))] = (cl->start << 5) | (cl->count - 1);
                }
                ncl += nl;
            }
            __syncwarp();
        }
    }
    if (ncl > 0 && !stage_cells<PERIODIC>(ws, ncl, pos4, cx, cy, cz, Lx, Ly, Lz, radkern, st, cap)) return false;
    st.maxrel = fmaxf(st.maxrel, __shfl_xor_sync(FULLMASK, st.maxrel, 16));
    st.maxrel = fmaxf(st.maxrel, __shfl_xor_sync(FULLMASK, st.maxrel, 8));
    st.maxrel = fmaxf(st.maxrel, __shfl_xor_sync(FULLMASK, st.maxrel, 4));
    st.maxrel = fmaxf(st.maxrel, __shfl_xor_sync(FULLMASK, st.maxrel, 2));
    st.maxrel = fmaxf(st.maxrel, __shfl_xor_sync(FULLMASK, st.maxrel, 1));
    __syncwarp();
    return true;
}

__device__ __forceinline__ float prefilter_slack(float maxrel);
__device__ __forceinline__ float prefilter_limit(float rc, float slack);

// hit masks for one round of staged candidates [base, base + nchunk*32): lane = candidate, loop over targets
template <bool SYM>
__device__ __forceinline__ void build_masks(WarpShared &ws, const Staged &st, int base, int nchunk, int ntargets, float slack)
{
    const int lane = lane_id();
    for (int c = 0; c < nchunk; c++) {
        const int i = base + c * 32 + lane;
        const bool valid = i < st.n;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) r = st.pos[i];
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

// generic transpose reduction for N = 8,16,32 partial sums per lane; lane L ends with the total of v[L >> (5 - log2 N)]
template <int N>
__device__ __forceinline__ double warp_transpose_reduce(double (&v)[N])
{
    const int lane = lane_id();
    int s = 16;
#pragma unroll
    for (int cnt = N; cnt > 1; cnt >>= 1, s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < cnt / 2; k++) {
            const double send = upper ? v[k] : v[k + cnt / 2];
            const double keep = upper ? v[k + cnt / 2] : v[k];
            v[k] = keep + __shfl_xor_sync(FULLMASK, send, s);
        }
    }
    double r = v[0];
    for (; s >= 1; s >>= 1) r += __shfl_xor_sync(FULLMASK, r, s);
    return r;
}
