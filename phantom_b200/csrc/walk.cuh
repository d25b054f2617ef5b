// walk.cuh -- warp-cooperative tree walk and pair-candidate scan shared by the density and force passes.
//
// Replaces getneigh + cache_neighbours (src/main/kdtree.F90:1221-1347, :1175-1213): one WARP owns one leaf cell.
//   * walk: the warp pops up to 16 tree nodes per step; lane L tests child (L&1) of node (L>>1) against the cell's
//     box grown by the search radius (periodic minimum-image gaps), pushes internal hits on a shared-memory stack
//     and appends the particles of leaf hits to the warp's candidate list (global scratch, L1/L2 resident).
//   * scan: for each target particle of the cell the lanes stride over the candidate list, apply the EXACT
//     reference distance test (non-contracted IEEE mul/add in the reference's association order, dens.F90:671-679 /
//     force.F90:1271-1287) and compact passing pairs through a shared-memory ring so that the expensive
//     pair body always runs with (nearly) full warps.
#pragma once
#include "common.cuh"

#define WALK_STACK 256
#define QRING 64

struct WarpShared {
    int stack[WALK_STACK];
    int qj[QRING];
    double qdx[QRING], qdy[QRING], qdz[QRING], qr2[QRING];
    double sums[48];
};

// squared minimum-image gap between two boxes
template <bool PERIODIC>
__device__ __forceinline__ double box_gap2(const double *tlo, const double *thi, double slo0, double slo1, double slo2, double shi0, double shi1,
                                           double shi2, double Lx, double Ly, double Lz)
{
    double g2 = 0.;
    {
        const double d1 = slo0 - thi[0], d2 = tlo[0] - shi0;
        double g = fmax(0., fmax(d1, d2));
        if (PERIODIC) g = fmax(0., fmin(g, fmin(d1, d2) + Lx));
        g2 += g * g;
    }
    {
        const double d1 = slo1 - thi[1], d2 = tlo[1] - shi1;
        double g = fmax(0., fmax(d1, d2));
        if (PERIODIC) g = fmax(0., fmin(g, fmin(d1, d2) + Ly));
        g2 += g * g;
    }
    {
        const double d1 = slo2 - thi[2], d2 = tlo[2] - shi2;
        double g = fmax(0., fmax(d1, d2));
        if (PERIODIC) g = fmax(0., fmin(g, fmin(d1, d2) + Lz));
        g2 += g * g;
    }
    return g2;
}

// Returns the number of candidate particles written to list[], or -1 when cap is exceeded.
// rcut_t: search radius of the target cell (radkern*hmax, already including any safety margin)
// SYM: also open nodes whose own radkern*hmax reaches the target box (force pass, get_hj of kdtree.F90:1288-1291)
template <bool SYM, bool PERIODIC>
__device__ int warp_walk(const TreeNode *__restrict__ nodes, const Cell *__restrict__ cells, int ncells, const double *tlo, const double *thi,
                         double rcut_t, double radkern, double Lx, double Ly, double Lz, int *__restrict__ list, int cap, int *stack)
{
    const int lane = lane_id();
    const double safe = 1.0 + 1e-9;   // guards the gap arithmetic's rounding; the exact test follows in the scan
    if (ncells == 1) {
        const int cnt = cells[0].count;
        if (cnt > cap) return -1;
        for (int k = lane; k < cnt; k += 32) list[k] = cells[0].start + k;
        __syncwarp();
        return cnt;
    }
    int sp = 1, nlist = 0;
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
        if (node >= 0) {
            const TreeNode *nd = &nodes[node];
            child = nd->child[slot];
            double rc = rcut_t;
            if (SYM) rc = fmax(rc, radkern * nd->hmax[slot]);
            rc *= safe;
            const double g2 = box_gap2<PERIODIC>(tlo, thi, nd->lo[slot][0], nd->lo[slot][1], nd->lo[slot][2], nd->hi[slot][0], nd->hi[slot][1],
                                                 nd->hi[slot][2], Lx, Ly, Lz);
            hit = g2 < rc * rc;
        }
        const unsigned mint = __ballot_sync(FULLMASK, hit && child >= 0);
        const unsigned mleaf = __ballot_sync(FULLMASK, hit && child < 0);
        if (sp + __popc(mint) > WALK_STACK) return -1;
        if (hit && child >= 0) stack[sp + __popc(mint & ((1u << lane) - 1))] = child;
        sp += __popc(mint);
        if (mleaf) {
            int cnt = 0, start = 0;
            if (hit && child < 0) { const Cell *cl = &cells[~child]; cnt = cl->count; start = cl->start; }
            int incl = cnt;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, s); if (lane >= s) incl += t; }
            const int total = __shfl_sync(FULLMASK, incl, 31);
            if (nlist + total > cap) return -1;
            int *dst = list + nlist + (incl - cnt);
            for (int k = 0; k < cnt; k++) dst[k] = start + k;
            nlist += total;
        }
        __syncwarp();
    }
    return nlist;
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
