// gravity.cu -- tree self-gravity by the fast multipole method: replaces the GRAVITY branches of maketree
// (construct_node moments, src/main/kdtree.F90:612-790), getneigh_dual + open_nodes + node_interaction (:1357-1700),
// compute_M2L (:1702-1781), propagate_fnode_to_node (:1527-1560), expand_fgrav_in_taylor_series (:1799-1840) and the
// 1/r^2 sum over trial neighbours outside both kernels (force.F90:1992-2053).
//
// Why a second tree.  The far field depends on WHICH node pairs are accepted as well separated, i.e. on the tree itself;
// to agree with the reference to 1e-8 (not to the ~1e-3 of the multipole acceptance error) the gravity pass uses the
// reference's own tree: top-down bisection at the centre of mass along the longest axis of the particle bounding box, leaves
// of <= 10 particles (kdtree.F90:531-929).  Only the construction is redesigned for the GPU:
//   build : LEVEL-synchronous.  Every level is three coalesced passes over the particles (segmented warp reductions feeding
//           one atomic per node and warp: mass, centre of mass, bounding box; then r2max and quadrupole moments and the
//           left/right flag; then a stable partition through one device-wide prefix sum) and two small node kernels.
//   hmax  : node hmax follows the reference's history (built from the h before the density pass, raised by set_hmaxcell to
//           1.01*max(h) every time a leaf re-walks during the h-rho iteration, neigh_kdtree.f90:115-131, dens.F90:1275-1289):
//           the density kernel logs each particle's h per iteration and k_g_hmax replays that per leaf.
//   walk  : the reference walks (dst ancestor, src) pairs once per LEAF and caches the ancestors' results; here the same
//           interaction lists are produced once per NODE, breadth first: node d tests the sources its parent could not accept
//           (one warp per node, lane = source): accepted -> M2L into the lane's 20 Taylor coefficients, rejected -> the source
//           (if a leaf) or its two children go to d's own list.  Leaves keep opening until only leaf sources remain = the
//           P2P list.  F(d) = sum of M2L + L2L(F(parent)) is finished in the same kernel, so no separate downward pass exists.
//   P2P   : one warp per leaf, lane = source leaf, <= 10 targets in registers: Newtonian m_j/r^2 for every pair that is NOT
//           an SPH neighbour pair (those carry softened gravity and are summed by k_force), then L2P of F(leaf).
// Result: gacc[i] = {fx, fy, fz, phi} in the caller's particle order, consumed by the epilogue of k_force.
// FP64-pipe bound (P2P: ~25 flop per pair, M2L: ~130 flop per accepted node pair); no tensor-core shaped work.
#include "common.cuh"
#include "sphkern.cuh"
#include <cub/cub.cuh>
#include <float.h>
#include <string.h>

namespace {

constexpr int MINPART = 10;   // kdtree.F90:45
constexpr int LENF = 20;      // lenfgrav (kdtree.F90:47)

struct __align__(16) GNode {        // kdnode (dtype_kdtree.F90:53-68) + the particle range (inoderange)
    double xcen[3], size, hmax, mass, quads[6];
    int left, right, parent, start, count, level, flags, pad;      // left < 0: leaf ; flags bit0: leaf holds an active particle
};

struct GBuild {                     // per-node accumulators of the level being built
    double sm, sx, sy, sz, q[6], hmaxP;
    unsigned long long lo[3], hi[3], r2max;
    double pivot; int axis, split, nl, pad;
};

__device__ __forceinline__ unsigned long long enc_ord(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_ord(unsigned long long b)
{
    unsigned long long r = (b & 0x8000000000000000ull) ? (b & 0x7fffffffffffffffull) : ~b;
    return __longlong_as_double((long long)r);
}

// ---- segmented warp reductions over runs of equal keys (slots are ordered by node, so equal keys are contiguous) ----
__device__ __forceinline__ bool seg_bounds(int key, int &last)
{
    const unsigned peers = __match_any_sync(FULLMASK, key);
    last = 31 - __clz(peers);
    return lane_id() == (__ffs(peers) - 1);
}
__device__ __forceinline__ double seg_sum(double v, int last)
{
    const int lane = lane_id();
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const double o = __shfl_down_sync(FULLMASK, v, off); if (lane + off <= last) v += o; }
    return v;
}
__device__ __forceinline__ double seg_min(double v, int last)
{
    const int lane = lane_id();
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const double o = __shfl_down_sync(FULLMASK, v, off); if (lane + off <= last) v = fmin(v, o); }
    return v;
}
__device__ __forceinline__ double seg_max(double v, int last)
{
    const int lane = lane_id();
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const double o = __shfl_down_sync(FULLMASK, v, off); if (lane + off <= last) v = fmax(v, o); }
    return v;
}

// ---- build ---------------------------------------------------------------------------------------------------------------
// live particles in caller order -> slot arrays (construct_root_node, kdtree.F90:429-456)
__global__ void k_g_init(int64_t n, const double *__restrict__ xyzh, const int8_t *__restrict__ iphase, const DevParams dp, double4 *__restrict__ pos,
                         double *__restrict__ mass, int *__restrict__ gid, int *__restrict__ pnode, const int *__restrict__ rank)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 x = reinterpret_cast<const double4 *>(xyzh)[i];
    if (x.w < DBL_MIN) return;
    const int s = rank[i];
    pos[s] = x; mass[s] = dp.p.massoftype[abs((int)iphase[i])]; gid[s] = (int)i; pnode[s] = 0;
}
__global__ void k_g_liveflag(int64_t n, const double *__restrict__ xyzh, int *__restrict__ flag)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (xyzh[4 * i + 3] < DBL_MIN) ? 0 : 1;
}

// (the node range [lvl[L], lvl[L+1]) of the level being built is read from the device: the host does not wait for it, grids are sized
// for the largest range the level can have)
__global__ void k_g_node_reset(const int *__restrict__ lvl, int L, GBuild *__restrict__ gb)
{
    const int n0 = lvl[L], n1 = lvl[L + 1];
    const int d = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n1) return;
    GBuild b; memset(&b, 0, sizeof b);
    for (int k = 0; k < 3; k++) { b.lo[k] = ~0ull; b.hi[k] = 0ull; }
    gb[d] = b;
}

// block-level combine for the upper levels, where a whole 256-thread block lies inside one node: one atomic per block and value
// instead of one per warp (the root alone would otherwise serialise N/32 atomics on a single address)
template <int NV>
__device__ __forceinline__ bool block_uniform_sum(int key, bool head, double (&v)[NV], double (*sh)[NV], int *shkey)
{
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const bool wuni = __all_sync(FULLMASK, key == __shfl_sync(FULLMASK, key, 0)) && key >= 0;
    if (lane == 0) shkey[w] = wuni ? key : -2 - w;
    __syncthreads();
    bool buni = true;
    for (int k = 1; k < (int)(blockDim.x >> 5); k++) buni = buni && (shkey[k] == shkey[0]);
    buni = buni && shkey[0] >= 0;
    if (buni) {
        if (head) for (int k = 0; k < NV; k++) sh[w][k] = v[k];
        __syncthreads();
        if (threadIdx.x == 0) for (int k = 0; k < NV; k++) { double t = sh[0][k]; for (int q = 1; q < (int)(blockDim.x >> 5); q++) t += sh[q][k]; v[k] = t; }
    }
    return buni;
}

// pass 1: mass, mass-weighted position, bounding box per node of the current level (kdtree.F90:654-666, :867-900)
__global__ void __launch_bounds__(256) k_g_sums(int nlive, const int *__restrict__ pnode, const double4 *__restrict__ pos, const double *__restrict__ mass, double dfac,
                         GBuild *__restrict__ gb)
{
    __shared__ double sh[8][4]; __shared__ double shm[8][6]; __shared__ int shkey[8];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    int key = -1; double4 x = make_double4(0., 0., 0., 0.); double m = 0.;
    if (s < nlive) { key = pnode[s]; if (key >= 0) { x = pos[s]; m = mass[s]; } }
    int last; const bool head = seg_bounds(key, last);
    const double fac = m * dfac;
    double v[4] = {seg_sum(m, last), seg_sum(fac * x.x, last), seg_sum(fac * x.y, last), seg_sum(fac * x.z, last)};
    double lo[3] = {seg_min(x.x, last), seg_min(x.y, last), seg_min(x.z, last)}, hi[3] = {seg_max(x.x, last), seg_max(x.y, last), seg_max(x.z, last)};
    const bool buni = block_uniform_sum<4>(key, head, v, sh, shkey);
    if (buni) {                                   // min/max of the block through the same shared staging
        const int w = threadIdx.x >> 5;
        if (head) for (int k = 0; k < 3; k++) { shm[w][k] = lo[k]; shm[w][3 + k] = hi[k]; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 1; q < 8; q++) for (int k = 0; k < 3; k++) { lo[k] = fmin(lo[k], shm[q][k]); hi[k] = fmax(hi[k], shm[q][3 + k]); }
        }
    }
    if ((buni && threadIdx.x == 0) || (!buni && head && key >= 0)) {
        GBuild *b = &gb[key];
        atomicAdd(&b->sm, v[0]); atomicAdd(&b->sx, v[1]); atomicAdd(&b->sy, v[2]); atomicAdd(&b->sz, v[3]);
        atomicMin(&b->lo[0], enc_ord(lo[0])); atomicMin(&b->lo[1], enc_ord(lo[1])); atomicMin(&b->lo[2], enc_ord(lo[2]));
        atomicMax(&b->hi[0], enc_ord(hi[0])); atomicMax(&b->hi[1], enc_ord(hi[1])); atomicMax(&b->hi[2], enc_ord(hi[2]));
    }
}

// node kernel A: centre of mass, split decision, axis = first longest bbox axis, pivot = COM on that axis (kdtree.F90:668-700, :792-818)
__global__ void k_g_nodes_a(const int *__restrict__ lvl, int L, GNode *__restrict__ nodes, GBuild *__restrict__ gb, double dfac, unsigned long long *cnt)
{
    const int n0 = lvl[L], n1 = lvl[L + 1];
    const int d = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n1) return;
    GBuild &b = gb[d];
    GNode &nd = nodes[d];
    const double den = b.sm * dfac;
    if (!(b.sm > 0.)) { atomicMax(&cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_ARG); return; }   // mtree: totmass_node==0
    nd.xcen[0] = b.sx / den; nd.xcen[1] = b.sy / den; nd.xcen[2] = b.sz / den;
    nd.mass = b.sm;
    const bool split = nd.count > MINPART;
    int axis = 0;
    double best = dec_ord(b.hi[0]) - dec_ord(b.lo[0]);
    for (int k = 1; k < 3; k++) { const double e = dec_ord(b.hi[k]) - dec_ord(b.lo[k]); if (e > best) { best = e; axis = k; } }
    b.axis = axis; b.split = split ? 1 : 0; b.pivot = nd.xcen[axis];
}

// pass 2: size^2 = max |x - xcen|^2, quadrupole moments (kdtree.F90:734-752) and the left/right flag of sort_particles_in_cell (:937-1001)
// (QUADS = false: the tree is built for the reference-compatible neighbour mode alone, which needs centres, sizes and topology)
template <bool QUADS>
__global__ void __launch_bounds__(256) k_g_moments(int nlive, const int *__restrict__ pnode, const double4 *__restrict__ pos, const double *__restrict__ mass,
                            const GNode *__restrict__ nodes, GBuild *__restrict__ gb, int *__restrict__ flag)
{
    __shared__ double sh[8][7]; __shared__ int shkey[8];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    int key = -1; double4 x = make_double4(0., 0., 0., 0.); double m = 0.;
    if (s < nlive) { key = pnode[s]; if (key >= 0) { x = pos[s]; m = mass[s]; } }
    double dx = 0., dy = 0., dz = 0.; int fl = 0;
    if (key >= 0) {
        const GNode &nd = nodes[key];
        dx = x.x - nd.xcen[0]; dy = x.y - nd.xcen[1]; dz = x.z - nd.xcen[2];
        const GBuild &b = gb[key];
        if (b.split) { const double xa = (b.axis == 0) ? x.x : (b.axis == 1 ? x.y : x.z); fl = (xa <= b.pivot) ? 1 : 0; }
    }
    if (s < nlive) flag[s] = fl;
    int last; const bool head = seg_bounds(key, last);
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    double r2m = seg_max(r2, last);
    double v[7] = {0., 0., 0., 0., 0., 0., 0.};
    if (QUADS) {
        v[0] = seg_sum(m * (dx * dx), last); v[1] = seg_sum(m * (dx * dy), last); v[2] = seg_sum(m * (dx * dz), last);
        v[3] = seg_sum(m * (dy * dy), last); v[4] = seg_sum(m * (dy * dz), last); v[5] = seg_sum(m * (dz * dz), last);
    }
    const bool buni = block_uniform_sum<7>(key, head, v, sh, shkey);
    if (buni) {                                   // block maximum of r2 through slot 6
        const int w = threadIdx.x >> 5;
        __syncthreads();
        if (head) sh[w][6] = r2m;
        __syncthreads();
        if (threadIdx.x == 0) for (int q = 1; q < 8; q++) r2m = fmax(r2m, sh[q][6]);
    }
    if ((buni && threadIdx.x == 0) || (!buni && head && key >= 0)) {
        GBuild *b = &gb[key];
        atomicMax(&b->r2max, (unsigned long long)__double_as_longlong(r2m));
        if (QUADS) {
            atomicAdd(&b->q[0], v[0]); atomicAdd(&b->q[1], v[1]); atomicAdd(&b->q[2], v[2]);
            atomicAdd(&b->q[3], v[3]); atomicAdd(&b->q[4], v[4]); atomicAdd(&b->q[5], v[5]);
        }
    }
}

// node kernel B: finish the node record; split nodes get two children appended after the current level
__global__ void k_g_nodes_b(const int *__restrict__ lvl, int L, GNode *__restrict__ nodes, GBuild *__restrict__ gb, const int *__restrict__ flag, const int *__restrict__ scan,
                            int *nnodes, int maxnodes, unsigned long long *cnt)
{
    const int n0 = lvl[L], n1 = lvl[L + 1];
    const int d = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n1) return;
    GBuild &b = gb[d];
    GNode &nd = nodes[d];
    nd.size = sqrt(__longlong_as_double((long long)b.r2max)) + DBL_EPSILON;        // kdtree.F90:782
    for (int k = 0; k < 6; k++) nd.quads[k] = b.q[k];
    nd.hmax = 0.; nd.flags = 0;
    if (!b.split) { nd.left = nd.right = -1; return; }
    const int e = nd.start + nd.count - 1;
    int nl = scan[e] + flag[e] - scan[nd.start];
    if (nl == 0 || nl == nd.count) { nl = nd.count / 2; b.split = 2; }           // kdtree.F90:856-865: all on one side -> halve by position
    b.nl = nl;
    const int base = atomicAdd(nnodes, 2);
    if (base + 2 > maxnodes) { atomicMax(&cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); b.split = 0; nd.left = nd.right = -1; return; }
    nd.left = base; nd.right = base + 1;
    GNode l; memset(&l, 0, sizeof l);
    l.parent = d; l.level = nd.level + 1; l.left = l.right = -1;
    GNode r = l;
    l.start = nd.start; l.count = nl; r.start = nd.start + nl; r.count = nd.count - nl;
    nodes[base] = l; nodes[base + 1] = r;
}

// end of level L: the nodes allocated during it are level L + 1
__global__ void k_g_level_end(const int *__restrict__ nnodes, int *__restrict__ lvl, int L) { lvl[L + 2] = *nnodes; }

// pass 3: stable partition of every split node; particles of finished leaves keep their slot
__global__ void k_g_scatter(int nlive, const int *__restrict__ pnode, const double4 *__restrict__ pos, const double *__restrict__ mass,
                            const int *__restrict__ gid, const GNode *__restrict__ nodes, const GBuild *__restrict__ gb, const int *__restrict__ flag,
                            const int *__restrict__ scan, int *__restrict__ pnode2, double4 *__restrict__ pos2, double *__restrict__ mass2,
                            int *__restrict__ gid2)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    const int key = pnode[s];
    int dst = s, nk = key;
    if (key >= 0) {
        const GBuild &b = gb[key];
        const GNode &nd = nodes[key];
        if (b.split == 1) {
            const int rank = scan[s] - scan[nd.start];                         // lefts before me inside the node
            if (flag[s]) { dst = nd.start + rank; nk = nd.left; }
            else { dst = nd.start + b.nl + (s - nd.start - rank); nk = nd.right; }
        } else if (b.split == 2) {
            nk = (s - nd.start < b.nl) ? nd.left : nd.right;
        } else nk = ~key;                                                       // leaf: remember it as ~node
    }
    pnode2[dst] = nk; pos2[dst] = pos[s]; mass2[dst] = mass[s]; gid2[dst] = gid[s];
}

// ---- node hmax: replay of set_hmaxcell during the density iterations (see header) -------------------------------------------
__global__ void k_g_hmax_leaf(int nn, GNode *__restrict__ nodes, GBuild *__restrict__ gb, const int *__restrict__ gid, const double *__restrict__ hbuild,
                              const int *__restrict__ hits, const double *__restrict__ hhist, int64_t npart, int hk, const int8_t *__restrict__ iphase,
                              int ind_ts, int64_t own_lo, int64_t own_hi)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= nn) return;
    GNode &nd = nodes[d];
    if (nd.left >= 0) { gb[d].hmaxP = 0.; return; }
    double hb = 0.; int nits = 0, act = 0;
    for (int s = nd.start; s < nd.start + nd.count; s++) {
        const int i = gid[s];
        hb = fmax(hb, hbuild[i]);
        nits = max(nits, hits[i]);
        const bool own = (i >= own_lo && i < own_hi);                            // multi-GPU: this rank evaluates only leaves holding its own particles
        if (own) act |= 2;
        if (own && (!ind_ts || iphase[i] > 0)) act |= 1;
    }
    double cellh = hb, hset = hb, hmaxset = hb;
    for (int k = 1; k <= nits - 1; k++) {                                       // iterations after which the cell was not yet converged
        double hm = 0.;
        for (int s = nd.start; s < nd.start + nd.count; s++) {
            const int i = gid[s];
            const int ni = hits[i];
            if (ni <= 0) continue;                                              // not an active target of the density pass
            const int kk = min(min(k, ni - 1), hk);                             // h after iteration kk (frozen once converged)
            hm = fmax(hm, kk == 0 ? hbuild[i] : hhist[(size_t)(kk - 1) * npart + i]);
        }
        const double cand = 1.01 * hm;                                          // compute_hmax (dens.F90:1275-1289)
        if (cand > cellh) { hset = cand; hmaxset = fmax(hmaxset, cand); }       // set_hmaxcell (neigh_kdtree.f90:115-131)
        cellh = cand;
    }
    nd.hmax = hset; nd.flags = act;
    gb[d].hmaxP = hmaxset;
}
__global__ void k_g_hmax_up(int n0, int n1, GNode *__restrict__ nodes, GBuild *__restrict__ gb)
{
    const int d = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n1) return;
    GNode &nd = nodes[d];
    if (nd.left < 0) return;
    const double h = fmax(gb[nd.left].hmaxP, gb[nd.right].hmaxP);
    nd.hmax = h; gb[d].hmaxP = h;
    nd.flags = (nodes[nd.left].flags | nodes[nd.right].flags) & 2;               // some owned particle below
}

// ---- FMM pieces ------------------------------------------------------------------------------------------------------------
// compute_M2L (kdtree.F90:1702-1781): Taylor coefficients of -q0/r (+ quadrupole) about the destination centre
__device__ __forceinline__ void m2l(double dx, double dy, double dz, double dr1, double q0, const double *__restrict__ quads, double (&f)[LENF])
{
    const double dr12 = dr1 * dr1, dx2 = dx * dx, dx3 = dx * dx2, dy2 = dy * dy, dy3 = dy * dy2, dz2 = dz * dz, dz3 = dz * dz2;
    const double g0 = -dr1, g1 = dr12 * g0, g2 = -3. * dr12 * g1, g3 = -5. * dr12 * g2;
    const double g2dx = g2 * dx, g2dy = g2 * dy, g2dz = g2 * dz;
    double D3[10], D2[6], D1[3];
    D3[0] = 3. * g2dx + g3 * dx3; D3[1] = g2dy + g3 * dx2 * dy; D3[2] = g2dz + g3 * dx2 * dz; D3[3] = g2dx + g3 * dy2 * dx;
    D3[4] = g3 * dx * dy * dz; D3[5] = g2dx + g3 * dz2 * dx; D3[6] = 3. * g2dy + g3 * dy3; D3[7] = g2dz + g3 * dy2 * dz;
    D3[8] = g2dy + g3 * dz2 * dy; D3[9] = 3. * g2dz + g3 * dz3;
    D2[0] = g1 + g2 * dx2; D2[1] = g2dx * dy; D2[2] = g2dx * dz; D2[3] = g1 + g2 * dy2; D2[4] = g2dy * dz; D2[5] = g1 + g2 * dz2;
    D1[0] = g1 * dx; D1[1] = g1 * dy; D1[2] = g1 * dz;
    const double qxx = quads[0], qxy = quads[1], qxz = quads[2], qyy = quads[3], qyz = quads[4], qzz = quads[5];
    f[0] += D1[0] * q0 + 0.5 * (D3[0] * qxx + 2. * (D3[1] * qxy + D3[2] * qxz + D3[4] * qyz) + D3[3] * qyy + D3[5] * qzz);
    f[1] += D1[1] * q0 + 0.5 * (D3[1] * qxx + 2. * (D3[3] * qxy + D3[4] * qxz + D3[7] * qyz) + D3[6] * qyy + D3[8] * qzz);
    f[2] += D1[2] * q0 + 0.5 * (D3[2] * qxx + 2. * (D3[4] * qxy + D3[5] * qxz + D3[8] * qyz) + D3[7] * qyy + D3[9] * qzz);
#pragma unroll
    for (int k = 0; k < 6; k++) f[3 + k] += D2[k] * q0;
#pragma unroll
    for (int k = 0; k < 10; k++) f[9 + k] += D3[k] * q0;
    f[19] += g0 * q0 - 0.5 * (D2[0] * qxx + D2[3] * qyy + D2[5] * qzz + 2. * (D2[1] * qxy + D2[2] * qxz + D2[4] * qyz));
}

// propagate_fnode_to_node (kdtree.F90:1527-1560): second-order shift of the parent's expansion to the child centre
__device__ __forceinline__ void l2l(double (&f)[LENF], const double *__restrict__ fs, double dx, double dy, double dz)
{
    f[0] = fs[0] + dx * (fs[3] + 0.5 * (dx * fs[9] + dy * fs[10] + dz * fs[11])) + dy * (fs[4] + 0.5 * (dx * fs[10] + dy * fs[12] + dz * fs[13])) +
           dz * (fs[5] + 0.5 * (dx * fs[11] + dy * fs[13] + dz * fs[14]));
    f[1] = fs[1] + dx * (fs[4] + 0.5 * (dx * fs[10] + dy * fs[12] + dz * fs[13])) + dy * (fs[6] + 0.5 * (dx * fs[12] + dy * fs[15] + dz * fs[16])) +
           dz * (fs[7] + 0.5 * (dx * fs[13] + dy * fs[16] + dz * fs[17]));
    f[2] = fs[2] + dx * (fs[5] + 0.5 * (dx * fs[11] + dy * fs[13] + dz * fs[14])) + dy * (fs[7] + 0.5 * (dx * fs[13] + dy * fs[16] + dz * fs[17])) +
           dz * (fs[8] + 0.5 * (dx * fs[14] + dy * fs[17] + dz * fs[18]));
    f[3] = fs[3] + dx * fs[9] + dy * fs[10] + dz * fs[11];
    f[4] = fs[4] + dx * fs[10] + dy * fs[12] + dz * fs[13];
    f[5] = fs[5] + dx * fs[11] + dy * fs[13] + dz * fs[14];
    f[6] = fs[6] + dx * fs[12] + dy * fs[15] + dz * fs[16];
    f[7] = fs[7] + dx * fs[13] + dy * fs[16] + dz * fs[17];
    f[8] = fs[8] + dx * fs[14] + dy * fs[17] + dz * fs[18];
#pragma unroll
    for (int k = 9; k < 19; k++) f[k] = fs[k];
    f[19] = fs[19] + dx * (fs[0] + 0.5 * (dx * fs[3] + dy * fs[4] + dz * fs[5])) + dy * (fs[1] + 0.5 * (dx * fs[4] + dy * fs[6] + dz * fs[7])) +
            dz * (fs[2] + 0.5 * (dx * fs[5] + dy * fs[7] + dz * fs[8]));
}

// expand_fgrav_in_taylor_series (kdtree.F90:1799-1840), including the reference's `dz*dfxy` term of the potential (:1836)
__device__ __forceinline__ void l2p(const double *__restrict__ fn, double dx, double dy, double dz, double &fx, double &fy, double &fz, double &pot)
{
    const double dfxx = fn[3], dfxy = fn[4], dfxz = fn[5], dfyy = fn[6], dfyz = fn[7], dfzz = fn[8];
    const double xxx = fn[9], xxy = fn[10], xxz = fn[11], xyy = fn[12], xyz = fn[13], xzz = fn[14], yyy = fn[15], yyz = fn[16], yzz = fn[17], zzz = fn[18];
    fx = fn[0] + dx * (dfxx + 0.5 * (dx * xxx + dy * xxy + dz * xxz)) + dy * (dfxy + 0.5 * (dx * xxy + dy * xyy + dz * xyz)) +
         dz * (dfxz + 0.5 * (dx * xxz + dy * xyz + dz * xzz));
    fy = fn[1] + dx * (dfxy + 0.5 * (dx * xxy + dy * xyy + dz * xyz)) + dy * (dfyy + 0.5 * (dx * xyy + dy * yyy + dz * yyz)) +
         dz * (dfyz + 0.5 * (dx * xyz + dy * yyz + dz * yzz));
    fz = fn[2] + dx * (dfxz + 0.5 * (dx * xxz + dy * xyz + dz * xzz)) + dy * (dfyz + 0.5 * (dx * xyz + dy * yyz + dz * yzz)) +
         dz * (dfzz + 0.5 * (dx * xzz + dy * yzz + dz * zzz));
    pot = fn[19] - (dx * (fx - 0.5 * (dx * dfxx + dy * dfxy + dz * dfxy)) + dy * (fy - 0.5 * (dx * dfxy + dy * dfyy + dz * dfyz)) +
                    dz * (fz - 0.5 * (dx * dfxz + dy * dfyz + dz * dfzz)));
}

struct WalkArgs {
    const GNode *nodes; int n0, n1;              // nodes of this level
    const int *lst_prev; const long long *lstoff; const int *lstcnt_in;   // lists written by the previous level (indexed by parent)
    int *lst_cur; unsigned long long *lst_ptr; unsigned long long lst_cap;   // this level's list pool
    long long *outoff; int *outcnt;              // per node: offset / count of its own list (internal) in lst_cur
    int *p2p; unsigned long long *p2p_ptr; unsigned long long p2p_cap; long long *p2poff; int *p2pcnt;   // persistent P2P lists of the leaves
    double *fnode;                               // (LENF, nnodes)
    double tree_acc2, radkern; unsigned long long *cnt;
};

// node_interaction (kdtree.F90:1653-1700): 0 = accepted (well separated), 1 = rejected
__device__ __forceinline__ int mac(const GNode &nd, const GNode &ns, double tree_acc2, double radkern, double &dx, double &dy, double &dz, double &r2)
{
    dx = nd.xcen[0] - ns.xcen[0]; dy = nd.xcen[1] - ns.xcen[1]; dz = nd.xcen[2] - ns.xcen[2];
    r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double rcut = fmax(__dmul_rn(nd.hmax, radkern), __dmul_rn(ns.hmax, radkern));
    const double ss = __dadd_rn(nd.size, ns.size);
    const double sr = __dadd_rn(ss, rcut);
    const bool wellsep = (__dmul_rn(tree_acc2, r2) > __dmul_rn(ss, ss)) && (r2 > __dmul_rn(sr, sr));
    return wellsep ? 0 : 1;
}

// one warp per node of the level (see header).  Leaves run the source loop twice: first to count their P2P list, then,
// after reserving exactly that many slots in the persistent pool, to fill it and do the M2L sums.
__global__ void __launch_bounds__(128) k_g_walk(const WalkArgs a)
{
    const int lane = lane_id();
    const int d = a.n0 + (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    if (d >= a.n1) return;
    const GNode nd = a.nodes[d];
    if (!(nd.flags & 2)) { if (lane == 0) { a.outoff[d] = 0; a.outcnt[d] = 0; } return; }     // no particle of this rank below: nobody needs F(d)
    const bool dleaf = nd.left < 0;
    const int *pend; int npend; int rootlist = 0;
    if (nd.parent < 0) { pend = &rootlist; npend = 1; }
    else { pend = a.lst_prev + a.lstoff[nd.parent]; npend = a.lstcnt_in[nd.parent]; }
    // working region in this level's pool: out-list of an internal node (<= 2 entries per rejected source), or the open queue of a leaf
    const unsigned long long cap = dleaf ? (unsigned long long)(2 * npend + 256) : (unsigned long long)(2 * npend);
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(a.lst_ptr, cap);
    off = __shfl_sync(FULLMASK, off, 0);
    if (off + cap > a.lst_cap) { if (lane == 0) atomicMax(&a.cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); return; }
    int *work = a.lst_cur + off;
    double f[LENF];
#pragma unroll
    for (int k = 0; k < LENF; k++) f[k] = 0.;
    int nout = 0, np2p = 0, nm2l = 0;
    int *p2pdst = nullptr;
    const int npass = dleaf ? 2 : 1;
    for (int pass = 0; pass < npass; pass++) {
        const bool final_pass = (pass == npass - 1);
        int qn = 0;                                  // leaf: entries waiting in the open queue
        np2p = 0;
        int idx = 0;
        while (true) {
            // next batch of sources: first the parent's list, then (leaf) whatever the queue holds
            int s = -1;
            if (idx < npend) { if (idx + lane < npend) s = pend[idx + lane]; idx += 32; }
            else if (dleaf && qn > 0) { const int take = min(32, qn); if (lane < take) s = work[qn - 1 - lane]; qn -= take; }
            else break;
            int verdict = -1;                        // -1 nothing, 0 accepted, 1 rejected
            bool sleaf = false; int sl = -1, sr = -1;
            if (s >= 0) {
                const GNode &ns = a.nodes[s];
                sl = ns.left; sr = ns.right; sleaf = sl < 0;
                if (s == d) verdict = 1;             // self interaction: always opened (kdtree.F90:1407-1411)
                else {
                    double dx, dy, dz, r2;
                    verdict = mac(nd, ns, a.tree_acc2, a.radkern, dx, dy, dz, r2);
                    if (verdict == 0 && final_pass) { m2l(dx, dy, dz, rsqrt_pos(r2), ns.mass, ns.quads, f); nm2l++; }
                }
            }
            const unsigned mleaf = __ballot_sync(FULLMASK, verdict == 1 && sleaf);
            const unsigned mint = __ballot_sync(FULLMASK, verdict == 1 && !sleaf);
            const unsigned lt = (1u << lane) - 1u;
            if (dleaf) {                             // open_nodes with isdstleaf (kdtree.F90:1617-1640)
                if (verdict == 1 && sleaf && final_pass) p2pdst[np2p + __popc(mleaf & lt)] = s;
                np2p += __popc(mleaf);
                if ((unsigned long long)(qn + 2 * __popc(mint)) > cap) { if (lane == 0) atomicMax(&a.cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); return; }
                if (verdict == 1 && !sleaf) { const int q = qn + 2 * __popc(mint & lt); work[q] = sl; work[q + 1] = sr; }
                qn += 2 * __popc(mint);
                __syncwarp();
            } else {                                 // the destination descends: leaf sources stay, internal sources are replaced by their children
                const int q = nout + __popc(mleaf & lt) + 2 * __popc(mint & lt);
                if (verdict == 1 && sleaf) work[q] = s;
                if (verdict == 1 && !sleaf) { work[q] = sl; work[q + 1] = sr; }
                nout += __popc(mleaf) + 2 * __popc(mint);
            }
        }
        if (dleaf && pass == 0) {                    // reserve the exact P2P list
            unsigned long long po = 0;
            if (lane == 0) po = atomicAdd(a.p2p_ptr, (unsigned long long)np2p);
            po = __shfl_sync(FULLMASK, po, 0);
            if (po + np2p > a.p2p_cap) { if (lane == 0) atomicMax(&a.cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); return; }
            p2pdst = a.p2p + po;
            if (lane == 0) { a.p2poff[d] = (long long)po; a.p2pcnt[d] = np2p; }
        }
    }
    if (lane == 0) { a.outoff[d] = (long long)off; a.outcnt[d] = nout; }
    // F(d) = sum of accepted M2L + L2L(F(parent)) (kdtree.F90:1427-1457)
#pragma unroll
    for (int k = 0; k < LENF; k++) f[k] = warp_sum(f[k]);
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) nm2l += __shfl_xor_sync(FULLMASK, nm2l, sft);
    if (lane == 0 && nm2l) atomicAdd(&a.cnt[CNT_NM2L], (unsigned long long)nm2l);
    if (lane == 0) {
        double *fd = a.fnode + (size_t)LENF * d;
        if (nd.parent >= 0) {
            const GNode &np = a.nodes[nd.parent];
            double g[LENF];
            l2l(g, a.fnode + (size_t)LENF * nd.parent, nd.xcen[0] - np.xcen[0], nd.xcen[1] - np.xcen[1], nd.xcen[2] - np.xcen[2]);
#pragma unroll
            for (int k = 0; k < LENF; k++) fd[k] = g[k] + f[k];
        } else {
#pragma unroll
            for (int k = 0; k < LENF; k++) fd[k] = f[k];
        }
    }
}

// capacity the next level will ask of its list pool
__global__ void k_g_need(int n0, int n1, const GNode *__restrict__ nodes, const int *__restrict__ outcnt, unsigned long long *need)
{
    const int d = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (d < n1) { const GNode &nd = nodes[d]; if (nd.flags & 2) v = 2ull * (unsigned long long)outcnt[nd.parent] + (nd.left < 0 ? 256ull : 0ull); }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(FULLMASK, v, s);
    if (lane_id() == 0 && v) atomicAdd(need, v);
}

// ---- P2P + L2P: one warp per leaf ---------------------------------------------------------------------------------------------
struct P2PArgs {
    const GNode *nodes; int nn; const int *p2p; const long long *p2poff; const int *p2pcnt; const double4 *pos; const double *mass; const int *gid; const double *fnode;
    double4 *gacc; unsigned long long *cnt; int64_t own_lo, own_hi;
};

#define P2P_BATCH 32                       // source leaves staged per batch
#define P2P_SLOTS (P2P_BATCH * MINPART)    // source particles per batch (upper bound)

// One warp per leaf.  The source leaves of the P2P list are staged batch by batch into shared memory as a dense particle array
// (lane = source leaf copies its <= 10 particles), then lane = source PARTICLE: full lanes regardless of the leaves' fill.
// The <= 10 targets of the leaf live in registers and are broadcast across the lanes; per-target sums are reduced once at the end.
template <int K>
__global__ void __launch_bounds__(128, 4) k_g_p2p(const P2PArgs a)
{
    typedef SphKern<K> KF;
    extern __shared__ double4 p2p_smem[];
    __shared__ double4 tgt_smem[4][MINPART];                                             // targets {x, y, z, 1/h^2}: broadcast reads
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    double4 *spos = p2p_smem + (size_t)wib * P2P_SLOTS;                                  // {x, y, z, 1/h^2}
    double *smass = reinterpret_cast<double *>(p2p_smem + 4 * P2P_SLOTS) + (size_t)wib * P2P_SLOTS;
    double4 *tg = tgt_smem[wib];
    const int d = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    if (d >= a.nn) return;
    const GNode nd = a.nodes[d];
    if (nd.left >= 0 || !(nd.flags & 1)) return;             // leaves with an active particle only (force.F90:509)
    const int nt = nd.count;                                  // <= MINPART targets
    double fx[MINPART], fy[MINPART], fz[MINPART], ph[MINPART];
#pragma unroll
    for (int t = 0; t < MINPART; t++) fx[t] = fy[t] = fz[t] = ph[t] = 0.;
    if (lane < MINPART) {
        const double4 p = a.pos[nd.start + min(lane, nt - 1)];
        const double h1 = 1. / fabs(p.w);
        tg[lane] = make_double4(p.x, p.y, p.z, h1 * h1);       // same expression as k_force_prep: the SPH-pair test below must be the one k_force uses
    }
    __syncwarp();
    const int *lst = a.p2p + a.p2poff[d];
    const int nl = a.p2pcnt[d];
    unsigned long long npairs = 0;
    // particle range of this lane's source leaf in the NEXT batch: list entry -> node -> range are two of the three dependent reads of
    // the staging, fetched while the current batch is being summed
    int nstart = 0, ncount = 0;
    if (lane < nl) { const GNode &ns = a.nodes[lst[lane]]; nstart = ns.start; ncount = ns.count; }
    for (int base = 0; base < nl; base += P2P_BATCH) {
        // ---- stage: lane = source leaf
        const int sstart = nstart, scount = ncount;
        nstart = 0; ncount = 0;
        if (base + P2P_BATCH + lane < nl) { const GNode &ns = a.nodes[lst[base + P2P_BATCH + lane]]; nstart = ns.start; ncount = ns.count; }
        int off = scount;                                     // inclusive warp scan of the counts
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const int o = __shfl_up_sync(FULLMASK, off, sft); if (lane >= sft) off += o; }
        const int total = __shfl_sync(FULLMASK, off, 31);
        off -= scount;
        for (int k = 0; k < scount; k++) {
            const double4 pj = a.pos[sstart + k];
            const double hj1 = 1. / fabs(pj.w);
            spos[off + k] = make_double4(pj.x, pj.y, pj.z, hj1 * hj1);
            smass[off + k] = a.mass[sstart + k];
        }
        // where (if anywhere) this leaf's own particles sit in the batch: source slot selfoff + t is target t itself
        const bool selfleaf = (sstart == nd.start) && scount > 0;
        const unsigned selfmask = __ballot_sync(FULLMASK, selfleaf);
        const int selfoff = selfmask ? __shfl_sync(FULLMASK, off, __ffs(selfmask) - 1) : -(1 << 20);
        __syncwarp();
        // ---- pairs: lane = source particle
        for (int j = lane; j < total; j += 32) {
            const double4 pj = spos[j];
            const double mj = smass[j];
            const int tself = j - selfoff;                    // index of the target this source is (out of range if none)
#pragma unroll
            for (int t = 0; t < MINPART; t++) {
                if (t < nt) {
                    const double4 ti = tg[t];
                    const double dx = ti.x - pj.x, dy = ti.y - pj.y, dz = ti.z - pj.z;
                    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    const double q2i = __dmul_rn(r2, ti.w), q2j = __dmul_rn(r2, pj.w);
                    const bool sph = (q2i < KF::radkern2) || (q2j < KF::radkern2);       // force.F90:1287: handled by k_force
                    const bool use = !sph && (t != tself);
                    const double rinv = use ? rsqrt_pos(r2) : 0.;
                    const double mr3 = mj * rinv * rinv * rinv;                            // force.F90:2020-2043
                    fx[t] -= dx * mr3; fy[t] -= dy * mr3; fz[t] -= dz * mr3;
                    ph[t] -= mj * rinv;
                    npairs += use ? 1 : 0;
                }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int t = 0; t < MINPART; t++) { fx[t] = warp_sum(fx[t]); fy[t] = warp_sum(fy[t]); fz[t] = warp_sum(fz[t]); ph[t] = warp_sum(ph[t]); }
    // L2P at every particle of the leaf (force.F90:2909-2927), lane = target
    double ox = 0., oy = 0., oz = 0., op = 0., px = 0., py = 0., pz = 0.;
#pragma unroll
    for (int t = 0; t < MINPART; t++) if (lane == t) { ox = fx[t]; oy = fy[t]; oz = fz[t]; op = ph[t]; }
    if (lane < nt) { const double4 ti = tg[lane]; px = ti.x; py = ti.y; pz = ti.z; }
    if (lane < nt) {
        double gx, gy, gz, gp;
        l2p(a.fnode + (size_t)LENF * d, px - nd.xcen[0], py - nd.xcen[1], pz - nd.xcen[2], gx, gy, gz, gp);
        const int64_t gi = a.gid[nd.start + lane];
        if (gi >= a.own_lo && gi < a.own_hi) a.gacc[gi - a.own_lo] = make_double4(ox + gx, oy + gy, oz + gz, op + gp);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) npairs += __shfl_xor_sync(FULLMASK, npairs, s);
    if (lane == 0) atomicAdd(&a.cnt[CNT_NGRAVPAIRS], npairs);
}

}  // namespace

static inline int nblk(int64_t n, int b) { return (int)((n + b - 1) / b); }
#define GL(c, kern, grid, block, ...) do { kern<<<(grid), (block), 0, (c)->stream>>>(__VA_ARGS__); (c)->launches++; } while (0)

struct GravState {
    DevBuf<GNode> nodes; DevBuf<GBuild> gb;
    DevBuf<double4> pos[2]; DevBuf<double> mass[2]; DevBuf<int> gid[2], pnode[2];
    DevBuf<int> flag, scan, lst[2], outcnt, p2p, p2pcnt;
    DevBuf<long long> outoff, p2poff;
    DevBuf<double> fnode;
    DevBuf<unsigned long long> ptrs;          // [0] list pool pointer, [1] P2P pool pointer, [2] need, [3] nnodes (as int)
    DevBuf<char> cubtmp;
    std::vector<int> level_start;              // nodes of level L are [level_start[L], level_start[L+1])
    DevBuf<int> lvl;                           // the same on the device while the tree is being built (grav_build)
    int levels_prev = 0;                       // depth of the previous build: from where on the host watches for the last level
    int nn = 0, cur = 0, nlive = 0;
    // multi-GPU: the gathered particle set of all ranks (positions, h history) in rank order; this rank owns [own_lo, own_hi)
    int64_t nglobal = 0, own_lo = 0;
    DevBuf<double> gsend, grecv, g_xyzh, g_hbuild, g_hhist; DevBuf<int> g_hits; DevBuf<int8_t> g_iphase; DevBuf<long long> g_counts;
    void release()
    {
        nodes.release(); gb.release(); for (int k = 0; k < 2; k++) { pos[k].release(); mass[k].release(); gid[k].release(); pnode[k].release(); lst[k].release(); }
        flag.release(); scan.release(); outoff.release(); outcnt.release(); p2p.release(); p2poff.release(); p2pcnt.release(); fnode.release(); ptrs.release();
        cubtmp.release(); gsend.release(); grecv.release(); g_xyzh.release(); g_hbuild.release(); g_hhist.release(); g_hits.release(); g_iphase.release(); g_counts.release();
    }
};

void gravity_release(sphgpu_ctx *c)
{
    if (c->grav) { c->grav->release(); delete c->grav; c->grav = nullptr; }
}

// level-synchronous construction of the reference-topology tree (positions and masses only: valid until the next build_tree)
struct GravInput { int64_t n; const double *xyzh; const int8_t *iphase; const double *hbuild; const int *hits; const double *hhist; int64_t own_lo, own_hi; };

static GravInput grav_input(sphgpu_ctx *c, GravState &g)
{
    GravInput in;
    if (g.nglobal > 0) { in.n = g.nglobal; in.xyzh = g.g_xyzh.p; in.iphase = g.g_iphase.p; in.hbuild = g.g_hbuild.p; in.hits = g.g_hits.p; in.hhist = g.g_hhist.p; in.own_lo = g.own_lo; in.own_hi = g.own_lo + c->nlocal; }
    else { in.n = c->npart; in.xyzh = c->xyzh.p; in.iphase = c->iphase.p; in.hbuild = c->h_build.p; in.hits = c->h_its.p; in.hhist = c->h_hist.p; in.own_lo = 0; in.own_hi = c->npart; }
    return in;
}

static int grav_build(sphgpu_ctx *c, GravState &g, const GravInput &in)
{
    const int64_t n = in.n;
    const sphgpu_params &p = c->hp.p;
    cudaStream_t st = c->stream;
    CUDA_TRY(c, g.flag.ensure(n)); CUDA_TRY(c, g.scan.ensure(n)); CUDA_TRY(c, g.ptrs.ensure(8));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, g.flag.p, g.scan.p, (int)n, st);
    CUDA_TRY(c, g.cubtmp.ensure(tb + 256));
    // live particles keep the caller's order (construct_root_node)
    GL(c, k_g_liveflag, nblk(n, 256), 256, n, in.xyzh, g.flag.p);
    size_t tbb = g.cubtmp.cap;
    CUDA_TRY(c, cub::DeviceScan::ExclusiveSum(g.cubtmp.p, tbb, g.flag.p, g.scan.p, (int)n, st));
    c->launches++;
    int lastf = 0, lasts = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&lastf, g.flag.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(&lasts, g.scan.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    const int nlive = lastf + lasts;
    if (nlive <= 0) { c->err = "gravity tree: no live particles"; return SPHGPU_ERR_NOPART; }
    g.nlive = nlive;
    for (int k = 0; k < 2; k++) { CUDA_TRY(c, g.pos[k].ensure(nlive)); CUDA_TRY(c, g.mass[k].ensure(nlive)); CUDA_TRY(c, g.gid[k].ensure(nlive)); CUDA_TRY(c, g.pnode[k].ensure(nlive)); }
    GL(c, k_g_init, nblk(n, 256), 256, n, in.xyzh, in.iphase, c->hp, g.pos[0].p, g.mass[0].p, g.gid[0].p, g.pnode[0].p, g.scan.p);
    const int maxnodes = nlive + 1024;          // leaves hold > 1 particle except in degenerate splits; overflow is reported
    CUDA_TRY(c, g.nodes.ensure(maxnodes)); CUDA_TRY(c, g.gb.ensure(maxnodes));
    // dfac = 1/massoftype(igas), or the first massive type when there is no gas (kdtree.F90:612-624)
    double pm = p.massoftype[IGAS];
    if (!(pm > 0.)) { pm = 0.; for (int t = 2; t < SPHGPU_MAXTYPES; t++) if (p.massoftype[t] > 0.) { pm = p.massoftype[t]; break; } }
    const double dfac = pm > 0. ? 1. / pm : 1.;
    GNode root; memset(&root, 0, sizeof root);
    root.parent = -1; root.left = root.right = -1; root.start = 0; root.count = nlive; root.level = 0;
    CUDA_TRY(c, cudaMemcpyAsync(g.nodes.p, &root, sizeof root, cudaMemcpyHostToDevice, st));
    int *nnodes_d = reinterpret_cast<int *>(g.ptrs.p + 3);
    int one = 1;
    CUDA_TRY(c, cudaMemcpyAsync(nnodes_d, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    // level table on the device: lvl[L] = first node of level L.  The host reads it (and the error flag) only from the depth of the
    // previous build on: a tree of the same particles a step later is as deep, give or take a level, and the levels above that are
    // enqueued without a round trip each.  Levels enqueued beyond the last one find an empty node range and sweep finished leaves.
    CUDA_TRY(c, g.lvl.ensure(132));
    { const int init[2] = {0, 1}; CUDA_TRY(c, cudaMemcpyAsync(g.lvl.p, init, sizeof init, cudaMemcpyHostToDevice, st)); }
    g.level_start.clear();
    const int first_check = g.levels_prev > 3 ? g.levels_prev - 2 : 0;
    int cur = 0, nlevels = -1;
    std::vector<int> hl(132, 0);
    for (int level = 0; level < 128; level++) {
        // most nodes a level can hold: 2^level, and no more than there are nodes at all
        const long long ub = level < 30 ? std::min<long long>(1ll << level, maxnodes) : maxnodes;
        GL(c, k_g_node_reset, nblk(ub, 128), 128, g.lvl.p, level, g.gb.p);
        GL(c, k_g_sums, nblk(nlive, 256), 256, nlive, g.pnode[cur].p, g.pos[cur].p, g.mass[cur].p, dfac, g.gb.p);
        GL(c, k_g_nodes_a, nblk(ub, 128), 128, g.lvl.p, level, g.nodes.p, g.gb.p, dfac, c->counters.p);
        if (p.gravity) GL(c, k_g_moments<true>, nblk(nlive, 256), 256, nlive, g.pnode[cur].p, g.pos[cur].p, g.mass[cur].p, g.nodes.p, g.gb.p, g.flag.p);
        else GL(c, k_g_moments<false>, nblk(nlive, 256), 256, nlive, g.pnode[cur].p, g.pos[cur].p, g.mass[cur].p, g.nodes.p, g.gb.p, g.flag.p);
        tbb = g.cubtmp.cap;
        CUDA_TRY(c, cub::DeviceScan::ExclusiveSum(g.cubtmp.p, tbb, g.flag.p, g.scan.p, nlive, st));
        c->launches++;
        GL(c, k_g_nodes_b, nblk(ub, 128), 128, g.lvl.p, level, g.nodes.p, g.gb.p, g.flag.p, g.scan.p, nnodes_d, maxnodes, c->counters.p);
        GL(c, k_g_scatter, nblk(nlive, 256), 256, nlive, g.pnode[cur].p, g.pos[cur].p, g.mass[cur].p, g.gid[cur].p, g.nodes.p, g.gb.p, g.flag.p, g.scan.p,
           g.pnode[1 - cur].p, g.pos[1 - cur].p, g.mass[1 - cur].p, g.gid[1 - cur].p);
        GL(c, k_g_level_end, 1, 1, nnodes_d, g.lvl.p, level);
        cur = 1 - cur;
        if (level < first_check) continue;
        unsigned long long err = 0;
        CUDA_TRY(c, cudaMemcpyAsync(hl.data(), g.lvl.p, sizeof(int) * (level + 3), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaMemcpyAsync(&err, c->counters.p + CNT_ERR, sizeof err, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
        if (err) { c->err = err == SPHGPU_ERR_OVERFLOW ? "gravity tree: number of nodes exceeds array dimensions" : "gravity tree: totmass_node==0"; return (int)err; }
        for (int L = 0; L <= level; L++) if (hl[L + 2] == hl[L + 1]) { nlevels = L + 1; break; }      // level L split nothing: it is the last one
        if (nlevels > 0) break;
    }
    if (nlevels <= 0) { c->err = "gravity tree: more than 128 levels"; return SPHGPU_ERR_OVERFLOW; }
    for (int L = 0; L <= nlevels; L++) g.level_start.push_back(hl[L]);
    g.levels_prev = nlevels;
    const int n1 = hl[nlevels];
    g.nn = n1; g.cur = cur;
    return SPHGPU_OK;
}

int gravity_run(sphgpu_ctx *c)
{
    const sphgpu_params &p = c->hp.p;
    if (p.periodic) { c->err = "gravity: self-gravity with periodic boundaries is not supported (as in the reference)"; return SPHGPU_ERR_ARG; }
    if (!c->grav) c->grav = new GravState();
    GravState &g = *c->grav;
    cudaStream_t st = c->stream;
    const GravInput in = grav_input(c, g);
    const int64_t n = in.n;
    cudaEventRecord(c->ev[12], st);
    CUDA_TRY(c, cudaMemsetAsync(c->counters.p + CNT_ERR, 0, 2 * sizeof(unsigned long long), st));
    CUDA_TRY(c, cudaMemsetAsync(c->counters.p + CNT_NGRAVPAIRS, 0, 2 * sizeof(unsigned long long), st));
    if (!c->grav_tree_valid) { TRY(grav_build(c, g, in)); c->grav_tree_valid = true; }
    const int nn = g.nn, cur = g.cur, nlive = g.nlive;
    const int nlev = (int)g.level_start.size() - 1;
    // node hmax as the reference's tree holds it at force time
    GL(c, k_g_hmax_leaf, nblk(nn, 128), 128, nn, g.nodes.p, g.gb.p, g.gid[cur].p, in.hbuild, in.hits, in.hhist, n, SPHGPU_HHIST, in.iphase,
       p.ind_timesteps, in.own_lo, in.own_hi);
    for (int L = nlev - 1; L >= 0; L--) {
        const int a0 = g.level_start[L], a1 = g.level_start[L + 1];
        GL(c, k_g_hmax_up, nblk(a1 - a0, 128), 128, a0, a1, g.nodes.p, g.gb.p);
    }
    // breadth-first dual walk
    CUDA_TRY(c, g.outoff.ensure(nn)); CUDA_TRY(c, g.outcnt.ensure(nn)); CUDA_TRY(c, g.p2poff.ensure(nn)); CUDA_TRY(c, g.p2pcnt.ensure(nn));
    CUDA_TRY(c, g.fnode.ensure((size_t)LENF * nn));
    const unsigned long long p2pcap = (unsigned long long)c->grav_p2p_per_particle * (unsigned long long)nlive + 4096ull;
    CUDA_TRY(c, g.p2p.ensure(p2pcap));
    CUDA_TRY(c, cudaMemsetAsync(g.ptrs.p, 0, 3 * sizeof(unsigned long long), st));
    CUDA_TRY(c, cudaMemsetAsync(g.p2pcnt.p, 0, sizeof(int) * (size_t)nn, st));
    unsigned long long need = 512;
    for (int L = 0; L < nlev; L++) {
        const int a0 = g.level_start[L], a1 = g.level_start[L + 1];
        DevBuf<int> &pool = g.lst[L & 1];
        CUDA_TRY(c, pool.ensure(need + 64));
        CUDA_TRY(c, cudaMemsetAsync(g.ptrs.p, 0, sizeof(unsigned long long), st));
        WalkArgs w;
        w.nodes = g.nodes.p; w.n0 = a0; w.n1 = a1; w.lst_prev = g.lst[(L + 1) & 1].p; w.lstoff = g.outoff.p; w.lstcnt_in = g.outcnt.p;
        w.lst_cur = pool.p; w.lst_ptr = g.ptrs.p; w.lst_cap = pool.cap; w.outoff = g.outoff.p; w.outcnt = g.outcnt.p;
        w.p2p = g.p2p.p; w.p2p_ptr = g.ptrs.p + 1; w.p2p_cap = g.p2p.cap; w.p2poff = g.p2poff.p; w.p2pcnt = g.p2pcnt.p;
        w.fnode = g.fnode.p; w.tree_acc2 = p.tree_accuracy * p.tree_accuracy; w.radkern = c->hp.kc.radkern; w.cnt = c->counters.p;
        GL(c, k_g_walk, nblk((int64_t)(a1 - a0) * 32, 128), 128, w);
        if (L + 1 < nlev) {
            CUDA_TRY(c, cudaMemsetAsync(g.ptrs.p + 2, 0, sizeof(unsigned long long), st));
            GL(c, k_g_need, nblk(g.level_start[L + 2] - a1, 128), 128, a1, g.level_start[L + 2], g.nodes.p, g.outcnt.p, g.ptrs.p + 2);
        }
        unsigned long long hp[3], err = 0;
        CUDA_TRY(c, cudaMemcpyAsync(hp, g.ptrs.p, sizeof hp, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaMemcpyAsync(&err, c->counters.p + CNT_ERR, sizeof err, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
        CUDA_TRY(c, cudaGetLastError());
        if (err) { c->err = "gravity: interaction list pool overflow (raise option grav_p2p_per_particle)"; return SPHGPU_ERR_OVERFLOW; }
        need = hp[2] + 512;
    }
    CUDA_TRY(c, c->gacc.ensure(c->npart));
    P2PArgs a;
    a.nodes = g.nodes.p; a.nn = nn; a.p2p = g.p2p.p; a.p2poff = g.p2poff.p; a.p2pcnt = g.p2pcnt.p; a.pos = g.pos[cur].p; a.mass = g.mass[cur].p;
    a.gid = g.gid[cur].p; a.fnode = g.fnode.p; a.gacc = c->gacc.p; a.cnt = c->counters.p; a.own_lo = in.own_lo; a.own_hi = in.own_hi;
    cudaEventRecord(c->ev[13], st);
    const size_t p2psmem = (size_t)4 * P2P_SLOTS * (sizeof(double4) + sizeof(double));
    if (p.kernel == 0) {
        CUDA_TRY(c, cudaFuncSetAttribute(k_g_p2p<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2psmem));
        k_g_p2p<0><<<nblk((int64_t)nn * 32, 128), 128, p2psmem, st>>>(a);
    } else {
        CUDA_TRY(c, cudaFuncSetAttribute(k_g_p2p<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2psmem));
        k_g_p2p<1><<<nblk((int64_t)nn * 32, 128), 128, p2psmem, st>>>(a);
    }
    c->launches++;
    cudaEventRecord(c->ev[14], st);
    unsigned long long hg[2];
    CUDA_TRY(c, cudaMemcpyAsync(hg, c->counters.p + CNT_NGRAVPAIRS, sizeof hg, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    CUDA_TRY(c, cudaGetLastError());
    c->npairs_gravity = (int64_t)hg[0]; c->nm2l = (int64_t)hg[1];
    { float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[12], c->ev[14]); c->ms_gravity[0] = ms; cudaEventElapsedTime(&ms, c->ev[13], c->ev[14]); c->ms_gravity[1] = ms; }
    return SPHGPU_OK;
}

// ---- reference-compatible neighbour mode (common.cuh: refcompat): the reference's tree with its node hmax at force time, compact ----
__global__ void k_ref_export(int nn, const GNode *__restrict__ nodes, const int *__restrict__ gid, RefNode *__restrict__ out, int *__restrict__ leaf_of)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= nn) return;
    const GNode &nd = nodes[d];
    RefNode r;
    r.xcen[0] = nd.xcen[0]; r.xcen[1] = nd.xcen[1]; r.xcen[2] = nd.xcen[2]; r.size = nd.size; r.hmax = nd.hmax; r.parent = nd.parent; r.pad = 0;
    out[d] = r;
    if (nd.left < 0) for (int s = nd.start; s < nd.start + nd.count; s++) leaf_of[gid[s]] = d;
}
// leaf of every sorted slot; stored as ~leaf when the leaf's hmax covers the particle's own h: every node on its path to the root then
// has hmax >= h_j (hmax only grows upwards), so any i inside j's kernel opens them all (|x_i - c_i| <= size_i, |x_j - c_n| <= size_n,
// r_ij < radkern h_j <= radkern hmax_n) and ref_walk_reaches need not walk.  Only the particles a leaf's hmax does NOT cover -- inactive
// members above 1.01 x the active members' h, members that converged before the others -- can be missed by the reference.
__global__ void k_ref_leaf_sorted(int64_t nlive, const int *__restrict__ perm, const int *__restrict__ leaf_of, const RefNode *__restrict__ nodes,
                                  const double4 *__restrict__ pos4, int *__restrict__ out)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    const int leaf = leaf_of[perm[s]];
    out[s] = (fabs(pos4[s].w) <= nodes[leaf].hmax) ? ~leaf : leaf;
}

int refcompat_prepare(sphgpu_ctx *c)
{
    if (c->ref_valid) return SPHGPU_OK;
    if (!c->grav) c->grav = new GravState();
    GravState &g = *c->grav;
    const GravInput in = grav_input(c, g);
    if (g.nglobal > 0) { c->err = "refcompat: not available together with the gathered multi-GPU gravity set"; return SPHGPU_ERR_STATE; }
    CUDA_TRY(c, cudaMemsetAsync(c->counters.p + CNT_ERR, 0, 2 * sizeof(unsigned long long), c->stream));
    if (!c->grav_tree_valid) { TRY(grav_build(c, g, in)); c->grav_tree_valid = true; }
    const int nn = g.nn, cur = g.cur;
    const int nlev = (int)g.level_start.size() - 1;
    GL(c, k_g_hmax_leaf, nblk(nn, 128), 128, nn, g.nodes.p, g.gb.p, g.gid[cur].p, in.hbuild, in.hits, in.hhist, in.n, SPHGPU_HHIST, in.iphase,
       c->hp.p.ind_timesteps, in.own_lo, in.own_hi);
    for (int L = nlev - 1; L >= 0; L--) {
        const int a0 = g.level_start[L], a1 = g.level_start[L + 1];
        GL(c, k_g_hmax_up, nblk(a1 - a0, 128), 128, a0, a1, g.nodes.p, g.gb.p);
    }
    CUDA_TRY(c, c->ref_nodes.ensure(nn)); CUDA_TRY(c, c->ref_leaf.ensure(c->npart)); CUDA_TRY(c, c->ref_leaf_sorted.ensure(c->npart));
    CUDA_TRY(c, cudaMemsetAsync(c->ref_leaf.p, 0, sizeof(int) * (size_t)c->npart, c->stream));
    GL(c, k_ref_export, nblk(nn, 128), 128, nn, g.nodes.p, g.gid[cur].p, c->ref_nodes.p, c->ref_leaf.p);
    GL(c, k_ref_leaf_sorted, nblk(c->nlive, 256), 256, c->nlive, c->perm.p, c->ref_leaf.p, c->ref_nodes.p, c->pos4.p, c->ref_leaf_sorted.p);
    CUDA_TRY(c, cudaGetLastError());
    c->ref_valid = true;
    return SPHGPU_OK;
}

// ---- multi-GPU: every rank evaluates the FMM for its own particles on the tree of the WHOLE particle set ---------------------------
// (replaces the global tree + remote cell export of maketreeglobal / mpi_force for the gravity terms, kdtree.F90:2044-2300).
// The ranks all-gather 13 doubles per owned particle {x,y,z,h, iphase, h at build_tree, iterations, h history(6)}; each rank then
// builds the same tree, so the result is the single-GPU result to round-off.  Walk, M2L and P2P are restricted to nodes that hold
// particles of this rank; only the tree construction is replicated.
#define GREC (7 + SPHGPU_HHIST)

__global__ void k_gg_pack(int64_t nlocal, int64_t npart, const double *__restrict__ xyzh, const int8_t *__restrict__ iphase, const double *__restrict__ hbuild,
                          const int *__restrict__ hits, const double *__restrict__ hhist, double *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nlocal) return;
    double *o = out + (size_t)GREC * i;
    for (int k = 0; k < 4; k++) o[k] = xyzh[4 * i + k];
    o[4] = (double)iphase[i]; o[5] = hbuild[i]; o[6] = (double)hits[i];
    for (int k = 0; k < SPHGPU_HHIST; k++) o[7 + k] = hhist[(size_t)k * npart + i];
}

// recv = nranks blocks of `stride` records (padded); counts/offs per rank
__global__ void k_gg_unpack(int nranks, int64_t stride, int64_t nglobal, const long long *__restrict__ counts, const long long *__restrict__ offs,
                            const double *__restrict__ in, double *__restrict__ xyzh, int8_t *__restrict__ iphase, double *__restrict__ hbuild,
                            int *__restrict__ hits, double *__restrict__ hhist)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)nranks * stride) return;
    const int r = (int)(t / stride); const int64_t k = t - (int64_t)r * stride;
    if (k >= counts[r]) return;
    const int64_t gi = offs[r] + k;
    const double *o = in + (size_t)GREC * t;
    for (int q = 0; q < 4; q++) xyzh[4 * gi + q] = o[q];
    iphase[gi] = (int8_t)o[4]; hbuild[gi] = o[5]; hits[gi] = (int)o[6];
    for (int q = 0; q < SPHGPU_HHIST; q++) hhist[(size_t)q * nglobal + gi] = o[7 + q];
}

int gravity_gather_pack(sphgpu_ctx *c, void **sendptr, int *record_doubles)
{
    if (!c->grav) c->grav = new GravState();
    GravState &g = *c->grav;
    const int64_t nl = c->nlocal;
    CUDA_TRY(c, g.gsend.ensure((size_t)GREC * nl));
    GL(c, k_gg_pack, nblk(nl, 256), 256, nl, c->npart, c->xyzh.p, c->iphase.p, c->h_build.p, c->h_its.p, c->h_hist.p, g.gsend.p);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *sendptr = g.gsend.p; *record_doubles = GREC;
    return SPHGPU_OK;
}

int gravity_gather_recvbuf(sphgpu_ctx *c, int nranks, int64_t stride, void **recvptr)
{
    if (!c->grav) c->grav = new GravState();
    GravState &g = *c->grav;
    CUDA_TRY(c, g.grecv.ensure((size_t)GREC * nranks * stride));
    *recvptr = g.grecv.p;
    return SPHGPU_OK;
}

int gravity_gather_unpack(sphgpu_ctx *c, int nranks, int myrank, int64_t stride, const int64_t *counts)
{
    if (!c->grav) return SPHGPU_ERR_STATE;
    GravState &g = *c->grav;
    std::vector<long long> hc(nranks), ho(nranks);
    long long tot = 0;
    for (int r = 0; r < nranks; r++) { hc[r] = counts[r]; ho[r] = tot; tot += counts[r]; }
    if (counts[myrank] != c->nlocal) { c->err = "gravity gather: counts[myrank] differs from the number of owned particles"; return SPHGPU_ERR_ARG; }
    CUDA_TRY(c, g.g_counts.ensure(2 * nranks));
    CUDA_TRY(c, cudaMemcpyAsync(g.g_counts.p, hc.data(), sizeof(long long) * nranks, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(g.g_counts.p + nranks, ho.data(), sizeof(long long) * nranks, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, g.g_xyzh.ensure(4 * (size_t)tot)); CUDA_TRY(c, g.g_iphase.ensure(tot)); CUDA_TRY(c, g.g_hbuild.ensure(tot)); CUDA_TRY(c, g.g_hits.ensure(tot));
    CUDA_TRY(c, g.g_hhist.ensure((size_t)SPHGPU_HHIST * tot));
    GL(c, k_gg_unpack, nblk((int64_t)nranks * stride, 256), 256, nranks, stride, (int64_t)tot, g.g_counts.p, g.g_counts.p + nranks, g.grecv.p, g.g_xyzh.p,
       g.g_iphase.p, g.g_hbuild.p, g.g_hits.p, g.g_hhist.p);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    g.nglobal = tot; g.own_lo = ho[myrank];
    c->grav_tree_valid = false;
    return SPHGPU_OK;
}

// node records of the gravity tree, for the parity tests against the reference construction:
// rec = {xcen[3], size, hmax, mass, quads[6]} ; irec = {left, right, parent, start, count, level} ; ids = particle ids (1-based) by slot
int64_t gravity_tree_dump(sphgpu_ctx *c, int64_t maxnodes, double *rec12, int32_t *irec6, int32_t *ids)
{
    if (!c->grav || !c->grav_tree_valid) return -1;
    GravState &g = *c->grav;
    if (!rec12) return g.nn;
    std::vector<GNode> h((size_t)g.nn);
    if (cudaMemcpy(h.data(), g.nodes.p, sizeof(GNode) * (size_t)g.nn, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    const int64_t m = g.nn < maxnodes ? g.nn : maxnodes;
    for (int64_t k = 0; k < m; k++) {
        const GNode &nd = h[(size_t)k];
        double *r = rec12 + 12 * k;
        r[0] = nd.xcen[0]; r[1] = nd.xcen[1]; r[2] = nd.xcen[2]; r[3] = nd.size; r[4] = nd.hmax; r[5] = nd.mass;
        for (int q = 0; q < 6; q++) r[6 + q] = nd.quads[q];
        int32_t *ir = irec6 + 6 * k;
        ir[0] = nd.left; ir[1] = nd.right; ir[2] = nd.parent; ir[3] = nd.start; ir[4] = nd.count; ir[5] = nd.level;
    }
    if (ids) {
        std::vector<int> hid((size_t)g.nlive);
        if (cudaMemcpy(hid.data(), g.gid[g.cur].p, sizeof(int) * (size_t)g.nlive, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        for (int k = 0; k < g.nlive; k++) ids[k] = hid[(size_t)k] + 1;
    }
    return g.nn;
}
