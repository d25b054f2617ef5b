// tree.cu -- GPU spatial tree for neighbour finding: replaces maketree (src/main/kdtree.F90:117-314).
//
// B200-first design, not the reference's top-down COM-split kd-tree:
//   1. 63-bit Morton keys (21 bits/axis, cubic normalisation) + cub radix sort      [HBM-bound, ~24 B/particle/pass]
//   2. leaf cells = maximal octree nodes holding <= max_cell particles, found from the common-prefix
//      lengths of adjacent sorted keys inside a +-max_cell window (no pointer structure needed)
//   3. binary radix tree (Karras 2012) over the cell keys, one thread per internal node
//   4. bottom-up refit of child boxes + hmax with one atomic flag per node
// The tree shape differs from the reference's; what must agree (and is tested) are the neighbour *sets*.
// Periodic wrap and the NaN check of construct_root_node (kdtree.F90:385-401) are done in k_wrap_count.
#include "common.cuh"
#include <cub/cub.cuh>
#include <float.h>

namespace {

__device__ __forceinline__ unsigned long long enc_ordered(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dec_ordered(unsigned long long b)
{
    unsigned long long r = (b & 0x8000000000000000ull) ? (b & 0x7fffffffffffffffull) : ~b;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)r);
#else
    double d; memcpy(&d, &r, 8); return d;
#endif
}

// periodic wrap (boundary.f90:123-157), NaN check (kdtree.F90:391), live count, bounding box
// (+ h as the tree is built with: start of the node-hmax history of the gravity tree and restore point of a halo-widening retry)
__global__ void k_wrap_count(int64_t n, double *__restrict__ xyzh, DevParams dp, unsigned long long *cnt, unsigned long long *bbox_enc,
                             double *__restrict__ h_build, int *__restrict__ h_its)
{
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    int nlive = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double4 x = reinterpret_cast<double4 *>(xyzh)[i];
        h_build[i] = x.w; h_its[i] = 0;
        if (!(x.w < DBL_MIN)) {
            if (dp.p.periodic) {
                bool ch = false;
                if (x.x < dp.p.xmin) { x.x += dp.dxbound; ch = true; } else if (x.x > dp.p.xmax) { x.x -= dp.dxbound; ch = true; }
                if (x.y < dp.p.ymin) { x.y += dp.dybound; ch = true; } else if (x.y > dp.p.ymax) { x.y -= dp.dybound; ch = true; }
                if (x.z < dp.p.zmin) { x.z += dp.dzbound; ch = true; } else if (x.z > dp.p.zmax) { x.z -= dp.dzbound; ch = true; }
                if (ch) reinterpret_cast<double4 *>(xyzh)[i] = x;
            }
            if (isnan(x.x) || isnan(x.y) || isnan(x.z)) { atomicMax(&cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_NAN); atomicMax(&cnt[CNT_ERRID], (unsigned long long)(i + 1)); }
            nlive++;
            lo[0] = fmin(lo[0], x.x); lo[1] = fmin(lo[1], x.y); lo[2] = fmin(lo[2], x.z);
            hi[0] = fmax(hi[0], x.x); hi[1] = fmax(hi[1], x.y); hi[2] = fmax(hi[2], x.z);
        }
    }
    for (int k = 0; k < 3; k++) { lo[k] = warp_min(lo[k]); hi[k] = warp_max(hi[k]); }
    int tot = nlive;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) tot += __shfl_xor_sync(FULLMASK, tot, s);
    if (lane_id() == 0) {
        if (tot) atomicAdd(&cnt[CNT_NLIVE], (unsigned long long)tot);
        if (tot) for (int k = 0; k < 3; k++) { atomicMin(&bbox_enc[k], enc_ordered(lo[k])); atomicMax(&bbox_enc[3 + k], enc_ordered(hi[k])); }
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// 3-D Hilbert index of 21-bit coordinates (Skilling 2004, "Programming the Hilbert curve": axes -> transposed index).  Like Morton
// keys, every key prefix is an octree (binary-radix) node, so the cell and tree construction below is unchanged; unlike Morton order,
// consecutive keys are always spatial neighbours, which keeps the candidate sets of consecutively processed target groups overlapping
// (L1/L2 reuse of the neighbour records).
__device__ __forceinline__ void hilbert_transpose(unsigned &x0, unsigned &x1, unsigned &x2)
{
    const unsigned M = 1u << 20;
    for (unsigned Q = M; Q > 1; Q >>= 1) {
        const unsigned P = Q - 1;
        if (x0 & Q) x0 ^= P;                                   // i = 0: invert
        if (x1 & Q) x0 ^= P; else { const unsigned t = (x0 ^ x1) & P; x0 ^= t; x1 ^= t; }
        if (x2 & Q) x0 ^= P; else { const unsigned t = (x0 ^ x2) & P; x0 ^= t; x2 ^= t; }
    }
    x1 ^= x0; x2 ^= x1;                                        // Gray encode
    unsigned t = 0;
    for (unsigned Q = M; Q > 1; Q >>= 1) if (x2 & Q) t ^= Q - 1;
    x0 ^= t; x1 ^= t; x2 ^= t;
}

// key normalisation computed per thread from the device bounding box (no host round trip): the periodic box when there is one
// (stable across steps), else the particles' bounding box; cubic, slightly enlarged
__device__ __forceinline__ void key_frame(const unsigned long long *bbox_enc, const DevParams &dp, double &x0, double &y0, double &z0, double &inv_scale)
{
    double lo[3], hi[3];
    for (int k = 0; k < 3; k++) { lo[k] = dec_ordered(bbox_enc[k]); hi[k] = dec_ordered(bbox_enc[3 + k]); }
    if (dp.p.periodic) {
        lo[0] = fmin(lo[0], dp.p.xmin); lo[1] = fmin(lo[1], dp.p.ymin); lo[2] = fmin(lo[2], dp.p.zmin);
        hi[0] = fmax(hi[0], dp.p.xmax); hi[1] = fmax(hi[1], dp.p.ymax); hi[2] = fmax(hi[2], dp.p.zmax);
    }
    double scale = fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    if (!(scale > 0.)) scale = 1.0;
    scale *= 1.0000001;
    x0 = lo[0]; y0 = lo[1]; z0 = lo[2]; inv_scale = 1.0 / scale;
}

__global__ void k_keys(int64_t n, const double *__restrict__ xyzh, const int8_t *__restrict__ iphase, const unsigned long long *__restrict__ bbox_enc,
                       const __grid_constant__ DevParams dp, unsigned long long *__restrict__ keys, int *__restrict__ idx, int hilbert)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x0, y0, z0, inv_scale;
    key_frame(bbox_enc, dp, x0, y0, z0, inv_scale);
    const double4 x = reinterpret_cast<const double4 *>(xyzh)[i];
    unsigned long long key;
    if (x.w < DBL_MIN) key = ~0ull;   // dead/accreted particles sort to the end (part.F90:931)
    else {
        const double s = 2097152.0;   // 2^21
        double ux = fmin(fmax((x.x - x0) * inv_scale, 0.0), 0.99999994) * s;
        double uy = fmin(fmax((x.y - y0) * inv_scale, 0.0), 0.99999994) * s;
        double uz = fmin(fmax((x.z - z0) * inv_scale, 0.0), 0.99999994) * s;
        unsigned ix = (unsigned)ux, iy = (unsigned)uy, iz = (unsigned)uz;
        if (hilbert) hilbert_transpose(ix, iy, iz);
        key = (spread21((unsigned long long)ix) << 2) | (spread21((unsigned long long)iy) << 1) | spread21((unsigned long long)iz);
        key = ((unsigned long long)sort_class(iphase[i]) << 61) | (key >> 2);     // bit 63 stays clear (cpl counts 63 key bits)
        key &= ~0xFFFFull;            // 15 bits per axis decide the order (6 radix passes instead of 8); ties are split by index
    }
    keys[i] = key;
    idx[i] = (int)i;
}

__global__ void k_gather_pos(int64_t n, const int *__restrict__ perm, const double *__restrict__ xyzh, const int8_t *__restrict__ iphase,
                             double4 *__restrict__ pos4, int8_t *__restrict__ stype, const unsigned long long *__restrict__ keys,
                             unsigned char *__restrict__ cpl, unsigned long long *cnt, int boundary_is_gas)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nlive = (int64_t)cnt[CNT_NLIVE];
    if (s >= n) return;
    if (s >= nlive) { cpl[s] = 0; return; }
    const int i = perm[s];
    pos4[s] = reinterpret_cast<const double4 *>(xyzh)[i];
    const int8_t ph = iphase[i];
    stype[s] = ph;
    // anything but gas makes the set "multitype" (general kernels); boundary particles of the gas mass behave as gas in every sum
    // (dens.F90:717-723, force.F90:1539) and differ only in being inactive, which the fast kernels honour through get_partinfo
    const int ta = ph < 0 ? -ph : ph;
    if (!(ta == IGAS || (boundary_is_gas && ta == IBOUNDARY))) cnt[CNT_MULTITYPE] = 1ull;
    { const int sc = sort_class(ph); if (sc == 1) cnt[CNT_CLASS1] = 1ull; else if (sc == 2) cnt[CNT_CLASS2] = 1ull; }
    if (ph <= 0) cnt[CNT_ANYINACTIVE] = 1ull;
    // common leading bits (of the 63 key bits) with the previous key: 0..63
    unsigned char c = 0;
    if (s > 0) {
        unsigned long long x = keys[s] ^ keys[s - 1];
        c = x ? (unsigned char)(__clzll((long long)x) - 1) : (unsigned char)63;
    }
    cpl[s] = c;
}

// leaf-cell boundaries: particle s starts a cell iff its predecessor is not in the same maximal binary-radix node
// (key prefix) holding <= tmax particles.  Prefix nodes of a Morton key are boxes (each bit halves one axis).
__global__ void k_cell_flags(int64_t n, const unsigned long long *__restrict__ cnt, const unsigned char *__restrict__ cpl, int tmax, int *__restrict__ flag)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nlive = (int64_t)cnt[CNT_NLIVE];
    if (s >= n) return;
    if (s >= nlive) { flag[s] = 0; return; }
    // depth = shallowest d at which at most tmax particles share their first d key bits with s.  Particle s-k shares
    // mL_k = min cpl(s-k+1..s) bits with s, particle s+k shares mR_k = min cpl(s+1..s+k); both sequences fall with k, so the tmax-th
    // largest of all of them comes out of a merge of at most tmax steps, and depth is one more than it (0 if there are fewer).
    int kL = 1, kR = 1;
    int curL = (s - 1 >= 0) ? (int)cpl[s] : -1, curR = (s + 1 < nlive) ? (int)cpl[s + 1] : -1;
    int depth = 0;
    for (int t = 1; t <= tmax; t++) {
        if (curL < 0 && curR < 0) break;                     // fewer than tmax others in reach: everything fits at depth 0
        if (curL >= curR) {
            if (t == tmax) depth = curL + 1;
            kL++;
            curL = (kL <= tmax && s - kL >= 0) ? min(curL, (int)cpl[s - kL + 1]) : -1;
        } else {
            if (t == tmax) depth = curR + 1;
            kR++;
            curR = (kR <= tmax && s + kR < nlive) ? min(curR, (int)cpl[s + kR]) : -1;
        }
    }
    if (depth > 64) depth = 64;
    if (depth < 2) depth = 2;                   // the two class bits: a cell never holds two sort classes
    int f;
    if (s == 0) f = 1;
    else if (depth == 64) f = (cpl[s] < 63) || (s % tmax == 0);   // > tmax identical keys: split the run arbitrarily
    else f = (int)cpl[s] < depth;
    flag[s] = f;
}

// M = number of leaf cells = scan[nlive-1] + flag[nlive-1], left on the device (CNT_NCELLS); M beyond the allocated capacity raises CNT_CELLOVER
__global__ void k_cell_count(const int *__restrict__ flag, const int *__restrict__ scan, unsigned long long *cnt, long long cap)
{
    const long long nlive = (long long)cnt[CNT_NLIVE];
    const unsigned long long M = nlive > 0 ? (unsigned long long)(scan[nlive - 1] + flag[nlive - 1]) : 0ull;
    cnt[CNT_NCELLS] = M;
    cnt[CNT_CELLOVER] = ((long long)M > cap) ? 1ull : 0ull;
}

__global__ void k_cell_starts(int64_t n, const unsigned long long *__restrict__ cnt, const int *__restrict__ flag, const int *__restrict__ scan, Cell *__restrict__ cells)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n || s >= (int64_t)cnt[CNT_NLIVE] || cnt[CNT_CELLOVER]) return;
    if (flag[s]) cells[scan[s]].start = (int)s;
}

__global__ void k_cell_props(const unsigned long long *__restrict__ cnt, Cell *__restrict__ cells, const double4 *__restrict__ pos4, const int8_t *__restrict__ stype,
                             const unsigned long long *__restrict__ keys, unsigned long long *__restrict__ cellkeys, int sbta, int use_dust, int ind_ts)
{
    int64_t cidx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t ncells = (int64_t)cnt[CNT_NCELLS], nlive = (int64_t)cnt[CNT_NLIVE];
    if (cidx >= ncells || cnt[CNT_CELLOVER]) return;
    Cell c = cells[cidx];
    const int end = (cidx + 1 < ncells) ? cells[cidx + 1].start : (int)nlive;
    c.count = end - c.start;
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX}, hmax = 0.;
    int nact = 0;
    for (int s = c.start; s < end; s++) {
        const double4 x = pos4[s];
        lo[0] = fmin(lo[0], x.x); lo[1] = fmin(lo[1], x.y); lo[2] = fmin(lo[2], x.z);
        hi[0] = fmax(hi[0], x.x); hi[1] = fmax(hi[1], x.y); hi[2] = fmax(hi[2], x.z);
        hmax = fmax(hmax, x.w);
        bool a, g, d; int t;
        get_partinfo_d(stype[s], sbta, use_dust, a, g, d, t);
        (void)ind_ts;
        if (a) nact++;
    }
    for (int k = 0; k < 3; k++) { c.lo[k] = lo[k]; c.hi[k] = hi[k]; }
    c.hmax = hmax; c.active = nact; c.parent = -1;
    cells[cidx] = c;
    cellkeys[cidx] = keys[c.start];
}

// ---- Karras binary radix tree over the M cell keys -------------------------------------------
__device__ __forceinline__ int delta_k(const unsigned long long *keys, int M, int i, int j)
{
    if (j < 0 || j >= M) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_radix_tree(const unsigned long long *__restrict__ cnt, const unsigned long long *__restrict__ keys, TreeNode *__restrict__ nodes, TreeNodeF *__restrict__ nodesf, Cell *__restrict__ cells)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = (int)cnt[CNT_NCELLS];
    if (i >= M - 1 || cnt[CNT_CELLOVER]) return;
    const int d = (delta_k(keys, M, i, i + 1) - delta_k(keys, M, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta_k(keys, M, i, i - d);
    int lmax = 2;
    while (delta_k(keys, M, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta_k(keys, M, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta_k(keys, M, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta_k(keys, M, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    int cl, cr;
    if (lo == gamma) { cl = ~gamma; cells[gamma].parent = i; } else { cl = gamma; nodes[gamma].parent = i; }
    if (hi == gamma + 1) { cr = ~(gamma + 1); cells[gamma + 1].parent = i; } else { cr = gamma + 1; nodes[gamma + 1].parent = i; }
    nodes[i].child[0] = cl; nodes[i].child[1] = cr;
    nodesf[i].child[0] = cl; nodesf[i].child[1] = cr;
    if (i == 0) nodes[0].parent = -1;
}

// bottom-up refit: the second thread to arrive at a node merges its two child boxes into the grandparent slot
__device__ __forceinline__ void write_nodef(TreeNodeF *nf, int slot, const double *lo, const double *hi, double hmax)
{
    for (int k = 0; k < 3; k++) { nf->lo[slot][k] = __double2float_rd(lo[k]); nf->hi[slot][k] = __double2float_ru(hi[k]); }
    nf->hmax[slot] = __double2float_ru(hmax);
}

__global__ void k_refit(const unsigned long long *__restrict__ tcnt, const Cell *__restrict__ cells, TreeNode *nodes, TreeNodeF *nodesf, int *flags)
{
    int cidx = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = (int)tcnt[CNT_NCELLS];
    if (cidx >= M || tcnt[CNT_CELLOVER]) return;
    const Cell c = cells[cidx];
    double lo[3] = {c.lo[0], c.lo[1], c.lo[2]}, hi[3] = {c.hi[0], c.hi[1], c.hi[2]}, hmax = c.hmax;
    int cnt = c.count, start = c.start, act = c.active;
    int me = ~cidx, parent = c.parent;
    while (parent >= 0) {
        TreeNode *nd = &nodes[parent];
        const int slot = (nd->child[0] == me) ? 0 : 1;
        for (int k = 0; k < 3; k++) { nd->lo[slot][k] = lo[k]; nd->hi[slot][k] = hi[k]; }
        nd->hmax[slot] = hmax;
        nd->cnt[slot] = cnt; nd->start[slot] = start; nd->act[slot] = act;
        write_nodef(&nodesf[parent], slot, lo, hi, hmax);
        // the walk's copy names a leaf child by its packed particle range, so that a leaf hit needs no second (dependent) read
        if (me < 0) nodesf[parent].child[slot] = (int)(0x80000000u | ((unsigned)start << 5) | (unsigned)(cnt - 1));
        __threadfence();
        if (atomicAdd(&flags[parent], 1) == 0) return;     // first arrival: sibling not ready yet
        __threadfence();
        const int o = 1 - slot;
        const volatile TreeNode *vn = nd;
        for (int k = 0; k < 3; k++) { lo[k] = fmin(lo[k], vn->lo[o][k]); hi[k] = fmax(hi[k], vn->hi[o][k]); }
        hmax = fmax(hmax, vn->hmax[o]);
        cnt += vn->cnt[o]; start = min(start, vn->start[o]); act += vn->act[o];
        me = parent;
        parent = vn->parent;
    }
}

// target groups: maximal subtrees holding <= gmax particles (one lane per target in the pair kernels)
// (a subtree that spans two sort classes is never a group: the classes are key prefixes, so it is one of the top nodes of the tree)
__device__ __forceinline__ bool mixed_classes(const int8_t *__restrict__ stype, int start, int count)
{
    return sort_class(stype[start]) != sort_class(stype[start + count - 1]);
}

__global__ void k_groups(const unsigned long long *__restrict__ cnt, int gmax, const Cell *__restrict__ cells, const TreeNode *__restrict__ nodes, const int8_t *__restrict__ stype,
                         Cell *__restrict__ groups, unsigned long long *ngroups)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = (int)cnt[CNT_NCELLS];
    if (t >= 2 * M - 1 || cnt[CNT_CELLOVER]) return;
    int me, parent, total, start;
    if (t < M) { me = ~t; parent = cells[t].parent; total = cells[t].count; start = cells[t].start; }
    else { const int i = t - M; me = i; parent = nodes[i].parent; total = nodes[i].cnt[0] + nodes[i].cnt[1]; start = min(nodes[i].start[0], nodes[i].start[1]); }
    if (total > gmax) return;
    if (t >= M && mixed_classes(stype, start, total)) return;
    Cell g;
    if (parent >= 0) {
        const TreeNode &pn = nodes[parent];
        const int ptotal = pn.cnt[0] + pn.cnt[1];
        if (ptotal <= gmax && !mixed_classes(stype, min(pn.start[0], pn.start[1]), ptotal)) return;          // the parent is (inside) a group already
        const int slot = (pn.child[0] == me) ? 0 : 1;
        for (int k = 0; k < 3; k++) { g.lo[k] = pn.lo[slot][k]; g.hi[k] = pn.hi[slot][k]; }
        g.hmax = pn.hmax[slot]; g.start = pn.start[slot]; g.count = pn.cnt[slot]; g.active = pn.act[slot]; g.parent = parent;
    } else if (t < M) {
        g = cells[t];                                         // a single cell is the whole tree
    } else {
        const TreeNode &nd = nodes[me];                       // the root itself fits in one group
        for (int k = 0; k < 3; k++) { g.lo[k] = fmin(nd.lo[0][k], nd.lo[1][k]); g.hi[k] = fmax(nd.hi[0][k], nd.hi[1][k]); }
        g.hmax = fmax(nd.hmax[0], nd.hmax[1]); g.start = min(nd.start[0], nd.start[1]); g.count = total; g.active = nd.act[0] + nd.act[1]; g.parent = -1;
    }
    const unsigned long long slotg = atomicAdd(ngroups, 1ull);
    groups[slotg] = g;
}

// target groups with better fill: inside every maximal subtree holding <= smax particles the Morton-consecutive leaf cells are packed
// greedily into groups of <= gmax particles (a maximal <= gmax subtree holds 22.6 of 32 particles on average away from power-of-two
// lattices, tools/group_fill.py).  A group is then a run of whole cells: box, hmax and active count are merged from its cells.
// scan[s] = index of the cell that starts at sorted slot s (k_cell_starts).  One thread per subtree: two passes (count, then write into
// a contiguous block of group slots so that consecutive groups stay spatial neighbours).
__global__ void k_groups_packed(const unsigned long long *__restrict__ cnt, int gmax, int smax, const Cell *__restrict__ cells, const TreeNode *__restrict__ nodes, const int *__restrict__ scan,
                                const int8_t *__restrict__ stype, Cell *__restrict__ groups, unsigned long long *ngroups)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = (int)cnt[CNT_NCELLS];
    if (t >= 2 * M - 1 || cnt[CNT_CELLOVER]) return;
    int parent, total, start;
    if (t < M) { parent = cells[t].parent; total = cells[t].count; start = cells[t].start; }
    else { const TreeNode &nd = nodes[t - M]; parent = nd.parent; total = nd.cnt[0] + nd.cnt[1]; start = min(nd.start[0], nd.start[1]); }
    if (total > smax && t >= M) return;                     // (a single cell above smax cannot occur: cells hold <= 32 particles)
    if (t >= M && mixed_classes(stype, start, total)) return;
    if (parent >= 0) {
        const int ptotal = nodes[parent].cnt[0] + nodes[parent].cnt[1];
        if (ptotal <= smax && !mixed_classes(stype, min(nodes[parent].start[0], nodes[parent].start[1]), ptotal)) return;      // not maximal
    }
    const int c0 = (t < M) ? t : scan[start];
    int ng = 0, acc = 0;
    for (int cidx = c0, left = total; left > 0; cidx++) {   // pass 1: number of groups
        const int cnt = cells[cidx].count;
        if (acc == 0 || acc + cnt > gmax) { ng++; acc = 0; }
        acc += cnt; left -= cnt;
    }
    unsigned long long slot = atomicAdd(ngroups, (unsigned long long)ng);
    Cell g; g.count = 0;
    for (int cidx = c0, left = total; left > 0; cidx++) {   // pass 2: merge and write
        const Cell cc = cells[cidx];
        if (g.count > 0 && g.count + cc.count > gmax) { groups[slot++] = g; g.count = 0; }
        if (g.count == 0) { g = cc; g.parent = parent; }
        else {
            for (int k = 0; k < 3; k++) { g.lo[k] = fmin(g.lo[k], cc.lo[k]); g.hi[k] = fmax(g.hi[k], cc.hi[k]); }
            g.hmax = fmax(g.hmax, cc.hmax); g.count += cc.count; g.active += cc.active;
        }
        left -= cc.count;
    }
    if (g.count > 0) groups[slot] = g;
}

__global__ void k_cell_hmax(int64_t ncells, Cell *__restrict__ cells, const double4 *__restrict__ pos4)
{
    int64_t cidx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cidx >= ncells) return;
    const int start = cells[cidx].start, count = cells[cidx].count;
    double hmax = 0.;
    for (int s = start; s < start + count; s++) hmax = fmax(hmax, pos4[s].w);
    cells[cidx].hmax = hmax;
}

}  // namespace

#define LAUNCH(c, kern, grid, block, ...)                         \
    do {                                                          \
        kern<<<(grid), (block), 0, (c)->stream>>>(__VA_ARGS__);   \
        (c)->launches++;                                          \
    } while (0)

static inline int nblk(int64_t n, int b) { return (int)((n + b - 1) / b); }

// groups of the current tree; cap = number of cells the grids are sized for (the kernels read the true count on the device).
// The group count is left in counters[CNT_COUNT-1]: tree_build reads it with its other counts in ONE round trip.
static int build_groups(sphgpu_ctx *c, int64_t cap)
{
    CUDA_TRY(c, c->groups.ensure(cap));
    unsigned long long *ng = c->counters.p + CNT_COUNT - 1;
    CUDA_TRY(c, cudaMemsetAsync(ng, 0, sizeof(unsigned long long), c->stream));
    if (c->group_pack > c->max_cell && c->max_leaf <= c->max_cell)
        LAUNCH(c, k_groups_packed, nblk(2 * cap - 1, 128), 128, c->counters.p, c->max_cell, c->group_pack, c->cells.p, c->nodes.p, c->cellid_scan.p, c->stype.p, c->groups.p, ng);
    else
        LAUNCH(c, k_groups, nblk(2 * cap - 1, 128), 128, c->counters.p, c->max_cell, c->cells.p, c->nodes.p, c->stype.p, c->groups.p, ng);
    return SPHGPU_OK;
}

int tree_refit_hmax(sphgpu_ctx *c)
{
    const int M = (int)c->ncells;
    c->hscale = 1.; c->wl_force_ok = false;
    LAUNCH(c, k_cell_hmax, nblk(M, 128), 128, M, c->cells.p, c->pos4.p);
    if (M > 1) {
        CUDA_TRY(c, cudaMemsetAsync(c->nodeflag.p, 0, sizeof(int) * (size_t)M, c->stream));
        LAUNCH(c, k_refit, nblk(M, 128), 128, c->counters.p, c->cells.p, c->nodes.p, c->nodesf.p, c->nodeflag.p);
    }
    // the structure is unchanged and grouping depends on particle counts only, so the number of groups is too: no read-back
    // (the order of the groups may differ, which nothing depends on)
    TRY(build_groups(c, M));
    CUDA_TRY(c, cudaGetLastError());
    return SPHGPU_OK;
}

// One host round trip per build. Everything that used to be read back mid-build (live count and bounding box for the key
// frame, number of leaf cells for the allocation sizes) stays in c->counters; kernels are launched on grids sized from the
// particle count / the cell capacity and read the true counts on the device. The single read at the end returns error
// state, nlive, ncells, ngroups and the multi-type flag. The cell arrays are sized from the previous build (+25%) or
// nlive/3; a build that finds more cells than that raises CNT_CELLOVER, skips the cell-indexed kernels and is repeated once
// with the worst case (one cell per particle).
int tree_build(sphgpu_ctx *c)
{
    const int64_t n = c->npart;
    if (n <= 0) { c->err = "build_tree: no particles"; return SPHGPU_ERR_NOPART; }
    c->tree_valid = false; c->dens_valid = false;
    const sphgpu_params &p = c->hp.p;
    CUDA_TRY(c, c->keys.ensure(n)); CUDA_TRY(c, c->keys_alt.ensure(n));
    CUDA_TRY(c, c->perm.ensure(n)); CUDA_TRY(c, c->perm_alt.ensure(n));
    CUDA_TRY(c, c->counters.ensure(CNT_COUNT)); CUDA_TRY(c, c->dscal.ensure(DS_COUNT));
    CUDA_TRY(c, c->pos4.ensure(n)); CUDA_TRY(c, c->stype.ensure(n)); CUDA_TRY(c, c->cpl.ensure(n));
    CUDA_TRY(c, c->cellflag.ensure(n)); CUDA_TRY(c, c->cellid_scan.ensure(n));
    CUDA_TRY(c, c->h_build.ensure(n)); CUDA_TRY(c, c->h_its.ensure(n)); if (p.gravity) CUDA_TRY(c, c->h_hist.ensure((size_t)SPHGPU_HHIST * n));
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, c->keys_alt.p, c->keys.p, c->perm_alt.p, c->perm.p, (int)n, 16, 64, c->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, c->cellflag.p, c->cellid_scan.p, (int)n, c->stream);
    CUDA_TRY(c, c->cubtemp.ensure(tb > tb2 ? tb : tb2));
    int64_t cap = c->ncells > 0 ? c->ncells + c->ncells / 4 + 1024 : 0;
    if (cap < n / 3 + 1024) cap = n / 3 + 1024;
    if (cap > n) cap = n;
    unsigned long long hc[CNT_COUNT];
    for (int attempt = 0; attempt < 2; attempt++) {
        CUDA_TRY(c, c->cells.ensure(cap)); CUDA_TRY(c, c->cellkeys.ensure(cap)); CUDA_TRY(c, c->nodes.ensure(cap)); CUDA_TRY(c, c->nodesf.ensure(cap));
        CUDA_TRY(c, c->nodeflag.ensure(cap));
        CUDA_TRY(c, cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * CNT_COUNT, c->stream));
        // bbox accumulators (ordered encoding): min slots start at all-ones, max slots at zero
        unsigned long long *bbox_enc = c->counters.p + 16;
        CUDA_TRY(c, cudaMemsetAsync(bbox_enc, 0xff, sizeof(unsigned long long) * 3, c->stream));
        LAUNCH(c, k_wrap_count, c->numSMs * 8, 256, n, c->xyzh.p, c->hp, c->counters.p, bbox_enc, c->h_build.p, c->h_its.p);
        LAUNCH(c, k_keys, nblk(n, 256), 256, n, c->xyzh.p, c->iphase.p, bbox_enc, c->hp, c->keys_alt.p, c->perm_alt.p, c->hilbert ? 1 : 0);
        size_t tbb = c->cubtemp.cap;
        CUDA_TRY(c, cub::DeviceRadixSort::SortPairs(c->cubtemp.p, tbb, c->keys_alt.p, c->keys.p, c->perm_alt.p, c->perm.p, (int)n, 16, 64, c->stream));
        c->launches += 7;
        LAUNCH(c, k_gather_pos, nblk(n, 256), 256, n, c->perm.p, c->xyzh.p, c->iphase.p, c->pos4.p, c->stype.p, c->keys.p, c->cpl.p, c->counters.p,
               (p.massoftype[IBOUNDARY] == p.massoftype[IGAS]) ? 1 : 0);
        LAUNCH(c, k_cell_flags, nblk(n, 128), 128, n, c->counters.p, c->cpl.p, c->max_leaf, c->cellflag.p);
        tbb = c->cubtemp.cap;
        CUDA_TRY(c, cub::DeviceScan::ExclusiveSum(c->cubtemp.p, tbb, c->cellflag.p, c->cellid_scan.p, (int)n, c->stream));
        c->launches += 2;
        LAUNCH(c, k_cell_count, 1, 1, c->cellflag.p, c->cellid_scan.p, c->counters.p, (long long)cap);
        LAUNCH(c, k_cell_starts, nblk(n, 256), 256, n, c->counters.p, c->cellflag.p, c->cellid_scan.p, c->cells.p);
        LAUNCH(c, k_cell_props, nblk(cap, 128), 128, c->counters.p, c->cells.p, c->pos4.p, c->stype.p, c->keys.p, c->cellkeys.p,
               p.set_boundaries_to_active, p.dust, p.ind_timesteps);
        if (cap > 1) {
            LAUNCH(c, k_radix_tree, nblk(cap - 1, 128), 128, c->counters.p, c->cellkeys.p, c->nodes.p, c->nodesf.p, c->cells.p);
            CUDA_TRY(c, cudaMemsetAsync(c->nodeflag.p, 0, sizeof(int) * (size_t)cap, c->stream));
            LAUNCH(c, k_refit, nblk(cap, 128), 128, c->counters.p, c->cells.p, c->nodes.p, c->nodesf.p, c->nodeflag.p);
        }
        TRY(build_groups(c, cap));
        CUDA_TRY(c, cudaMemcpyAsync(hc, c->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (hc[CNT_ERR]) {
            char buf[128]; snprintf(buf, sizeof buf, "maketree: NaN in particle position, likely caused by NaN in force (particle %llu)", hc[CNT_ERRID]);
            c->err = buf; return (int)hc[CNT_ERR];
        }
        if (hc[CNT_NLIVE] == 0) { c->err = "maketree: no particles or all particles dead/accreted"; return SPHGPU_ERR_NOPART; }
        if (!hc[CNT_CELLOVER]) break;
        if (attempt == 1) { c->err = "build_tree: cell capacity exceeded twice"; return SPHGPU_ERR_ARG; }
        cap = (int64_t)hc[CNT_NCELLS];
    }
    c->nlive = (int64_t)hc[CNT_NLIVE];
    c->ncells = (int64_t)hc[CNT_NCELLS];
    c->ngroups = (int64_t)hc[CNT_COUNT - 1];
    c->multitype = hc[CNT_MULTITYPE] != 0;
    c->class_mask = 1 | (hc[CNT_CLASS1] ? 2 : 0) | (hc[CNT_CLASS2] ? 4 : 0);
    c->any_inactive = hc[CNT_ANYINACTIVE] != 0;
    c->grav_tree_valid = false; c->hscale = 1.; c->wl_force_ok = false;
    CUDA_TRY(c, cudaGetLastError());
    c->tree_valid = true;
    return SPHGPU_OK;
}
