// neigh.cu -- exact neighbour sets of the current tree/h state in CSR form, for the bit-exact set parity tests
// (the reference's own check is test_neigh.f90:264-367: tree neighbour counts == O(N^2) brute force).
// Uses the same walk and the same non-contracted distance test as the density / force kernels.
#include "walk.cuh"
#include "sphkern.cuh"
#include <cub/cub.cuh>
#include <float.h>

namespace {

template <bool PERIODIC, bool SYM, bool FILL>
__global__ void __launch_bounds__(128) k_neigh(const TreeNodeF *nodes, const Cell *cells, int ncells, const double4 *pos4, const int *perm, double radkern,
                                               double Lx, double Ly, double Lz, int *stage_idx, int scratch_per_warp, int max_leaf, unsigned long long *cnt,
                                               int *counts, const long long *offsets, int *out)
{
    __shared__ WarpShared wsh[4];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    WarpShared &ws = wsh[wib];
    int *clist = stage_idx + (size_t)(blockIdx.x * 4 + wib) * scratch_per_warp;
    const double radkern2 = radkern * radkern;
    const unsigned lt_mask = (1u << lane) - 1;
    while (true) {
        int cellid = 0;
        if (lane == 0) cellid = (int)atomicAdd(&cnt[CNT_WORK], 1ull);
        cellid = __shfl_sync(FULLMASK, cellid, 0);
        if (cellid >= ncells) break;
        const Cell cell = cells[cellid];
        float tlo[3], thi[3];
        for (int k = 0; k < 3; k++) { tlo[k] = __double2float_rd(cell.lo[k]); thi[k] = __double2float_ru(cell.hi[k]); }
        const double cx = 0.5 * (cell.lo[0] + cell.hi[0]), cy = 0.5 * (cell.lo[1] + cell.hi[1]), cz = 0.5 * (cell.lo[2] + cell.hi[2]);
        float reach = 0.f;
        const int ncl = warp_walk<SYM, PERIODIC>(nodes, cells, ncells, tlo, thi, __double2float_ru(radkern * cell.hmax), (float)radkern, (float)Lx, (float)Ly,
                                                 (float)Lz, ws.stack, clist, scratch_per_warp, reach);
        if (ncl < 0) { if (lane == 0) atomicMax(&cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); break; }
        const FilterScale fs = filter_scale(0.f, reach);     // the FP16 filter words are staged but not used here: every candidate gets the exact test
        int nfound[32];                                      // per target of the cell (max_leaf <= 32), kept across rounds
        for (int t = 0; t < 32; t++) nfound[t] = 0;
        for (int cellpos = 0; cellpos < ncl;) {
          const int nlist = stage_round<PERIODIC, false>(ws, clist, ncl, cellpos, reinterpret_cast<const double2 *>(pos4), 2, cx, cy, cz, Lx, Ly, Lz, (float)radkern, max_leaf, fs);
          const int *list = ws.sidx;
          for (int t = 0; t < cell.count; t++) {
            const int s = cell.start + t;
            const double4 pi = pos4[s];
            const double hi1 = 1. / pi.w, hi21 = hi1 * hi1;
            const int iorig = perm[s];
            int n = nfound[t];
            for (int c0 = 0; c0 < nlist; c0 += 32) {
                const int idx = c0 + lane;
                bool pass = false; int j = 0;
                if (idx < nlist) {
                    j = list[idx];
                    const double4 pj = pos4[j];
                    double dx, dy, dz;
                    const double r2 = pair_r2<PERIODIC>(pi.x, pi.y, pi.z, pj, Lx, Ly, Lz, dx, dy, dz);
                    pass = __dmul_rn(r2, hi21) < radkern2;
                    if (SYM) { const double hj1 = 1. / pj.w; pass = pass || (__dmul_rn(r2, hj1 * hj1) < radkern2); }
                    pass = pass && (j != s);
                }
                const unsigned m = __ballot_sync(FULLMASK, pass);
                if (FILL && pass) out[offsets[iorig] + n + __popc(m & lt_mask)] = perm[j] + 1;
                n += __popc(m);
            }
            nfound[t] = n;
            if (!FILL && lane == 0) counts[iorig] = n;
          }
          __syncwarp();
        }
    }
}

}  // namespace

int64_t neighbour_sets_run(sphgpu_ctx *c, int symmetric, int64_t *offsets, int32_t *list, int64_t maxlist)
{
    if (!c->tree_valid) { c->err = "neighbour_sets: build_tree has not been called"; return -1; }
    if (c->hscale != 1.) { if (tree_refit_hmax(c) != SPHGPU_OK) return -1; }      // the cells' hmax must hold the current h
    const int64_t n = c->npart;
    const int grid = c->numSMs * 4;
    if (c->stage_idx.ensure((size_t)grid * 4 * c->scratch_per_warp) != cudaSuccess) return -1;
    DevBuf<int> counts; DevBuf<long long> offs; DevBuf<int> out; DevBuf<char> tmp;
    if (counts.ensure(n + 1) != cudaSuccess || offs.ensure(n + 1) != cudaSuccess) return -1;
    cudaMemsetAsync(counts.p, 0, sizeof(int) * (n + 1), c->stream);
    cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 16, c->stream);
    const double R = c->hp.kc.radkern, Lx = c->hp.dxbound, Ly = c->hp.dybound, Lz = c->hp.dzbound;
    const bool per = c->hp.p.periodic;
#define NEIGH_LAUNCH(FILLV)                                                                                                                       \
    do {                                                                                                                                          \
        if (per && symmetric) k_neigh<true, true, FILLV><<<grid, 128, 0, c->stream>>>(c->nodesf.p, c->cells.p, (int)c->ncells, c->pos4.p, c->perm.p, R, Lx, Ly, Lz, c->stage_idx.p, c->scratch_per_warp, c->max_leaf, c->counters.p, counts.p, offs.p, out.p); \
        else if (per) k_neigh<true, false, FILLV><<<grid, 128, 0, c->stream>>>(c->nodesf.p, c->cells.p, (int)c->ncells, c->pos4.p, c->perm.p, R, Lx, Ly, Lz, c->stage_idx.p, c->scratch_per_warp, c->max_leaf, c->counters.p, counts.p, offs.p, out.p); \
        else if (symmetric) k_neigh<false, true, FILLV><<<grid, 128, 0, c->stream>>>(c->nodesf.p, c->cells.p, (int)c->ncells, c->pos4.p, c->perm.p, R, Lx, Ly, Lz, c->stage_idx.p, c->scratch_per_warp, c->max_leaf, c->counters.p, counts.p, offs.p, out.p); \
        else k_neigh<false, false, FILLV><<<grid, 128, 0, c->stream>>>(c->nodesf.p, c->cells.p, (int)c->ncells, c->pos4.p, c->perm.p, R, Lx, Ly, Lz, c->stage_idx.p, c->scratch_per_warp, c->max_leaf, c->counters.p, counts.p, offs.p, out.p); \
        c->launches++;                                                                                                                            \
    } while (0)
    NEIGH_LAUNCH(false);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, offs.p, (int)(n + 1), c->stream);
    tmp.ensure(tb);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, offs.p, (int)(n + 1), c->stream);
    long long total = 0;
    cudaMemcpyAsync(&total, offs.p + n, sizeof(long long), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    int64_t ret = total;
    if (total > maxlist) ret = -total;
    else {
        out.ensure(total + 1);
        cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 16, c->stream);
        NEIGH_LAUNCH(true);
        cudaMemcpyAsync(list, out.p, sizeof(int) * total, cudaMemcpyDeviceToHost, c->stream);
        std::vector<long long> ho(n + 1);
        cudaMemcpyAsync(ho.data(), offs.p, sizeof(long long) * (n + 1), cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        for (int64_t i = 0; i <= n; i++) offsets[i] = ho[i];
    }
    unsigned long long err = 0;
    cudaMemcpy(&err, c->counters.p + CNT_ERR, sizeof err, cudaMemcpyDeviceToHost);
    counts.release(); offs.release(); out.release(); tmp.release();
    if (cudaGetLastError() != cudaSuccess || err) { c->err = "neighbour_sets: kernel failure or scratch overflow"; return -1; }
    return ret;
}
