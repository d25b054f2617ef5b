// density.cu -- density pass with the h-rho Newton-Raphson iteration: replaces densityiterate
// (src/main/dens.F90:117-568) with get_density_sums (:578-857), finish_cell/finish_rhosum (:1382-1507),
// compute_hmax (:1275-1289) and store_results (:1511-1681).
//
// Mapping to the hardware: a persistent grid of warps pulls leaf cells (<= 32 particles) from an atomic work counter;
// per cell the warp walks the tree and stages the candidates once (walk.cuh), builds per-target FP32 hit masks
// (lane = candidate), then switches to lane = TARGET: each lane walks its own hit masks and accumulates its own
// 28 (+10 MHD) FP64 sums in registers -- no cross-lane reduction.  The h-rho Newton-Raphson step, the convergence
// test and store_results are lane-parallel.  The kernel is FP64-pipe bound.
//
// Semantics kept from the reference: self term added analytically (dens.F90:1491-1492), neighbours of the same
// (base) type only (:717-723), iteration per particle with clamp to +-20% and the tolh/omega test (:1425-1440),
// final h = hrho(rho) (:1595), real*4 rounding of gradh before it is reused in `term` (:1601,:1610,:1631),
// exact-linear gradients (:866-967), maxdensits = 100 (:101).
// Deliberate difference: divcurlB uses h_j as it was at the START of the pass for every neighbour; the reference
// reads xyzh(4,j) while other threads update it (fast_divcurlB race, config.F90:215-217), so its own result is
// thread-order dependent.
#include "walk.cuh"
#include "sphkern.cuh"
#include <float.h>
#include <string.h>

namespace {

struct DensArgs {
    const TreeNodeF *nodes; const Cell *cells; int ncells; const Cell *groups; int ngroups;
    double4 *pos4; const double4 *vel4, *acc4, *bev4; const int8_t *stype; const int *perm;
    const double4 *drec;     // fast path: 3 (MHD: 4) x 32 B per particle {x,y,z,h} {vx,vy,vz,ax} {ay,az,Bx,By} [{Bz,psi,-,-}], a = f + fext, B = (B/rho) rho(h)
    int *stage_idx; int multitype; int max_leaf; int class_mask; double hmax_global;
    WalkLists wl;       // cell lists prepared by k_walk_lists for the first pass of every group
    double *hnew; int *s_nneigh;                                   // sorted order: new h, neighbour count (< 0: not an active target)
    double *xyzh;                                                  // caller's order (fast path writes the new h directly)
    float *gradh, *divcurlv, *dvdx, *alphaind, *divcurlB; double *dustfrac;   // caller's order, written directly by the pair kernel (index perm[s])
    double *h_hist; int *h_its; int64_t npart;     // GRAV: per-particle h after every iteration, for the node-hmax replay of gravity.cu
    int scratch_per_warp; unsigned long long *cnt; double *dscal;
    double margin; int icall;
    double mask_margin;   // > 0: the FP16 hit masks are built for radkern h (1 + mask_margin) so that the next h-rho iteration can reuse them
};

// DENS_STAGE = 1: the fast kernel keeps the candidates' FP64 positions in shared memory (measured 1.53 vs 1.64 ms on turb 128^3 against
// reading them with the rest of the record)
#ifndef DENS_STAGE
#define DENS_STAGE 1
#endif

enum {  // slots of the per-lane partial sums (order of dens.F90:51-95)
    S_RHO = 0, S_GRADH, S_GRADSOFT, S_DIVV, S_DVXDX, S_DVXDY, S_DVXDZ, S_DVYDX, S_DVYDY, S_DVYDZ, S_DVZDX, S_DVZDY, S_DVZDZ,
    S_DAXDX, S_DAXDY, S_DAXDZ, S_DAYDX, S_DAYDY, S_DAYDZ, S_DAZDX, S_DAZDY, S_DAZDZ, S_RXX, S_RXY, S_RXZ, S_RYY, S_RYZ, S_RZZ, S_RHODUST
};
// MHD sums: div B and the three curl B components (the nine dB_a/dx_b sums of dens.F90:806-829 are only ever combined into the curl, :975-989)
enum { B_DIVB = 0, B_CURLX, B_CURLY, B_CURLZ, B_COUNT };

__global__ void k_gather_dens(int64_t nlive, const int *__restrict__ perm, const double *__restrict__ vxyzu, const double *__restrict__ fxyzu,
                              const double *__restrict__ fext, const double *__restrict__ Bevol, int nvu, int mhd, const double4 *__restrict__ pos4,
                              double4 *__restrict__ vel4, double4 *__restrict__ acc4, double4 *__restrict__ bev4, double *__restrict__ hnew,
                              int *__restrict__ s_nneigh, double4 *__restrict__ drec, double pmass, double hfact)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    const int i = perm[s];
    const double *v = vxyzu + (size_t)nvu * i, *f = fxyzu + (size_t)nvu * i, *fe = fext + 3 * (size_t)i;
    const double4 vv = make_double4(v[0], v[1], v[2], nvu >= 4 ? v[3] : 0.);
    const double4 aa = make_double4(f[0] + fe[0], f[1] + fe[1], f[2] + fe[2], 0.);     // dens.F90:1353-1355
    if (drec) {                                                  // packed record of the single-type fast path (96 / 128 bytes)
        double4 *r = drec + (mhd ? 4 : 3) * (size_t)s;
        const double4 pp = pos4[s];
        double4 be = make_double4(0., 0., 0., 0.);
        double rho = 0.;
        if (mhd) {
            be = reinterpret_cast<const double4 *>(Bevol)[i];
            rho = rhoh_d(pp.w, pmass, hfact);                               // rho_j of dens.F90:810, the same for every pair j enters
            bev4[s] = be;
        }
        r[0] = pp; r[1] = make_double4(vv.x, vv.y, vv.z, aa.x); r[2] = make_double4(aa.y, aa.z, be.x * rho, be.y * rho);
        if (mhd) r[3] = make_double4(be.z * rho, be.w, 0., 0.);
    } else {
        vel4[s] = vv; acc4[s] = aa;
        if (mhd) bev4[s] = reinterpret_cast<const double4 *>(Bevol)[i];
    }
    hnew[s] = pos4[s].w;
    s_nneigh[s] = -1;
}

// the new h of the active targets goes back only after the pass has succeeded (an overflow retry must start from the old h)
__global__ void k_scatter_h(int64_t nlive, const int *__restrict__ perm, const int *__restrict__ s_nneigh, const double *__restrict__ hnew,
                            double4 *__restrict__ pos4, double *__restrict__ xyzh)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    if (s_nneigh[s] < 0) return;     // not an active target: keep stored values (test_derivs.f90:233-238)
    const double h = hnew[s];
    pos4[s].w = h;                                       // treecache(4,:) refresh (dens.F90:1596)
    xyzh[4 * (size_t)perm[s] + 3] = h;
}

// overflow retry of the fast path: back to the h the tree was built with (k_wrap_count keeps it in caller order)
__global__ void k_restore_h_sorted(int64_t nlive, const int *__restrict__ perm, const double *__restrict__ h_build, double4 *__restrict__ pos4,
                                   double *__restrict__ xyzh)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    const int i = perm[s];
    const double h = h_build[i];
    pos4[s].w = h;
    xyzh[4 * (size_t)i + 3] = h;
}

// pair body: lane = target, j = this lane's next neighbour candidate (slot < 0: none).  Branch-free so that two
// neighbours per trip give the scheduler two independent FP64 dependency chains: the exact reference membership test
// (dens.F90:675-679, :650) becomes a 0/1 weight on m_j, through which every same-type sum of get_density_sums scales.
template <int K, bool PERIODIC, bool MHD, bool GRAV>
__device__ __forceinline__ void dens_pair(double (&v)[29], double (&w)[B_COUNT], int &nneighi, int slot, const int *__restrict__ idxlist, int s,
                                          const double4 &pi, double hi, double hi1, double hi21, int itypei, bool gasi, const double4 &vi,
                                          const double4 &ai, const double4 &bi, const DensArgs &a, const DevParams &dp, bool use_da, double Lx, double Ly,
                                          double Lz)
{
    typedef SphKern<K> KF;
    const int j = (slot >= 0) ? idxlist[slot] : s;
    const double4 pj = ldg256(a.pos4 + j);
    const double4 vj = ldg256(a.vel4 + j);
    const double4 aj = ldg256(a.acc4 + j);
    double dx, dy, dz;
    const double r2 = pair_r2<PERIODIC>(pi.x, pi.y, pi.z, pj, Lx, Ly, Lz, dx, dy, dz);
    const double q2i = __dmul_rn(r2, hi21);                                   // dens.F90:675
    const bool isn = (q2i < KF::radkern2) && (j != s);                        // :679, :650 (exact membership)
    // rij = sqrt(rij2), rij1 = 1/(rij + epsilon) (dens.F90:688,:746) from one reciprocal square root
    const double r2s = isn ? r2 : 1.0;
    const double rinv = rsqrt_pos(r2s);
    const double rij = r2s * rinv;
    const double qi = rij * hi1;
    double wabi, grkerni;
    KF::get_kernel(isn ? q2i : 1.0, qi, wabi, grkerni);
    int itypej = IGAS;
    if (a.multitype) itypej = abs((int)a.stype[j]);
    const int basej = (itypej == IBOUNDARY) ? IGAS : itypej;
    const bool same_type = (itypei == itypej) || (basej == itypei);
    const double pmj = dp.p.massoftype[itypej];
    const double pmassj = (isn && same_type) ? pmj : 0.;
    nneighi += (isn && same_type) ? 1 : 0;
    const double dwdhi = (-qi * grkerni - 3. * wabi);
    v[S_RHO] += wabi * pmassj;
    v[S_GRADH] += dwdhi * pmassj;
    if (GRAV) v[S_GRADSOFT] += KF::dphidh(isn ? q2i : 1.0, qi) * pmassj;
    const double rij1 = rinv - DBL_EPSILON * rinv * rinv;
    const double rij1grkern = rij1 * grkerni;
    const double runix = dx * rij1grkern * pmassj, runiy = dy * rij1grkern * pmassj, runiz = dz * rij1grkern * pmassj;
    const double dvx = vi.x - vj.x, dvy = vi.y - vj.y, dvz = vi.z - vj.z;
    v[S_DIVV] += dvx * runix + dvy * runiy + dvz * runiz;
    v[S_DVXDX] += dvx * runix; v[S_DVXDY] += dvx * runiy; v[S_DVXDZ] += dvx * runiz;
    v[S_DVYDX] += dvy * runix; v[S_DVYDY] += dvy * runiy; v[S_DVYDZ] += dvy * runiz;
    v[S_DVZDX] += dvz * runix; v[S_DVZDY] += dvz * runiy; v[S_DVZDZ] += dvz * runiz;
    if (use_da) {
        const double g = gasi ? 1. : 0.;                                      // gas_gas (dens.F90:722, :777)
        const double dax = (ai.x - aj.x) * g, day = (ai.y - aj.y) * g, daz = (ai.z - aj.z) * g;
        v[S_DAXDX] += dax * runix; v[S_DAXDY] += dax * runiy; v[S_DAXDZ] += dax * runiz;
        v[S_DAYDX] += day * runix; v[S_DAYDY] += day * runiy; v[S_DAYDZ] += day * runiz;
        v[S_DAZDX] += daz * runix; v[S_DAZDY] += daz * runiy; v[S_DAZDZ] += daz * runiz;
    }
    v[S_RXX] -= dx * runix; v[S_RXY] -= dx * runiy; v[S_RXZ] -= dx * runiz;
    v[S_RYY] -= dy * runiy; v[S_RYZ] -= dy * runiz; v[S_RZZ] -= dz * runiz;
    if (MHD) {
        const double g = gasi ? 1. : 0.;
        const double pmassi = dp.p.massoftype[itypei];
        const double rhoi = rhoh_d(hi, pmassi, dp.p.hfact);
        const double rhoj = rhoh_d(pj.w, pmj, dp.p.hfact);
        const double4 bj = ldg256(a.bev4 + j);
        const double dBx = (bi.x * rhoi - bj.x * rhoj) * g, dBy = (bi.y * rhoi - bj.y * rhoj) * g, dBz = (bi.z * rhoi - bj.z * rhoj) * g;
        w[B_DIVB] += dBx * runix + dBy * runiy + dBz * runiz;
        w[B_CURLX] += dBz * runiy - dBy * runiz;          // dBz/dy - dBy/dz
        w[B_CURLY] += dBx * runiz - dBz * runix;          // dBx/dz - dBz/dx
        w[B_CURLZ] += dBy * runix - dBx * runiy;          // dBy/dx - dBx/dy
    }
    if (dp.p.dust) v[S_RHODUST] += (isn && !same_type && gasi && itypej == IDUST) ? wabi : 0.;
}


// gas target, candidate of the dust class: the only density sum a pair of different classes enters is rho_dust (dens.F90:738-741);
// W is evaluated exactly as in dens_pair
template <int K, bool PERIODIC>
__device__ __forceinline__ void dens_pair_rhodust(double &rhodust, int slot, const int *__restrict__ idxlist, int s, const double4 &pi, double hi1, double hi21,
                                                  bool gasi, const DensArgs &a, double Lx, double Ly, double Lz)
{
    typedef SphKern<K> KF;
    const int j = (slot >= 0) ? idxlist[slot] : s;
    const double4 pj = ldg256(a.pos4 + j);
    double dx, dy, dz;
    const double r2 = pair_r2<PERIODIC>(pi.x, pi.y, pi.z, pj, Lx, Ly, Lz, dx, dy, dz);
    const double q2i = __dmul_rn(r2, hi21);
    const bool isn = (q2i < KF::radkern2) && (slot >= 0);
    const double r2s = isn ? r2 : 1.0;
    const double rinv = rsqrt_pos(r2s);
    const double qi = (r2s * rinv) * hi1;
    double wabi, grkerni;
    KF::get_kernel(isn ? q2i : 1.0, qi, wabi, grkerni);
    rhodust += (isn && gasi && abs((int)a.stype[j]) == IDUST) ? wabi : 0.;
}

// fast path (every particle has the same type and mass, no dust): same sums as dens_pair for TWO neighbours of the lane at once,
// written phase by phase over both so that two independent FP64 dependency chains are in flight.  No branches: the kernel is
// evaluated as truncated powers, a non-member (exact test fails, j == s, or the padding of an odd hit count) enters with weight 0
// on m_j, through which every sum scales, and r = 0 gives 1/r := 0 (rsqrt_pos).  The neighbour comes as one packed record; the
// minimum-image wrap is skipped for interior target groups.
template <int K, bool PERIODIC, bool MHD, bool GRAV, bool STAGED>
__device__ __forceinline__ void dens_pair2_fast(double (&v)[29], double (&w)[B_COUNT], int &nneighi, int slot0, int slot1, int myslot, unsigned rec2_s,
                                                unsigned rec1_s, unsigned sidx_s, double xi, double yi, double zi, double hi1, double hi21,
                                                const double4 &vi, const double4 &ai, const double4 &bi, const double4 *__restrict__ drec, double pmass0,
                                                bool use_da, bool interior, double Lx, double Ly, double Lz)
{
    typedef SphKern<K> KF;
    // a lane with an odd number of hits evaluates its first neighbour twice, the second time with weight 0
    const int sl[2] = {slot0, slot1 >= 0 ? slot1 : slot0};
    const bool live[2] = {slot0 != myslot, slot1 >= 0 && slot1 != myslot};
    double4 pj[2], vj[2], aj[2], bj[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        // the position (head of the dependency chain) comes from the staging block, v and a from the packed record in global memory
        const double4 *rj = drec + (MHD ? 4 : 3) * (size_t)lds_u32(sidx_s + 4u * (unsigned)sl[k]);
        if (STAGED) {
            const double2 XY = lds_d2(rec2_s + 16u * (unsigned)sl[k]);
            pj[k] = make_double4(XY.x, XY.y, lds_d(rec1_s + 8u * (unsigned)sl[k]), 0.);
        } else pj[k] = ldg256(rj);
        const double4 B = ldg256(rj + 1);                                    // {vx,vy,vz,ax} in one 256-bit load
        vj[k] = make_double4(B.x, B.y, B.z, 0.);
        if (MHD) {
            const double4 C = ldg256(rj + 2);                                // {ay,az,Bx,By}; B_j = (B/rho)_j rho(h_j), formed by k_gather_dens
            aj[k] = make_double4(B.w, C.x, C.y, 0.); bj[k] = make_double4(C.z, C.w, rj[3].x, 0.);
        } else {
            const double2 C = *reinterpret_cast<const double2 *>(rj + 2);    // {ay,az}
            aj[k] = make_double4(B.w, C.x, C.y, 0.);
        }
    }
    double dx[2], dy[2], dz[2];
#pragma unroll
    for (int k = 0; k < 2; k++) { dx[k] = xi - pj[k].x; dy[k] = yi - pj[k].y; dz[k] = zi - pj[k].z; }
    if (PERIODIC && !interior) {                                              // dens.F90:666-670
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (fabs(dx[k]) > 0.5 * Lx) dx[k] = dx[k] - copysign(Lx, dx[k]);
            if (fabs(dy[k]) > 0.5 * Ly) dy[k] = dy[k] - copysign(Ly, dy[k]);
            if (fabs(dz[k]) > 0.5 * Lz) dz[k] = dz[k] - copysign(Lz, dz[k]);
        }
    }
    double r2[2], q2i[2], pmass[2], rinv[2], qi[2], wabi[2], grkerni[2], g[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        r2[k] = __dadd_rn(__dadd_rn(__dmul_rn(dx[k], dx[k]), __dmul_rn(dy[k], dy[k])), __dmul_rn(dz[k], dz[k]));
        q2i[k] = __dmul_rn(r2[k], hi21);                                      // dens.F90:675
        const bool isn = (q2i[k] < KF::radkern2) && live[k];                  // :679, :650 (exact membership) -> 0/1 weight on m_j
        pmass[k] = isn ? pmass0 : 0.;
        nneighi += isn ? 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < 2; k++) rinv[k] = rsqrt_pos(r2[k]);
#pragma unroll
    for (int k = 0; k < 2; k++) { qi[k] = (r2[k] * rinv[k]) * hi1; KF::get_kernel_bf(qi[k], wabi[k], grkerni[k]); }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        v[S_RHO] += wabi[k] * pmass[k];
        v[S_GRADH] += (-qi[k] * grkerni[k] - 3. * wabi[k]) * pmass[k];
        if (GRAV) v[S_GRADSOFT] += KF::dphidh(fmin(q2i[k], KF::radkern2), fmin(qi[k], KF::radkern)) * pmass[k];
        g[k] = fma(-DBL_EPSILON * rinv[k], rinv[k], rinv[k]) * (grkerni[k] * pmass[k]);   // rij1 = 1/(rij + epsilon) (dens.F90:746), times grkern m_j
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double runix = dx[k] * g[k], runiy = dy[k] * g[k], runiz = dz[k] * g[k];
        const double dvx = vi.x - vj[k].x, dvy = vi.y - vj[k].y, dvz = vi.z - vj[k].z;
        v[S_DIVV] += dvx * runix + dvy * runiy + dvz * runiz;
        v[S_DVXDX] += dvx * runix; v[S_DVXDY] += dvx * runiy; v[S_DVXDZ] += dvx * runiz;
        v[S_DVYDX] += dvy * runix; v[S_DVYDY] += dvy * runiy; v[S_DVYDZ] += dvy * runiz;
        v[S_DVZDX] += dvz * runix; v[S_DVZDY] += dvz * runiy; v[S_DVZDZ] += dvz * runiz;
        if (use_da) {
            const double dax = ai.x - aj[k].x, day = ai.y - aj[k].y, daz = ai.z - aj[k].z;
            v[S_DAXDX] += dax * runix; v[S_DAXDY] += dax * runiy; v[S_DAXDZ] += dax * runiz;
            v[S_DAYDX] += day * runix; v[S_DAYDY] += day * runiy; v[S_DAYDZ] += day * runiz;
            v[S_DAZDX] += daz * runix; v[S_DAZDY] += daz * runiy; v[S_DAZDZ] += daz * runiz;
        }
        v[S_RXX] -= dx[k] * runix; v[S_RXY] -= dx[k] * runiy; v[S_RXZ] -= dx[k] * runiz;
        v[S_RYY] -= dy[k] * runiy; v[S_RYZ] -= dy[k] * runiz; v[S_RZZ] -= dz[k] * runiz;
        if (MHD) {
            const double dBx = bi.x - bj[k].x, dBy = bi.y - bj[k].y, dBz = bi.z - bj[k].z;   // bi = (B/rho)_i rho(h_i) of this iteration (dens.F90:806-812)
            w[B_DIVB] += dBx * runix + dBy * runiy + dBz * runiz;
            w[B_CURLX] += dBz * runiy - dBy * runiz;          // dBz/dy - dBy/dz
            w[B_CURLY] += dBx * runiz - dBz * runix;          // dBx/dz - dBz/dx
            w[B_CURLZ] += dBy * runix - dBx * runiy;          // dBy/dx - dBx/dy
        }
    }
}

__device__ __forceinline__ void exactlinear_d(double &gx, double &gy, double &gz, double dAx, double dAy, double dAz, const double *rm, double ddenom)
{
    gx = (dAx * rm[0] + dAy * rm[1] + dAz * rm[2]) * ddenom;
    gy = (dAx * rm[1] + dAy * rm[3] + dAz * rm[4]) * ddenom;
    gz = (dAx * rm[2] + dAy * rm[4] + dAz * rm[5]) * ddenom;
}

// Staging block of the density kernel.  Two instantiations of the fast path, picked per call from the candidate counts of the previous
// pass: up to 384 candidates per group (cubic lattices) fit one round WITH their FP64 positions in shared memory; larger candidate sets
// (close-packed lattices ~510, glass-like and centrally condensed distributions ~700) take rounds of 768 without the staged positions,
// because a second round costs more than the staging saves (its hits are lopsided over the lanes and the warp pays the maximum).
#ifndef DENS_ROUND
#define DENS_ROUND ROUND_DEFAULT
#endif
#ifndef DENS_ROUND_BIG
#define DENS_ROUND_BIG 768
#endif
template <bool FAST, bool BIG> struct DensShared { typedef WarpSharedGeneral type; };
template <> struct DensShared<true, false> { typedef WarpSharedT<DENS_ROUND, DENS_STAGE, DENS_STAGE, !DENS_STAGE> type; };
template <> struct DensShared<true, true> { typedef WarpSharedT<DENS_ROUND_BIG, 0, 0> type; };

template <int K, bool PERIODIC, bool MHD, bool GRAV, bool FAST, bool BIG, bool REUSE>
#ifndef DENS_MINB
#define DENS_MINB 3
#endif
#ifndef DENS_MHD_MINB
#define DENS_MHD_MINB 3
#endif
#ifndef DENS_NPAIR
#define DENS_NPAIR 2
#endif
__global__ void __launch_bounds__(128, (FAST && !MHD) ? DENS_MINB : ((FAST && MHD) ? DENS_MHD_MINB : 3)) k_density(const DensArgs a, const __grid_constant__ DevParams dp)
{
    typedef SphKern<K> KF;
    typedef typename DensShared<FAST, BIG>::type WS;
    constexpr bool STAGED = FAST && WS::P2 > 0;
    constexpr bool RU = FAST && REUSE;          // keep the staged round and its masks over the h-rho iterations (see below)
    extern __shared__ __align__(16) unsigned char dens_smem[];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    WS &ws = reinterpret_cast<WS *>(dens_smem)[wib];
    const unsigned ws_s = ws_shared_addr(ws);
    const unsigned hm_lane = ws_s + (unsigned)offsetof(WS, hm) + 4u * lane, sidx_s = ws_s + (unsigned)offsetof(WS, sidx);
    const unsigned rec2_s = ws_s + (unsigned)offsetof(WS, rec2), rec1_s = ws_s + (unsigned)offsetof(WS, rec1);
    const int gwarp = blockIdx.x * 4 + wib;
    int *clist = a.stage_idx + (size_t)gwarp * a.scratch_per_warp;      // cell list of the current group (the only global scratch)
    constexpr int DSTRIDE = MHD ? 8 : 6;                                 // double2 per packed record of the fast path
    const double2 *posrec = reinterpret_cast<const double2 *>(FAST ? a.drec : a.pos4);
    const int pstride = FAST ? DSTRIDE : 2;
    const double Lx = dp.dxbound, Ly = dp.dybound, Lz = dp.dzbound;
    const float fLx = (float)Lx, fLy = (float)Ly, fLz = (float)Lz;
    const double radkern = KF::radkern;
    const double halfLmin = 0.5 * fmin(Lx, fmin(Ly, Lz));
    const bool use_da = dp.nalpha > 1;
    // per-lane statistics (dens.F90:104-106), reduced once at kernel exit
    unsigned long long st_pairs = 0, st_trial = 0, st_ncalc = 0, st_nact = 0, st_np = 0, st_nwalk = 0, st_surv = 0;
    int st_maxact = 0, st_maxtrial = 0;
    double st_rhomax = 0., st_hused = 0., st_hgrow = 0.;

    while (true) {
        int cellid = 0;
        if (lane == 0) cellid = (int)atomicAdd(&a.cnt[CNT_WORK], 1ull);
        cellid = __shfl_sync(FULLMASK, cellid, 0);
        if (cellid >= a.ngroups) break;
        if (a.wl.order) cellid = a.wl.order[cellid];
        const Cell cell = a.groups[cellid];
        if (cell.active == 0) continue;                              // dens.F90:302
        const double cx = 0.5 * (cell.lo[0] + cell.hi[0]), cy = 0.5 * (cell.lo[1] + cell.hi[1]), cz = 0.5 * (cell.lo[2] + cell.hi[2]);
        const double halfext = 0.5 * fmax(cell.hi[0] - cell.lo[0], fmax(cell.hi[1] - cell.lo[1], cell.hi[2] - cell.lo[2]));
        float tlo[3], thi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { tlo[k] = __double2float_rd(cell.lo[k]); thi[k] = __double2float_ru(cell.hi[k]); }
        // ---- lane = target: start_cell (dens.F90:1293-1378)
        const int s = cell.start + min(lane, cell.count - 1);
        bool act = false, gasi = true, dusti = false; int itypei = IGAS;
        if (lane < cell.count) get_partinfo_d(a.stype[s], dp.p.set_boundaries_to_active, dp.p.dust, act, gasi, dusti, itypei);
        const double4 pi = a.pos4[s];
        double4 vi, ai, bi = make_double4(0., 0., 0., 0.);
        double4 bevi = make_double4(0., 0., 0., 0.);            // (B/rho, psi) of the target; the fast pair body takes B = (B/rho) rho(h) of the current iterate
        if (FAST) {
            const double4 *r = a.drec + (MHD ? 4 : 3) * (size_t)s;
            const double4 B = r[1], C = r[2];
            vi = make_double4(B.x, B.y, B.z, 0.); ai = make_double4(B.w, C.x, C.y, 0.);
            if (MHD && gasi) bevi = a.bev4[s];
        }
        else { vi = a.vel4[s]; ai = a.acc4[s]; if (MHD && gasi) bi = a.bev4[s]; }
        const double pmassi = dp.p.massoftype[itypei];
        const float xif = (float)(pi.x - cx), yif = (float)(pi.y - cy), zif = (float)(pi.z - cz);
        double h = pi.w;
        const double h_old = h;
        bool conv = !act;                                            // inactive / boundary lanes take no part (dens.F90:1329)
        bool failed = false;
        double v[29], w[B_COUNT];
        int nneighi = 0, its_lane = 0, nlist_last = 0;

        double hmax_list = cell.hmax * a.margin;
        double rcut_list = radkern * hmax_list;
        bool interior = false;   // no candidate of this group lies across the periodic boundary: the minimum-image branch can be skipped
        // the FP32 filter works on nearest images relative to the cell centre: only valid while the search sphere of every
        // target stays inside half a box length; otherwise every candidate goes to the exact test
        bool wide = PERIODIC && (halfext + rcut_list >= 0.999 * halfLmin);
        float reach = 0.f;
        const int *cl = clist;                                       // list in use: the prepared one, or this warp's slice after a walk in here
        int ncl = a.wl.ncl[cellid];
        if (ncl >= 0) { cl = a.wl.list + (size_t)cellid * a.wl.cap; reach = a.wl.reach[cellid]; }
        else ncl = warp_walk<false, PERIODIC>(a.nodes, a.cells, a.ncells, tlo, thi, __double2float_ru(rcut_list), (float)radkern, fLx, fLy, fLz, ws.walk_stack(),
                                              clist, a.scratch_per_warp, reach);
        st_nwalk += (lane == 0);
        if (ncl < 0) { if (lane == 0) atomicMax(&a.cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); break; }
        if (FAST && PERIODIC) {
            const double rr = rcut_list * 1.0001;
            interior = cell.lo[0] - rr > dp.p.xmin && cell.hi[0] + rr < dp.p.xmax && cell.lo[1] - rr > dp.p.ymin && cell.hi[1] + rr < dp.p.ymax &&
                       cell.lo[2] - rr > dp.p.zmin && cell.hi[2] + rr < dp.p.zmax;
        }

        // fast path, group staged in ONE round: the staged candidates stay in shared memory over the h-rho iterations, and so do the hit
        // masks while every unconverged target's kernel radius stays inside the radius its mask was built for (the exact FP64 test in
        // the pair body decides membership with the current h either way)
        bool staged_all = false; int nr_saved = 0;
        for (int its = 1;; its++) {                                  // local_its (dens.F90:338-373): the cell iterates until every particle converged
            // compute_hmax / redo_neighbours (dens.F90:1275-1289, :343-347)
            const double hneed = warp_max(conv ? 0. : h);
            st_hused = fmax(st_hused, hneed);
            if (radkern * hneed > rcut_list) {
                hmax_list = hneed * a.margin * 1.01;
                rcut_list = radkern * hmax_list;
                wide = PERIODIC && (halfext + rcut_list >= 0.999 * halfLmin);
                ncl = warp_walk<false, PERIODIC>(a.nodes, a.cells, a.ncells, tlo, thi, __double2float_ru(rcut_list), (float)radkern, fLx, fLy, fLz, ws.walk_stack(),
                                                 clist, a.scratch_per_warp, reach);
                cl = clist;
                staged_all = false;
                st_nwalk += (lane == 0);
                if (ncl < 0) { if (lane == 0) atomicMax(&a.cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); failed = true; }
                if (FAST && PERIODIC) {
                    const double rr = rcut_list * 1.0001;
                    interior = cell.lo[0] - rr > dp.p.xmin && cell.hi[0] + rr < dp.p.xmax && cell.lo[1] - rr > dp.p.ymin && cell.hi[1] + rr < dp.p.ymax &&
                               cell.lo[2] - rr > dp.p.zmin && cell.hi[2] + rr < dp.p.zmax;
                }
            }
            if (__any_sync(FULLMASK, failed)) break;
            const bool wrapf = FAST && wide && rcut_list < 0.9 * halfLmin;      // thin periodic box: the filter takes the minimum image per pair (walk.cuh)
            FilterScale fs = filter_scale((float)halfext, reach);
            if (wrapf) filter_scale_wrap(fs);
            // converged / inactive targets get an empty mask; a wide periodic search beyond that switches the filter off
            const double rfilt = RU ? radkern * h * (1. + a.mask_margin) : radkern * h;
            const FilterTarget ft = filter_target(fs, xif, yif, zif, conv ? 0.f : ((wide && !wrapf) ? -1.f : __double2float_ru(rfilt)));
            const bool masks_ok = RU && staged_all && its > 1 && __all_sync(FULLMASK, conv || radkern * h <= (double)ws.rmask[lane]);
            if (!conv) {
#pragma unroll
                for (int k = 0; k < 29; k++) v[k] = 0.;
#pragma unroll
                for (int k = 0; k < B_COUNT; k++) w[k] = 0.;
                nneighi = 0;
                its_lane = its;
            }
            const double hi1 = 1. / h, hi21 = hi1 * hi1;
            if (FAST && MHD) { const double rhoi = rhoh_d(h, pmassi, dp.p.hfact); bi = make_double4(bevi.x * rhoi, bevi.y * rhoi, bevi.z * rhoi, bevi.w); }
            int nlist = 0;
            // General kernel: the candidates are taken one SORT CLASS at a time (cells, target groups and therefore rounds hold one
            // class each), so the kind of pair is the same for the whole warp: same class -> the full sums; gas targets and dust
            // candidates -> rho_dust alone; every other combination enters no density sum (dens.F90:717-741) and is not even staged.
            const int ci = FAST ? 0 : sort_class(a.stype[cell.start]);
            for (int cj = 0; cj < (FAST ? 1 : 3); cj++) {
            if (!FAST) {
                if (!((a.class_mask >> cj) & 1)) continue;
                if (cj != ci && !(ci == 0 && cj == 1 && dp.p.dust)) continue;
            }
            for (int cellpos = 0; cellpos < ncl;) {                  // rounds of <= ROUND candidates staged in shared memory
                auto stage_rec = [&](int slot, int, const double2 &xy, const double2 &zw) {      // fast path: FP64 positions to shared memory
                    if (STAGED) { ws.rec2[0][slot] = xy; ws.rec1[0][slot] = zw.x; }
                };
                int nr;
                bool reuse = false;
                if (RU && staged_all) { nr = nr_saved; cellpos = ncl; reuse = masks_ok; }
                else {
                    const int cp0 = cellpos;
                    nr = stage_round<PERIODIC, false>(ws, cl, ncl, cellpos, posrec, pstride, cx, cy, cz, Lx, Ly, Lz, (float)radkern, a.max_leaf, fs, interior,
                                                      cell.start, stage_rec, FAST ? -1 : cj, a.stype);
                    if (RU) { staged_all = (cp0 == 0 && cellpos >= ncl); nr_saved = nr; }
                }
                if (!FAST && nr == 0) continue;
                nlist += nr;
                unsigned nz;
                if (RU && reuse) nz = conv ? 0u : ws.nzsave[lane];
                else {
                    if (wrapf) { const FilterWrap fw = filter_wrap(fs, (float)dp.dxbound, (float)dp.dybound, (float)dp.dzbound); nz = build_masks<false, true>(ws, nr, ft, &fw); }
                    else nz = build_masks<false>(ws, nr, ft);
                    if (RU) { ws.nzsave[lane] = nz; ws.rmask[lane] = conv ? 0.f : ((wide && !wrapf) ? 3.0e38f : __double2float_rd(rfilt)); }
                }
                int c = -1; unsigned m = 0u;
                if (FAST) {
                    int surv = 0;
                    const int myslot = ws.selfslot[lane];
                    while (true) {      // two neighbours per trip
                        int slot0, slot1;
                        next_hits2(hm_lane, nz, c, m, slot0, slot1);
                        if (slot0 < 0) break;
                        surv += 1 + (slot1 >= 0);
                        dens_pair2_fast<K, PERIODIC, MHD, GRAV, STAGED>(v, w, nneighi, slot0, slot1, myslot, rec2_s, rec1_s, sidx_s, pi.x, pi.y, pi.z, hi1, hi21,
                                                                           vi, ai, bi, a.drec, pmassi, use_da, interior, Lx, Ly, Lz);
                    }
                    st_surv += surv;
                } else if (cj == ci) {
                    while (true) {      // two neighbours per trip: independent dependency chains, loads of both in flight
                        int slot0, slot1;
                        next_hits2(hm_lane, nz, c, m, slot0, slot1);
                        if (slot0 < 0) break;
                        st_surv += 1 + (slot1 >= 0);
                        dens_pair<K, PERIODIC, MHD, GRAV>(v, w, nneighi, slot0, ws.sidx, s, pi, h, hi1, hi21, itypei, gasi, vi, ai, bi, a, dp, use_da, Lx, Ly, Lz);
                        dens_pair<K, PERIODIC, MHD, GRAV>(v, w, nneighi, slot1, ws.sidx, s, pi, h, hi1, hi21, itypei, gasi, vi, ai, bi, a, dp, use_da, Lx, Ly, Lz);
                    }
                } else {
                    while (true) {
                        int slot0, slot1;
                        next_hits2(hm_lane, nz, c, m, slot0, slot1);
                        if (slot0 < 0) break;
                        st_surv += 1 + (slot1 >= 0);
                        dens_pair_rhodust<K, PERIODIC>(v[S_RHODUST], slot0, ws.sidx, s, pi, hi1, hi21, gasi, a, Lx, Ly, Lz);
                        dens_pair_rhodust<K, PERIODIC>(v[S_RHODUST], slot1, ws.sidx, s, pi, hi1, hi21, gasi, a, Lx, Ly, Lz);
                    }
                }
                __syncwarp();
            }
            }
            nlist_last = nlist;
            if (!conv) {
                st_trial += (unsigned long long)nlist;
                // finish_rhosum + finish_cell (dens.F90:1470-1507, :1401-1462)
                const double hi31 = hi1 * hi21, hi41 = hi21 * hi21;
                const double rhoi = KF::cnormk * (v[S_RHO] + KF::wab0 * pmassi) * hi31;
                const double gradhi = KF::cnormk * (v[S_GRADH] + KF::gradh0 * pmassi) * hi41;
                const double rhohi = rhoh_d(h, pmassi, dp.p.hfact);
                const double dhdrhoi = -h / (3. * rhohi);
                const double omegai = 1. - dhdrhoi * gradhi;
                const double func = rhohi - rhoi;
                double dfdh1;
                if (omegai > DBL_MIN) dfdh1 = dhdrhoi / omegai;
                else dfdh1 = dhdrhoi / fabs(omegai + DBL_EPSILON);
                double hnew = h - func * dfdh1;
                if (hnew > 1.2 * h) hnew = 1.2 * h;
                else if (hnew < 0.8 * h) hnew = 0.8 * h;
                conv = ((fabs(hnew - h) / h_old) < dp.p.tolh) && (omegai > 0.) && (h > 0.);
                if (a.icall == 0) conv = true;
                if (!conv) {
                    if (its >= 100) {                               // maxdensits, dens.F90:1443-1456
                        atomicMax(&a.cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_NOCONVERGE);
                        atomicMax(&a.cnt[CNT_ERRID], (unsigned long long)(a.perm[s] + 1));
                        failed = true; conv = true;
                    } else {
                        h = hnew;
                        if (a.h_hist && its <= SPHGPU_HHIST) a.h_hist[(size_t)(its - 1) * a.npart + a.perm[s]] = hnew;
                    }
                }
            }
            if (__all_sync(FULLMASK, conv)) break;
        }
        // ---- store_results (dens.F90:1511-1681), lane = target ----
        if (act && !failed) {
            const double hi1 = 1. / h, hi21 = hi1 * hi1, hi31 = hi1 * hi21, hi41 = hi21 * hi21;
            const double rho = KF::cnormk * (v[S_RHO] + KF::wab0 * pmassi) * hi31;
            double gradhi = KF::cnormk * (v[S_GRADH] + KF::gradh0 * pmassi) * hi41;
            const double rhohi = rhoh_d(h, pmassi, dp.p.hfact);
            const double dhdrhoi = -h / (3. * rhohi);
            const double omegai = 1. - dhdrhoi * gradhi;
            gradhi = 1. / omegai;
            const double hfin = dp.p.hfact * pow(pmassi / fabs(rho), 1.0 / 3.0);    // hrho, part.F90:845
            const size_t io = (size_t)a.perm[s];                                      // results go straight to the caller's arrays
            // the new h too in the fast path (nothing in this pass reads another particle's pos4.w or xyzh; an overflow retry restores
            // the h of build_tree first); the general path reads h_j in its pair loop and keeps the new h aside until the pass is over
            if (FAST) { a.pos4[s].w = hfin; a.xyzh[4 * io + 3] = hfin; }
            else a.hnew[s] = hfin;
            st_hgrow = fmax(st_hgrow, hfin / h_old);
            const float gradh4 = (float)gradhi;
            a.gradh[(size_t)dp.ngradh * io] = gradh4;
            if (GRAV) {
                double gradsofti = (v[S_GRADSOFT] + KF::dphidh0 * pmassi) * hi21;
                gradsofti = gradsofti * dhdrhoi;
                a.gradh[(size_t)dp.ngradh * io + 1] = (float)gradsofti;
            }
            gradhi = (double)gradh4;                                                  // dens.F90:1610
            const double rho1i = 1. / rho;
            if (dp.p.dust) a.dustfrac[io] = gasi ? (KF::cnormk * dp.p.massoftype[IDUST] * v[S_RHODUST] * hi31) * rho1i : 0.;   // dens.F90:1612-1624
            const double term = KF::cnormk * gradhi * rho1i * hi41;
            const double rxx = v[S_RXX], rxy = v[S_RXY], rxz = v[S_RXZ], ryy = v[S_RYY], ryz = v[S_RYZ], rzz = v[S_RZZ];
            const double denom = rxx * ryy * rzz + 2. * rxy * rxz * ryz - rxx * ryz * ryz - ryy * rxz * rxz - rzz * rxy * rxy;
            double rm[6];
            rm[0] = ryy * rzz - ryz * ryz; rm[1] = rxz * ryz - rzz * rxy; rm[2] = rxy * ryz - rxz * ryy;
            rm[3] = rzz * rxx - rxz * rxz; rm[4] = rxy * rxz - rxx * ryz; rm[5] = rxx * ryy - rxy * rxy;
            const double divv = -v[S_DIVV] * term;
            double dv[9], divcurlv5 = 0.;
            if (fabs(denom) > DBL_MIN) {
                const double ddenom = 1. / denom;
                exactlinear_d(dv[0], dv[1], dv[2], v[S_DVXDX], v[S_DVXDY], v[S_DVXDZ], rm, ddenom);
                exactlinear_d(dv[3], dv[4], dv[5], v[S_DVYDX], v[S_DVYDY], v[S_DVYDZ], rm, ddenom);
                exactlinear_d(dv[6], dv[7], dv[8], v[S_DVZDX], v[S_DVZDY], v[S_DVZDZ], rm, ddenom);
#pragma unroll
                for (int k = 0; k < 9; k++) dv[k] = -dv[k];
                if (use_da) {
                    double ax, ay, az, bx, by, bz, cxx, cyy, czz;
                    exactlinear_d(ax, ay, az, v[S_DAXDX], v[S_DAXDY], v[S_DAXDZ], rm, ddenom);
                    exactlinear_d(bx, by, bz, v[S_DAYDX], v[S_DAYDY], v[S_DAYDZ], rm, ddenom);
                    exactlinear_d(cxx, cyy, czz, v[S_DAZDX], v[S_DAZDY], v[S_DAZDZ], rm, ddenom);
                    const double div_a = -(ax + by + czz);
                    divcurlv5 = div_a - (dv[0] * dv[0] + dv[4] * dv[4] + dv[8] * dv[8] + 2. * (dv[1] * dv[3] + dv[2] * dv[6] + dv[5] * dv[7]));
                }
            } else {
                dv[0] = -term * v[S_DVXDX]; dv[1] = -term * v[S_DVXDY]; dv[2] = -term * v[S_DVXDZ];
                dv[3] = -term * v[S_DVYDX]; dv[4] = -term * v[S_DVYDY]; dv[5] = -term * v[S_DVYDZ];
                dv[6] = -term * v[S_DVZDX]; dv[7] = -term * v[S_DVZDY]; dv[8] = -term * v[S_DVZDZ];
                if (use_da) {
                    const double div_a = -term * (v[S_DAXDX] + v[S_DAYDY] + v[S_DAZDZ]);
                    divcurlv5 = div_a - (dv[0] * dv[0] + dv[4] * dv[4] + dv[8] * dv[8] + 2. * (dv[1] * dv[3] + dv[2] * dv[6] + dv[5] * dv[7]));
                }
            }
            a.divcurlv[io] = (float)divv;
            if (dp.nalpha >= 3) a.alphaind[3 * io + 2] = (float)divcurlv5;
#pragma unroll
            for (int k = 0; k < 9; k++) a.dvdx[9 * io + k] = (float)dv[k];
            if (MHD) {
                float *o = a.divcurlB + 4 * io;
                if (gasi) {
                    o[0] = (float)(-w[B_DIVB] * term);
                    o[1] = (float)(-w[B_CURLX] * term);
                    o[2] = (float)(-w[B_CURLY] * term);
                    o[3] = (float)(-w[B_CURLZ] * term);
                } else { o[0] = o[1] = o[2] = o[3] = 0.f; }
            }
            const int nn = nneighi + 1;   // + self
            a.s_nneigh[s] = nn;
            if (a.h_hist) a.h_its[a.perm[s]] = its_lane;
            st_rhomax = fmax(st_rhomax, rho);
            st_pairs += (unsigned long long)nneighi * its_lane;
            st_ncalc += its_lane; st_nact += nn; st_np += 1;
            st_maxact = max(st_maxact, nn); st_maxtrial = max(st_maxtrial, nlist_last);
        }
        __syncwarp();
    }
    // ---- statistics: one atomic per warp
    st_rhomax = warp_max(st_rhomax); st_hused = warp_max(st_hused); st_hgrow = warp_max(st_hgrow);
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) {
        st_pairs += __shfl_xor_sync(FULLMASK, st_pairs, sft); st_trial += __shfl_xor_sync(FULLMASK, st_trial, sft);
        st_ncalc += __shfl_xor_sync(FULLMASK, st_ncalc, sft); st_nact += __shfl_xor_sync(FULLMASK, st_nact, sft);
        st_np += __shfl_xor_sync(FULLMASK, st_np, sft); st_nwalk += __shfl_xor_sync(FULLMASK, st_nwalk, sft);
        st_surv += __shfl_xor_sync(FULLMASK, st_surv, sft);
        st_maxact = max(st_maxact, __shfl_xor_sync(FULLMASK, st_maxact, sft)); st_maxtrial = max(st_maxtrial, __shfl_xor_sync(FULLMASK, st_maxtrial, sft));
    }
    if (lane == 0) {
        atomicAdd(&a.cnt[CNT_NPAIRS], st_pairs); atomicAdd(&a.cnt[CNT_NTRIAL], st_trial); atomicAdd(&a.cnt[CNT_NCALC], st_ncalc);
        atomicAdd(&a.cnt[CNT_NACT], st_nact); atomicAdd(&a.cnt[CNT_NP], st_np); atomicAdd(&a.cnt[CNT_NWALK], st_nwalk);
        atomicAdd(&a.cnt[CNT_NSURV], st_surv);
        atomicMax(&a.cnt[CNT_MAXACT], (unsigned long long)st_maxact); atomicMax(&a.cnt[CNT_MAXTRIAL], (unsigned long long)st_maxtrial);
        atomic_max_pos(&a.dscal[DS_RHOMAX], st_rhomax); atomic_max_pos(&a.dscal[DS_HUSED], st_hused); atomic_max_pos(&a.dscal[DS_HGROW], st_hgrow);
    }
}

// grid < 0: only query the resident CTAs/SM of the instantiation; otherwise launch on `grid` CTAs
template <int K, bool PERIODIC, bool MHD, bool GRAV, bool FAST, bool BIG, bool REUSE>
int launch_density2(sphgpu_ctx *c, const DensArgs &a, int grid)
{
    const size_t smem = 4 * sizeof(typename DensShared<FAST, BIG>::type);
    cudaFuncSetAttribute(k_density<K, PERIODIC, MHD, GRAV, FAST, BIG, REUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (grid < 0) {
        int bps = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_density<K, PERIODIC, MHD, GRAV, FAST, BIG, REUSE>, 128, smem);
        return bps < 1 ? 1 : bps;
    }
    k_density<K, PERIODIC, MHD, GRAV, FAST, BIG, REUSE><<<grid, 128, smem, c->stream>>>(a, c->hp);
    c->launches++;
    return 0;
}
template <int K, bool PERIODIC, bool MHD, bool GRAV, bool FAST>
int launch_density(sphgpu_ctx *c, const DensArgs &a, int grid)
{
    // the previous pass staged more candidates per group than one small round holds: take the big rounds
    // (and a set that iterated last time -- a.mask_margin > 0 -- takes the instantiation that keeps round and masks over the iterations;
    // it exists with the big rounds only: disordered sets are the ones that iterate, and the extra state costs the lattice kernel spills)
    if (FAST && c->dens_trial_max > DENS_ROUND && c->dens_trial_hint > 0.8 * DENS_ROUND) {
        if (c->dens_reuse) return launch_density2<K, PERIODIC, MHD, GRAV, FAST, FAST, FAST>(c, a, grid);
        return launch_density2<K, PERIODIC, MHD, GRAV, FAST, FAST, false>(c, a, grid);
    }
    return launch_density2<K, PERIODIC, MHD, GRAV, FAST, false, false>(c, a, grid);
}

template <int K, bool PERIODIC, bool FAST>
int dispatch_density3(sphgpu_ctx *c, const DensArgs &a, int grid)
{
    const bool mhd = c->hp.p.mhd, grav = c->hp.p.gravity;
    if (mhd && grav) return launch_density<K, PERIODIC, true, true, FAST>(c, a, grid);
    if (mhd) return launch_density<K, PERIODIC, true, false, FAST>(c, a, grid);
    if (grav) return launch_density<K, PERIODIC, false, true, FAST>(c, a, grid);
    return launch_density<K, PERIODIC, false, false, FAST>(c, a, grid);
}

// general path: several particle types (boundary, dust) in the set
bool density_is_general(const sphgpu_ctx *c) { return c->multitype || c->hp.p.dust || c->force_general; }

int dispatch_density(sphgpu_ctx *c, const DensArgs &a, int grid)
{
    const sphgpu_params &p = c->hp.p;
    const bool fast = !density_is_general(c);
    if (p.kernel == 0) {
        if (p.periodic) return fast ? dispatch_density3<0, true, true>(c, a, grid) : dispatch_density3<0, true, false>(c, a, grid);
        return fast ? dispatch_density3<0, false, true>(c, a, grid) : dispatch_density3<0, false, false>(c, a, grid);
    }
    if (p.periodic) return fast ? dispatch_density3<1, true, true>(c, a, grid) : dispatch_density3<1, true, false>(c, a, grid);
    return fast ? dispatch_density3<1, false, true>(c, a, grid) : dispatch_density3<1, false, false>(c, a, grid);
}

}  // namespace

static inline int nblk(int64_t n, int b) { return (int)((n + b - 1) / b); }

int density_run(sphgpu_ctx *c, int icall, sphgpu_scalars *out)
{
    if (!c->tree_valid) { c->err = "densityiterate: build_tree has not been called"; return SPHGPU_ERR_STATE; }
    const int64_t n = c->npart, nl = c->nlive;
    const sphgpu_params &p = c->hp.p;
    CUDA_TRY(c, c->vel4.ensure(n)); CUDA_TRY(c, c->acc4.ensure(n)); if (p.mhd) CUDA_TRY(c, c->bev4.ensure(n));
    CUDA_TRY(c, c->hnew.ensure(n)); CUDA_TRY(c, c->s_nneigh.ensure(n));
    DensArgs a;
    memset(&a, 0, sizeof a);
    c->dens_reuse = c->last_dens.np > 0 && (double)c->last_dens.nrhocalc > 1.2 * (double)c->last_dens.np;
    const int grid = c->numSMs * dispatch_density(c, a, -1);     // persistent grid = resident CTAs/SM x SMs
    const bool fast = !density_is_general(c);
    if (fast) CUDA_TRY(c, c->drec.ensure(4 * (size_t)n));
    CUDA_TRY(c, c->stage_idx.ensure((size_t)grid * 4 * c->scratch_per_warp));
    k_gather_dens<<<nblk(nl, 256), 256, 0, c->stream>>>(nl, c->perm.p, c->vxyzu.p, c->fxyzu.p, c->fext.p, c->Bevol.p, c->hp.nvu, p.mhd, c->pos4.p, c->vel4.p,
                                                        c->acc4.p, c->bev4.p, c->hnew.p, c->s_nneigh.p, fast ? c->drec.p : nullptr, p.massoftype[IGAS], p.hfact);
    c->launches++;
    a.drec = c->drec.p;
    a.nodes = c->nodesf.p; a.cells = c->cells.p; a.ncells = (int)c->ncells; a.groups = c->groups.p; a.ngroups = (int)c->ngroups;
    a.pos4 = c->pos4.p; a.vel4 = c->vel4.p; a.acc4 = c->acc4.p; a.bev4 = c->bev4.p; a.stype = c->stype.p; a.perm = c->perm.p;
    a.hnew = c->hnew.p; a.s_nneigh = c->s_nneigh.p; a.xyzh = c->xyzh.p;
    a.gradh = c->gradh.p; a.divcurlv = c->divcurlv.p; a.dvdx = c->dvdx.p; a.alphaind = c->alphaind.p; a.divcurlB = c->divcurlB.p; a.dustfrac = c->dustfrac.p;
    a.multitype = c->multitype ? 1 : 0; a.max_leaf = c->max_leaf; a.class_mask = c->class_mask; a.hmax_global = 0.;
    a.cnt = c->counters.p; a.dscal = c->dscal.p;
    a.margin = c->list_margin; a.icall = icall;
    // a set that needed a second h-rho iteration last time will need one again: build the masks a little wide and keep them
    a.mask_margin = c->dens_reuse ? 0.005 : 0.;
    // the node-hmax replay of the reference tree (self-gravity, and the reference-compatible neighbour mode) needs every particle's h history
    const bool loghist = p.gravity || refcompat_on(c);
    if (loghist) CUDA_TRY(c, c->h_hist.ensure((size_t)SPHGPU_HHIST * n));
    a.h_hist = loghist ? c->h_hist.p : nullptr; a.h_its = c->h_its.p; a.npart = n;
    c->ref_valid = false;
    c->grav_tree_valid = false;            // h changes below: the gravity tree caches h
    {   // ONE symmetric walk serves both passes: with the list margin on every radius it holds the density candidates now and the force
        // candidates as long as no h grows by more than the margin (force_run checks hscale against wl_cover)
        const double R = (p.kernel == 0 ? SphKern<0>::radkern : SphKern<1>::radkern), hs0 = fmax(c->hscale, 1.);
        TRY(walk_lists_run(c, true, R * c->list_margin * hs0, R * c->list_margin * hs0, a.wl));
        TRY(walk_order_run(c, a.wl));
        c->wl_force_ok = true; c->wl_cover = hs0 * c->list_margin; c->wl_ordered = a.wl.order != nullptr;
    }
    unsigned long long hc[16]; double hrhomax, hused, hgrow = 0.;
    for (int attempt = 0;; attempt++) {
        CUDA_TRY(c, cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 16, c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->dscal.p + DS_RHOMAX, 0, 3 * sizeof(double), c->stream));
        a.stage_idx = c->stage_idx.p; a.scratch_per_warp = c->scratch_per_warp;
        cudaEventRecord(c->ev[8], c->stream);
        dispatch_density(c, a, grid);
        cudaEventRecord(c->ev[9], c->stream);
        CUDA_TRY(c, cudaMemcpyAsync(hc, c->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(&hrhomax, c->dscal.p + DS_RHOMAX, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(&hused, c->dscal.p + DS_HUSED, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(&hgrow, c->dscal.p + DS_HGROW, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaGetLastError());
        // a group whose search sphere covers more leaf cells than a warp's cell list holds (strongly non-uniform h): grow the lists and
        // repeat the pass -- nothing has been scattered to the caller's arrays yet
        if (hc[CNT_ERR] == SPHGPU_ERR_OVERFLOW && attempt < 3 && c->scratch_per_warp < (1 << 20)) {
            c->scratch_per_warp *= 8;
            CUDA_TRY(c, c->stage_idx.ensure((size_t)grid * 4 * c->scratch_per_warp));
            if (fast) {                        // groups that finished have already stored their new h: start again from the tree's h
                k_restore_h_sorted<<<nblk(nl, 256), 256, 0, c->stream>>>(nl, c->perm.p, c->h_build.p, c->pos4.p, c->xyzh.p);
                c->launches++;
            }
            continue;
        }
        break;
    }
    { float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[8], c->ev[9]); c->ms_kernel[0] = ms; }
    if (hc[CNT_ERR] == SPHGPU_ERR_NOCONVERGE) {
        char buf[160]; snprintf(buf, sizeof buf, "densityiterate: could not converge in density on particle %llu", hc[CNT_ERRID]);
        c->err = buf; return SPHGPU_ERR_NOCONVERGE;
    }
    if (hc[CNT_ERR]) { c->err = "densityiterate: neighbour scratch overflow (raise scratch_per_warp)"; return (int)hc[CNT_ERR]; }
    if (!fast) {
        k_scatter_h<<<nblk(nl, 256), 256, 0, c->stream>>>(nl, c->perm.p, c->s_nneigh.p, c->hnew.p, c->pos4.p, c->xyzh.p);
        c->launches++;
    }
    sphgpu_scalars &sc = c->last_dens;
    memset(&sc, 0, sizeof sc);
    sc.rhomax = hrhomax; sc.np = (int64_t)hc[CNT_NP];
    c->dens_hmax_used = hused; c->dens_hgrow = hgrow;
    // set_hmaxcell (neigh_kdtree.f90:115-131): the force walk needs hmax >= the new h.  When no h grew by more than 2 % the tree's
    // hmax are inflated by that factor instead of refitted (saves the bottom-up pass); otherwise refit.
    if (hgrow <= 1.02 && !c->always_refit) c->hscale = fmax(c->hscale, 1.) * fmax(hgrow, 1.) * (1. + 1e-12);
    else TRY(tree_refit_hmax(c));
    sc.trialmean = sc.np ? (double)hc[CNT_NTRIAL] / (double)hc[CNT_NCALC] : -1.;
    if (sc.np) { c->dens_trial_hint = sc.trialmean; c->dens_trial_max = (double)hc[CNT_MAXTRIAL]; }
    sc.actualmean = sc.np ? (double)hc[CNT_NACT] / (double)sc.np : -1.;
    sc.maxtrial = (int64_t)hc[CNT_MAXTRIAL]; sc.maxactual = (int64_t)hc[CNT_MAXACT]; sc.nrhocalc = (int64_t)hc[CNT_NCALC];
    sc.nactualtot = (int64_t)hc[CNT_NACT]; sc.ncalls_neigh = (int64_t)hc[CNT_NWALK]; sc.npairs_density = (int64_t)hc[CNT_NPAIRS];
    if (out) *out = sc;
    c->dens_valid = true;
    return SPHGPU_OK;
}
