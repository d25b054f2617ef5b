// force.cu -- force pass: replaces force (src/main/force.F90:193-905) with start_cell/get_stress (:2172-2514, :2068-2168),
// compute_forces (:914-2060) and finish_cell_and_store_results (:2649-3330), for hydro + artificial viscosity
// (Cullen-Dehnen alpha per particle or constant, optional disc viscosity) + artificial conductivity + ideal MHD with
// artificial resistivity and hyperbolic/parabolic div-B cleaning; and, in the XTRA instantiation, the softened
// self-gravity of SPH-neighbour pairs (:1303-1339; the far field comes from gravity.cu), two-fluid gas-dust drag
// (:1852-1989, get_ts of dust.f90:161-276, reconstruct_dv :3338-3372) and individual timestep bins (:1346-1358, :3272-3310).
//
// Same warp-per-leaf-cell skeleton as the density pass (walk.cuh) with the symmetric neighbour criterion
// q2i < R^2 .or. q2j < R^2 (force.F90:1287).  What the reference recomputes per pair for particle j
// (rho_j, get_stress -> P_j/rho_j^2, v_wave,j, the gradW prefactor h_j^-4 cnormk gradh_j: force.F90:1455-1504, :1327)
// is computed once per particle by k_force_prep into 32-byte records so that a pair costs four sector gathers.
#include "walk.cuh"
#include "sphkern.cuh"
#include <float.h>
#include <string.h>

namespace {

struct ForceArgs {
    const TreeNodeF *nodes; const Cell *cells; int ncells; const Cell *groups; int ngroups;
    int *stage_idx; int multitype; int max_leaf; int class_mask;
    int iso1; double iso_pmass, iso_hfact, iso_polyk, iso_cs, iso_cnormk, iso_alpha;   // two-sector records of the isothermal fast path (iso1_derive)
    WalkLists wl;       // cell lists prepared by k_walk_lists
    const double4 *pos4, *vel4, *recC, *recD, *recE; const double2 *hinv; const int8_t *stype; const int *perm;
    double4 *s_fxyzu, *s_dB; float *s_divvf, *s_divBsymm; int *s_done;
    int scratch_per_warp; unsigned long long *cnt; double *dscal;
    int icall;
    double hscale;           // >= largest growth of any h since the tree's hmax were last refitted (1 after a refit): the walk inflates hmax by it
    // XTRA (gravity / dust / individual timesteps) only
    const double4 *frec;     // fast path: 5 x 32 B per particle {x,y,z,1/h} {v,gradW factor} {P/rho^2.., v_wave, alpha v_wave, 1/rho} {P,u,cs,alpha} {B,psi}
    const double *gsoft; const float *dvdx9; const double4 *gacc; float *s_poten; double *s_tstop;
    const int8_t *s_ibinold, *s_ibin; int *s_wake; int8_t *s_ibinnew;
    int nbinmax, ibinnow_m1, istepfrac;
    // reference-compatible neighbour mode (common.cuh: refcompat): reference tree nodes, leaf of every sorted slot; refnodes == NULL: exact mode
    const RefNode *refnodes; const int *refleaf; double ref_radkern, ref_tree_acc2; int ref_gravity;
};

struct XtraSums { double fdx, fdy, fdz, tsmin; int ibin_neigh; };

enum { A_FX = 0, A_FY, A_FZ, A_DRHODT, A_DUDTDISS, A_DENDTDISS, A_DIVBSYM, A_DBX, A_DBY, A_DBZ, A_DIVBDIFF, A_POT };

// one thread per sorted particle: the per-particle part of start_cell + get_stress
__global__ void k_force_prep(int64_t nlive, const int *__restrict__ perm, const double4 *__restrict__ pos4, const int8_t *__restrict__ stype,
                             const double *__restrict__ vxyzu, const double *__restrict__ Bevol, const double *__restrict__ eos_vars,
                             const float *__restrict__ alphaind, const float *__restrict__ gradh, double4 *__restrict__ vel4, double4 *__restrict__ recC,
                             double4 *__restrict__ recD, double4 *__restrict__ recE, double2 *__restrict__ hinv, int *__restrict__ s_done,
                             const __grid_constant__ DevParams dp, unsigned long long *cnt, double *__restrict__ gsoft, const float *__restrict__ dvdx,
                             float *__restrict__ dvdx9, const int8_t *__restrict__ ibin, const int8_t *__restrict__ ibin_old,
                             const int8_t *__restrict__ ibin_wake, int8_t *__restrict__ s_ibin, int8_t *__restrict__ s_ibinold, int *__restrict__ s_wake,
                             double4 *__restrict__ frec, int iso1)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    const sphgpu_params &p = dp.p;
    const int i = perm[s];
    const double h = pos4[s].w;
    bool act, gas, dust; int itype;
    get_partinfo_d(stype[s], p.set_boundaries_to_active, p.dust, act, gas, dust, itype);
    const int itypej = abs((int)stype[s]);                       // neighbours are looked up by raw type (force.F90:1363)
    const double pmass = p.massoftype[itypej];
    if (h < 0.) { atomicMax(&cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_NEGH); atomicMax(&cnt[CNT_ERRID], (unsigned long long)(i + 1)); }
    const double h1 = 1. / fabs(h);                              // rhoanddhdrho (part.F90:805-817)
    const double h21 = h1 * h1;
    const double hf = p.hfact * h1;
    const double rho = pmass * (hf * hf * hf);
    const double rho1 = 1. / rho;
    const double *v = vxyzu + (size_t)dp.nvu * i;
    if (!frec) {                                                 // separate records: general kernel only
        vel4[s] = make_double4(v[0], v[1], v[2], dp.nvu >= 4 ? v[3] : 0.);
        hinv[s] = make_double2(h1, h21);
    }
    double pro2 = 0., vwave = 0., alpha = 0., pr = 0., cs = 0., gradhfac = 0.;
    double4 E = make_double4(0., 0., 0., 0.);
    const double gh = (double)gradh[(size_t)dp.ngradh * i];
    gradhfac = h21 * h21 * dp.kc.cnormk * (gh > 0. ? gh : 1.);   // force.F90:2616-2621 resets a zero gradh to 1
    if (gas) {
        pr = eos_vars[7 * (size_t)i]; cs = eos_vars[7 * (size_t)i + 1];
        alpha = p.const_av ? p.alpha : (double)alphaind[3 * (size_t)i];
        if (p.mhd) {                                             // get_stress (force.F90:2132-2155)
            const double4 B = reinterpret_cast<const double4 *>(Bevol)[i];
            E = make_double4(B.x * rho, B.y * rho, B.z * rho, B.w);
            const double bx = E.x * rho1, by = E.y * rho1, bz = E.z * rho1;
            const double Bro2 = bx * bx + by * by + bz * bz;
            vwave = sqrt(cs * cs + Bro2 * rho);
            pro2 = pr * rho1 * rho1 + 0.5 * Bro2;
        } else { pro2 = pr * rho1 * rho1; vwave = cs; }
    }
    if (!frec) {
        recC[s] = make_double4(pro2, vwave, alpha, pr);
        recD[s] = make_double4(rho1, gradhfac, pmass, cs);
        if (p.mhd) recE[s] = E;
    }
    s_done[s] = 0;
    if (frec) {                                                  // packed j-records of the all-gas fast path
        const double4 x = pos4[s];
        const int fstride = iso1 ? 2 : (p.mhd ? 5 : ((dp.nvu >= 4 || p.gravity) ? 4 : 3));
        double4 *r = frec + fstride * (size_t)s;
        r[0] = make_double4(x.x, x.y, x.z, h1);
        if (iso1) {          // gradh and alpha as the real*4 the reference stores; everything else of the record follows from h (iso1_derive)
            const float ghf = gradh[(size_t)dp.ngradh * i];
            const float alf = p.const_av ? 0.f : alphaind[3 * (size_t)i];
            r[1] = make_double4(v[0], v[1], v[2], __hiloint2double(__float_as_int(alf), __float_as_int(ghf > 0.f ? ghf : 1.f)));
        } else {
            r[1] = make_double4(v[0], v[1], v[2], gradhfac);
            r[2] = make_double4(pro2, vwave, alpha * vwave, rho1);
            if (fstride >= 4) r[3] = make_double4(pr, dp.nvu >= 4 ? v[3] : 0., p.gravity ? (double)gradh[(size_t)dp.ngradh * i + 1] : cs, alpha);
            if (fstride >= 5) r[4] = E;
        }
    }
    if (p.gravity) gsoft[s] = (double)gradh[(size_t)dp.ngradh * i + 1];
    if (p.dust) {
#pragma unroll
        for (int k = 0; k < 9; k++) dvdx9[9 * (size_t)s + k] = dvdx[9 * (size_t)i + k];
    }
    if (p.ind_timesteps) { s_ibin[s] = ibin[i]; s_ibinold[s] = ibin_old[i]; s_wake[s] = (int)ibin_wake[i]; }
}

// 1/x for a normal x > 0: two Newton steps on the hardware seed, no special-case branch
__device__ __forceinline__ double rcp_pos(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// get_ts (dust.f90:161-276): stopping time of a gas-dust pair
// (not inlined: the Stokes branch carries a pow(); one copy keeps the general kernel's pair loops inside the instruction cache)
__device__ __noinline__ double get_ts_d(const sphgpu_params &p, double rhogas, double rhodust, double spsoundgas, double dv2)
{
    const double pi = 3.14159265358979323846264338327950288;
    const double rhosum = rhogas + rhodust;
    const double sgrain = p.grainsize, densgrain = p.graindens;
    if (p.idrag == 1) {
        const double cste_mu = sqrt(2. / (pi * p.gamma)), coeff_gei_1 = sqrt(8. / (pi * p.gamma));
        double dragcoeff, f;
        const double lambda = p.seff / rhogas;
        const double kn_eff = (sgrain > 0.) ? 9. * lambda / (4. * sgrain) : DBL_MAX;
        if (kn_eff >= 1.) {                               // Epstein, with the Kwok (1975) supersonic correction
            dragcoeff = (densgrain > DBL_MIN) ? coeff_gei_1 * spsoundgas / (densgrain * sgrain) : DBL_MAX;
            f = (spsoundgas > 0. && dv2 > 0.) ? sqrt(1. + 9. * pi / 128. * dv2 / (spsoundgas * spsoundgas)) : 1.;
        } else {                                          // Stokes, three Reynolds-number regimes
            const double viscmol_nu = cste_mu * lambda * spsoundgas;
            const double abs_dv = sqrt(dv2);
            const double Re_dust = 2. * sgrain * abs_dv / viscmol_nu;
            if (Re_dust <= 1.) { dragcoeff = 4.5 * viscmol_nu / (densgrain * sgrain * sgrain); f = 1.; }
            else if (Re_dust <= 800.) { dragcoeff = 9. / (densgrain * sgrain * pow(Re_dust, 0.6)); f = abs_dv; }
            else { dragcoeff = 0.163075 / (densgrain * sgrain); f = abs_dv; }
        }
        const double ts1 = (dragcoeff == DBL_MAX) ? DBL_MAX : dragcoeff * f * rhosum;
        return (ts1 > 0.) ? 1. / ts1 : DBL_MAX;
    }
    if (p.idrag == 2) return (p.K_code > 0.) ? rhogas * rhodust / (p.K_code * rhosum) : DBL_MAX;
    if (p.idrag == 3) return p.K_code;
    return 0.;
}

// slope of the velocity along the pair direction (reconstruct_dv, force.F90:3348-3354)
__device__ __forceinline__ double recon_slope(const float *__restrict__ d, double dx, double dy, double dz, double rx, double ry, double rz)
{
    return dx * (rx * (double)d[0] + ry * (double)d[3] + rz * (double)d[6]) + dy * (rx * (double)d[1] + ry * (double)d[4] + rz * (double)d[7]) +
           dz * (rx * (double)d[2] + ry * (double)d[5] + rz * (double)d[8]);
}

// Would the reference's force walk from the leaf of i reach the leaf of j?  getneigh (kdtree.F90:1221-1347) opens a node when
// r2 < (size_i + size_n + max(rcut_i, radkern hmax_n))^2 between the node centres (minimum image, get_sep :1468-1501), or by the
// gravity opening criterion; j's leaf is reached iff every node on the path from the root down to it is opened.
template <bool PERIODIC>
__device__ __noinline__ bool ref_walk_reaches(const ForceArgs &a, int si, int sj, double Lx, double Ly, double Lz)
{
    int n = a.refleaf[sj];
    if (n < 0) return true;                                  // j's leaf hmax covers h_j (k_ref_leaf_sorted): every node on its path opens
    const int li = a.refleaf[si];
    const RefNode ci = a.refnodes[li < 0 ? ~li : li];
    const double rcuti = a.ref_radkern * ci.hmax;
    while (n >= 0) {
        const RefNode nd = a.refnodes[n];
        double dx = ci.xcen[0] - nd.xcen[0], dy = ci.xcen[1] - nd.xcen[1], dz = ci.xcen[2] - nd.xcen[2];
        if (PERIODIC) {
            if (fabs(dx) > 0.5 * Lx) dx -= copysign(Lx, dx);
            if (fabs(dy) > 0.5 * Ly) dy -= copysign(Ly, dy);
            if (fabs(dz) > 0.5 * Lz) dz -= copysign(Lz, dz);
        }
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double rcut = fmax(rcuti, a.ref_radkern * nd.hmax);
        const double s = ci.size + nd.size;
        const bool open = (r2 < (s + rcut) * (s + rcut)) || (a.ref_gravity && a.ref_tree_acc2 * r2 < s * s);
        if (!open) return false;
        n = nd.parent;
    }
    return true;
}

// pair body: lane = target, j = this lane's next neighbour candidate (slot < 0: none).  Branch-free: the exact membership
// test (force.F90:1271-1287, :1230) and the gas-gas condition (:1539) become zero weights on grad W_i, grad W_j, through
// which every sum of compute_forces scales; two calls per trip give two independent FP64 dependency chains.
// CP = kind of pair, the same for the whole warp because target groups and staged rounds hold one sort class each (tree.cu):
// 0 gas-gas (hydro, MHD, conductivity), 1 gas-dust or dust-gas (drag), 2 anything else; gravity, the wake-up of individual-timestep
// neighbours, the pair count and the signal speed are common to all three.
template <int K, bool PERIODIC, bool MHD, bool XTRA, int CP>
__device__ __forceinline__ void force_pair(double (&f)[12], double &vsigmax, int &npair, int slot, const int *__restrict__ idxlist, int s,
                                           const double4 &pi, double hi, double hi1, double hi21, bool gasi, const double4 &vi, const double4 &Ci,
                                           const double4 &Di, const double4 &Ei, const ForceArgs &a, const DevParams &dp, double Lx, double Ly, double Lz,
                                           XtraSums &xs, int itypei)
{
    typedef SphKern<K> KF;
    const sphgpu_params &p = dp.p;
    const int j = (slot >= 0) ? idxlist[slot] : s;
    const double4 pj = ldg256(a.pos4 + j);
    const double2 hj = a.hinv[j];
    const double4 Dj = ldg256(a.recD + j);
    double4 Cj = make_double4(0., 0., 0., 0.);
    if (CP == 0) Cj = ldg256(a.recC + j);
    const double4 vj = ldg256(a.vel4 + j);
    double dx, dy, dz;
    const double r2 = pair_r2<PERIODIC>(pi.x, pi.y, pi.z, pj, Lx, Ly, Lz, dx, dy, dz);
    const double hj1 = hj.x, hj21 = hj.y;
    const double q2i = __dmul_rn(r2, hi21), q2j = __dmul_rn(r2, hj21);        // force.F90:1272, :1285
    bool isn = (q2i < KF::radkern2 || q2j < KF::radkern2) && (j != s);        // :1287, :1230
    // a pair that only j's kernel reaches exists for the reference only if its walk got to j's leaf (see ref_walk_reaches)
    if (XTRA && a.refnodes && isn && !(q2i < KF::radkern2)) isn = ref_walk_reaches<PERIODIC>(a, s, j, Lx, Ly, Lz);
    npair += isn ? 1 : 0;
    int itypej = IGAS;
    if (a.multitype) itypej = abs((int)a.stype[j]);
    const bool gg = isn && (CP == 0);                                         // gas-gas condition (force.F90:1539)
    const double r2s = isn ? r2 : 1.0;
    const double rij1 = rsqrt_pos(r2s);                                        // force.F90:1293-1299 (0 for a vanishing separation)
    const double rij = r2s * rij1;
    const double qi = rij * hi1, qj = rij * hj1;
    const double pmassj = Dj.z, pmassi = Di.z;
    double grkerni, grkernj;
    if (XTRA) {
        // kernel gradients for every pair type (gravity and drag use them), masked to gas-gas for the hydro sums below
        const bool ki = isn && (q2i < KF::radkern2), kj = isn && (q2j < KF::radkern2);
        const double gri = ki ? KF::grkern(q2i, qi) * Di.y : 0., grj = kj ? KF::grkern(q2j, qj) * Dj.y : 0.;
        if (p.ind_timesteps && isn && itypej != IBOUNDARY) {                  // force.F90:1346-1358
            if (a.s_wake[j] < a.ibinnow_m1) atomicMax(&a.s_wake[j], a.ibinnow_m1);
            xs.ibin_neigh = max(xs.ibin_neigh, (int)a.s_ibinold[j]);
        }
        if (p.gravity && isn) {                                               // force.F90:1303-1339, :1522
            double phii, fgravi, fgravj;
            if (ki) { double fmi; KF::softening(q2i, qi, phii, fmi); phii *= hi1; fgravi = fmi * hi21 + a.gsoft[s] * gri; }
            else { phii = -rij1; fgravi = rij1 * rij1; }
            if (kj) { double phij, fmj; KF::softening(q2j, qj, phij, fmj); fgravj = fmj * hj21 + a.gsoft[j] * grj; }
            else fgravj = rij1 * rij1;
            const double fgrav = 0.5 * (Dj.z * fgravi + Di.z * fgravj);
            f[A_FX] -= dx * rij1 * fgrav; f[A_FY] -= dy * rij1 * fgrav; f[A_FZ] -= dz * rij1 * fgrav;
            f[A_POT] += Dj.z * phii;
        }
        if (CP == 1 && p.dust && p.idrag > 0 && isn) {                        // force.F90:1864-1970 (explicit drag, large grains)
            const bool gas_dust = gasi;                                       // else dust-gas: the warp's targets are of one class
            {
                const double rx = dx * rij1, ry = dy * rij1, rz = dz * rij1;
                const double pv = (vi.x - vj.x) * rx + (vi.y - vj.y) * ry + (vi.z - vj.z) * rz;
                const double sl = recon_slope(a.dvdx9 + 9 * (size_t)s, dx, dy, dz, rx, ry, rz);
                const double sr = recon_slope(a.dvdx9 + 9 * (size_t)j, dx, dy, dz, rx, ry, rz);
                double slope = 0.;                                            // Van Leer MC limiter (force.F90:3419-3438), irecon = 1
                if (sl * sr > 0.) slope = copysign(1.0, sl) * fmin(fmin(fabs(0.5 * (sl + sr)), 2. * fabs(sl)), 2. * fabs(sr));
                const double projvstar = pv - slope;
                const double dv2 = projvstar * projvstar;
                const double wdrag = (q2i < q2j) ? KF::wdrag(q2i, qi) * hi21 * hi1 * KF::cnormk_drag : KF::wdrag(q2j, qj) * hj21 * hj1 * KF::cnormk_drag;
                const double rhoi = 1. / Di.x, rhoj = 1. / Dj.x;
                const double ts = get_ts_d(p, gas_dust ? rhoi : rhoj, gas_dust ? rhoj : rhoi, gas_dust ? Di.w : Dj.w, dv2);
                const double dragterm = 3. * Dj.z / ((rhoi + rhoj) * ts) * projvstar * wdrag;
                xs.tsmin = fmin(xs.tsmin, ts);
                xs.fdx -= dragterm * rx; xs.fdy -= dragterm * ry; xs.fdz -= dragterm * rz;
                if (gas_dust && dp.nvu >= 4) f[A_DUDTDISS] += dragterm * pv;
            }
        }
        grkerni = gg ? gri : 0.; grkernj = gg ? grj : 0.;
    } else {
        const bool ini = gg && (q2i < KF::radkern2), inj = gg && (q2j < KF::radkern2);
        grkerni = ini ? KF::grkern(q2i, qi) * Di.y : 0.;                      // :1301-1302
        grkernj = inj ? KF::grkern(q2j, qj) * Dj.y : 0.;                      // :1325-1327
    }
    const double runix = dx * rij1, runiy = dy * rij1, runiz = dz * rij1;
    const double dvx = vi.x - vj.x, dvy = vi.y - vj.y, dvz = vi.z - vj.z;
    const double projv = dvx * runix + dvy * runiy + dvz * runiz;
    if (CP != 0) {                                                            // force.F90:1446-1448: signal speed of a pair that is not gas-gas
        vsigmax = fmax(vsigmax, isn ? fmax(-projv, 0.) : 0.);
        return;
    }
    bool usej = (q2j < KF::radkern2);
    if (MHD) usej = true;
    if (p.dust) usej = true;
    if (dp.nvu >= 4 && !p.gravity) usej = true;                               // :1343-1345
    const double rho1i = Di.x, rho1j = usej ? Dj.x : 0.;
    const double vwavei = Ci.y, alphai = Ci.z, pri = Ci.w, pro2i = Ci.x;
    const double beta = p.beta;
    const double vsigi = fmax(vwavei - beta * projv, 0.);                     // :1423-1426
    const double vsigavi = fmax(alphai * vwavei - beta * projv, 0.);
    const double vwavej = usej ? Cj.y : 0., pro2j = usej ? Cj.x : 0., prj = usej ? Cj.w : 0.;
    const double alphaj = (usej && !p.const_av) ? Cj.z : alphai;
    const double vsigj = usej ? fmax(vwavej - beta * projv, 0.) : 0.;         // :1501-1504
    const double vsigavj = usej ? fmax(alphaj * vwavej - beta * projv, 0.) : 0.;
    vsigmax = fmax(vsigmax, isn ? fmax(vsigi, vsigj) : 0.);
    double qrho2i = 0., qrho2j = 0., dudtdissi;
    if (p.disc_viscosity) {                                                   // force.F90:1555-1579
        const double hjv = 1. / hj1, csi = Di.w, csj = Dj.w;
        const double bi_ = (projv < 0.) ? (alphai * csi - beta * projv) : alphai * csi;
        const double bj_ = (projv < 0.) ? (alphaj * csj - beta * projv) : alphaj * csj;
        qrho2i = -0.5 * rho1i * bi_ * hi * rij1 * projv;
        qrho2j = usej ? -0.5 * rho1j * bj_ * hjv * rij1 * projv : 0.;
        dudtdissi = -0.5 * pmassj * rho1i * alphai * csi * hi * rij1 * (projv * projv) * grkerni;
    } else {
        const double ap = (projv < 0.) ? projv : 0.;                          // force.F90:1581-1592: approaching pairs only
        qrho2i = -0.5 * rho1i * vsigavi * ap;
        qrho2j = -0.5 * rho1j * vsigavj * ap;
        dudtdissi = pmassj * qrho2i * projv * grkerni;
    }
    const double gradp = pmassj * ((pro2i + qrho2i) * grkerni + (pro2j + qrho2j) * grkernj);
    double projsx = 0., projsy = 0., projsz = 0.;
    double dudtresist = 0.;
    if (dp.nvu >= 4) {                                                        // artificial conductivity, force.F90:1606-1624
        const double denij = vi.w - vj.w;
        double vsigu;
        if (p.gravity) vsigu = fabs(projv);
        else { const double xu = fabs(pri - prj) * (2. * rho1i * Dj.x) * rcp_pos(rho1i + Dj.x); vsigu = xu * rsqrt_pos(xu); }      // sqrt(|dP| 2/(rho_i + rho_j))
        const double auterm = 0.5 * pmassi * rho1i * p.alphau, autermj = 0.5 * pmassj * rho1j * p.alphau;
        f[A_DENDTDISS] += vsigu * denij * (auterm * grkerni + autermj * grkernj);
    }
    if (MHD) {                                                                // force.F90:1428-1444, :1626-1672, :2132-2150
        const double4 Ej = ldg256(a.recE + j);
        const double Bxi = Ei.x, Byi = Ei.y, Bzi = Ei.z, psii = Ei.w;
        const double Bxj = Ej.x, Byj = Ej.y, Bzj = Ej.z, psij = Ej.w;
        const double dBx = Bxi - Bxj, dBy = Byi - Byj, dBz = Bzi - Bzj;
        const double projBi = Bxi * runix + Byi * runiy + Bzi * runiz;
        const double projBj = Bxj * runix + Byj * runiy + Bzj * runiz;
        const double projdB = dBx * runix + dBy * runiy + dBz * runiz;
        const double dB2 = dBx * dBx + dBy * dBy + dBz * dBz;
        f[A_DIVBDIFF] += -pmassj * projdB * grkerni;
        const double rho21i = rho1i * rho1i, rho21j = rho1j * rho1j;
        const double avBterm = 0.5 * pmassi * rho1i * p.alphaB * rho1i, avBtermj = 0.5 * pmassj * rho1j * p.alphaB * rho1j;
        const double tx = dvx - projv * runix, ty = dvy - projv * runiy, tz = dvz - projv * runiz;
        const double t2 = tx * tx + ty * ty + tz * tz;
        const double vsigB = t2 * rsqrt_pos(t2);
        const double dBdissterm = (avBterm * grkerni + avBtermj * grkernj) * vsigB;
        if (p.iresistive_heating > 0) dudtresist = -0.5 * dB2 * dBdissterm;
        const double pmjrho21grkerni = pmassj * rho21i * grkerni, pmjrho21grkernj = pmassj * rho21j * grkernj;
        const double termi = pmjrho21grkerni * projBi;
        f[A_DIVBSYM] += termi + pmjrho21grkernj * projBj;
        const double dBrhoterm = -termi;
        const double dpsiterm = p.overcleanfac * (pmjrho21grkerni * psii * vwavei + pmjrho21grkernj * psij * vwavej);
        f[A_DBX] += dBrhoterm * dvx + dBdissterm * dBx - dpsiterm * runix;
        f[A_DBY] += dBrhoterm * dvy + dBdissterm * dBy - dpsiterm * runiy;
        f[A_DBZ] += dBrhoterm * dvz + dBdissterm * dBz - dpsiterm * runiz;
        // anisotropic Maxwell stress S_ab = -m B_a B_b / rho^2 projected on the pair direction (force.F90:1677-1684)
        const double si = -pmassi * rho21i * projBi * grkerni, sj = -pmassj * rho21j * projBj * grkernj;
        projsx = si * Bxi + sj * Bxj; projsy = si * Byi + sj * Byj; projsz = si * Bzi + sj * Bzj;
    }
    f[A_FX] += -runix * gradp - projsx;
    f[A_FY] += -runiy * gradp - projsy;
    f[A_FZ] += -runiz * gradp - projsz;
    f[A_DRHODT] += projv * grkerni;
    if (dp.nvu >= 4) f[A_DUDTDISS] += dudtdissi + dudtresist;
}


// ---- fast path: every particle is gas, no gravity / dust / individual timesteps / disc viscosity ---------------------------------
// Same sums as force_pair above, restructured for instruction count: one packed 96/128/160-byte record per neighbour instead of
// five gathers, the minimum-image wrap is skipped for target groups whose search region lies inside the box, and terms that vanish
// with grad W (q2 >= R^2) are not masked separately.
// staging block of the fast force kernel.  FORCE_STAGE = 1 keeps the candidates' {x, y} {z, 1/h} (first 32 bytes of the packed record) in
// shared memory; measured slower than reading the whole record with 256-bit loads at four resident CTAs (turb 128^3: 1.45 vs 1.38 ms),
// and much slower with MHD, where three resident CTAs of the staged block leave ~28 KB of L1 (mhdblast force 23.5 -> 29.1 ms).
#ifndef FORCE_STAGE
#define FORCE_STAGE 0
#endif
// candidates per round: 384 hold every group of a cubic lattice in one round; close-packed / glass-like sets stage 500-700 per group and
// take rounds of 768 (chosen per call from the candidate counts of the last density pass, like the density kernel)
#ifndef FORCE_ROUND
#define FORCE_ROUND ROUND_DEFAULT
#endif
#ifndef FORCE_ROUND_BIG
#define FORCE_ROUND_BIG 768
#endif
// Isothermal (ieos = 1) hydro without self-gravity: P = polyk rho and c_s = sqrt(polyk) (eos.f90, cons2prim.cu) make every entry of
// the third sector a function of h, so the neighbour record shrinks from three 32-byte sectors to two -- {x, y, z, 1/h} {v, (gradh,
// alpha) as the two real*4 the reference stores} -- and the L1 data pipe, which bounds this kernel, carries a third less.  Target
// and neighbour side go through the same function.  1/rho by a branch-free reciprocal (two Newton steps on the hardware seed).
struct Iso1 { double gradhfac, pro2, avw, rho1; };
__device__ __forceinline__ Iso1 iso1_derive(double h1, double packed, const ForceArgs &a)
{
    Iso1 r;
    const double gh = (double)__int_as_float(__double2loint(packed));
    const double al = a.iso1 == 2 ? a.iso_alpha : (double)__int_as_float(__double2hiint(packed));     // iso1 == 2: constant alpha
    const double h21 = h1 * h1, hf = a.iso_hfact * h1;
    r.gradhfac = h21 * h21 * a.iso_cnormk * gh;                  // force.F90:1327, as k_force_prep
    r.rho1 = rcp_pos(a.iso_pmass * (hf * hf * hf));              // rhoanddhdrho (part.F90:805-817)
    r.pro2 = a.iso_polyk * r.rho1;                               // P / rho^2
    r.avw = al * a.iso_cs;
    return r;
}

template <bool MHD, bool BIG> struct ForceFastSharedT { typedef WarpSharedT<BIG ? FORCE_ROUND_BIG : FORCE_ROUND, (FORCE_STAGE && !MHD && !BIG) ? 2 : 0, 0, !(FORCE_STAGE && !MHD && !BIG)> type; };

template <int K, bool PERIODIC, bool MHD, bool ADIA, bool GRAV, bool INDTS, bool BIG, bool ISO1>
#ifndef FORCE_MINB
#define FORCE_MINB 4
#endif
#ifndef FORCE_MHD_MINB
#define FORCE_MHD_MINB 3
#endif
__global__ void __launch_bounds__(128, MHD ? FORCE_MHD_MINB : (GRAV ? 3 : FORCE_MINB)) k_force_fast(const ForceArgs a, const __grid_constant__ DevParams dp)
{
    typedef SphKern<K> KF;
    typedef typename ForceFastSharedT<MHD, BIG>::type WS;
    extern __shared__ __align__(16) unsigned char force_smem[];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    WS &ws = reinterpret_cast<WS *>(force_smem)[wib];
    const unsigned ws_s = ws_shared_addr(ws);
    const unsigned hm_lane = ws_s + (unsigned)offsetof(WS, hm) + 4u * lane, sidx_s = ws_s + (unsigned)offsetof(WS, sidx);
    const unsigned rec2_s = ws_s + (unsigned)offsetof(WS, rec2);
    const int gwarp = blockIdx.x * 4 + wib;
    int *clist = a.stage_idx + (size_t)gwarp * a.scratch_per_warp;      // cell list of the current group (the only global scratch)
    const double Lx = dp.dxbound, Ly = dp.dybound, Lz = dp.dzbound;
    const float fLx = (float)Lx, fLy = (float)Ly, fLz = (float)Lz;
    const double halfLmin = 0.5 * fmin(Lx, fmin(Ly, Lz));
    const sphgpu_params &p = dp.p;
    unsigned long long st_pairs = 0, st_trial = 0;
    double st_dtc = 1.e29, st_dtf = 1.e29, st_dtmax = 0.;
    int st_nbinmax = 0, st_ncheckbin = 0;
    constexpr bool indts = INDTS;                                    // individual timestep bins (force.F90:1346-1358, :3272-3310)
    const float hmax_global = (a.ncells > 1) ? fmaxf(a.nodes[0].hmax[0], a.nodes[0].hmax[1]) : 0.f;
    const double pmass = p.massoftype[IGAS];
    const double beta = p.beta;
    constexpr bool USEJ = MHD || (ADIA && !GRAV);                    // force.F90:1343-1345 without dust
    static_assert(!ISO1 || (!MHD && !ADIA && !GRAV), "two-sector records: isothermal hydro only");
    constexpr int FSTRIDE = ISO1 ? 2 : (MHD ? 5 : ((ADIA || GRAV) ? 4 : 3));     // double4 per packed record

    while (true) {
        int cellid = 0;
        if (lane == 0) cellid = (int)atomicAdd(&a.cnt[CNT_WORK], 1ull);
        cellid = __shfl_sync(FULLMASK, cellid, 0);
        if (cellid >= a.ngroups) break;
        if (a.wl.order) cellid = a.wl.order[cellid];
        const Cell cell = a.groups[cellid];
        if (cell.active == 0) continue;                              // force.F90:509
        const double cx = 0.5 * (cell.lo[0] + cell.hi[0]), cy = 0.5 * (cell.lo[1] + cell.hi[1]), cz = 0.5 * (cell.lo[2] + cell.hi[2]);
        const double halfext = 0.5 * fmax(cell.hi[0] - cell.lo[0], fmax(cell.hi[1] - cell.lo[1], cell.hi[2] - cell.lo[2]));
        float tlo[3], thi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { tlo[k] = __double2float_rd(cell.lo[k]); thi[k] = __double2float_ru(cell.hi[k]); }
        const double rreach = KF::radkern * fmax(cell.hmax, (double)hmax_global) * a.hscale * 1.0001;
        const double rcut = KF::radkern * cell.hmax * a.hscale;
        float reach = 0.f;
        const int *cl = clist;                                       // list in use: the prepared one, or this warp's slice after a walk in here
        int ncl = a.wl.ncl[cellid];
        if (ncl >= 0) { cl = a.wl.list + (size_t)cellid * a.wl.cap; reach = a.wl.reach[cellid]; }
        else ncl = warp_walk<true, PERIODIC>(a.nodes, a.cells, a.ncells, tlo, thi, __double2float_ru(rcut), __double2float_ru(KF::radkern * a.hscale),
                                             fLx, fLy, fLz, ws.walk_stack(), clist, a.scratch_per_warp, reach);
        if (ncl < 0) { if (lane == 0) atomicMax(&a.cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); break; }
        // Largest kernel radius of any target or candidate of THIS group: the walk's reach (search radius + extent over the cells it
        // hit) bounds it locally, radkern x the global hmax bounds it anyway.  The FP16 filter, which works on nearest images relative
        // to the group centre, is valid while halfext + that radius < L/2; a thin periodic box (the shock tube's y, z) whose large-h
        // side would fail this must not switch the filter off for the groups of the small-h side.
        const double rloc = fmin(rreach, (double)reach * 1.0001);
        const bool wide = PERIODIC && (halfext + rloc >= 0.999 * halfLmin);
        const bool wrapf = wide && rloc < 0.9 * halfLmin;           // kernel radii well below half the box: the filter wraps per pair (walk.cuh)
        // no pair of this group can straddle the periodic boundary: |xi - xj| <= L/2 for every candidate that survives the prefilter
        const bool interior = !PERIODIC || (cell.lo[0] - rloc > p.xmin && cell.hi[0] + rloc < p.xmax && cell.lo[1] - rloc > p.ymin &&
                                            cell.hi[1] + rloc < p.ymax && cell.lo[2] - rloc > p.zmin && cell.hi[2] + rloc < p.zmax);
        int nlist = 0;
        FilterScale fs = filter_scale((float)halfext, reach);
        if (wrapf) filter_scale_wrap(fs);
        // ---- lane = target (start_cell, force.F90:2172-2514; per-particle part done by k_force_prep)
        const int s = cell.start + min(lane, cell.count - 1);
        bool act = false;
        { bool g_, d_; int t_; if (lane < cell.count) get_partinfo_d(a.stype[s], p.set_boundaries_to_active, 0, act, g_, d_, t_); }
        const double4 *ri = a.frec + FSTRIDE * (size_t)s;
        const double4 T0 = ri[0];
        double4 T1 = ri[1], T2;
        if (ISO1) { const Iso1 di = iso1_derive(T0.w, T1.w, a); T1.w = di.gradhfac; T2 = make_double4(di.pro2, a.iso_cs, di.avw, di.rho1); }
        else T2 = ri[2];
        double4 T3 = make_double4(0., 0., 0., 0.), T4 = T3;
        if (ADIA || MHD || GRAV) T3 = ri[3];
        if (MHD) T4 = ri[4];
        const double xi = T0.x, yi = T0.y, zi = T0.z, hi1 = T0.w, hi21 = hi1 * hi1;
        const double h = a.pos4[s].w;
        const double gi = T1.w, pro2i = T2.x, vwavei = T2.y, avwi = T2.z, rho1i = T2.w, pri = T3.x, alphai = T3.w;
        const double hrho1i = -0.5 * rho1i;
        // force.F90:2255: inactive targets skipped (empty masks); a wide periodic search switches the filter off
        const FilterTarget ft = filter_target(fs, (float)(xi - cx), (float)(yi - cy), (float)(zi - cz),
                                              act ? ((wide && !wrapf) ? -1.f : __double2float_ru(KF::radkern * h)) : 0.f);
        double fpot = 0.;
        double fx = 0., fy = 0., fz = 0., drhodt = 0., dudtdiss = 0., dendtdiss = 0., divBsym = 0., dBx = 0., dBy = 0., dBz = 0., divBdiff = 0.;
        double vsigmax = 0.;
        int npair = 0, ibin_neigh = 0;
        for (int cellpos = 0; cellpos < ncl;) {                     // rounds of <= ROUND candidates staged in shared memory
            auto stage_rec = [&](int slot, int, const double2 &xy, const double2 &zw) { if (WS::P2 > 0) { ws.rec2[0][slot] = xy; ws.rec2[WS::P2 - 1][slot] = zw; } };
            const int nr = stage_round<PERIODIC, true>(ws, cl, ncl, cellpos, reinterpret_cast<const double2 *>(a.frec), 2 * FSTRIDE, cx, cy, cz, Lx, Ly, Lz,
                                                       (float)KF::radkern, a.max_leaf, fs, PERIODIC && interior, cell.start, stage_rec);
            const int myslot = ws.selfslot[lane];
            nlist += nr;
            const int nchunk = (nr + 31) >> 5;
            unsigned nz;
            if (wrapf) { const FilterWrap fw = filter_wrap(fs, (float)dp.dxbound, (float)dp.dybound, (float)dp.dzbound); nz = build_masks<true, true>(ws, nr, ft, &fw); }
            else nz = wide ? build_masks<false>(ws, nr, ft) : build_masks<true>(ws, nr, ft);
            int c = -1; unsigned m = 0u;
            // Two neighbours per trip, written phase by phase over both so that two independent FP64 dependency chains and both
            // records are in flight.  No branches: grad W is evaluated as truncated powers (zero beyond the support), r = 0 gives
            // 1/r := 0, so a non-member (exact test fails, j == s, or the self padding of an odd hit count) adds exact zeros;
            // only the pair count, the signal-speed maximum and the softened gravity need the membership flags.
            auto pair2 = [&](int slot0, int slot1) {
                // a lane with an odd number of hits evaluates its first neighbour twice, the second time as a non-member
                const int sl[2] = {slot0, slot1 >= 0 ? slot1 : slot0};
                const bool live[2] = {slot0 != myslot, slot1 >= 0 && slot1 != myslot};
                int jj[2];
                double4 R0[2], R1[2], R2[2], R3[2], E[2];
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    // {x,y,z,1/h} (head of the dependency chain) from the staging block, the rest of the packed record from global memory
                    jj[k] = (int)lds_u32(sidx_s + 4u * (unsigned)sl[k]);
                    const double4 *rj = a.frec + FSTRIDE * (size_t)jj[k];
                    if (WS::P2 > 0) {
                        const double2 XY = lds_d2(rec2_s + 16u * (unsigned)sl[k]), ZW = lds_d2(rec2_s + 16u * (unsigned)(WS::ROUND + sl[k]));
                        R0[k] = make_double4(XY.x, XY.y, ZW.x, ZW.y);
                    } else R0[k] = ldg256(rj);
                    R1[k] = ldg256(rj + 1);
                    if (!ISO1) R2[k] = ldg256(rj + 2);
                    if (ADIA || GRAV) R3[k] = ldg256(rj + 3);
                    if (MHD) E[k] = ldg256(rj + 4);
                }
                if (ISO1) {
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const Iso1 dj = iso1_derive(R0[k].w, R1[k].w, a);
                        R1[k].w = dj.gradhfac; R2[k] = make_double4(dj.pro2, a.iso_cs, dj.avw, dj.rho1);
                    }
                }
                double dx[2], dy[2], dz[2];
#pragma unroll
                for (int k = 0; k < 2; k++) { dx[k] = xi - R0[k].x; dy[k] = yi - R0[k].y; dz[k] = zi - R0[k].z; }
                if (PERIODIC && !interior) {                            // force.F90:1266-1270
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        if (fabs(dx[k]) > 0.5 * Lx) dx[k] = dx[k] - copysign(Lx, dx[k]);
                        if (fabs(dy[k]) > 0.5 * Ly) dy[k] = dy[k] - copysign(Ly, dy[k]);
                        if (fabs(dz[k]) > 0.5 * Lz) dz[k] = dz[k] - copysign(Lz, dz[k]);
                    }
                }
                double r2[2], rij1[2], grkerni[2], grkernj[2];
                bool ini[2], inj[2], isn[2], dropj[2] = {false, false};
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    r2[k] = __dadd_rn(__dadd_rn(__dmul_rn(dx[k], dx[k]), __dmul_rn(dy[k], dy[k])), __dmul_rn(dz[k], dz[k]));
                    const double hj1 = R0[k].w;
                    const double q2i = __dmul_rn(r2[k], hi21), q2j = __dmul_rn(r2[k], __dmul_rn(hj1, hj1));       // force.F90:1272, :1285
                    ini[k] = (q2i < KF::radkern2) && live[k]; inj[k] = (q2j < KF::radkern2) && live[k];          // :1287, :1230 (exact membership)
                    if (indts && a.refnodes && inj[k] && !ini[k]) {       // only j's kernel reaches: did the reference's walk find j's leaf?
                        const int jj0 = (int)lds_u32(sidx_s + 4u * (unsigned)sl[k]);
                        if (!ref_walk_reaches<PERIODIC>(a, s, jj0, Lx, Ly, Lz)) { inj[k] = false; dropj[k] = true; }
                    }
                    isn[k] = ini[k] || inj[k];
                    npair += isn[k] ? 1 : 0;
                    if (indts && isn[k] && abs((int)a.stype[jj[k]]) != IBOUNDARY) {   // j neighbours an active particle: wake flag, Saitoh-Makino input
                        if (a.s_wake[jj[k]] < a.ibinnow_m1) atomicMax(&a.s_wake[jj[k]], a.ibinnow_m1);
                        ibin_neigh = max(ibin_neigh, (int)a.s_ibinold[jj[k]]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 2; k++) rij1[k] = rsqrt_pos(r2[k]);                 // force.F90:1293-1299
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const double rij = r2[k] * rij1[k];
                    // the padding repeats a real neighbour: it (and the self pair) enter with weight 0 on both gradients
                    grkerni[k] = KF::grkern_bf(rij * hi1) * (live[k] ? gi : 0.);         // :1301-1302
                    grkernj[k] = KF::grkern_bf(rij * R0[k].w) * ((live[k] && !(indts && dropj[k])) ? R1[k].w : 0.);   // :1325-1327
                }
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const double runix = dx[k] * rij1[k], runiy = dy[k] * rij1[k], runiz = dz[k] * rij1[k];
                    double fgrav = 0.;
                    if (GRAV) {                                            // softened gravity of SPH-neighbour pairs, force.F90:1303-1339, :1522
                        const double hj1 = R0[k].w, rij = r2[k] * rij1[k];
                        const double q2i = __dmul_rn(r2[k], hi21), q2j = __dmul_rn(r2[k], __dmul_rn(hj1, hj1));
                        // both sides evaluated and selected (no branch between lanes that are inside one kernel, the other, both or none)
                        double phis, fmi, phij, fmj;
                        KF::softening_in(fmin(q2i, KF::radkern2), fmin(rij * hi1, KF::radkern), rij1[k] * h, phis, fmi);
                        KF::softening_in(fmin(q2j, KF::radkern2), fmin(rij * hj1, KF::radkern), rij1[k] * rcp_pos(hj1), phij, fmj);
                        const double newt = rij1[k] * rij1[k];
                        const double phii = ini[k] ? phis * hi1 : -rij1[k];
                        const double fgravi = ini[k] ? fmi * hi21 + T3.z * grkerni[k] : newt;
                        const double fgravj = inj[k] ? fmj * (hj1 * hj1) + R3[k].z * grkernj[k] : newt;
                        fgrav = isn[k] ? 0.5 * pmass * (fgravi + fgravj) : 0.;
                        fpot += isn[k] ? pmass * phii : 0.;
                    }
                    const double dvx = T1.x - R1[k].x, dvy = T1.y - R1[k].y, dvz = T1.z - R1[k].z;
                    const double projv = dvx * runix + dvy * runiy + dvz * runiz;
                    const double bp = beta * projv;
                    const double vwavej = R2[k].y;
                    // :1423-1426, :1501-1504: vsigmax >= 0 makes the clip at 0 implicit; j-side only when its terms are used (:1343-1345)
                    double vs = vwavei - bp;
                    if (USEJ || inj[k]) vs = dmax(vs, vwavej - bp);
                    vsigmax = dmax(vsigmax, isn[k] ? vs : 0.);
                    const double ap = (projv < 0.) ? projv : 0.;          // force.F90:1581-1592: approaching pairs only
                    const double qrho2i = hrho1i * dmax(avwi - bp, 0.) * ap;
                    const double rho1j = R2[k].w;
                    const double qrho2j = (-0.5 * rho1j) * dmax(R2[k].z - bp, 0.) * ap;
                    const double gradp = pmass * ((pro2i + qrho2i) * grkerni[k] + (R2[k].x + qrho2j) * grkernj[k]);
                    double projsx = 0., projsy = 0., projsz = 0.;
                    if (ADIA) {                                            // artificial conductivity, force.F90:1606-1624
                        const double denij = T3.y - R3[k].y;
                        // sqrt(|dP| 2/(rho_i + rho_j)) without the library's division and square root (branch-free reciprocal and rsqrt)
                        const double xu = fabs(pri - R3[k].x) * (2. * rho1i * rho1j) * rcp_pos(rho1i + rho1j);
                        const double vsigu = GRAV ? fabs(projv) : xu * rsqrt_pos(xu);
                        const double auterm = 0.5 * pmass * rho1i * p.alphau, autermj = 0.5 * pmass * rho1j * p.alphau;
                        dendtdiss += vsigu * denij * (auterm * grkerni[k] + autermj * grkernj[k]);
                        dudtdiss += pmass * qrho2i * projv * grkerni[k];
                    }
                    if (MHD) {                                             // force.F90:1428-1444, :1626-1684, :2132-2150
                        const double Bxi = T4.x, Byi = T4.y, Bzi = T4.z, psii = T4.w;
                        const double dBxx = Bxi - E[k].x, dByy = Byi - E[k].y, dBzz = Bzi - E[k].z;
                        const double projBi = Bxi * runix + Byi * runiy + Bzi * runiz;
                        const double projBj = E[k].x * runix + E[k].y * runiy + E[k].z * runiz;
                        const double projdB = dBxx * runix + dByy * runiy + dBzz * runiz;
                        divBdiff += -pmass * projdB * grkerni[k];
                        const double rho21i = rho1i * rho1i, rho21j = rho1j * rho1j;
                        const double avBterm = 0.5 * pmass * rho1i * p.alphaB * rho1i, avBtermj = 0.5 * pmass * rho1j * p.alphaB * rho1j;
                        const double tx = dvx - projv * runix, ty = dvy - projv * runiy, tz = dvz - projv * runiz;
                        const double t2 = tx * tx + ty * ty + tz * tz;
                        const double vsigB = t2 * rsqrt_pos(t2);                 // |v_ij x r|, branch-free
                        const double dBdissterm = (avBterm * grkerni[k] + avBtermj * grkernj[k]) * vsigB;
                        if (ADIA && p.iresistive_heating > 0) dudtdiss += -0.5 * (dBxx * dBxx + dByy * dByy + dBzz * dBzz) * dBdissterm;
                        const double pmjrho21grkerni = pmass * rho21i * grkerni[k], pmjrho21grkernj = pmass * rho21j * grkernj[k];
                        const double termi = pmjrho21grkerni * projBi;
                        divBsym += termi + pmjrho21grkernj * projBj;
                        const double dpsiterm = p.overcleanfac * (pmjrho21grkerni * psii * vwavei + pmjrho21grkernj * E[k].w * vwavej);
                        dBx += -termi * dvx + dBdissterm * dBxx - dpsiterm * runix;
                        dBy += -termi * dvy + dBdissterm * dByy - dpsiterm * runiy;
                        dBz += -termi * dvz + dBdissterm * dBzz - dpsiterm * runiz;
                        const double si = -pmass * rho21i * projBi * grkerni[k], sj = -pmass * rho21j * projBj * grkernj[k];   // Maxwell stress, :1677-1684
                        projsx = si * Bxi + sj * E[k].x; projsy = si * Byi + sj * E[k].y; projsz = si * Bzi + sj * E[k].z;
                    }
                    fx += -runix * (gradp + fgrav) - projsx;
                    fy += -runiy * (gradp + fgrav) - projsy;
                    fz += -runiz * (gradp + fgrav) - projsz;
                    drhodt += projv * grkerni[k];
                }
            };
            while (true) {
                int slot0, slot1;
                next_hits2(hm_lane, nz, c, m, slot0, slot1);
                if (slot0 < 0) break;
                pair2(slot0, slot1);
            }
            __syncwarp();
        }
        // ---- finish_cell_and_store_results (force.F90:2649-3330), lane = target, gas only ----
        if (act) {
            st_pairs += npair; st_trial += nlist;
            double dtc = p.dtmax, dtf = 1.e29, dtclean = 1.e29;
            double fxyz4 = 0.;
            double4 dB = make_double4(0., 0., 0., 0.);
            float divBsymm4 = 0.f;
            const double rhoi = 1. / rho1i;
            if (GRAV) {                                              // force.F90:2909-2927: far field (L2P + distant P2P) from gravity.cu
                const double4 g = a.gacc[a.perm[s]];
                double potensoft0, dum;
                KF::softening(0., 0., potensoft0, dum);
                fx += g.x; fy += g.y; fz += g.z;
                a.s_poten[s] = (float)(0.5 * pmass * (fpot + pmass * potensoft0 * hi1) + 0.5 * pmass * g.w);
            }
            if (MHD) {                                               // force.F90:2939-2965
                const double B2i = T4.x * T4.x + T4.y * T4.y + T4.z * T4.z;
                double frac_divB = 0.;
                if (B2i > 0.0) {
                    const double betai = 2.0 * pri / B2i;
                    if (betai < 2.0) frac_divB = 1.0;
                    else if (betai < 10.0) frac_divB = (10.0 - betai) * 0.125;
                }
                fx -= T4.x * divBsym * frac_divB; fy -= T4.y * divBsym * frac_divB; fz -= T4.z * divBsym * frac_divB;
                divBsymm4 = (float)(rhoi * divBsym);
            }
            const double drhodti = pmass * drhodt;
            const double divvi = -drhodti * rho1i;
            if (ADIA) {                                              // force.F90:3024-3095 (ien_type = energy, fac = 1)
                const double pdv_work = pri * rho1i * rho1i * drhodti;
                if (p.ipdv_heating > 0) fxyz4 += pdv_work;
                if (p.ishock_heating > 0) fxyz4 += dudtdiss;
                fxyz4 += dendtdiss;
            }
            if (MHD) {                                               // force.F90:3103-3125
                dB.x = dBx; dB.y = dBy; dB.z = dBz;
                if (p.psidecayfac > 0.) {
                    const double vcleani = p.overcleanfac * vwavei;
                    const double dtau = p.psidecayfac * vcleani * hi1;
                    dB.w = -vcleani * divBdiff * rho1i - T4.w * dtau - 0.5 * T4.w * divvi;
                    dtclean = p.C_cour * h / (vcleani + DBL_MIN);
                }
            }
            const double vsigdtc = fmax(vsigmax, vwavei);
            if (vsigdtc > DBL_MIN) dtc = p.C_cour * h / (vsigdtc * fmax(p.alpha, 1.0));          // force.F90:3138-3141
            if (ADIA) {
                const double eni = T3.y;
                if (eni + dtc * fxyz4 < DBL_EPSILON && eni > DBL_EPSILON) fxyz4 = fxyz4 / (1. - dtc * fxyz4 / eni);       // :3144-3148
            }
            const double f2i = fx * fx + fy * fy + fz * fz;
            if (fabs(f2i) > DBL_EPSILON) dtf = p.C_force * sqrt(h / sqrt(f2i));                  // force.F90:3217-3219
            a.s_fxyzu[s] = make_double4(fx, fy, fz, fxyz4);
            a.s_divvf[s] = (float)divvi;
            if (MHD) { a.s_dB[s] = dB; a.s_divBsymm[s] = divBsymm4; }
            a.s_done[s] = 2;
            if (indts) {                                             // force.F90:3272-3310 + get_newbin (utils_indtimesteps.f90:230-287)
                double dti = dtc;
                const double dtitmp = fmin(dtf, dtclean);
                if (dtitmp < dti + DBL_MIN && dtitmp < p.dtmax) dti = dtitmp;
                const int ibin_oldi = (int)a.s_ibin[s];
                int ibin_newi;
                if (dti > p.dtmax) ibin_newi = 0;
                else if (dti < DBL_MIN) ibin_newi = 30;
                else ibin_newi = max((int)(log(2. * p.dtmax / dti) * 1.4426950408889634 - DBL_EPSILON), 0);
                int ibini = ibin_oldi;
                if (ibin_newi > ibin_oldi) ibini = ibin_newi;
                else if (ibin_newi < ibin_oldi && ibin_oldi <= a.nbinmax && a.icall < 2) {
                    if (a.istepfrac % (1 << (a.nbinmax - (ibin_oldi - 1))) == 0) ibini = ibini - 1;
                }
                ibini = max(ibini, ibin_neigh - 1);                  // Saitoh-Makino limiter
                a.s_ibinnew[s] = (int8_t)ibini;
                st_nbinmax = max(st_nbinmax, ibini); st_ncheckbin += 1;
            } else {
                st_dtc = fmin(st_dtc, dtc);
                st_dtf = fmin(st_dtf, fmin(dtf, dtclean));
                st_dtmax = fmax(st_dtmax, dtc);
            }
            (void)alphai;
        }
        __syncwarp();
    }
    st_dtc = warp_min(st_dtc); st_dtf = warp_min(st_dtf); st_dtmax = warp_max(st_dtmax);
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) { st_pairs += __shfl_xor_sync(FULLMASK, st_pairs, sft); st_trial += __shfl_xor_sync(FULLMASK, st_trial, sft); }
    if (indts) {
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) { st_nbinmax = max(st_nbinmax, __shfl_xor_sync(FULLMASK, st_nbinmax, sft)); st_ncheckbin += __shfl_xor_sync(FULLMASK, st_ncheckbin, sft); }
        if (lane == 0 && st_ncheckbin) { atomicMax(&a.cnt[CNT_NBINMAX], (unsigned long long)st_nbinmax); atomicAdd(&a.cnt[CNT_NCHECKBIN], (unsigned long long)st_ncheckbin); }
    }
    if (lane == 0) {
        atomicAdd(&a.cnt[CNT_NPAIRS], st_pairs); atomicAdd(&a.cnt[CNT_NTRIAL], st_trial);
        atomic_min_pos(&a.dscal[DS_DTCOURANT], st_dtc); atomic_min_pos(&a.dscal[DS_DTFORCE], st_dtf);
        atomic_min_pos(&a.dscal[DS_DTMINI], st_dtc); atomic_max_pos(&a.dscal[DS_DTMAXI], st_dtmax);
    }
}

#ifndef XTRA_MINB
#define XTRA_MINB 3
#endif
template <int K, bool PERIODIC, bool MHD, bool XTRA>
__global__ void __launch_bounds__(128, XTRA ? XTRA_MINB : 4) k_force(const ForceArgs a, const __grid_constant__ DevParams dp)
{
    typedef SphKern<K> KF;
    typedef WarpSharedGeneral WS;
    extern __shared__ __align__(16) unsigned char forceg_smem[];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    WS &ws = reinterpret_cast<WS *>(forceg_smem)[wib];
    const unsigned hm_lane = ws_shared_addr(ws) + (unsigned)offsetof(WS, hm) + 4u * lane;
    const int gwarp = blockIdx.x * 4 + wib;
    int *clist = a.stage_idx + (size_t)gwarp * a.scratch_per_warp;      // cell list of the current group (the only global scratch)
    const double Lx = dp.dxbound, Ly = dp.dybound, Lz = dp.dzbound;
    const float fLx = (float)Lx, fLy = (float)Ly, fLz = (float)Lz;
    const double halfLmin = 0.5 * fmin(Lx, fmin(Ly, Lz));
    const sphgpu_params &p = dp.p;
    unsigned long long st_pairs = 0, st_trial = 0;
    double st_dtc = 1.e29, st_dtf = 1.e29, st_dtmin = 1.e29, st_dtmax = 0.;
    int st_nbinmax = 0, st_ncheckbin = 0;
    const float hmax_global = (a.ncells > 1) ? fmaxf(a.nodes[0].hmax[0], a.nodes[0].hmax[1]) : 0.f;

    while (true) {
        int cellid = 0;
        if (lane == 0) cellid = (int)atomicAdd(&a.cnt[CNT_WORK], 1ull);
        cellid = __shfl_sync(FULLMASK, cellid, 0);
        if (cellid >= a.ngroups) break;
        if (a.wl.order) cellid = a.wl.order[cellid];
        const Cell cell = a.groups[cellid];
        if (cell.active == 0) continue;                              // force.F90:509
        const double cx = 0.5 * (cell.lo[0] + cell.hi[0]), cy = 0.5 * (cell.lo[1] + cell.hi[1]), cz = 0.5 * (cell.lo[2] + cell.hi[2]);
        const double halfext = 0.5 * fmax(cell.hi[0] - cell.lo[0], fmax(cell.hi[1] - cell.lo[1], cell.hi[2] - cell.lo[2]));
        float tlo[3], thi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { tlo[k] = __double2float_rd(cell.lo[k]); thi[k] = __double2float_ru(cell.hi[k]); }
        const double rcut = KF::radkern * cell.hmax * a.hscale;
        float reach = 0.f;
        const int *cl = clist;                                       // list in use: the prepared one, or this warp's slice after a walk in here
        int ncl = a.wl.ncl[cellid];
        if (ncl >= 0) { cl = a.wl.list + (size_t)cellid * a.wl.cap; reach = a.wl.reach[cellid]; }
        else ncl = warp_walk<true, PERIODIC>(a.nodes, a.cells, a.ncells, tlo, thi, __double2float_ru(rcut), __double2float_ru(KF::radkern * a.hscale),
                                             fLx, fLy, fLz, ws.stack, clist, a.scratch_per_warp, reach);
        if (ncl < 0) { if (lane == 0) atomicMax(&a.cnt[CNT_ERR], (unsigned long long)SPHGPU_ERR_OVERFLOW); break; }
        // (largest kernel radius of this group's targets and candidates: bounded locally by the walk's reach, see k_force_fast)
        const double rloc = fmin(KF::radkern * fmax(cell.hmax, (double)hmax_global) * a.hscale, (double)reach) * 1.0001;
        const bool wide = PERIODIC && (halfext + rloc >= 0.999 * halfLmin);
        int nlist = 0;
        const FilterScale fs = filter_scale((float)halfext, reach);
        // ---- lane = target: start_cell (force.F90:2172-2514); the per-particle part was done by k_force_prep
        const int s = cell.start + min(lane, cell.count - 1);
        bool act = false, gasi = true, dusti = false; int itypei = IGAS;
        if (lane < cell.count) get_partinfo_d(a.stype[s], p.set_boundaries_to_active, p.dust, act, gasi, dusti, itypei);
        const double4 pi = a.pos4[s], vi = a.vel4[s], Ci = a.recC[s], Di = a.recD[s];
        double4 Ei = make_double4(0., 0., 0., 0.);
        if (MHD) Ei = a.recE[s];
        const double h = pi.w;
        const double2 hv = a.hinv[s];
        const double hi1 = hv.x, hi21 = hv.y;
        // force.F90:2255: inactive targets skipped (empty masks); a wide periodic search switches the filter off
        const FilterTarget ft = filter_target(fs, (float)(pi.x - cx), (float)(pi.y - cy), (float)(pi.z - cz),
                                              act ? (wide ? -1.f : __double2float_ru(KF::radkern * h)) : 0.f);
        double f[12];
#pragma unroll
        for (int k = 0; k < 12; k++) f[k] = 0.;
        double vsigmax = 0.;
        int npair = 0;
        XtraSums xs; xs.fdx = xs.fdy = xs.fdz = 0.; xs.tsmin = 1.e29; xs.ibin_neigh = 0;
        // The candidates are taken one SORT CLASS at a time (cells, target groups and therefore rounds hold one class each, tree.cu), so
        // the kind of pair -- gas-gas, gas-dust, other -- is the same for the whole warp and each kind has its own branch-free body.
        const int ci = sort_class(a.stype[cell.start]);
        for (int cj = 0; cj < 3; cj++) {
            if (!((a.class_mask >> cj) & 1)) continue;
            const int cp = (ci == 0 && cj == 0) ? 0 : ((ci + cj == 1) ? 1 : 2);
            for (int cellpos = 0; cellpos < ncl;) {                 // rounds of <= ROUND candidates staged in shared memory
                const int nr = stage_round<PERIODIC, false>(ws, cl, ncl, cellpos, reinterpret_cast<const double2 *>(a.pos4), 2, cx, cy, cz, Lx, Ly, Lz, (float)KF::radkern,
                                                            a.max_leaf, fs, false, 0, NoRecord(), cj, a.stype);
                if (nr == 0) continue;
                nlist += nr;
                unsigned nz = wide ? build_masks<false>(ws, nr, ft) : build_masks<true>(ws, nr, ft);     // (no per-pair wrap here: this kernel is instruction-fetch bound)
                int c = -1; unsigned m = 0u;
#define FORCE_PAIR_LOOP(CP)                                                                                                                                          \
                while (true) {      /* two neighbours per trip, every lane on the same path */                                                                      \
                    int sl[2];                                                                                                                                       \
                    next_hits2(hm_lane, nz, c, m, sl[0], sl[1]);                                                                                                     \
                    if (sl[0] < 0) break;                                                                                                                            \
                    if (CP == 0) {  /* gas-gas: both bodies inlined side by side, two independent FP64 chains */                                                    \
                        force_pair<K, PERIODIC, MHD, XTRA, CP>(f, vsigmax, npair, sl[0], ws.sidx, s, pi, h, hi1, hi21, gasi, vi, Ci, Di, Ei, a, dp, Lx, Ly, Lz, xs, itypei); \
                        force_pair<K, PERIODIC, MHD, XTRA, CP>(f, vsigmax, npair, sl[1], ws.sidx, s, pi, h, hi1, hi21, gasi, vi, Ci, Di, Ei, a, dp, Lx, Ly, Lz, xs, itypei); \
                    } else {        /* one copy of the body: the general kernel is bound by instruction fetch, not by the FP64 pipe */                              \
                        _Pragma("unroll 1")                                                                                                                          \
                        for (int k = 0; k < 2; k++)                                                                                                                  \
                            force_pair<K, PERIODIC, MHD, XTRA, CP>(f, vsigmax, npair, sl[k], ws.sidx, s, pi, h, hi1, hi21, gasi, vi, Ci, Di, Ei, a, dp, Lx, Ly, Lz, xs, itypei); \
                    }                                                                                                                                                \
                }
                if (cp == 0) { FORCE_PAIR_LOOP(0) } else if (cp == 1) { FORCE_PAIR_LOOP(1) } else { FORCE_PAIR_LOOP(2) }
#undef FORCE_PAIR_LOOP
                __syncwarp();
            }
        }
        // ---- finish_cell_and_store_results (force.F90:2649-3330), lane = target ----
        if (act) {
            st_pairs += npair; st_trial += nlist;
            const double pmassi = Di.z;
            double fx = f[A_FX], fy = f[A_FY], fz = f[A_FZ];
            double dtc = p.dtmax, dtf = 1.e29, dtclean = 1.e29, dtdrag = 1.e29;
            double fxyz4 = 0., divvi = 0.;
            if (XTRA && p.gravity) {                                 // force.F90:2909-2927: far field (L2P + distant P2P) from gravity.cu
                const double4 g = a.gacc[a.perm[s]];
                double potensoft0, dum;
                KF::softening(0., 0., potensoft0, dum);
                fx += g.x; fy += g.y; fz += g.z;
                a.s_poten[s] = (float)(0.5 * pmassi * (f[A_POT] + pmassi * potensoft0 * hi1) + 0.5 * pmassi * g.w);
            }
            double4 dB = make_double4(0., 0., 0., 0.);
            float divBsymm4 = 0.f;
            if (gasi) {
                const double rho1i = Di.x, rhoi = 1. / rho1i, pri = Ci.w, vwavei = Ci.y;
                if (MHD) {                                           // force.F90:2939-2965
                    const double B2i = Ei.x * Ei.x + Ei.y * Ei.y + Ei.z * Ei.z;
                    const double divBsymmi = f[A_DIVBSYM];
                    double frac_divB = 0.;
                    if (B2i > 0.0) {
                        const double betai = 2.0 * pri / B2i;
                        if (betai < 2.0) frac_divB = 1.0;
                        else if (betai < 10.0) frac_divB = (10.0 - betai) * 0.125;
                    }
                    fx -= Ei.x * divBsymmi * frac_divB; fy -= Ei.y * divBsymmi * frac_divB; fz -= Ei.z * divBsymmi * frac_divB;
                    divBsymm4 = (float)(rhoi * divBsymmi);
                }
                const double drhodti = pmassi * f[A_DRHODT];
                divvi = -drhodti * rho1i;
                if (dp.nvu >= 4) {                                   // force.F90:3024-3095 (ien_type = energy, fac = rho/rhogas = 1)
                    const double pdv_work = pri * rho1i * rho1i * drhodti;
                    if (p.ipdv_heating > 0) fxyz4 += pdv_work;
                    if (p.ishock_heating > 0) fxyz4 += f[A_DUDTDISS];
                    fxyz4 += f[A_DENDTDISS];
                }
                if (MHD) {                                           // force.F90:3103-3125
                    dB.x = f[A_DBX]; dB.y = f[A_DBY]; dB.z = f[A_DBZ];
                    if (p.psidecayfac > 0.) {
                        const double vcleani = p.overcleanfac * vwavei;
                        const double dtau = p.psidecayfac * vcleani * hi1;
                        const double psii = Ei.w;
                        dB.w = -vcleani * f[A_DIVBDIFF] * rho1i - psii * dtau - 0.5 * psii * divvi;
                        dtclean = p.C_cour * h / (vcleani + DBL_MIN);
                    }
                }
                const double vsigdtc = fmax(vsigmax, vwavei);
                if (vsigdtc > DBL_MIN) dtc = p.C_cour * h / (vsigdtc * fmax(p.alpha, 1.0));      // force.F90:3138-3141
                if (dp.nvu >= 4) {
                    const double eni = vi.w;
                    if (eni + dtc * fxyz4 < DBL_EPSILON && eni > DBL_EPSILON) fxyz4 = fxyz4 / (1. - dtc * fxyz4 / eni);   // :3144-3148
                }
            } else {
                if (vsigmax > DBL_MIN) dtc = p.C_cour * h / vsigmax;
            }
            const double f2i = fx * fx + fy * fy + fz * fz;
            if (fabs(f2i) > DBL_EPSILON) dtf = p.C_force * sqrt(h / sqrt(f2i));                  // force.F90:3217-3219
            if (XTRA && p.dust) {                                    // force.F90:2978-2988, :3239-3246
                fx += xs.fdx; fy += xs.fdy; fz += xs.fdz;
                a.s_tstop[s] = xs.tsmin;
                dtdrag = 0.9 * xs.tsmin;
            }
            a.s_fxyzu[s] = make_double4(fx, fy, fz, fxyz4);
            a.s_divvf[s] = (float)divvi;
            if (MHD) { a.s_dB[s] = dB; a.s_divBsymm[s] = divBsymm4; }
            a.s_done[s] = gasi ? 2 : 1;
            if (XTRA && p.ind_timesteps) {                           // force.F90:3272-3310 + get_newbin (utils_indtimesteps.f90:230-287)
                double dti = dtc;
                const double dtitmp = fmin(fmin(dtf, dtclean), dtdrag);
                if (dtitmp < dti + DBL_MIN && dtitmp < p.dtmax) dti = dtitmp;
                const int ibin_oldi = (int)a.s_ibin[s];
                int ibin_newi;
                if (dti > p.dtmax) ibin_newi = 0;
                else if (dti < DBL_MIN) ibin_newi = 30;
                else ibin_newi = max((int)(log(2. * p.dtmax / dti) * 1.4426950408889634 - DBL_EPSILON), 0);
                int ibini = ibin_oldi;
                if (ibin_newi > ibin_oldi) ibini = ibin_newi;
                else if (ibin_newi < ibin_oldi && ibin_oldi <= a.nbinmax && a.icall < 2) {
                    if (a.istepfrac % (1 << (a.nbinmax - (ibin_oldi - 1))) == 0) ibini = ibini - 1;
                }
                ibini = max(ibini, xs.ibin_neigh - 1);               // Saitoh-Makino limiter
                a.s_ibinnew[s] = (int8_t)ibini;
                st_nbinmax = max(st_nbinmax, ibini); st_ncheckbin += 1;
            } else {
                st_dtc = fmin(st_dtc, dtc);
                st_dtf = fmin(st_dtf, fmin(fmin(dtf, dtdrag), dtclean));
                st_dtmin = fmin(st_dtmin, dtc); st_dtmax = fmax(st_dtmax, dtc);
            }
        }
        __syncwarp();
    }
    st_dtc = warp_min(st_dtc); st_dtf = warp_min(st_dtf); st_dtmin = warp_min(st_dtmin); st_dtmax = warp_max(st_dtmax);
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) { st_pairs += __shfl_xor_sync(FULLMASK, st_pairs, sft); st_trial += __shfl_xor_sync(FULLMASK, st_trial, sft); }
    if (XTRA) {
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) { st_nbinmax = max(st_nbinmax, __shfl_xor_sync(FULLMASK, st_nbinmax, sft)); st_ncheckbin += __shfl_xor_sync(FULLMASK, st_ncheckbin, sft); }
        if (lane == 0 && st_ncheckbin) { atomicMax(&a.cnt[CNT_NBINMAX], (unsigned long long)st_nbinmax); atomicAdd(&a.cnt[CNT_NCHECKBIN], (unsigned long long)st_ncheckbin); }
    }
    if (lane == 0) {
        atomicAdd(&a.cnt[CNT_NPAIRS], st_pairs); atomicAdd(&a.cnt[CNT_NTRIAL], st_trial);
        atomic_min_pos(&a.dscal[DS_DTCOURANT], st_dtc); atomic_min_pos(&a.dscal[DS_DTFORCE], st_dtf);
        atomic_min_pos(&a.dscal[DS_DTMINI], st_dtmin); atomic_max_pos(&a.dscal[DS_DTMAXI], st_dtmax);
    }
}

__global__ void k_scatter_force(int64_t nlive, const int *__restrict__ perm, const int *__restrict__ s_done, const double4 *__restrict__ s_fxyzu,
                                const double4 *__restrict__ s_dB, const float *__restrict__ s_divvf, const float *__restrict__ s_divBsymm,
                                double *__restrict__ fxyzu, double *__restrict__ dBevol, float *__restrict__ divcurlv, float *__restrict__ divBsymm, int nvu,
                                int mhd, int gravity, int dust, int ind_ts, int driving, const float *__restrict__ s_poten, float *__restrict__ poten,
                                const double *__restrict__ s_tstop, double *__restrict__ tstop, const int8_t *__restrict__ s_ibinnew,
                                int8_t *__restrict__ ibin, const int *__restrict__ s_wake, int8_t *__restrict__ ibin_wake)
{
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nlive) return;
    const int d = s_done[s];
    const int i = perm[s];
    if (ind_ts) ibin_wake[i] = (int8_t)s_wake[s];                    // neighbours of active particles, active or not
    if (d == 0) return;
    if (gravity) poten[i] = s_poten[s];
    if (dust) tstop[i] = s_tstop[s];
    if (ind_ts) ibin[i] = s_ibinnew[s];
    const double4 f = s_fxyzu[s];
    double *fi = fxyzu + (size_t)nvu * i;
    if (driving) { fi[0] += f.x; fi[1] += f.y; fi[2] += f.z; }      // the driving routine initialised the force (force.F90:2969-2973)
    else { fi[0] = f.x; fi[1] = f.y; fi[2] = f.z; }
    if (d == 2) {
        if (nvu >= 4) fi[3] = f.w;
        divcurlv[i] = s_divvf[s];                                    // force.F90:2999
        if (mhd) { reinterpret_cast<double4 *>(dBevol)[i] = s_dB[s]; divBsymm[i] = s_divBsymm[s]; }
    }
}

// two-sector records (iso1_derive): isothermal equation of state whose P and c_s this library computed itself (cons2prim_run since
// the last upload of eos_vars), energy not evolved, no MHD, no self-gravity
static bool force_iso1(const sphgpu_ctx *c)
{
    const sphgpu_params &p = c->hp.p;
    return c->eos_on_device && p.ieos == 1 && c->hp.nvu < 4 && !p.mhd && !p.gravity && !c->no_iso1;
}

// grid < 0: only query the resident CTAs/SM of the instantiation; otherwise launch on `grid` CTAs
template <int K, bool PERIODIC, bool MHD>
int launch_force_general(sphgpu_ctx *c, const ForceArgs &a, int grid)
{
    const size_t smem = 4 * sizeof(WarpSharedGeneral);
    cudaFuncSetAttribute(k_force<K, PERIODIC, MHD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (grid < 0) {
        int bps = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_force<K, PERIODIC, MHD, true>, 128, smem);
        return bps < 1 ? 1 : bps;
    }
    k_force<K, PERIODIC, MHD, true><<<grid, 128, smem, c->stream>>>(a, c->hp);
    c->launches++;
    return 0;
}
template <int K, bool PERIODIC, bool MHD, bool ADIA, bool GRAV, bool INDTS, bool BIG, bool ISO1>
int launch_force_fast4(sphgpu_ctx *c, const ForceArgs &a, int grid)
{
    const size_t smem = 4 * sizeof(typename ForceFastSharedT<MHD, BIG>::type);
    cudaFuncSetAttribute(k_force_fast<K, PERIODIC, MHD, ADIA, GRAV, INDTS, BIG, ISO1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (grid < 0) {
        int bps = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_force_fast<K, PERIODIC, MHD, ADIA, GRAV, INDTS, BIG, ISO1>, 128, smem);
        return bps < 1 ? 1 : bps;
    }
    k_force_fast<K, PERIODIC, MHD, ADIA, GRAV, INDTS, BIG, ISO1><<<grid, 128, smem, c->stream>>>(a, c->hp);
    c->launches++;
    return 0;
}
template <int K, bool PERIODIC, bool MHD, bool ADIA, bool GRAV, bool INDTS, bool BIG>
int launch_force_fast3(sphgpu_ctx *c, const ForceArgs &a, int grid)
{
    if (!MHD && !ADIA && !GRAV && force_iso1(c)) return launch_force_fast4<K, PERIODIC, MHD, ADIA, GRAV, INDTS, BIG, !MHD && !ADIA && !GRAV>(c, a, grid);
    return launch_force_fast4<K, PERIODIC, MHD, ADIA, GRAV, INDTS, BIG, false>(c, a, grid);
}
template <int K, bool PERIODIC, bool MHD, bool ADIA, bool GRAV, bool INDTS>
int launch_force_fast2(sphgpu_ctx *c, const ForceArgs &a, int grid)
{
    if (c->dens_trial_max > FORCE_ROUND && c->dens_trial_hint > 0.8 * FORCE_ROUND) return launch_force_fast3<K, PERIODIC, MHD, ADIA, GRAV, INDTS, true>(c, a, grid);
    return launch_force_fast3<K, PERIODIC, MHD, ADIA, GRAV, INDTS, false>(c, a, grid);
}
template <int K, bool PERIODIC, bool MHD, bool ADIA, bool GRAV>
int launch_force_fast(sphgpu_ctx *c, const ForceArgs &a, int grid)
{
    if (c->hp.p.ind_timesteps) return launch_force_fast2<K, PERIODIC, MHD, ADIA, GRAV, true>(c, a, grid);
    return launch_force_fast2<K, PERIODIC, MHD, ADIA, GRAV, false>(c, a, grid);
}

// general path: anything beyond all-gas hydro / MHD (boundary or dust particles, gravity, individual timesteps, disc viscosity)
bool force_is_general(const sphgpu_ctx *c)
{
    const sphgpu_params &p = c->hp.p;
    return p.dust || p.disc_viscosity || c->multitype || c->force_general;
}

template <int K, bool PERIODIC>
int dispatch_force2(sphgpu_ctx *c, const ForceArgs &a, int grid)
{
    const sphgpu_params &p = c->hp.p;
    if (force_is_general(c)) return p.mhd ? launch_force_general<K, PERIODIC, true>(c, a, grid) : launch_force_general<K, PERIODIC, false>(c, a, grid);
    const bool adia = c->hp.nvu >= 4;
    if (p.gravity) {       // self-gravity is not periodic (gravity_run rejects it): PERIODIC instantiations are not needed
        if (PERIODIC) return launch_force_general<K, PERIODIC, false>(c, a, grid);
        if (p.mhd) return adia ? launch_force_fast<K, false, true, true, true>(c, a, grid) : launch_force_fast<K, false, true, false, true>(c, a, grid);
        return adia ? launch_force_fast<K, false, false, true, true>(c, a, grid) : launch_force_fast<K, false, false, false, true>(c, a, grid);
    }
    if (p.mhd) return adia ? launch_force_fast<K, PERIODIC, true, true, false>(c, a, grid) : launch_force_fast<K, PERIODIC, true, false, false>(c, a, grid);
    return adia ? launch_force_fast<K, PERIODIC, false, true, false>(c, a, grid) : launch_force_fast<K, PERIODIC, false, false, false>(c, a, grid);
}

int dispatch_force(sphgpu_ctx *c, const ForceArgs &a, int grid)
{
    const sphgpu_params &p = c->hp.p;
    if (p.kernel == 0) return p.periodic ? dispatch_force2<0, true>(c, a, grid) : dispatch_force2<0, false>(c, a, grid);
    return p.periodic ? dispatch_force2<1, true>(c, a, grid) : dispatch_force2<1, false>(c, a, grid);
}

}  // namespace

static inline int nblk(int64_t n, int b) { return (int)((n + b - 1) / b); }

int force_run(sphgpu_ctx *c, int icall, double dt, sphgpu_scalars *out)
{
    (void)dt;
    if (!c->tree_valid) { c->err = "force: build_tree has not been called"; return SPHGPU_ERR_STATE; }
    const int64_t n = c->npart, nl = c->nlive;
    const sphgpu_params &p = c->hp.p;
    if (p.gravity) TRY(gravity_run(c));                                    // far field -> c->gacc (canonical order)
    CUDA_TRY(c, c->vel4.ensure(n)); CUDA_TRY(c, c->frecC.ensure(n)); CUDA_TRY(c, c->frecD.ensure(n)); if (p.mhd) CUDA_TRY(c, c->frecE.ensure(n));
    CUDA_TRY(c, c->hnew.ensure(2 * n));                                    // reused as double2 hinv
    CUDA_TRY(c, c->s_fxyzu.ensure(n)); if (p.mhd) { CUDA_TRY(c, c->s_dB.ensure(n)); CUDA_TRY(c, c->s_divBsymm.ensure(n)); }
    CUDA_TRY(c, c->s_divvf.ensure(n)); CUDA_TRY(c, c->s_nneigh.ensure(n));
    if (p.gravity) { CUDA_TRY(c, c->s_gsoft.ensure(n)); CUDA_TRY(c, c->s_poten.ensure(n)); }
    if (p.dust) { CUDA_TRY(c, c->s_dvdx.ensure(9 * n)); CUDA_TRY(c, c->s_tstop.ensure(n)); }
    if (p.ind_timesteps) { CUDA_TRY(c, c->s_ibin.ensure(n)); CUDA_TRY(c, c->s_ibinold.ensure(n)); CUDA_TRY(c, c->s_ibinnew.ensure(n)); CUDA_TRY(c, c->s_wake.ensure(n)); }
    const bool fast = !force_is_general(c);
    if (fast) CUDA_TRY(c, c->frec.ensure(5 * (size_t)n));
    ForceArgs a;
    memset(&a, 0, sizeof a);
    const int grid = c->numSMs * dispatch_force(c, a, -1);
    CUDA_TRY(c, c->stage_idx.ensure((size_t)grid * 4 * c->scratch_per_warp));
    double2 *hinv = reinterpret_cast<double2 *>(c->hnew.p);
    a.nodes = c->nodesf.p; a.cells = c->cells.p; a.ncells = (int)c->ncells; a.groups = c->groups.p; a.ngroups = (int)c->ngroups;
    a.pos4 = c->pos4.p; a.vel4 = c->vel4.p; a.recC = c->frecC.p; a.recD = c->frecD.p; a.recE = c->frecE.p; a.hinv = hinv; a.stype = c->stype.p; a.perm = c->perm.p;
    a.s_fxyzu = c->s_fxyzu.p; a.s_dB = c->s_dB.p; a.s_divvf = c->s_divvf.p; a.s_divBsymm = c->s_divBsymm.p; a.s_done = c->s_nneigh.p;
    a.multitype = c->multitype ? 1 : 0; a.max_leaf = c->max_leaf; a.class_mask = c->class_mask;
    a.cnt = c->counters.p; a.dscal = c->dscal.p; a.icall = icall;
    a.frec = c->frec.p;
    a.gsoft = c->s_gsoft.p; a.dvdx9 = c->s_dvdx.p; a.gacc = c->gacc.p; a.s_poten = c->s_poten.p; a.s_tstop = c->s_tstop.p;
    a.s_ibinold = c->s_ibinold.p; a.s_ibin = c->s_ibin.p; a.s_wake = c->s_wake.p; a.s_ibinnew = c->s_ibinnew.p;
    a.hscale = c->hscale;
    a.iso1 = (fast && force_iso1(c)) ? (p.const_av ? 2 : 1) : 0;
    a.iso_pmass = p.massoftype[IGAS]; a.iso_hfact = p.hfact; a.iso_polyk = p.polyk; a.iso_cs = sqrt(p.polyk); a.iso_cnormk = c->hp.kc.cnormk; a.iso_alpha = p.alpha;
    a.nbinmax = c->nbinmax; a.ibinnow_m1 = c->ibinnow - 1; a.istepfrac = c->istepfrac;
    if (p.ind_timesteps && refcompat_on(c) && c->dens_valid) {           // reference-compatible neighbour sets (common.cuh: refcompat)
        TRY(refcompat_prepare(c));
        a.refnodes = c->ref_nodes.p; a.refleaf = c->ref_leaf_sorted.p; a.ref_radkern = c->hp.kc.radkern; a.ref_tree_acc2 = p.tree_accuracy * p.tree_accuracy;
        a.ref_gravity = p.gravity;
    }
    if (c->wl_force_ok && c->hscale <= c->wl_cover) {                      // the lists of the density pass still cover every pair
        a.wl.list = c->wl_list.p; a.wl.ncl = c->wl_ncl.p; a.wl.reach = c->wl_reach.p; a.wl.cap = c->walk_cap;
        a.wl.order = c->wl_ordered ? c->wl_order.p : nullptr;
    } else {
        const double R = (p.kernel == 0 ? SphKern<0>::radkern : SphKern<1>::radkern);
        TRY(walk_lists_run(c, true, R * c->hscale, R * c->hscale, a.wl));
        TRY(walk_order_run(c, a.wl));
        c->wl_force_ok = true; c->wl_cover = c->hscale; c->wl_ordered = a.wl.order != nullptr;
    }
    unsigned long long hc[16]; double hd[4];
    for (int attempt = 0;; attempt++) {
        CUDA_TRY(c, cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 16, c->stream));
        const double init[4] = {1.e29, 1.e29, 1.e29, 0.};
        CUDA_TRY(c, cudaMemcpyAsync(c->dscal.p + DS_DTCOURANT, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
        k_force_prep<<<nblk(nl, 256), 256, 0, c->stream>>>(nl, c->perm.p, c->pos4.p, c->stype.p, c->vxyzu.p, c->Bevol.p, c->eos_vars.p, c->alphaind.p, c->gradh.p,
                                                           c->vel4.p, c->frecC.p, c->frecD.p, c->frecE.p, hinv, c->s_nneigh.p, c->hp, c->counters.p,
                                                           c->s_gsoft.p, c->dvdx.p, c->s_dvdx.p, c->ibin.p, c->ibin_old.p, c->ibin_wake.p, c->s_ibin.p,
                                                           c->s_ibinold.p, c->s_wake.p, fast ? c->frec.p : nullptr, a.iso1);
        c->launches++;
        a.stage_idx = c->stage_idx.p; a.scratch_per_warp = c->scratch_per_warp;
        cudaEventRecord(c->ev[10], c->stream);
        dispatch_force(c, a, grid);
        cudaEventRecord(c->ev[11], c->stream);
        CUDA_TRY(c, cudaMemcpyAsync(hc, c->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(hd, c->dscal.p + DS_DTCOURANT, sizeof(hd), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaGetLastError());
        // cell lists too short for some group (strongly non-uniform h): grow them and repeat -- nothing has been scattered yet
        if (hc[CNT_ERR] == SPHGPU_ERR_OVERFLOW && attempt < 3 && c->scratch_per_warp < (1 << 20)) {
            c->scratch_per_warp *= 8;
            CUDA_TRY(c, c->stage_idx.ensure((size_t)grid * 4 * c->scratch_per_warp));
            continue;
        }
        break;
    }
    { float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[10], c->ev[11]); c->ms_kernel[1] = ms; }
    if (hc[CNT_ERR] == SPHGPU_ERR_NEGH) {
        char buf[128]; snprintf(buf, sizeof buf, "force: negative smoothing length on particle %llu", hc[CNT_ERRID]);
        c->err = buf; return SPHGPU_ERR_NEGH;
    }
    if (hc[CNT_ERR]) { c->err = "force: neighbour scratch overflow (raise scratch_per_warp)"; return (int)hc[CNT_ERR]; }
    k_scatter_force<<<nblk(nl, 256), 256, 0, c->stream>>>(nl, c->perm.p, c->s_nneigh.p, c->s_fxyzu.p, c->s_dB.p, c->s_divvf.p, c->s_divBsymm.p, c->fxyzu.p,
                                                          c->dBevol.p, c->divcurlv.p, c->divBsymm.p, c->hp.nvu, p.mhd, p.gravity, p.dust, p.ind_timesteps, p.driving,
                                                          c->s_poten.p, c->poten.p, c->s_tstop.p, c->tstop.p, c->s_ibinnew.p, c->ibin.p, c->s_wake.p,
                                                          c->ibin_wake.p);
    c->launches++;
    sphgpu_scalars &sc = c->last_force;
    memset(&sc, 0, sizeof sc);
    sc.dtcourant = hd[0]; sc.dtforce = hd[1]; sc.dtmini = hd[2]; sc.dtmaxi = hd[3];
    sc.npairs_force = (int64_t)hc[CNT_NPAIRS];
    sc.nbinmaxnew = p.ind_timesteps ? (hc[CNT_NCHECKBIN] ? (int64_t)hc[CNT_NBINMAX] : (int64_t)c->nbinmax) : 0;   // force.F90:856-861
    if (out) {
        sphgpu_scalars o = c->last_dens;
        o.dtcourant = sc.dtcourant; o.dtforce = sc.dtforce; o.dtmini = sc.dtmini; o.dtmaxi = sc.dtmaxi; o.npairs_force = sc.npairs_force;
        o.nbinmaxnew = sc.nbinmaxnew; o.npairs_gravity = c->npairs_gravity; o.nm2l = c->nm2l;
        *out = o;
    }
    return SPHGPU_OK;
}
