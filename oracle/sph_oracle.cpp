/*
 * sph_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see sph_oracle.h).
 *
 * CPU restatement of Phantom's derivs hot path.  Every routine cites the
 * reference file:line it follows.  The arithmetic keeps the reference's
 * operation order and its real*4 roundings; compile with
 *     g++ -O2 -ffp-contract=off -fopenmp
 * (the reference is built with gfortran -O3 -fdefault-real-8, no -ffast-math,
 * no FMA contraction: build/Makefile_defaults_gfortran:15-23).
 *
 * "parity unpinned": no Fortran compiler is available, so the reference
 * binary itself was never run against this file; it is pinned against the
 * known answers of the reference's own tests (tests/test_oracle_*.py).
 */
#include "sph_oracle.h"
#include <cmath>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int igas = 1, iboundary = 3, idust = 7;          // part.F90:428-438
constexpr int minpart = 10;                                // config.F90:113
constexpr int maxdepth = 32;                               // kdtree.F90:42
constexpr int irootnode = 1;                               // kdtree.F90:40
constexpr int isizecellcache = 1000;                       // dens.F90:99
constexpr int maxcellcache = 1000;                         // force.F90:58
constexpr int maxdensits = 100;                            // dens.F90:101
constexpr int lenfgrav = 20;                               // dtype_kdtree.F90:21
constexpr double bignumber = 1.e29;                        // timestep.f90
constexpr double pi = 3.14159265358979323846264338327950288419716939937510582097494459; // physcon.f90

// ---- rhosum indices (dens.F90:51-95), 0-based -------------------------------
enum {
    irhoi = 0, igradhi, igradsofti, idivvi,
    idvxdxi, idvxdyi, idvxdzi, idvydxi, idvydyi, idvydzi, idvzdxi, idvzdyi, idvzdzi,
    idaxdxi, idaxdyi, idaxdzi, idaydxi, idaydyi, idaydzi, idazdxi, idazdyi, idazdzi,
    irxxi, irxyi, irxzi, iryyi, iryzi, irzzi,
    idivBi, idBxdxi, idBxdyi, idBxdzi, idBydxi, idBydyi, idBydzi, idBzdxi, idBzdyi, idBzdzi,
    irhodusti, maxrhosum
};
// ---- fsum indices (force.F90:141-182) with maxdustsmall=1 ---------------------
enum {
    f_fx = 0, f_fy, f_fz, f_pot, f_drhodt, f_dudtdiss, f_dendtdiss, f_divBsym,
    f_dBx, f_dBy, f_dBz, f_divBdiff, f_fdragx, f_fdragy, f_fdragz, maxfsum
};

struct Node {  // kdnode, dtype_kdtree.F90:53-68
    double xcen[3], size, hmax, mass;
    int leftchild, rightchild, parent;
    double quads[6];
    int tobecached, cached;
};

struct KernelConsts { double radkern, radkern2, cnormk, wab0, gradh0, dphidh0, cnormk_drag, hfact_default; };

inline KernelConsts kernel_consts(int k)
{
    KernelConsts c;
    if (k == 0) {   // kernel_cubic.f90:23-30
        c.radkern = 2.; c.radkern2 = 4.; c.cnormk = 1. / pi; c.wab0 = 1.; c.gradh0 = -3. * c.wab0;
        c.dphidh0 = 1.4; c.cnormk_drag = 10. / (9. * pi); c.hfact_default = 1.2;
    } else {        // kernel_quintic.f90:23-30
        c.radkern = 3.; c.radkern2 = 9.; c.cnormk = 1. / (120. * pi); c.wab0 = 66.; c.gradh0 = -3. * c.wab0;
        c.dphidh0 = 239. / 210.; c.cnormk_drag = 1. / (168. * pi); c.hfact_default = 1.0;
    }
    return c;
}

inline double pow2(double x) { return x * x; }
inline double pow3(double x) { return x * x * x; }
inline double pow4(double x) { double x2 = x * x; return x2 * x2; }
inline double pow5(double x) { double x2 = x * x; return x2 * x2 * x; }

// get_kernel (kernel_cubic.f90:35-52, kernel_quintic.f90:33-56)
inline void get_kernel(int k, double q2, double q, double &w, double &gr)
{
    if (k == 0) {
        if (q < 1.) { w = 0.75 * q2 * q - 1.5 * q2 + 1.; gr = q * (2.25 * q - 3.); }
        else if (q < 2.) { w = -0.25 * pow3(q - 2.); gr = -0.75 * pow2(q - 2.); }
        else { w = 0.; gr = 0.; }
    } else {
        if (q < 1.) { double q4 = q2 * q2; w = -10. * q4 * q + 30. * q4 - 60. * q2 + 66.; gr = q * (-50. * q2 * q + 120. * q2 - 120.); }
        else if (q < 2.) { w = -pow5(q - 3.) + 6. * pow5(q - 2.); gr = -5. * pow4(q - 3.) + 30. * pow4(q - 2.); }
        else if (q < 3.) { w = -pow5(q - 3.); gr = -5. * pow4(q - 3.); }
        else { w = 0.; gr = 0.; }
    }
}
inline double grkern(int k, double q2, double q) { double w, g; get_kernel(k, q2, q, w, g); return g; }
inline double wkern(int k, double q2, double q) { double w, g; get_kernel(k, q2, q, w, g); return w; }

// get_kernel_grav1 dphidh part (kernel_cubic.f90:80-103, kernel_quintic.f90:88-123)
inline double dphidh_kernel(int k, double q2, double q)
{
    if (k == 0) {
        if (q < 1.) { double q4 = q2 * q2; return -0.6 * q4 * q + 1.5 * q4 - 2. * q2 + 1.4; }
        else if (q < 2.) { double q4 = q2 * q2; return 0.2 * q4 * q - 1.5 * q4 + 4. * q2 * q - 4. * q2 + 1.6; }
        return 0.;
    } else {
        double q4 = q2 * q2, q6 = q4 * q2;
        if (q < 1.) return q6 * q / 21. - q6 / 6. + q4 / 2. - 11. * q2 / 10. + 239. / 210.;
        else if (q < 2.) return -q6 * q / 42. + q6 / 4. - q4 * q + 7. * q4 / 4. - 5. * q2 * q / 6. - 17. * q2 / 20. + 473. / 420.;
        else if (q < 3.) return q6 * q / 210. - q6 / 12. + 3. * q4 * q / 5. - 9. * q4 / 4. + 9. * q2 * q / 2. - 81. * q2 / 20. + 243. / 140.;
        return 0.;
    }
}

// kernel_softening (kernel_cubic.f90:105-124, kernel_quintic.f90:125-156)
inline void kernel_softening(int k, double q2, double q, double &potensoft, double &fsoft)
{
    if (k == 0) {
        if (q < 1.) {
            double q4 = q2 * q2;
            potensoft = q4 * q / 10. - 3. * q4 / 10. + 2. * q2 / 3. - 7. / 5.;
            fsoft = q * (15. * q2 * q - 36. * q2 + 40.) / 30.;
        } else if (q < 2.) {
            double q4 = q2 * q2, q6 = q4 * q2;
            potensoft = (q * (-q4 * q + 9. * q4 - 30. * q2 * q + 40. * q2 - 48.) + 2.) / (30. * q);
            fsoft = (-5. * q6 + 36. * q4 * q - 90. * q4 + 80. * q2 * q - 2.) / (30. * q2);
        } else { potensoft = -1. / q; fsoft = 1. / q2; }
    } else {
        if (q < 1.) {
            double q4 = q2 * q2, q6 = q4 * q2;
            potensoft = -q6 * q / 168. + q6 / 42. - q4 / 10. + 11. * q2 / 30. - 239. / 210.;
            fsoft = q * (-35. * q4 * q + 120. * q4 - 336. * q2 + 616.) / 840.;
        } else if (q < 2.) {
            double q4 = q2 * q2, q6 = q4 * q2, q8 = q6 * q2;
            potensoft = (q * (5. * q6 * q - 60. * q6 + 280. * q4 * q - 588. * q4 + 350. * q2 * q + 476. * q2 - 1892.) - 5.) / (1680. * q);
            fsoft = (35. * q8 - 360. * q6 * q + 1400. * q6 - 2352. * q4 * q + 1050. * q4 + 952. * q2 * q + 5.) / (1680. * q2);
        } else if (q < 3.) {
            double q4 = q2 * q2, q6 = q4 * q2, q8 = q6 * q2;
            potensoft = (q * (-q6 * q + 20. * q6 - 168. * q4 * q + 756. * q4 - 1890. * q2 * q + 2268. * q2 - 2916.) + 507.) / (1680. * q);
            fsoft = (-7. * q8 + 120. * q6 * q - 840. * q6 + 3024. * q4 * q - 5670. * q4 + 4536. * q2 * q - 507.) / (1680. * q2);
        } else { potensoft = -1. / q; fsoft = 1. / q2; }
    }
}

// wkern_drag (kernel_cubic.f90:146-158, kernel_quintic.f90:188-203)
inline double wkern_drag(int k, double q2, double q)
{
    if (k == 0) {
        if (q < 1.) return q2 * (0.75 * q2 * q - 1.5 * q2 + 1.);
        else if (q < 2.) return -0.25 * q2 * pow3(q - 2.);
        return 0.;
    } else {
        if (q < 1.) { double q4 = q2 * q2; return q2 * (-10. * q4 * q + 30. * q4 - 60. * q2 + 66.); }
        else if (q < 2.) return q2 * (-pow5(q - 3.) + 6. * pow5(q - 2.));
        else if (q < 3.) return -q2 * pow5(q - 3.);
        return 0.;
    }
}

// get_partinfo (part.F90:1026-1067), two-fluid dust with one large grain type
inline void get_partinfo(const oracle_params &p, int8_t iphasei, bool &isactive, bool &isgas, bool &isdust, int &itype)
{
    if (iphasei >= 0) { isactive = true; itype = iphasei; }
    else { isactive = false; itype = -iphasei; }
    isgas = (itype == igas || itype == iboundary);
    isdust = p.dust ? (itype == idust) : false;
    if (itype == iboundary) {
        if (p.set_boundaries_to_active) { isactive = true; itype = igas; }
        else isactive = false;
    }
}
inline int iamtype(int8_t iphasei) { return std::abs((int)iphasei); }          // part.F90:1078
inline int ibasetype(int itype) { return itype == iboundary ? igas : itype; }   // part.F90:1118
inline bool iamboundary(int itype) { return itype == iboundary; }

}  // namespace

struct oracle_ctx {
    oracle_params p;
    KernelConsts kc;
    std::string err;
    // tree storage (kdtree.F90:31-35, neigh_kdtree.f90:28-35, part.F90 treecache)
    int64_t npart = 0, ncells = 0, ncellsmax = 0;
    int maxlevel_indexed = 0, maxlevel = 0, minlevel = 0;
    std::vector<Node> node;
    std::vector<int> leaf_is_active, inodeparts;
    std::vector<int> inoderange;   // (2, ncellsmax+1) 1-based ranges into inodeparts/treecache
    std::vector<double> treecache; // (5, npart)
    std::vector<double> fnodecache;
    double dxbound, dybound, dzbound, hdlx, hdly, hdlz;
    // density statistics (dens.F90:104-106)
    int64_t nneightry = 0, nneighact = 0, maxneightry = 0, ncalc = 0, nptot = -1, ncalls_neigh = 0;
    int maxneighact = 0;

    // helpers with the module-level constants bound
    inline double rhoh(double hi, double pmassi) const { return pmassi * pow3(p.hfact / std::fabs(hi)); }  // part.F90:779
    inline double dhdrho(double hi, double pmassi) const { double rhoi = rhoh(hi, pmassi); return -hi / (3. * rhoi); }  // part.F90:791
    inline double hrho(double rhoi, double pmassi) const { return p.hfact * std::pow(pmassi / std::fabs(rhoi), 1.0 / 3.0); } // part.F90:845
    inline int nvu() const { return p.isothermal ? 3 : 4; }
    inline int ngradh() const { return p.gravity ? 2 : 1; }
    inline int nalpha() const { return p.const_av ? 0 : 3; }
    void set_bounds()
    {
        dxbound = p.xmax - p.xmin; dybound = p.ymax - p.ymin; dzbound = p.zmax - p.zmin;  // boundary.f90:102-108
        hdlx = 0.5 * dxbound; hdly = 0.5 * dybound; hdlz = 0.5 * dzbound;
    }
};

namespace {

// get_sep (kdtree.F90:1468-1501)
inline void get_sep(const oracle_ctx &c, const double *x1, const double *x2, double &dx, double &dy, double &dz,
                    double &xoffset, double &yoffset, double &zoffset, double &r2)
{
    xoffset = 0.; yoffset = 0.; zoffset = 0.;
    dx = x1[0] - x2[0]; dy = x1[1] - x2[1]; dz = x1[2] - x2[2];
    if (c.p.periodic) {
        if (std::fabs(dx) > c.hdlx) { xoffset = std::copysign(c.dxbound, dx); dx = dx - xoffset; }
        if (std::fabs(dy) > c.hdly) { yoffset = std::copysign(c.dybound, dy); dy = dy - yoffset; }
        if (std::fabs(dz) > c.hdlz) { zoffset = std::copysign(c.dzbound, dz); dz = dz - zoffset; }
    }
    r2 = dx * dx + dy * dy + dz * dz;
}

// ------------------------------------------------------------------------------
//  kd-tree build: maketree (kdtree.F90:117-314), serial semantics (numthreads=1)
// ------------------------------------------------------------------------------
struct BuildStack { int node, parent, level, npnode; double xmin[3], xmax[3]; };

#define TC(k, i) c.treecache[5 * (size_t)((i)-1) + (k)-1]          /* treecache(k,i), 1-based */
#define IR(k, n) c.inoderange[2 * (size_t)(n) + (k)-1]             /* inoderange(k,n) */

// sort_particles_in_cell (kdtree.F90:937-1001)
void sort_particles_in_cell(oracle_ctx &c, int iaxis, int imin, int imax, int &min_l, int &max_l, int &min_r, int &max_r,
                            int &nl, int &nr, double xpivot)
{
    int i = imin, j = imax;
    double xi_coord = TC(iaxis, i), xj_coord = TC(iaxis, j);
    bool i_lt_pivot = xi_coord <= xpivot, j_lt_pivot = xj_coord <= xpivot;
    while (i < j) {
        if (i_lt_pivot) {
            i = i + 1; xi_coord = TC(iaxis, i); i_lt_pivot = xi_coord <= xpivot;
        } else {
            if (!j_lt_pivot) {
                j = j - 1; xj_coord = TC(iaxis, j); j_lt_pivot = xj_coord <= xpivot;
            } else {
                std::swap(c.inodeparts[i - 1], c.inodeparts[j - 1]);
                for (int k = 1; k <= 5; k++) std::swap(TC(k, i), TC(k, j));
                i = i + 1; j = j - 1;
                xi_coord = TC(iaxis, i); xj_coord = TC(iaxis, j);
                i_lt_pivot = xi_coord <= xpivot; j_lt_pivot = xj_coord <= xpivot;
            }
        }
    }
    if (!i_lt_pivot) i = i - 1;
    if (j_lt_pivot) j = j + 1;
    min_l = imin; max_l = i; min_r = j; max_r = imax;
    nl = max_l - min_l + 1; nr = max_r - min_r + 1;
}

// construct_node (kdtree.F90:531-929), non-MPI, non-APR, no sink tree
int construct_node(oracle_ctx &c, int nnode, int mymum, int level, double *xmini, double *xmaxi, int npnode,
                   int &il, int &ir, int &nl, int &nr, double *xminl, double *xmaxl, double *xminr, double *xmaxr,
                   bool &wassplit)
{
    Node &nodeentry = c.node[nnode];
    bool nodeisactive = false;
    int npcounter = 0;
    if (IR(1, nnode) > 0) {
        for (int i = IR(1, nnode); i <= IR(2, nnode); i++)
            if (c.inodeparts[i - 1] > 0) { nodeisactive = true; break; }
        npcounter = IR(2, nnode) - IR(1, nnode) + 1;
    }
    if (npcounter != npnode) { c.err = "maketree: expected number of particles in node differed from actual number"; return 1; }
    ir = 0; il = 0; nl = 0; nr = 0; wassplit = false;
    if (npnode < 1) return 0;

    double hmax = 0., xcofm = 0., ycofm = 0., zcofm = 0.;
    // kdtree.F90:612-624 : centre of mass relative to the gas particle mass
    double pmassi = c.p.massoftype[igas], dfac, totmass_node = 0.;
    if (pmassi > 0.) dfac = 1. / pmassi;
    else {
        // massoftype(maxloc(npartoftype(2:maxtypes))+1): first non-gas type with mass > 0 stands in
        pmassi = 0.;
        for (int t = 2; t < ORACLE_MAXTYPES; t++) if (c.p.massoftype[t] > 0.) { pmassi = c.p.massoftype[t]; break; }
        dfac = pmassi > 0. ? 1. / pmassi : 1.;
    }
    int i1 = IR(1, nnode);
    for (int i = i1; i <= i1 + npnode - 1; i++) {     // kdtree.F90:654-666
        double xi = TC(1, i), yi = TC(2, i), zi = TC(3, i), hi = TC(4, i);
        pmassi = TC(5, i);
        double fac = pmassi * dfac;
        hmax = std::max(hmax, hi);
        totmass_node = totmass_node + pmassi;
        xcofm = xcofm + fac * xi; ycofm = ycofm + fac * yi; zcofm = zcofm + fac * zi;
    }
    double xyzcofm[3] = {xcofm, ycofm, zcofm};
    if (totmass_node > 0.) { double d = totmass_node * dfac; xyzcofm[0] /= d; xyzcofm[1] /= d; xyzcofm[2] /= d; }
    if (totmass_node <= 0.) { c.err = "mtree: totmass_node==0"; return 1; }
    double x0[3] = {xyzcofm[0], xyzcofm[1], xyzcofm[2]};
    double r2max = 0., quads[6] = {0, 0, 0, 0, 0, 0};
    for (int i = i1; i <= i1 + npnode - 1; i++) {     // kdtree.F90:734-752
        double dx = TC(1, i) - x0[0], dy = TC(2, i) - x0[1], dz = TC(3, i) - x0[2];
        double dr2 = dx * dx + dy * dy + dz * dz;
        r2max = std::max(r2max, dr2);
        if (c.p.gravity) {
            pmassi = TC(5, i);
            quads[0] += pmassi * (dx * dx); quads[1] += pmassi * (dx * dy); quads[2] += pmassi * (dx * dz);
            quads[3] += pmassi * (dy * dy); quads[4] += pmassi * (dy * dz); quads[5] += pmassi * (dz * dz);
        }
    }
    nodeentry.xcen[0] = x0[0]; nodeentry.xcen[1] = x0[1]; nodeentry.xcen[2] = x0[2];
    nodeentry.size = std::sqrt(r2max) + DBL_EPSILON;   // kdtree.F90:782
    nodeentry.hmax = hmax;
    nodeentry.parent = mymum;
    nodeentry.mass = totmass_node;
    for (int k = 0; k < 6; k++) nodeentry.quads[k] = quads[k];
    nodeentry.tobecached = 1; nodeentry.cached = 0;

    wassplit = (npnode > minpart);                      // kdtree.F90:792
    if (!wassplit) {
        nodeentry.leftchild = 0; nodeentry.rightchild = 0;
        c.maxlevel = std::max(level, c.maxlevel);
        c.minlevel = std::min(level, c.minlevel);
        if (c.p.ind_timesteps) c.leaf_is_active[nnode] = nodeisactive ? 1 : -1;   // kdtree.F90:801-813
        else c.leaf_is_active[nnode] = 1;
    } else {
        // iaxis = maxloc(xmaxi - xmini,1): first maximum (kdtree.F90:815)
        int iaxis = 1; double best = xmaxi[0] - xmini[0];
        for (int k = 1; k < 3; k++) if (xmaxi[k] - xmini[k] > best) { best = xmaxi[k] - xmini[k]; iaxis = k + 1; }
        double xpivot = xyzcofm[iaxis - 1];
        if (level < c.maxlevel_indexed) { il = 2 * nnode; ir = il + 1; }  // kdtree.F90:820-822
        else {
            c.ncells = c.ncells + 2; ir = (int)c.ncells; il = ir - 1;
            if (ir > c.ncellsmax) { c.err = "maketree: number of nodes exceeds array dimensions"; return 1; }
        }
        nodeentry.leftchild = il; nodeentry.rightchild = ir;
        c.leaf_is_active[nnode] = 0;
        sort_particles_in_cell(c, iaxis, IR(1, nnode), IR(2, nnode), IR(1, il), IR(2, il), IR(1, ir), IR(2, ir), nl, nr, xpivot);
        if (nr + nl != npnode) { c.err = "maketree: number of left + right != parent while splitting"; return 1; }
        if (nl == npnode || nr == npnode) {            // kdtree.F90:856-865
            nl = npnode / 2;
            IR(1, il) = IR(1, nnode); IR(2, il) = IR(1, nnode) + nl - 1;
            IR(1, ir) = IR(1, nnode) + nl; IR(2, ir) = IR(2, nnode);
            nr = npnode - nl;
        }
        for (int k = 0; k < 3; k++) { xminl[k] = xmaxl[k] = TC(k + 1, IR(1, il)); }
        for (int ip = IR(1, il) + 1; ip <= IR(2, il); ip++)
            for (int k = 0; k < 3; k++) { xminl[k] = std::min(xminl[k], TC(k + 1, ip)); xmaxl[k] = std::max(xmaxl[k], TC(k + 1, ip)); }
        for (int k = 0; k < 3; k++) { xminr[k] = xmaxr[k] = TC(k + 1, IR(1, ir)); }
        for (int ip = IR(1, ir) + 1; ip <= IR(2, ir); ip++)
            for (int k = 0; k < 3; k++) { xminr[k] = std::min(xminr[k], TC(k + 1, ip)); xmaxr[k] = std::max(xmaxr[k], TC(k + 1, ip)); }
    }
    return 0;
}

// maketree (kdtree.F90:117-314) + construct_root_node (kdtree.F90:347-488)
int maketree(oracle_ctx &c, int64_t np, double *xyzh, const int8_t *iphase)
{
    c.npart = np;
    c.ncellsmax = 2 * np;                                // config.F90:393 with maxp = npart
    if (c.ncellsmax < 16) c.ncellsmax = 16;
    c.node.assign(c.ncellsmax + 2, Node());
    c.leaf_is_active.assign(c.ncellsmax + 2, 0);
    c.inoderange.assign(2 * (c.ncellsmax + 2), 0);
    c.inodeparts.assign(np, 0);
    c.treecache.assign(5 * (size_t)np, 0.);
    if (c.p.gravity) c.fnodecache.assign((size_t)lenfgrav * (c.ncellsmax + 2), 0.);

    // --- construct_root_node
    double xminp[3] = {xyzh[0], xyzh[1], xyzh[2]}, xmaxp[3] = {xyzh[0], xyzh[1], xyzh[2]};
    for (int64_t i = 0; i < np; i++) {
        double *x = xyzh + 4 * i;
        if (!(x[3] < DBL_MIN)) {                         // .not.isdead_or_accreted (part.F90:931)
            if (c.p.periodic) {                          // cross_boundary (boundary.f90:123-157)
                if (x[0] < c.p.xmin) x[0] += c.dxbound; else if (x[0] > c.p.xmax) x[0] -= c.dxbound;
                if (x[1] < c.p.ymin) x[1] += c.dybound; else if (x[1] > c.p.ymax) x[1] -= c.dybound;
                if (x[2] < c.p.zmin) x[2] += c.dzbound; else if (x[2] > c.p.zmax) x[2] -= c.dzbound;
            }
            if (std::isnan(x[0]) || std::isnan(x[1]) || std::isnan(x[2])) { c.err = "maketree: NaN in particle position"; return 1; }
            for (int k = 0; k < 3; k++) { xminp[k] = std::min(xminp[k], x[k]); xmaxp[k] = std::max(xmaxp[k], x[k]); }
        }
    }
    int nproot = 0;
    for (int64_t i = 0; i < np; i++) {                   // kdtree.F90:429-456
        const double *x = xyzh + 4 * i;
        if (!(x[3] < DBL_MIN)) {
            nproot++;
            if (c.p.ind_timesteps) c.inodeparts[nproot - 1] = (iphase[i] > 0) ? (int)(i + 1) : -(int)(i + 1);
            else c.inodeparts[nproot - 1] = (int)(i + 1);
            for (int k = 1; k <= 4; k++) TC(k, nproot) = x[k - 1];
            TC(5, nproot) = c.p.massoftype[iamtype(iphase[i])];
        }
    }
    if (nproot != 0) { IR(1, irootnode) = 1; IR(2, irootnode) = nproot; }
    else { c.err = "maketree: no particles or all particles dead/accreted"; return 1; }

    c.ncells = 1; c.maxlevel = 0; c.minlevel = maxdepth - 1;
    c.maxlevel_indexed = (int)(std::log((double)(c.ncellsmax + 1)) / std::log(2.)) - 1;   // kdtree.F90:174
    c.ncells = ((int64_t)1 << (c.maxlevel_indexed + 1)) - 1;

    std::vector<BuildStack> stack(512);
    int istack = 1;
    stack[0].node = irootnode; stack[0].parent = 0; stack[0].level = 0; stack[0].npnode = nproot;
    for (int k = 0; k < 3; k++) { stack[0].xmin[k] = xminp[k]; stack[0].xmax[k] = xmaxp[k]; }
    while (istack > 0) {                                  // over_stack, kdtree.F90:264-290
        BuildStack s = stack[istack - 1]; istack--;
        int il, ir, nl, nr; bool wassplit;
        double xminl[3], xmaxl[3], xminr[3], xmaxr[3];
        if (construct_node(c, s.node, s.parent, s.level, s.xmin, s.xmax, s.npnode, il, ir, nl, nr, xminl, xmaxl, xminr, xmaxr, wassplit)) return 1;
        if (wassplit) {
            if (istack + 2 > 512) { c.err = "maketree: stack size exceeded"; return 1; }
            BuildStack &L = stack[istack++]; L.node = il; L.parent = s.node; L.level = s.level + 1; L.npnode = nl;
            for (int k = 0; k < 3; k++) { L.xmin[k] = xminl[k]; L.xmax[k] = xmaxl[k]; }
            BuildStack &R = stack[istack++]; R.node = ir; R.parent = s.node; R.level = s.level + 1; R.npnode = nr;
            for (int k = 0; k < 3; k++) { R.xmin[k] = xminr[k]; R.xmax[k] = xmaxr[k]; }
        }
    }
    if (c.maxlevel < c.maxlevel_indexed) c.ncells = ((int64_t)1 << (c.maxlevel + 1)) - 1;   // kdtree.F90:298-300
    return 0;
}

// set_hmaxcell (neigh_kdtree.f90:115-131)
inline void set_hmaxcell(oracle_ctx &c, int inode, double hmaxcell)
{
    int n = inode;
#pragma omp critical(crit_node_hmax)
    {
        c.node[n].hmax = hmaxcell;
        while (c.node[n].parent != 0) { n = c.node[n].parent; c.node[n].hmax = std::max(c.node[n].hmax, hmaxcell); }
    }
}

// cache_neighbours (kdtree.F90:1175-1213); xyzcache is (maxcache, ixyzcachesize) column major
inline void cache_neighbours(const oracle_ctx &c, int &nneigh, int isrc, int ixyzcachesize, int maxcache, int *listneigh,
                             double *xyzcache, double xoffset, double yoffset, double zoffset)
{
    const int i1 = c.inoderange[2 * (size_t)isrc], i2 = c.inoderange[2 * (size_t)isrc + 1];
    const int npnode = i2 - i1 + 1;
    int num_to_cache;
    if (nneigh + npnode <= ixyzcachesize) num_to_cache = npnode;
    else if (nneigh < ixyzcachesize) num_to_cache = ixyzcachesize - nneigh;
    else num_to_cache = 0;
    for (int ipart = 1; ipart <= num_to_cache; ipart++) {
        const size_t src = (size_t)(i1 + ipart - 2);  // 0-based slot of treecache(:,i1+ipart-1)
        listneigh[nneigh + ipart - 1] = std::abs(c.inodeparts[src]);
        double *xc = xyzcache + (size_t)maxcache * (nneigh + ipart - 1);
        xc[0] = c.treecache[5 * src + 0] + xoffset;
        xc[1] = c.treecache[5 * src + 1] + yoffset;
        xc[2] = c.treecache[5 * src + 2] + zoffset;
        if (maxcache >= 4) xc[3] = 1. / c.treecache[5 * src + 3];
    }
    for (int ipart = num_to_cache + 1; ipart <= npnode; ipart++)
        listneigh[nneigh + ipart - 1] = std::abs(c.inodeparts[(size_t)(i1 + ipart - 2)]);
    nneigh += npnode;
}

// compute_M2L (kdtree.F90:1702-1781)
inline void compute_M2L(double dx, double dy, double dz, double dr1, double q0, const double *quads, double *fnode)
{
    double dr12 = dr1 * dr1, dx2 = dx * dx, dx3 = dx * dx2, dy2 = dy * dy, dy3 = dy * dy2, dz2 = dz * dz, dz3 = dz * dz2;
    double g0 = -dr1, g1 = 1. * dr12 * g0, g2 = -3. * dr12 * g1, g3 = -5. * dr12 * g2;
    double g2dx = g2 * dx, g2dy = g2 * dy, g2dz = g2 * dz;
    double D3[10], D2[6], D1[3];
    D3[0] = 3. * g2dx + g3 * dx3; D3[1] = g2dy + g3 * dx2 * dy; D3[2] = g2dz + g3 * dx2 * dz; D3[3] = g2dx + g3 * dy2 * dx;
    D3[4] = g3 * dx * dy * dz; D3[5] = g2dx + g3 * dz2 * dx; D3[6] = 3. * g2dy + g3 * dy3; D3[7] = g2dz + g3 * dy2 * dz;
    D3[8] = g2dy + g3 * dz2 * dy; D3[9] = 3. * g2dz + g3 * dz3;
    D2[0] = g1 + g2 * dx2; D2[1] = g2dx * dy; D2[2] = g2dx * dz; D2[3] = g1 + g2 * dy2; D2[4] = g2dy * dz; D2[5] = g1 + g2 * dz2;
    D1[0] = g1 * dx; D1[1] = g1 * dy; D1[2] = g1 * dz;
    double qxx = quads[0], qxy = quads[1], qxz = quads[2], qyy = quads[3], qyz = quads[4], qzz = quads[5];
    fnode[0] = fnode[0] + (D1[0] * q0 + 0.5 * (D3[0] * qxx + 2. * (D3[1] * qxy + D3[2] * qxz + D3[4] * qyz) + D3[3] * qyy + D3[5] * qzz));
    fnode[1] = fnode[1] + (D1[1] * q0 + 0.5 * (D3[1] * qxx + 2. * (D3[3] * qxy + D3[4] * qxz + D3[7] * qyz) + D3[6] * qyy + D3[8] * qzz));
    fnode[2] = fnode[2] + (D1[2] * q0 + 0.5 * (D3[2] * qxx + 2. * (D3[4] * qxy + D3[5] * qxz + D3[8] * qyz) + D3[7] * qyy + D3[9] * qzz));
    for (int k = 0; k < 6; k++) fnode[3 + k] = fnode[3 + k] + D2[k] * q0;
    for (int k = 0; k < 10; k++) fnode[9 + k] = fnode[9 + k] + D3[k] * q0;
    fnode[19] = fnode[19] + g0 * q0 - 0.5 * (D2[0] * qxx + D2[3] * qyy + D2[5] * qzz + 2 * (D2[1] * qxy + D2[2] * qxz + D2[4] * qyz));
}

// propagate_fnode_to_node (kdtree.F90:1527-1560): fnode = L2L(fnode_sup)
inline void propagate_fnode_to_node(double *fnode, const double *fs, double dx, double dy, double dz)
{
    double f[lenfgrav];
    // fs[] is 0-based: fs[k-1] = fnode_sup(k)
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
    for (int k = 9; k < 19; k++) f[k] = fs[k];
    f[19] = fs[19] + dx * (fs[0] + 0.5 * (dx * fs[3] + dy * fs[4] + dz * fs[5])) + dy * (fs[1] + 0.5 * (dx * fs[4] + dy * fs[6] + dz * fs[7])) +
            dz * (fs[2] + 0.5 * (dx * fs[5] + dy * fs[7] + dz * fs[8]));
    for (int k = 0; k < lenfgrav; k++) fnode[k] = f[k];
}

// expand_fgrav_in_taylor_series (kdtree.F90:1799-1840).  NB: the potential line uses dz*dfxy exactly as the reference does (:1836)
inline void expand_fgrav_in_taylor_series(const double *fnode, double dx, double dy, double dz, double &fxi, double &fyi, double &fzi, double &poti)
{
    fxi = fnode[0]; fyi = fnode[1]; fzi = fnode[2];
    double dfxx = fnode[3], dfxy = fnode[4], dfxz = fnode[5], dfyy = fnode[6], dfyz = fnode[7], dfzz = fnode[8];
    double d2fxxx = fnode[9], d2fxxy = fnode[10], d2fxxz = fnode[11], d2fxyy = fnode[12], d2fxyz = fnode[13], d2fxzz = fnode[14];
    double d2fyyy = fnode[15], d2fyyz = fnode[16], d2fyzz = fnode[17], d2fzzz = fnode[18];
    poti = fnode[19];
    fxi = fxi + dx * (dfxx + 0.5 * (dx * d2fxxx + dy * d2fxxy + dz * d2fxxz)) + dy * (dfxy + 0.5 * (dx * d2fxxy + dy * d2fxyy + dz * d2fxyz)) +
          dz * (dfxz + 0.5 * (dx * d2fxxz + dy * d2fxyz + dz * d2fxzz));
    fyi = fyi + dx * (dfxy + 0.5 * (dx * d2fxxy + dy * d2fxyy + dz * d2fxyz)) + dy * (dfyy + 0.5 * (dx * d2fxyy + dy * d2fyyy + dz * d2fyyz)) +
          dz * (dfyz + 0.5 * (dx * d2fxyz + dy * d2fyyz + dz * d2fyzz));
    fzi = fzi + dx * (dfxz + 0.5 * (dx * d2fxxz + dy * d2fxyz + dz * d2fxzz)) + dy * (dfyz + 0.5 * (dx * d2fxyz + dy * d2fyyz + dz * d2fyzz)) +
          dz * (dfzz + 0.5 * (dx * d2fxzz + dy * d2fyzz + dz * d2fzzz));
    poti = poti - (dx * (fxi - 0.5 * (dx * dfxx + dy * dfxy + dz * dfxy)) + dy * (fyi - 0.5 * (dx * dfxy + dy * dfyy + dz * dfyz)) +
                   dz * (fzi - 0.5 * (dx * dfxz + dy * dfyz + dz * dfzz)));
}

// getneigh (kdtree.F90:1221-1347), local tree only
void getneigh(const oracle_ctx &c, const double *xpos, double xsizei, double rcuti, int *listneigh, int &nneigh, double *xyzcache,
              int ixyzcachesize, int maxcache, bool get_hj, bool get_f, double *fnode)
{
    const double tree_acc2 = c.p.tree_accuracy * c.p.tree_accuracy;
    if (fnode) for (int k = 0; k < lenfgrav; k++) fnode[k] = 0.;
    double rcut = rcuti;
    nneigh = 0;
    int nstack[maxdepth * 2 + 8];
    int istack = 1; nstack[0] = irootnode;
    bool open_tree_node = false;
    while (istack != 0) {
        int n = nstack[istack - 1]; istack--;
        const Node &nd = c.node[n];
        double dx, dy, dz, xoffset, yoffset, zoffset, r2;
        get_sep(c, xpos, nd.xcen, dx, dy, dz, xoffset, yoffset, zoffset, r2);
        double xsizej = nd.size;
        int il = nd.leftchild, ir = nd.rightchild;
        if (get_hj) { double rcutj = c.kc.radkern * nd.hmax; rcut = std::max(rcuti, rcutj); }
        double rcut2 = pow2(xsizei + xsizej + rcut);
        if (c.p.gravity) open_tree_node = tree_acc2 * r2 < pow2(xsizei + xsizej);
        if ((r2 < rcut2) || open_tree_node) {
            if (c.leaf_is_active[n] != 0) {
                cache_neighbours(c, nneigh, n, ixyzcachesize, maxcache, listneigh, xyzcache, xoffset, yoffset, zoffset);
            } else {
                if (il != 0) nstack[istack++] = il;
                if (ir != 0) nstack[istack++] = ir;
            }
        } else if (c.p.gravity && get_f) {
            double dr = 1. / std::sqrt(r2);
            compute_M2L(dx, dy, dz, dr, nd.mass, nd.quads, fnode);
        }
    }
}

// getneigh_dual (kdtree.F90:1357-1461) + open_nodes (:1591) + node_interaction (:1653).
// The fnodecache optimisation (:1427-1449) stores the branch value computed by the first
// thread to reach an ancestor; every leaf under that ancestor computes the identical
// interaction list for it, so caching does not change values and is omitted here.
void getneigh_dual(const oracle_ctx &c, const double *xpos, double xsizei, double rcuti, int *listneigh, int &nneigh, double *xyzcache,
                   int ixyzcachesize, int maxcache, double *fnode, int icell)
{
    (void)xpos; (void)xsizei; (void)rcuti;
    const double tree_acc2 = c.p.tree_accuracy * c.p.tree_accuracy;
    int branch[maxdepth + 2], nparents = 1;
    { int j = icell; branch[0] = j; while (c.node[j].parent != 0) { j = c.node[j].parent; branch[nparents++] = j; } }
    std::vector<double> fnode_branch((size_t)lenfgrav * (maxdepth + 2), 0.);
    double fnode_acc[lenfgrav]; for (int k = 0; k < lenfgrav; k++) fnode_acc[k] = 0.;
    nneigh = 0;
    // each level keeps at most 2 pending entries per depth of src descent; 4*maxdepth is ample
    std::vector<int> stack(3 * 64 * maxdepth);
    int istack = 1;
    stack[0] = irootnode; stack[1] = irootnode; stack[2] = nparents;
    while (istack > 0) {
        int idst = stack[3 * (istack - 1)], isrc = stack[3 * (istack - 1) + 1], idstbranch = stack[3 * (istack - 1) + 2];
        istack--;
        bool stackit; double xoffset = 0., yoffset = 0., zoffset = 0.;
        if (idst == isrc) stackit = true;
        else {
            const Node &nd = c.node[idst], &ns = c.node[isrc];
            double dx, dy, dz, r2;
            get_sep(c, nd.xcen, ns.xcen, dx, dy, dz, xoffset, yoffset, zoffset, r2);
            double rcut_src = ns.hmax * c.kc.radkern, rcut_dst = nd.hmax * c.kc.radkern;
            double rcut = std::max(rcut_dst, rcut_src);
            double rcut2 = pow2(nd.size + ns.size + rcut);
            bool wellsep = (tree_acc2 * r2 > pow2(nd.size + ns.size)) && (r2 > rcut2);
            if (wellsep) {
                double dr1 = 1. / std::sqrt(r2);
                compute_M2L(dx, dy, dz, dr1, ns.mass, ns.quads, &fnode_branch[(size_t)lenfgrav * (idstbranch - 1)]);
                stackit = false;
            } else stackit = true;
        }
        if (stackit) {   // open_nodes
            const Node &ns = c.node[isrc];
            int ibranchnext; bool isdstleaf;
            if (idstbranch - 1 > 0) { ibranchnext = idstbranch - 1; isdstleaf = false; }
            else { ibranchnext = idstbranch; isdstleaf = true; }
            int idstnext = branch[ibranchnext - 1];
            if (c.leaf_is_active[isrc] != 0) {
                if (isdstleaf) cache_neighbours(c, nneigh, isrc, ixyzcachesize, maxcache, listneigh, xyzcache, xoffset, yoffset, zoffset);
                else { int *s = &stack[3 * istack++]; s[0] = idstnext; s[1] = isrc; s[2] = ibranchnext; }
            } else {
                if (ns.leftchild != 0) { int *s = &stack[3 * istack++]; s[0] = idstnext; s[1] = ns.leftchild; s[2] = ibranchnext; }
                if (ns.rightchild != 0) { int *s = &stack[3 * istack++]; s[0] = idstnext; s[1] = ns.rightchild; s[2] = ibranchnext; }
            }
            if ((size_t)(3 * (istack + 2)) > stack.size()) stack.resize(stack.size() * 2);
        }
    }
    for (int i = nparents; i >= 2; i--) {        // downward pass (kdtree.F90:1427-1457)
        int iparent = branch[i - 1];
        double dx, dy, dz, xo, yo, zo, r2;
        get_sep(c, c.node[branch[i - 2]].xcen, c.node[iparent].xcen, dx, dy, dz, xo, yo, zo, r2);
        double ftmp[lenfgrav];
        for (int k = 0; k < lenfgrav; k++) ftmp[k] = fnode_acc[k] + fnode_branch[(size_t)lenfgrav * (i - 1) + k];
        propagate_fnode_to_node(fnode_acc, ftmp, dx, dy, dz);
    }
    for (int k = 0; k < lenfgrav; k++) fnode[k] = fnode_acc[k] + fnode_branch[k];
}

// get_neighbour_list (neigh_kdtree.f90:218-292), non-MPI
inline void get_neighbour_list(const oracle_ctx &c, int inode, int *listneigh, int &nneigh, double *xyzcache, int ixyzcachesize, int maxcache,
                               bool getj, double *f, const double *cell_xpos = nullptr, double cell_xsizei = 0., double cell_rcuti = 0.)
{
    double xpos[3], xsizei, rcuti;
    if (cell_xpos) { xpos[0] = cell_xpos[0]; xpos[1] = cell_xpos[1]; xpos[2] = cell_xpos[2]; xsizei = cell_xsizei; rcuti = cell_rcuti; }
    else {  // get_cell_location (neigh_kdtree.f90:354-365)
        const Node &nd = c.node[inode];
        xpos[0] = nd.xcen[0]; xpos[1] = nd.xcen[1]; xpos[2] = nd.xcen[2]; xsizei = nd.size; rcuti = c.kc.radkern * nd.hmax;
    }
    bool get_f = (c.p.gravity && f != nullptr);
    if (get_f) getneigh_dual(c, xpos, xsizei, rcuti, listneigh, nneigh, xyzcache, ixyzcachesize, maxcache, f, inode);
    else getneigh(c, xpos, xsizei, rcuti, listneigh, nneigh, xyzcache, ixyzcachesize, maxcache, getj, false, nullptr);
}

// ------------------------------------------------------------------------------
//  density pass
// ------------------------------------------------------------------------------
struct CellDens {  // mpi_dens.F90:63-86
    int icell, npcell, nits, nneightry;
    int arr_index[minpart]; int8_t iphase[minpart]; int nneigh[minpart];
    double xpos[3], xsizei, rcuti, hmax;
    double h[minpart], h_old[minpart];
    // xpartvec (dens.F90:33-48)
    double x[minpart], y[minpart], z[minpart], vx[minpart], vy[minpart], vz[minpart], en[minpart];
    double Bx[minpart], By[minpart], Bz[minpart], psi[minpart], fx[minpart], fy[minpart], fz[minpart];
    double rhosums[minpart][maxrhosum];
};

struct DensArrays {
    double *xyzh; const double *vxyzu, *fxyzu, *fext, *Bevol; const int8_t *iphase;
    float *divcurlv, *divcurlB, *alphaind, *gradh, *dvdx; double *dustfrac;
};

// get_density_sums (dens.F90:578-857) with isizeneighcache = 0
inline void get_density_sums(const oracle_ctx &c, int i, const CellDens &cell, int ip, double hi, double hi1, double hi21, int iamtypei,
                             bool iamgasi, const int *listneigh, int nneigh, int &nneighi, const double *xyzcache, double *rhosum,
                             bool getdv, bool getdB, const DensArrays &a, int64_t &npairs)
{
    const oracle_params &p = c.p;
    const int nvu = c.nvu();
    for (int k = 0; k < maxrhosum; k++) rhosum[k] = 0.;
    nneighi = 1;  // self (ignoreself = .true. for local cells)
    bool same_type = true, gas_gas = true;
    double dphidhi = 0.;
    const double xi = cell.x[ip], yi = cell.y[ip], zi = cell.z[ip];
    const double fxi = cell.fx[ip], fyi = cell.fy[ip], fzi = cell.fz[ip];
    const bool nalpha_gt1 = c.nalpha() > 1;
    for (int n = 1; n <= nneigh; n++) {
        const int j = listneigh[n - 1];
        if (j == i) continue;
        double dx, dy, dz;
        if (n <= isizecellcache) {
            const double *xc = xyzcache + 3 * (size_t)(n - 1);
            dx = xi - xc[0]; dy = yi - xc[1]; dz = zi - xc[2];
        } else {
            const double *xj = a.xyzh + 4 * (size_t)(j - 1);
            dx = xi - xj[0]; dy = yi - xj[1]; dz = zi - xj[2];
        }
        if (p.periodic) {   // dens.F90:666-670
            if (std::fabs(dx) > 0.5 * c.dxbound) dx = dx - c.dxbound * std::copysign(1.0, dx);
            if (std::fabs(dy) > 0.5 * c.dybound) dy = dy - c.dybound * std::copysign(1.0, dy);
            if (std::fabs(dz) > 0.5 * c.dzbound) dz = dz - c.dzbound * std::copysign(1.0, dz);
        }
        const double rij2 = dx * dx + dy * dy + dz * dz;
        const double q2i = rij2 * hi21;
        if (q2i < c.kc.radkern2) {
            const double rij = std::sqrt(rij2);
            const double qi = rij * hi1;
            double wabi, grkerni;
            get_kernel(p.kernel, q2i, qi, wabi, grkerni);
            if (p.gravity) dphidhi = dphidh_kernel(p.kernel, q2i, qi);
            const int8_t iphasej = a.iphase[j - 1];
            const int iamtypej = iamtype(iphasej);
            const bool iamdustj = p.dust && (iamtypej == idust);
            same_type = ((iamtypei == iamtypej) || (ibasetype(iamtypej) == iamtypei));
            gas_gas = (iamgasi && same_type);
            const double pmassi = p.massoftype[iamtypei], pmassj = p.massoftype[iamtypej];
            if (same_type) {
                npairs++;
                const double dwdhi = (-qi * grkerni - 3. * wabi);
                rhosum[irhoi] = rhosum[irhoi] + wabi * pmassj;
                rhosum[igradhi] = rhosum[igradhi] + dwdhi * pmassj;
                rhosum[igradsofti] = rhosum[igradsofti] + dphidhi * pmassj;
                nneighi = nneighi + 1;
                if (getdv || getdB) {
                    const double rij1 = 1. / (rij + DBL_EPSILON);
                    const double rij1grkern = rij1 * grkerni;
                    const double runix = dx * rij1grkern * pmassj, runiy = dy * rij1grkern * pmassj, runiz = dz * rij1grkern * pmassj;
                    if (getdv) {
                        const double *vj = a.vxyzu + (size_t)nvu * (j - 1);
                        const double dvx = cell.vx[ip] - vj[0], dvy = cell.vy[ip] - vj[1], dvz = cell.vz[ip] - vj[2];
                        const double projv = dvx * runix + dvy * runiy + dvz * runiz;
                        rhosum[idivvi] = rhosum[idivvi] + projv;
                        rhosum[idvxdxi] += dvx * runix; rhosum[idvxdyi] += dvx * runiy; rhosum[idvxdzi] += dvx * runiz;
                        rhosum[idvydxi] += dvy * runix; rhosum[idvydyi] += dvy * runiy; rhosum[idvydzi] += dvy * runiz;
                        rhosum[idvzdxi] += dvz * runix; rhosum[idvzdyi] += dvz * runiy; rhosum[idvzdzi] += dvz * runiz;
                        if (nalpha_gt1 && gas_gas) {
                            const double *fj = a.fxyzu + (size_t)nvu * (j - 1), *fe = a.fext + 3 * (size_t)(j - 1);
                            const double fxj = fj[0] + fe[0], fyj = fj[1] + fe[1], fzj = fj[2] + fe[2];
                            const double dax = fxi - fxj, day = fyi - fyj, daz = fzi - fzj;
                            rhosum[idaxdxi] += dax * runix; rhosum[idaxdyi] += dax * runiy; rhosum[idaxdzi] += dax * runiz;
                            rhosum[idaydxi] += day * runix; rhosum[idaydyi] += day * runiy; rhosum[idaydzi] += day * runiz;
                            rhosum[idazdxi] += daz * runix; rhosum[idazdyi] += daz * runiy; rhosum[idazdzi] += daz * runiz;
                        }
                        rhosum[irxxi] -= dx * runix; rhosum[irxyi] -= dx * runiy; rhosum[irxzi] -= dx * runiz;
                        rhosum[iryyi] -= dy * runiy; rhosum[iryzi] -= dy * runiz; rhosum[irzzi] -= dz * runiz;
                    }
                    if (getdB && gas_gas) {   // dens.F90:806-829
                        const double rhoi = c.rhoh(hi, pmassi);
                        const double rhoj = c.rhoh(a.xyzh[4 * (size_t)(j - 1) + 3], pmassj);
                        const double *Bj = a.Bevol + 4 * (size_t)(j - 1);
                        const double dBx = cell.Bx[ip] * rhoi - Bj[0] * rhoj, dBy = cell.By[ip] * rhoi - Bj[1] * rhoj, dBz = cell.Bz[ip] * rhoi - Bj[2] * rhoj;
                        const double projdB = dBx * runix + dBy * runiy + dBz * runiz;
                        rhosum[idivBi] += projdB;
                        rhosum[idBxdxi] += dBx * runix; rhosum[idBxdyi] += dBx * runiy; rhosum[idBxdzi] += dBx * runiz;
                        rhosum[idBydxi] += dBy * runix; rhosum[idBydyi] += dBy * runiy; rhosum[idBydzi] += dBy * runiz;
                        rhosum[idBzdxi] += dBz * runix; rhosum[idBzdyi] += dBz * runiy; rhosum[idBzdzi] += dBz * runiz;
                    }
                }
            } else if (p.dust && (iamgasi && iamdustj)) {
                rhosum[irhodusti + iamtypej - idust] += wabi;
            }
        }
    }
}

// finish_rhosum (dens.F90:1470-1507)
inline void finish_rhosum(const oracle_ctx &c, const double *rhosum, double pmassi, double hi, bool iterating, double &rhoi, double &gradhi,
                          double &rhohi, double &gradsofti, double &dhdrhoi_out, double &omegai_out)
{
    const double hi1 = 1. / hi, hi21 = hi1 * hi1, hi31 = hi1 * hi21, hi41 = hi21 * hi21;
    rhoi = c.kc.cnormk * (rhosum[irhoi] + c.kc.wab0 * pmassi) * hi31;
    gradhi = c.kc.cnormk * (rhosum[igradhi] + c.kc.gradh0 * pmassi) * hi41;
    const double dhdrhoi = c.dhdrho(hi, pmassi);
    const double omegai = 1. - dhdrhoi * gradhi;
    gradhi = 1. / omegai;
    if (iterating) { rhohi = c.rhoh(hi, pmassi); dhdrhoi_out = dhdrhoi; omegai_out = omegai; }
    else { gradsofti = (rhosum[igradsofti] + c.kc.dphidh0 * pmassi) * hi21; gradsofti = gradsofti * dhdrhoi; }
}

inline void exactlinear(double &gAx, double &gAy, double &gAz, double dAx, double dAy, double dAz, const double *rm, double ddenom)
{   // dens.F90:1086-1101
    gAx = (dAx * rm[0] + dAy * rm[1] + dAz * rm[2]) * ddenom;
    gAy = (dAx * rm[1] + dAy * rm[3] + dAz * rm[4]) * ddenom;
    gAz = (dAx * rm[2] + dAy * rm[4] + dAz * rm[5]) * ddenom;
}

// start_cell (dens.F90:1293-1378)
void dens_start_cell(const oracle_ctx &c, CellDens &cell, const DensArrays &a)
{
    const int nvu = c.nvu();
    cell.npcell = 0;
    const int i1 = c.inoderange[2 * (size_t)cell.icell], i2 = c.inoderange[2 * (size_t)cell.icell + 1];
    for (int ip = i1; ip <= i2; ip++) {
        const int i = c.inodeparts[ip - 1];
        if (i < 0) continue;
        bool iactivei, iamgasi, iamdusti; int iamtypei;
        get_partinfo(c.p, a.iphase[i - 1], iactivei, iamgasi, iamdusti, iamtypei);
        if (!iactivei) continue;
        const int n = cell.npcell++;
        cell.arr_index[n] = ip; cell.iphase[n] = a.iphase[i - 1];
        const double *x = a.xyzh + 4 * (size_t)(i - 1), *v = a.vxyzu + (size_t)nvu * (i - 1);
        cell.x[n] = x[0]; cell.y[n] = x[1]; cell.z[n] = x[2];
        cell.h[n] = x[3]; cell.h_old[n] = x[3];
        cell.vx[n] = v[0]; cell.vy[n] = v[1]; cell.vz[n] = v[2];
        cell.en[n] = nvu >= 4 ? v[3] : 0.;
        const double *f = a.fxyzu + (size_t)nvu * (i - 1), *fe = a.fext + 3 * (size_t)(i - 1);
        cell.fx[n] = f[0] + fe[0]; cell.fy[n] = f[1] + fe[1]; cell.fz[n] = f[2] + fe[2];
        if (c.p.mhd) {
            if (iamgasi) { const double *B = a.Bevol + 4 * (size_t)(i - 1); cell.Bx[n] = B[0]; cell.By[n] = B[1]; cell.Bz[n] = B[2]; cell.psi[n] = B[3]; }
            else { cell.Bx[n] = cell.By[n] = cell.Bz[n] = cell.psi[n] = 0.; }
        }
    }
}

// compute_cell (dens.F90:1203-1271)
void dens_compute_cell(const oracle_ctx &c, CellDens &cell, const int *listneigh, int nneigh, bool getdv, bool getdB, const DensArrays &a,
                       const double *xyzcache, int64_t &npairs)
{
    for (int i = 0; i < cell.npcell; i++) {
        const int lli = c.inodeparts[cell.arr_index[i] - 1];
        bool iactivei, iamgasi, iamdusti; int iamtypei;
        get_partinfo(c.p, cell.iphase[i], iactivei, iamgasi, iamdusti, iamtypei);
        const double hi = cell.h[i], hi1 = 1. / hi, hi21 = hi1 * hi1;
        int nneighi;
        get_density_sums(c, lli, cell, i, hi, hi1, hi21, iamtypei, iamgasi, listneigh, nneigh, nneighi, xyzcache, cell.rhosums[i], getdv, getdB, a, npairs);
        cell.nneightry = nneigh;
        cell.nneigh[i] = nneighi;
    }
}

// finish_cell (dens.F90:1382-1466); returns 1 on non-convergence (fatal)
int dens_finish_cell(oracle_ctx &c, CellDens &cell, bool &cell_converged)
{
    cell.nits = cell.nits + 1;
    cell_converged = true;
    for (int i = 0; i < cell.npcell; i++) {
        const double hi = cell.h[i], hi_old = cell.h_old[i];
        bool iactivei, iamgasi, iamdusti; int iamtypei;
        get_partinfo(c.p, cell.iphase[i], iactivei, iamgasi, iamdusti, iamtypei);
        const double pmassi = c.p.massoftype[iamtypei];
        double rhoi, gradhi, rhohi, gradsofti, dhdrhoi, omegai;
        finish_rhosum(c, cell.rhosums[i], pmassi, hi, true, rhoi, gradhi, rhohi, gradsofti, dhdrhoi, omegai);
        const double func = rhohi - rhoi;
        double dfdh1;
        if (omegai > DBL_MIN) dfdh1 = dhdrhoi / omegai;
        else dfdh1 = dhdrhoi / std::fabs(omegai + DBL_EPSILON);
        double hnew = hi - func * dfdh1;
        if (hnew > 1.2 * hi) hnew = 1.2 * hi;
        else if (hnew < 0.8 * hi) hnew = 0.8 * hi;
        const bool converged = ((std::fabs(hnew - hi) / hi_old) < c.p.tolh && omegai > 0. && hi > 0.);
        if (cell_converged) cell_converged = converged;
        if (!converged && cell.nits >= maxdensits) {
#pragma omp critical(oracle_err)
            {
                char buf[256];
                snprintf(buf, sizeof buf, "densityiterate: could not converge in density on particle %d (error %g)",
                         c.inodeparts[cell.arr_index[i] - 1], std::fabs(hnew - hi) / hi_old);
                c.err = buf;
            }
            return 1;
        }
        cell.h[i] = converged ? hi : hnew;
    }
    return 0;
}

// compute_hmax (dens.F90:1275-1289)
inline void dens_compute_hmax(const oracle_ctx &c, CellDens &cell, bool &redo_neighbours)
{
    redo_neighbours = false;
    if (cell.npcell > 0) {
        const double hmax_old = cell.hmax;
        double hm = cell.h[0];
        for (int i = 1; i < cell.npcell; i++) hm = std::max(hm, cell.h[i]);
        const double hmax = 1.01 * hm;
        if (hmax > hmax_old) redo_neighbours = true;
        cell.hmax = hmax;
        cell.rcuti = c.kc.radkern * hmax;
    }
}

// store_results (dens.F90:1511-1681), fast_divcurlB => calculate_density and calculate_divcurlB both true
void dens_store_results(oracle_ctx &c, const CellDens &cell, bool getdv, bool getdB, const DensArrays &a, double &rhomax,
                        int64_t &nneightry, int64_t &nneighact, int64_t &maxneightry, int &maxneighact, int64_t &np, int64_t &ncalc)
{
    const int ngradh = c.ngradh(), nalpha = c.nalpha();
    for (int i = 0; i < cell.npcell; i++) {
        const int lli = c.inodeparts[cell.arr_index[i] - 1];
        const double hi = cell.h[i];
        const double *rhosum = cell.rhosums[i];
        const double hi1 = 1. / hi, hi21 = hi1 * hi1, hi31 = hi1 * hi21, hi41 = hi21 * hi21;
        bool iactivei, iamgasi, iamdusti; int iamtypei;
        get_partinfo(c.p, cell.iphase[i], iactivei, iamgasi, iamdusti, iamtypei);
        const double pmassi = c.p.massoftype[iamtypei];
        double rhoi, gradhi, rhohi, gradsofti = 0., d1, d2;
        finish_rhosum(c, rhosum, pmassi, hi, false, rhoi, gradhi, rhohi, gradsofti, d1, d2);
        a.xyzh[4 * (size_t)(lli - 1) + 3] = c.hrho(rhoi, pmassi);                          // dens.F90:1595
        c.treecache[5 * (size_t)(cell.arr_index[i] - 1) + 3] = a.xyzh[4 * (size_t)(lli - 1) + 3];
        a.gradh[(size_t)ngradh * (lli - 1)] = (float)gradhi;                                 // real(gradhi,kind=4)
        if (c.p.gravity) a.gradh[(size_t)ngradh * (lli - 1) + 1] = (float)gradsofti;
        rhomax = std::max(rhomax, rhoi);
        // calculate_divcurlB branch
        gradhi = (double)a.gradh[(size_t)ngradh * (lli - 1)];                                // dens.F90:1610 (re-read real*4)
        const double rho1i = 1. / rhoi;
        if (c.p.dust && a.dustfrac) {                                                        // two-fluid: dust-to-gas ratio on gas particles
            a.dustfrac[lli - 1] = 0.;
            if (iamgasi) { const double rhodusti = c.kc.cnormk * c.p.massoftype[idust] * rhosum[irhodusti] * hi31; a.dustfrac[lli - 1] = rhodusti * rho1i; }
        }
        const double term = c.kc.cnormk * gradhi * rho1i * hi41;
        double denom = 0., rmatrix[6] = {0, 0, 0, 0, 0, 0};
        if (getdv) {
            // calculate_rmatrix_from_sums (dens.F90:866-891)
            const double rxxi = rhosum[irxxi], rxyi = rhosum[irxyi], rxzi = rhosum[irxzi], ryyi = rhosum[iryyi], ryzi = rhosum[iryzi], rzzi = rhosum[irzzi];
            denom = rxxi * ryyi * rzzi + 2. * rxyi * rxzi * ryzi - rxxi * ryzi * ryzi - ryyi * rxzi * rxzi - rzzi * rxyi * rxyi;
            rmatrix[0] = ryyi * rzzi - ryzi * ryzi; rmatrix[1] = rxzi * ryzi - rzzi * rxyi; rmatrix[2] = rxyi * ryzi - rxzi * ryyi;
            rmatrix[3] = rzzi * rxxi - rxzi * rxzi; rmatrix[4] = rxyi * rxzi - rxxi * ryzi; rmatrix[5] = rxxi * ryyi - rxyi * rxyi;
            // calculate_divcurlv_from_sums (dens.F90:899-967), ndivcurlv = 1
            const double divv = -rhosum[idivvi] * term;
            double divcurlv5 = 0.;
            double dv[9];
            const bool exact = std::fabs(denom) > DBL_MIN;
            if (exact) {
                const double ddenom = 1. / denom;
                double g[3];
                for (int r = 0; r < 3; r++) {
                    exactlinear(g[0], g[1], g[2], rhosum[idvxdxi + 3 * r], rhosum[idvxdxi + 3 * r + 1], rhosum[idvxdxi + 3 * r + 2], rmatrix, ddenom);
                    dv[3 * r] = -g[0]; dv[3 * r + 1] = -g[1]; dv[3 * r + 2] = -g[2];
                }
                if (nalpha >= 2) {
                    double gax[3], gay[3], gaz[3];
                    exactlinear(gax[0], gax[1], gax[2], rhosum[idaxdxi], rhosum[idaxdyi], rhosum[idaxdzi], rmatrix, ddenom);
                    exactlinear(gay[0], gay[1], gay[2], rhosum[idaydxi], rhosum[idaydyi], rhosum[idaydzi], rmatrix, ddenom);
                    exactlinear(gaz[0], gaz[1], gaz[2], rhosum[idazdxi], rhosum[idazdyi], rhosum[idazdzi], rmatrix, ddenom);
                    const double div_a = -(gax[0] + gay[1] + gaz[2]);
                    divcurlv5 = div_a - (dv[0] * dv[0] + dv[4] * dv[4] + dv[8] * dv[8] + 2. * (dv[1] * dv[3] + dv[2] * dv[6] + dv[5] * dv[7]));
                }
            } else {
                for (int k = 0; k < 9; k++) dv[k] = -term * rhosum[idvxdxi + k];
                if (nalpha >= 2) {
                    const double div_a = -term * (rhosum[idaxdxi] + rhosum[idaydyi] + rhosum[idazdzi]);
                    divcurlv5 = div_a - (dv[0] * dv[0] + dv[4] * dv[4] + dv[8] * dv[8] + 2. * (dv[1] * dv[3] + dv[2] * dv[6] + dv[5] * dv[7]));
                }
            }
            a.divcurlv[lli - 1] = (float)divv;
            if (nalpha >= 3) a.alphaind[3 * (size_t)(lli - 1) + 2] = (float)divcurlv5;
            // calculate_strain_from_sums with use_exact_linear = .not.realviscosity = .true. (dens.F90:997-1052)
            double dvdxi[9];
            if (exact) { for (int k = 0; k < 9; k++) dvdxi[k] = dv[k]; }
            else { for (int k = 0; k < 9; k++) dvdxi[k] = -rhosum[idvxdxi + k] * term; }
            for (int k = 0; k < 9; k++) a.dvdx[9 * (size_t)(lli - 1) + k] = (float)dvdxi[k];
        } else {
            a.divcurlv[lli - 1] = -(float)(rhosum[idivvi] * term);
            if (nalpha >= 2) a.alphaind[3 * (size_t)(lli - 1) + 1] = 0.f;
        }
        if (c.p.mhd && iamgasi && getdB) {     // calculate_divcurlB_from_sums (dens.F90:975-989)
            float *dcB = a.divcurlB + 4 * (size_t)(lli - 1);
            dcB[0] = (float)(-rhosum[idivBi] * term);
            dcB[1] = (float)(-(rhosum[idBzdyi] - rhosum[idBydzi]) * term);
            dcB[2] = (float)(-(rhosum[idBxdzi] - rhosum[idBzdxi]) * term);
            dcB[3] = (float)(-(rhosum[idBydxi] - rhosum[idBxdyi]) * term);
        }
        nneightry += cell.nneightry;
        nneighact += cell.nneigh[i];
        maxneightry = std::max(maxneightry, (int64_t)cell.nneightry);
        maxneighact = std::max(maxneighact, cell.nneigh[i]);
    }
    np += cell.npcell;
    ncalc += (int64_t)cell.npcell * cell.nits;
}

// ------------------------------------------------------------------------------
//  force pass
// ------------------------------------------------------------------------------
struct ForceArrays {
    const double *xyzh, *vxyzu; double *fxyzu; float *divcurlv; const float *divcurlB; const double *Bevol; double *dBevol;
    const double *fext, *eos_vars; const float *alphaind, *gradh, *dvdx; const int8_t *iphase; const double *dustfrac;
    float *poten, *divBsymm; double *tstop; int8_t *ibin, *ibin_wake; const int8_t *ibin_old;
};

struct PartForce {   // the slice of xpartvec (force.F90:62-139) this path uses
    int ip_index; int8_t iphase;
    double x, y, z, h, vx, vy, vz, en, Bevolx, Bevoly, Bevolz, psi, gradh1, gradh2, alpha, vwave, rho, rhogas, spsound, temp;
    double sxx, sxy, sxz, syy, syz, szz, pr, pro2, dvdx[9];
    double fsum[maxfsum], vsigmax, tsmin; int ibinneigh;
};

// get_stress (force.F90:2068-2168) without physical viscosity / radiation
inline void get_stress(const oracle_ctx &c, double pri, double spsoundi, double rhoi, double rho1i, double pmassi, double Bxi, double Byi, double Bzi,
                       double &pro2i, double &vwavei, double &sxxi, double &sxyi, double &sxzi, double &syyi, double &syzi, double &szzi)
{
    sxxi = sxyi = sxzi = syyi = syzi = szzi = 0.;
    const double stressiso = 0.;
    if (c.p.mhd) {
        const double Brhoxi = Bxi * rho1i, Brhoyi = Byi * rho1i, Brhozi = Bzi * rho1i;
        const double Bro2i = Brhoxi * Brhoxi + Brhoyi * Brhoyi + Brhozi * Brhozi;
        const double valfven2i = Bro2i * rhoi;
        vwavei = std::sqrt(spsoundi * spsoundi + valfven2i);
        sxxi = sxxi - pmassi * Brhoxi * Brhoxi; sxyi = sxyi - pmassi * Brhoxi * Brhoyi; sxzi = sxzi - pmassi * Brhoxi * Brhozi;
        syyi = syyi - pmassi * Brhoyi * Brhoyi; syzi = syzi - pmassi * Brhoyi * Brhozi; szzi = szzi - pmassi * Brhozi * Brhozi;
        pro2i = (pri + 0.) * rho1i * rho1i + stressiso + 0.5 * Bro2i;
    } else {
        pro2i = (pri + 0.) * rho1i * rho1i + stressiso;
        vwavei = spsoundi;
    }
}

// get_ts (dust.f90:161-276), idrag = 1 (Epstein/Stokes), 2 (const K), 3 (const ts)
inline void get_ts(const oracle_ctx &c, double sgrain, double densgrain, double rhogas, double rhodust, double spsoundgas, double dv2, double &ts, int &iregime);

// reconstruct_dv (force.F90:3338-3380)
inline double slope_limiter(double sl, double sr, int ilimiter);

// compute_forces (force.F90:914-2060) for hydro + MHD + gravity P2P + two-fluid drag
void compute_forces(const oracle_ctx &c, int i, bool iamgasi, bool iamdusti, PartForce &pf, double hi, double hi1, double hi21, double hi41,
                    double gradhi, double gradsofti, double pmassi, const int *listneigh, int nneigh, const double *xyzcache,
                    const ForceArrays &a, int ibinnow_m1, int64_t &npairs);

}  // namespace

#include "sph_oracle_force.inc"

// ------------------------------------------------------------------------------
//  C entry points
// ------------------------------------------------------------------------------
extern "C" {

oracle_ctx *oracle_create(const oracle_params *p)
{
    oracle_ctx *c = new oracle_ctx();
    c->p = *p; c->kc = kernel_consts(p->kernel); c->set_bounds();
    return c;
}
void oracle_destroy(oracle_ctx *c) { delete c; }
void oracle_set_params(oracle_ctx *c, const oracle_params *p) { c->p = *p; c->kc = kernel_consts(p->kernel); c->set_bounds(); }
void oracle_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int oracle_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
const char *oracle_last_error(oracle_ctx *c) { return c->err.c_str(); }

int oracle_build_tree(oracle_ctx *c, int64_t npart, double *xyzh, const int8_t *iphase)
{
    c->err.clear();
    return maketree(*c, npart, xyzh, iphase);
}
int64_t oracle_tree_ncells(oracle_ctx *c) { return c->ncells; }
int oracle_tree_get_node(oracle_ctx *c, int64_t n, double *rec, int32_t *irec)
{
    if (n < 1 || n > c->ncellsmax) return 1;
    const Node &nd = c->node[n];
    rec[0] = nd.xcen[0]; rec[1] = nd.xcen[1]; rec[2] = nd.xcen[2]; rec[3] = nd.size; rec[4] = nd.hmax; rec[5] = nd.mass;
    for (int k = 0; k < 6; k++) rec[6 + k] = nd.quads[k];
    irec[0] = nd.leftchild; irec[1] = nd.rightchild; irec[2] = nd.parent; irec[3] = c->leaf_is_active[n];
    irec[4] = c->inoderange[2 * n]; irec[5] = c->inoderange[2 * n + 1];
    return 0;
}
int oracle_tree_get_inodeparts(oracle_ctx *c, int32_t *out)
{
    for (int64_t i = 0; i < c->npart; i++) out[i] = c->inodeparts[i];
    return 0;
}
int64_t oracle_get_neighbour_list(oracle_ctx *c, int64_t icell, int getj, int32_t *list, int64_t maxlist)
{
    std::vector<int> listneigh(c->npart + 16);
    std::vector<double> xyzcache(4 * (size_t)maxcellcache);
    int nneigh = 0;
    get_neighbour_list(*c, (int)icell, listneigh.data(), nneigh, xyzcache.data(), maxcellcache, getj ? 4 : 3, getj != 0, nullptr);
    for (int64_t k = 0; k < nneigh && k < maxlist; k++) list[k] = listneigh[k];
    return nneigh;
}

int oracle_densityiterate(oracle_ctx *cp, int icall, int64_t npart, double *xyzh, const double *vxyzu, const double *fxyzu,
                          const double *fext, const double *Bevol, const int8_t *iphase, float *divcurlv, float *divcurlB,
                          float *alphaind, float *gradh, float *dvdx, double *dustfrac, oracle_scalars *out)
{
    oracle_ctx &c = *cp;
    c.err.clear();
    if (npart != c.npart) { c.err = "densityiterate: npart differs from tree"; return 1; }
    // dens.F90:201-204
    const bool getdv = ((!c.p.const_av) && (icall <= 1 || icall == 3)) || (c.p.dust != 0);
    const bool getdB = (c.p.mhd != 0);
    DensArrays a{xyzh, vxyzu, fxyzu, fext, Bevol, iphase, divcurlv, divcurlB, alphaind, gradh, dvdx, dustfrac};
    int64_t nneightry = 0, nneighact = 0, maxneightry = 0, np = 0, ncalc = 0, ncalls_neigh = 0, npairs = 0;
    int maxneighact = 0, failed = 0;
    double rhomax = 0.;
#pragma omp parallel default(shared) reduction(+ : nneightry, nneighact, np, ncalc, ncalls_neigh, npairs) \
    reduction(max : maxneightry, maxneighact, rhomax)
    {
        std::vector<int> listneigh(c.npart + 16);
        std::vector<double> xyzcache(3 * (size_t)isizecellcache);
        CellDens cell;
#pragma omp for schedule(dynamic, 16)
        for (int64_t icell = 1; icell <= c.ncells; icell++) {
            if (c.leaf_is_active[icell] <= 0 || failed) continue;
            int nneigh = 0;
            get_neighbour_list(c, (int)icell, listneigh.data(), nneigh, xyzcache.data(), isizecellcache, 3, false, nullptr);
            cell.icell = (int)icell; cell.nits = 0; cell.nneightry = 0;
            for (int k = 0; k < minpart; k++) cell.nneigh[k] = 0;
            dens_start_cell(c, cell, a);
            { const Node &nd = c.node[icell]; cell.xpos[0] = nd.xcen[0]; cell.xpos[1] = nd.xcen[1]; cell.xpos[2] = nd.xcen[2];
              cell.xsizei = nd.size; cell.rcuti = c.kc.radkern * nd.hmax; cell.hmax = nd.hmax; }
            dens_compute_cell(c, cell, listneigh.data(), nneigh, getdv, getdB, a, xyzcache.data(), npairs);
            bool converged = false;
            while (!converged) {                              // local_its (dens.F90:338-373)
                if (dens_finish_cell(c, cell, converged)) { failed = 1; break; }
                bool redo_neighbours;
                dens_compute_hmax(c, cell, redo_neighbours);
                if (icall == 0) converged = true;
                if (!converged) {
                    if (redo_neighbours) {
                        set_hmaxcell(c, cell.icell, cell.hmax);
                        get_neighbour_list(c, -1, listneigh.data(), nneigh, xyzcache.data(), isizecellcache, 3, false, nullptr, cell.xpos, cell.xsizei, cell.rcuti);
                        ncalls_neigh = ncalls_neigh + 1;
                    }
                    dens_compute_cell(c, cell, listneigh.data(), nneigh, getdv, getdB, a, xyzcache.data(), npairs);
                }
            }
            if (failed) continue;
            dens_store_results(c, cell, getdv, getdB, a, rhomax, nneightry, nneighact, maxneightry, maxneighact, np, ncalc);
        }
    }
    if (failed) return 1;
    c.nptot = np; c.nneightry = nneightry; c.nneighact = nneighact; c.maxneightry = maxneightry; c.maxneighact = maxneighact;
    c.ncalc = ncalc; c.ncalls_neigh = ncalls_neigh;
    if (out) {
        out->rhomax = rhomax; out->np = np;
        out->trialmean = np > 0 ? (double)nneightry / (double)np : -1; out->actualmean = np > 0 ? (double)nneighact / (double)np : -1;
        out->maxtrial = maxneightry; out->maxactual = maxneighact; out->nrhocalc = ncalc; out->nactualtot = nneighact;
        out->ncalls_neigh = ncalls_neigh; out->npairs_density = npairs;
    }
    return 0;
}

// cons2prim_everything (cons2prim.f90:274-456) for ieos 1,2,3, no radiation / non-ideal MHD / one-fluid dust
int oracle_cons2prim(oracle_ctx *cp, int64_t npart, const double *xyzh, const double *vxyzu, const float *dvdx, const double *Bevol,
                     const int8_t *iphase, double *eos_vars, float *alphaind, double *Bxyz)
{
    oracle_ctx &c = *cp;
    const oracle_params &p = c.p;
    const int nvu = c.nvu();
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < npart; i++) {
        const double *x = xyzh + 4 * i;
        if (x[3] < DBL_MIN) continue;
        bool iactivei, iamgasi, iamdusti; int iamtypei;
        get_partinfo(p, iphase[i], iactivei, iamgasi, iamdusti, iamtypei);
        const double hi = x[3], pmassi = p.massoftype[iamtypei];
        const double rhoi = c.rhoh(hi, pmassi);
        const double rhogas = rhoi;
        if (!iamgasi) continue;
        double ponrhoi, spsoundi;
        // equationofstate (eos.f90:183-256)
        if (p.ieos == 1) { ponrhoi = p.polyk; spsoundi = std::sqrt(ponrhoi); }
        else if (p.ieos == 2) {
            if (nvu >= 4) {
                const double eni = vxyzu[nvu * i + 3];
                if (eni < 0.) bad = 1;
                if (p.gamma > 1.0001) ponrhoi = (p.gamma - 1.) * eni; else ponrhoi = 2. / 3. * eni;
            } else ponrhoi = p.polyk * std::pow(rhogas, p.gamma - 1.);
            spsoundi = std::sqrt(p.gamma * ponrhoi);
        } else {  // ieos == 3
            ponrhoi = p.polyk * std::pow(x[0] * x[0] + x[1] * x[1] + x[2] * x[2], -p.qfacdisc);
            ponrhoi = std::max(ponrhoi, p.cs_min * p.cs_min);
            spsoundi = std::sqrt(ponrhoi);
        }
        double *ev = eos_vars + 7 * i;
        ev[0] = ponrhoi * rhogas; ev[1] = spsoundi; ev[2] = p.temp_coef_mu * ponrhoi;   // igasP, ics, itemp (cons2prim.f90:390-392); the other rows are left alone
        if (c.nalpha() >= 2) {
            // xi_limiter (shock_capturing.f90:151-178)
            const float *d = dvdx + 9 * i;
            const double dvxdx = d[0], dvxdy = d[1], dvxdz = d[2], dvydx = d[3], dvydy = d[4], dvydz = d[5], dvzdx = d[6], dvzdy = d[7], dvzdz = d[8];
            const double divv = dvxdx + dvydy + dvzdz;
            const double curlvx = dvzdy - dvydz, curlvy = dvxdz - dvzdx, curlvz = dvydx - dvxdy;
            const double fac = pow2(std::max(-divv, 0.));
            const double traceS = curlvx * curlvx + curlvy * curlvy + curlvz * curlvz;
            const double xi_lim = (fac + traceS > DBL_EPSILON) ? fac / (fac + traceS) : 1.;
            // get_alphaloc (shock_capturing.f90:131-143)
            const double divvdti = (double)alphaind[3 * i + 2];
            const double source = 10. * hi * hi * xi_lim * std::max(-divvdti, 0.);
            const double temp = spsoundi * spsoundi;
            double alphaloc;
            if (temp > DBL_EPSILON) alphaloc = std::max(std::min(source / temp, p.alphamax), p.alpha);
            else alphaloc = p.alpha;
            alphaind[3 * i + 1] = (float)alphaloc;
        }
        if (p.mhd && Bxyz) {
            const double *B = Bevol + 4 * i;
            Bxyz[3 * i] = B[0] * rhoi; Bxyz[3 * i + 1] = B[1] * rhoi; Bxyz[3 * i + 2] = B[2] * rhoi;
        }
    }
    if (bad) { c.err = "eos: utherm < 0"; return 1; }
    return 0;
}

int64_t oracle_neighbour_sets(oracle_ctx *cp, int64_t npart, const double *xyzh, const int8_t *iphase, int symmetric,
                              int64_t *offsets, int32_t *list, int64_t maxlist)
{
    oracle_ctx &c = *cp;
    (void)iphase;
    std::vector<std::vector<int>> sets(npart);
#pragma omp parallel
    {
        std::vector<int> listneigh(c.npart + 16);
        std::vector<double> xyzcache(4 * (size_t)maxcellcache);
#pragma omp for schedule(dynamic, 16)
        for (int64_t icell = 1; icell <= c.ncells; icell++) {
            if (c.leaf_is_active[icell] == 0) continue;
            int nneigh = 0;
            getneigh(c, c.node[icell].xcen, c.node[icell].size, c.kc.radkern * c.node[icell].hmax, listneigh.data(), nneigh, xyzcache.data(),
                     maxcellcache, 4, symmetric != 0, false, nullptr);
            const int i1 = c.inoderange[2 * icell], i2 = c.inoderange[2 * icell + 1];
            for (int ip = i1; ip <= i2; ip++) {
                const int i = std::abs(c.inodeparts[ip - 1]);
                const double *xi = xyzh + 4 * (size_t)(i - 1);
                const double hi1 = 1. / xi[3], hi21 = hi1 * hi1;
                std::vector<int> &s = sets[i - 1];
                for (int n = 1; n <= nneigh; n++) {
                    const int j = listneigh[n - 1];
                    if (j == i) continue;
                    double dx, dy, dz, hj1;
                    if (n <= maxcellcache) { const double *xc = &xyzcache[4 * (size_t)(n - 1)]; dx = xi[0] - xc[0]; dy = xi[1] - xc[1]; dz = xi[2] - xc[2]; hj1 = xc[3]; }
                    else { const double *xj = xyzh + 4 * (size_t)(j - 1); dx = xi[0] - xj[0]; dy = xi[1] - xj[1]; dz = xi[2] - xj[2]; hj1 = 1. / xj[3]; }
                    if (c.p.periodic) {
                        if (std::fabs(dx) > 0.5 * c.dxbound) dx = dx - c.dxbound * std::copysign(1.0, dx);
                        if (std::fabs(dy) > 0.5 * c.dybound) dy = dy - c.dybound * std::copysign(1.0, dy);
                        if (std::fabs(dz) > 0.5 * c.dzbound) dz = dz - c.dzbound * std::copysign(1.0, dz);
                    }
                    const double rij2 = dx * dx + dy * dy + dz * dz;
                    const double q2i = rij2 * hi21;
                    bool isn = q2i < c.kc.radkern2;
                    if (symmetric) { const double hj21 = hj1 * hj1; isn = isn || (rij2 * hj21 < c.kc.radkern2); }
                    if (isn) s.push_back(j);
                }
                std::sort(s.begin(), s.end());
            }
        }
    }
    int64_t tot = 0;
    for (int64_t i = 0; i < npart; i++) { offsets[i] = tot; tot += (int64_t)sets[i].size(); }
    offsets[npart] = tot;
    if (tot > maxlist) return -tot;
    for (int64_t i = 0; i < npart; i++) std::copy(sets[i].begin(), sets[i].end(), list + offsets[i]);
    return tot;
}

int64_t oracle_neighbour_counts_bruteforce(oracle_ctx *cp, int64_t npart, const double *xyzh, int symmetric, int32_t *counts)
{
    oracle_ctx &c = *cp;
    int64_t tot = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tot)
    for (int64_t i = 0; i < npart; i++) {
        const double *xi = xyzh + 4 * i;
        const double hi1 = 1. / xi[3], hi21 = hi1 * hi1;
        int n = 0;
        for (int64_t j = 0; j < npart; j++) {
            if (j == i) continue;
            const double *xj = xyzh + 4 * j;
            double dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
            if (c.p.periodic) {
                if (std::fabs(dx) > 0.5 * c.dxbound) dx = dx - c.dxbound * std::copysign(1.0, dx);
                if (std::fabs(dy) > 0.5 * c.dybound) dy = dy - c.dybound * std::copysign(1.0, dy);
                if (std::fabs(dz) > 0.5 * c.dzbound) dz = dz - c.dzbound * std::copysign(1.0, dz);
            }
            const double rij2 = dx * dx + dy * dy + dz * dz;
            bool isn = rij2 * hi21 < c.kc.radkern2;
            if (symmetric) { const double hj1 = 1. / xj[3]; isn = isn || (rij2 * (hj1 * hj1) < c.kc.radkern2); }
            if (isn) n++;
        }
        counts[i] = n; tot += n;
    }
    return tot;
}

void oracle_kernel(int kernel, double q2, double q, double *w, double *gr, double *dphidh, double *potensoft, double *fsoft, double *wdrag)
{
    get_kernel(kernel, q2, q, *w, *gr);
    *dphidh = dphidh_kernel(kernel, q2, q);
    kernel_softening(kernel, q2, q, *potensoft, *fsoft);
    *wdrag = wkern_drag(kernel, q2, q);
}
void oracle_kernel_constants(int kernel, double *radkern, double *cnormk, double *wab0, double *gradh0, double *dphidh0, double *cnormk_drag,
                             double *hfact_default)
{
    KernelConsts k = kernel_consts(kernel);
    *radkern = k.radkern; *cnormk = k.cnormk; *wab0 = k.wab0; *gradh0 = k.gradh0; *dphidh0 = k.dphidh0; *cnormk_drag = k.cnormk_drag;
    *hfact_default = k.hfact_default;
}
void oracle_compute_M2L(double dx, double dy, double dz, double dr, double totmass, const double *quads, double *fnode20)
{
    compute_M2L(dx, dy, dz, dr, totmass, quads, fnode20);
}
void oracle_expand_fgrav(const double *fnode20, double dx, double dy, double dz, double *out4)
{
    expand_fgrav_in_taylor_series(fnode20, dx, dy, dz, out4[0], out4[1], out4[2], out4[3]);
}
void oracle_propagate_fnode(double *dst, const double *src, double dx, double dy, double dz) { propagate_fnode_to_node(dst, src, dx, dy, dz); }

// ran2 / get_random (random.f90:38-114); the second state is module-level and reset when the seed is negative
// st_calcAccel (forcing.f90:728-830)
void oracle_forcing(oracle_ctx *cp, int64_t npart, const double *xyzh, const int8_t *iphase, double *fxyzu, int nmodes, const double *mode,
                    const double *ampl, const double *aka, const double *akb, double amplfac, double solweightnorm, int correct_mean_force)
{
    const oracle_ctx &c = *cp;
    const int nvu = c.nvu();
    double fm0 = 0., fm1 = 0., fm2 = 0.;
#pragma omp parallel for schedule(static) reduction(+ : fm0, fm1, fm2)
    for (int64_t i = 0; i < npart; i++) {
        const bool active = c.p.ind_timesteps ? (iphase[i] > 0) : true;
        if (!(active || correct_mean_force)) continue;
        const double *x = xyzh + 4 * i;
        double fxi = 0., fyi = 0., fzi = 0.;
        for (int m = 0; m < nmodes; m++) {
            const double kdotx = mode[3 * m] * x[0] + mode[3 * m + 1] * x[1] + mode[3 * m + 2] * x[2];
            const double re = std::cos(kdotx), im = std::sin(kdotx);
            fxi += ampl[m] * (aka[3 * m] * re - akb[3 * m] * im);
            fyi += ampl[m] * (aka[3 * m + 1] * re - akb[3 * m + 1] * im);
            fzi += ampl[m] * (aka[3 * m + 2] * re - akb[3 * m + 2] * im);
        }
        fxi = 2. * amplfac * solweightnorm * fxi; fyi = 2. * amplfac * solweightnorm * fyi; fzi = 2. * amplfac * solweightnorm * fzi;
        if (active) { fxyzu[nvu * i] = fxi; fxyzu[nvu * i + 1] = fyi; fxyzu[nvu * i + 2] = fzi; }
        if (correct_mean_force) { fm0 += fxi; fm1 += fyi; fm2 += fzi; }
    }
    if (correct_mean_force) {
        fm0 /= (double)npart; fm1 /= (double)npart; fm2 /= (double)npart;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < npart; i++) {
            const bool active = c.p.ind_timesteps ? (iphase[i] > 0) : true;
            if (active) { fxyzu[nvu * i] -= fm0; fxyzu[nvu * i + 1] -= fm1; fxyzu[nvu * i + 2] -= fm2; }
        }
    }
}

double oracle_ran2(int32_t *s1p)
{
    static int32_t s2 = 123456789;
    int32_t s1 = *s1p;
    if (s1 < 0) s2 = 123456789;
    int32_t k = s1 / 53668;
    s1 = 40014 * (s1 - k * 53668) - k * 12211;
    if (s1 < 0) s1 = s1 + 2147483563;
    k = s2 / 52774;
    s2 = 40692 * (s2 - k * 52774) - k * 3791;
    if (s2 < 0) s2 = s2 + 2147483399;
    int32_t z = s1 - s2;
    if (z < 1) z = z + 2147483562;
    *s1p = s1;
    return (double)z / 2147483563.0;
}

}  // extern "C"
