/* roofline_constants.h -- ALGORITHMIC work per unit used by bench.py's roofline.
 * Only real interacting pairs are counted (work any correct implementation must do); FMA = 2 flop,
 * sqrt = div = 1.  Pair counts are measured by the kernels at run time (sphgpu_scalars.npairs_*).
 * Derived by reading the loops cited in DESIGN.md; frozen here so the numbers cannot drift silently. */
#ifndef ROOFLINE_CONSTANTS_H
#define ROOFLINE_CONSTANTS_H
#define FLOP_DENS_PAIR_HYDRO      95   /* dens.F90:675-803 rho,gradh + 9 dv + 9 da + 6 r sums (getdv, nalpha=3) */
#define FLOP_DENS_PAIR_CONSTAV    35   /* no da/dvdx sums */
#define FLOP_DENS_PAIR_MHD_EXTRA  35   /* dens.F90:806-829 */
#define FLOP_DENS_EPILOGUE        60   /* per particle per iteration: finish_rhosum + finish_cell (dens.F90:1401-1507) */
#define FLOP_FORCE_PAIR_ADIABATIC 140  /* force.F90:1287-1741 hydro + AV + conductivity */
#define FLOP_FORCE_PAIR_ISOTHERMAL 110
#define FLOP_FORCE_PAIR_MHD_EXTRA 150  /* force.F90:1428-1444,1626-1684 */
#define FLOP_FORCE_EPILOGUE       150  /* finish_cell_and_store_results (force.F90:2939-3223) */
#define FLOP_FORCE_PAIR_GRAV_EXTRA 30  /* softened gravity of SPH-neighbour pairs (force.F90:1303-1339) */
#define FLOP_FORCE_PAIR_DRAG_EXTRA 80  /* two-fluid drag pair incl. reconstruct_dv and get_ts (force.F90:1864-1970) */
#define FLOP_GRAV_P2P_PAIR         25  /* Newtonian m/r^2 pair outside both kernels (force.F90:1992-2053) */
#define FLOP_GRAV_M2L             130  /* compute_M2L per accepted node pair (kdtree.F90:1702-1781) */
#define BYTES_TREE_PER_PARTICLE   150  /* keys + sort passes + gather */
#define BYTES_DENS_PER_PARTICLE   180  /* each array touched once */
#define BYTES_DENS_MHD_EXTRA       48
#define BYTES_C2P_PER_PARTICLE    120
#define BYTES_FORCE_PER_PARTICLE  175
#define BYTES_FORCE_MHD_EXTRA      85
#endif
