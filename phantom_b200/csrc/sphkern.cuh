// sphkern.cuh -- SPH kernel functions as device code, specialised at compile time on the kernel id
// (the reference selects one kernel module at link time, build/Makefile:283-290).
//   K = 0: M4 cubic   (src/main/kernel_cubic.f90:35-158)   radkern = 2
//   K = 1: M6 quintic (src/main/kernel_quintic.f90:33-206)  radkern = 3
// The kernels are piecewise polynomials: "kernel tables" reduce to immediate coefficients.
#pragma once

template <int K> struct SphKern;

__device__ __forceinline__ double p2(double x) { return x * x; }
__device__ __forceinline__ double p3(double x) { return x * x * x; }
__device__ __forceinline__ double p4(double x) { double y = x * x; return y * y; }
__device__ __forceinline__ double p5(double x) { double y = x * x; return y * y * x; }
// max(x, 0) on the integer pipe (sign bit of the high word), keeping the FP64 pipe for the sums
__device__ __forceinline__ double relu_d(double x)
{
    const int hi = __double2hiint(x), keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(x) & keep);
}

template <> struct SphKern<0> {
    static constexpr double radkern = 2.0, radkern2 = 4.0;
    static constexpr double cnormk = 0.31830988618379067153776752674502872;   // 1/pi
    static constexpr double wab0 = 1.0, gradh0 = -3.0, dphidh0 = 1.4;
    static constexpr double cnormk_drag = 10. / (9. * 3.14159265358979323846264338327950288);
    __device__ __forceinline__ static void get_kernel(double q2, double q, double &w, double &gr)
    {
        if (q < 1.) { w = 0.75 * q2 * q - 1.5 * q2 + 1.; gr = q * (2.25 * q - 3.); }
        else if (q < 2.) { w = -0.25 * p3(q - 2.); gr = -0.75 * p2(q - 2.); }
        else { w = 0.; gr = 0.; }
    }
    // the same spline written as truncated powers, w = (2-q)+^3/4 - (1-q)+^3: no branches, zero beyond the support, any q >= 0
    __device__ __forceinline__ static void get_kernel_bf(double q, double &w, double &gr)
    {
        const double a = relu_d(2. - q), b = relu_d(1. - q);
        const double a2 = a * a, b2 = b * b;
        w = fma(0.25 * a, a2, -(b2 * b));
        gr = fma(-0.75, a2, 3. * b2);
    }
    __device__ __forceinline__ static double grkern_bf(double q)
    {
        const double a = relu_d(2. - q), b = relu_d(1. - q);
        return fma(-0.75 * a, a, 3. * (b * b));
    }
    __device__ __forceinline__ static double grkern(double q2, double q)
    {
        if (q < 1.) return q * (2.25 * q - 3.);
        else if (q < 2.) return -0.75 * p2(q - 2.);
        return 0.;
    }
    __device__ __forceinline__ static double dphidh(double q2, double q)
    {
        const double q4 = q2 * q2;
        if (q < 1.) return -0.6 * q4 * q + 1.5 * q4 - 2. * q2 + 1.4;
        else if (q < 2.) return 0.2 * q4 * q - 1.5 * q4 + 4. * q2 * q - 4. * q2 + 1.6;
        return 0.;
    }
    __device__ __forceinline__ static void softening(double q2, double q, double &pot, double &fs)
    {
        if (q < 1.) {
            const double q4 = q2 * q2;
            pot = q4 * q / 10. - 3. * q4 / 10. + 2. * q2 / 3. - 7. / 5.;
            fs = q * (15. * q2 * q - 36. * q2 + 40.) / 30.;
        } else if (q < 2.) {
            const double q4 = q2 * q2, q6 = q4 * q2;
            pot = (q * (-q4 * q + 9. * q4 - 30. * q2 * q + 40. * q2 - 48.) + 2.) / (30. * q);
            fs = (-5. * q6 + 36. * q4 * q - 90. * q4 + 80. * q2 * q - 2.) / (30. * q2);
        } else { pot = -1. / q; fs = 1. / q2; }
    }
    // the same inside the support (q < 2), both pieces evaluated and selected: no branch between lanes on different pieces, no divisions
    // (qinv = 1/q from the caller's 1/r); force.F90:1303-1339 calls it only for q2 < radkern2
    __device__ __forceinline__ static void softening_in(double q2, double q, double qinv, double &pot, double &fs)
    {
        const double potA = fma(q2, fma(q2, fma(q, 0.1, -0.3), 2. / 3.), -1.4);
        const double fsA = q * fma(q2, fma(q, 0.5, -1.2), 4. / 3.);
        const double polB = fma(q2, fma(q, fma(q, 9. - q, -30.), 40.), -48.);
        const double potB = fma(polB, 1. / 30., qinv * (1. / 15.));
        const double pfB = fma(q2 * q, fma(q, fma(q, fma(q, -5., 36.), -90.), 80.), -2.);
        const double fsB = pfB * (qinv * qinv) * (1. / 30.);
        const bool inner = q < 1.;
        pot = inner ? potA : potB;
        fs = inner ? fsA : fsB;
    }
    __device__ __forceinline__ static double wdrag(double q2, double q)
    {
        if (q < 1.) return q2 * (0.75 * q2 * q - 1.5 * q2 + 1.);
        else if (q < 2.) return -0.25 * q2 * p3(q - 2.);
        return 0.;
    }
};

template <> struct SphKern<1> {
    static constexpr double radkern = 3.0, radkern2 = 9.0;
    static constexpr double cnormk = 1. / (120. * 3.14159265358979323846264338327950288);
    static constexpr double wab0 = 66.0, gradh0 = -198.0, dphidh0 = 239. / 210.;
    static constexpr double cnormk_drag = 1. / (168. * 3.14159265358979323846264338327950288);
    __device__ __forceinline__ static void get_kernel(double q2, double q, double &w, double &gr)
    {
        if (q < 1.) { const double q4 = q2 * q2; w = -10. * q4 * q + 30. * q4 - 60. * q2 + 66.; gr = q * (-50. * q2 * q + 120. * q2 - 120.); }
        else if (q < 2.) { w = -p5(q - 3.) + 6. * p5(q - 2.); gr = -5. * p4(q - 3.) + 30. * p4(q - 2.); }
        else if (q < 3.) { w = -p5(q - 3.); gr = -5. * p4(q - 3.); }
        else { w = 0.; gr = 0.; }
    }
    // truncated powers: w = (3-q)+^5 - 6 (2-q)+^5 + 15 (1-q)+^5
    __device__ __forceinline__ static void get_kernel_bf(double q, double &w, double &gr)
    {
        const double a = relu_d(3. - q), b = relu_d(2. - q), c = relu_d(1. - q);
        const double a2 = a * a, b2 = b * b, c2 = c * c;
        const double a4 = a2 * a2, b4 = b2 * b2, c4 = c2 * c2;
        w = fma(a4, a, fma(-6. * b4, b, 15. * c4 * c));
        gr = fma(-5., a4, fma(30., b4, -75. * c4));
    }
    __device__ __forceinline__ static double grkern_bf(double q)
    {
        const double a = relu_d(3. - q), b = relu_d(2. - q), c = relu_d(1. - q);
        const double a2 = a * a, b2 = b * b, c2 = c * c;
        return fma(-5. * a2, a2, fma(30. * b2, b2, -75. * (c2 * c2)));
    }
    __device__ __forceinline__ static double grkern(double q2, double q)
    {
        if (q < 1.) return q * (-50. * q2 * q + 120. * q2 - 120.);
        else if (q < 2.) return -5. * p4(q - 3.) + 30. * p4(q - 2.);
        else if (q < 3.) return -5. * p4(q - 3.);
        return 0.;
    }
    __device__ __forceinline__ static double dphidh(double q2, double q)
    {
        const double q4 = q2 * q2, q6 = q4 * q2;
        if (q < 1.) return q6 * q / 21. - q6 / 6. + q4 / 2. - 11. * q2 / 10. + 239. / 210.;
        else if (q < 2.) return -q6 * q / 42. + q6 / 4. - q4 * q + 7. * q4 / 4. - 5. * q2 * q / 6. - 17. * q2 / 20. + 473. / 420.;
        else if (q < 3.) return q6 * q / 210. - q6 / 12. + 3. * q4 * q / 5. - 9. * q4 / 4. + 9. * q2 * q / 2. - 81. * q2 / 20. + 243. / 140.;
        return 0.;
    }
    __device__ __forceinline__ static void softening(double q2, double q, double &pot, double &fs)
    {
        if (q < 1.) {
            const double q4 = q2 * q2, q6 = q4 * q2;
            pot = -q6 * q / 168. + q6 / 42. - q4 / 10. + 11. * q2 / 30. - 239. / 210.;
            fs = q * (-35. * q4 * q + 120. * q4 - 336. * q2 + 616.) / 840.;
        } else if (q < 2.) {
            const double q4 = q2 * q2, q6 = q4 * q2, q8 = q6 * q2;
            pot = (q * (5. * q6 * q - 60. * q6 + 280. * q4 * q - 588. * q4 + 350. * q2 * q + 476. * q2 - 1892.) - 5.) / (1680. * q);
            fs = (35. * q8 - 360. * q6 * q + 1400. * q6 - 2352. * q4 * q + 1050. * q4 + 952. * q2 * q + 5.) / (1680. * q2);
        } else if (q < 3.) {
            const double q4 = q2 * q2, q6 = q4 * q2, q8 = q6 * q2;
            pot = (q * (-q6 * q + 20. * q6 - 168. * q4 * q + 756. * q4 - 1890. * q2 * q + 2268. * q2 - 2916.) + 507.) / (1680. * q);
            fs = (-7. * q8 + 120. * q6 * q - 840. * q6 + 3024. * q4 * q - 5670. * q4 + 4536. * q2 * q - 507.) / (1680. * q2);
        } else { pot = -1. / q; fs = 1. / q2; }
    }
    __device__ __forceinline__ static void softening_in(double q2, double q, double, double &pot, double &fs) { softening(q2, q, pot, fs); }
    __device__ __forceinline__ static double wdrag(double q2, double q)
    {
        if (q < 1.) { const double q4 = q2 * q2; return q2 * (-10. * q4 * q + 30. * q4 - 60. * q2 + 66.); }
        else if (q < 2.) return q2 * (-p5(q - 3.) + 6. * p5(q - 2.));
        else if (q < 3.) return -q2 * p5(q - 3.);
        return 0.;
    }
};
