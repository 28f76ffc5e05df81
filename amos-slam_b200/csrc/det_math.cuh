// det_math.cuh -- bit-exact device arithmetic for the orientation / rBRIEF stage.
//
// fast_atan2_deg : cv::fastAtan2 (OpenCV, degrees) as called at /root/reference/src/ORBextractor.cc:160.
//                  Float Horner polynomial, coefficients are FLOAT products ck * (180/pi), every op
//                  rounded to float, NO fma contraction (nvcc contracts by default, hence the
//                  explicit __fmul_rn/__fadd_rn/__fdiv_rn).
// det_sincos     : sin/cos of the float angle used at ORBextractor.cc:181, evaluated in double with
//                  explicit fma() (Cody-Waite pi/2 reduction + fdlibm kernel polynomials) and rounded
//                  to float: the correctly rounded float value for all practical purposes, and the
//                  same operation sequence as the CPU oracle, so both sides agree bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <float.h>

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    // (float)ck * (float)(180/pi), folded at compile time in float
    const float r2d = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * r2d;
    const float p3 = -0.3258083974640975f * r2d;
    const float p5 = 0.1555786518463281f * r2d;
    const float p7 = -0.04432655554792128f * r2d;
    const float eps = (float)DBL_EPSILON;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        float t = __fadd_rn(__fmul_rn(p7, c2), p5);
        t = __fadd_rn(__fmul_rn(t, c2), p3);
        t = __fadd_rn(__fmul_rn(t, c2), p1);
        a = __fmul_rn(t, c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        float t = __fadd_rn(__fmul_rn(p7, c2), p5);
        t = __fadd_rn(__fmul_rn(t, c2), p3);
        t = __fadd_rn(__fmul_rn(t, c2), p1);
        a = __fsub_rn(90.f, __fmul_rn(t, c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

__device__ __forceinline__ void det_sincos(float xf, float* s_out, float* c_out) {
    const double x = (double)xf;
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_1 = 1.57079632673412561417e+00;
    const double PIO2_2 = 6.07710050650619224932e-11;
    const double PIO2_3 = 2.02226624879595063154e-21;
    double kd = rint(__dmul_rn(x, TWO_OVER_PI));
    int k = (int)kd;
    double r = fma(-kd, PIO2_1, x);
    r = fma(-kd, PIO2_2, r);
    r = fma(-kd, PIO2_3, r);
    const double z = __dmul_rn(r, r);
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double ps = fma(z, S6, S5); ps = fma(z, ps, S4); ps = fma(z, ps, S3); ps = fma(z, ps, S2); ps = fma(z, ps, S1);
    double sn = fma(__dmul_rn(r, z), ps, r);
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double pc = fma(z, C6, C5); pc = fma(z, pc, C4); pc = fma(z, pc, C3); pc = fma(z, pc, C2); pc = fma(z, pc, C1);
    double cs = fma(__dmul_rn(z, z), pc, fma(-0.5, z, 1.0));
    double s, c;
    switch (k & 3) {
        case 0: s = sn; c = cs; break;
        case 1: s = cs; c = -sn; break;
        case 2: s = -sn; c = -cs; break;
        default: s = -cs; c = sn; break;
    }
    *s_out = (float)s; *c_out = (float)c;   // double -> float: round to nearest even (cvt.rn.f32.f64)
}
