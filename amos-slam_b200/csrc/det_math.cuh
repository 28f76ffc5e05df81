// det_math.cuh -- bit-exact device arithmetic for the orientation / rBRIEF stage.
//
// fast_atan2_deg : cv::fastAtan2 (OpenCV, degrees) as called at /root/reference/src/ORBextractor.cc:160.
//                  Float Horner polynomial, coefficients are FLOAT products ck * (180/pi), every op
//                  rounded to float, NO fma contraction (nvcc contracts by default, hence the
//                  explicit __fmul_rn/__fadd_rn/__fdiv_rn).
// det_sincos     : sin/cos of the float angle used at ORBextractor.cc:181 = glibc's sincosf (FMA variant),
//                  restated operation by operation (see the comment at the function).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

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

// __inv_pio4 of glibc (4/pi as overlapping 32-bit words), for the large-argument reduction
__constant__ uint32_t c_inv_pio4[24] = {0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u, 0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u,
                                        0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u, 0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u};

// glibc sincosf, FMA ifunc variant (__sincosf_fma of glibc 2.39; sysdeps/ieee754/flt-32/s_sincosf.c, sincosf_poly.h): the function the
// reference build calls for `cos(angle)`, `sin(angle)` at ORBextractor.cc:181 (GCC merges the pair into one sincosf).  Double-precision
// polynomial after a pi/2 reduction; in that variant every `a + b * c` is one fused multiply-add and every bare product is rounded, which
// is what fma() / __dmul_rn spell out here.  The CPU-side restatement of the same sequence is bit-exact against the live libm for all 2^32 inputs
// (the CPU tests); the GPU tests sweep this function against it on the device.
__device__ __forceinline__ void det_sincos(float y, float* sinp, float* cosp) {
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;       // __sincosf_table: 2/pi * 2^24, pi/2
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    const uint32_t xi = __float_as_uint(y), top = (xi >> 20) & 0x7ffu;              // abstop12
    double x = (double)y;
    int n = 0, q = 0;                                                              // q = quadrant index that selects sign[] and the table
    if (top < 0x3f4u) {                                                            // |y| < pi/4
        if (top < 0x398u) { *sinp = y; *cosp = 1.0f; return; }                     // |y| < 2^-12
    } else if (top < 0x42fu) {                                                     // |y| < 120: reduce_fast
        const double r = __dmul_rn(x, HPI_INV);
        n = (__double2int_rz(r) + 0x800000) >> 24;
        x = fma(-(double)n, HPI, x);
        q = n;
    } else if (top < 0x7f8u) {                                                     // reduce_large (never reached by an orientation angle, kept for fidelity)
        const int i0 = (xi >> 26) & 15, shift = (xi >> 23) & 7;
        const uint32_t m = ((xi & 0xffffffu) | 0x800000u) << shift;
        unsigned long long res0 = (uint32_t)(m * c_inv_pio4[i0]);
        const unsigned long long res1 = (unsigned long long)m * c_inv_pio4[i0 + 4], res2 = (unsigned long long)m * c_inv_pio4[i0 + 8];
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        const unsigned long long nn = (res0 + (1ULL << 61)) >> 62;
        res0 -= nn << 62;
        x = __dmul_rn((double)(long long)res0, 0x1.921FB54442D18p-62);
        n = (int)nn; q = n + (int)(xi >> 31);
    } else { *sinp = *cosp = __fsub_rn(y, y); return; }                            // inf / nan
    const double s = ((q + 1) & 2) ? -1.0 : 1.0;                                   // sign[q & 3] = {1, -1, -1, 1}
    const bool neg = (q & 2) != 0;                                                 // __sincosf_table[1]: cosine polynomial negated
    const double c0 = neg ? -C0 : C0, c1 = neg ? -C1 : C1, c2 = neg ? -C2 : C2, c3 = neg ? -C3 : C3, c4 = neg ? -C4 : C4;
    // sincosf_poly(x * s, x * x, p, n, sinp, cosp)
    const double xs = __dmul_rn(x, s), x2 = __dmul_rn(x, x);
    const double s1v = fma(x2, S3, S2), c2v = fma(x2, c4, c3);
    const double x3 = __dmul_rn(x2, xs), x4 = __dmul_rn(x2, x2);
    const double x5 = __dmul_rn(x2, x3), x6 = __dmul_rn(x2, x4);
    const double c1v = fma(x2, c1, c0);
    const double sv = fma(x3, S1, xs), cv = fma(x4, c2, c1v);
    const float fs = (float)fma(s1v, x5, sv), fc = (float)fma(c2v, x6, cv);      // cvt.rn.f32.f64
    if (n & 1) { *sinp = fc; *cosp = fs; } else { *sinp = fs; *cosp = fc; }
}
