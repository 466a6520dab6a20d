// prl_libm.h -- sinf/cosf as the reference computes them.
//
// rapier2d's `enhanced-determinism` feature (reference Cargo.toml:19) routes nalgebra's
// UnitComplex::new through the `libm` crate (0.2.7, musl/FreeBSD-msun derived) instead of the
// platform libm, so rotation integration and `ColliderBuilder::rotation` (common.rs:123, 133)
// evaluate sinf/cosf with double-precision minimax kernels after a pi/2 argument reduction.
// Restated here once for host (scene build) and device (pose integration) so both sides of the
// batch agree bit for bit. Must be compiled without FMA contraction.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PRL_HD __host__ __device__ __forceinline__
#else
#define PRL_HD inline
#endif

namespace prl {

PRL_HD uint32_t f32_bits(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
#endif
}

// |x| <= pi/4 kernels, evaluated in double and rounded once
PRL_HD float msun_ksin(double x) {
    const double s1 = -0.166666666416265235595, s2 = 0.0083333293858894631756,
                 s3 = -0.000198393348360966317347, s4 = 0.0000027183114939898219064;
    double z = x * x;
    double w = z * z;
    double r = s3 + z * s4;
    double s = z * x;
    return (float)((x + s * (s1 + z * s2)) + s * w * r);
}
PRL_HD float msun_kcos(double x) {
    const double c0 = -0.499999997251031003120, c1 = 0.0416666233237390631894,
                 c2 = -0.00138867637746099294692, c3 = 0.0000243904487962774090654;
    double z = x * x;
    double w = z * z;
    double r = c2 + z * c3;
    return (float)(((1.0 + z * c0) + w * c1) + (w * z) * r);
}
// medium-size reduction (|x| < 2^28 * pi/2); huge arguments never occur: |angvel * dt| is tiny
PRL_HD int msun_rem_pio2(float x, double* y) {
    const double toint = 1.5 / 2.220446049250313e-16, inv_pio2 = 6.36619772367581382433e-01,
                 pio2_hi = 1.57079631090164184570e+00, pio2_lo = 1.58932547735281966916e-08;
    double xd = (double)x;
    double fn = xd * inv_pio2 + toint - toint;
    *y = xd - fn * pio2_hi - fn * pio2_lo;
    return (int)fn;
}

PRL_HD float libm_sinf(float x) {
    const double pio2 = 1.57079632679489661923;
    double xd = (double)x;
    uint32_t ix = f32_bits(x);
    bool neg = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    if (ix <= 0x3f490fdau) {           // |x| <= pi/4
        if (ix < 0x39800000u) return x;  // |x| < 2^-12
        return msun_ksin(xd);
    }
    if (ix <= 0x407b53d1u) {           // |x| <= 5pi/4
        if (ix <= 0x4016cbe3u) return neg ? -msun_kcos(xd + 1.0 * pio2) : msun_kcos(xd - 1.0 * pio2);
        return msun_ksin(neg ? -(xd + 2.0 * pio2) : -(xd - 2.0 * pio2));
    }
    if (ix <= 0x40e231d5u) {           // |x| <= 9pi/4
        if (ix <= 0x40afeddfu) return neg ? msun_kcos(xd + 3.0 * pio2) : -msun_kcos(xd - 3.0 * pio2);
        return msun_ksin(neg ? xd + 4.0 * pio2 : xd - 4.0 * pio2);
    }
    if (ix >= 0x7f800000u) return x - x;
    double y;
    int n = msun_rem_pio2(x, &y);
    switch (n & 3) {
        case 0: return msun_ksin(y);
        case 1: return msun_kcos(y);
        case 2: return msun_ksin(-y);
        default: return -msun_kcos(y);
    }
}

PRL_HD float libm_cosf(float x) {
    const double pio2 = 1.57079632679489661923;
    double xd = (double)x;
    uint32_t ix = f32_bits(x);
    bool neg = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    if (ix <= 0x3f490fdau) {
        if (ix < 0x39800000u) return 1.0f;
        return msun_kcos(xd);
    }
    if (ix <= 0x407b53d1u) {
        if (ix > 0x4016cbe3u) return -msun_kcos(neg ? xd + 2.0 * pio2 : xd - 2.0 * pio2);
        return neg ? msun_ksin(xd + 1.0 * pio2) : msun_ksin(1.0 * pio2 - xd);
    }
    if (ix <= 0x40e231d5u) {
        if (ix > 0x40afeddfu) return msun_kcos(neg ? xd + 4.0 * pio2 : xd - 4.0 * pio2);
        return neg ? msun_ksin(-xd - 3.0 * pio2) : msun_ksin(xd - 3.0 * pio2);
    }
    if (ix >= 0x7f800000u) return x - x;
    double y;
    int n = msun_rem_pio2(x, &y);
    switch (n & 3) {
        case 0: return msun_kcos(y);
        case 1: return msun_ksin(-y);
        case 2: return -msun_kcos(y);
        default: return msun_ksin(y);
    }
}

}  // namespace prl
