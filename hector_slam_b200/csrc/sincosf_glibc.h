// sincosf_glibc.h — sinf / cosf evaluated the way glibc 2.39's x86-64 FMA build does.
//
// Why: the reference calls sinf/cosf on the pose angle (OccGridMapUtil.h:70-71, Rotation2Df at
// :351, OccGridMapBase.h:130-131).  glibc's single-precision sin/cos (Szabolcs Nagy's routines,
// sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h) are NOT correctly rounded: they differ
// from the correctly rounded value on 1.3 % of the arguments in [-pi, pi] (measured here).  One
// ulp of cos/sin moves a map coordinate of ~600 cells by ~6e-5 cell, and on a sparsely mapped
// grid that can move the matcher's answer by more than the 1e-4 parity bar (measured: 1.3e-3 on
// one step of the 40 Hz stream test).  So the pose angle's sine and cosine are evaluated with
// glibc's own algorithm: argument reduced with one fused multiply-add against pi/2, degree-7 /
// degree-8 polynomials in DOUBLE with exactly the multiply / fused-multiply-add sequence of the
// FMA-enabled build (transcribed from the disassembly of __sinf_fma / __cosf_fma in
// libm.so.6 2.39-0ubuntu8.5, which every AVX2+FMA host selects through ifunc), constants read
// from that library's __sincosf_table.  tests/test_sincosf_glibc.py checks the host build of
// this header against the running libm bit for bit over tens of millions of arguments.
//
// Provenance / licence: this header is a restatement of the GNU C Library's single-precision
// sin/cos algorithm (glibc 2.39, sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,sincosf.h,
// sincosf_data.c}; Copyright (C) Free Software Foundation, Inc., contributed by Arm Ltd.,
// licensed LGPL-2.1-or-later).  The polynomial coefficients and the operation order are glibc's
// (they have to be, that is the point); no glibc source text is included.  Treat this file as
// LGPL-2.1-or-later derived material when redistributing.
// Platform note: "as glibc evaluates them" means the x86-64 FMA ifunc variant; on an aarch64 host
// (Grace) glibc's generic build uses the same algorithm with compiler-chosen contraction, so a
// reference running there may differ in the last bit and the table below would need re-pinning.
//
// Arguments with |y| >= 120 (glibc's table-driven large-argument reduction) fall back to a
// correctly rounded evaluation; pose angles are normalised to (-pi, pi] once per level.
#ifndef HSB_SINCOSF_GLIBC_H
#define HSB_SINCOSF_GLIBC_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HSB_HD __host__ __device__ __forceinline__
#else
#define HSB_HD static inline
#endif

namespace hsb {

#if defined(__CUDA_ARCH__)
#define HSB_DMUL(a, b) __dmul_rn((a), (b))
#define HSB_DFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define HSB_DMUL(a, b) ((a) * (b))
#define HSB_DFMA(a, b, c) fma((a), (b), (c))
#endif

// __sincosf_table[0]; table[1] has c0..c4 negated
#define HSB_SC_HPI_INV 0x1.45f306dc9c883p+23 /* 2/pi * 2^24 */
#define HSB_SC_HPI 0x1.921fb54442d18p+0      /* pi/2 */
#define HSB_SC_C0 0x1.0000000000000p+0
#define HSB_SC_C1 -0x1.ffffffd0c621cp-2
#define HSB_SC_C2 0x1.55553e1068f19p-5
#define HSB_SC_C3 -0x1.6c087e89a359dp-10
#define HSB_SC_C4 0x1.99343027bf8c3p-16
#define HSB_SC_S1 -0x1.555545995a603p-3
#define HSB_SC_S2 0x1.1107605230bc4p-7
#define HSB_SC_S3 -0x1.994eb3774cf24p-13

HSB_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

// sinf_poly with n even: x + x^3*s1 + x^7*(s2 + x^2*s3)
HSB_HD double sc_sin_poly(double xs, double x2) {
  const double s1_ = HSB_DFMA(x2, HSB_SC_S3, HSB_SC_S2);
  const double x3 = HSB_DMUL(x2, xs);
  const double x7 = HSB_DMUL(x2, x3);
  const double s = HSB_DFMA(x3, HSB_SC_S1, xs);
  return HSB_DFMA(s1_, x7, s);
}
// sinf_poly with n odd: (c0 + x^2*c1) + x^4*c2 + x^6*(c3 + x^2*c4); `neg` selects table[1]
HSB_HD double sc_cos_poly(double x2, bool neg) {
  const double sg = neg ? -1.0 : 1.0;  // negating every coefficient negates every partial result exactly
  const double x4 = HSB_DMUL(x2, x2);
  const double c1_ = HSB_DFMA(x2, sg * HSB_SC_C1, sg * HSB_SC_C0);
  const double c2_ = HSB_DFMA(x2, sg * HSB_SC_C4, sg * HSB_SC_C3);
  const double x6 = HSB_DMUL(x2, x4);
  const double c = HSB_DFMA(x4, sg * HSB_SC_C2, c1_);
  return HSB_DFMA(c2_, x6, c);
}

// reduce_fast: n = round(x * 2/pi), xr = x - n*pi/2 (one fused operation)
HSB_HD double sc_reduce(double x, int* np) {
  const double r = HSB_DMUL(x, HSB_SC_HPI_INV);
  const int n = ((int32_t)r + 0x800000) >> 24;  // cvttsd2si, then arithmetic shift
  *np = n;
  return HSB_DFMA(-(double)n, HSB_SC_HPI, x);
}

HSB_HD float sinf_glibc(float y) {
  const double x = (double)y;
  const uint32_t top = (f32_bits(y) >> 20) & 0x7ffu;
  if (top <= 0x3f3u) {                 // |y| < pi/4
    if (top <= 0x397u) return y;       // |y| < 2^-12
    return (float)sc_sin_poly(x, HSB_DMUL(x, x));
  }
  if (top <= 0x42eu) {                 // |y| < 120
    int n;
    const double xr = sc_reduce(x, &n);
    const double x2 = HSB_DMUL(xr, xr);
    if ((n & 1) == 0) {
      const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign[4] = {1,-1,-1,1}
      return (float)sc_sin_poly(HSB_DMUL(xr, sign), x2);
    }
    return (float)sc_cos_poly(x2, (n & 2) != 0);
  }
  return (float)sin(x);
}

HSB_HD float cosf_glibc(float y) {
  const double x = (double)y;
  const uint32_t top = (f32_bits(y) >> 20) & 0x7ffu;
  if (top <= 0x3f3u) {
    if (top <= 0x397u) return 1.0f;
    return (float)sc_cos_poly(HSB_DMUL(x, x), false);
  }
  if (top <= 0x42eu) {
    int n;
    const double xr = sc_reduce(x, &n);
    const double x2 = HSB_DMUL(xr, xr);
    if ((n & 1) != 0) {
      const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
      return (float)sc_sin_poly(HSB_DMUL(xr, sign), x2);
    }
    return (float)sc_cos_poly(x2, (n & 2) != 0);
  }
  return (float)cos(x);
}

// sinf_glibc and cosf_glibc of the same argument with ONE argument reduction: the two polynomials are the ones the
// separate functions evaluate (n even: sin from the sine polynomial, cos from the cosine polynomial; n odd: swapped),
// so both results are bit-identical to the separate calls (tests/test_sincosf_glibc.py).
HSB_HD void sincosf_glibc(float y, float* sp, float* cp) {
  const double x = (double)y;
  const uint32_t top = (f32_bits(y) >> 20) & 0x7ffu;
  if (top <= 0x3f3u) {                 // |y| < pi/4
    if (top <= 0x397u) {               // |y| < 2^-12
      *sp = y;
      *cp = 1.0f;
      return;
    }
    const double x2 = HSB_DMUL(x, x);
    *sp = (float)sc_sin_poly(x, x2);
    *cp = (float)sc_cos_poly(x2, false);
    return;
  }
  if (top <= 0x42eu) {                 // |y| < 120
    int n;
    const double xr = sc_reduce(x, &n);
    const double x2 = HSB_DMUL(xr, xr);
    const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const float a = (float)sc_sin_poly(HSB_DMUL(xr, sign), x2);
    const float b = (float)sc_cos_poly(x2, (n & 2) != 0);
    const bool even = (n & 1) == 0;
    *sp = even ? a : b;
    *cp = even ? b : a;
    return;
  }
  *sp = (float)sin(x);
  *cp = (float)cos(x);
}

}  // namespace hsb
#endif
