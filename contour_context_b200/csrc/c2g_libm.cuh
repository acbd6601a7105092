// c2g_libm.cuh — bit-exact restatements of the glibc 2.39 libm functions whose result bits feed descriptors.
//
// The reference calls glibc through std:: — std::exp(double) inside gaussPDF<float> (include/tools/algos.h:53-56, retrieval
// keys), std::atan2(float, float) in the BCI build (include/cont2/contour_mng.h:860) and std::acos(float) in
// checkConstellCorrespSim (:1191-1192).  CUDA's libdevice versions differ from glibc by 1-2 ulp on a fraction of the
// inputs, which is enough to break "bit-exact retrieval keys", so the glibc algorithms themselves run on the device:
//   * exp      : glibc sysdeps/ieee754/dbl-64/e_exp.c (Szabolcs Nagy's table-driven exp, N = 128, degree-5 polynomial); the
//                table and coefficients are extracted from this image's libm.so.6 by tools/extract_glibc_exp_table.py.
//                x86-64 glibc selects an FMA-compiled variant (__exp_fma) at run time on CPUs with FMA, and the two variants
//                differ in 0.07 % of the results, so BOTH are provided and c2g_create() probes the host's exp() to pick
//                the one this machine's reference build would use (mode 0 = libdevice fallback if neither matches).
//   * atan2f/atanf, acosf : glibc sysdeps/ieee754/flt-32/{e_atan2f,s_atanf,e_acosf}.c (fdlibm float ports, no FMA).
// Every function is __host__ __device__ so tests/test_libm.py can compare the very same code with the host's libm
// (100 % agreement on tens of millions of inputs, see the test).  Compiled with -fmad=false: only the explicit fma() calls
// of the FMA variant fuse.
#pragma once
#include <math.h>
#include <stdint.h>

#include "c2g_exp_table.inc"

#ifdef __CUDACC__
#define C2G_LM __host__ __device__ __forceinline__
#else
#define C2G_LM inline
#endif

C2G_LM uint32_t c2g_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  union {
    float f;
    uint32_t u;
  } c;
  c.f = f;
  return c.u;
#endif
}
C2G_LM float c2g_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union {
    float f;
    uint32_t u;
  } c;
  c.u = u;
  return c.f;
#endif
}
C2G_LM uint64_t c2g_d2u(double d) {
#ifdef __CUDA_ARCH__
  return (uint64_t) __double_as_longlong(d);
#else
  union {
    double d;
    uint64_t u;
  } c;
  c.d = d;
  return c.u;
#endif
}
C2G_LM double c2g_u2d(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long) u);
#else
  union {
    double d;
    uint64_t u;
  } c;
  c.u = u;
  return c.d;
#endif
}

// ---- exp -----------------------------------------------------------------------------------------------------------
// tab: the 256-word __exp_data.tab (device: global/constant copy, host: static array)
template <bool FMA>
C2G_LM double c2g_exp_glibc(double x, const uint64_t *tab) {
  const uint32_t abstop = (uint32_t) (c2g_d2u(x) >> 52) & 0x7ffu;
  if (abstop - 0x3c9u >= 0x408u - 0x3c9u) {  // |x| < 2^-54 or |x| >= 512 (or NaN/Inf)
    if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + x;  // tiny x
    return exp(x);  // |x| >= 512: never reached by gaussPDF arguments; keep the library semantics
  }
  const double z = C2G_EXP_INVLN2N * x;
  double kd = z + C2G_EXP_SHIFT;
  const uint64_t ki = c2g_d2u(kd);
  kd -= C2G_EXP_SHIFT;
  double r;
  if (FMA)
    r = fma(kd, C2G_EXP_NEGLN2LON, fma(kd, C2G_EXP_NEGLN2HIN, x));
  else
    r = x + kd * C2G_EXP_NEGLN2HIN + kd * C2G_EXP_NEGLN2LON;
  const uint64_t idx = 2 * (ki % 128);
  const uint64_t top = ki << (52 - 7);
  const double tail = c2g_u2d(tab[idx]);
  const uint64_t sbits = tab[idx + 1] + top;
  const double r2 = r * r;
  double tmp;
  if (FMA)
    tmp = fma(r2 * r2, fma(r, C2G_EXP_C5, C2G_EXP_C4), fma(r2, fma(r, C2G_EXP_C3, C2G_EXP_C2), tail + r));
  else
    tmp = tail + r + r2 * (C2G_EXP_C2 + r * C2G_EXP_C3) + r2 * r2 * (C2G_EXP_C4 + r * C2G_EXP_C5);
  const double scale = c2g_u2d(sbits);
  return FMA ? fma(scale, tmp, scale) : scale + scale * tmp;
}

// Main path of c2g_exp_glibc without its range branches (straight-line code that the compiler can interleave over several
// independent arguments): *special is set when the argument needs the full function (|x| < 2^-54, |x| >= 512, NaN / Inf);
// the value returned in that case is meaningless but its computation is harmless.
template <bool FMA>
C2G_LM double c2g_exp_glibc_main(double x, const uint64_t *tab, bool *special) {
  const uint32_t abstop = (uint32_t) (c2g_d2u(x) >> 52) & 0x7ffu;
  if (abstop - 0x3c9u >= 0x408u - 0x3c9u) *special = true;
  const double z = C2G_EXP_INVLN2N * x;
  double kd = z + C2G_EXP_SHIFT;
  const uint64_t ki = c2g_d2u(kd);
  kd -= C2G_EXP_SHIFT;
  double r;
  if (FMA)
    r = fma(kd, C2G_EXP_NEGLN2LON, fma(kd, C2G_EXP_NEGLN2HIN, x));
  else
    r = x + kd * C2G_EXP_NEGLN2HIN + kd * C2G_EXP_NEGLN2LON;
  const uint64_t idx = 2 * (ki % 128);
  const uint64_t top = ki << (52 - 7);
  const double tail = c2g_u2d(tab[idx]);
  const uint64_t sbits = tab[idx + 1] + top;
  const double r2 = r * r;
  double tmp;
  if (FMA)
    tmp = fma(r2 * r2, fma(r, C2G_EXP_C5, C2G_EXP_C4), fma(r2, fma(r, C2G_EXP_C3, C2G_EXP_C2), tail + r));
  else
    tmp = tail + r + r2 * (C2G_EXP_C2 + r * C2G_EXP_C3) + r2 * r2 * (C2G_EXP_C4 + r * C2G_EXP_C5);
  const double scale = c2g_u2d(sbits);
  return FMA ? fma(scale, tmp, scale) : scale + scale * tmp;
}

// mode: 0 libdevice/libm exp, 1 glibc algorithm without FMA, 2 glibc algorithm with FMA (x86-64 __exp_fma)
C2G_LM double c2g_exp(double x, int mode, const uint64_t *tab) {
  if (mode == 2) return c2g_exp_glibc<true>(x, tab);
  if (mode == 1) return c2g_exp_glibc<false>(x, tab);
  return exp(x);
}

// ---- atanf / atan2f (glibc flt-32/s_atanf.c, e_atan2f.c) --------------------------------------------------------------
C2G_LM float c2g_atanf(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  const int32_t hx = (int32_t) c2g_f2u(x);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;
    return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x39800000) return x;  // |x| < 2^-12
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) {
        id = 0;
        x = (2.0f * x - 1.0f) / (2.0f + x);
      } else {
        id = 1;
        x = (x - 1.0f) / (x + 1.0f);
      }
    } else {
      if (ix < 0x401c0000) {
        id = 2;
        x = (x - 1.5f) / (1.0f + 1.5f * x);
      } else {
        id = 3;
        x = -1.0f / x;
      }
    }
  }
  const float z = x * x;
  const float w = z * z;
  const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) return x - x * (s1 + s2);
  const float zz = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return hx < 0 ? -zz : zz;
}

C2G_LM float c2g_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const int32_t hx = (int32_t) c2g_f2u(x), hy = (int32_t) c2g_f2u(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return c2g_atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    if (m == 0 || m == 1) return y;
    return m == 2 ? pi + tiny : -pi - tiny;
  }
  if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      if (m == 0) return pi_o_4 + tiny;
      if (m == 1) return -pi_o_4 - tiny;
      if (m == 2) return 3.0f * pi_o_4 + tiny;
      return -3.0f * pi_o_4 - tiny;
    }
    if (m == 0) return 0.0f;
    if (m == 1) return -0.0f;
    return m == 2 ? pi + tiny : -pi - tiny;
  }
  if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60)
    z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60)
    z = 0.0f;
  else
    z = c2g_atanf(fabsf(y / x));
  if (m == 0) return z;
  if (m == 1) return c2g_u2f(c2g_f2u(z) ^ 0x80000000u);
  if (m == 2) return pi - (z - pi_lo);
  return (z - pi_lo) - pi;
}

// ---- acosf (glibc flt-32/e_acosf.c) --------------------------------------------------------------------------------------
C2G_LM float c2g_acosf(float x) {
  const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
              pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
              pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
              qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
  const int32_t hx = (int32_t) c2g_f2u(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
  if (ix > 0x3f800000) return (x - x) / (x - x);
  if (ix < 0x3f000000) {  // |x| < 0.5
    if (ix <= 0x32800000) return pio2_hi + pio2_lo;
    const float z = x * x;
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = p / q;
    return pio2_hi - (x - (pio2_lo - x * r));
  }
  if (hx < 0) {  // x < -0.5
    const float z = (one + x) * 0.5f;
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float s = sqrtf(z);
    const float r = p / q;
    const float w = r * s - pio2_lo;
    return pi - 2.0f * (s + w);
  }
  const float z = (one - x) * 0.5f;  // x > 0.5
  const float s = sqrtf(z);
  const float df = c2g_u2f(c2g_f2u(s) & 0xfffff000u);
  const float c = (z - df * df) / (s + df);
  const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
  const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
  const float r = p / q;
  const float w = r * s + c;
  return 2.0f * (df + w);
}
