// pgb_mathf.h — bit-defined FP32 elementary functions and FP32 Box-Muller for the F32 precision mode
// (pgb_config.precision = PGB_PRECISION_F32), compiled identically for the host (gcc, -ffp-contract=off) and for sm_100a.
//
// Why not the MUFU approximations (__expf / __sinf / __logf): the resampling and ancestor indices are a discontinuous
// function of the weights, and the parity contract of this repository is "indices bit-exact against the oracle".  A
// hardware approximation cannot be reproduced on the CPU; a short polynomial built from IEEE-754 correctly rounded
// FP32 operations (+, -, *, /, sqrt, fma) in one fixed order can — and costs about as many instructions as the MUFU
// sequence with its range reduction (FP32 issues at the full rate, FP64 at half).  Accuracy (tests/test_mathf.py against
// float64 references): exp <= 1 ulp on [-87, 88], sin / cos <= 1 ulp of the FP32 result for |x| <= 2^15 (beyond: the FP64
// routine of pgb_math.h, rounded), log <= 1 ulp on (0, 1], i.e. the "stated FP32 tolerance" is FP32 round-off.
// Coefficients: Cephes single-precision (expf, sinf, cosf, logf), public domain.
#ifndef PGB_MATHF_H
#define PGB_MATHF_H

#include "pgb_math.h"

#if defined(__CUDA_ARCH__)
#define PGF_MUL(a, b) __fmul_rn((a), (b))
#define PGF_ADD(a, b) __fadd_rn((a), (b))
#define PGF_SUB(a, b) __fsub_rn((a), (b))
#define PGF_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define PGF_DIV(a, b) __fdiv_rn((a), (b))
#define PGF_SQRT(a) __fsqrt_rn((a))
#else
#define PGF_MUL(a, b) ((float)(a) * (float)(b))
#define PGF_ADD(a, b) ((float)(a) + (float)(b))
#define PGF_SUB(a, b) ((float)(a) - (float)(b))
#define PGF_FMA(a, b, c) fmaf((a), (b), (c))
#define PGF_DIV(a, b) ((float)(a) / (float)(b))
#define PGF_SQRT(a) sqrtf((a))
#endif

PGB_HD uint32_t pgf_f2u(float x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__float_as_int(x);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  return u;
#endif
}
PGB_HD float pgf_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __int_as_float((int)u);
#else
  float x;
  memcpy(&x, &u, 4);
  return x;
#endif
}

// exp(x).  x < -87 returns +0 (a DEFINED flush: e^-87 = 1.6e-38 is at the edge of the normal range; the weights this
// produces are relative to the step's maximum weight 1), x > 88.5 returns +Inf, NaN propagates.
PGB_HD float pgf_exp(float x) {
  if (!(x >= -87.0f)) return (x != x) ? x : 0.0f;
  if (x > 88.5f) return pgf_u2f(0x7f800000u);
  const float t = PGF_FMA(x, 1.44269504088896341f, 12582912.0f);  // 1.5 * 2^23: the low mantissa bits hold k
  const float kf = PGF_SUB(t, 12582912.0f);
  const int k = (int)(pgf_f2u(t) - 0x4b400000u);
  float r = PGF_FMA(-kf, 0.693359375f, x);         // ln2 split: 0x1.63p-1 (exact product for |k| < 2^15)
  r = PGF_FMA(-kf, -2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = PGF_FMA(p, r, 1.3981999507e-3f);
  p = PGF_FMA(p, r, 8.3334519073e-3f);
  p = PGF_FMA(p, r, 4.1665795894e-2f);
  p = PGF_FMA(p, r, 1.6666665459e-1f);
  p = PGF_FMA(p, r, 5.0000001201e-1f);
  const float e = PGF_ADD(PGF_FMA(PGF_MUL(r, r), p, r), 1.0f);
  // 2^k in two exact steps (k in [-126, 128]): both factors are normal powers of two
  const int k1 = k >> 1, k2 = k - k1;
  return PGF_MUL(PGF_MUL(e, pgf_u2f((uint32_t)(k1 + 127) << 23)), pgf_u2f((uint32_t)(k2 + 127) << 23));
}

// exp(x) of an FP32 argument returned as an FP64 value: the FP32 polynomial of pgf_exp, scaled by 2^k in FP64 (exact), so
// the result does not underflow where FP32 would (down to e^-700).  Used for the transition density of the ancestor
// weights (particle_Gibbs.jl:178-179): those are NOT relative to a maximum, and exp(c0 - |r|^2/2) of every particle can
// lie below FLT_MIN at a step where the model fits the reference trajectory poorly — the FP64 normalisation that follows
// then still sees their ratios.  x < -700 returns +0 (defined flush), x > 700 +Inf, NaN propagates.
PGB_HD double pgf_exp_wide(float x) {
  if (!(x >= -700.0f)) return (x != x) ? (double)x : 0.0;
  if (x > 700.0f) return (double)pgf_u2f(0x7f800000u);
  const float t = PGF_FMA(x, 1.44269504088896341f, 12582912.0f);
  const float kf = PGF_SUB(t, 12582912.0f);
  const int k = (int)(pgf_f2u(t) - 0x4b400000u);  // |k| <= 1010
  float r = PGF_FMA(-kf, 0.693359375f, x);
  r = PGF_FMA(-kf, -2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = PGF_FMA(p, r, 1.3981999507e-3f);
  p = PGF_FMA(p, r, 8.3334519073e-3f);
  p = PGF_FMA(p, r, 4.1665795894e-2f);
  p = PGF_FMA(p, r, 1.6666665459e-1f);
  p = PGF_FMA(p, r, 5.0000001201e-1f);
  const float e = PGF_ADD(PGF_FMA(PGF_MUL(r, r), p, r), 1.0f);
  return PGB_MUL((double)e, pgb_u2d((uint64_t)(k + 1023) << 52));  // exact: e has 24 bits, 2^k is a normal double
}

// r = x - k*pi/2 (three-constant Cody-Waite, exact products for k < 2^15), returns k mod 4
PGB_HD int pgf_rem_pio2(float x, float* r) {
  const float t = PGF_FMA(x, 0.636619772367581343f, 12582912.0f);
  const float kf = PGF_SUB(t, 12582912.0f);
  float r1 = PGF_FMA(-kf, 1.5703125f, x);
  r1 = PGF_FMA(-kf, 4.837512969970703125e-4f, r1);
  *r = PGF_FMA(-kf, 7.54978995489188216e-8f, r1);
  return (int)(pgf_f2u(t) - 0x4b400000u) & 3;
}
PGB_HD float pgf_ksin(float r, float z) {
  float p = -1.9515295891e-4f;
  p = PGF_FMA(p, z, 8.3321608736e-3f);
  p = PGF_FMA(p, z, -1.6666654611e-1f);
  return PGF_FMA(PGF_MUL(r, z), p, r);
}
PGB_HD float pgf_kcos(float z) {
  float p = 2.443315711809948e-5f;
  p = PGF_FMA(p, z, -1.388731625493765e-3f);
  p = PGF_FMA(p, z, 4.166664568298827e-2f);
  return PGF_ADD(PGF_FMA(-0.5f, z, 1.0f), PGF_MUL(PGF_MUL(z, z), p));
}
PGB_HD void pgf_sincos(float x, float* s, float* c) {
  if (!(fabsf(x) <= 32768.0f)) {  // rare: beyond the exact range of the three-constant reduction (also NaN / Inf)
    double sd, cd;
    pgb_sincos((double)x, &sd, &cd);
    *s = (float)sd;
    *c = (float)cd;
    return;
  }
  float r;
  const int q = pgf_rem_pio2(x, &r);
  const float z = PGF_MUL(r, r);
  const float ks = pgf_ksin(r, z), kc = pgf_kcos(z);
  const float s0 = (q & 1) ? kc : ks;
  const float c0 = (q & 1) ? ks : kc;
  *s = (q & 2) ? -s0 : s0;
  *c = ((q + 1) & 2) ? -c0 : c0;
}
PGB_HD float pgf_sin(float x) {
  if (!(fabsf(x) <= 32768.0f)) return (float)pgb_sin((double)x);
  float r;
  const int q = pgf_rem_pio2(x, &r);
  const float z = PGF_MUL(r, r);
  const float v = (q & 1) ? pgf_kcos(z) : pgf_ksin(r, z);
  return (q & 2) ? -v : v;
}
PGB_HD float pgf_cos(float x) {
  if (!(fabsf(x) <= 32768.0f)) return (float)pgb_cos((double)x);
  float r;
  const int q = pgf_rem_pio2(x, &r);
  const float z = PGF_MUL(r, r);
  const float v = (q & 1) ? pgf_ksin(r, z) : pgf_kcos(z);
  return ((q + 1) & 2) ? -v : v;
}

// log(x) for a finite, positive, NORMAL x (the Box-Muller uniforms are >= 2^-24)
PGB_HD float pgf_log_pos_normal(float x) {
  const uint32_t u = pgf_f2u(x);
  int e = (int)(u >> 23) - 126;                       // x = m * 2^e, m in [0.5, 1)
  float m = pgf_u2f((u & 0x007fffffu) | 0x3f000000u);
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = PGF_SUB(PGF_ADD(m, m), 1.0f);
  } else {
    m = PGF_SUB(m, 1.0f);
  }
  const float z = PGF_MUL(m, m);
  float p = 7.0376836292e-2f;
  p = PGF_FMA(p, m, -1.1514610310e-1f);
  p = PGF_FMA(p, m, 1.1676998740e-1f);
  p = PGF_FMA(p, m, -1.2420140846e-1f);
  p = PGF_FMA(p, m, 1.4249322787e-1f);
  p = PGF_FMA(p, m, -1.6668057665e-1f);
  p = PGF_FMA(p, m, 2.0000714765e-1f);
  p = PGF_FMA(p, m, -2.4999993993e-1f);
  p = PGF_FMA(p, m, 3.3333331174e-1f);
  const float ef = (float)e;
  float y = PGF_MUL(PGF_MUL(m, z), p);
  y = PGF_FMA(ef, -2.12194440e-4f, y);
  y = PGF_FMA(-0.5f, z, y);
  return PGF_FMA(ef, 0.693359375f, PGF_ADD(m, y));
}

// uniform in (0, 1) on a 2^-23 grid, never 0 or 1: ((w >> 9) + 0.5) * 2^-23 (exact in FP32)
PGB_HD float pgf_u23(uint32_t w) { return PGF_MUL(PGF_ADD((float)(w >> 9), 0.5f), 1.1920928955078125e-7f); }

// Box-Muller pair from two 32-bit words
PGB_HD void pgf_normal_pair(uint32_t w0, uint32_t w1, float* z0, float* z1) {
  const float u1 = pgf_u23(w0), u2 = pgf_u23(w1);
  const float rad = PGF_SQRT(PGF_MUL(-2.0f, pgf_log_pos_normal(u1)));
  float s, c;
  pgf_sincos(PGF_MUL(6.28318530717958648f, u2), &s, &c);
  *z0 = PGF_MUL(rad, c);
  *z1 = PGF_MUL(rad, s);
}

// The F32 mode's standard normals of particle j (all n_x <= 4 dims) from ONE Philox block: words (0, 1) -> dims (0, 1),
// words (2, 3) -> dims (2, 3).  Counter = (j, t, sweep, purpose << 24 | chain) as in pgb_stream_block.
PGB_HD void pgf_stream_normals(uint64_t seed, uint32_t purpose, uint32_t chain, uint32_t sweep, uint32_t t, uint32_t j,
                               int n_x, float* z) {
  const pgb_u32x4 b = pgb_stream_block(seed, purpose, chain, sweep, t, j);
  float a0, a1;
  pgf_normal_pair(b.v[0], b.v[1], &a0, &a1);
  z[0] = a0;
  if (n_x > 1) z[1] = a1;
  if (n_x > 2) {
    pgf_normal_pair(b.v[2], b.v[3], &a0, &a1);
    z[2] = a0;
    if (n_x > 3) z[3] = a1;
  }
}

#endif  // PGB_MATHF_H
