// pgb_lmath.cuh — lean one-element device versions of the pgb_math.h routines for the chunk-stationary sweep kernel
// (pgb_v2.cuh).  Same arithmetic, same order, same bits as the scalar definitions: the common case runs a short
// straight-line sequence, everything unusual (NaN, overflow / underflow range, huge arguments) branches to the
// scalar __noinline__ routine.  Latency is hidden by the 28 resident warps of that kernel, not by interleaving
// several elements per thread, so the hot code stays small enough for the 32 KB instruction cache.
// Bit identity with the host build is checked on the GPU by tests/test_gpu_parity.py::test_lean_math_* .
#pragma once
#include "pgb_math.h"

namespace pgb {

// exp: the |x| <= 700 range covers every weight / density exponent that does not underflow or overflow
__device__ __forceinline__ double lexp(double x) {
  if (!(fabs(x) <= 700.0)) return pgb_exp(x);
  const double t = __fma_rn(x, 0x1.71547652b82fep+0, 0x1.8p52);
  const double kd = __dsub_rn(t, 0x1.8p52);
  const int k = __double2loint(t);  // t = 1.5*2^52 + k exactly: the low word of its mantissa is k (two's complement)
  double r = __fma_rn(-kd, 0x1.62e42fee00000p-1, x);
  r = __fma_rn(-kd, 0x1.a39ef35793c76p-33, r);
  double p = 0x1.93974a8c07c9dp-37;
  p = __fma_rn(p, r, 0x1.6124613a86d09p-33);
  p = __fma_rn(p, r, 0x1.1eed8eff8d898p-29);
  p = __fma_rn(p, r, 0x1.ae64567f544e4p-26);
  p = __fma_rn(p, r, 0x1.27e4fb7789f5cp-22);
  p = __fma_rn(p, r, 0x1.71de3a556c734p-19);
  p = __fma_rn(p, r, 0x1.a01a01a01a01ap-16);
  p = __fma_rn(p, r, 0x1.a01a01a01a01ap-13);
  p = __fma_rn(p, r, 0x1.6c16c16c16c17p-10);
  p = __fma_rn(p, r, 0x1.1111111111111p-7);
  p = __fma_rn(p, r, 0x1.5555555555555p-5);
  p = __fma_rn(p, r, 0x1.5555555555555p-3);
  p = __fma_rn(p, r, 0.5);
  const double r2 = __dmul_rn(r, r);
  const double e = __dadd_rn(__fma_rn(r2, p, r), 1.0);
  return __dmul_rn(e, __hiloint2double((k + 1023) << 20, 0));  // |k| <= 1010: one exact-or-single-rounding scaling
}

// r = x - k*pi/2 (three-constant Cody-Waite), q = k mod 4; |x| < 1e15 assumed by the callers below
__device__ __forceinline__ int lrem_pio2(double x, double& r) {
  const double t = __fma_rn(x, 0x1.45f306dc9c883p-1, 0x1.8p52);
  const double kd = __dsub_rn(t, 0x1.8p52);
  double r1 = __fma_rn(-kd, 0x1.921fb54400000p+0, x);
  r1 = __fma_rn(-kd, 0x1.0b4611a600000p-34, r1);
  r = __fma_rn(-kd, 0x1.3198a2e037073p-69, r1);
  return __double2loint(t) & 3;
}
__device__ __forceinline__ double lksin(double r, double z) {
  double p = 1.58969099521155010221e-10;
  p = __fma_rn(p, z, -2.50507602534068634195e-08);
  p = __fma_rn(p, z, 2.75573137070700676789e-06);
  p = __fma_rn(p, z, -1.98412698298579493134e-04);
  p = __fma_rn(p, z, 8.33333333332248946124e-03);
  p = __fma_rn(p, z, -1.66666666666666324348e-01);
  return __fma_rn(__dmul_rn(r, z), p, r);
}
__device__ __forceinline__ double lkcos(double z) {
  double p = -1.13596475577881948265e-11;
  p = __fma_rn(p, z, 2.08757232129817482790e-09);
  p = __fma_rn(p, z, -2.75573143513906633035e-07);
  p = __fma_rn(p, z, 2.48015872894767294178e-05);
  p = __fma_rn(p, z, -1.38888888888741095749e-03);
  p = __fma_rn(p, z, 4.16666666666666019037e-02);
  const double hz = __dmul_rn(0.5, z);
  const double w = __dsub_rn(1.0, hz);
  const double tail = __dadd_rn(__dsub_rn(__dsub_rn(1.0, w), hz), __dmul_rn(__dmul_rn(z, z), p));
  return __dadd_rn(w, tail);
}
__device__ __forceinline__ void lsincos(double x, double& s, double& c) {
  if (!(fabs(x) < 1e15)) {
    pgb_sincos(x, &s, &c);
    return;
  }
  double r;
  const int q = lrem_pio2(x, r);
  const double z = __dmul_rn(r, r);
  const double ks = lksin(r, z), kc = lkcos(z);
  double s0 = (q & 1) ? kc : ks;
  double c0 = (q & 1) ? ks : kc;
  s = (q & 2) ? -s0 : s0;
  c = ((q + 1) & 2) ? -c0 : c0;
}
// sin only / cos only: the quadrant picks ONE of the two kernels (warp-divergent only when lanes sit in different
// quadrants; both sides are short)
__device__ __forceinline__ double lsin(double x) {
  if (!(fabs(x) < 1e15)) return pgb_sin(x);
  double r;
  const int q = lrem_pio2(x, r);
  const double z = __dmul_rn(r, r);
  const double v = (q & 1) ? lkcos(z) : lksin(r, z);
  return (q & 2) ? -v : v;
}
__device__ __forceinline__ double lcos(double x) {
  if (!(fabs(x) < 1e15)) return pgb_cos(x);
  double r;
  const int q = lrem_pio2(x, r);
  const double z = __dmul_rn(r, r);
  const double v = (q & 1) ? lksin(r, z) : lkcos(z);
  return ((q + 1) & 2) ? -v : v;
}

// log of a finite, positive, NORMAL argument (the Box-Muller uniforms are >= 2^-54)
__device__ __forceinline__ double llog_pos_normal(double x) {
  const int hi = __double2hiint(x);
  int e = (hi >> 20) - 1023;
  double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  if (m > 0x1.6a09e667f3bcdp+0) {
    m = __dmul_rn(m, 0.5);
    e += 1;
  }
  const double dk = (double)e;
  const double f = __dsub_rn(m, 1.0);
  const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
  const double z = __dmul_rn(s, s);
  const double w = __dmul_rn(z, z);
  const double t1 = __dmul_rn(w, __fma_rn(w, __fma_rn(w, 1.531383769920937332e-01, 2.222219843214978396e-01),
                                          3.999999999940941908e-01));
  const double t2 = __dmul_rn(z, __fma_rn(w, __fma_rn(w, __fma_rn(w, 1.479819860511658591e-01, 1.818357216161805012e-01),
                                                      2.857142874366239149e-01),
                                          6.666666666666735130e-01));
  const double R = __dadd_rn(t2, t1);
  const double hfsq = __dmul_rn(0.5, __dmul_rn(f, f));
  const double inner = __fma_rn(s, __dadd_rn(hfsq, R), __dmul_rn(dk, 0x1.a39ef35793c76p-33));
  return __dsub_rn(__dmul_rn(dk, 0x1.62e42fee00000p-1), __dsub_rn(__dsub_rn(hfsq, inner), f));
}

// Philox4x32-10, one block (same as pgb_philox4x32_10)
__device__ __forceinline__ void lphilox(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// Box-Muller pair from one Philox block (same as pgb_normal_pair)
__device__ __forceinline__ void lnormal_pair(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, double& z0, double& z1) {
  const double u1 = pgb_u53(b0, b1);
  const double u2 = pgb_u53(b2, b3);
  const double rad = __dsqrt_rn(__dmul_rn(-2.0, llog_pos_normal(u1)));
  double s, c;
  lsincos(__dmul_rn(0x1.921fb54442d18p+2, u2), s, c);
  z0 = __dmul_rn(rad, c);
  z1 = __dmul_rn(rad, s);
}

}  // namespace pgb
