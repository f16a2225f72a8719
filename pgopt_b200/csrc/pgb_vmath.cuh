// pgb_vmath.cuh — K-wide, branch-free device versions of the pgb_math.h routines.
//
// Same arithmetic, same order, same results bit for bit as the scalar routines (checked on the GPU against the
// host build over random and edge inputs, tests/test_gpu_parity.py::test_vector_math_*), but written
// "structure of arrays": every elementary step is applied to all K independent lanes before the next step, and
// special cases are resolved with selects at the end instead of early returns.  That gives the scheduler K
// independent dependency chains per thread (the scalar Horner chains are strictly serial) and removes the
// divergence bookkeeping (BSSY/BSYNC/BRA) from the hot loops.
#pragma once
#include "pgb_math.h"

namespace pgb {

// PGB_SMALL_CODE=1 makes the K-wide entry points forward to the scalar __noinline__ routines of pgb_math.h (one
// call per lane, compact code).  Measured on B200 (scripts/ab_variants.py, N = 2^20, known basis): interleaved +
// fully inlined 195 us/step, compact 225 us/step, compact + non-inlined helpers 277 us/step — the kernel is bound by
// per-warp dependency latency, not by instruction fetch, so the interleaved versions are the default.
#ifndef PGB_SMALL_CODE
#define PGB_SMALL_CODE 0
#endif

template <int K>
__device__ __forceinline__ void vexp_wide(const double (&x)[K], double (&y)[K]) {
  double xc[K], kd[K], r[K], p[K];
  int ki[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // clamp keeps the integer part in range; results outside the clamp are fixed up below
    xc[k] = fmin(fmax(x[k], -746.0), 710.0);
    double t = __fma_rn(xc[k], 0x1.71547652b82fep+0, 0x1.8p52);
    kd[k] = __dsub_rn(t, 0x1.8p52);
    ki[k] = (int)kd[k];
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    r[k] = __fma_rn(-kd[k], 0x1.62e42fee00000p-1, xc[k]);
    r[k] = __fma_rn(-kd[k], 0x1.a39ef35793c76p-33, r[k]);
    p[k] = 0x1.93974a8c07c9dp-37;
  }
  const double C[12] = {0x1.6124613a86d09p-33, 0x1.1eed8eff8d898p-29, 0x1.ae64567f544e4p-26, 0x1.27e4fb7789f5cp-22,
                        0x1.71de3a556c734p-19, 0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-13, 0x1.6c16c16c16c17p-10,
                        0x1.1111111111111p-7,  0x1.5555555555555p-5,  0x1.5555555555555p-3,  0.5};
#pragma unroll
  for (int j = 0; j < 12; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = __fma_rn(p[k], r[k], C[j]);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double r2 = __dmul_rn(r[k], r[k]);
    double e = __dadd_rn(__fma_rn(r2, p[k], r[k]), 1.0);
    // e * 2^ki in two exact-or-single-rounding steps (identical to the scalar routine's case split)
    int a = ki[k] >> 1, b = ki[k] - a;
    double v = __dmul_rn(__dmul_rn(e, pgb_pow2i(a)), pgb_pow2i(b));
    v = (x[k] > 709.782712893384) ? __longlong_as_double(0x7ff0000000000000ll) : v;
    v = (x[k] < -746.0) ? 0.0 : v;
    y[k] = (x[k] != x[k]) ? x[k] : v;
  }
}

// log for finite, positive, NORMAL arguments (the Box-Muller uniforms are >= 2^-54)
template <int K>
__device__ __forceinline__ void vlog_pos_normal_wide(const double (&x)[K], double (&y)[K]) {
  double f[K], s[K], dk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    uint64_t ux = (uint64_t)__double_as_longlong(x[k]);
    int e = (int)(ux >> 52) - 1023;
    double m = __longlong_as_double((long long)((ux & 0x000fffffffffffffull) | 0x3ff0000000000000ull));
    const bool big = m > 0x1.6a09e667f3bcdp+0;
    m = big ? __dmul_rn(m, 0.5) : m;
    e += big ? 1 : 0;
    dk[k] = (double)e;
    f[k] = __dsub_rn(m, 1.0);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) s[k] = __ddiv_rn(f[k], __dadd_rn(2.0, f[k]));
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double z = __dmul_rn(s[k], s[k]);
    double w = __dmul_rn(z, z);
    double t1 = __dmul_rn(w, __fma_rn(w, __fma_rn(w, 1.531383769920937332e-01, 2.222219843214978396e-01),
                                      3.999999999940941908e-01));
    double t2 = __dmul_rn(z, __fma_rn(w, __fma_rn(w, __fma_rn(w, 1.479819860511658591e-01, 1.818357216161805012e-01),
                                                  2.857142874366239149e-01),
                                      6.666666666666735130e-01));
    double R = __dadd_rn(t2, t1);
    double hfsq = __dmul_rn(0.5, __dmul_rn(f[k], f[k]));
    double inner = __fma_rn(s[k], __dadd_rn(hfsq, R), __dmul_rn(dk[k], 0x1.a39ef35793c76p-33));
    y[k] = __dsub_rn(__dmul_rn(dk[k], 0x1.62e42fee00000p-1), __dsub_rn(__dsub_rn(hfsq, inner), f[k]));
  }
}

// sin and cos of K arguments (|x| < 2^50 handled exactly like the scalar routine; beyond 1e15 -> NaN)
template <int K>
__device__ __forceinline__ void vsincos_wide(const double (&x)[K], double (&sn)[K], double (&cs)[K]) {
  double r[K];
  int q[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = __fma_rn(x[k], 0x1.45f306dc9c883p-1, 0x1.8p52);
    double kd = __dsub_rn(t, 0x1.8p52);
    q[k] = (int)((long long)kd & 3);
    double r1 = __fma_rn(-kd, 0x1.921fb54400000p+0, x[k]);
    r1 = __fma_rn(-kd, 0x1.0b4611a600000p-34, r1);
    r[k] = __fma_rn(-kd, 0x1.3198a2e037073p-69, r1);
  }
  double ps[K], pc[K], z[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    z[k] = __dmul_rn(r[k], r[k]);
    ps[k] = 1.58969099521155010221e-10;
    pc[k] = -1.13596475577881948265e-11;
  }
  const double S[5] = {-2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
                       8.33333333332248946124e-03, -1.66666666666666324348e-01};
  const double Cc[5] = {2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
                        -1.38888888888741095749e-03, 4.16666666666666019037e-02};
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      ps[k] = __fma_rn(ps[k], z[k], S[j]);
      pc[k] = __fma_rn(pc[k], z[k], Cc[j]);
    }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double ks = __fma_rn(__dmul_rn(r[k], z[k]), ps[k], r[k]);
    double hz = __dmul_rn(0.5, z[k]);
    double w = __dsub_rn(1.0, hz);
    double tail = __dadd_rn(__dsub_rn(__dsub_rn(1.0, w), hz), __dmul_rn(__dmul_rn(z[k], z[k]), pc[k]));
    double kc = __dadd_rn(w, tail);
    double s0 = (q[k] & 1) ? kc : ks;
    double c0 = (q[k] & 1) ? ks : kc;
    s0 = (q[k] & 2) ? -s0 : s0;
    c0 = ((q[k] + 1) & 2) ? -c0 : c0;
    const bool ok = fabs(x[k]) < 1e15;
    double bad = __dsub_rn(x[k], x[k]);
    sn[k] = ok ? s0 : bad;
    cs[k] = ok ? c0 : bad;
  }
}

// K Philox4x32-10 blocks with a common key
template <int K>
__device__ __forceinline__ void vphilox_wide(pgb_u32x4 (&ctr)[K], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      uint32_t hi0 = __umulhi(0xD2511F53u, ctr[k].v[0]), lo0 = 0xD2511F53u * ctr[k].v[0];
      uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[k].v[2]), lo1 = 0xCD9E8D57u * ctr[k].v[2];
      pgb_u32x4 n;
      n.v[0] = hi1 ^ ctr[k].v[1] ^ k0;
      n.v[1] = lo1;
      n.v[2] = hi0 ^ ctr[k].v[3] ^ k1;
      n.v[3] = lo0;
      ctr[k] = n;
    }
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// K Box-Muller pairs from K Philox blocks (same as pgb_normal_pair)
template <int K>
__device__ __forceinline__ void vnormal_pairs_wide(const pgb_u32x4 (&bits)[K], double (&z0)[K], double (&z1)[K]) {
  double u1[K], ang[K], lg[K], s[K], c[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    u1[k] = pgb_u53(bits[k].v[0], bits[k].v[1]);
    ang[k] = __dmul_rn(0x1.921fb54442d18p+2, pgb_u53(bits[k].v[2], bits[k].v[3]));
  }
  vlog_pos_normal_wide<K>(u1, lg);
  vsincos_wide<K>(ang, s, c);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double rad = __dsqrt_rn(__dmul_rn(-2.0, lg[k]));
    z0[k] = __dmul_rn(rad, c[k]);
    z1[k] = __dmul_rn(rad, s[k]);
  }
}


template <int K>
__device__ __forceinline__ void vexp(const double (&x)[K], double (&y)[K]) {
#if PGB_SMALL_CODE
#pragma unroll
  for (int k = 0; k < K; ++k) y[k] = pgb_exp(x[k]);
#else
  vexp_wide<K>(x, y);
#endif
}
template <int K>
__device__ __forceinline__ void vsincos(const double (&x)[K], double (&sn)[K], double (&cs)[K]) {
#if PGB_SMALL_CODE
#pragma unroll
  for (int k = 0; k < K; ++k) pgb_sincos(x[k], &sn[k], &cs[k]);
#else
  vsincos_wide<K>(x, sn, cs);
#endif
}
template <int K>
__device__ __forceinline__ void vlog_pos_normal(const double (&x)[K], double (&y)[K]) {
#if PGB_SMALL_CODE
#pragma unroll
  for (int k = 0; k < K; ++k) y[k] = pgb_log(x[k]);
#else
  vlog_pos_normal_wide<K>(x, y);
#endif
}
template <int K>
__device__ __forceinline__ void vphilox(pgb_u32x4 (&ctr)[K], uint32_t k0, uint32_t k1) {
#if PGB_SMALL_CODE
#pragma unroll
  for (int k = 0; k < K; ++k) ctr[k] = pgb_philox4x32_10(ctr[k], k0, k1);
#else
  vphilox_wide<K>(ctr, k0, k1);
#endif
}
static __device__ __noinline__ void normal_pair_noinline(pgb_u32x4 bits, double* z0, double* z1) { pgb_normal_pair(bits, z0, z1); }
template <int K>
__device__ __forceinline__ void vnormal_pairs(const pgb_u32x4 (&bits)[K], double (&z0)[K], double (&z1)[K]) {
#if PGB_SMALL_CODE
#pragma unroll
  for (int k = 0; k < K; ++k) normal_pair_noinline(bits[k], &z0[k], &z1[k]);
#else
  vnormal_pairs_wide<K>(bits, z0, z1);
#endif
}

}  // namespace pgb
