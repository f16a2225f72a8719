// pgb_math.h — bit-defined FP64 elementary functions and counter-based RNG,
// compiled identically for the host (gcc, -ffp-contract=off) and for sm_100a (nvcc).
//
// Why this exists: resampling / ancestor indices are a discontinuous function of the
// particle weights, so GPU-vs-CPU index parity needs *bit-identical* exp/sin/cos/log on
// both sides.  glibc, libdevice and Julia's Base.Math all differ in the last bit on a
// few per cent of inputs, so neither side may call its platform libm.  Everything here
// is built from IEEE-754 correctly-rounded primitives only (+, -, *, /, sqrt, fma) in
// one fixed evaluation order; contraction is switched off on the host and every
// product/sum that must NOT fuse goes through PGB_MUL/PGB_ADD (round-to-nearest
// intrinsics on the device), so the result does not depend on compiler flags there.
//
// Accuracy (checked in tests/test_math.py against mpmath): exp, log, sin, cos < 1.5 ulp
// on the domains the particle filter uses.  The reference (Julia Base: exp/sin/cos,
// Julia/src/particle_Gibbs.jl:198, examples/PG_OCP_known_basis_functions_Altro.jl:34)
// is itself only "< 1 ulp, not correctly rounded", so agreement with real Julia is at
// the few-ulp level either way (SURVEY.md §7 hard part 2).
#ifndef PGB_MATH_H
#define PGB_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define PGB_HD static __host__ __device__ __forceinline__
// larger routines: one copy per kernel instead of one per call site (code size / instruction cache)
#define PGB_HD_NOINLINE static __host__ __device__ __noinline__
// bulky helpers of the exact-scan algebra / phase functions: inline or not is a tuning knob (code size vs calls)
#ifndef PGB_NOINLINE_HELPERS
#define PGB_NOINLINE_HELPERS 0
#endif
#if PGB_NOINLINE_HELPERS
#define PGB_HD_HELPER static __host__ __device__ __noinline__
#define PGB_DEV_HELPER __device__ __noinline__
#else
#define PGB_HD_HELPER static __host__ __device__ __forceinline__
#define PGB_DEV_HELPER __device__ __forceinline__
#endif
#else
#define PGB_HD_HELPER static inline
#define PGB_HD static inline
#define PGB_HD_NOINLINE static inline
#endif

#if defined(__CUDA_ARCH__)
#define PGB_MUL(a, b) __dmul_rn((a), (b))
#define PGB_ADD(a, b) __dadd_rn((a), (b))
#define PGB_SUB(a, b) __dsub_rn((a), (b))
#define PGB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PGB_DIV(a, b) __ddiv_rn((a), (b))
#define PGB_SQRT(a) __dsqrt_rn((a))
#else
// host: the translation unit MUST be compiled with -ffp-contract=off (oracle/Makefile,
// __graft_entry__.build); pgb_math_selftest() detects a violation at run time.
#define PGB_MUL(a, b) ((a) * (b))
#define PGB_ADD(a, b) ((a) + (b))
#define PGB_SUB(a, b) ((a) - (b))
#define PGB_FMA(a, b, c) fma((a), (b), (c))
#define PGB_DIV(a, b) ((a) / (b))
#define PGB_SQRT(a) sqrt((a))
#endif

PGB_HD uint64_t pgb_d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
PGB_HD double pgb_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}

// 2^k for -1022 <= k <= 1023
PGB_HD double pgb_pow2i(int k) { return pgb_u2d((uint64_t)(k + 1023) << 52); }

// ---------------------------------------------------------------------------------
// exp: k = rint(x*log2(e)); r = x - k*ln2 (two-constant Cody-Waite, exact first step);
// degree-14 Taylor/Horner on |r| <= 0.3466 (remainder 4e-18), scaled by 2^k.
// ---------------------------------------------------------------------------------
PGB_HD_NOINLINE double pgb_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return pgb_u2d(0x7ff0000000000000ull);
  if (x < -746.0) return 0.0;
  const double SHIFT = 0x1.8p52;
  double t = PGB_FMA(x, 0x1.71547652b82fep+0, SHIFT);
  double kd = PGB_SUB(t, SHIFT);
  int k = (int)kd;
  double r = PGB_FMA(-kd, 0x1.62e42fee00000p-1, x);
  r = PGB_FMA(-kd, 0x1.a39ef35793c76p-33, r);
  double p = 0x1.93974a8c07c9dp-37;            // 1/14!
  p = PGB_FMA(p, r, 0x1.6124613a86d09p-33);    // 1/13!
  p = PGB_FMA(p, r, 0x1.1eed8eff8d898p-29);    // 1/12!
  p = PGB_FMA(p, r, 0x1.ae64567f544e4p-26);    // 1/11!
  p = PGB_FMA(p, r, 0x1.27e4fb7789f5cp-22);    // 1/10!
  p = PGB_FMA(p, r, 0x1.71de3a556c734p-19);    // 1/9!
  p = PGB_FMA(p, r, 0x1.a01a01a01a01ap-16);    // 1/8!
  p = PGB_FMA(p, r, 0x1.a01a01a01a01ap-13);    // 1/7!
  p = PGB_FMA(p, r, 0x1.6c16c16c16c17p-10);    // 1/6!
  p = PGB_FMA(p, r, 0x1.1111111111111p-7);     // 1/5!
  p = PGB_FMA(p, r, 0x1.5555555555555p-5);     // 1/4!
  p = PGB_FMA(p, r, 0x1.5555555555555p-3);     // 1/3!
  p = PGB_FMA(p, r, 0.5);                      // 1/2!
  double r2 = PGB_MUL(r, r);
  double e = PGB_ADD(PGB_FMA(r2, p, r), 1.0);
  if (k >= -1021 && k <= 1023) return PGB_MUL(e, pgb_pow2i(k));
  if (k < -1021) return PGB_MUL(PGB_MUL(e, pgb_pow2i(k + 1000)), 0x1p-1000);
  return PGB_MUL(PGB_MUL(e, pgb_pow2i(k - 1)), 2.0);
}

// ---------------------------------------------------------------------------------
// log (x > 0): x = 2^k * m, m in [sqrt(1/2), sqrt(2)); f = m-1; s = f/(2+f);
// log(m) = f - (f^2/2 - s*(f^2/2 + R(s^2)))   (the classic 7-term even polynomial).
// ---------------------------------------------------------------------------------
PGB_HD_NOINLINE double pgb_log(double x) {
  if (x != x || x < 0.0) return pgb_u2d(0x7ff8000000000000ull);
  if (x == 0.0) return pgb_u2d(0xfff0000000000000ull);
  uint64_t ux = pgb_d2u(x);
  if (ux >= 0x7ff0000000000000ull) return x;
  int k = 0;
  if (ux < 0x0010000000000000ull) {  // subnormal
    x = PGB_MUL(x, 0x1p54);
    ux = pgb_d2u(x);
    k = -54;
  }
  k += (int)(ux >> 52) - 1023;
  uint64_t frac = ux & 0x000fffffffffffffull;
  double m = pgb_u2d(frac | 0x3ff0000000000000ull);  // [1,2)
  if (m > 0x1.6a09e667f3bcdp+0) {                      // > sqrt(2)
    m = PGB_MUL(m, 0.5);
    k += 1;
  }
  double f = PGB_SUB(m, 1.0);
  double s = PGB_DIV(f, PGB_ADD(2.0, f));
  double z = PGB_MUL(s, s);
  double w = PGB_MUL(z, z);
  double t1 = PGB_MUL(w, PGB_FMA(w, PGB_FMA(w, 1.531383769920937332e-01, 2.222219843214978396e-01),
                                 3.999999999940941908e-01));
  double t2 = PGB_MUL(
      z, PGB_FMA(w, PGB_FMA(w, PGB_FMA(w, 1.479819860511658591e-01, 1.818357216161805012e-01),
                            2.857142874366239149e-01),
                 6.666666666666735130e-01));
  double R = PGB_ADD(t2, t1);
  double hfsq = PGB_MUL(0.5, PGB_MUL(f, f));
  double dk = (double)k;
  // k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f)
  double inner = PGB_FMA(s, PGB_ADD(hfsq, R), PGB_MUL(dk, 0x1.a39ef35793c76p-33));
  return PGB_SUB(PGB_MUL(dk, 0x1.62e42fee00000p-1), PGB_SUB(PGB_SUB(hfsq, inner), f));
}

// ---------------------------------------------------------------------------------
// sin / cos: three-constant Cody-Waite reduction by pi/2 (33+33+53 bits; first step is
// exact for |x| < 2^20*pi/2), then the classic degree-13/14 odd/even kernels on
// |r| <= pi/4.  Deterministic for all finite x with |x| < 1e15; accurate (< 1.5 ulp)
// for |x| <~ 1e6, which covers cos(3*x1), sin(2*x2) and the Hilbert-space basis
// arguments pi*j*(z+L)/(2L).
// ---------------------------------------------------------------------------------
PGB_HD double pgb_ksin(double r) {
  double z = PGB_MUL(r, r);
  double p = 1.58969099521155010221e-10;
  p = PGB_FMA(p, z, -2.50507602534068634195e-08);
  p = PGB_FMA(p, z, 2.75573137070700676789e-06);
  p = PGB_FMA(p, z, -1.98412698298579493134e-04);
  p = PGB_FMA(p, z, 8.33333333332248946124e-03);
  p = PGB_FMA(p, z, -1.66666666666666324348e-01);
  return PGB_FMA(PGB_MUL(r, z), p, r);
}
PGB_HD double pgb_kcos(double r) {
  double z = PGB_MUL(r, r);
  double p = -1.13596475577881948265e-11;
  p = PGB_FMA(p, z, 2.08757232129817482790e-09);
  p = PGB_FMA(p, z, -2.75573143513906633035e-07);
  p = PGB_FMA(p, z, 2.48015872894767294178e-05);
  p = PGB_FMA(p, z, -1.38888888888741095749e-03);
  p = PGB_FMA(p, z, 4.16666666666666019037e-02);
  double hz = PGB_MUL(0.5, z);
  double w = PGB_SUB(1.0, hz);
  double tail = PGB_ADD(PGB_SUB(PGB_SUB(1.0, w), hz), PGB_MUL(PGB_MUL(z, z), p));
  return PGB_ADD(w, tail);
}
// r = x - k*pi/2, returns k mod 4
PGB_HD int pgb_rem_pio2(double x, double* r) {
  const double SHIFT = 0x1.8p52;
  double t = PGB_FMA(x, 0x1.45f306dc9c883p-1, SHIFT);
  double kd = PGB_SUB(t, SHIFT);
  double r1 = PGB_FMA(-kd, 0x1.921fb54400000p+0, x);
  r1 = PGB_FMA(-kd, 0x1.0b4611a600000p-34, r1);
  r1 = PGB_FMA(-kd, 0x1.3198a2e037073p-69, r1);
  *r = r1;
  return (int)(((long long)kd) & 3);
}
PGB_HD_NOINLINE double pgb_sin(double x) {
  if (!(fabs(x) < 1e15)) return PGB_SUB(x, x);  // inf/nan/out of domain -> nan
  double r;
  int q = pgb_rem_pio2(x, &r);
  switch (q) {
    case 0: return pgb_ksin(r);
    case 1: return pgb_kcos(r);
    case 2: return -pgb_ksin(r);
    default: return -pgb_kcos(r);
  }
}
PGB_HD_NOINLINE double pgb_cos(double x) {
  if (!(fabs(x) < 1e15)) return PGB_SUB(x, x);
  double r;
  int q = pgb_rem_pio2(x, &r);
  switch (q) {
    case 0: return pgb_kcos(r);
    case 1: return -pgb_ksin(r);
    case 2: return -pgb_kcos(r);
    default: return pgb_ksin(r);
  }
}
PGB_HD_NOINLINE void pgb_sincos(double x, double* s, double* c) {
  if (!(fabs(x) < 1e15)) {
    *s = *c = PGB_SUB(x, x);
    return;
  }
  double r;
  int q = pgb_rem_pio2(x, &r);
  double ks = pgb_ksin(r), kc = pgb_kcos(r);
  switch (q) {
    case 0: *s = ks; *c = kc; break;
    case 1: *s = kc; *c = -ks; break;
    case 2: *s = -ks; *c = -kc; break;
    default: *s = -kc; *c = ks; break;
  }
}

// ---------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the uniform /
// normal maps built on it.  One call = 128 random bits = two 53-bit uniforms = one
// Box-Muller pair.  Stream addressing is by (key = seed, counter = purpose/sweep/t/j),
// so the CPU oracle and the kernels materialise identical values without moving them.
// ---------------------------------------------------------------------------------
typedef struct { uint32_t v[4]; } pgb_u32x4;

PGB_HD void pgb_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}
PGB_HD_NOINLINE pgb_u32x4 pgb_philox4x32_10(pgb_u32x4 ctr, uint32_t k0, uint32_t k1) {
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0, lo0, hi1, lo1;
    pgb_mulhilo32(0xD2511F53u, ctr.v[0], &hi0, &lo0);
    pgb_mulhilo32(0xCD9E8D57u, ctr.v[2], &hi1, &lo1);
    pgb_u32x4 n;
    n.v[0] = hi1 ^ ctr.v[1] ^ k0;
    n.v[1] = lo1;
    n.v[2] = hi0 ^ ctr.v[3] ^ k1;
    n.v[3] = lo0;
    ctr = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ctr;
}
// 53-bit uniform in (0,1): (m + 0.5) * 2^-53, never 0 or 1
PGB_HD double pgb_u53(uint32_t hi, uint32_t lo) {
  uint64_t m = (((uint64_t)hi) << 21) | (((uint64_t)lo) >> 11);
  return PGB_MUL(PGB_ADD((double)m, 0.5), 0x1p-53);
}
// Julia-style uniform in [0,1) (rand() can return 0): m * 2^-53
PGB_HD double pgb_u53_halfopen(uint32_t hi, uint32_t lo) {
  uint64_t m = (((uint64_t)hi) << 21) | (((uint64_t)lo) >> 11);
  return PGB_MUL((double)m, 0x1p-53);
}
// Box-Muller pair from one Philox block
PGB_HD void pgb_normal_pair(pgb_u32x4 bits, double* z0, double* z1) {
  double u1 = pgb_u53(bits.v[0], bits.v[1]);
  double u2 = pgb_u53(bits.v[2], bits.v[3]);
  double rad = PGB_SQRT(PGB_MUL(-2.0, pgb_log(u1)));
  double s, c;
  pgb_sincos(PGB_MUL(0x1.921fb54442d18p+2, u2), &s, &c);
  *z0 = PGB_MUL(rad, c);
  *z1 = PGB_MUL(rad, s);
}

// Stream purposes (counter word 3, high byte) — see DESIGN.md "random streams".
enum {
  PGB_STREAM_Z0 = 1,     // initial-state normals      (sweep, j)
  PGB_STREAM_Z = 2,      // process-noise normals      (sweep, t, j)
  PGB_STREAM_URES = 3,   // resampling uniform         (sweep, t)
  PGB_STREAM_UANC = 4,   // ancestor-sampling uniform  (sweep, t)
  PGB_STREAM_USTAR = 5,  // trajectory-pick uniform    (sweep)
  PGB_STREAM_MNIW = 6,   // Bartlett factor + matrix normal (sweep, idx)
  PGB_STREAM_ROLL = 7    // scenario rollout           (scenario, kind, t)
};
// counter layout: v0 = j (particle / element pair index), v1 = t, v2 = sweep (low 32),
// v3 = purpose<<24 | (chain & 0xffffff)
PGB_HD pgb_u32x4 pgb_stream_block(uint64_t seed, uint32_t purpose, uint32_t chain, uint32_t sweep,
                                  uint32_t t, uint32_t j) {
  pgb_u32x4 c;
  c.v[0] = j;
  c.v[1] = t;
  c.v[2] = sweep;
  c.v[3] = (purpose << 24) | (chain & 0x00ffffffu);
  return pgb_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Returns 0 when the host build honours "no contraction" (a*b+c stays two roundings).
PGB_HD int pgb_math_selftest(void) {
  volatile double a = 0x1.0000001p0, b = 0x1.0000001p0, c = -0x1.0000002p0;
  double unfused = PGB_ADD(PGB_MUL(a, b), c);
  double fused = PGB_FMA(a, b, c);
  // a*b = 1 + 2^-27 + 2^-56 rounds to 1 + 2^-27 (ulp 2^-52), so unfused = 0, fused = 2^-56
  return (unfused == 0.0 && fused == 0x1p-56) ? 0 : 1;
}

// Inverse of a symmetric positive definite matrix (column-major n x n), host side only: the dense-V form of
// `I / V` (Julia/src/particle_Gibbs.jl:64) is a constant of the sampler, evaluated once when the prior is set.  Lower
// Cholesky factor (right-looking, fma updates in increasing pivot order), then column j = L' \ (L \ e_j), upper triangle
// copied from the lower one (LAPACK potri convention).  `work` holds n*n + n doubles.  Returns 1 when V is not positive
// definite (Julia: PosDefException / SingularException).  Shared by the library (pgb_set_prior_dense) and the oracle.
static inline int pgb_host_spd_inverse(const double* V, int n, double* inv, double* work) {
  double* L = work;
  double* col = work + (size_t)n * (size_t)n;
  for (size_t e = 0; e < (size_t)n * (size_t)n; ++e) L[e] = V[e];
  for (int k = 0; k < n; ++k) {
    double d = L[k + (size_t)n * k];
    if (!(d > 0.0)) return 1;
    d = sqrt(d);
    L[k + (size_t)n * k] = d;
    for (int i = k + 1; i < n; ++i) L[i + (size_t)n * k] = L[i + (size_t)n * k] / d;
    for (int j = k + 1; j < n; ++j)
      for (int i = j; i < n; ++i) L[i + (size_t)n * j] = fma(-L[i + (size_t)n * k], L[j + (size_t)n * k], L[i + (size_t)n * j]);
  }
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) col[i] = (i == j) ? 1.0 : 0.0;
    for (int d = 0; d < n; ++d) {
      double acc = col[d];
      for (int e = 0; e < d; ++e) acc = fma(-L[d + (size_t)n * e], col[e], acc);
      col[d] = acc / L[d + (size_t)n * d];
    }
    for (int d = n - 1; d >= 0; --d) {
      double acc = col[d];
      for (int e = n - 1; e > d; --e) acc = fma(-L[e + (size_t)n * d], col[e], acc);
      col[d] = acc / L[d + (size_t)n * d];
    }
    for (int i = 0; i < n; ++i) inv[i + (size_t)n * j] = col[i];
  }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i) inv[i + (size_t)n * j] = inv[j + (size_t)n * i];
  return 0;
}

#endif  // PGB_MATH_H
