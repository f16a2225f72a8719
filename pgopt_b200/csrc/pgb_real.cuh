// pgb_real.cuh — the per-particle arithmetic of the sweep kernel (pgb_v2.cuh) written once over the real type RT:
// RT = double is the FP64 path (pgb_lmath.cuh routines, reciprocal-based exact divisions), RT = float the F32 precision
// mode (pgb_config.precision = PGB_PRECISION_F32): bit-defined FP32 functions of pgb_mathf.h, IEEE FP32 division, model
// constants rounded to FP32 once.  Same products, same fma order in both — the oracle's (oracle/pg_oracle.c: eval_f /
// eval_f_f32, log_weight / log_weight_f32).
#pragma once
#include "pgb_lmath.cuh"
#include "pgb_mathf.h"
#include "pgb_persist.cuh"

namespace pgb {

__device__ __forceinline__ double rfma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float rfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double rmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float rmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double radd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float radd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double rsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float rsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double rexp(double x) { return lexp(x); }
__device__ __forceinline__ float rexp(float x) { return pgf_exp(x); }
// the transition density as an FP64 value (F32 mode: FP32 polynomial, FP64 exponent range — pgf_exp_wide)
__device__ __forceinline__ double rexp_wide(double x) { return lexp(x); }
__device__ __forceinline__ double rexp_wide(float x) { return pgf_exp_wide(x); }
__device__ __forceinline__ double rsin(double x) { return lsin(x); }
__device__ __forceinline__ float rsin(float x) { return pgf_sin(x); }
__device__ __forceinline__ double rcos(double x) { return lcos(x); }
__device__ __forceinline__ float rcos(float x) { return pgf_cos(x); }

// division by a value shared by many particles: FP64 through the reciprocal (div_inv: exact), FP32 by the IEEE divide
template <typename RT> struct RDiv;
template <> struct RDiv<double> {
  InvDiv iv;
  __device__ __forceinline__ void set(double d) { iv = make_invdiv(d); }
  __device__ __forceinline__ double operator()(double a) const { return div_inv(a, iv); }
};
template <> struct RDiv<float> {
  float d;
  __device__ __forceinline__ void set(double v) { d = (float)v; }
  __device__ __forceinline__ float operator()(float a) const { return __fdiv_rn(a, d); }
};

// F = A*phi(x, u) for ONE particle; BASIS 0 = known basis (n_x = 2), 1 = Hilbert basis with 5 x 5 x 5 functions over
// z = [u; x1; x2], 2 = general Hilbert basis.  A_sm: the chain's A in shared memory as RT (rounded once for RT = float);
// CA = 1 (single chain): A as constant-bank operands (cA) instead.  tku: scaled sines of the input dimensions.
template <int NX, int BASIS, int CA, typename RT>
__device__ __forceinline__ void r_eval_F(const FilterDev& fd, const RT* __restrict__ A_sm, const RT* __restrict__ cA,
                                         const RT (&x)[NX], const double* __restrict__ u, const RT* __restrict__ tku,
                                         RT (&F)[NX]) {
  if (BASIS == 0) {
    const RT x0 = x[0], x1 = x[NX > 1 ? 1 : 0], u0 = (RT)u[0];
    RT phi[5];
    phi[0] = rmul((RT)0.1, x0);
    phi[1] = rmul((RT)0.1, x1);
    phi[2] = u0;
    phi[3] = rmul(rmul((RT)0.01, rcos(rmul((RT)3.0, x0))), x1);
    phi[4] = rmul(rmul((RT)0.1, rsin(rmul((RT)2.0, x1))), u0);
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      RT acc = (RT)0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) acc = rfma(A_sm[d + NX * i], phi[i], acc);
      F[d] = acc;
    }
  } else if (BASIS == 1) {
    RT t1[5], t2[5];
    {
      const RT zl1 = radd(x[0], (RT)fd.hg_L[1]), zl2 = radd(x[NX > 1 ? 1 : 0], (RT)fd.hg_L[2]);
      const RT l1 = (RT)fd.hg_lsi[1], l2 = (RT)fd.hg_lsi[2];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        t1[j] = rmul(l1, rsin(rmul((RT)fd.hg_pij2l[1][j], zl1)));
        t2[j] = rmul(l2, rsin(rmul((RT)fd.hg_pij2l[2][j], zl2)));
      }
    }
    RT acc[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) acc[d] = (RT)0.0;
#pragma unroll
    for (int j0 = 0; j0 < 5; ++j0) {
      const RT t0 = tku[j0];  // = (1 * t_0): the product with 1.0 is exact
#pragma unroll
      for (int j1 = 0; j1 < 5; ++j1) {
        const RT pa = rmul(t0, t1[j1]);
#pragma unroll
        for (int j2 = 0; j2 < 5; ++j2) {
          const int i = (j0 * 5 + j1) * 5 + j2;
          const RT q = rmul(pa, t2[j2]);
#pragma unroll
          for (int d = 0; d < NX; ++d) acc[d] = rfma(CA ? cA[d + NX * i] : A_sm[d + NX * i], q, acc[d]);
        }
      }
    }
#pragma unroll
    for (int d = 0; d < NX; ++d) F[d] = acc[d];
  } else {
    const int nz = fd.n_z, nu = fd.n_u, n_phi = fd.n_phi;
    RT tk[PGB_MAX_Z_][MAXC];
    for (int k = 0; k < nz; ++k) {
      if (k < nu) {
        for (int j = 0; j < fd.hg_count[k]; ++j) tk[k][j] = tku[k * MAXC + j];
      } else {
        const RT zl = radd(x[k - nu], (RT)fd.hg_L[k]);
        const RT lk = (RT)fd.hg_lsi[k];
        for (int j = 0; j < fd.hg_count[k]; ++j) tk[k][j] = rmul(lk, rsin(rmul((RT)fd.hg_pij2l[k][j], zl)));
      }
    }
    if constexpr (sizeof(RT) == 8) {
      hilbert_apply<NX>(fd, A_sm, tk, F);
    } else {
      // phi_i = ((1*t_0)*t_1)*... with dimension 0 slowest; F_d = fma(A[d,i], phi_i, F_d) in increasing i (hilbert_apply)
      RT acc[NX];
#pragma unroll
      for (int d = 0; d < NX; ++d) acc[d] = (RT)0.0;
      int dg[PGB_MAX_Z_];
#pragma unroll
      for (int k = 0; k < PGB_MAX_Z_; ++k) dg[k] = 0;
      for (int i = 0; i < n_phi; ++i) {
        RT p = (RT)1.0;
#pragma unroll
        for (int k = 0; k < PGB_MAX_Z_; ++k)
          if (k < nz) p = rmul(p, tk[k][dg[k]]);
#pragma unroll
        for (int d = 0; d < NX; ++d) acc[d] = rfma(A_sm[d + NX * i], p, acc[d]);
        for (int k = nz - 1; k >= 0; --k) {  // odometer: last dimension fastest
          if (++dg[k] < fd.hg_count[k]) break;
          dg[k] = 0;
        }
      }
#pragma unroll
      for (int d = 0; d < NX; ++d) F[d] = acc[d];
    }
  }
}

// scaled sines of the input dimensions (the same for every particle of a step)
template <int BASIS, typename RT>
__device__ __forceinline__ void r_input_sines(const FilterDev& fd, const double* __restrict__ up, RT* __restrict__ tku) {
  if (BASIS == 1) {
    const RT zl = radd((RT)up[0], (RT)fd.hg_L[0]);
    const RT l0 = (RT)fd.hg_lsi[0];
#pragma unroll
    for (int j = 0; j < 5; ++j) tku[j] = rmul(l0, rsin(rmul((RT)fd.hg_pij2l[0][j], zl)));
  } else if (BASIS == 2) {
    for (int k = 0; k < fd.n_u; ++k) {
      const RT zl = radd((RT)up[k], (RT)fd.hg_L[k]);
      const RT lk = (RT)fd.hg_lsi[k];
      for (int j = 0; j < fd.hg_count[k]; ++j) tku[k * MAXC + j] = rmul(lk, rsin(rmul((RT)fd.hg_pij2l[k][j], zl)));
    }
  }
}

// measurement log-weight (:194 | :196); FP64: the scalar-R quotient through the shared reciprocal (exact)
template <int NX>
__device__ __forceinline__ double r_log_weight(const FilterDev& fd, const double (&x)[NX], const double* __restrict__ y,
                                               const RDiv<double>& divR) {
  if (fd.n_y == 1) {
    double g = 0.0;
#pragma unroll
    for (int d = 0; d < NX; ++d) g = __fma_rn(fd.C[d], x[d], g);
    const double dd = __dsub_rn(g, y[0]);
    return divR(__dmul_rn(-__dmul_rn(dd, dd), 0.5));
  }
  return log_weight<NX>(fd, x, y);
}
template <int NX>
__device__ __forceinline__ float r_log_weight(const FilterDev& fd, const float (&x)[NX], const double* __restrict__ y,
                                              const RDiv<float>& divR) {
  const int ny = fd.n_y;
  if (ny == 1) {
    float g = 0.0f;
#pragma unroll
    for (int d = 0; d < NX; ++d) g = __fmaf_rn((float)fd.C[d], x[d], g);
    const float dd = __fsub_rn(g, (float)y[0]);
    return divR(__fmul_rn(-__fmul_rn(dd, dd), 0.5f));
  }
  float r[4];
  for (int o = 0; o < ny; ++o) {
    float g = 0.0f;
#pragma unroll
    for (int d = 0; d < NX; ++d) g = __fmaf_rn((float)fd.C[o + ny * d], x[d], g);
    r[o] = __fsub_rn((float)y[o], g);
  }
  for (int d = 0; d < ny; ++d) {  // forward substitution with (float)chol(R)
    float acc = r[d];
    for (int e = 0; e < d; ++e) acc = __fmaf_rn(-(float)fd.Rchol[d + ny * e], r[e], acc);
    r[d] = __fdiv_rn(acc, (float)fd.Rchol[d + ny * d]);
  }
  float s = 0.0f;
  for (int o = 0; o < ny; ++o) s = __fmaf_rn(r[o], r[o], s);
  return __fmul_rn(-s, 0.5f);
}

// standard normals of output slot j (all NX dims): Philox (FP64: block j*NP + p -> dims (2p, 2p+1); F32: one block per
// particle, words (0,1) -> dims (0,1), (2,3) -> dims (2,3)) or the injected stream
template <int NX>
__device__ __forceinline__ void r_normals(const FilterDev& fd, uint32_t stream_word, uint32_t t, int64_t j,
                                          const double* __restrict__ inj, double (&z)[NX]) {
  if (fd.philox) {
    constexpr int NP = (NX + 1) / 2;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      uint32_t c0 = (uint32_t)(j * NP + p), c1 = t, c2 = fd.sweep, c3 = stream_word;
      lphilox(c0, c1, c2, c3, (uint32_t)fd.seed, (uint32_t)(fd.seed >> 32));
      double z0, z1;
      lnormal_pair(c0, c1, c2, c3, z0, z1);
      z[2 * p] = z0;
      if (2 * p + 1 < NX) z[2 * p + 1] = z1;
    }
  } else {
#pragma unroll
    for (int d = 0; d < NX; ++d) z[d] = inj[j * NX + d];
  }
}
template <int NX>
__device__ __forceinline__ void r_normals(const FilterDev& fd, uint32_t stream_word, uint32_t t, int64_t j,
                                          const double* __restrict__ inj, float (&z)[NX]) {
  if (fd.philox) {
    uint32_t c0 = (uint32_t)j, c1 = t, c2 = fd.sweep, c3 = stream_word;
    lphilox(c0, c1, c2, c3, (uint32_t)fd.seed, (uint32_t)(fd.seed >> 32));
    float a0, a1;
    pgf_normal_pair(c0, c1, &a0, &a1);
    z[0] = a0;
    if (NX > 1) z[NX > 1 ? 1 : 0] = a1;
    if (NX > 2) {
      pgf_normal_pair(c2, c3, &a0, &a1);
      z[NX > 2 ? 2 : 0] = a0;
      if (NX > 3) z[NX > 3 ? 3 : 0] = a1;
    }
  } else {
#pragma unroll
    for (int d = 0; d < NX; ++d) z[d] = (float)inj[j * NX + d];
  }
}

// row t of the particle history as RT (the buffer is sized in FP64 elements; the F32 mode uses its first half)
template <typename RT>
__device__ __forceinline__ RT* r_x_row(const FilterDev& fd, int chain, int64_t t) {
  const int64_t r = (fd.x_rows == fd.T) ? t : (t & 1);
  return reinterpret_cast<RT*>(fd.xh) + ((int64_t)chain * fd.x_rows + r) * fd.n_x * fd.N;
}

}  // namespace pgb
