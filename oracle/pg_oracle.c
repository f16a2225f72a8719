/* pg_oracle.c — CPU restatement of PGopt's particle-Gibbs hot path.  TEST INFRASTRUCTURE ONLY,
 * PARITY UNPINNED (see pg_oracle.h).  Single-threaded, FP64, every evaluation order written out.
 *
 * Orders fixed here because the reference leaves them to Julia Base / OpenBLAS / Distributions.jl:
 *   sum(v)            balanced binary tree over the index range padded with +0.0 to a power of
 *                     two (Julia's own sum is pairwise with SIMD leaf blocks; LLVM-dependent)
 *   A*phi, C*x, L*z   acc = 0; acc = fma(a_i, b_i, acc) in increasing i
 *   L \ r             forward substitution, acc = r_d; acc = fma(-L[d,e], z_e, acc); z_d = acc/L[d,d]
 *   exp/sin/cos/log   pgb_math.h (shared with the kernels, < 2 ulp)
 * The sequential recurrences the reference *does* pin (q = q + W[n], u = u + 1/N) are kept strictly
 * sequential here.
 */
#include "pg_oracle.h"

#include <stdlib.h>
#include <string.h>

#include "../pgopt_b200/csrc/pgb_mathf.h"

#define LOG2PI 1.8378770664093453 /* Float64(log(2*pi)) */

/* ------------------------------------------------------------------ math wrappers */
void orc_vexp(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = pgb_exp(x[i]); }
void orc_vlog(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = pgb_log(x[i]); }
void orc_vsin(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = pgb_sin(x[i]); }
void orc_vcos(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = pgb_cos(x[i]); }
void orc_vnormal_pairs(uint64_t seed, uint32_t purpose, uint32_t chain, uint32_t sweep, uint32_t t,
                       int64_t npairs, double* z) {
  for (int64_t j = 0; j < npairs; ++j)
    pgb_normal_pair(pgb_stream_block(seed, purpose, chain, sweep, t, (uint32_t)j), &z[2 * j], &z[2 * j + 1]);
}
void orc_vmathf(int fn, const float* x, float* y, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    y[i] = fn == 0 ? pgf_exp(x[i]) : fn == 1 ? pgf_log_pos_normal(x[i]) : fn == 2 ? pgf_sin(x[i]) : fn == 3 ? pgf_cos(x[i]) : (float)pgf_exp_wide(x[i]);
}
void orc_vnormal_pairs_f32(uint64_t seed, uint32_t purpose, uint32_t chain, uint32_t sweep, uint32_t t, int64_t n,
                           int n_x, float* z) {
  for (int64_t j = 0; j < n; ++j) pgf_stream_normals(seed, purpose, chain, sweep, t, (uint32_t)j, n_x, &z[j * n_x]);
}
int orc_selftest(void) { return pgb_math_selftest(); }

/* ------------------------------------------------------------------ sum */
static double psum_rec(const double* v, int64_t lo, int64_t size, int64_t n) {
  if (lo >= n) return 0.0;
  if (size == 1) return v[lo];
  int64_t h = size >> 1;
  return psum_rec(v, lo, h, n) + psum_rec(v, lo + h, h, n);
}
double orc_pairwise_sum(const double* v, int64_t n) {
  if (n <= 0) return 0.0;
  int64_t size = 1;
  while (size < n) size <<= 1;
  return psum_rec(v, 0, size, n);
}

/* ------------------------------------------------------------------ systematic_resampling
 * particle_Gibbs.jl:17-34.  `rnd` stands for the rand() at :21.  */
int orc_systematic_resampling(const double* W_in, int64_t n_w, int64_t n_draw, double rnd, int64_t* idx) {
  int status = 0;
  double* W = (double*)malloc(sizeof(double) * (size_t)(n_w > 0 ? n_w : 1));
  double s = orc_pairwise_sum(W_in, n_w); /* :19 */
  for (int64_t i = 0; i < n_w; ++i) {
    W[i] = W_in[i] / s;
    if (!(W[i] >= 0.0 && W[i] <= 1.7976931348623157e308)) status |= 4; /* NaN/Inf weights: the reference carries on silently */
  }
  double inc = 1.0 / (double)n_draw;
  double u = inc * rnd; /* :21  u = 1 / N * rand() */
  double q = 0.0;       /* :23 */
  int64_t n = 0;        /* :24 */
  for (int64_t i = 0; i < n_draw; ++i) { /* :25 */
    while (q < u) {                      /* :26 */
      n = n + 1;                         /* :27 */
      if (n > n_w) {                     /* Julia: BoundsError at W[n] */
        status |= 1;
        n = n_w;
        break;
      }
      q = q + W[n - 1]; /* :28 */
    }
    if (n == 0) status |= 2; /* idx 0: the reference returns it and fails later on indexing */
    idx[i] = n;              /* :30 */
    u = u + inc;             /* :31 */
  }
  free(W);
  return status;
}

/* ------------------------------------------------------------------ basis functions */
void orc_phi(const orc_model* m, const double* x, const double* u, double* phi) {
  if (m->basis == ORC_BASIS_KNOWN_VB) {
    /* examples/PG_OCP_known_basis_functions_Altro.jl:34
     * [0.1*x1  0.1*x2  u1  0.01*cos(3*x1).*x2  0.1*sin(2*x2).*u1]  (left-assoc products) */
    phi[0] = 0.1 * x[0];
    phi[1] = 0.1 * x[1];
    phi[2] = u[0];
    phi[3] = (0.01 * pgb_cos(3.0 * x[0])) * x[1];
    phi[4] = (0.1 * pgb_sin(2.0 * x[1])) * u[0];
    return;
  }
  /* examples/PG_OCP_generic_basis_functions_Altro.jl:85-99 (phi_sampling):
   *   z = vcat(u, x); phi = ones; for k: phi .*= L_sqrt_inv[k] .* sin.(pi_j_over_2L[:,k] * (z[k] + L[k]))
   * with pi_j_over_2L = (pi*j)/(2*L) (:86) and L_sqrt_inv = 1/sqrt(L) (:85). */
  int n_z = m->n_u + m->n_x;
  double z[ORC_MAX_Z];
  for (int k = 0; k < m->n_u; ++k) z[k] = u[k];
  for (int k = 0; k < m->n_x; ++k) z[m->n_u + k] = x[k];
  const double PI = 3.141592653589793;
  for (int i = 0; i < m->n_phi; ++i) {
    double p = 1.0;
    for (int k = 0; k < n_z; ++k) {
      double j = (double)((i / m->stride[k]) % m->count[k] + 1);
      double pij2l = (PI * j) / (2.0 * m->L[k]);
      double lsi = 1.0 / sqrt(m->L[k]);
      double arg = pij2l * (z[k] + m->L[k]);
      p = p * (lsi * pgb_sin(arg));
    }
    phi[i] = p;
  }
}

/* f(x,u) = A*phi(x,u)   particle_Gibbs.jl:156 */
static void eval_f(const orc_model* m, const double* A, const double* x, const double* u, double* phi_tmp,
                   double* out) {
  orc_phi(m, x, u, phi_tmp);
  for (int d = 0; d < m->n_x; ++d) {
    double acc = 0.0;
    for (int i = 0; i < m->n_phi; ++i) acc = fma(A[d + m->n_x * i], phi_tmp[i], acc);
    out[d] = acc;
  }
}

/* f(x,u) = PG_samples[k].A * phi(x,u) in the scenario rollout (optimal_control_Ipopt.jl:62,:121).  Julia hands this
 * product to BLAS gemv, whose summation order is not pinned by the reference; here it is DEFINED as a blocked sum:
 * the columns are cut into 128 runs of c = ceil(n_phi/128) consecutive columns, run j accumulates
 * p_j = fma(A[d, i], phi[i], p_j) over its columns in increasing i (empty runs give +0.0), and the 128 partial sums
 * are combined by the balanced binary tree over j.  (The filter's per-particle product, eval_f above, stays strictly
 * sequential in i.) */
#define ORC_ROLL_LANES 128
static void eval_f_blocked(const orc_model* m, const double* A, const double* x, const double* u, double* phi_tmp,
                           double* out) {
  orc_phi(m, x, u, phi_tmp);
  double part[ORC_ROLL_LANES];
  const int c = (m->n_phi + ORC_ROLL_LANES - 1) / ORC_ROLL_LANES;
  for (int d = 0; d < m->n_x; ++d) {
    for (int j = 0; j < ORC_ROLL_LANES; ++j) {
      double acc = 0.0;
      for (int i = j * c; i < (j + 1) * c && i < m->n_phi; ++i) acc = fma(A[d + m->n_x * i], phi_tmp[i], acc);
      part[j] = acc;
    }
    for (int w = 1; w < ORC_ROLL_LANES; w <<= 1)
      for (int j = 0; j < ORC_ROLL_LANES; j += 2 * w) part[j] = part[j] + part[j + w];
    out[d] = part[0];
  }
}

/* ------------------------------------------------------------------ small dense helpers */
/* Cholesky, right-looking / outer-product form: every entry (i,j) receives the updates
 * a_ij = fma(-l_ik, l_jk, a_ij) in increasing k — the order a parallel right-looking
 * factorisation performs them in as well.  Column-major, lower; upper part zeroed. */
int orc_cholesky_lower(double* A, int n) {
  for (int k = 0; k < n; ++k) {
    double d = A[k + n * k];
    if (!(d > 0.0)) return 1; /* Julia: PosDefException */
    d = sqrt(d);
    A[k + n * k] = d;
    for (int i = k + 1; i < n; ++i) A[i + n * k] = A[i + n * k] / d;
    for (int j = k + 1; j < n; ++j)
      for (int i = j; i < n; ++i) A[i + n * j] = fma(-A[i + n * k], A[j + n * k], A[i + n * j]);
  }
  for (int j = 1; j < n; ++j)
    for (int i = 0; i < j; ++i) A[i + n * j] = 0.0;
  return 0;
}
/* solve L y = b (forward), in place */
static void fwd_subst(const double* L, int n, double* b) {
  for (int d = 0; d < n; ++d) {
    double acc = b[d];
    for (int e = 0; e < d; ++e) acc = fma(-L[d + n * e], b[e], acc);
    b[d] = acc / L[d + n * d];
  }
}
/* solve L' y = b (backward), in place */
static void bwd_subst_t(const double* L, int n, double* b) {
  for (int d = n - 1; d >= 0; --d) {
    double acc = b[d];
    for (int e = n - 1; e > d; --e) acc = fma(-L[e + n * d], b[e], acc);
    b[d] = acc / L[d + n * d];
  }
}
/* inverse of an SPD matrix from its lower Cholesky factor: column j = L'\(L\e_j);
 * result symmetrised by copying the lower triangle (LAPACK potri convention). */
static void spd_inverse_from_chol(const double* L, int n, double* inv) {
  double* col = (double*)malloc(sizeof(double) * (size_t)n);
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) col[i] = (i == j) ? 1.0 : 0.0;
    fwd_subst(L, n, col);
    bwd_subst_t(L, n, col);
    for (int i = 0; i < n; ++i) inv[i + n * j] = col[i];
  }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i) inv[i + n * j] = inv[j + n * i];
  free(col);
}

/* ------------------------------------------------------------------ random streams */
static double stream_uniform(const orc_streams* s, uint32_t purpose, uint32_t sweep, uint32_t t) {
  pgb_u32x4 b = pgb_stream_block(s->seed, purpose, s->chain, sweep, t, 0);
  return pgb_u53_halfopen(b.v[0], b.v[1]);
}
/* normals for particle j (all n_x dims): pairs (2p, 2p+1) come from block j*npairs + p */
static void stream_normals(const orc_streams* s, uint32_t purpose, uint32_t sweep, uint32_t t, int64_t j,
                           int n_x, double* z) {
  int npairs = (n_x + 1) / 2;
  for (int p = 0; p < npairs; ++p) {
    double z0, z1;
    pgb_normal_pair(pgb_stream_block(s->seed, purpose, s->chain, sweep, t, (uint32_t)(j * npairs + p)), &z0, &z1);
    z[2 * p] = z0;
    if (2 * p + 1 < n_x) z[2 * p + 1] = z1;
  }
}

/* measurement log-weight, particle_Gibbs.jl:193-197 with g(x,u) = C*x */
static double log_weight(const orc_model* m, const double* C, const double* R, const double* Rchol,
                         const double* x, const double* y) {
  int n_y = m->n_y, n_x = m->n_x;
  if (n_y == 1) {
    double g = 0.0;
    for (int d = 0; d < n_x; ++d) g = fma(C[n_y * d], x[d], g);
    double dd = g - y[0];
    return -(dd * dd) / 2.0 / R[0]; /* :194  -(d)^2 / 2 / R */
  }
  /* :196  -sum(r .* (R \ r)) / 2 ; R \ r realised as a Cholesky solve: r'R^-1 r = |L^-1 r|^2 */
  double r[ORC_MAX_Z];
  for (int o = 0; o < n_y; ++o) {
    double g = 0.0;
    for (int d = 0; d < n_x; ++d) g = fma(C[o + n_y * d], x[d], g);
    r[o] = y[o] - g;
  }
  fwd_subst(Rchol, n_y, r);
  double sacc = 0.0;
  for (int o = 0; o < n_y; ++o) sacc = fma(r[o], r[o], sacc);
  return -sacc / 2.0;
}

/* ------------------------------------------------------------------ one PG sweep */
/* ------------------------------------------------------------------ F32 precision mode (PGB_PRECISION_F32)
 * The north star's "FP32 particle state": particle states, basis functions, A*phi, process noise, log-weights and the two
 * exponentials per particle are evaluated in FP32 with the bit-defined functions of pgb_mathf.h; every model constant is
 * rounded to FP32 ONCE ((float)A, (float)chol(Q), (float)C, ...).  The un-normalised weights e = exp(lw - max) and the
 * transition densities are widened to FP64 and everything downstream — the two normalisations, the sequential
 * cumulative sums of systematic resampling / ancestor sampling, the backward trace, the statistics, MNIW — is the FP64
 * code above, so index parity with the kernels is again bit-exact.  The history arrays keep FP64 storage holding
 * FP32 values. */
static void orc_phi_f32(const orc_model* m, const float* x, const double* u, float* phi) {
  if (m->basis == ORC_BASIS_KNOWN_VB) {
    const float u0 = (float)u[0];
    phi[0] = PGF_MUL(0.1f, x[0]);
    phi[1] = PGF_MUL(0.1f, x[1]);
    phi[2] = u0;
    phi[3] = PGF_MUL(PGF_MUL(0.01f, pgf_cos(PGF_MUL(3.0f, x[0]))), x[1]);
    phi[4] = PGF_MUL(PGF_MUL(0.1f, pgf_sin(PGF_MUL(2.0f, x[1]))), u0);
    return;
  }
  int n_z = m->n_u + m->n_x;
  float z[ORC_MAX_Z];
  for (int k = 0; k < m->n_u; ++k) z[k] = (float)u[k];
  for (int k = 0; k < m->n_x; ++k) z[m->n_u + k] = x[k];
  const double PI = 3.141592653589793;
  for (int i = 0; i < m->n_phi; ++i) {
    float p = 1.0f;
    for (int k = 0; k < n_z; ++k) {
      double j = (double)((i / m->stride[k]) % m->count[k] + 1);
      const float pij2l = (float)((PI * j) / (2.0 * m->L[k])); /* the FP64 constants of orc_phi, rounded once */
      const float lsi = (float)(1.0 / sqrt(m->L[k]));
      const float arg = PGF_MUL(pij2l, PGF_ADD(z[k], (float)m->L[k]));
      p = PGF_MUL(p, PGF_MUL(lsi, pgf_sin(arg)));
    }
    phi[i] = p;
  }
}
static void eval_f_f32(const orc_model* m, const double* A, const float* x, const double* u, float* phi_tmp, float* out) {
  orc_phi_f32(m, x, u, phi_tmp);
  for (int d = 0; d < m->n_x; ++d) {
    float acc = 0.0f;
    for (int i = 0; i < m->n_phi; ++i) acc = PGF_FMA((float)A[d + m->n_x * i], phi_tmp[i], acc);
    out[d] = acc;
  }
}
static float log_weight_f32(const orc_model* m, const double* C, const double* R, const double* Rchol, const float* x,
                            const double* y) {
  int n_y = m->n_y, n_x = m->n_x;
  if (n_y == 1) {
    float g = 0.0f;
    for (int d = 0; d < n_x; ++d) g = PGF_FMA((float)C[n_y * d], x[d], g);
    const float dd = PGF_SUB(g, (float)y[0]);
    return PGF_DIV(PGF_MUL(-PGF_MUL(dd, dd), 0.5f), (float)R[0]);
  }
  float r[ORC_MAX_Z];
  for (int o = 0; o < n_y; ++o) {
    float g = 0.0f;
    for (int d = 0; d < n_x; ++d) g = PGF_FMA((float)C[o + n_y * d], x[d], g);
    r[o] = PGF_SUB((float)y[o], g);
  }
  for (int d = 0; d < n_y; ++d) { /* forward substitution with (float)chol(R) */
    float acc = r[d];
    for (int e = 0; e < d; ++e) acc = PGF_FMA(-(float)Rchol[d + n_y * e], r[e], acc);
    r[d] = PGF_DIV(acc, (float)Rchol[d + n_y * d]);
  }
  float sacc = 0.0f;
  for (int o = 0; o < n_y; ++o) sacc = PGF_FMA(r[o], r[o], sacc);
  return PGF_MUL(-sacc, 0.5f);
}
static void stream_normals_f32(const orc_streams* s, uint32_t purpose, uint32_t sweep, uint32_t t, int64_t j, int n_x, float* z) {
  pgf_stream_normals(s->seed, purpose, s->chain, sweep, t, (uint32_t)j, n_x, z);
}

static int sweep_impl(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y, const double* C,
              const double* R, const double* x0_mean, const double* x0_chol, const double* A, const double* Q,
              double* x_prim, int first, const orc_streams* s, uint32_t sweep_index, int64_t* a_hist,
              double* x_hist, double* w_hist, double* w_last, double* x_last, double* Phi, double* Psi,
              double* Sigma, int64_t* star_out, int f32);

int orc_sweep(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y, const double* C,
              const double* R, const double* x0_mean, const double* x0_chol, const double* A, const double* Q,
              double* x_prim, int first, const orc_streams* s, uint32_t sweep_index, int64_t* a_hist,
              double* x_hist, double* w_hist, double* w_last, double* x_last, double* Phi, double* Psi,
              double* Sigma, int64_t* star_out) {
  return sweep_impl(m, N, T, u, y, C, R, x0_mean, x0_chol, A, Q, x_prim, first, s, sweep_index, a_hist, x_hist, w_hist, w_last,
                    x_last, Phi, Psi, Sigma, star_out, 0);
}
int orc_sweep_f32(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y, const double* C,
                  const double* R, const double* x0_mean, const double* x0_chol, const double* A, const double* Q,
                  double* x_prim, int first, const orc_streams* s, uint32_t sweep_index, int64_t* a_hist,
                  double* x_hist, double* w_hist, double* w_last, double* x_last, double* Phi, double* Psi,
                  double* Sigma, int64_t* star_out) {
  return sweep_impl(m, N, T, u, y, C, R, x0_mean, x0_chol, A, Q, x_prim, first, s, sweep_index, a_hist, x_hist, w_hist, w_last,
                    x_last, Phi, Psi, Sigma, star_out, 1);
}

static int sweep_impl(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y, const double* C,
              const double* R, const double* x0_mean, const double* x0_chol, const double* A, const double* Q,
              double* x_prim, int first, const orc_streams* s, uint32_t sweep_index, int64_t* a_hist,
              double* x_hist, double* w_hist, double* w_last, double* x_last, double* Phi, double* Psi,
              double* Sigma, int64_t* star_out, int f32) {
  const int n_x = m->n_x, n_u = m->n_u, n_y = m->n_y, n_phi = m->n_phi;
  int status = 0;
  /* full histories are needed for the backward trace (:213-218) */
  double* xh = (double*)malloc(sizeof(double) * (size_t)(T * N * n_x));   /* x_pf[:, n, t] at (t*N+n)*n_x */
  int64_t* ah = (int64_t*)calloc((size_t)(T * N), sizeof(int64_t));       /* a[t, n] 1-based at t*N+n */
  double* w = (double*)malloc(sizeof(double) * (size_t)N);
  double* wprev = (double*)malloc(sizeof(double) * (size_t)N);
  double* lw = (double*)malloc(sizeof(double) * (size_t)N);
  double* waN = (double*)malloc(sizeof(double) * (size_t)N);
  double* F = (double*)malloc(sizeof(double) * (size_t)(N * n_x));
  double* phi_tmp = (double*)malloc(sizeof(double) * (size_t)n_phi);
  float* phi_tmp_f = (float*)malloc(sizeof(float) * (size_t)n_phi);
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
  double Lq[ORC_MAX_Z * ORC_MAX_Z], Rchol[ORC_MAX_Z * ORC_MAX_Z], z[ORC_MAX_Z], r[ORC_MAX_Z];

  /* mvn_v = MvNormal(zeros, Q) (:157) -> lower Cholesky factor of Q */
  memcpy(Lq, Q, sizeof(double) * (size_t)(n_x * n_x));
  if (orc_cholesky_lower(Lq, n_x)) { status = 16; goto done; }
  if (n_y > 1) {
    memcpy(Rchol, R, sizeof(double) * (size_t)(n_y * n_y));
    if (orc_cholesky_lower(Rchol, n_y)) { status = 16; goto done; }
  }
  /* c0 = -(n*log(2pi) + logdet(Q))/2 with logdet = 2*sum(log(diag(L)))  (Distributions mvnormal_c0) */
  double logdet = 0.0;
  for (int d = 0; d < n_x; ++d) logdet = logdet + pgb_log(Lq[d + n_x * d]);
  logdet = 2.0 * logdet;
  double c0 = -((double)n_x * LOG2PI + logdet) / 2.0;

  /* x_pf[:, end, :] .= x_prim (:160) */
  for (int64_t t = 0; t < T; ++t)
    for (int d = 0; d < n_x; ++d) xh[(t * N + (N - 1)) * n_x + d] = f32 ? (double)(float)x_prim[d + n_x * t] : x_prim[d + n_x * t];
  /* initial states (:163-165): rand(x_init_dist) = mean + L*z */
  float zf[ORC_MAX_Z], xf[ORC_MAX_Z], Ff[ORC_MAX_Z];
  for (int64_t n = 0; n < N - 1; ++n) {
    if (f32) {
      if (s->philox) stream_normals_f32(s, PGB_STREAM_Z0, sweep_index, 0, n, n_x, zf);
      else for (int d = 0; d < n_x; ++d) zf[d] = (float)s->Z0[n * n_x + d];
      for (int d = 0; d < n_x; ++d) {
        float acc = 0.0f;
        for (int e = 0; e <= d; ++e) acc = PGF_FMA((float)x0_chol[d + n_x * e], zf[e], acc);
        xh[(0 * N + n) * n_x + d] = (double)PGF_ADD((float)x0_mean[d], acc);
      }
      continue;
    }
    if (s->philox) stream_normals(s, PGB_STREAM_Z0, sweep_index, 0, n, n_x, z);
    else for (int d = 0; d < n_x; ++d) z[d] = s->Z0[n * n_x + d];
    for (int d = 0; d < n_x; ++d) {
      double acc = 0.0;
      for (int e = 0; e <= d; ++e) acc = fma(x0_chol[d + n_x * e], z[e], acc);
      xh[(0 * N + n) * n_x + d] = x0_mean[d] + acc;
    }
  }

  for (int64_t t = 0; t < T; ++t) { /* :168 */
    double* xt = &xh[t * N * n_x];
    if (t >= 1) {
      const double* xp = &xh[(t - 1) * N * n_x];
      const double* up = &u[n_u * (t - 1)];
      int64_t ndraw = first ? N : N - 1;
      double ures = s->philox ? stream_uniform(s, PGB_STREAM_URES, sweep_index, (uint32_t)t) : s->Ures[t];
      /* f evaluated once per particle; the reference evaluates it on the gathered columns (:175) and
       * again on all columns (:179) — identical values, f is a pure per-column function */
      for (int64_t n = 0; n < N; ++n) {
        if (f32) {
          for (int d = 0; d < n_x; ++d) xf[d] = (float)xp[n * n_x + d];
          eval_f_f32(m, A, xf, up, phi_tmp_f, Ff);
          for (int d = 0; d < n_x; ++d) F[n * n_x + d] = (double)Ff[d];
        } else {
          eval_f(m, A, &xp[n * n_x], up, phi_tmp, &F[n * n_x]);
        }
      }
      status |= orc_systematic_resampling(wprev, N, ndraw, ures, idx); /* :172 | :185 */
      for (int64_t j = 0; j < ndraw; ++j) {
        int64_t aj = idx[j] < 1 ? 1 : idx[j];
        ah[t * N + j] = idx[j];
        if (f32) {
          if (s->philox) stream_normals_f32(s, PGB_STREAM_Z, sweep_index, (uint32_t)t, j, n_x, zf);
          else for (int d = 0; d < n_x; ++d) zf[d] = (float)s->Z[t * (N * n_x) + j * n_x + d];
          for (int d = 0; d < n_x; ++d) {
            float acc = 0.0f;
            for (int e = 0; e <= d; ++e) acc = PGF_FMA((float)Lq[d + n_x * e], zf[e], acc);
            xt[j * n_x + d] = (double)PGF_ADD((float)F[(aj - 1) * n_x + d], acc);
          }
          continue;
        }
        if (s->philox) stream_normals(s, PGB_STREAM_Z, sweep_index, (uint32_t)t, j, n_x, z);
        else for (int d = 0; d < n_x; ++d) z[d] = s->Z[t * (N * n_x) + j * n_x + d];
        for (int d = 0; d < n_x; ++d) { /* :175 | :188   f(x[a]) + L*z */
          double acc = 0.0;
          for (int e = 0; e <= d; ++e) acc = fma(Lq[d + n_x * e], z[e], acc);
          xt[j * n_x + d] = F[(aj - 1) * n_x + d] + acc;
        }
      }
      if (!first) {
        /* ancestor sampling (:178-181): waN = w .* pdf(MvNormal(x_prim_t, Q), f(x_{t-1})) */
        double uanc = s->philox ? stream_uniform(s, PGB_STREAM_UANC, sweep_index, (uint32_t)t) : s->Uanc[t];
        const double* xref = &xt[(N - 1) * n_x];
        for (int64_t n = 0; n < N; ++n) {
          if (f32) {
            float rf[ORC_MAX_Z], sqf = 0.0f;
            for (int d = 0; d < n_x; ++d) {
              float acc = PGF_SUB((float)F[n * n_x + d], (float)xref[d]);
              for (int e = 0; e < d; ++e) acc = PGF_FMA(-(float)Lq[d + n_x * e], rf[e], acc);
              rf[d] = PGF_DIV(acc, (float)Lq[d + n_x * d]);
              sqf = PGF_FMA(rf[d], rf[d], sqf);
            }
            waN[n] = wprev[n] * pgf_exp_wide(PGF_SUB((float)c0, PGF_MUL(sqf, 0.5f))); /* FP32 polynomial, FP64 exponent range */
            continue;
          }
          for (int d = 0; d < n_x; ++d) r[d] = F[n * n_x + d] - xref[d];
          fwd_subst(Lq, n_x, r);
          double sq = 0.0;
          for (int d = 0; d < n_x; ++d) sq = fma(r[d], r[d], sq);
          waN[n] = wprev[n] * pgb_exp(c0 - sq / 2.0);
        }
        double sw = orc_pairwise_sum(waN, N); /* :180 */
        for (int64_t n = 0; n < N; ++n) waN[n] = waN[n] / sw;
        int64_t one;
        status |= orc_systematic_resampling(waN, N, 1, uanc, &one); /* :181 */
        ah[t * N + (N - 1)] = one;
      }
    }
    /* weights (:193-199) */
    double mx = -INFINITY;
    for (int64_t n = 0; n < N; ++n) {
      if (f32) {
        for (int d = 0; d < n_x; ++d) xf[d] = (float)xt[n * n_x + d];
        lw[n] = (double)log_weight_f32(m, C, R, Rchol, xf, &y[n_y * t]);
      } else {
        lw[n] = log_weight(m, C, R, Rchol, &xt[n * n_x], &y[n_y * t]);
      }
      if (lw[n] > mx) mx = lw[n];
    }
    for (int64_t n = 0; n < N; ++n) w[n] = f32 ? (double)pgf_exp(PGF_SUB((float)lw[n], (float)mx)) : pgb_exp(lw[n] - mx);
    double sw = orc_pairwise_sum(w, N);
    for (int64_t n = 0; n < N; ++n) w[n] = w[n] / sw;
    if (w_hist) memcpy(&w_hist[t * N], w, sizeof(double) * (size_t)N);
    memcpy(wprev, w, sizeof(double) * (size_t)N);
  }
  if (w_last) memcpy(w_last, w, sizeof(double) * (size_t)N);                                   /* :206 */
  if (x_last) memcpy(x_last, &xh[(T - 1) * N * n_x], sizeof(double) * (size_t)(N * n_x));     /* :207 */

  /* trajectory to condition on (:213-218) */
  {
    double ustar = s->philox ? stream_uniform(s, PGB_STREAM_USTAR, sweep_index, 0) : s->Ustar;
    int64_t star;
    status |= orc_systematic_resampling(w, N, 1, ustar, &star);
    if (star < 1) star = 1;
    if (star_out) *star_out = star;
    for (int d = 0; d < n_x; ++d) x_prim[d + n_x * (T - 1)] = xh[((T - 1) * N + (star - 1)) * n_x + d];
    for (int64_t t = T - 2; t >= 0; --t) {
      star = ah[(t + 1) * N + (star - 1)];
      if (star < 1) star = 1;
      for (int d = 0; d < n_x; ++d) x_prim[d + n_x * t] = xh[(t * N + (star - 1)) * n_x + d];
    }
  }
  /* statistics (:221-225): zeta = x_prim[:,2:T]; z = phi(x_prim[:,1:T-1], u[:,1:T-1]) */
  if (Phi && Psi && Sigma) {
    int64_t Tm1 = T - 1;
    double* zeta = (double*)malloc(sizeof(double) * (size_t)(n_x * Tm1));
    double* zz = (double*)malloc(sizeof(double) * (size_t)(n_phi * Tm1));
    double* prod = (double*)malloc(sizeof(double) * (size_t)(Tm1 > 0 ? Tm1 : 1));
    for (int64_t t = 0; t < Tm1; ++t) {
      for (int d = 0; d < n_x; ++d) zeta[d + n_x * t] = x_prim[d + n_x * (t + 1)];
      orc_phi(m, &x_prim[n_x * t], &u[n_u * t], &zz[n_phi * t]);
    }
    for (int a = 0; a < n_x; ++a)
      for (int b = 0; b < n_x; ++b) {
        for (int64_t t = 0; t < Tm1; ++t) prod[t] = zeta[a + n_x * t] * zeta[b + n_x * t];
        Phi[a + n_x * b] = orc_pairwise_sum(prod, Tm1);
      }
    for (int a = 0; a < n_x; ++a)
      for (int i = 0; i < n_phi; ++i) {
        for (int64_t t = 0; t < Tm1; ++t) prod[t] = zeta[a + n_x * t] * zz[i + n_phi * t];
        Psi[a + n_x * i] = orc_pairwise_sum(prod, Tm1);
      }
    for (int i = 0; i < n_phi; ++i)
      for (int j = 0; j < n_phi; ++j) {
        for (int64_t t = 0; t < Tm1; ++t) prod[t] = zz[i + n_phi * t] * zz[j + n_phi * t];
        Sigma[i + n_phi * j] = orc_pairwise_sum(prod, Tm1);
      }
    free(zeta); free(zz); free(prod);
  }
  if (a_hist) memcpy(a_hist, ah, sizeof(int64_t) * (size_t)(T * N));
  if (x_hist) memcpy(x_hist, xh, sizeof(double) * (size_t)(T * N * n_x));
done:
  free(xh); free(ah); free(w); free(wprev); free(lw); free(waN); free(F); free(phi_tmp); free(idx); free(phi_tmp_f);
  return status;
}

/* ------------------------------------------------------------------ MNIW_sample
 * particle_Gibbs.jl:56-81.  V diagonal (both reference examples).  The `/` (LU in Julia) solves are
 * realised by a Cholesky factorisation of Sigbar (SPD); Distributions.jl internals restated:
 *   rand(InverseWishart(df, Psi)) = inv(rand(Wishart(df, inv(Psi)))),
 *   rand(Wishart(df, S)) = (L_S*B)(L_S*B)' with B the Bartlett factor,
 *   rand(MatrixNormal(M, U, V)) = M + chol(U).L * X * chol(V).U. */
static double mt_chi2(const orc_streams* s, uint32_t sweep, uint32_t which, double df) {
  /* Marsaglia & Tsang (2000) gamma(a = df/2, 1), a >= 1; chi2 = 2*gamma */
  double a = df / 2.0;
  double d = a - 1.0 / 3.0;
  double c = 1.0 / sqrt(9.0 * d);
  for (uint32_t attempt = 0;; ++attempt) {
    double x, unused;
    pgb_normal_pair(pgb_stream_block(s->seed, PGB_STREAM_MNIW, s->chain, sweep, which, 2 * attempt), &x, &unused);
    pgb_u32x4 b = pgb_stream_block(s->seed, PGB_STREAM_MNIW, s->chain, sweep, which, 2 * attempt + 1);
    double uu = pgb_u53(b.v[0], b.v[1]);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (pgb_log(uu) < 0.5 * x * x + d - d * v + d * pgb_log(v)) return 2.0 * (d * v);
  }
}

/* V_diag: the diagonal of V (the reference's examples); Vinv != NULL: I / V of a dense V (pgb_host_spd_inverse) */
static int mniw_sample_impl(int n_x, int n_phi, const double* Phi, const double* Psi, const double* Sigma,
                            const double* V_diag, const double* Vinv, const double* Lambda_Q, double ell_Q, int64_t Tm1,
                            const orc_streams* s, uint32_t sweep_index, double* A_out, double* Q_out) {
  int status = 0;
  size_t pp = (size_t)n_phi * (size_t)n_phi;
  double* Ls = (double*)malloc(sizeof(double) * pp);      /* chol(Sigbar) */
  double* Sinv = (double*)malloc(sizeof(double) * pp);    /* left_cov */
  double* Lv = (double*)malloc(sizeof(double) * pp);      /* chol(left_cov_sym) */
  double* M = (double*)malloc(sizeof(double) * (size_t)(n_x * n_phi));
  double* col = (double*)malloc(sizeof(double) * (size_t)n_phi);
  double* X = (double*)malloc(sizeof(double) * (size_t)(n_x * n_phi));
  double* LX = (double*)malloc(sizeof(double) * (size_t)(n_x * n_phi));
  double covM[ORC_MAX_Z * ORC_MAX_Z], Lc[ORC_MAX_Z * ORC_MAX_Z], S[ORC_MAX_Z * ORC_MAX_Z];
  double B[ORC_MAX_Z * ORC_MAX_Z], LB[ORC_MAX_Z * ORC_MAX_Z], Wm[ORC_MAX_Z * ORC_MAX_Z], Lq[ORC_MAX_Z * ORC_MAX_Z];

  /* Sigbar = Sigma + I / V   (:64) */
  memcpy(Ls, Sigma, sizeof(double) * pp);
  if (Vinv) {
    for (size_t e = 0; e < pp; ++e) Ls[e] = Ls[e] + Vinv[e];
  } else {
    for (int i = 0; i < n_phi; ++i) Ls[i + n_phi * i] = Ls[i + n_phi * i] + 1.0 / V_diag[i];
  }
  if (orc_cholesky_lower(Ls, n_phi)) { status = 16; goto done; }
  /* post_mean = Psibar / Sigbar  (:67): row d solves Sigbar m = psi_d */
  for (int d = 0; d < n_x; ++d) {
    for (int i = 0; i < n_phi; ++i) col[i] = Psi[d + n_x * i];
    fwd_subst(Ls, n_phi, col);
    bwd_subst_t(Ls, n_phi, col);
    for (int i = 0; i < n_phi; ++i) M[d + n_x * i] = col[i];
  }
  /* cov_M = Lambda_Q + Phibar - (Psibar/Sigbar)*Psibar'  (:65), symmetrised (:66) */
  for (int a = 0; a < n_x; ++a)
    for (int b = 0; b < n_x; ++b) {
      double acc = 0.0;
      for (int i = 0; i < n_phi; ++i) acc = fma(M[a + n_x * i], Psi[b + n_x * i], acc);
      covM[a + n_x * b] = (Lambda_Q[a + n_x * b] + Phi[a + n_x * b]) - acc;
    }
  for (int a = 0; a < n_x; ++a)
    for (int b = 0; b < n_x; ++b) Lc[a + n_x * b] = 0.5 * (covM[a + n_x * b] + covM[b + n_x * a]);
  /* left_cov = I / Sigbar (:68), symmetrised (:69) */
  spd_inverse_from_chol(Ls, n_phi, Sinv);
  for (int j = 0; j < n_phi; ++j)
    for (int i = 0; i < n_phi; ++i) Lv[i + n_phi * j] = 0.5 * (Sinv[i + n_phi * j] + Sinv[j + n_phi * i]);
  if (orc_cholesky_lower(Lv, n_phi)) { status = 16; goto done; }

  /* Q ~ InverseWishart(ell_Q + T*n_x, cov_M_sym)   (:72-73) */
  double df = ell_Q + (double)Tm1 * (double)n_x;
  if (orc_cholesky_lower(Lc, n_x)) { status = 16; goto done; }
  spd_inverse_from_chol(Lc, n_x, S); /* S = inv(cov_M_sym) */
  if (orc_cholesky_lower(S, n_x)) { status = 16; goto done; } /* S now holds L_S */
  memset(B, 0, sizeof(B));
  if (s->philox) {
    for (int i = 0; i < n_x; ++i) B[i + n_x * i] = sqrt(mt_chi2(s, sweep_index, (uint32_t)i, df - (double)i));
    int e = 0;
    for (int c = 0; c < n_x; ++c)
      for (int r = c + 1; r < n_x; ++r, ++e) {
        double z0, z1;
        pgb_normal_pair(pgb_stream_block(s->seed, PGB_STREAM_MNIW, s->chain, sweep_index, 200, (uint32_t)e), &z0, &z1);
        B[r + n_x * c] = z0;
      }
  } else {
    for (int c = 0; c < n_x; ++c)
      for (int r = c; r < n_x; ++r) B[r + n_x * c] = s->bartlett[r + n_x * c];
  }
  for (int r = 0; r < n_x; ++r) /* LB = L_S * B */
    for (int c = 0; c < n_x; ++c) {
      double acc = 0.0;
      for (int k = 0; k < n_x; ++k) acc = fma(S[r + n_x * k], B[k + n_x * c], acc);
      LB[r + n_x * c] = acc;
    }
  for (int r = 0; r < n_x; ++r) /* W = LB * LB' */
    for (int c = 0; c < n_x; ++c) {
      double acc = 0.0;
      for (int k = 0; k < n_x; ++k) acc = fma(LB[r + n_x * k], LB[c + n_x * k], acc);
      Wm[r + n_x * c] = acc;
    }
  if (orc_cholesky_lower(Wm, n_x)) { status = 16; goto done; }
  spd_inverse_from_chol(Wm, n_x, Q_out);

  /* A ~ MatrixNormal(post_mean, Q, left_cov_sym) = M + chol(Q).L * X * chol(V).U   (:76-77) */
  memcpy(Lq, Q_out, sizeof(double) * (size_t)(n_x * n_x));
  if (orc_cholesky_lower(Lq, n_x)) { status = 16; goto done; }
  if (s->philox) {
    int tot = n_x * n_phi;
    for (int p = 0; p < (tot + 1) / 2; ++p) {
      double z0, z1;
      pgb_normal_pair(pgb_stream_block(s->seed, PGB_STREAM_MNIW, s->chain, sweep_index, 300, (uint32_t)p), &z0, &z1);
      X[2 * p] = z0;
      if (2 * p + 1 < tot) X[2 * p + 1] = z1;
    }
  } else {
    memcpy(X, s->Xmn, sizeof(double) * (size_t)(n_x * n_phi));
  }
  for (int r = 0; r < n_x; ++r) /* LX = L_Q * X */
    for (int i = 0; i < n_phi; ++i) {
      double acc = 0.0;
      for (int k = 0; k <= r; ++k) acc = fma(Lq[r + n_x * k], X[k + n_x * i], acc);
      LX[r + n_x * i] = acc;
    }
  for (int r = 0; r < n_x; ++r) /* (LX * Lv')[r,i] = sum_{k<=i} LX[r,k] * Lv[i,k] */
    for (int i = 0; i < n_phi; ++i) {
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc = fma(LX[r + n_x * k], Lv[i + n_phi * k], acc);
      A_out[r + n_x * i] = M[r + n_x * i] + acc;
    }
done:
  free(Ls); free(Sinv); free(Lv); free(M); free(col); free(X); free(LX);
  return status;
}

int orc_mniw_sample(int n_x, int n_phi, const double* Phi, const double* Psi, const double* Sigma,
                    const double* V_diag, const double* Lambda_Q, double ell_Q, int64_t Tm1,
                    const orc_streams* s, uint32_t sweep_index, double* A_out, double* Q_out) {
  return mniw_sample_impl(n_x, n_phi, Phi, Psi, Sigma, V_diag, NULL, Lambda_Q, ell_Q, Tm1, s, sweep_index, A_out, Q_out);
}

/* I / V of a dense symmetric positive definite V (column-major n x n): 0 ok, 1 not positive definite */
int orc_spd_inverse(const double* V, int n, double* inv) {
  double* work = (double*)malloc(sizeof(double) * ((size_t)n * (size_t)n + (size_t)n));
  int rc = pgb_host_spd_inverse(V, n, inv, work);
  free(work);
  return rc;
}

/* MNIW_sample with a dense V (particle_Gibbs.jl:56 takes any V; :64 Sigbar = Sigma + I / V) */
int orc_mniw_sample_dense(int n_x, int n_phi, const double* Phi, const double* Psi, const double* Sigma,
                          const double* V, const double* Lambda_Q, double ell_Q, int64_t Tm1,
                          const orc_streams* s, uint32_t sweep_index, double* A_out, double* Q_out) {
  double* Vinv = (double*)malloc(sizeof(double) * (size_t)n_phi * (size_t)n_phi);
  int status = orc_spd_inverse(V, n_phi, Vinv) ? 16 : 0;
  if (!status)
    status = mniw_sample_impl(n_x, n_phi, Phi, Psi, Sigma, NULL, Vinv, Lambda_Q, ell_Q, Tm1, s, sweep_index, A_out, Q_out);
  free(Vinv);
  return status;
}

/* ------------------------------------------------------------------ full sampler */
static int particle_gibbs_impl(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y,
                               const double* C, const double* R, const double* x0_mean, const double* x0_chol,
                               const double* A_init, const double* Q_init, const double* V_diag, const double* Vinv,
                               const double* Lambda_Q, double ell_Q, int64_t K, int64_t K_b, int64_t k_d,
                               uint64_t seed, uint32_t chain, double* x_prim, double* A_s, double* Q_s, double* w_s,
                               double* x_s);

int orc_particle_gibbs(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y,
                       const double* C, const double* R, const double* x0_mean, const double* x0_chol,
                       const double* A_init, const double* Q_init, const double* V_diag,
                       const double* Lambda_Q, double ell_Q, int64_t K, int64_t K_b, int64_t k_d,
                       uint64_t seed, uint32_t chain, double* x_prim, double* A_s, double* Q_s, double* w_s,
                       double* x_s) {
  return particle_gibbs_impl(m, N, T, u, y, C, R, x0_mean, x0_chol, A_init, Q_init, V_diag, NULL, Lambda_Q, ell_Q, K, K_b,
                             k_d, seed, chain, x_prim, A_s, Q_s, w_s, x_s);
}

/* particle_Gibbs with a dense prior covariance V (n_phi x n_phi, column-major) */
int orc_particle_gibbs_dense(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y,
                             const double* C, const double* R, const double* x0_mean, const double* x0_chol,
                             const double* A_init, const double* Q_init, const double* V,
                             const double* Lambda_Q, double ell_Q, int64_t K, int64_t K_b, int64_t k_d,
                             uint64_t seed, uint32_t chain, double* x_prim, double* A_s, double* Q_s, double* w_s,
                             double* x_s) {
  double* Vinv = (double*)malloc(sizeof(double) * (size_t)m->n_phi * (size_t)m->n_phi);
  int status = orc_spd_inverse(V, m->n_phi, Vinv) ? 16 : 0;
  if (!status)
    status = particle_gibbs_impl(m, N, T, u, y, C, R, x0_mean, x0_chol, A_init, Q_init, NULL, Vinv, Lambda_Q, ell_Q, K,
                                 K_b, k_d, seed, chain, x_prim, A_s, Q_s, w_s, x_s);
  free(Vinv);
  return status;
}

static int particle_gibbs_impl(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y,
                               const double* C, const double* R, const double* x0_mean, const double* x0_chol,
                               const double* A_init, const double* Q_init, const double* V_diag, const double* Vinv,
                               const double* Lambda_Q, double ell_Q, int64_t K, int64_t K_b, int64_t k_d,
                               uint64_t seed, uint32_t chain, double* x_prim, double* A_s, double* Q_s, double* w_s,
                               double* x_s) {
  const int n_x = m->n_x, n_phi = m->n_phi;
  int64_t K_total = K_b + 1 + (K - 1) * (k_d + 1); /* :116 */
  int status = 0;
  double* A = (double*)malloc(sizeof(double) * (size_t)(n_x * n_phi));
  double* Q = (double*)malloc(sizeof(double) * (size_t)(n_x * n_x));
  double* Phi = (double*)malloc(sizeof(double) * (size_t)(n_x * n_x));
  double* Psi = (double*)malloc(sizeof(double) * (size_t)(n_x * n_phi));
  double* Sigma = (double*)malloc(sizeof(double) * (size_t)n_phi * (size_t)n_phi);
  double* w_last = (double*)malloc(sizeof(double) * (size_t)N);
  double* x_last = (double*)malloc(sizeof(double) * (size_t)(N * n_x));
  memcpy(A, A_init, sizeof(double) * (size_t)(n_x * n_phi));
  memcpy(Q, Q_init, sizeof(double) * (size_t)(n_x * n_x));
  orc_streams s;
  memset(&s, 0, sizeof(s));
  s.philox = 1;
  s.seed = seed;
  s.chain = chain;
  int64_t cur = 0;
  for (int64_t k = 1; k <= K_total; ++k) { /* :154 */
    status |= orc_sweep(m, N, T, u, y, C, R, x0_mean, x0_chol, A, Q, x_prim, k == 1, &s, (uint32_t)k, NULL, NULL,
                        NULL, w_last, x_last, Phi, Psi, Sigma, NULL);
    if (K_b < k && ((k - (K_b + 1)) % (k_d + 1) == 0) && cur < K) { /* :203 */
      memcpy(&A_s[cur * n_x * n_phi], A, sizeof(double) * (size_t)(n_x * n_phi));
      memcpy(&Q_s[cur * n_x * n_x], Q, sizeof(double) * (size_t)(n_x * n_x));
      if (w_s) memcpy(&w_s[cur * N], w_last, sizeof(double) * (size_t)N);
      if (x_s) memcpy(&x_s[cur * N * n_x], x_last, sizeof(double) * (size_t)(N * n_x));
      ++cur;
    }
    status |= mniw_sample_impl(n_x, n_phi, Phi, Psi, Sigma, V_diag, Vinv, Lambda_Q, ell_Q, T - 1, &s, (uint32_t)k, A, Q); /* :226 */
    if (status & 16) break;
  }
  free(A); free(Q); free(Phi); free(Psi); free(Sigma); free(w_last); free(x_last);
  return status;
}

/* ------------------------------------------------------------------ scenario rollout
 * optimal_control_Ipopt.jl:47-64 (x0), :67-74 (v), :77-83 (e), :114-125 (X_init, Y_init).
 * Streams: scenario k uses chain = k; purposes ROLL with t = 0 (star), 1 (x0 noise),
 * 2 (v, j = step*npairs+p), 3 (e). */
static int rollout_impl(const orc_model* m, int64_t K, int64_t n_models, int64_t H, int64_t N, const double* A,
                        const double* Q, const double* w_m1, const double* x_m1, const double* u_m1, const double* C,
                        const double* Le, const double* u_init, uint64_t seed, uint32_t k_offset, const double* x0_in,
                        const double* v_in, const double* e_in, double* x0, double* v, double* e, double* X, double* Y);

int orc_rollout(const orc_model* m, int64_t K, int64_t H, int64_t N, const double* A, const double* Q,
                const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                const double* u_init, uint64_t seed, uint32_t k_offset, double* x0, double* v, double* e, double* X,
                double* Y) {
  return rollout_impl(m, K, 0, H, N, A, Q, w_m1, x_m1, u_m1, C, Le, u_init, seed, k_offset, NULL, NULL, NULL, x0, v, e, X, Y);
}

/* The same with the x_vec_0 / v_vec / e_vec keyword arguments of solve_PG_OCP_Ipopt (optimal_control_Ipopt.jl:37):
 * `if x_vec_0 === nothing` (:48), `if v_vec === nothing` (:67), `if e_vec === nothing` (:77) — a supplied array replaces
 * the draw, everything else (:114-125) is unchanged.  This is the second entry of the pre-solve step (:88-112). */
int orc_rollout_supplied(const orc_model* m, int64_t K, int64_t H, int64_t N, const double* A, const double* Q,
                         const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                         const double* u_init, uint64_t seed, uint32_t k_offset, const double* x0_in, const double* v_in,
                         const double* e_in, double* x0, double* v, double* e, double* X, double* Y) {
  return rollout_impl(m, K, 0, H, N, A, Q, w_m1, x_m1, u_m1, C, Le, u_init, seed, k_offset, x0_in, v_in, e_in, x0, v, e, X, Y);
}

/* test_prediction (particle_Gibbs.jl:253-314): every one of the K models is simulated k_n times over the test
 * input.  Simulation (k, kn) is scenario s = k + K*kn of a rollout over K*k_n scenarios that reads model s mod K
 * (the order of the reference's reshape (n_y, T_test, K*k_n) at :301-302, k fastest): star ~ w_m1 (:288),
 * x_1 = f(x_m1[:, star], u_m1) + v (:290), x_t = f(x_{t-1}, u_test[:, t-1]) + v (:295), y_t = g(x_t) + e (:297) -
 * i.e. orc_rollout with H = T_test and u_init = u_test: x_loop[:, t] = X[t-1], y_loop[:, t] = Y[t-1] (the reference
 * leaves x_loop[:, T_test+1] undefined; here it holds one more transition).
 * mean_rmse = sqrt(sum((y_sim - y_test)^2) / (n_y*T_test*K*k_n)), errors matched in TIME (y_sim[:, t, s] against
 * y_test[:, t]), the sum taken by the balanced tree over the column-major (n_y, T_test, K*k_n) array.  DEVIATION from
 * the Julia source, stated: :311 subtracts repeat(y_test', 1, 1, K*k_n), an array of shape (T_test, n_y, K*k_n), from
 * the (n_y, T_test, K*k_n) simulations; for n_y = 1 Julia broadcasts that to T_test x T_test x S cross terms (every
 * simulated output against every test output) and for n_y > 1 it throws.  The time-matched value is the quantity the
 * MATLAB twin computes (MATLAB/src/test_prediction.m) and the one a prediction RMSE means; the drop-in therefore
 * prints a different number than the Julia reference does. */
int orc_test_prediction(const orc_model* m, int64_t K, int64_t k_n, int64_t T_test, int64_t N, const double* A,
                        const double* Q, const double* w_m1, const double* x_m1, const double* u_m1, const double* C,
                        const double* Le, const double* u_test, const double* y_test, uint64_t seed, double* x_sim,
                        double* y_sim, double* mean_rmse) {
  const int64_t S = K * k_n, M = S * T_test * m->n_y;
  int st = rollout_impl(m, S, K, T_test, N, A, Q, w_m1, x_m1, u_m1, C, Le, u_test, seed, 0, NULL, NULL, NULL, NULL, NULL, NULL, x_sim, y_sim);
  double* sq = (double*)malloc(sizeof(double) * (size_t)M);
  for (int64_t s2 = 0; s2 < S; ++s2)
    for (int64_t t = 0; t < T_test; ++t)
      for (int o = 0; o < m->n_y; ++o) {
        const int64_t i = (s2 * T_test + t) * m->n_y + o;
        const double d = y_sim[i] - y_test[t * m->n_y + o];
        sq[i] = d * d;
      }
  *mean_rmse = sqrt(orc_pairwise_sum(sq, M) / (double)M);
  free(sq);
  return st;
}

static int rollout_impl(const orc_model* m, int64_t K, int64_t n_models, int64_t H, int64_t N, const double* A,
                        const double* Q, const double* w_m1, const double* x_m1, const double* u_m1, const double* C,
                        const double* Le, const double* u_init, uint64_t seed, uint32_t k_offset, const double* x0_in,
                        const double* v_in, const double* e_in, double* x0, double* v, double* e, double* X, double* Y) {
  const int n_x = m->n_x, n_u = m->n_u, n_y = m->n_y, n_phi = m->n_phi;
  int status = 0;
  double* phi_tmp = (double*)malloc(sizeof(double) * (size_t)n_phi);
  double Lq[ORC_MAX_Z * ORC_MAX_Z], z[ORC_MAX_Z], f[ORC_MAX_Z];
  const double* Lr = Le; /* lower factor the measurement noise is unwhitened with (see header) */
  double *x0_own = NULL, *v_own = NULL, *e_own = NULL; /* optional outputs */
  if (!x0) x0 = x0_own = (double*)malloc(sizeof(double) * (size_t)(K * n_x));
  if (!v) v = v_own = (double*)malloc(sizeof(double) * (size_t)(K * H * n_x));
  if (!e) e = e_own = (double*)malloc(sizeof(double) * (size_t)(K * H * n_y));
  for (int64_t k = 0; k < K && !(status & 16); ++k) {
    const int64_t km = n_models > 0 ? k % n_models : k; /* model (PG sample) scenario k simulates */
    const double* Ak = &A[km * n_x * n_phi];
    orc_streams s;
    memset(&s, 0, sizeof(s));
    s.philox = 1;
    s.seed = seed;
    s.chain = k_offset + (uint32_t)k; /* scenario id: a shard of the K scenarios keeps its global stream */
    if (!(x0_in && v_in)) { /* Q is only used by the draws (:55, :71) */
      memcpy(Lq, &Q[km * n_x * n_x], sizeof(double) * (size_t)(n_x * n_x));
      if (orc_cholesky_lower(Lq, n_x)) { status |= 16; break; }
    }
    if (x0_in) { /* :48 */
      for (int d = 0; d < n_x; ++d) X[(k * (H + 1) + 0) * n_x + d] = x0[k * n_x + d] = x0_in[k * n_x + d];
    } else {
      /* star = systematic_resampling(w_m1, 1); x_m1 = x_m1[:, star]   (:58-59) */
      int64_t star;
      double us = stream_uniform(&s, PGB_STREAM_ROLL, 0, 0);
      status |= orc_systematic_resampling(&w_m1[km * N], N, 1, us, &star);
      if (star < 1) star = 1;
      /* x0 = f(x_m1, u_m1) + rand(mvn)   (:62) */
      eval_f_blocked(m, Ak, &x_m1[(km * N + (star - 1)) * n_x], &u_m1[km * n_u], phi_tmp, f);
      stream_normals(&s, PGB_STREAM_ROLL, 0, 1, 0, n_x, z);
      for (int d = 0; d < n_x; ++d) {
        double acc = 0.0;
        for (int e2 = 0; e2 <= d; ++e2) acc = fma(Lq[d + n_x * e2], z[e2], acc);
        x0[k * n_x + d] = f[d] + acc;
        X[(k * (H + 1) + 0) * n_x + d] = x0[k * n_x + d];
      }
    }
    for (int64_t t = 0; t < H; ++t) {
      if (v_in) { /* :67 */
        for (int d = 0; d < n_x; ++d) v[(k * H + t) * n_x + d] = v_in[(k * H + t) * n_x + d];
      } else { /* v_vec[:, :, k] = rand(mvn_v, H)  (:72) */
        stream_normals(&s, PGB_STREAM_ROLL, 0, 2, t, n_x, z);
        for (int d = 0; d < n_x; ++d) {
          double acc = 0.0;
          for (int e2 = 0; e2 <= d; ++e2) acc = fma(Lq[d + n_x * e2], z[e2], acc);
          v[(k * H + t) * n_x + d] = acc;
        }
      }
      if (e_in) { /* :77 */
        for (int o = 0; o < n_y; ++o) e[(k * H + t) * n_y + o] = e_in[(k * H + t) * n_y + o];
      } else {
        stream_normals(&s, PGB_STREAM_ROLL, 0, 3, t, n_y, z); /* e_vec (:81) */
        for (int o = 0; o < n_y; ++o) {
          double acc = 0.0;
          for (int e2 = 0; e2 <= o; ++e2) acc = fma(Lr[o + n_y * e2], z[e2], acc);
          e[(k * H + t) * n_y + o] = acc;
        }
      }
    }
    for (int64_t t = 0; t < H; ++t) { /* :121-122 */
      const double* xt = &X[(k * (H + 1) + t) * n_x];
      eval_f_blocked(m, Ak, xt, &u_init[n_u * t], phi_tmp, f);
      for (int d = 0; d < n_x; ++d) X[(k * (H + 1) + t + 1) * n_x + d] = f[d] + v[(k * H + t) * n_x + d];
      for (int o = 0; o < n_y; ++o) {
        double g = 0.0;
        for (int d = 0; d < n_x; ++d) g = fma(C[o + n_y * d], xt[d], g);
        Y[(k * H + t) * n_y + o] = g + e[(k * H + t) * n_y + o];
      }
    }
  }
  free(phi_tmp);
  free(x0_own); free(v_own); free(e_own);
  return status;
}

/* x_T = x_m1[:, star], star = systematic_resampling(w_m1, 1) (optimal_control_Ipopt.jl:58-59) for ONE sample, with the
 * uniform of scenario stream `cid` — the first draw rollout_impl makes for that scenario.  This is what the owning GPU
 * computes before the packed (A, Q, x_T) records are gathered. */
int orc_draw_xT(int64_t N, int n_x, const double* w_m1, const double* x_m1 /* N x n_x */, uint64_t seed, uint32_t cid,
                double* xT) {
  orc_streams s;
  memset(&s, 0, sizeof(s));
  s.philox = 1;
  s.seed = seed;
  s.chain = cid;
  int64_t star;
  double us = stream_uniform(&s, PGB_STREAM_ROLL, 0, 0);
  int status = orc_systematic_resampling(w_m1, N, 1, us, &star);
  if (star < 1) star = 1;
  for (int d = 0; d < n_x; ++d) xT[d] = x_m1[(star - 1) * n_x + d];
  return status;
}

/* ---- one step of the stacked K-scenario model and its Jacobians ---------------------------
 * RobotDynamics.discrete_dynamics[!] of PGS_model_obj (optimal_control_Altro.jl:79-107):
 *   x_vec_n[k] = PG_samples[k].A * phi(x_vec[k], u) + v_vec[:, t+1, k]          (t: 0-based time index)
 * and the Jacobians RobotDynamics.@autodiff (:51) obtains by forward-mode differentiation of exactly that expression:
 *   d x_vec_n[k] / d x_vec[k] = A_k * dphi/dx  (block diagonal, n_x x n_x per scenario),   d x_vec_n[k] / d u = A_k * dphi/du.
 * phi_variant 0: the rounding of phi_sampling (orc_phi); 1: the rounding of the examples' phi_opt
 * (examples/PG_OCP_generic_basis_functions_Altro.jl:204-217): arg = (pi*j*(z + L)) / (2L).  Derivatives are the analytic
 * ones: d/dz [lsi*sin(arg)] = lsi * (cos(arg) * w) with w = (pi*j)/(2L); every product over the augmented dimensions
 * runs left to right (k = 0 .. n_z-1) with the differentiated factor in its place.  A_k * (.) uses the rollout's blocked
 * order (eval_f_blocked): 128 runs of consecutive columns, fma in increasing i, balanced tree over the runs — so x_next
 * with variant 0 equals the X[t+1] the rollout computes from the same state, bit for bit.  Against ForwardDiff the
 * Jacobians agree to a few ulp (different association), which is the tolerance the tests state. */
static void phi_and_jac(const orc_model* m, const double* x, const double* u, int variant, double* phi, double* dphi /* n_z x n_phi */) {
  const int n_z = m->n_u + m->n_x, n_phi = m->n_phi;
  if (m->basis == ORC_BASIS_KNOWN_VB) {
    /* z = [u; x1; x2]; rows of dphi: d/du, d/dx1, d/dx2 */
    const double c3 = pgb_cos(3.0 * x[0]), s3 = pgb_sin(3.0 * x[0]), s2 = pgb_sin(2.0 * x[1]), c2 = pgb_cos(2.0 * x[1]);
    phi[0] = 0.1 * x[0]; phi[1] = 0.1 * x[1]; phi[2] = u[0]; phi[3] = (0.01 * c3) * x[1]; phi[4] = (0.1 * s2) * u[0];
    for (int i = 0; i < 3 * 5; ++i) dphi[i] = 0.0;
    dphi[0 + 3 * 2] = 1.0;                         /* d phi_3 / du */
    dphi[0 + 3 * 4] = 0.1 * s2;                    /* d phi_5 / du */
    dphi[1 + 3 * 0] = 0.1;                         /* d phi_1 / dx1 */
    dphi[1 + 3 * 3] = (0.01 * (-(s3 * 3.0))) * x[1]; /* d phi_4 / dx1 */
    dphi[2 + 3 * 1] = 0.1;                         /* d phi_2 / dx2 */
    dphi[2 + 3 * 3] = 0.01 * c3;                   /* d phi_4 / dx2 */
    dphi[2 + 3 * 4] = (0.1 * (c2 * 2.0)) * u[0];   /* d phi_5 / dx2 */
    return;
  }
  double z[ORC_MAX_Z];
  for (int k = 0; k < m->n_u; ++k) z[k] = u[k];
  for (int k = 0; k < m->n_x; ++k) z[m->n_u + k] = x[k];
  const double PI = 3.141592653589793;
  double tk[ORC_MAX_Z][16], dk[ORC_MAX_Z][16];
  for (int k = 0; k < n_z; ++k)
    for (int j = 0; j < m->count[k]; ++j) {
      const double jj = (double)(j + 1), w = (PI * jj) / (2.0 * m->L[k]), lsi = 1.0 / sqrt(m->L[k]);
      const double arg = variant ? ((PI * jj) * (z[k] + m->L[k])) / (2.0 * m->L[k]) : w * (z[k] + m->L[k]);
      tk[k][j] = lsi * pgb_sin(arg);
      dk[k][j] = lsi * (pgb_cos(arg) * w);
    }
  for (int i = 0; i < n_phi; ++i) {
    int dg[ORC_MAX_Z];
    for (int k = 0; k < n_z; ++k) dg[k] = (i / m->stride[k]) % m->count[k];
    double p = 1.0;
    for (int k = 0; k < n_z; ++k) p = p * tk[k][dg[k]];
    phi[i] = p;
    for (int mm = 0; mm < n_z; ++mm) {
      double q = 1.0;
      for (int k = 0; k < n_z; ++k) q = q * (k == mm ? dk[k][dg[k]] : tk[k][dg[k]]);
      dphi[mm + n_z * i] = q;
    }
  }
}
/* out = A (n_x x n_phi, column-major) * vec, in the rollout's blocked order; vec read with stride */
static double blocked_dot(const double* A, int n_x, int d, const double* vec, int stride, int n_phi) {
  double part[ORC_ROLL_LANES];
  const int c = (n_phi + ORC_ROLL_LANES - 1) / ORC_ROLL_LANES;
  for (int j = 0; j < ORC_ROLL_LANES; ++j) {
    double acc = 0.0;
    for (int i = j * c; i < (j + 1) * c && i < n_phi; ++i) acc = fma(A[d + n_x * i], vec[(int64_t)stride * i], acc);
    part[j] = acc;
  }
  for (int w = 1; w < ORC_ROLL_LANES; w <<= 1)
    for (int j = 0; j < ORC_ROLL_LANES; j += 2 * w) part[j] = part[j] + part[j + w];
  return part[0];
}
int orc_dynamics(const orc_model* m, int64_t K, const double* A /* K x n_x x n_phi col-major */, const double* x_vec /* K x n_x */,
                 const double* u, const double* v /* K x H x n_x or NULL */, int64_t H, int64_t t, int variant,
                 double* x_next /* K x n_x */, double* dfdx /* K x (n_x x n_x col-major) or NULL */,
                 double* dfdu /* K x (n_x x n_u col-major) or NULL */) {
  const int n_x = m->n_x, n_u = m->n_u, n_phi = m->n_phi, n_z = n_x + n_u;
  double* phi = (double*)malloc(sizeof(double) * (size_t)n_phi);
  double* dphi = (double*)malloc(sizeof(double) * (size_t)n_phi * (size_t)n_z);
  for (int64_t k = 0; k < K; ++k) {
    const double* Ak = &A[k * n_x * n_phi];
    phi_and_jac(m, &x_vec[k * n_x], u, variant, phi, dphi);
    for (int d = 0; d < n_x; ++d) {
      const double f = blocked_dot(Ak, n_x, d, phi, 1, n_phi);
      x_next[k * n_x + d] = v ? f + v[(k * H + t) * n_x + d] : f;
      if (dfdu)
        for (int c = 0; c < n_u; ++c) dfdu[k * n_x * n_u + d + n_x * c] = blocked_dot(Ak, n_x, d, dphi + c, n_z, n_phi);
      if (dfdx)
        for (int c = 0; c < n_x; ++c) dfdx[k * n_x * n_x + d + n_x * c] = blocked_dot(Ak, n_x, d, dphi + n_u + c, n_z, n_phi);
    }
  }
  free(phi); free(dphi);
  return 0;
}

/* ---- scenario screening (optimal_control_Ipopt.jl:334-338; h_scenario of the examples, ..._Ipopt.jl:198-214) ---------- */
/* Julia's isless on Float64: NaN is the largest value, -0.0 sorts before 0.0. */
static int orc_isless(double a, double b) {
  if (isnan(a)) return 0;
  if (isnan(b)) return 1;
  if (a == 0.0 && b == 0.0) return signbit(a) && !signbit(b);
  return a < b;
}

int orc_scenario_constraint_max(int64_t K, int64_t H, int32_t n_y, const double* Y, const double* y_min,
                                const double* y_max, double* h_max, int64_t* order) {
  int64_t n_fin = 0;
  for (int64_t i = 0; i < (int64_t)n_y * H; ++i) n_fin += (isfinite(y_min[i]) != 0) + (isfinite(y_max[i]) != 0);
  if (n_fin == 0) return 1;
  for (int64_t k = 0; k < K; ++k) {
    double m = -INFINITY;
    int nan = 0;
    for (int64_t t = 0; t < H; ++t)      /* loop order of h_scenario: t outer, n inner, min before max */
      for (int32_t n = 0; n < n_y; ++n) {
        double y = Y[(k * H + t) * n_y + n];
        if (isfinite(y_min[n + (int64_t)n_y * t])) { double h = y_min[n + (int64_t)n_y * t] - y; if (isnan(h)) nan = 1; else if (h > m || (h == m && !signbit(h))) m = h; }
        if (isfinite(y_max[n + (int64_t)n_y * t])) { double h = y - y_max[n + (int64_t)n_y * t]; if (isnan(h)) nan = 1; else if (h > m || (h == m && !signbit(h))) m = h; }
      }
    h_max[k] = nan ? NAN : m;
  }
  if (order) { /* stable insertion of each element: sortperm(...; rev=true) keeps ties in their original order */
    for (int64_t k = 0; k < K; ++k) {
      int64_t j = k;
      while (j > 0 && orc_isless(h_max[order[j - 1] - 1], h_max[k])) { order[j] = order[j - 1]; --j; }
      order[j] = k + 1;
    }
  }
  return 0;
}
