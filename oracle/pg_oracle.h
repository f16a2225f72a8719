/* pg_oracle.h — CPU restatement of PGopt's particle-Gibbs hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a plain, single-threaded C restatement of
 *   Julia/src/particle_Gibbs.jl:17-34    systematic_resampling
 *   Julia/src/particle_Gibbs.jl:56-81    MNIW_sample
 *   Julia/src/particle_Gibbs.jl:114-237  particle_Gibbs (CPF-AS sweep, trace, statistics)
 *   Julia/src/optimal_control_Ipopt.jl:47-83,114-125   scenario sampling + forward rollout
 * and of the two basis definitions in
 *   Julia/examples/PG_OCP_known_basis_functions_Altro.jl:34
 *   Julia/examples/PG_OCP_generic_basis_functions_Altro.jl:52-99
 * (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * PARITY UNPINNED: the reference has no tests, golden vectors or fixtures, and neither Julia nor
 * MATLAB exists in the build image or on the GPU box, so this restatement cannot be checked
 * against outputs of the reference itself.  It is pinned only by hand-traced known-answer cases
 * derived from the reference source (tests/test_oracle_kat.py) and by libm cross-checks.  The
 * arithmetic the reference delegates to Distributions.jl / OpenBLAS / Julia Base (summation order
 * of sum(), gemm order, exp/sin/cos last bits, LU vs Cholesky) is not pinned by anything upstream;
 * every such order is fixed HERE and documented next to the code.
 */
#ifndef PG_ORACLE_H
#define PG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_BASIS_KNOWN_VB = 0, ORC_BASIS_HILBERT_GP = 1 };
#define ORC_MAX_Z 8

typedef struct {
  int32_t basis;
  int32_t n_x, n_u, n_y, n_phi;
  /* Hilbert-space GP basis over z = [u; x] (n_z = n_u + n_x dims).  For augmented dimension k,
   * basis row i (0-based) uses the 1-based frequency index j = (i / stride[k]) % count[k] + 1. */
  int32_t count[ORC_MAX_Z];
  int32_t stride[ORC_MAX_Z];
  double L[ORC_MAX_Z];
} orc_model;

typedef struct {
  int32_t philox;      /* 1: counter-based streams (seed, chain); 0: injected arrays below */
  uint64_t seed;
  uint32_t chain;
  /* injected streams for ONE sweep (layout = order of consumption in the reference) */
  const double* Z0;    /* [n_x*(N-1)]  initial-state normals, particle-major           :163-165 */
  const double* Ures;  /* [T]  index t (0-based step t>=1 used)                         :172|185 */
  const double* Z;     /* [T][n_x*N]  step t (>=1) normals, particle-major (N-1 or N)   :175|188 */
  const double* Uanc;  /* [T]                                                          :181 */
  double Ustar;        /*                                                              :213 */
  const double* bartlett; /* [n_x*n_x] lower-triangular Bartlett factor, column-major   :73 */
  const double* Xmn;      /* [n_x*n_phi] matrix-normal normals, column-major            :77 */
} orc_streams;

/* math helpers (vectorised wrappers over pgb_math.h, for tests) */
void orc_vexp(const double* x, double* y, int64_t n);
void orc_vlog(const double* x, double* y, int64_t n);
void orc_vsin(const double* x, double* y, int64_t n);
void orc_vcos(const double* x, double* y, int64_t n);
void orc_vnormal_pairs(uint64_t seed, uint32_t purpose, uint32_t chain, uint32_t sweep, uint32_t t,
                       int64_t npairs, double* z);
int orc_selftest(void);

double orc_pairwise_sum(const double* v, int64_t n);

/* returns status bits: 1 the reference would raise BoundsError (q_N < u_i), 2 an index 0 was returned,
 * 4 a normalised weight is NaN/Inf/negative (the reference carries on silently) */
int orc_systematic_resampling(const double* W, int64_t n_w, int64_t n_draw, double rnd,
                              int64_t* idx /* 1-based */);

void orc_phi(const orc_model* m, const double* x, const double* u, double* phi);

int orc_cholesky_lower(double* A, int n); /* in place, column-major, upper part zeroed */

/* One PG sweep (particle_Gibbs.jl:155-226 for a single k).  `first` selects the k==1 branch.
 * Histories are caller-allocated (may be NULL): a[T*N] (1-based, row t; row 0 unused),
 * x_pf[T][n_x][N]... see pg_oracle.c for layouts.  x_prim is updated in place. */
int orc_sweep(const orc_model* m, int64_t N, int64_t T, const double* u /*n_u x T*/,
              const double* y /*n_y x T*/, const double* C /*n_y x n_x*/, const double* R /*n_y x n_y*/,
              const double* x0_mean, const double* x0_chol /*lower n_x x n_x*/,
              const double* A, const double* Q, double* x_prim /*n_x x T*/, int first,
              const orc_streams* s, uint32_t sweep_index,
              int64_t* a_hist /*T x N or NULL*/, double* x_hist /*T x n_x x N or NULL*/,
              double* w_hist /*T x N or NULL*/, double* w_last /*N*/, double* x_last /*n_x x N*/,
              double* Phi, double* Psi, double* Sigma, int64_t* star_out);

int orc_mniw_sample(int n_x, int n_phi, const double* Phi, const double* Psi, const double* Sigma,
                    const double* V_diag, const double* Lambda_Q, double ell_Q, int64_t Tm1,
                    const orc_streams* s, uint32_t sweep_index, double* A_out, double* Q_out);

/* Dense prior covariance V (n_phi x n_phi, column-major, symmetric positive definite; particle_Gibbs.jl:56 takes any V):
 * Sigbar = Sigma + I / V (:64) with I / V = orc_spd_inverse(V) (Cholesky + column solves, pgb_math.h). */
int orc_spd_inverse(const double* V, int n, double* inv);
int orc_mniw_sample_dense(int n_x, int n_phi, const double* Phi, const double* Psi, const double* Sigma,
                          const double* V, const double* Lambda_Q, double ell_Q, int64_t Tm1,
                          const orc_streams* s, uint32_t sweep_index, double* A_out, double* Q_out);
int orc_particle_gibbs_dense(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y,
                             const double* C, const double* R, const double* x0_mean, const double* x0_chol,
                             const double* A_init, const double* Q_init, const double* V,
                             const double* Lambda_Q, double ell_Q, int64_t K, int64_t K_b, int64_t k_d,
                             uint64_t seed, uint32_t chain, double* x_prim, double* A_s, double* Q_s, double* w_s,
                             double* x_s);

/* orc_sweep in the F32 precision mode (PGB_PRECISION_F32; see pg_oracle.c): FP32 particle states / basis functions /
 * noise / log-weights / exponentials with the bit-defined functions of pgb_mathf.h, FP64 normalisation and cumulative
 * sums.  Same arguments; the history arrays hold FP32 values widened to FP64. */
int orc_sweep_f32(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y, const double* C,
                  const double* R, const double* x0_mean, const double* x0_chol, const double* A, const double* Q,
                  double* x_prim, int first, const orc_streams* s, uint32_t sweep_index, int64_t* a_hist,
                  double* x_hist, double* w_hist, double* w_last, double* x_last, double* Phi, double* Psi,
                  double* Sigma, int64_t* star_out);
/* FP32 math wrappers for tests/test_mathf.py: fn 0 exp, 1 log (positive normal), 2 sin, 3 cos */
void orc_vmathf(int fn, const float* x, float* y, int64_t n);
void orc_vnormal_pairs_f32(uint64_t seed, uint32_t purpose, uint32_t chain, uint32_t sweep, uint32_t t, int64_t n,
                           int n_x, float* z /* n x n_x */);

/* Full sampler (particle_Gibbs.jl:114-237) with Philox streams. Outputs K samples. */
int orc_particle_gibbs(const orc_model* m, int64_t N, int64_t T, const double* u, const double* y,
                       const double* C, const double* R, const double* x0_mean, const double* x0_chol,
                       const double* A_init, const double* Q_init, const double* V_diag,
                       const double* Lambda_Q, double ell_Q, int64_t K, int64_t K_b, int64_t k_d,
                       uint64_t seed, uint32_t chain, double* x_prim /* n_x x T in/out */,
                       double* A_s /*K x n_x x n_phi*/, double* Q_s, double* w_s /*K x N*/,
                       double* x_s /*K x n_x x N*/);

/* Scenario sampling + rollout (optimal_control_Ipopt.jl:47-83,114-125).
 * Le = lower factor L with e = L*z.  NOTE (reference quirk, kept): the reference builds
 * MvNormal(zeros(n_y), R) (:79); with the examples' *scalar* R = 0.1 Distributions.jl 0.25 reads a
 * scalar second argument as a standard deviation, so e = R*z (Le = [R]); with a matrix R it is a
 * covariance (Le = chol(R).L).  The caller (shim) makes that choice and passes Le. */
int orc_rollout(const orc_model* m, int64_t K, int64_t H, int64_t N, const double* A /*K x n_x x n_phi*/,
                const double* Q /*K x n_x x n_x*/, const double* w_m1 /*K x N*/,
                const double* x_m1 /*K x n_x x N*/, const double* u_m1 /*K x n_u*/,
                const double* C, const double* Le, const double* u_init /*n_u x H*/,
                uint64_t seed, uint32_t k_offset /* global id of scenario 0 (stream id) */, double* x0 /*K x n_x*/,
                double* v /*K x H x n_x*/,
                double* e /*K x H x n_y*/, double* X /*K x (H+1) x n_x*/, double* Y /*K x H x n_y*/);

/* orc_rollout with the reference's x_vec_0 / v_vec / e_vec keyword arguments (optimal_control_Ipopt.jl:48, :67, :77):
 * a non-NULL x0_in / v_in / e_in (layouts of x0 / v / e) replaces the corresponding draw. */
int orc_rollout_supplied(const orc_model* m, int64_t K, int64_t H, int64_t N, const double* A, const double* Q,
                         const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                         const double* u_init, uint64_t seed, uint32_t k_offset, const double* x0_in, const double* v_in,
                         const double* e_in, double* x0, double* v, double* e, double* X, double* Y);

/* x_T = x_m1[:, star] with star = systematic_resampling(w_m1, 1) (optimal_control_Ipopt.jl:58-59), scenario stream cid */
int orc_draw_xT(int64_t N, int n_x, const double* w_m1, const double* x_m1 /* N x n_x */, uint64_t seed, uint32_t cid,
                double* xT);

/* One step of the stacked scenario model and its Jacobians (optimal_control_Altro.jl:79-107, @autodiff :51); see pg_oracle.c.
 * v: K x H x n_x (the rollout's layout) or NULL (no noise); t: 0-based time index; variant 0 phi_sampling, 1 phi_opt rounding. */
int orc_dynamics(const orc_model* m, int64_t K, const double* A, const double* x_vec, const double* u, const double* v,
                 int64_t H, int64_t t, int variant, double* x_next, double* dfdx, double* dfdu);

/* test_prediction (particle_Gibbs.jl:253-314): K models x k_n simulations over the test input, simulation (k, kn) =
 * scenario k + K*kn with its own stream; x_sim K*k_n x (T_test+1) x n_x, y_sim K*k_n x T_test x n_y (scenario-major),
 * mean_rmse = the time-matched RMSE (NOT the cross-term broadcast the Julia line :311 evaluates; see pg_oracle.c). */
int orc_test_prediction(const orc_model* m, int64_t K, int64_t k_n, int64_t T_test, int64_t N, const double* A,
                        const double* Q, const double* w_m1, const double* x_m1, const double* u_m1, const double* C,
                        const double* Le, const double* u_test /*n_u x T_test*/, const double* y_test /*n_y x T_test*/,
                        uint64_t seed, double* x_sim, double* y_sim, double* mean_rmse);

/* Scenario screening of the greedy support-sub-sample search (optimal_control_Ipopt.jl:334-338, Altro.jl:496-500) for the
 * examples' box output constraints (examples/PG_OCP_known_basis_functions_Ipopt.jl:198-214): h_max[k] = maximum over the
 * finite entries (n, t) of y_min[n,t] - y[n,t,k] and y[n,t,k] - y_max[n,t] (NaN propagates like Julia's maximum), then
 * order = sortperm(h_max; rev=true) (stable, isless ordering), 1-based.  Y is scenario-major K x H x n_y; y_min / y_max
 * are n_y x H column-major.  Returns 1 when no bound is finite (maximum of an empty collection throws in Julia). */
int orc_scenario_constraint_max(int64_t K, int64_t H, int32_t n_y, const double* Y, const double* y_min,
                                const double* y_max, double* h_max, int64_t* order);

#ifdef __cplusplus
}
#endif
#endif
