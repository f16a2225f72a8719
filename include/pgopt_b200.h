/* pgopt_b200.h — C ABI of libpgopt_b200.so: the B200-native (sm_100a) replacement for the
 * data-parallel hot path of TUM-ITR/PGopt.
 *
 * The reference has no FFI/plugin boundary: its boundary is the Julia function signatures
 *   particle_Gibbs(u, y, K, K_b, k_d, N, phi, Lambda_Q, ell_Q, Q_init, V, A_init, x_init_dist, g, R; x_prim)
 *                                                   Julia/src/particle_Gibbs.jl:114
 *   systematic_resampling(W, N)                     Julia/src/particle_Gibbs.jl:17
 *   MNIW_sample(Phi, Psi, Sigma, V, Lambda_Q, ell_Q, T)   Julia/src/particle_Gibbs.jl:56
 *   solve_PG_OCP_Ipopt(...; x_vec_0, v_vec, e_vec, u_init)  Julia/src/optimal_control_Ipopt.jl:37
 * A thin `ccall` shim (julia/PGoptB200.jl) keeps those signatures and calls the
 * entry points below.  Plain pointers and sizes only; all matrices are column-major FP64 exactly as
 * Julia passes them; indices returned to the caller are 1-based Int64 like the reference's.
 * Every function returns 0 on success or a PGB_ERR_* code; pgb_last_error() gives the message.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * PGB_ERR_CUDA.
 *
 * Ownership: the library owns all device memory behind a handle; host buffers are borrowed for
 * the duration of a call.  A handle is bound to one device and one stream and is not re-entrant;
 * distinct handles may be driven from distinct host threads.
 */
#ifndef PGOPT_B200_H
#define PGOPT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGB_MAX_Z 8 /* max augmented dimension n_u + n_x of the Hilbert-space basis */

enum {
  PGB_OK = 0,
  PGB_ERR_INVALID = 1,     /* bad argument / unsupported configuration */
  PGB_ERR_CUDA = 2,        /* CUDA runtime error (no device, OOM, launch failure) */
  PGB_ERR_NOT_POSDEF = 3,  /* Cholesky failed (Julia: PosDefException) */
  PGB_ERR_STATE = 4        /* call sequence error (data/prior/state not set) */
};

/* basis-function families: the C ABI cannot take Julia closures, so phi is a descriptor */
enum {
  PGB_BASIS_KNOWN_VB = 0,  /* examples/PG_OCP_known_basis_functions_Altro.jl:34  (n_x=2,n_u=1,n_phi=5) */
  PGB_BASIS_HILBERT_GP = 1 /* examples/PG_OCP_generic_basis_functions_Altro.jl:52-99 (phi_sampling) */
};

/* status bits reported by pgb_get_status (sticky per handle until cleared) */
enum {
  PGB_STATUS_RESAMPLE_OVERRUN = 1, /* q_N < u_i: the reference raises BoundsError at :28 */
  PGB_STATUS_INDEX_ZERO = 2,       /* u_1 <= 0: the reference returns index 0 */
  PGB_STATUS_NONFINITE = 4,        /* NaN/Inf weight or weight sum */
  PGB_STATUS_NOT_POSDEF = 16
};

enum { PGB_X_HISTORY_AUTO = 0, PGB_X_HISTORY_STORE = 1, PGB_X_HISTORY_REPLAY = 2 };
/* Arithmetic of the particle filter.  F64: everything in FP64 (the reference).  F32 ("FP32 particle state" of the
 * north star): particle states, basis functions, A*phi, process noise, log-weights and the two exponentials per
 * particle in FP32 — bit-defined FP32 elementary functions shared with the oracle (pgb_mathf.h), so indices stay
 * bit-exact against the oracle's F32 mode — while weight normalisation, the cumulative sums of the two systematic
 * resampling draws, the backward trace, statistics and the MNIW draw stay FP64.  8 n_x + 12 bytes per particle-step. */
enum { PGB_PRECISION_F64 = 0, PGB_PRECISION_F32 = 1 };

typedef struct {
  int32_t device;     /* CUDA device ordinal */
  int32_t n_x, n_u, n_y, n_phi;
  int32_t basis;      /* PGB_BASIS_* */
  /* Hilbert-space basis over z = [u; x]: for augmented dimension k, basis row i (0-based) uses the
   * 1-based frequency j = (i / hg_stride[k]) % hg_count[k] + 1; strides must be the row-major
   * mixed radix of hg_count (dimension 0 slowest), which is what the reference constructs
   * (examples/PG_OCP_generic_basis_functions_Altro.jl:59-82). */
  int32_t hg_count[PGB_MAX_Z];
  int32_t hg_stride[PGB_MAX_Z];
  double hg_L[PGB_MAX_Z];
  int64_t N;          /* particles (slot N is the conditioned trajectory)  particle_Gibbs.jl:160 */
  int64_t T;          /* training length */
  int32_t n_chains;   /* independent PG chains batched on this handle (>= 1) */
  int32_t keep_w_history; /* 1: keep w[t,:] for every t (tests; costs T*N*8 bytes per chain) */
  /* The reference allocates x_pf[n_x,N,T] (particle_Gibbs.jl:159) but reads only x_pf[:,:,t-1] while filtering, one
   * ancestor lineage for the new x_prim (:213-218) and x_pf[:,:,end] for the sample.  PGB_X_HISTORY_REPLAY keeps two
   * rows of x and recomputes that lineage from the stored ancestor indices and the random streams (bit-identical);
   * PGB_X_HISTORY_STORE keeps everything (needed for the x output of pgb_get_history; 8*n_x*N*T bytes per chain);
   * PGB_X_HISTORY_AUTO stores when the history is at most 1 GiB. */
  int32_t x_history;
  int32_t precision;  /* PGB_PRECISION_F64 (0, default) or PGB_PRECISION_F32: see pgb_set_precision notes below */
} pgb_config;

typedef struct pgb_handle pgb_handle;

const char* pgb_version(void);
const char* pgb_last_error(const pgb_handle* h); /* h may be NULL: last create/standalone error */

int pgb_create(pgb_handle** out, const pgb_config* cfg);
void pgb_destroy(pgb_handle* h);

/* u: n_u x T, y: n_y x T (shared by all chains)                      particle_Gibbs.jl:114 args 1,2 */
int pgb_set_data(pgb_handle* h, const double* u, const double* y);
/* g(x,u) = C*x with C n_y x n_x; R n_y x n_y measurement-noise covariance   :193-197 */
int pgb_set_measurement(pgb_handle* h, const double* C, const double* R);
/* x_init_dist = MvNormal(mean, cov): pass mean and the lower Cholesky factor of cov   :163-165 */
int pgb_set_init_dist(pgb_handle* h, const double* mean, const double* chol_lower);
/* MNIW prior: V diagonal (n_phi), Lambda_Q n_x x n_x, ell_Q              :56, :226 */
int pgb_set_prior(pgb_handle* h, const double* V_diag, const double* Lambda_Q, double ell_Q);
/* the same with a dense V (n_phi x n_phi, column-major, symmetric positive definite): MNIW_sample takes any V (:56) and
 * forms Sigbar = Sigma + I / V (:64).  I / V is a constant of the sampler: it is evaluated here, once, on the host
 * (Cholesky + column solves) and added on the device in every draw.  PGB_ERR_NOT_POSDEF when V has no Cholesky factor. */
int pgb_set_prior_dense(pgb_handle* h, const double* V, const double* Lambda_Q, double ell_Q);
/* chain state: A n_x x n_phi, Q n_x x n_x, x_prim n_x x T (NULL = zeros, :126-128);
 * sweep_index = k of the next sweep (1 selects the plain bootstrap filter branch, :183-189). */
int pgb_set_state(pgb_handle* h, int32_t chain, const double* A, const double* Q, const double* x_prim,
                  int64_t sweep_index);
int pgb_get_state(pgb_handle* h, int32_t chain, double* A, double* Q, double* x_prim, int64_t* sweep_index);

/* random streams: counter-based Philox4x32-10 (chain c uses stream id chain_base + c) ... */
int pgb_set_rng_philox(pgb_handle* h, uint64_t seed, uint32_t chain_base);
/* ... or pre-drawn streams for the NEXT sweep of chain `chain` (parity testing; layout = order of
 * consumption in the reference): Z0[n_x*(N-1)], Ures[T], Z[T][n_x*N], Uanc[T], Ustar,
 * bartlett[n_x*n_x] (lower, column-major), Xmn[n_x*n_phi].  Host pointers, copied. */
int pgb_set_rng_injected(pgb_handle* h, int32_t chain, const double* Z0, const double* Ures, const double* Z,
                         const double* Uanc, double Ustar, const double* bartlett, const double* Xmn);

/* Run n_sweeps full PG sweeps (CPF-AS filter, trace, statistics, MNIW draw) on all chains,
 * entirely on the device.  With injected streams n_sweeps must be 1.      particle_Gibbs.jl:154-227 */
int pgb_sweep(pgb_handle* h, int64_t n_sweeps);
/* Split form for tests / benchmarks: filter (+ trace + statistics) only, and the MNIW draw only. */
int pgb_sweep_filter(pgb_handle* h);
int pgb_sweep_mniw(pgb_handle* h);

/* The whole sampler (particle_Gibbs.jl:114-237) for chain-batched handles: runs
 * K_total = K_b + 1 + (K-1)(k_d+1) sweeps and stores K samples per chain.  Outputs (host, may be
 * NULL): A_s[chain][k][n_x*n_phi], Q_s[chain][k][n_x*n_x], w_s[chain][k][N], x_s[chain][k][n_x*N]
 * (x_m1 column-major n_x x N), u_m1[n_u]. */
int pgb_particle_gibbs(pgb_handle* h, int64_t K, int64_t K_b, int64_t k_d, double* A_s, double* Q_s, double* w_s,
                       double* x_s, double* u_m1);

/* current PG_sample fields of a chain (PGopt.jl:23-29; particle_Gibbs.jl:204-208): the (A, Q) the
 * last filter ran with, w[T,:], x_pf[:,:,T], u[:,T] */
int pgb_get_sample(pgb_handle* h, int32_t chain, double* A, double* Q, double* w_m1, double* x_m1, double* u_m1);
/* test-only: a[T x N] (row t = step t, 1-based ancestors, row 0 zero), x[T][N][n_x] (= x_pf[:,n,t]),
 * w[T][N] (needs keep_w_history).  Any pointer may be NULL. */
int pgb_get_history(pgb_handle* h, int32_t chain, int64_t* a, double* x, double* w);
int pgb_get_stats(pgb_handle* h, int32_t chain, double* Phi, double* Psi, double* Sigma, int64_t* star);
/* status bits (PGB_STATUS_*), number of resampling steps that needed the exact serial fallback,
 * and number of exact scans performed.  clear != 0 resets them. */
int pgb_get_status(pgb_handle* h, int32_t chain, int32_t* status_bits, int64_t* n_fallback, int64_t* n_scans,
                   int32_t clear);
int pgb_synchronize(pgb_handle* h);

/* Measurement support (bench.py).  pgb_timer_start/stop bracket work on the handle's stream with
 * CUDA events and return the elapsed device time.  pgb_profile(h, 1) additionally brackets every
 * kernel launch of the filter with an event pair; pgb_get_profile returns, per kernel class
 * (PGB_PROF_*), the number of launches and the summed device time since profiling was switched on.
 * pgb_get_launch_count returns the number of kernels this handle has launched so far. */
enum { PGB_PROF_PREP = 0, PGB_PROF_NORM, PGB_PROF_SCAN_LOCAL, PGB_PROF_RESOLVE, PGB_PROF_EXPAND, PGB_PROF_SEQ,
       PGB_PROF_EXPAND_FB, PGB_PROF_WEIGHTS, PGB_PROF_INIT, PGB_PROF_TAIL, PGB_PROF_MNIW, PGB_PROF_NCLASS };
/* Filter kernel of a sweep.  -1 (default): automatic - when the particle history is stored, the warp-per-chain kernel
 * for N <= 32 (the reference's examples run N = 30) and the one-CTA-per-chain kernel for N <= 1024; otherwise mode 4.
 * 4: the chunk-stationary sweep kernel (one cooperative launch per sweep, pgb_v2.cuh); 1: the 1024-particle tile
 * persistent kernel of round 1 (cross-check); 0: one launch per phase (13 launches per time step, per-kernel
 * profiling); 3: the warp / one-CTA-per-chain kernels (fall back to 1 when N > 1024 or the history is replayed).
 * All of them are bit-identical. */
int pgb_set_mode(pgb_handle* h, int32_t persistent);
int pgb_timer_start(pgb_handle* h);
int pgb_timer_stop(pgb_handle* h, double* elapsed_ms);
int pgb_profile(pgb_handle* h, int32_t on);
int pgb_get_profile(pgb_handle* h, int64_t* launches /*[PGB_PROF_NCLASS]*/, double* ms /*[PGB_PROF_NCLASS]*/);
int pgb_get_launch_count(pgb_handle* h, int64_t* n);

/* Stand-alone helpers matching the exported Julia functions; host buffers in and out. */
/* systematic_resampling(W, N) with rnd = the rand() of :21; idx is 1-based.  *status gets PGB_STATUS_* bits. */
int pgb_systematic_resampling(int32_t device, const double* W, int64_t n_w, int64_t n_draw, double rnd,
                              int64_t* idx, int32_t* status);
/* device time of the kernels of the last pgb_systematic_resampling (tile sums + exact scan + index search), CUDA events */
int pgb_resampling_last_ms(double* kernel_ms);
/* MNIW_sample(Phi, Psi, Sigma, V, Lambda_Q, ell_Q, T) with the Bartlett factor and the matrix-normal
 * normals injected (bartlett/Xmn non-NULL) or drawn from Philox (seed, sweep_index). */
int pgb_mniw_sample(int32_t device, int32_t n_x, int32_t n_phi, const double* Phi, const double* Psi,
                    const double* Sigma, const double* V_diag, const double* Lambda_Q, double ell_Q, int64_t T,
                    const double* bartlett, const double* Xmn, uint64_t seed, uint32_t chain, uint32_t sweep_index,
                    double* A_out, double* Q_out);
/* MNIW_sample with a dense V (n_phi x n_phi, column-major); otherwise as pgb_mniw_sample */
int pgb_mniw_sample_dense(int32_t device, int32_t n_x, int32_t n_phi, const double* Phi, const double* Psi,
                          const double* Sigma, const double* V, const double* Lambda_Q, double ell_Q, int64_t T,
                          const double* bartlett, const double* Xmn, uint64_t seed, uint32_t chain,
                          uint32_t sweep_index, double* A_out, double* Q_out);

/* Scenario sampling + forward rollout (optimal_control_Ipopt.jl:47-83,114-125) for K posterior
 * samples.  cfg supplies the model (N = particles per sample, T ignored).  Inputs (host):
 * A[K][n_x*n_phi], Q[K][n_x*n_x], w_m1[K][N], x_m1[K][n_x*N], u_m1[K][n_u], C, Le (lower factor the
 * measurement noise is unwhitened with: [R] for the reference's scalar-R call, chol(R) otherwise),
 * u_init[n_u*H].  Outputs (host, may be NULL): x0[K][n_x], v[K][H][n_x], e[K][H][n_y],
 * X[K][H+1][n_x], Y[K][H][n_y].  Scenario k uses Philox stream (seed, chain = k_offset + k). */
int pgb_rollout(const pgb_config* cfg, int64_t K, int64_t H, const double* A, const double* Q, const double* w_m1,
                const double* x_m1, const double* u_m1, const double* C, const double* Le, const double* u_init,
                uint64_t seed, uint32_t k_offset, double* x0, double* v, double* e, double* X, double* Y);
/* test_prediction(PG_samples, phi, g, R, k_n, u_test, y_test) (Julia/src/particle_Gibbs.jl:253-314): every one of the
 * K models is simulated k_n times over the test input (T_test steps) on the rollout kernel; simulation (k, kn) is
 * scenario s = k + K*kn (the order of the reference's reshape at :301-302) with Philox stream (seed, chain = s).
 * Inputs as for pgb_rollout (A, Q, w_m1, x_m1, u_m1 hold K entries), u_test[n_u*T_test], y_test[n_y*T_test].
 * Outputs (host; x_sim / y_sim may be NULL): x_sim[K*k_n][T_test+1][n_x] (the reference leaves its last column
 * undefined, :278; here it holds one more transition), y_sim[K*k_n][T_test][n_y], *mean_rmse = the time-matched RMSE
 * sqrt(mean((y_sim[:, t, s] - y_test[:, t])^2)), its sum of squares reduced on the device in the oracle's tree order.
 * Stated deviation: the Julia line :311 subtracts repeat(y_test', 1, 1, K*k_n) (shape T_test x n_y x S), which for
 * n_y = 1 broadcasts to T_test x T_test cross terms and for n_y > 1 throws; the MATLAB twin matches in time. */
int pgb_test_prediction(const pgb_config* cfg, int64_t K, int64_t k_n, int64_t T_test, const double* A,
                        const double* Q, const double* w_m1, const double* x_m1, const double* u_m1, const double* C,
                        const double* Le, const double* u_test, const double* y_test, uint64_t seed, double* x_sim,
                        double* y_sim, double* mean_rmse);
/* Scenario screening of the greedy support-sub-sample search (Julia/src/optimal_control_Ipopt.jl:334-338,
 * optimal_control_Altro.jl:496-500) for box output constraints, the h_scenario of every reference example
 * (examples/PG_OCP_known_basis_functions_Ipopt.jl:198-214): h_max[k] = maximum over the finite bounds (n, t) of
 * y_min[n,t] - y[n,t,k] and y[n,t,k] - y_max[n,t]; order (may be NULL) = sortperm(h_max; rev=true), 1-based.
 * Y[K][H][n_y] as written by pgb_rollout; y_min / y_max [n_y*H] column-major, +-Inf = no bound.  All bounds
 * infinite -> PGB_ERR_INVALID (Julia's maximum of an empty collection throws). */
int pgb_scenario_constraint_max(int32_t device, int64_t K, int64_t H, int32_t n_y, const double* Y,
                                const double* y_min, const double* y_max, double* h_max, int64_t* order);
/* Device-side time (CUDA events, ms) of the last pgb_scenario_constraint_max call of this process: H2D copies, kernel. */
int pgb_scenario_constraint_last_ms(double* h2d_ms, double* kernel_ms);
/* Device-side time (CUDA events, ms) of the last pgb_rollout / pgb_test_prediction call of this process: allocation + host-to-device copies
 * of the inputs, and the rollout kernel itself. */
int pgb_rollout_last_ms(double* h2d_ms, double* kernel_ms);

/* The same with the reference's x_vec_0 / v_vec / e_vec keyword arguments (optimal_control_Ipopt.jl:37, :48, :67, :77): a
 * non-NULL x0_in[K][n_x] / v_in[K][H][n_x] / e_in[K][H][n_y] is used instead of the corresponding draw — the re-entry of
 * the pre-solve step (:88-112) and the route for identical-stream parity runs fed from Julia.  With x0_in the arguments
 * w_m1 / x_m1 / u_m1 may be NULL; with x0_in and v_in also Q.  Supplied arrays are echoed into x0 / v / e. */
int pgb_rollout_supplied(const pgb_config* cfg, int64_t K, int64_t H, const double* A, const double* Q,
                         const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                         const double* u_init, uint64_t seed, uint32_t k_offset, const double* x0_in, const double* v_in,
                         const double* e_in, double* x0, double* v, double* e, double* X, double* Y);

/* ---- device-resident pipeline: sampler -> samples -> rollout -> screening --------------------------------------
 * The reference reads PG_samples[k].A in place (optimal_control_Ipopt.jl:47-64).  A pgb_samples set keeps K PG_sample
 * records (PGopt.jl:23-29) and the scenario arrays of its last rollout in device memory, allocated once. */
typedef struct pgb_samples pgb_samples;
/* cfg: device, model, N (particles per sample); K: capacity = number of scenarios */
int pgb_samples_create(pgb_samples** out, const pgb_config* cfg, int64_t K);
void pgb_samples_destroy(pgb_samples* s);
const char* pgb_samples_last_error(const pgb_samples* s);
/* host <-> device for samples [k0, k0+n); layouts of pgb_rollout; NULL pointers are skipped */
int pgb_samples_set(pgb_samples* s, int64_t k0, int64_t n, const double* A, const double* Q, const double* w_m1,
                    const double* x_m1, const double* u_m1);
int pgb_samples_get(pgb_samples* s, int64_t k0, int64_t n, double* A, double* Q, double* w_m1, double* x_m1,
                    double* u_m1);
/* slot k <- the current PG_sample of `chain` of handle h (same device), device to device (particle_Gibbs.jl:204-208) */
int pgb_samples_store(pgb_samples* s, int64_t k, pgb_handle* h, int32_t chain);
/* pgb_particle_gibbs with the kept samples left on the device: chain c, sample k -> slot c*K + k of `out` */
int pgb_particle_gibbs_dev(pgb_handle* h, int64_t K, int64_t K_b, int64_t k_d, pgb_samples* out);
/* x_T[k] = x_m1[:, star_k] with star_k = systematic_resampling(w_m1, 1) (optimal_control_Ipopt.jl:58-59), drawn where
 * the sample lives from Philox stream (seed, k_offset + k) — the stream pgb_rollout uses for the same draw; a later
 * pgb_rollout_dev with the same (seed, k_offset) reuses it.  x_T (host, may be NULL): [K][n_x]. */
int pgb_samples_draw_xT(pgb_samples* s, uint64_t seed, uint32_t k_offset, double* x_T);
/* pgb_rollout_supplied on the resident samples (scenario k = sample k); results stay on the device.  x0_in / v_in / e_in
 * are host arrays or NULL; reuse bits 1 / 2 / 4 keep the x_vec_0 / v_vec / e_vec of the previous rollout of this set
 * (same H) — the pre-solve re-entry with a new u_init (optimal_control_Ipopt.jl:88-125) without moving anything. */
int pgb_rollout_dev(pgb_samples* s, int64_t H, const double* C, const double* Le, const double* u_init, uint64_t seed,
                    uint32_t k_offset, int32_t reuse, const double* x0_in, const double* v_in, const double* e_in);
/* scenario arrays of the last pgb_rollout_dev -> host; any pointer may be NULL; layouts of pgb_rollout */
int pgb_scenarios_get(pgb_samples* s, double* x0, double* v, double* e, double* X, double* Y);
/* pgb_scenario_constraint_max on the Y of the last pgb_rollout_dev, read where the rollout left it */
int pgb_scenario_constraint_max_dev(pgb_samples* s, const double* y_min, const double* y_max, double* h_max,
                                    int64_t* order);
/* One step of the stacked K-scenario model and its Jacobians on the resident samples:
 * RobotDynamics.discrete_dynamics[!] of PGS_model_obj, x_vec_n[k] = A_k*phi(x_vec[k], u) + v_vec[:, t+1, k]
 * (Julia/src/optimal_control_Altro.jl:79-107), and what RobotDynamics.@autodiff (:51) derives from it.
 * x_vec: host [K*n_x] in the reference's stacking (scenario k = rows n_x*(k-1)+1 .. n_x*k); u [n_u]; t >= 0 adds the
 * process noise column t (0-based) the last pgb_rollout_dev left on the device, t < 0 adds none; phi_variant 0 = the
 * rounding of phi_sampling, 1 = of the examples' phi_opt (examples/PG_OCP_generic_basis_functions_Altro.jl:204-217).
 * Outputs (host): x_next [K*n_x]; dfdx [K][n_x*n_x] — the diagonal blocks of the (K n_x) x (K n_x) Jacobian, column-major;
 * dfdu [K][n_x*n_u] — the rows of the (K n_x) x n_u Jacobian, column-major per scenario.  dfdx / dfdu may be NULL.
 * With phi_variant 0, x_next is bit-identical to the X[t+1] the rollout computes from the same state. */
int pgb_dynamics_dev(pgb_samples* s, const double* x_vec, const double* u, int64_t t, int32_t phi_variant, double* x_next,
                     double* dfdx, double* dfdu);
int pgb_dynamics_last_ms(pgb_samples* s, double* kernel_ms);
/* device time (CUDA events, ms) of the last rollout kernel / screening kernel / input upload of this set */
int pgb_samples_last_ms(pgb_samples* s, double* rollout_kernel_ms, double* screening_kernel_ms, double* upload_ms);
/* page-locked host memory for the caller's I/O arrays (NULL on failure) */
void* pgb_host_alloc(size_t bytes);
void pgb_host_free(void* p);

/* ---- several GPUs of one node: independent chains and independent scenarios -------------
 * One process drives the listed devices (one host thread per device while they work, in-process ncclCommInitAll);
 * cfg->device is ignored and cfg->n_chains is the TOTAL number of chains, dealt to the devices in contiguous balanced
 * blocks.  Global chain c always uses Philox stream id chain_base + c and global sample g = c*K + k the scenario
 * stream id g, so results do not depend on the number of devices.  The only collective is the NCCL all-gather at the
 * end, device buffer to device buffer. */
typedef struct pgb_multi pgb_multi;
int pgb_multi_create(pgb_multi** out, const pgb_config* cfg, const int32_t* devices, int32_t n_devices);
void pgb_multi_destroy(pgb_multi* m);
const char* pgb_multi_last_error(const pgb_multi* m);
int pgb_multi_n_devices(const pgb_multi* m);
int pgb_multi_set_data(pgb_multi* m, const double* u, const double* y);
int pgb_multi_set_measurement(pgb_multi* m, const double* C, const double* R);
int pgb_multi_set_init_dist(pgb_multi* m, const double* mean, const double* chol_lower);
int pgb_multi_set_prior(pgb_multi* m, const double* V_diag, const double* Lambda_Q, double ell_Q);
int pgb_multi_set_prior_dense(pgb_multi* m, const double* V, const double* Lambda_Q, double ell_Q);
/* chain: global index, or -1 = every chain */
int pgb_multi_set_state(pgb_multi* m, int32_t chain, const double* A, const double* Q, const double* x_prim,
                        int64_t sweep_index);
int pgb_multi_set_rng_philox(pgb_multi* m, uint64_t seed, uint32_t chain_base);
/* packed posterior-sample record (also the wire format of the gather and of pgb_samples_pack):
 * [A n_x*n_phi column-major | Q n_x*n_x | x_T n_x | u_m1 n_u] doubles */
int64_t pgb_packed_record_doubles(int32_t n_x, int32_t n_phi, int32_t n_u);
/* particle_Gibbs on every chain; the kept samples stay on the device that produced them; x_T = x_m1[:, star] is drawn
 * there (optimal_control_Ipopt.jl:58-59, seed roll_seed, stream id c*K + k) and the packed records are all-gathered to
 * every device.  packed (host, may be NULL): [n_chains*K][record], chain-major. */
int pgb_multi_particle_gibbs(pgb_multi* m, int64_t K, int64_t K_b, int64_t k_d, uint64_t roll_seed, double* packed);
int pgb_multi_get_packed(pgb_multi* m, int32_t which_device, double* packed);
/* PG_sample k of global chain c from the device that holds it (w_m1 / x_m1 are never gathered) */
int pgb_multi_get_sample(pgb_multi* m, int32_t chain, int64_t k, double* A, double* Q, double* w_m1, double* x_m1,
                         double* u_m1);
/* K_total host samples dealt to the devices in contiguous blocks and kept resident (the configs[4] shape); layouts of
 * pgb_rollout; x_T drawn on the owning device with (roll_seed, stream id = scenario index) */
int pgb_multi_set_samples(pgb_multi* m, int64_t K_total, const double* A, const double* Q, const double* w_m1,
                          const double* x_m1, const double* u_m1, uint64_t roll_seed);
/* scenario rollout of all resident samples, each device its own, then one all-gather of x_vec_0 / X / Y; host outputs
 * (may be NULL) in global order, layouts of pgb_rollout — each device sends its own block over its own PCIe link while
 * the NVLink gather runs.  seed = the roll_seed the x_T were drawn with. */
int pgb_multi_rollout(pgb_multi* m, int64_t H, const double* C, const double* Le, const double* u_init, uint64_t seed,
                      double* x0, double* X, double* Y);
/* the gathered scenario arrays as device which_device of the list holds them after pgb_multi_rollout (every device holds
 * all of them: what a solver running on any GPU reads); host outputs may be NULL */
int pgb_multi_get_scenarios(pgb_multi* m, int32_t which_device, double* x0, double* X, double* Y);
/* discrete_dynamics[!] and its Jacobians (optimal_control_Altro.jl:79-107, :51; see pgb_dynamics_dev) on the sharded resident
 * samples: every device evaluates its own scenarios from / into its block of the host arrays (layouts of pgb_dynamics_dev
 * over all K scenarios); independent scenarios, no collective.  kernel_ms_max (optional): slowest device's kernel. */
int pgb_multi_dynamics(pgb_multi* m, const double* x_vec, const double* u, int64_t t, int32_t phi_variant, double* x_next,
                       double* dfdx, double* dfdu, double* kernel_ms_max);
/* device-side times (CUDA events, ms): NCCL gather of the packed samples; slowest device's rollout kernel; NCCL gather
 * of the scenario arrays */
int pgb_multi_last_ms(pgb_multi* m, double* sample_gather_ms, double* rollout_kernel_ms_max, double* scenario_gather_ms);

/* ---- sample serialisation and checkpoint / resume ----------------------------------
 * The reference keeps PG_sample only in memory (PGopt.jl:23-29) and warm-starts through the x_prim keyword
 * (particle_Gibbs.jl:114,126-128).  Packed layout, little-endian, version 1:
 *   64-byte header {char magic[8] = "PGB200S\0"; u32 version; u32 flags (bit 0: particles follow, bit 1: x_T valid); i32 n_x, n_u, n_phi,
 *   n_y; i64 N; i64 K; u64 reserved[2]} | K records [A | Q | x_T | u_m1] (pgb_packed_record_doubles) | if flags&1:
 *   K x [w_m1 N | x_m1 n_x*N column-major].  x_T is filled by pgb_samples_draw_xT (zeros before). */
int64_t pgb_samples_pack_bytes(const pgb_samples* s, int32_t with_particles);
int pgb_samples_pack(pgb_samples* s, int32_t with_particles, void* buf, int64_t buf_bytes);
/* the inverse: validates magic / version / shapes against the set (PGB_ERR_INVALID on mismatch) */
int pgb_samples_unpack(pgb_samples* s, const void* buf, int64_t buf_bytes);
/* Sampler checkpoint: everything a later pgb_checkpoint_load on a handle of the same configuration needs to continue
 * the chains bit-identically (per chain A, Q, x_prim; sweep index; Philox seed and chain_base).  Same header with
 * magic "PGB200C\0". */
int64_t pgb_checkpoint_bytes(const pgb_handle* h);
int pgb_checkpoint_save(pgb_handle* h, void* buf, int64_t buf_bytes);
int pgb_checkpoint_load(pgb_handle* h, const void* buf, int64_t buf_bytes);

/* ---- test hooks (not part of the drop-in surface) ---------------------------------------------
 * pgb_test_force_fallback(1) makes every exact scan take the serial-walk fallback path, (2) makes the ancestor
 * draw use the exact scan instead of the certified approximate shortcut (results must not change); pgb_test_math evaluates the device build of pgb_math.h: fn 0 exp, 1 log, 2 sin, 3 cos,
 * 4 Box-Muller pairs from Philox (x[0] = seed, y gets 2n normals of purpose Z, sweep 1, t 5);
 * 10..14: the same through the 4-wide branch-free versions of pgb_vmath.cuh (11: positive normal arguments);
 * 30..34: the FP32 functions of pgb_mathf.h (exp, log, sin, cos, wide-range exp) on (float)x, result widened; 35: y gets
 * 4n FP32 normals (pgf_stream_normals of particles 0..n-1, x[0] = seed). */
void pgb_test_force_fallback(int32_t on);
int pgb_test_math(int32_t device, int32_t fn, const double* x, double* y, int64_t n);
/* DFMA micro-benchmark (8 independent chains per thread, all SMs): measured FP64 FMA rate of the device. */
int pgb_test_fp64_peak(int32_t device, double* tflops);
/* rollout kernel choice for the tests: -1 automatic (default), 0 CTA per scenario, 1 warp per scenario */
void pgb_test_rollout_kernel(int32_t which);
/* dynamics kernel choice for the tests: -1 automatic (default), 0 k_dynamics (strided loads, every shape), 1 / 2 / 3
 * k_dynamics_bulk with at most 4 / 2 / 8 warps per CTA (cp.async.bulk staging of A_k; Hilbert basis, even n_x,
 * n_z = 3 | 5, else k_dynamics) */
void pgb_test_dynamics_kernel(int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* PGOPT_B200_H */
