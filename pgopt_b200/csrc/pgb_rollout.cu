// pgb_rollout.cu — scenario sampling + forward rollout (optimal_control_Ipopt.jl:47-83,114-125), test_prediction
// (particle_Gibbs.jl:253-314) and scenario screening (optimal_control_Ipopt.jl:334-338): host side.
#include "pgb_internal.h"
#include "pgb_rollout.cuh"

using namespace pgb;

static int g_rollout_grid[3] = {0, 0, 0};  // last warp-kernel launch: {CTAs, CTAs per SM, runs of A in shared memory}
static int g_rollout_force = -1;   // pgb_test_rollout_kernel: -1 automatic, 0 CTA per scenario, 1 warp per scenario
static int g_rollout_kernel = -1;  // kernel the last pgb_rollout ran: 1 = warp per scenario, 0 = CTA per scenario
static float g_rollout_ms[2] = {0.f, 0.f};  // device time of the last pgb_rollout: {H2D copies, rollout kernel}

template <int NX, int NI>
static cudaError_t launch_rollout_ni(const FilterDev& fd, const RolloutDev& rd, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_rollout<NX, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_rollout<NX, NI><<<(unsigned)rd.K, R_TPB, smem, st>>>(fd, rd);
  return cudaGetLastError();
}
// n_phi <= 128: one register-resident column of A per thread; otherwise five (n_phi <= 640 entirely in registers,
// e.g. 5^4 = 625), any further columns are read from L2
template <int NX>
static cudaError_t launch_rollout(const FilterDev& fd, const RolloutDev& rd, size_t smem, cudaStream_t st) {
  return fd.n_phi <= R_TPB ? launch_rollout_ni<NX, 1>(fd, rd, smem, st) : launch_rollout_ni<NX, 5>(fd, rd, smem, st);
}

// Warp-per-scenario kernel (default): A_k resident over 32 lanes (RS of each lane's 4 runs in shared memory, the
// rest in registers), persistent grid of independent warps.
static size_t rollout_warp_smem(int nx, int ny, int nu, int64_t H, int ni, int rs) {
  return sizeof(RolloutWShared) +
         ((size_t)H * nu + (size_t)RW_WPB * (PGB_MAX_Z_ * MAXC + (size_t)H * (nx + ny) + (size_t)rs * ni * nx * 32 + (size_t)(H + 2) * nx)) * 8;
}
template <int NX, int NI, int RS>
static cudaError_t launch_rollout_warp_ni(const FilterDev& fd, const RolloutDev& rd, int n_sm, cudaStream_t st) {
  const size_t smem = rollout_warp_smem(NX, fd.n_y, fd.n_u, rd.H, NI, RS);
  cudaError_t e = cudaFuncSetAttribute(k_rollout_warp<NX, NI, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rollout_warp<NX, NI, RS>, RW_WPB * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (rd.K + RW_WPB - 1) / RW_WPB;
  if (grid > (int64_t)n_sm * per_sm) grid = (int64_t)n_sm * per_sm;
  g_rollout_grid[0] = (int)grid; g_rollout_grid[1] = per_sm; g_rollout_grid[2] = RS;
  k_rollout_warp<NX, NI, RS><<<(unsigned)grid, RW_WPB * 32, smem, st>>>(fd, rd);
  return cudaGetLastError();
}
template <int NX>
static cudaError_t launch_rollout_warp(const FilterDev& fd, const RolloutDev& rd, int n_sm, int rs, cudaStream_t st) {
  if (fd.n_phi <= 128) return launch_rollout_warp_ni<NX, 1, 0>(fd, rd, n_sm, st);
  (void)rs;
  return launch_rollout_warp_ni<NX, RW_MAXNI, 2>(fd, rd, n_sm, st);
}

// Kernel choice + launch on device-resident inputs (everything in rd is a device pointer), on stream st.
static int rollout_launch(pgb_handle* h, int device, const FilterDev& fd, const RolloutDev& rd, cudaStream_t st) {
  const int nx = fd.n_x, np = fd.n_phi, nu = fd.n_u, ny = fd.n_y;
  const int64_t H = rd.H, N = rd.N;
  const size_t smem = sizeof(RolloutShared) + ((size_t)H * (nx + ny + nu) + (size_t)(N < R_WSTAGE ? N : R_WSTAGE)) * 8;
  if (smem > 200 * 1024 || rd.K > 0x7fffffffLL)
    return set_err(h, PGB_ERR_INVALID, "rollout: horizon H = %lld (or K) too large", (long long)H);
  // one warp per scenario whenever A_k fits the registers of 32 lanes (n_phi <= 128 * RW_MAXNI) and the per-warp noise
  // buffers fit shared memory; pgb_test_rollout_kernel forces one (tests run both)
  const int rs = 2;  // runs (of 4 per lane) of A_k kept in shared memory (0 / 2 / 3 measured: 4.8 / 4.7 / 5.1 ms at C5)
  const size_t smem_w = rollout_warp_smem(nx, ny, nu, H, np <= 128 ? 1 : RW_MAXNI, np <= 128 ? 0 : rs);
  bool use_warp = np <= 128 * RW_MAXNI && smem_w <= 100 * 1024;
  if (g_rollout_force == 0) use_warp = false;  // test hook: the CTA-per-scenario kernel on a shape the warp kernel covers
  else if (g_rollout_force == 1 && !use_warp)
    return set_err(h, PGB_ERR_INVALID, "rollout: warp kernel needs n_phi <= %d and a shorter horizon", 128 * RW_MAXNI);
  g_rollout_kernel = use_warp ? 1 : 0;
  if (use_warp) {
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    switch (nx) {
      case 1: CK(h, launch_rollout_warp<1>(fd, rd, n_sm, rs, st)); break;
      case 2: CK(h, launch_rollout_warp<2>(fd, rd, n_sm, rs, st)); break;
      case 3: CK(h, launch_rollout_warp<3>(fd, rd, n_sm, rs, st)); break;
      default: CK(h, launch_rollout_warp<4>(fd, rd, n_sm, rs, st)); break;
    }
  } else {
    switch (nx) {
      case 1: CK(h, launch_rollout<1>(fd, rd, smem, st)); break;
      case 2: CK(h, launch_rollout<2>(fd, rd, smem, st)); break;
      case 3: CK(h, launch_rollout<3>(fd, rd, smem, st)); break;
      default: CK(h, launch_rollout<4>(fd, rd, smem, st)); break;
    }
  }
  return PGB_OK;
}

// Shared body of pgb_rollout / pgb_rollout_supplied (n_models = 0, y_test = null) and pgb_test_prediction (K scenarios
// over n_models models, scenario k simulating model k mod n_models; sum_sq_err = tree sum of (Y - y_test)^2).  Host
// buffers in and out; x0_in / v_in / e_in (may be NULL) are the reference's x_vec_0 / v_vec / e_vec kwargs.
static int rollout_run(const pgb_config* cfg, int64_t K, int64_t n_models, int64_t H, const double* A, const double* Q,
                       const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                       const double* u_init, uint64_t seed, uint32_t k_offset, const double* x0_in, const double* v_in,
                       const double* e_in, double* x0, double* v, double* e, double* X, double* Y, const double* y_test,
                       double* sum_sq_err) {
  int rc = pgb_validate_cfg(cfg);
  if (rc) return rc;
  const int64_t KM = n_models > 0 ? n_models : K;  // number of models (PG samples) behind the K scenarios
  const bool need_draw0 = (x0_in == nullptr);      // star / x_m1 / u_m1 are only read for the x_vec_0 draw (:48-64)
  if (!A || !C || !Le || !u_init || K < 1 || H < 1 || cfg->N < 1) return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  if (need_draw0 && (!w_m1 || !x_m1 || !u_m1)) return set_err(nullptr, PGB_ERR_INVALID, "w_m1 / x_m1 / u_m1 needed to draw x_vec_0");
  if ((need_draw0 || !v_in) && !Q) return set_err(nullptr, PGB_ERR_INVALID, "Q needed to draw x_vec_0 / v_vec");
  // scenario k draws from Philox stream id k_offset + k, a 24-bit field of the counter (pgb_math.h: pgb_stream_block)
  if ((uint64_t)k_offset + (uint64_t)K > (1ull << 24))
    return set_err(nullptr, PGB_ERR_INVALID, "k_offset + K = %llu exceeds the 2^24 Philox stream ids",
                   (unsigned long long)k_offset + (unsigned long long)K);
  if (int rcd = pgb_require_device()) return rcd;
  TmpHandle tmp;
  pgb_handle* h = &tmp;
  CK(h, cudaSetDevice(cfg->device));
  const int nx = cfg->n_x, np = cfg->n_phi, nu = cfg->n_u, ny = cfg->n_y;
  const int64_t N = cfg->N;
  FilterDev fd;
  memset(&fd, 0, sizeof(fd));
  pgb_fill_model(fd, cfg);
  for (int i = 0; i < ny * nx; ++i) fd.C[i] = C[i];
  RolloutDev rd;
  memset(&rd, 0, sizeof(rd));
  rd.K = K; rd.H = H; rd.N = N; rd.seed = seed; rd.k_offset = k_offset; rd.n_models = n_models;
  for (int i = 0; i < ny * ny; ++i) rd.Le[i] = Le[i];
  const int m = nx * np;
  double *dA, *dQ, *dw = nullptr, *dx = nullptr, *dum = nullptr, *dui;
  cudaEvent_t ev[3];
  for (auto& evt : ev) CK(h, cudaEventCreate(&evt));
  struct EvCleanup {
    cudaEvent_t* ev;
    ~EvCleanup() { for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]); }
  } ev_cleanup{ev};
  CK(h, cudaEventRecord(ev[0], 0));
  CK(h, dalloc(h, &dA, (size_t)KM * m, false));
  CK(h, dalloc(h, &dQ, (size_t)KM * nx * nx, Q == nullptr));
  CK(h, dalloc(h, &dui, (size_t)nu * H, false));
  CK(h, dalloc(h, &rd.status, 1));
  CK(h, cudaMemcpy(dA, A, (size_t)KM * m * 8, cudaMemcpyHostToDevice));
  if (Q) CK(h, cudaMemcpy(dQ, Q, (size_t)KM * nx * nx * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(dui, u_init, (size_t)nu * H * 8, cudaMemcpyHostToDevice));
  if (need_draw0) {
    CK(h, dalloc(h, &dw, (size_t)KM * N, false));
    CK(h, dalloc(h, &dx, (size_t)KM * nx * N, false));
    CK(h, dalloc(h, &dum, (size_t)KM * nu, false));
    CK(h, cudaMemcpy(dw, w_m1, (size_t)KM * N * 8, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(dx, x_m1, (size_t)KM * nx * N * 8, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(dum, u_m1, (size_t)KM * nu * 8, cudaMemcpyHostToDevice));
  }
  double *dx0i = nullptr, *dvi = nullptr, *dei = nullptr;
  if (x0_in) { CK(h, dalloc(h, &dx0i, (size_t)K * nx, false)); CK(h, cudaMemcpy(dx0i, x0_in, (size_t)K * nx * 8, cudaMemcpyHostToDevice)); }
  if (v_in) { CK(h, dalloc(h, &dvi, (size_t)K * H * nx, false)); CK(h, cudaMemcpy(dvi, v_in, (size_t)K * H * nx * 8, cudaMemcpyHostToDevice)); }
  if (e_in) { CK(h, dalloc(h, &dei, (size_t)K * H * ny, false)); CK(h, cudaMemcpy(dei, e_in, (size_t)K * H * ny * 8, cudaMemcpyHostToDevice)); }
  rd.x0_in = dx0i; rd.v_in = dvi; rd.e_in = dei;
  if (x0 && !x0_in) CK(h, dalloc(h, &rd.x0, (size_t)K * nx, false));
  if (v && !v_in) CK(h, dalloc(h, &rd.v, (size_t)K * H * nx, false));
  if (e && !e_in) CK(h, dalloc(h, &rd.e, (size_t)K * H * ny, false));
  if (X) CK(h, dalloc(h, &rd.X, (size_t)K * (H + 1) * nx, false));
  if (Y || y_test) CK(h, dalloc(h, &rd.Y, (size_t)K * H * ny, false));
  CK(h, cudaEventRecord(ev[1], 0));
  rd.A = dA; rd.Q = dQ; rd.w_m1 = dw; rd.x_m1 = dx; rd.u_m1 = dum; rd.u_init = dui;
  if ((rc = rollout_launch(nullptr, cfg->device, fd, rd, 0))) return rc;
  CK(h, cudaEventRecord(ev[2], 0));
  double* d_sse = nullptr;
  if (y_test && sum_sq_err) {  // squared prediction errors, tree-summed on the device (particle_Gibbs.jl:311)
    const int64_t M = K * H * ny, ntile = (M + TILE - 1) / TILE;
    double *d_yt, *d_tiles;
    CK(h, dalloc(h, &d_yt, (size_t)H * ny, false));
    CK(h, dalloc(h, &d_tiles, (size_t)ntile, false));
    CK(h, dalloc(h, &d_sse, 1, false));
    CK(h, cudaMemcpy(d_yt, y_test, (size_t)H * ny * 8, cudaMemcpyHostToDevice));
    k_sqerr_tiles<<<(unsigned)ntile, TPB>>>(rd.Y, d_yt, M, H * ny, d_tiles);
    k_tree_final<<<1, TPB>>>(d_tiles, ntile, d_sse);
    CK(h, cudaGetLastError());
  }
  CK(h, cudaDeviceSynchronize());
  if (d_sse) CK(h, cudaMemcpy(sum_sq_err, d_sse, 8, cudaMemcpyDeviceToHost));
  cudaEventElapsedTime(&g_rollout_ms[0], ev[0], ev[1]);
  cudaEventElapsedTime(&g_rollout_ms[1], ev[1], ev[2]);
  int st = 0;
  CK(h, cudaMemcpy(&st, rd.status, 4, cudaMemcpyDeviceToHost));
  if (st & PGB_STATUS_NOT_POSDEF) return set_err(nullptr, PGB_ERR_NOT_POSDEF, "a scenario's Q is not positive definite");
  // supplied arrays are handed back unchanged, as the reference returns its kwargs
  if (x0) { if (x0_in) { if (x0 != x0_in) memcpy(x0, x0_in, (size_t)K * nx * 8); } else CK(h, cudaMemcpy(x0, rd.x0, (size_t)K * nx * 8, cudaMemcpyDeviceToHost)); }
  if (v) { if (v_in) { if (v != v_in) memcpy(v, v_in, (size_t)K * H * nx * 8); } else CK(h, cudaMemcpy(v, rd.v, (size_t)K * H * nx * 8, cudaMemcpyDeviceToHost)); }
  if (e) { if (e_in) { if (e != e_in) memcpy(e, e_in, (size_t)K * H * ny * 8); } else CK(h, cudaMemcpy(e, rd.e, (size_t)K * H * ny * 8, cudaMemcpyDeviceToHost)); }
  if (X) CK(h, cudaMemcpy(X, rd.X, (size_t)K * (H + 1) * nx * 8, cudaMemcpyDeviceToHost));
  if (Y) CK(h, cudaMemcpy(Y, rd.Y, (size_t)K * H * ny * 8, cudaMemcpyDeviceToHost));
  return PGB_OK;
}

extern "C" int pgb_rollout(const pgb_config* cfg, int64_t K, int64_t H, const double* A, const double* Q,
                           const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                           const double* u_init, uint64_t seed, uint32_t k_offset, double* x0, double* v, double* e,
                           double* X, double* Y) {
  return rollout_run(cfg, K, 0, H, A, Q, w_m1, x_m1, u_m1, C, Le, u_init, seed, k_offset, nullptr, nullptr, nullptr, x0, v, e,
                     X, Y, nullptr, nullptr);
}

extern "C" int pgb_rollout_supplied(const pgb_config* cfg, int64_t K, int64_t H, const double* A, const double* Q,
                                    const double* w_m1, const double* x_m1, const double* u_m1, const double* C,
                                    const double* Le, const double* u_init, uint64_t seed, uint32_t k_offset,
                                    const double* x0_in, const double* v_in, const double* e_in, double* x0, double* v,
                                    double* e, double* X, double* Y) {
  return rollout_run(cfg, K, 0, H, A, Q, w_m1, x_m1, u_m1, C, Le, u_init, seed, k_offset, x0_in, v_in, e_in, x0, v, e, X, Y,
                     nullptr, nullptr);
}

extern "C" int pgb_test_prediction(const pgb_config* cfg, int64_t K, int64_t k_n, int64_t T_test, const double* A,
                                   const double* Q, const double* w_m1, const double* x_m1, const double* u_m1,
                                   const double* C, const double* Le, const double* u_test, const double* y_test,
                                   uint64_t seed, double* x_sim, double* y_sim, double* mean_rmse) {
  if (K < 1 || k_n < 1 || T_test < 1 || !y_test || !mean_rmse) return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  double sse = 0.0;
  int rc = rollout_run(cfg, K * k_n, K, T_test, A, Q, w_m1, x_m1, u_m1, C, Le, u_test, seed, 0, nullptr, nullptr, nullptr,
                       nullptr, nullptr, nullptr, x_sim, y_sim, y_test, &sse);
  if (rc) return rc;
  *mean_rmse = std::sqrt(sse / (double)(K * k_n * T_test * cfg->n_y));
  return PGB_OK;
}

extern "C" int pgb_rollout_last_ms(double* h2d_ms, double* kernel_ms) {
  if (h2d_ms) *h2d_ms = g_rollout_ms[0];
  if (kernel_ms) *kernel_ms = g_rollout_ms[1];
  return PGB_OK;
}

// ------------------------------------------------------------------------------------------------
// scenario screening for the greedy support-sub-sample search (optimal_control_Ipopt.jl:334-338, Altro.jl:496-500)
// ------------------------------------------------------------------------------------------------
// One warp per scenario, grid-stride: the scenario's H*n_y outputs are contiguous (coalesced 256 B per warp request);
// the bounds (n_y x H column-major = the same element index) stay in L1.  HBM-bound: 8 B read per output value.
__device__ __forceinline__ double hmax_take(double m, double h) {
  return (h > m || (h == m && !signbit(h))) ? h : m;  // Julia's max on non-NaN values (0.0 above -0.0)
}
constexpr int SCREEN_U = 4;  // scenarios per warp iteration: independent loads in flight (the kernel is latency-bound otherwise)
__global__ void __launch_bounds__(256) k_scenario_constraint_max(int64_t K, int64_t E, const double* __restrict__ Y,
                                                                 const double* __restrict__ lo,
                                                                 const double* __restrict__ hi,
                                                                 double* __restrict__ h_max) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t k0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SCREEN_U; k0 < K;
       k0 += nwarp * SCREEN_U) {
    const double* y[SCREEN_U];
    double m[SCREEN_U];
    int nan[SCREEN_U];
#pragma unroll
    for (int u = 0; u < SCREEN_U; ++u) {
      y[u] = Y + (k0 + u < K ? k0 + u : K - 1) * E;  // tail: re-read the last scenario, never written
      m[u] = -INFINITY;
      nan[u] = 0;
    }
#pragma unroll 2
    for (int64_t e = lane; e < E; e += 32) {
      const double a = __ldg(lo + e), b = __ldg(hi + e);
      if (!isfinite(a) && !isfinite(b)) continue;  // column without a bound: its outputs are never read
      double v[SCREEN_U];
#pragma unroll
      for (int u = 0; u < SCREEN_U; ++u) v[u] = y[u][e];
#pragma unroll
      for (int u = 0; u < SCREEN_U; ++u) {
        if (isfinite(a)) { const double h = a - v[u]; if (isnan(h)) nan[u] = 1; else m[u] = hmax_take(m[u], h); }
        if (isfinite(b)) { const double h = v[u] - b; if (isnan(h)) nan[u] = 1; else m[u] = hmax_take(m[u], h); }
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int u = 0; u < SCREEN_U; ++u) {
        m[u] = hmax_take(m[u], __shfl_xor_sync(0xffffffffu, m[u], o));
        nan[u] |= __shfl_xor_sync(0xffffffffu, nan[u], o);
      }
    }
#pragma unroll
    for (int u = 0; u < SCREEN_U; ++u)
      if (lane == u && k0 + u < K) h_max[k0 + u] = nan[u] ? __longlong_as_double(0x7ff8000000000000ll) : m[u];
  }
}

static float g_screen_ms[2] = {0.f, 0.f};  // device time of the last pgb_scenario_constraint_max: {H2D copies, kernel}

static inline bool jl_isless(double a, double b) {  // Julia's isless on Float64
  if (std::isnan(a)) return false;
  if (std::isnan(b)) return true;
  if (a == 0.0 && b == 0.0) return std::signbit(a) && !std::signbit(b);
  return a < b;
}

extern "C" int pgb_scenario_constraint_max(int32_t device, int64_t K, int64_t H, int32_t n_y, const double* Y,
                                           const double* y_min, const double* y_max, double* h_max, int64_t* order) {
  if (!Y || !y_min || !y_max || !h_max || K < 1 || H < 1 || n_y < 1)
    return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  const int64_t E = H * (int64_t)n_y;
  int64_t n_fin = 0;
  for (int64_t i = 0; i < E; ++i) n_fin += (std::isfinite(y_min[i]) ? 1 : 0) + (std::isfinite(y_max[i]) ? 1 : 0);
  if (n_fin == 0)  // Julia: maximum of an empty collection throws (optimal_control_Ipopt.jl:336)
    return set_err(nullptr, PGB_ERR_INVALID, "ArgumentError: no finite output bound: maximum over an empty collection");
  if (int rcd = pgb_require_device()) return rcd;
  TmpHandle tmp;  // only used as an allocation list / error sink
  pgb_handle* h = &tmp;
  h->cfg.device = device;
  CK(h, cudaSetDevice(device));
  double *d_Y = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_h = nullptr;
  CK(h, dalloc(h, &d_Y, (size_t)(K * E)));
  CK(h, dalloc(h, &d_lo, (size_t)E));
  CK(h, dalloc(h, &d_hi, (size_t)E));
  CK(h, dalloc(h, &d_h, (size_t)K));
  cudaEvent_t ev[3];
  for (auto& v : ev) CK(h, cudaEventCreate(&v));
  struct EvCleanup {
    cudaEvent_t* ev;
    ~EvCleanup() { for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]); }
  } ev_cleanup{ev};
  CK(h, cudaEventRecord(ev[0], 0));
  CK(h, cudaMemcpy(d_Y, Y, (size_t)(K * E) * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(d_lo, y_min, (size_t)E * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(d_hi, y_max, (size_t)E * 8, cudaMemcpyHostToDevice));
  CK(h, cudaEventRecord(ev[1], 0));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int64_t want = (K + 8 * SCREEN_U - 1) / (8 * SCREEN_U);  // 8 warps x SCREEN_U scenarios per CTA iteration
  const int grid = (int)std::min<int64_t>(want, (int64_t)sms * 8);
  k_scenario_constraint_max<<<grid, 256>>>(K, E, d_Y, d_lo, d_hi, d_h);
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(ev[2], 0));
  CK(h, cudaEventSynchronize(ev[2]));
  cudaEventElapsedTime(&g_screen_ms[0], ev[0], ev[1]);
  cudaEventElapsedTime(&g_screen_ms[1], ev[1], ev[2]);
  CK(h, cudaMemcpy(h_max, d_h, (size_t)K * 8, cudaMemcpyDeviceToHost));
  if (order) {  // sortperm(h_scenario_max; rev=true): stable, ties keep their original order; 1-based
    for (int64_t k = 0; k < K; ++k) order[k] = k + 1;
    std::stable_sort(order, order + K, [&](int64_t a, int64_t b) { return jl_isless(h_max[b - 1], h_max[a - 1]); });
  }
  return PGB_OK;
}

extern "C" int pgb_scenario_constraint_last_ms(double* h2d_ms, double* kernel_ms) {
  if (h2d_ms) *h2d_ms = g_screen_ms[0];
  if (kernel_ms) *kernel_ms = g_screen_ms[1];
  return PGB_OK;
}


// ------------------------------------------------------------------------------------------------
// Device-resident pipeline: sampler -> pgb_samples -> rollout -> screening without a host round trip of the samples
// (the reference reads PG_samples[k].A in place, optimal_control_Ipopt.jl:47-64; VERDICT r1 "missing" items 3 and 5).
// ------------------------------------------------------------------------------------------------
#define NEEDS(s)                                                                  \
  if (!s) return set_err(nullptr, PGB_ERR_INVALID, "sample set is NULL");         \
  pgb_handle* h = &s->hh;                                                         \
  CK(h, cudaSetDevice(h->cfg.device))

extern "C" int pgb_samples_create(pgb_samples** out, const pgb_config* cfg, int64_t K) {
  if (!out) return set_err(nullptr, PGB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = pgb_validate_cfg(cfg);
  if (rc) return rc;
  if (K < 1 || cfg->N < 1) return set_err(nullptr, PGB_ERR_INVALID, "need K >= 1 and N >= 1");
  if (int rcd = pgb_require_device()) return rcd;
  pgb_samples* s = new pgb_samples();
  pgb_handle* h = &s->hh;
  h->cfg = *cfg;
  s->K = s->K_active = K;
  const int nx = cfg->n_x, np = cfg->n_phi, nu = cfg->n_u;
  const int64_t N = cfg->N;
  auto fail = [&](cudaError_t e, const char* what) {
    int r = set_err(nullptr, PGB_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    pgb_samples_destroy(s);
    return r;
  };
  cudaError_t e;
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail(e, "cudaSetDevice");
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  for (auto& evt : s->ev)
    if ((e = cudaEventCreate(&evt)) != cudaSuccess) return fail(e, "cudaEventCreate");
  if ((e = dalloc(h, &s->A, (size_t)K * nx * np)) != cudaSuccess) return fail(e, "cudaMalloc(A)");
  if ((e = dalloc(h, &s->Q, (size_t)K * nx * nx)) != cudaSuccess) return fail(e, "cudaMalloc(Q)");
  if ((e = dalloc(h, &s->w, (size_t)K * N)) != cudaSuccess) return fail(e, "cudaMalloc(w_m1)");
  if ((e = dalloc(h, &s->x, (size_t)K * nx * N)) != cudaSuccess) return fail(e, "cudaMalloc(x_m1)");
  if ((e = dalloc(h, &s->um, (size_t)K * nu)) != cudaSuccess) return fail(e, "cudaMalloc(u_m1)");
  if ((e = dalloc(h, &s->xT, (size_t)K * nx)) != cudaSuccess) return fail(e, "cudaMalloc(x_T)");
  if ((e = dalloc(h, &s->hmax, (size_t)K)) != cudaSuccess) return fail(e, "cudaMalloc(h_max)");
  if ((e = dalloc(h, &s->status, 1)) != cudaSuccess) return fail(e, "cudaMalloc(status)");
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(e, "cudaStreamSynchronize");
  *out = s;
  return PGB_OK;
}

extern "C" void pgb_samples_destroy(pgb_samples* s) {
  if (!s) return;
  pgb_handle* h = &s->hh;
  cudaSetDevice(h->cfg.device);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  for (void* p : h->allocs) cudaFree(p);
  for (auto& evt : s->ev)
    if (evt) cudaEventDestroy(evt);
  delete s;
}

extern "C" int pgb_samples_set(pgb_samples* s, int64_t k0, int64_t n, const double* A, const double* Q, const double* w_m1,
                               const double* x_m1, const double* u_m1) {
  NEEDS(s);
  if (k0 < 0 || n < 1 || k0 + n > s->K) return set_err(h, PGB_ERR_INVALID, "sample range [%lld, %lld) outside 0..%lld",
                                                        (long long)k0, (long long)(k0 + n), (long long)s->K);
  const int nx = h->cfg.n_x, np = h->cfg.n_phi, nu = h->cfg.n_u;
  const int64_t N = h->cfg.N;
  cudaStream_t st = h->stream;
  if (A) CK(h, cudaMemcpyAsync(s->A + (size_t)k0 * nx * np, A, (size_t)n * nx * np * 8, cudaMemcpyHostToDevice, st));
  if (Q) CK(h, cudaMemcpyAsync(s->Q + (size_t)k0 * nx * nx, Q, (size_t)n * nx * nx * 8, cudaMemcpyHostToDevice, st));
  if (w_m1) CK(h, cudaMemcpyAsync(s->w + (size_t)k0 * N, w_m1, (size_t)n * N * 8, cudaMemcpyHostToDevice, st));
  if (x_m1) CK(h, cudaMemcpyAsync(s->x + (size_t)k0 * nx * N, x_m1, (size_t)n * nx * N * 8, cudaMemcpyHostToDevice, st));
  if (u_m1) CK(h, cudaMemcpyAsync(s->um + (size_t)k0 * nu, u_m1, (size_t)n * nu * 8, cudaMemcpyHostToDevice, st));
  CK(h, cudaStreamSynchronize(st));
  if (w_m1 || x_m1) s->have_xT = false;
  if (w_m1 && x_m1) s->have_particles = true;
  return PGB_OK;
}

extern "C" int pgb_samples_get(pgb_samples* s, int64_t k0, int64_t n, double* A, double* Q, double* w_m1, double* x_m1,
                               double* u_m1) {
  NEEDS(s);
  if (k0 < 0 || n < 1 || k0 + n > s->K) return set_err(h, PGB_ERR_INVALID, "sample range outside the set");
  const int nx = h->cfg.n_x, np = h->cfg.n_phi, nu = h->cfg.n_u;
  const int64_t N = h->cfg.N;
  cudaStream_t st = h->stream;
  if (A) CK(h, cudaMemcpyAsync(A, s->A + (size_t)k0 * nx * np, (size_t)n * nx * np * 8, cudaMemcpyDeviceToHost, st));
  if (Q) CK(h, cudaMemcpyAsync(Q, s->Q + (size_t)k0 * nx * nx, (size_t)n * nx * nx * 8, cudaMemcpyDeviceToHost, st));
  if (w_m1) CK(h, cudaMemcpyAsync(w_m1, s->w + (size_t)k0 * N, (size_t)n * N * 8, cudaMemcpyDeviceToHost, st));
  if (x_m1) CK(h, cudaMemcpyAsync(x_m1, s->x + (size_t)k0 * nx * N, (size_t)n * nx * N * 8, cudaMemcpyDeviceToHost, st));
  if (u_m1) CK(h, cudaMemcpyAsync(u_m1, s->um + (size_t)k0 * nu, (size_t)n * nu * 8, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  return PGB_OK;
}

// x_pf[:, :, T] is kept as [n_x][N] by the filter; PG_sample.x_m1 is n_x x N column-major
template <typename RT>
static __global__ void k_soa_to_aos(const RT* __restrict__ src, double* __restrict__ dst, int64_t N, int nx) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  for (int d = 0; d < nx; ++d) dst[n * nx + d] = (double)src[(int64_t)d * N + n];
}

int pgb_samples_store_async(pgb_samples* s, int64_t k, pgb_handle* h, int chain) {
  pgb_handle* sh = &s->hh;
  const pgb_config& a = sh->cfg;
  const pgb_config& b = h->cfg;
  if (a.device != b.device || a.n_x != b.n_x || a.n_phi != b.n_phi || a.n_u != b.n_u || a.N != b.N)
    return set_err(h, PGB_ERR_INVALID, "sample set and handle differ in device / model / N");
  if (k < 0 || k >= s->K || chain < 0 || chain >= b.n_chains) return set_err(h, PGB_ERR_INVALID, "bad slot / chain");
  const int nx = b.n_x, np = b.n_phi, nu = b.n_u;
  const int64_t N = b.N, T = b.T;
  cudaStream_t st = h->stream;
  CK(h, cudaMemcpyAsync(s->A + (size_t)k * nx * np, h->d_A_used + (size_t)chain * nx * np, (size_t)nx * np * 8, cudaMemcpyDeviceToDevice, st));
  CK(h, cudaMemcpyAsync(s->Q + (size_t)k * nx * nx, h->d_Q_used + (size_t)chain * nx * nx, (size_t)nx * nx * 8, cudaMemcpyDeviceToDevice, st));
  CK(h, cudaMemcpyAsync(s->w + (size_t)k * N, h->fd.wbuf + (size_t)chain * N, (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
  CK(h, cudaMemcpyAsync(s->um + (size_t)k * nu, h->d_u + (size_t)nu * (T - 1), (size_t)nu * 8, cudaMemcpyDeviceToDevice, st));
  const size_t row = (h->fd.x_rows == T) ? (size_t)(T - 1) : (size_t)((T - 1) & 1);
  const size_t xoff = ((size_t)chain * h->fd.x_rows + row) * nx * N;
  if (h->fd.f32)  // F32 precision mode: the particle rows hold FP32 values
    k_soa_to_aos<float><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float*>(h->fd.xh) + xoff,
                                                                     s->x + (size_t)k * nx * N, N, nx);
  else
    k_soa_to_aos<double><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(h->fd.xh + xoff, s->x + (size_t)k * nx * N, N, nx);
  CK(h, cudaGetLastError());
  s->have_xT = false;
  s->have_particles = true;
  return PGB_OK;
}

extern "C" int pgb_samples_store(pgb_samples* s, int64_t k, pgb_handle* h, int32_t chain) {
  if (!s) return set_err(nullptr, PGB_ERR_INVALID, "sample set is NULL");
  NEED(h);
  int rc = pgb_samples_store_async(s, k, h, chain);
  if (rc) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return PGB_OK;
}

// x_T[k] = x_m1[:, star_k], star_k ~ systematic_resampling(w_m1, 1) (optimal_control_Ipopt.jl:58-59) drawn where the sample
// lives, with the Philox stream (seed, k_offset + k) the rollout itself would use; a later pgb_rollout_dev with the same
// (seed, k_offset) uses it instead of walking w_m1 again.  x_T (host, may be NULL): [K][n_x].
extern "C" int pgb_samples_draw_xT(pgb_samples* s, uint64_t seed, uint32_t k_offset, double* x_T) {
  NEEDS(s);
  const int64_t K = s->K_active;
  if ((uint64_t)k_offset + (uint64_t)K > (1ull << 24)) return set_err(h, PGB_ERR_INVALID, "k_offset + K exceeds the 2^24 Philox stream ids");
  if (!s->have_particles) return set_err(h, PGB_ERR_STATE, "the set holds no w_m1 / x_m1 to draw the star from");
  CK(h, cudaMemsetAsync(s->status, 0, 4, h->stream));
  k_draw_xT<<<(unsigned)((K + 3) / 4), 128, 0, h->stream>>>(K, h->cfg.N, h->cfg.n_x, s->w, s->x, seed, k_offset, s->xT, s->status);
  CK(h, cudaGetLastError());
  if (x_T) CK(h, cudaMemcpyAsync(x_T, s->xT, (size_t)K * h->cfg.n_x * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  s->have_xT = true;
  return PGB_OK;
}

static int samples_scen_alloc(pgb_samples* s, int64_t H) {
  pgb_handle* h = &s->hh;
  if (H <= s->H_cap) return PGB_OK;
  if (s->H_cap > 0) return set_err(h, PGB_ERR_INVALID, "horizon H = %lld exceeds the %lld this set was first rolled out with",
                                   (long long)H, (long long)s->H_cap);
  const int nx = h->cfg.n_x, ny = h->cfg.n_y, nu = h->cfg.n_u;
  const int64_t K = s->K;
  CK(h, dalloc(h, &s->x0, (size_t)K * nx, false));
  CK(h, dalloc(h, &s->v, (size_t)K * H * nx, false));
  CK(h, dalloc(h, &s->e, (size_t)K * H * ny, false));
  CK(h, dalloc(h, &s->X, (size_t)K * (H + 1) * nx, false));
  CK(h, dalloc(h, &s->Y, (size_t)K * H * ny, false));
  CK(h, dalloc(h, &s->ui, (size_t)nu * H, false));
  CK(h, dalloc(h, &s->ylo, (size_t)ny * H, false));
  CK(h, dalloc(h, &s->yhi, (size_t)ny * H, false));
  s->H_cap = H;
  return PGB_OK;
}

// pgb_rollout on the resident samples (scenario k = sample k).  x0_in / v_in / e_in: host arrays to use instead of the
// draws (the reference's x_vec_0 / v_vec / e_vec kwargs), or NULL; `reuse` bits 1 / 2 / 4 instead keep the x_vec_0 /
// v_vec / e_vec of the PREVIOUS rollout of this set, resident on the device — the pre-solve re-entry of
// optimal_control_Ipopt.jl:88-125 (same noise, new u_init) without moving anything.
extern "C" int pgb_rollout_dev(pgb_samples* s, int64_t H, const double* C, const double* Le, const double* u_init,
                               uint64_t seed, uint32_t k_offset, int32_t reuse, const double* x0_in, const double* v_in,
                               const double* e_in) {
  NEEDS(s);
  if (!C || !Le || !u_init || H < 1) return set_err(h, PGB_ERR_INVALID, "bad arguments");
  if ((uint64_t)k_offset + (uint64_t)s->K_active > (1ull << 24)) return set_err(h, PGB_ERR_INVALID, "k_offset + K exceeds the 2^24 Philox stream ids");
  if (reuse && (!s->have_scen || H != s->H)) return set_err(h, PGB_ERR_STATE, "reuse needs a previous rollout of this set with the same horizon");
  if (!x0_in && !(reuse & 1) && !s->have_xT && !s->have_particles)
    return set_err(h, PGB_ERR_STATE, "the set holds neither particles (w_m1, x_m1) nor x_T: x_vec_0 cannot be drawn; supply x0_in");
  int rc = samples_scen_alloc(s, H);
  if (rc) return rc;
  const int nx = h->cfg.n_x, ny = h->cfg.n_y, nu = h->cfg.n_u;
  const int64_t K = s->K_active;
  cudaStream_t st = h->stream;
  FilterDev fd;
  memset(&fd, 0, sizeof(fd));
  pgb_fill_model(fd, &h->cfg);
  for (int i = 0; i < ny * nx; ++i) fd.C[i] = s->C[i] = C[i];
  RolloutDev rd;
  memset(&rd, 0, sizeof(rd));
  rd.K = K; rd.H = H; rd.N = h->cfg.N; rd.seed = seed; rd.k_offset = k_offset; rd.n_models = 0;
  for (int i = 0; i < ny * ny; ++i) rd.Le[i] = s->Le[i] = Le[i];
  CK(h, cudaEventRecord(s->ev[0], st));
  CK(h, cudaMemcpyAsync(s->ui, u_init, (size_t)nu * H * 8, cudaMemcpyHostToDevice, st));
  if (x0_in) CK(h, cudaMemcpyAsync(s->x0, x0_in, (size_t)K * nx * 8, cudaMemcpyHostToDevice, st));
  if (v_in) CK(h, cudaMemcpyAsync(s->v, v_in, (size_t)K * H * nx * 8, cudaMemcpyHostToDevice, st));
  if (e_in) CK(h, cudaMemcpyAsync(s->e, e_in, (size_t)K * H * ny * 8, cudaMemcpyHostToDevice, st));
  CK(h, cudaMemsetAsync(s->status, 0, 4, st));
  rd.A = s->A; rd.Q = s->Q; rd.w_m1 = s->w; rd.x_m1 = s->x; rd.u_m1 = s->um; rd.u_init = s->ui;
  rd.x_star = s->have_xT ? s->xT : nullptr;
  rd.x0 = s->x0; rd.v = s->v; rd.e = s->e; rd.X = s->X; rd.Y = s->Y; rd.status = s->status;
  rd.x0_in = (x0_in || (reuse & 1)) ? s->x0 : nullptr;
  rd.v_in = (v_in || (reuse & 2)) ? s->v : nullptr;
  rd.e_in = (e_in || (reuse & 4)) ? s->e : nullptr;
  CK(h, cudaEventRecord(s->ev[1], st));
  if ((rc = rollout_launch(h, h->cfg.device, fd, rd, st))) return rc;
  CK(h, cudaEventRecord(s->ev[2], st));
  int stt = 0;
  CK(h, cudaMemcpyAsync(&stt, s->status, 4, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&s->last_ms[2], s->ev[0], s->ev[1]);
  cudaEventElapsedTime(&s->last_ms[0], s->ev[1], s->ev[2]);
  s->H = H;
  s->have_scen = true;
  if (stt & PGB_STATUS_NOT_POSDEF) return set_err(h, PGB_ERR_NOT_POSDEF, "a scenario's Q is not positive definite");
  return PGB_OK;
}

// scenario arrays of the last pgb_rollout_dev -> host (any pointer may be NULL); layouts of pgb_rollout
extern "C" int pgb_scenarios_get(pgb_samples* s, double* x0, double* v, double* e, double* X, double* Y) {
  NEEDS(s);
  if (!s->have_scen) return set_err(h, PGB_ERR_STATE, "no rollout has run on this set");
  const int nx = h->cfg.n_x, ny = h->cfg.n_y;
  const int64_t K = s->K_active, H = s->H;
  cudaStream_t st = h->stream;
  if (x0) CK(h, cudaMemcpyAsync(x0, s->x0, (size_t)K * nx * 8, cudaMemcpyDeviceToHost, st));
  if (v) CK(h, cudaMemcpyAsync(v, s->v, (size_t)K * H * nx * 8, cudaMemcpyDeviceToHost, st));
  if (e) CK(h, cudaMemcpyAsync(e, s->e, (size_t)K * H * ny * 8, cudaMemcpyDeviceToHost, st));
  if (X) CK(h, cudaMemcpyAsync(X, s->X, (size_t)K * (H + 1) * nx * 8, cudaMemcpyDeviceToHost, st));
  if (Y) CK(h, cudaMemcpyAsync(Y, s->Y, (size_t)K * H * ny * 8, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  return PGB_OK;
}

static int screen_order(const double* h_max, int64_t K, int64_t* order) {
  for (int64_t k = 0; k < K; ++k) order[k] = k + 1;
  std::stable_sort(order, order + K, [&](int64_t a, int64_t b) { return jl_isless(h_max[b - 1], h_max[a - 1]); });
  return PGB_OK;
}

// pgb_scenario_constraint_max on the Y the last pgb_rollout_dev left on the device
extern "C" int pgb_scenario_constraint_max_dev(pgb_samples* s, const double* y_min, const double* y_max, double* h_max,
                                               int64_t* order) {
  NEEDS(s);
  if (!y_min || !y_max || !h_max) return set_err(h, PGB_ERR_INVALID, "bad arguments");
  if (!s->have_scen) return set_err(h, PGB_ERR_STATE, "no rollout has run on this set");
  const int64_t K = s->K_active, E = s->H * (int64_t)h->cfg.n_y;
  int64_t n_fin = 0;
  for (int64_t i = 0; i < E; ++i) n_fin += (std::isfinite(y_min[i]) ? 1 : 0) + (std::isfinite(y_max[i]) ? 1 : 0);
  if (n_fin == 0) return set_err(h, PGB_ERR_INVALID, "ArgumentError: no finite output bound: maximum over an empty collection");
  cudaStream_t st = h->stream;
  CK(h, cudaMemcpyAsync(s->ylo, y_min, (size_t)E * 8, cudaMemcpyHostToDevice, st));
  CK(h, cudaMemcpyAsync(s->yhi, y_max, (size_t)E * 8, cudaMemcpyHostToDevice, st));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device);
  const int64_t want = (K + 8 * SCREEN_U - 1) / (8 * SCREEN_U);
  const int grid = (int)std::min<int64_t>(want, (int64_t)sms * 8);
  CK(h, cudaEventRecord(s->ev[0], st));
  k_scenario_constraint_max<<<grid, 256, 0, st>>>(K, E, s->Y, s->ylo, s->yhi, s->hmax);
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(s->ev[1], st));
  CK(h, cudaMemcpyAsync(h_max, s->hmax, (size_t)K * 8, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&s->last_ms[1], s->ev[0], s->ev[1]);
  if (order) screen_order(h_max, K, order);
  return PGB_OK;
}

extern "C" int pgb_samples_last_ms(pgb_samples* s, double* rollout_kernel_ms, double* screening_kernel_ms, double* upload_ms) {
  if (!s) return set_err(nullptr, PGB_ERR_INVALID, "sample set is NULL");
  if (rollout_kernel_ms) *rollout_kernel_ms = s->last_ms[0];
  if (screening_kernel_ms) *screening_kernel_ms = s->last_ms[1];
  if (upload_ms) *upload_ms = s->last_ms[2];
  return PGB_OK;
}
extern "C" const char* pgb_samples_last_error(const pgb_samples* s) { return pgb_last_error(s ? &s->hh : nullptr); }

// Page-locked host memory for the callers' I/O arrays (a pageable cudaMemcpy of the C5 inputs was 261 ms of a 0.74 s call)
extern "C" void* pgb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
extern "C" void pgb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" void pgb_test_rollout_kernel(int32_t which) { g_rollout_force = which; }
