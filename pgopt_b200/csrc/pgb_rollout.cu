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
static cudaError_t launch_rollout_ni(const FilterDev& fd, const RolloutDev& rd, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(k_rollout<NX, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_rollout<NX, NI><<<(unsigned)rd.K, R_TPB, smem>>>(fd, rd);
  return cudaGetLastError();
}
// n_phi <= 128: one register-resident column of A per thread; otherwise five (n_phi <= 640 entirely in registers,
// e.g. 5^4 = 625), any further columns are read from L2
template <int NX>
static cudaError_t launch_rollout(const FilterDev& fd, const RolloutDev& rd, size_t smem) {
  return fd.n_phi <= R_TPB ? launch_rollout_ni<NX, 1>(fd, rd, smem) : launch_rollout_ni<NX, 5>(fd, rd, smem);
}

// Warp-per-scenario kernel (default): A_k resident over 32 lanes (RS of each lane's 4 runs in shared memory, the
// rest in registers), persistent grid of independent warps.
static size_t rollout_warp_smem(int nx, int ny, int nu, int64_t H, int ni, int rs) {
  return sizeof(RolloutWShared) +
         ((size_t)H * nu + (size_t)RW_WPB * (PGB_MAX_Z_ * MAXC + (size_t)H * (nx + ny) + (size_t)rs * ni * nx * 32 + (size_t)(H + 2) * nx)) * 8;
}
template <int NX, int NI, int RS>
static cudaError_t launch_rollout_warp_ni(const FilterDev& fd, const RolloutDev& rd, int n_sm) {
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
  k_rollout_warp<NX, NI, RS><<<(unsigned)grid, RW_WPB * 32, smem>>>(fd, rd);
  return cudaGetLastError();
}
template <int NX>
static cudaError_t launch_rollout_warp(const FilterDev& fd, const RolloutDev& rd, int n_sm, int rs) {
  if (fd.n_phi <= 128) return launch_rollout_warp_ni<NX, 1, 0>(fd, rd, n_sm);
  (void)rs;
  return launch_rollout_warp_ni<NX, RW_MAXNI, 2>(fd, rd, n_sm);
}

// Shared body of pgb_rollout (n_models = 0, y_test = null) and pgb_test_prediction (K scenarios over n_models models,
// scenario k simulating model k mod n_models; sum_sq_err = tree sum of (Y - y_test)^2).
static int rollout_run(const pgb_config* cfg, int64_t K, int64_t n_models, int64_t H, const double* A, const double* Q,
                       const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                       const double* u_init, uint64_t seed, uint32_t k_offset, double* x0, double* v, double* e,
                       double* X, double* Y, const double* y_test, double* sum_sq_err) {
  int rc = pgb_validate_cfg(cfg);
  if (rc) return rc;
  const int64_t KM = n_models > 0 ? n_models : K;  // number of models (PG samples) behind the K scenarios
  if (!A || !Q || !w_m1 || !x_m1 || !u_m1 || !C || !Le || !u_init || K < 1 || H < 1 || cfg->N < 1)
    return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
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
  const size_t smem = sizeof(RolloutShared) + ((size_t)H * (nx + ny + nu) + (size_t)(N < R_WSTAGE ? N : R_WSTAGE)) * 8;
  if (smem > 200 * 1024 || K > 0x7fffffffLL)
    return set_err(nullptr, PGB_ERR_INVALID, "rollout: horizon H = %lld (or K) too large", (long long)H);
  double *dA, *dQ, *dw, *dx, *dum, *dui;
  cudaEvent_t ev[3];
  for (auto& evt : ev) CK(h, cudaEventCreate(&evt));
  CK(h, cudaEventRecord(ev[0], 0));
  CK(h, dalloc(h, &dA, (size_t)KM * m, false));
  CK(h, dalloc(h, &dQ, (size_t)KM * nx * nx, false));
  CK(h, dalloc(h, &dw, (size_t)KM * N, false));
  CK(h, dalloc(h, &dx, (size_t)KM * nx * N, false));
  CK(h, dalloc(h, &dum, (size_t)KM * nu, false));
  CK(h, dalloc(h, &dui, (size_t)nu * H, false));
  CK(h, dalloc(h, &rd.status, 1));
  CK(h, cudaMemcpy(dA, A, (size_t)KM * m * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(dQ, Q, (size_t)KM * nx * nx * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(dw, w_m1, (size_t)KM * N * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(dx, x_m1, (size_t)KM * nx * N * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(dum, u_m1, (size_t)KM * nu * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(dui, u_init, (size_t)nu * H * 8, cudaMemcpyHostToDevice));
  if (x0) CK(h, dalloc(h, &rd.x0, (size_t)K * nx, false));
  if (v) CK(h, dalloc(h, &rd.v, (size_t)K * H * nx, false));
  if (e) CK(h, dalloc(h, &rd.e, (size_t)K * H * ny, false));
  if (X) CK(h, dalloc(h, &rd.X, (size_t)K * (H + 1) * nx, false));
  if (Y || y_test) CK(h, dalloc(h, &rd.Y, (size_t)K * H * ny, false));
  CK(h, cudaEventRecord(ev[1], 0));
  rd.A = dA; rd.Q = dQ; rd.w_m1 = dw; rd.x_m1 = dx; rd.u_m1 = dum; rd.u_init = dui;
  // kernel choice: one warp per scenario whenever A_k fits the registers of 32 lanes (n_phi <= 128 * RW_MAXNI) and
  // the per-warp noise buffers fit shared memory; pgb_test_rollout_kernel forces one (tests run both)
  const int rs = 2;  // runs (of 4 per lane) of A_k kept in shared memory (0 / 2 / 3 measured: 4.8 / 4.7 / 5.1 ms at C5)
  const size_t smem_w = rollout_warp_smem(nx, ny, nu, H, np <= 128 ? 1 : RW_MAXNI, np <= 128 ? 0 : rs);
  bool use_warp = np <= 128 * RW_MAXNI && smem_w <= 100 * 1024;
  if (g_rollout_force == 0) use_warp = false;  // test hook: the CTA-per-scenario kernel on a shape the warp kernel covers
  else if (g_rollout_force == 1 && !use_warp)
    return set_err(nullptr, PGB_ERR_INVALID, "rollout: warp kernel needs n_phi <= %d and a shorter horizon", 128 * RW_MAXNI);
  g_rollout_kernel = use_warp ? 1 : 0;
  if (use_warp) {
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device);
    switch (nx) {
      case 1: CK(h, launch_rollout_warp<1>(fd, rd, n_sm, rs)); break;
      case 2: CK(h, launch_rollout_warp<2>(fd, rd, n_sm, rs)); break;
      case 3: CK(h, launch_rollout_warp<3>(fd, rd, n_sm, rs)); break;
      default: CK(h, launch_rollout_warp<4>(fd, rd, n_sm, rs)); break;
    }
  } else {
    switch (nx) {
      case 1: CK(h, launch_rollout<1>(fd, rd, smem)); break;
      case 2: CK(h, launch_rollout<2>(fd, rd, smem)); break;
      case 3: CK(h, launch_rollout<3>(fd, rd, smem)); break;
      default: CK(h, launch_rollout<4>(fd, rd, smem)); break;
    }
  }
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
  for (auto& evt : ev) cudaEventDestroy(evt);
  int st = 0;
  CK(h, cudaMemcpy(&st, rd.status, 4, cudaMemcpyDeviceToHost));
  if (st & PGB_STATUS_NOT_POSDEF) return set_err(nullptr, PGB_ERR_NOT_POSDEF, "a scenario's Q is not positive definite");
  if (x0) CK(h, cudaMemcpy(x0, rd.x0, (size_t)K * nx * 8, cudaMemcpyDeviceToHost));
  if (v) CK(h, cudaMemcpy(v, rd.v, (size_t)K * H * nx * 8, cudaMemcpyDeviceToHost));
  if (e) CK(h, cudaMemcpy(e, rd.e, (size_t)K * H * ny * 8, cudaMemcpyDeviceToHost));
  if (X) CK(h, cudaMemcpy(X, rd.X, (size_t)K * (H + 1) * nx * 8, cudaMemcpyDeviceToHost));
  if (Y) CK(h, cudaMemcpy(Y, rd.Y, (size_t)K * H * ny * 8, cudaMemcpyDeviceToHost));
  return PGB_OK;
}

extern "C" int pgb_rollout(const pgb_config* cfg, int64_t K, int64_t H, const double* A, const double* Q,
                           const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                           const double* u_init, uint64_t seed, uint32_t k_offset, double* x0, double* v, double* e,
                           double* X, double* Y) {
  return rollout_run(cfg, K, 0, H, A, Q, w_m1, x_m1, u_m1, C, Le, u_init, seed, k_offset, x0, v, e, X, Y, nullptr, nullptr);
}

extern "C" int pgb_test_prediction(const pgb_config* cfg, int64_t K, int64_t k_n, int64_t T_test, const double* A,
                                   const double* Q, const double* w_m1, const double* x_m1, const double* u_m1,
                                   const double* C, const double* Le, const double* u_test, const double* y_test,
                                   uint64_t seed, double* x_sim, double* y_sim, double* mean_rmse) {
  if (K < 1 || k_n < 1 || T_test < 1 || !y_test || !mean_rmse) return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  double sse = 0.0;
  int rc = rollout_run(cfg, K * k_n, K, T_test, A, Q, w_m1, x_m1, u_m1, C, Le, u_test, seed, 0, nullptr, nullptr, nullptr,
                       x_sim, y_sim, y_test, &sse);
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

extern "C" void pgb_test_rollout_kernel(int32_t which) { g_rollout_force = which; }
