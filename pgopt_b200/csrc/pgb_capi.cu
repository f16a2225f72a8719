// pgb_capi.cu — host side of libpgopt_b200.so: the C ABI of include/pgopt_b200.h, device memory
// management and the launch sequences.  No torch types, no CPU compute path: every entry point that
// computes needs a CUDA device and fails with PGB_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pgopt_b200.h"
#include "pgb_mniw.cuh"
#include "pgb_rollout.cuh"
#include "pgb_persist.cuh"
#include "pgb_persist_w.cuh"
#include "pgb_small.cuh"

using namespace pgb;

static thread_local std::string g_last_error;
static int g_force_fallback = 0;

struct pgb_handle {
  pgb_config cfg;
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;
  std::string err;
  FilterDev fd;
  MniwDev md;
  bool have_data = false, have_meas = false, have_init = false, have_prior = false, have_rng = false;
  std::vector<char> have_state;
  int64_t sweep_index = 1;
  // device
  double *d_u = nullptr, *d_y = nullptr, *d_Q = nullptr, *d_Vdiag = nullptr, *d_LambdaQ = nullptr;
  double *d_Z0 = nullptr, *d_Z = nullptr, *d_Ures = nullptr, *d_Uanc = nullptr, *d_Ustar = nullptr;
  double *d_bart = nullptr, *d_Xmn = nullptr;
  int32_t* d_star_idx = nullptr;
  int64_t* d_star_out = nullptr;
  long long* d_counters_main = nullptr;
  long long* d_counters_side = nullptr;
  std::vector<double> h_u;  // host copy of u (for u_m1)
  // the (A, Q) the last filter pass ran with (sample fields, particle_Gibbs.jl:204-205)
  double *d_A_used = nullptr, *d_Q_used = nullptr;
  double* d_replay_noise = nullptr;  // [nc][T][n_x] noise along the traced lineage (replay mode)
  bool injected_ready = false;
  // persistent (single cooperative kernel) filter path
  int persistent = -1;  // -1: automatic (warp-per-chain kernel when N <= 32, else the tile persistent kernel), 1: tile persistent kernel, 2: warp-tile persistent kernel, 0: one launch per phase, 3: warp-per-chain kernel (N <= 32 only)
  PersistDev pd;
  int num_sms = 0;
  // measurement
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool profile = false;
  int64_t n_launch = 0;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_class;   // class of pair i (events 2i, 2i+1)
  size_t ev_used = 0;
  int64_t prof_launches[PGB_PROF_NCLASS] = {0};
  double prof_ms[PGB_PROF_NCLASS] = {0};
};

// Bracket one launch: counts it and, in profiling mode, records an event pair around it.
struct LaunchScope {
  pgb_handle* h;
  cudaEvent_t e1 = nullptr;
  LaunchScope(pgb_handle* h_, int cls) : h(h_) {
    h->n_launch += 1;
    if (!h->profile) return;
    if (h->ev_used + 2 > h->ev_pool.size()) {
      for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->ev_pool.push_back(e);
      }
    }
    cudaEvent_t e0 = h->ev_pool[h->ev_used];
    e1 = h->ev_pool[h->ev_used + 1];
    h->ev_used += 2;
    h->ev_class.push_back(cls);
    cudaEventRecord(e0, h->stream);
  }
  ~LaunchScope() {
    if (e1) cudaEventRecord(e1, h->stream);
  }
};
#define LAUNCH(h, cls, ...) do { LaunchScope ls_(h, cls); __VA_ARGS__; } while (0)

static void profile_collect(pgb_handle* h) {
  if (!h->ev_used) return;
  cudaStreamSynchronize(h->stream);
  for (size_t i = 0; i < h->ev_used / 2; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]);
    h->prof_launches[h->ev_class[i]] += 1;
    h->prof_ms[h->ev_class[i]] += ms;
  }
  h->ev_used = 0;
  h->ev_class.clear();
}

static int set_err(pgb_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (h) h->err = buf;
  return code;
}

#define CK(h, call)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return set_err(h, PGB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <typename T>
static cudaError_t dalloc(pgb_handle* h, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (e != cudaSuccess) return e;
  if (h) h->allocs.push_back(q);
  if (zero) {
    e = cudaMemset(q, 0, count * sizeof(T));
    if (e != cudaSuccess) return e;
  }
  *p = (T*)q;
  return cudaSuccess;
}

extern "C" const char* pgb_version(void) { return "pgopt_b200 0.1.0 (sm_100a)"; }
extern "C" const char* pgb_last_error(const pgb_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

static cudaError_t alloc_scanws(pgb_handle* h, ScanWS* ws, int64_t N, int nc, long long** counters) {
  cudaError_t e;
  memset(ws, 0, sizeof(*ws));
  ws->N = N;
  ws->nt = (int)((N + TILE - 1) / TILE);
  const int64_t nt = ws->nt;
#define A_(field, count) if ((e = dalloc(h, &ws->field, (size_t)(count))) != cudaSuccess) return e;
  A_(S, nc) A_(tsum, nc * nt) A_(tile_prefix, nc * (nt + 1)) A_(thr_ex, nc * nt * TPB) A_(tile_agg, nc * nt)
  A_(tile_ex, nc * (nt + 1)) A_(recs, nc * nt * RMAX) A_(crec, (int64_t)nc * CMAX) A_(seeds, (int64_t)nc * (CMAX + 1))
  A_(qin_fb, nc * nt * TPB) A_(Wn, nc * nt * TILE) A_(ticket, nc) A_(flag, 2 * nc) A_(pend, nc) A_(counters, 2 * nc)
#undef A_
  if (counters) *counters = ws->counters;
  return cudaSuccess;
}

static int validate_cfg(const pgb_config* c) {
  if (!c) return set_err(nullptr, PGB_ERR_INVALID, "config is NULL");
  if (c->n_x < 1 || c->n_x > 4) return set_err(nullptr, PGB_ERR_INVALID, "n_x must be 1..4 (got %d)", c->n_x);
  if (c->n_u < 1 || c->n_u > 4) return set_err(nullptr, PGB_ERR_INVALID, "n_u must be 1..4 (got %d)", c->n_u);
  if (c->n_y < 1 || c->n_y > 4) return set_err(nullptr, PGB_ERR_INVALID, "n_y must be 1..4 (got %d)", c->n_y);
  if (c->basis == PGB_BASIS_KNOWN_VB) {
    if (c->n_x != 2 || c->n_u != 1 || c->n_phi != 5)
      return set_err(nullptr, PGB_ERR_INVALID, "KNOWN_VB basis needs n_x=2, n_u=1, n_phi=5");
  } else if (c->basis == PGB_BASIS_HILBERT_GP) {
    int nz = c->n_x + c->n_u;
    if (nz > PGB_MAX_Z) return set_err(nullptr, PGB_ERR_INVALID, "n_x + n_u > %d", PGB_MAX_Z);
    int64_t stride = 1;
    for (int k = nz - 1; k >= 0; --k) {
      if (c->hg_count[k] < 1 || c->hg_count[k] > MAXC)
        return set_err(nullptr, PGB_ERR_INVALID, "hg_count[%d] must be 1..%d", k, MAXC);
      if (c->hg_stride[k] != stride)
        return set_err(nullptr, PGB_ERR_INVALID, "hg_stride must be the row-major mixed radix of hg_count (dim %d)", k);
      if (!(c->hg_L[k] > 0.0)) return set_err(nullptr, PGB_ERR_INVALID, "hg_L[%d] must be > 0", k);
      stride *= c->hg_count[k];
    }
    if (stride != c->n_phi) return set_err(nullptr, PGB_ERR_INVALID, "n_phi != prod(hg_count)");
  } else {
    return set_err(nullptr, PGB_ERR_INVALID, "unknown basis id %d (arbitrary closures are not supported)", c->basis);
  }
  return PGB_OK;
}

static void fill_model(FilterDev& fd, const pgb_config* c) {
  fd.n_x = c->n_x; fd.n_u = c->n_u; fd.n_y = c->n_y; fd.n_phi = c->n_phi; fd.basis = c->basis;
  fd.n_z = c->n_x + c->n_u;
  if (c->basis == PGB_BASIS_HILBERT_GP) {
    const double PI = 3.141592653589793;
    for (int k = 0; k < fd.n_z; ++k) {
      fd.hg_count[k] = c->hg_count[k];
      fd.hg_L[k] = c->hg_L[k];
      volatile double sq = std::sqrt(c->hg_L[k]);
      fd.hg_lsi[k] = 1.0 / sq;  // L_sqrt_inv = 1 ./ sqrt.(L)   (generic example :85)
      for (int j = 0; j < c->hg_count[k]; ++j) {
        volatile double num = PI * (double)(j + 1);  // pi .* j_vec ./ (2 .* L)   (:86)
        volatile double den = 2.0 * c->hg_L[k];
        fd.hg_pij2l[k][j] = num / den;
      }
    }
  }
}

extern "C" int pgb_create(pgb_handle** out, const pgb_config* cfg) {
  if (!out) return set_err(nullptr, PGB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  if (cfg->N < 2 || cfg->T < 2 || cfg->n_chains < 1)
    return set_err(nullptr, PGB_ERR_INVALID, "need N >= 2, T >= 2, n_chains >= 1");
  if (cfg->N >= (1ll << 31) - TILE) return set_err(nullptr, PGB_ERR_INVALID, "N too large for Int32 ancestors");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return set_err(nullptr, PGB_ERR_CUDA, "no CUDA device available (%s); pgopt_b200 has no CPU fallback",
                   cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return set_err(nullptr, PGB_ERR_INVALID, "bad device ordinal %d", cfg->device);
  pgb_handle* h = new pgb_handle();
  h->cfg = *cfg;
#define CKC(call)                                                                       \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      int rc_ = set_err(nullptr, PGB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
      pgb_destroy(h);                                                                   \
      return rc_;                                                                       \
    }                                                                                   \
  } while (0)
  CKC(cudaSetDevice(cfg->device));
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const int64_t N = cfg->N, T = cfg->T;
  const int nc = cfg->n_chains, nx = cfg->n_x, np = cfg->n_phi, nu = cfg->n_u, ny = cfg->n_y;
  FilterDev& fd = h->fd;
  memset(&fd, 0, sizeof(fd));
  fill_model(fd, cfg);
  fd.N = N; fd.T = T; fd.nc = nc;
  fd.nt = (int)((N + TILE - 1) / TILE);
  const int64_t nt = fd.nt;
  CKC(dalloc(h, &h->d_u, (size_t)nu * T));
  CKC(dalloc(h, &h->d_y, (size_t)ny * T));
  fd.u = h->d_u; fd.y = h->d_y;
  CKC(dalloc(h, &fd.A, (size_t)nc * nx * np));
  CKC(dalloc(h, &h->d_Q, (size_t)nc * nx * nx));
  CKC(dalloc(h, &h->d_A_used, (size_t)nc * nx * np));
  CKC(dalloc(h, &h->d_Q_used, (size_t)nc * nx * nx));
  CKC(dalloc(h, &fd.Lq, (size_t)nc * 16));
  CKC(dalloc(h, &fd.c0, (size_t)nc));
  CKC(dalloc(h, &fd.xprim, (size_t)nc * nx * T));
  {  // x history: kept (needed by pgb_get_history) or rolling two rows + replay of the traced lineage
    const double hist_bytes = (double)nc * (double)T * nx * (double)N * 8.0;
    const bool store = cfg->x_history == PGB_X_HISTORY_STORE ||
                       (cfg->x_history == PGB_X_HISTORY_AUTO && hist_bytes <= 1024.0 * 1024.0 * 1024.0);
    fd.x_rows = store ? T : 2;
    if (T < 2) fd.x_rows = T;
  }
  CKC(dalloc(h, &fd.xh, (size_t)nc * fd.x_rows * nx * N, false));
  CKC(dalloc(h, &fd.lineage, (size_t)nc * T));
  CKC(dalloc(h, &h->d_replay_noise, (size_t)nc * T * nx));
  CKC(dalloc(h, &fd.ah, (size_t)nc * T * N, false));
  CKC(cudaMemset(fd.ah, 0xff, (size_t)nc * T * N * sizeof(int32_t)));
  CKC(dalloc(h, &fd.wbuf, (size_t)nc * N));
  CKC(dalloc(h, &fd.ebuf, (size_t)nc * N));
  CKC(dalloc(h, &fd.lwbuf, (size_t)nc * N));
  CKC(dalloc(h, &fd.Fbuf, (size_t)nc * nx * N));
  CKC(dalloc(h, &fd.waN, (size_t)nc * N));
  if (cfg->keep_w_history) CKC(dalloc(h, &fd.whist, (size_t)nc * T * N));
  CKC(dalloc(h, &fd.maxkey, (size_t)nc));
  CKC(dalloc(h, &fd.status, (size_t)nc));
  CKC(dalloc(h, &fd.S1, (size_t)nc));
  CKC(dalloc(h, &fd.S3, (size_t)nc));
  CKC(dalloc(h, &fd.tsum_e, (size_t)nc * nt));
  CKC(dalloc(h, &fd.tsum_a, (size_t)nc * nt));
  CKC(dalloc(h, &fd.ticket_e, (size_t)nc));
  CKC(dalloc(h, &fd.ticket_a, (size_t)nc));
  CKC(dalloc(h, &fd.utab_res, (size_t)nc * T, false));
  CKC(dalloc(h, &fd.utab_anc, (size_t)nc * T, false));
  CKC(dalloc(h, &fd.utab_star, (size_t)nc, false));
  CKC(alloc_scanws(h, &fd.main, N, nc, &h->d_counters_main));
  CKC(alloc_scanws(h, &fd.side, N, nc, &h->d_counters_side));
  CKC(dalloc(h, &h->d_star_idx, (size_t)nc));
  CKC(dalloc(h, &h->d_star_out, (size_t)nc));
  // MNIW
  MniwDev& md = h->md;
  memset(&md, 0, sizeof(md));
  md.nc = nc; md.n_x = nx; md.n_phi = np; md.Tm1 = T - 1;
  CKC(dalloc(h, &h->d_Vdiag, (size_t)np));
  CKC(dalloc(h, &h->d_LambdaQ, (size_t)nx * nx));
  md.V_diag = h->d_Vdiag; md.Lambda_Q = h->d_LambdaQ;
  CKC(dalloc(h, &md.zeta, (size_t)nc * nx * (T - 1)));
  CKC(dalloc(h, &md.zb, (size_t)nc * np * (T - 1)));
  CKC(dalloc(h, &md.Phi, (size_t)nc * nx * nx));
  CKC(dalloc(h, &md.Psi, (size_t)nc * nx * np));
  CKC(dalloc(h, &md.Sigma, (size_t)nc * np * np));
  CKC(dalloc(h, &md.Ls, (size_t)nc * np * np));
  CKC(dalloc(h, &md.Sinv, (size_t)nc * np * np));
  CKC(dalloc(h, &md.Lv, (size_t)nc * np * np));
  CKC(dalloc(h, &md.Mpost, (size_t)nc * nx * np));
  CKC(dalloc(h, &md.Xn, (size_t)nc * nx * np));
  CKC(dalloc(h, &md.LX, (size_t)nc * nx * np));
  md.A = fd.A; md.Q = h->d_Q; md.Lq = fd.Lq; md.c0 = fd.c0; md.status = fd.status;
  {
    PersistDev& pd = h->pd;
    memset(&pd, 0, sizeof(pd));
    CKC(dalloc(h, &pd.bar, 4));
    CKC(dalloc(h, &pd.pe, (size_t)nc * P_MAXNB));
    CKC(dalloc(h, &pd.pw, (size_t)nc * P_MAXNB));
    CKC(dalloc(h, &pd.pa, (size_t)nc * P_MAXNB));
    CKC(dalloc(h, &pd.pa2, (size_t)nc * P_MAXNB));
    pd.star_idx = h->d_star_idx;
    CKC(dalloc(h, &pd.heavy, (size_t)nc * P_HEAVY_MAX));
    CKC(dalloc(h, &pd.heavy_cnt, (size_t)nc));
    CKC(dalloc(h, &pd.prof, (size_t)1024 * P_NPROF));
    CKC(dalloc(h, &pd.cand, (size_t)2 * nc));
    CKC(dalloc(h, &pd.sm_slots, (size_t)1024));
    CKC(dalloc(h, &pd.cta_key, (size_t)4096));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, cfg->device));
    h->num_sms = prop.multiProcessorCount;
    const char* abl = getenv("PGB_ABLATE");
    fd.ablate = abl ? atoi(abl) : 0;
    const char* env = getenv("PGB_PERSISTENT");
    if (env && env[0] >= '0' && env[0] <= '3') h->persistent = env[0] - '0';
  }
  h->have_state.assign(nc, 0);
  h->h_u.assign((size_t)nu * T, 0.0);
#undef CKC
  *out = h;
  return PGB_OK;
}

extern "C" void pgb_destroy(pgb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  for (void* p : h->allocs) cudaFree(p);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
}

#define NEED(h) \
  if (!h) return set_err(nullptr, PGB_ERR_INVALID, "handle is NULL"); \
  CK(h, cudaSetDevice(h->cfg.device))

extern "C" int pgb_set_data(pgb_handle* h, const double* u, const double* y) {
  NEED(h);
  if (!u || !y) return set_err(h, PGB_ERR_INVALID, "u/y NULL");
  const size_t nu = (size_t)h->cfg.n_u * h->cfg.T, ny = (size_t)h->cfg.n_y * h->cfg.T;
  CK(h, cudaMemcpyAsync(h->d_u, u, nu * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->d_y, y, ny * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  memcpy(h->h_u.data(), u, nu * 8);
  h->have_data = true;
  return PGB_OK;
}

static int host_chol(double* A, int n) {
  for (int k = 0; k < n; ++k) {
    double d = A[k + n * k];
    if (!(d > 0.0)) return 1;
    d = std::sqrt(d);
    A[k + n * k] = d;
    for (int i = k + 1; i < n; ++i) A[i + n * k] = A[i + n * k] / d;
    for (int j = k + 1; j < n; ++j)
      for (int i = j; i < n; ++i) A[i + n * j] = std::fma(-A[i + n * k], A[j + n * k], A[i + n * j]);
  }
  for (int j = 1; j < n; ++j)
    for (int i = 0; i < j; ++i) A[i + n * j] = 0.0;
  return 0;
}

extern "C" int pgb_set_measurement(pgb_handle* h, const double* C, const double* R) {
  NEED(h);
  if (!C || !R) return set_err(h, PGB_ERR_INVALID, "C/R NULL");
  const int nx = h->cfg.n_x, ny = h->cfg.n_y;
  memset(h->fd.C, 0, sizeof(h->fd.C));
  memset(h->fd.R, 0, sizeof(h->fd.R));
  memset(h->fd.Rchol, 0, sizeof(h->fd.Rchol));
  for (int i = 0; i < ny * nx; ++i) h->fd.C[i] = C[i];
  for (int i = 0; i < ny * ny; ++i) h->fd.R[i] = h->fd.Rchol[i] = R[i];
  if (host_chol(h->fd.Rchol, ny)) return set_err(h, PGB_ERR_NOT_POSDEF, "R is not positive definite");
  h->have_meas = true;
  return PGB_OK;
}

extern "C" int pgb_set_init_dist(pgb_handle* h, const double* mean, const double* chol_lower) {
  NEED(h);
  if (!mean || !chol_lower) return set_err(h, PGB_ERR_INVALID, "mean/chol NULL");
  const int nx = h->cfg.n_x;
  memset(h->fd.x0_mean, 0, sizeof(h->fd.x0_mean));
  memset(h->fd.x0_chol, 0, sizeof(h->fd.x0_chol));
  for (int i = 0; i < nx; ++i) h->fd.x0_mean[i] = mean[i];
  for (int i = 0; i < nx * nx; ++i) h->fd.x0_chol[i] = chol_lower[i];
  h->have_init = true;
  return PGB_OK;
}

extern "C" int pgb_set_prior(pgb_handle* h, const double* V_diag, const double* Lambda_Q, double ell_Q) {
  NEED(h);
  if (!V_diag || !Lambda_Q) return set_err(h, PGB_ERR_INVALID, "V_diag/Lambda_Q NULL");
  CK(h, cudaMemcpyAsync(h->d_Vdiag, V_diag, (size_t)h->cfg.n_phi * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->d_LambdaQ, Lambda_Q, (size_t)h->cfg.n_x * h->cfg.n_x * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->md.ell_Q = ell_Q;
  h->have_prior = true;
  return PGB_OK;
}

extern "C" int pgb_set_state(pgb_handle* h, int32_t chain, const double* A, const double* Q, const double* x_prim,
                             int64_t sweep_index) {
  NEED(h);
  const int nc = h->cfg.n_chains, nx = h->cfg.n_x, np = h->cfg.n_phi;
  const int64_t T = h->cfg.T;
  if (chain < 0 || chain >= nc) return set_err(h, PGB_ERR_INVALID, "bad chain %d", chain);
  if (!A || !Q) return set_err(h, PGB_ERR_INVALID, "A/Q NULL");
  if (sweep_index < 1) return set_err(h, PGB_ERR_INVALID, "sweep_index must be >= 1");
  CK(h, cudaMemcpyAsync(h->fd.A + (size_t)chain * nx * np, A, (size_t)nx * np * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->d_Q + (size_t)chain * nx * nx, Q, (size_t)nx * nx * 8, cudaMemcpyHostToDevice, h->stream));
  if (x_prim)
    CK(h, cudaMemcpyAsync(h->fd.xprim + (size_t)chain * nx * T, x_prim, (size_t)nx * T * 8, cudaMemcpyHostToDevice, h->stream));
  else
    CK(h, cudaMemsetAsync(h->fd.xprim + (size_t)chain * nx * T, 0, (size_t)nx * T * 8, h->stream));
  k_chain_params<<<1, 32, 0, h->stream>>>(chain, nx, h->d_Q, h->fd.Lq, h->fd.c0, h->fd.status);
  CK(h, cudaGetLastError());
  CK(h, cudaStreamSynchronize(h->stream));
  h->sweep_index = sweep_index;
  h->have_state[chain] = 1;
  return PGB_OK;
}

extern "C" int pgb_get_state(pgb_handle* h, int32_t chain, double* A, double* Q, double* x_prim, int64_t* sweep_index) {
  NEED(h);
  const int nc = h->cfg.n_chains, nx = h->cfg.n_x, np = h->cfg.n_phi;
  const int64_t T = h->cfg.T;
  if (chain < 0 || chain >= nc) return set_err(h, PGB_ERR_INVALID, "bad chain %d", chain);
  CK(h, cudaStreamSynchronize(h->stream));
  if (A) CK(h, cudaMemcpy(A, h->fd.A + (size_t)chain * nx * np, (size_t)nx * np * 8, cudaMemcpyDeviceToHost));
  if (Q) CK(h, cudaMemcpy(Q, h->d_Q + (size_t)chain * nx * nx, (size_t)nx * nx * 8, cudaMemcpyDeviceToHost));
  if (x_prim) CK(h, cudaMemcpy(x_prim, h->fd.xprim + (size_t)chain * nx * T, (size_t)nx * T * 8, cudaMemcpyDeviceToHost));
  if (sweep_index) *sweep_index = h->sweep_index;
  return PGB_OK;
}

extern "C" int pgb_set_rng_philox(pgb_handle* h, uint64_t seed, uint32_t chain_base) {
  NEED(h);
  h->fd.philox = 1; h->fd.seed = seed; h->fd.chain_base = chain_base;
  h->md.philox = 1; h->md.seed = seed; h->md.chain_base = chain_base;
  h->have_rng = true;
  return PGB_OK;
}

extern "C" int pgb_set_rng_injected(pgb_handle* h, int32_t chain, const double* Z0, const double* Ures, const double* Z,
                                    const double* Uanc, double Ustar, const double* bartlett, const double* Xmn) {
  NEED(h);
  const int nc = h->cfg.n_chains, nx = h->cfg.n_x, np = h->cfg.n_phi;
  const int64_t N = h->cfg.N, T = h->cfg.T;
  if (chain < 0 || chain >= nc) return set_err(h, PGB_ERR_INVALID, "bad chain %d", chain);
  if (!Z0 || !Ures || !Z || !Uanc) return set_err(h, PGB_ERR_INVALID, "injected stream pointer NULL");
  if (!h->d_Z0) {
    CK(h, dalloc(h, &h->d_Z0, (size_t)nc * nx * (N - 1)));
    CK(h, dalloc(h, &h->d_Z, (size_t)nc * T * nx * N));
    CK(h, dalloc(h, &h->d_Ures, (size_t)nc * T));
    CK(h, dalloc(h, &h->d_Uanc, (size_t)nc * T));
    CK(h, dalloc(h, &h->d_Ustar, (size_t)nc));
    CK(h, dalloc(h, &h->d_bart, (size_t)nc * nx * nx));
    CK(h, dalloc(h, &h->d_Xmn, (size_t)nc * nx * np));
  }
  CK(h, cudaMemcpy(h->d_Z0 + (size_t)chain * nx * (N - 1), Z0, (size_t)nx * (N - 1) * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(h->d_Z + (size_t)chain * T * nx * N, Z, (size_t)T * nx * N * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(h->d_Ures + (size_t)chain * T, Ures, (size_t)T * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(h->d_Uanc + (size_t)chain * T, Uanc, (size_t)T * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(h->d_Ustar + chain, &Ustar, 8, cudaMemcpyHostToDevice));
  if (bartlett) CK(h, cudaMemcpy(h->d_bart + (size_t)chain * nx * nx, bartlett, (size_t)nx * nx * 8, cudaMemcpyHostToDevice));
  if (Xmn) CK(h, cudaMemcpy(h->d_Xmn + (size_t)chain * nx * np, Xmn, (size_t)nx * np * 8, cudaMemcpyHostToDevice));
  h->fd.philox = 0;
  h->fd.Z0 = h->d_Z0; h->fd.Z = h->d_Z; h->fd.Ures = h->d_Ures; h->fd.Uanc = h->d_Uanc; h->fd.Ustar = h->d_Ustar;
  h->md.philox = 0; h->md.bartlett = h->d_bart; h->md.Xmn = h->d_Xmn;
  h->have_rng = true;
  h->injected_ready = true;
  return PGB_OK;
}

static int check_ready(pgb_handle* h) {
  if (!h->have_data || !h->have_meas || !h->have_init || !h->have_rng)
    return set_err(h, PGB_ERR_STATE, "set_data / set_measurement / set_init_dist / set_rng_* must be called first");
  for (char c : h->have_state)
    if (!c) return set_err(h, PGB_ERR_STATE, "pgb_set_state missing for a chain");
  return PGB_OK;
}

// Launch sequence of one exact scan + consumer on stream s.
template <int NX, int MODE>
static void launch_scan(pgb_handle* h, const ScanWS& ws, const double* v, const pgb_utable* utab, int64_t utab_stride,
                        int64_t utab_index, int32_t* out_idx, int64_t out_stride, int64_t t, int first,
                        unsigned long long* maxkey_reset) {
  const FilterDev& fd = h->fd;
  cudaStream_t s = h->stream;
  dim3 gt(ws.nt, fd.nc), g1(1, fd.nc);
  LAUNCH(h, PGB_PROF_SCAN_LOCAL, (k_scan_local<<<gt, TPB, 0, s>>>(ws, v, fd.status)));
  LAUNCH(h, PGB_PROF_RESOLVE, (k_resolve<<<g1, TPB, 0, s>>>(ws, g_force_fallback & 1)));
  LAUNCH(h, PGB_PROF_EXPAND, (k_expand<NX, MODE><<<gt, TPB, 0, s>>>(fd, ws, v, utab, utab_stride, utab_index, out_idx,
                                                                    out_stride, t, first, 0)));
  LAUNCH(h, PGB_PROF_SEQ, (k_seq_exact<<<g1, 32, 0, s>>>(ws, v, fd.status, maxkey_reset)));
  LAUNCH(h, PGB_PROF_EXPAND_FB, (k_expand<NX, MODE><<<gt, TPB, 0, s>>>(fd, ws, v, utab, utab_stride, utab_index, out_idx,
                                                                       out_stride, t, first, 1)));
}

template <int NX>
static int run_filter_nx(pgb_handle* h) {
  FilterDev& fd = h->fd;
  MniwDev& md = h->md;
  cudaStream_t s = h->stream;
  const int64_t N = fd.N, T = fd.T;
  const int nc = fd.nc;
  const int first = (h->sweep_index == 1);
  fd.sweep = (uint32_t)h->sweep_index;
  md.sweep = fd.sweep;
  const size_t nA = (size_t)nc * NX * fd.n_phi * 8, nQ = (size_t)nc * NX * NX * 8;
  CK(h, cudaMemcpyAsync(h->d_A_used, fd.A, nA, cudaMemcpyDeviceToDevice, s));
  CK(h, cudaMemcpyAsync(h->d_Q_used, h->d_Q, nQ, cudaMemcpyDeviceToDevice, s));
  CK(h, cudaMemsetAsync(fd.maxkey, 0, (size_t)nc * 8, s));
  dim3 gt(fd.nt, nc);
  const size_t smA = (size_t)NX * fd.n_phi * 8;
  LAUNCH(h, PGB_PROF_INIT, (k_build_utabs<<<dim3((unsigned)((T + 127) / 128), nc), 128, 0, s>>>(fd, first)));
  LAUNCH(h, PGB_PROF_INIT, (k_init<NX><<<dim3((unsigned)((N + TPB - 1) / TPB), nc), TPB, 0, s>>>(fd, first)));
  LAUNCH(h, PGB_PROF_WEIGHTS, (k_weights<<<gt, TPB, 0, s>>>(fd)));
  for (int64_t t = 1; t < T; ++t) {
    LAUNCH(h, PGB_PROF_PREP, (k_prep<NX><<<gt, TPB, smA, s>>>(fd, t, first)));
    launch_scan<NX, MODE_PROPAGATE>(h, fd.main, fd.wbuf, fd.utab_res, T, t, nullptr, 0, t, first, fd.maxkey);
    if (!first) {
      LAUNCH(h, PGB_PROF_NORM, (k_norm<<<gt, TPB, 0, s>>>(fd)));
      launch_scan<NX, MODE_INDEX>(h, fd.side, fd.waN, fd.utab_anc, T, t, fd.ah + t * N + (N - 1), T * N, t, first, nullptr);
    }
    LAUNCH(h, PGB_PROF_WEIGHTS, (k_weights<<<gt, TPB, 0, s>>>(fd)));
  }
  // w_T, then star = systematic_resampling(w[end,:], 1)   (:213)
  LAUNCH(h, PGB_PROF_TAIL, (k_prep<NX><<<gt, TPB, smA, s>>>(fd, T, first)));
  launch_scan<NX, MODE_INDEX>(h, fd.main, fd.wbuf, fd.utab_star, 1, 0, h->d_star_idx, 1, 0, first, nullptr);
  if (fd.x_rows == T) {
    LAUNCH(h, PGB_PROF_TAIL, (k_trace<NX><<<(nc + 31) / 32, 32, 0, s>>>(fd, h->d_star_idx, h->d_star_out)));
  } else {
    LAUNCH(h, PGB_PROF_TAIL, (k_trace_replay<NX><<<nc, TR_TPB, (size_t)fd.n_phi * 8, s>>>(fd, h->d_star_idx, h->d_star_out,
                                                                                        h->d_replay_noise, first)));
  }
  // statistics (:221-225)
  LAUNCH(h, PGB_PROF_TAIL, (k_stats_z<NX><<<dim3((unsigned)((T - 1 + 127) / 128), nc), 128, 0, s>>>(fd, md)));
  const int64_t nent = (int64_t)NX * NX + (int64_t)NX * fd.n_phi + (int64_t)fd.n_phi * fd.n_phi;
  LAUNCH(h, PGB_PROF_TAIL, (k_stats<<<dim3((unsigned)nent, nc), TPB, 0, s>>>(md)));
  CK(h, cudaGetLastError());
  return PGB_OK;
}

// One cooperative launch for the whole time loop.  Returns -1 when the configuration does not fit
// (too many chains / tiles for the co-resident grid); the caller then uses the multi-kernel path.
template <int NX>
static int run_filter_persistent_nx(pgb_handle* h, bool warp_tiles) {
  FilterDev& fd = h->fd;
  MniwDev& md = h->md;
  PersistDev& pd = h->pd;
  cudaStream_t s = h->stream;
  const int64_t T = fd.T;
  const int nc = fd.nc;
  const size_t smA = (warp_tiles ? sizeof(WShared) : sizeof(PersistShared)) + (size_t)NX * fd.n_phi * 8;
  const void* kfn = warp_tiles ? (const void*)k_sweep_w<NX> : (const void*)k_sweep_persistent<NX>;
  int occ = 0;
  CK(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA));
  if (warp_tiles) CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_w<NX>, TPB, smA));
  else CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_persistent<NX>, TPB, smA));
  const int max_ctas = occ * h->num_sms;
  if (max_ctas > 4096) return -1;
  int nbmax = max_ctas / nc;
  if (nbmax < 1) return -1;
  int tpc = 1;
  if (warp_tiles) {  // one block per CTA
    if (nbmax > W_MAXNB) nbmax = W_MAXNB;
    while ((fd.nt + tpc - 1) / tpc > nbmax) tpc *= 2;
    if (tpc > W_MAXTPC) return -1;
    pd.tpc = tpc;
    pd.nb = (fd.nt + tpc - 1) / tpc;
    pd.npc = pd.nb;
  } else {  // finest blocks the per-chain tables allow, dealt over every co-resident CTA slot
    while ((fd.nt + tpc - 1) / tpc > P_MAXNB) tpc *= 2;
    pd.tpc = tpc;
    pd.nb = (fd.nt + tpc - 1) / tpc;
    pd.npc = pd.nb < nbmax ? pd.nb : nbmax;
    if (((pd.nb + pd.npc - 1) / pd.npc) * tpc > P_MAXTPC) return -1;
  }
  pd.force_fallback = g_force_fallback;
  int first = (h->sweep_index == 1);
  fd.sweep = (uint32_t)h->sweep_index;
  md.sweep = fd.sweep;
  const size_t nA = (size_t)nc * NX * fd.n_phi * 8, nQ = (size_t)nc * NX * NX * 8;
  CK(h, cudaMemcpyAsync(h->d_A_used, fd.A, nA, cudaMemcpyDeviceToDevice, s));
  CK(h, cudaMemcpyAsync(h->d_Q_used, h->d_Q, nQ, cudaMemcpyDeviceToDevice, s));
  CK(h, cudaMemsetAsync(fd.maxkey, 0, (size_t)nc * 8, s));
  CK(h, cudaMemsetAsync(pd.bar, 0, 4, s));
  CK(h, cudaMemsetAsync(pd.sm_slots, 0, (size_t)1024 * 4, s));
  CK(h, cudaMemsetAsync(pd.heavy_cnt, 0, (size_t)nc * 4, s));
  CK(h, cudaMemsetAsync(pd.cand, 0, (size_t)2 * nc * 4, s));
  CK(h, cudaMemsetAsync(fd.main.flag, 0, (size_t)2 * nc * 4, s));
  CK(h, cudaMemsetAsync(fd.side.flag, 0, (size_t)2 * nc * 4, s));
  LAUNCH(h, PGB_PROF_INIT, (k_build_utabs<<<dim3((unsigned)((T + 127) / 128), nc), 128, 0, s>>>(fd, first)));
  {
    LaunchScope ls(h, PGB_PROF_EXPAND);
    void* args[] = {(void*)&fd, (void*)&pd, (void*)&first};
    CK(h, cudaLaunchCooperativeKernel(kfn, dim3((unsigned)(nc * pd.npc)), dim3(TPB), args, smA, s));
  }
  if (fd.x_rows == T) {
    LAUNCH(h, PGB_PROF_TAIL, (k_trace<NX><<<(nc + 31) / 32, 32, 0, s>>>(fd, h->d_star_idx, h->d_star_out)));
  } else {
    LAUNCH(h, PGB_PROF_TAIL, (k_trace_replay<NX><<<nc, TR_TPB, (size_t)fd.n_phi * 8, s>>>(fd, h->d_star_idx, h->d_star_out,
                                                                                        h->d_replay_noise, first)));
  }
  LAUNCH(h, PGB_PROF_TAIL, (k_stats_z<NX><<<dim3((unsigned)((T - 1 + 127) / 128), nc), 128, 0, s>>>(fd, md)));
  const int64_t nent = (int64_t)NX * NX + (int64_t)NX * fd.n_phi + (int64_t)fd.n_phi * fd.n_phi;
  LAUNCH(h, PGB_PROF_TAIL, (k_stats<<<dim3((unsigned)nent, nc), TPB, 0, s>>>(md)));
  CK(h, cudaGetLastError());
  return PGB_OK;
}

static int64_t g_small_auto_max = CTA_MAXN;  // largest N the automatic mode gives to the warp / one-CTA-per-chain kernels
                                               // (measured faster than the tile kernel up to 1024, profiles/r01_mid_n_kernel.md)

// Small particle counts (N <= 32, stored history): one warp per chain, the whole time loop in one ordinary launch
// (pgb_small.cuh).  Returns -1 when the configuration does not qualify.
template <int NX>
static int run_filter_small_nx(pgb_handle* h, int64_t n_max) {
  FilterDev& fd = h->fd;
  MniwDev& md = h->md;
  cudaStream_t s = h->stream;
  const int64_t T = fd.T;
  const int nc = fd.nc;
  if (fd.N > n_max || fd.N > CTA_MAXN || fd.x_rows != T) return -1;
  const bool cta = fd.N > SMALL_MAXN;  // mid-size N: one CTA per chain (k_sweep_cta), thread = particle
  const int np_thr = (int)((fd.N + 31) / 32) * 32;
  const size_t smem = cta ? ((size_t)((NX * fd.n_phi + 1) & ~1) + (size_t)(NX + 4) * np_thr + 4) * 8 + sizeof(CtaShared)
                          : ((size_t)((NX * fd.n_phi + 1) & ~1) + (size_t)NX * 32 + 32) * 8;
  if (smem > 200 * 1024) return -1;
  int first = (h->sweep_index == 1);
  fd.sweep = (uint32_t)h->sweep_index;
  md.sweep = fd.sweep;
  const size_t nA = (size_t)nc * NX * fd.n_phi * 8, nQ = (size_t)nc * NX * NX * 8;
  CK(h, cudaMemcpyAsync(h->d_A_used, fd.A, nA, cudaMemcpyDeviceToDevice, s));
  CK(h, cudaMemcpyAsync(h->d_Q_used, h->d_Q, nQ, cudaMemcpyDeviceToDevice, s));
  if (cta) {
    CK(h, cudaFuncSetAttribute(k_sweep_cta<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(h, PGB_PROF_EXPAND, (k_sweep_cta<NX><<<nc, np_thr, smem, s>>>(fd, h->pd.star_idx, first)));
    LAUNCH(h, PGB_PROF_TAIL, (k_trace<NX><<<(nc + 31) / 32, 32, 0, s>>>(fd, h->d_star_idx, h->d_star_out)));
  } else {
    CK(h, cudaFuncSetAttribute(k_sweep_small<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(h, PGB_PROF_EXPAND, (k_sweep_small<NX><<<nc, 32, smem, s>>>(fd, h->pd.star_idx, first)));
    LAUNCH(h, PGB_PROF_TAIL, (k_trace_small<NX><<<nc, 32, 0, s>>>(fd, h->d_star_idx, h->d_star_out)));
  }
  LAUNCH(h, PGB_PROF_TAIL, (k_stats_z<NX><<<dim3((unsigned)((T - 1 + 127) / 128), nc), 128, 0, s>>>(fd, md)));
  const int64_t nent = (int64_t)NX * NX + (int64_t)NX * fd.n_phi + (int64_t)fd.n_phi * fd.n_phi;
  LAUNCH(h, PGB_PROF_TAIL, (k_stats<<<dim3((unsigned)nent, nc), TPB, 0, s>>>(md)));
  CK(h, cudaGetLastError());
  return PGB_OK;
}

static int run_filter(pgb_handle* h) {
  int mode0 = h->persistent;
  if (mode0 < 0 || mode0 == 3) {
    int rc;
    // the warp kernel up to N = 32, the one-CTA-per-chain kernel up to N = 1024 (automatic and mode 3)
    const int64_t n_max = (mode0 == 3) ? CTA_MAXN : g_small_auto_max;
    switch (h->cfg.n_x) {
      case 1: rc = run_filter_small_nx<1>(h, n_max); break;
      case 2: rc = run_filter_small_nx<2>(h, n_max); break;
      case 3: rc = run_filter_small_nx<3>(h, n_max); break;
      default: rc = run_filter_small_nx<4>(h, n_max); break;
    }
    if (rc >= 0) return rc;
    mode0 = 1;
  }
  for (int mode = mode0; mode >= 1; --mode) {  // a configuration that does not fit falls through
    int rc;
    const bool wt = (mode == 2);
    switch (h->cfg.n_x) {
      case 1: rc = run_filter_persistent_nx<1>(h, wt); break;
      case 2: rc = run_filter_persistent_nx<2>(h, wt); break;
      case 3: rc = run_filter_persistent_nx<3>(h, wt); break;
      default: rc = run_filter_persistent_nx<4>(h, wt); break;
    }
    if (rc >= 0) return rc;
  }
  switch (h->cfg.n_x) {
    case 1: return run_filter_nx<1>(h);
    case 2: return run_filter_nx<2>(h);
    case 3: return run_filter_nx<3>(h);
    default: return run_filter_nx<4>(h);
  }
}

static int run_mniw(pgb_handle* h) {
  if (!h->have_prior) return set_err(h, PGB_ERR_STATE, "pgb_set_prior must be called before the MNIW draw");
  LAUNCH(h, PGB_PROF_MNIW, (k_mniw<<<h->fd.nc, MNIW_TPB, 0, h->stream>>>(h->md)));
  CK(h, cudaGetLastError());
  h->sweep_index += 1;
  return PGB_OK;
}

static int check_posdef(pgb_handle* h) {
  std::vector<int> st(h->fd.nc);
  CK(h, cudaMemcpyAsync(st.data(), h->fd.status, (size_t)h->fd.nc * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  for (int v : st)
    if (v & PGB_STATUS_NOT_POSDEF)
      return set_err(h, PGB_ERR_NOT_POSDEF, "Cholesky failed on device (matrix not positive definite)");
  return PGB_OK;
}

extern "C" int pgb_sweep_filter(pgb_handle* h) {
  NEED(h);
  int rc = check_ready(h);
  if (rc) return rc;
  return run_filter(h);
}
extern "C" int pgb_sweep_mniw(pgb_handle* h) {
  NEED(h);
  return run_mniw(h);
}
extern "C" int pgb_sweep(pgb_handle* h, int64_t n_sweeps) {
  NEED(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!h->fd.philox && n_sweeps != 1) return set_err(h, PGB_ERR_INVALID, "injected streams cover exactly one sweep");
  for (int64_t i = 0; i < n_sweeps; ++i) {
    if ((rc = run_filter(h))) return rc;
    if ((rc = run_mniw(h))) return rc;
  }
  return check_posdef(h);
}
extern "C" int pgb_synchronize(pgb_handle* h) {
  NEED(h);
  CK(h, cudaStreamSynchronize(h->stream));
  return PGB_OK;
}

static int fetch_sample(pgb_handle* h, int chain, double* A, double* Q, double* w_m1, double* x_m1, double* u_m1,
                        std::vector<double>& tmp) {
  const int nx = h->cfg.n_x, np = h->cfg.n_phi, nu = h->cfg.n_u;
  const int64_t N = h->cfg.N, T = h->cfg.T;
  cudaStream_t s = h->stream;
  if (A) CK(h, cudaMemcpyAsync(A, h->d_A_used + (size_t)chain * nx * np, (size_t)nx * np * 8, cudaMemcpyDeviceToHost, s));
  if (Q) CK(h, cudaMemcpyAsync(Q, h->d_Q_used + (size_t)chain * nx * nx, (size_t)nx * nx * 8, cudaMemcpyDeviceToHost, s));
  if (w_m1) CK(h, cudaMemcpyAsync(w_m1, h->fd.wbuf + (size_t)chain * N, (size_t)N * 8, cudaMemcpyDeviceToHost, s));
  if (x_m1) {
    tmp.resize((size_t)nx * N);
    const size_t row = (h->fd.x_rows == T) ? (size_t)(T - 1) : (size_t)((T - 1) & 1);
    CK(h, cudaMemcpyAsync(tmp.data(), h->fd.xh + ((size_t)chain * h->fd.x_rows + row) * nx * N, (size_t)nx * N * 8,
                          cudaMemcpyDeviceToHost, s));
  }
  CK(h, cudaStreamSynchronize(s));
  if (x_m1)
    for (int64_t n = 0; n < N; ++n)
      for (int d = 0; d < nx; ++d) x_m1[d + nx * n] = tmp[(size_t)d * N + n];
  if (u_m1)
    for (int c = 0; c < nu; ++c) u_m1[c] = h->h_u[(size_t)nu * (T - 1) + c];
  return PGB_OK;
}

extern "C" int pgb_get_sample(pgb_handle* h, int32_t chain, double* A, double* Q, double* w_m1, double* x_m1,
                              double* u_m1) {
  NEED(h);
  if (chain < 0 || chain >= h->cfg.n_chains) return set_err(h, PGB_ERR_INVALID, "bad chain %d", chain);
  std::vector<double> tmp;
  return fetch_sample(h, chain, A, Q, w_m1, x_m1, u_m1, tmp);
}

extern "C" int pgb_particle_gibbs(pgb_handle* h, int64_t K, int64_t K_b, int64_t k_d, double* A_s, double* Q_s,
                                  double* w_s, double* x_s, double* u_m1) {
  NEED(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!h->fd.philox) return set_err(h, PGB_ERR_INVALID, "pgb_particle_gibbs needs Philox streams");
  if (!h->have_prior) return set_err(h, PGB_ERR_STATE, "pgb_set_prior missing");
  if (K < 1 || K_b < 0 || k_d < 0) return set_err(h, PGB_ERR_INVALID, "bad K/K_b/k_d");
  const int nc = h->cfg.n_chains, nx = h->cfg.n_x, np = h->cfg.n_phi;
  const int64_t N = h->cfg.N;
  const int64_t K_total = K_b + 1 + (K - 1) * (k_d + 1);  // particle_Gibbs.jl:116
  std::vector<double> tmp;
  int64_t cur = 0;
  for (int64_t k = h->sweep_index; k <= K_total; ++k) {
    if ((rc = run_filter(h))) return rc;
    if (K_b < k && ((k - (K_b + 1)) % (k_d + 1) == 0) && cur < K) {  // :203
      for (int c = 0; c < nc; ++c) {
        size_t o = (size_t)c * K + cur;
        rc = fetch_sample(h, c, A_s ? A_s + o * nx * np : nullptr, Q_s ? Q_s + o * nx * nx : nullptr,
                          w_s ? w_s + o * N : nullptr, x_s ? x_s + o * nx * N : nullptr, u_m1, tmp);
        if (rc) return rc;
      }
      ++cur;
    }
    if ((rc = run_mniw(h))) return rc;
    if ((k & 15) == 0 && (rc = check_posdef(h))) return rc;
  }
  return check_posdef(h);
}

extern "C" int pgb_get_history(pgb_handle* h, int32_t chain, int64_t* a, double* x, double* w) {
  NEED(h);
  const int nx = h->cfg.n_x;
  const int64_t N = h->cfg.N, T = h->cfg.T;
  if (chain < 0 || chain >= h->cfg.n_chains) return set_err(h, PGB_ERR_INVALID, "bad chain %d", chain);
  CK(h, cudaStreamSynchronize(h->stream));
  if (a) {
    std::vector<int32_t> t32((size_t)T * N);
    CK(h, cudaMemcpy(t32.data(), h->fd.ah + (size_t)chain * T * N, (size_t)T * N * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < t32.size(); ++i) a[i] = (int64_t)t32[i] + 1;
    for (int64_t n = 0; n < N; ++n) a[n] = 0;  // a[1,:] is never written by the reference
  }
  if (x) {
    if (h->fd.x_rows != T)
      return set_err(h, PGB_ERR_STATE, "the particle history x was not kept (x_history = replay); create the handle "
                                       "with x_history = PGB_X_HISTORY_STORE");
    std::vector<double> tx((size_t)T * nx * N);
    CK(h, cudaMemcpy(tx.data(), h->fd.xh + (size_t)chain * T * nx * N, tx.size() * 8, cudaMemcpyDeviceToHost));
    for (int64_t t = 0; t < T; ++t)
      for (int64_t n = 0; n < N; ++n)
        for (int d = 0; d < nx; ++d) x[(t * N + n) * nx + d] = tx[((size_t)t * nx + d) * N + n];
  }
  if (w) {
    if (!h->fd.whist) return set_err(h, PGB_ERR_STATE, "w history was not kept (keep_w_history = 0)");
    CK(h, cudaMemcpy(w, h->fd.whist + (size_t)chain * T * N, (size_t)T * N * 8, cudaMemcpyDeviceToHost));
  }
  return PGB_OK;
}

extern "C" int pgb_get_stats(pgb_handle* h, int32_t chain, double* Phi, double* Psi, double* Sigma, int64_t* star) {
  NEED(h);
  const int nx = h->cfg.n_x, np = h->cfg.n_phi;
  if (chain < 0 || chain >= h->cfg.n_chains) return set_err(h, PGB_ERR_INVALID, "bad chain %d", chain);
  CK(h, cudaStreamSynchronize(h->stream));
  if (Phi) CK(h, cudaMemcpy(Phi, h->md.Phi + (size_t)chain * nx * nx, (size_t)nx * nx * 8, cudaMemcpyDeviceToHost));
  if (Psi) CK(h, cudaMemcpy(Psi, h->md.Psi + (size_t)chain * nx * np, (size_t)nx * np * 8, cudaMemcpyDeviceToHost));
  if (Sigma) CK(h, cudaMemcpy(Sigma, h->md.Sigma + (size_t)chain * np * np, (size_t)np * np * 8, cudaMemcpyDeviceToHost));
  if (star) CK(h, cudaMemcpy(star, h->d_star_out + chain, 8, cudaMemcpyDeviceToHost));
  return PGB_OK;
}

extern "C" int pgb_get_status(pgb_handle* h, int32_t chain, int32_t* status_bits, int64_t* n_fallback, int64_t* n_scans,
                              int32_t clear) {
  NEED(h);
  if (chain < 0 || chain >= h->cfg.n_chains) return set_err(h, PGB_ERR_INVALID, "bad chain %d", chain);
  CK(h, cudaStreamSynchronize(h->stream));
  int st = 0;
  long long cm[2], cs[2];
  CK(h, cudaMemcpy(&st, h->fd.status + chain, 4, cudaMemcpyDeviceToHost));
  CK(h, cudaMemcpy(cm, h->d_counters_main + 2 * chain, 16, cudaMemcpyDeviceToHost));
  CK(h, cudaMemcpy(cs, h->d_counters_side + 2 * chain, 16, cudaMemcpyDeviceToHost));
  if (status_bits) *status_bits = st;
  if (n_fallback) *n_fallback = cm[0] + cs[0];
  if (n_scans) *n_scans = cm[1] + cs[1];
  if (clear) {
    CK(h, cudaMemset(h->fd.status + chain, 0, 4));
    CK(h, cudaMemset(h->d_counters_main + 2 * chain, 0, 16));
    CK(h, cudaMemset(h->d_counters_side + 2 * chain, 0, 16));
  }
  return PGB_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone helpers
// ------------------------------------------------------------------------------------------------
__global__ void k_build_one_utab(pgb_utable* ut, int64_t M, double rnd) { pgb_utable_build(ut, M, rnd); }

extern "C" int pgb_systematic_resampling(int32_t device, const double* W, int64_t n_w, int64_t n_draw, double rnd,
                                         int64_t* idx, int32_t* status) {
  if (!W || !idx || n_w < 1 || n_draw < 1) return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  if (n_w >= (1ll << 31) - TILE) return set_err(nullptr, PGB_ERR_INVALID, "n_w too large");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return set_err(nullptr, PGB_ERR_CUDA, "no CUDA device available (%s); pgopt_b200 has no CPU fallback",
                   cudaGetErrorString(e));
  pgb_handle tmp;  // only used as an allocation list / error sink
  pgb_handle* h = &tmp;
  h->cfg.device = device;
  struct Cleanup {
    pgb_handle* h;
    ~Cleanup() { for (void* p : h->allocs) cudaFree(p); }
  } cleanup{h};
  CK(h, cudaSetDevice(device));
  FilterDev fd;
  memset(&fd, 0, sizeof(fd));
  fd.nc = 1; fd.N = n_w; fd.T = 1; fd.n_x = 1; fd.n_y = 1; fd.n_u = 1;
  ScanWS ws;
  CK(h, alloc_scanws(h, &ws, n_w, 1, nullptr));
  fd.nt = ws.nt;
  double* d_W = nullptr;
  int32_t* d_idx = nullptr;
  pgb_utable* d_ut = nullptr;
  CK(h, dalloc(h, &d_W, (size_t)n_w));
  CK(h, dalloc(h, &d_idx, (size_t)n_draw));
  CK(h, dalloc(h, &d_ut, 1));
  CK(h, dalloc(h, &fd.status, 1));
  CK(h, cudaMemcpy(d_W, W, (size_t)n_w * 8, cudaMemcpyHostToDevice));
  cudaStream_t s = 0;
  h->stream = s;
  h->fd = fd;
  k_build_one_utab<<<1, 1, 0, s>>>(d_ut, n_draw, rnd);
  k_tile_sums<<<dim3(ws.nt, 1), TPB, 0, s>>>(ws, d_W);
  launch_scan<1, MODE_INDEX>(h, ws, d_W, d_ut, 1, 0, d_idx, n_draw, 0, 0, nullptr);
  CK(h, cudaGetLastError());
  CK(h, cudaDeviceSynchronize());
  std::vector<int32_t> t32((size_t)n_draw);
  int st = 0;
  CK(h, cudaMemcpy(t32.data(), d_idx, (size_t)n_draw * 4, cudaMemcpyDeviceToHost));
  CK(h, cudaMemcpy(&st, fd.status, 4, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n_draw; ++i) idx[i] = (int64_t)t32[i] + 1;
  if (status) {
    long long c[2];
    CK(h, cudaMemcpy(c, ws.counters, 16, cudaMemcpyDeviceToHost));
    *status = st | (c[0] ? 0x100 : 0);  // bit 8: the serial exact fallback ran
  }
  h->stream = nullptr;
  return PGB_OK;
}

extern "C" int pgb_mniw_sample(int32_t device, int32_t n_x, int32_t n_phi, const double* Phi, const double* Psi,
                               const double* Sigma, const double* V_diag, const double* Lambda_Q, double ell_Q, int64_t T,
                               const double* bartlett, const double* Xmn, uint64_t seed, uint32_t chain,
                               uint32_t sweep_index, double* A_out, double* Q_out) {
  if (!Phi || !Psi || !Sigma || !V_diag || !Lambda_Q || !A_out || !Q_out || n_x < 1 || n_x > 4 || n_phi < 1)
    return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return set_err(nullptr, PGB_ERR_CUDA, "no CUDA device available (%s); pgopt_b200 has no CPU fallback",
                   cudaGetErrorString(e));
  pgb_handle tmp;
  pgb_handle* h = &tmp;
  struct Cleanup {
    pgb_handle* h;
    ~Cleanup() { for (void* p : h->allocs) cudaFree(p); }
  } cleanup{h};
  CK(h, cudaSetDevice(device));
  MniwDev md;
  memset(&md, 0, sizeof(md));
  md.nc = 1; md.n_x = n_x; md.n_phi = n_phi; md.Tm1 = T; md.ell_Q = ell_Q;
  const size_t pp = (size_t)n_phi * n_phi, xp = (size_t)n_x * n_phi, xx = (size_t)n_x * n_x;
  double *dV, *dL, *dB = nullptr, *dX = nullptr, *dQ;
  CK(h, dalloc(h, &dV, (size_t)n_phi));
  CK(h, dalloc(h, &dL, xx));
  CK(h, dalloc(h, &md.Phi, xx)); CK(h, dalloc(h, &md.Psi, xp)); CK(h, dalloc(h, &md.Sigma, pp));
  CK(h, dalloc(h, &md.Ls, pp)); CK(h, dalloc(h, &md.Sinv, pp)); CK(h, dalloc(h, &md.Lv, pp));
  CK(h, dalloc(h, &md.Mpost, xp)); CK(h, dalloc(h, &md.Xn, xp)); CK(h, dalloc(h, &md.LX, xp));
  CK(h, dalloc(h, &md.A, xp)); CK(h, dalloc(h, &dQ, xx)); CK(h, dalloc(h, &md.Lq, 16)); CK(h, dalloc(h, &md.c0, 1));
  CK(h, dalloc(h, &md.status, 1));
  md.Q = dQ; md.V_diag = dV; md.Lambda_Q = dL;
  CK(h, cudaMemcpy(dV, V_diag, (size_t)n_phi * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(dL, Lambda_Q, xx * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(md.Phi, Phi, xx * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(md.Psi, Psi, xp * 8, cudaMemcpyHostToDevice));
  CK(h, cudaMemcpy(md.Sigma, Sigma, pp * 8, cudaMemcpyHostToDevice));
  if (bartlett && Xmn) {
    CK(h, dalloc(h, &dB, xx)); CK(h, dalloc(h, &dX, xp));
    CK(h, cudaMemcpy(dB, bartlett, xx * 8, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(dX, Xmn, xp * 8, cudaMemcpyHostToDevice));
    md.philox = 0; md.bartlett = dB; md.Xmn = dX;
  } else {
    md.philox = 1; md.seed = seed; md.chain_base = chain; md.sweep = sweep_index;
  }
  k_mniw<<<1, MNIW_TPB>>>(md);
  CK(h, cudaGetLastError());
  CK(h, cudaDeviceSynchronize());
  int st = 0;
  CK(h, cudaMemcpy(&st, md.status, 4, cudaMemcpyDeviceToHost));
  if (st & PGB_STATUS_NOT_POSDEF) return set_err(nullptr, PGB_ERR_NOT_POSDEF, "matrix not positive definite");
  CK(h, cudaMemcpy(A_out, md.A, xp * 8, cudaMemcpyDeviceToHost));
  CK(h, cudaMemcpy(Q_out, dQ, xx * 8, cudaMemcpyDeviceToHost));
  return PGB_OK;
}

static int g_rollout_grid[3] = {0, 0, 0};  // last warp-kernel launch: {CTAs, CTAs per SM, runs of A in shared memory}
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
  switch (rs) {
    case 0: return launch_rollout_warp_ni<NX, RW_MAXNI, 0>(fd, rd, n_sm);
    case 3: return launch_rollout_warp_ni<NX, RW_MAXNI, 3>(fd, rd, n_sm);
    default: return launch_rollout_warp_ni<NX, RW_MAXNI, 2>(fd, rd, n_sm);
  }
}

// Shared body of pgb_rollout (n_models = 0, y_test = null) and pgb_test_prediction (K scenarios over n_models models,
// scenario k simulating model k mod n_models; sum_sq_err = tree sum of (Y - y_test)^2).
static int rollout_run(const pgb_config* cfg, int64_t K, int64_t n_models, int64_t H, const double* A, const double* Q,
                       const double* w_m1, const double* x_m1, const double* u_m1, const double* C, const double* Le,
                       const double* u_init, uint64_t seed, uint32_t k_offset, double* x0, double* v, double* e,
                       double* X, double* Y, const double* y_test, double* sum_sq_err) {
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  const int64_t KM = n_models > 0 ? n_models : K;  // number of models (PG samples) behind the K scenarios
  if (!A || !Q || !w_m1 || !x_m1 || !u_m1 || !C || !Le || !u_init || K < 1 || H < 1 || cfg->N < 1)
    return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0)
    return set_err(nullptr, PGB_ERR_CUDA, "no CUDA device available (%s); pgopt_b200 has no CPU fallback",
                   cudaGetErrorString(ce));
  pgb_handle tmp;
  pgb_handle* h = &tmp;
  struct Cleanup {
    pgb_handle* h;
    ~Cleanup() { for (void* p : h->allocs) cudaFree(p); }
  } cleanup{h};
  CK(h, cudaSetDevice(cfg->device));
  const int nx = cfg->n_x, np = cfg->n_phi, nu = cfg->n_u, ny = cfg->n_y;
  const int64_t N = cfg->N;
  FilterDev fd;
  memset(&fd, 0, sizeof(fd));
  fill_model(fd, cfg);
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
  // the per-warp noise buffers fit shared memory; PGB_ROLLOUT_KERNEL=cta|warp forces one (tests run both)
  int rs = 2;  // runs (of 4 per lane) of A_k kept in shared memory; PGB_ROLLOUT_RS=0|2|3 (experiments)
  if (const char* e2 = getenv("PGB_ROLLOUT_RS")) rs = atoi(e2);
  if (rs != 0 && rs != 3) rs = 2;
  const size_t smem_w = rollout_warp_smem(nx, ny, nu, H, np <= 128 ? 1 : RW_MAXNI, np <= 128 ? 0 : rs);
  bool use_warp = np <= 128 * RW_MAXNI && smem_w <= 100 * 1024;
  if (const char* kk = getenv("PGB_ROLLOUT_KERNEL")) {
    if (!strcmp(kk, "cta")) use_warp = false;
    else if (!strcmp(kk, "warp") && !use_warp)
      return set_err(nullptr, PGB_ERR_INVALID, "rollout: warp kernel needs n_phi <= %d and a shorter horizon", 128 * RW_MAXNI);
  }
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
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return set_err(nullptr, PGB_ERR_CUDA, "no CUDA device available (%s); pgopt_b200 has no CPU fallback",
                   cudaGetErrorString(e));
  pgb_handle tmp;  // only used as an allocation list / error sink
  pgb_handle* h = &tmp;
  h->cfg.device = device;
  struct Cleanup {
    pgb_handle* h;
    ~Cleanup() { for (void* p : h->allocs) cudaFree(p); }
  } cleanup{h};
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

extern "C" int pgb_get_phase_cycles(pgb_handle* h, double* avg_cycles /*[16]*/, int32_t* n_ctas) {
  NEED(h);
  CK(h, cudaStreamSynchronize(h->stream));
  const int g = h->fd.nc * h->pd.npc;
  std::vector<long long> v((size_t)1024 * P_NPROF);
  CK(h, cudaMemcpy(v.data(), h->pd.prof, v.size() * 8, cudaMemcpyDeviceToHost));
  for (int i = 0; i < P_NPROF; ++i) {
    double s = 0;
    for (int c = 0; c < g && c < 1024; ++c) s += (double)v[(size_t)c * P_NPROF + i];
    avg_cycles[i] = g > 0 ? s / g : 0.0;
  }
  if (n_ctas) *n_ctas = g;
  return PGB_OK;
}
extern "C" int pgb_set_mode(pgb_handle* h, int32_t persistent) {
  NEED(h);
  h->persistent = persistent < 0 ? -1 : (persistent > 3 ? 3 : persistent);
  return PGB_OK;
}
extern "C" int pgb_timer_start(pgb_handle* h) {
  NEED(h);
  if (!h->ev0) { CK(h, cudaEventCreate(&h->ev0)); CK(h, cudaEventCreate(&h->ev1)); }
  CK(h, cudaEventRecord(h->ev0, h->stream));
  return PGB_OK;
}
extern "C" int pgb_timer_stop(pgb_handle* h, double* elapsed_ms) {
  NEED(h);
  if (!h->ev0 || !elapsed_ms) return set_err(h, PGB_ERR_STATE, "pgb_timer_start was not called");
  CK(h, cudaEventRecord(h->ev1, h->stream));
  CK(h, cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *elapsed_ms = ms;
  return PGB_OK;
}
extern "C" int pgb_profile(pgb_handle* h, int32_t on) {
  NEED(h);
  profile_collect(h);
  h->profile = on != 0;
  if (on) {
    for (int i = 0; i < PGB_PROF_NCLASS; ++i) { h->prof_launches[i] = 0; h->prof_ms[i] = 0.0; }
  }
  return PGB_OK;
}
extern "C" int pgb_get_profile(pgb_handle* h, int64_t* launches, double* ms) {
  NEED(h);
  profile_collect(h);
  for (int i = 0; i < PGB_PROF_NCLASS; ++i) {
    if (launches) launches[i] = h->prof_launches[i];
    if (ms) ms[i] = h->prof_ms[i];
  }
  return PGB_OK;
}
extern "C" int pgb_get_launch_count(pgb_handle* h, int64_t* n) {
  NEED(h);
  if (n) *n = h->n_launch;
  return PGB_OK;
}

// ------------------------------------------------------------------------------------------------
// test hooks
// ------------------------------------------------------------------------------------------------
extern "C" void pgb_test_force_fallback(int32_t on) { g_force_fallback = on; }

__global__ void k_test_math(int fn, const double* __restrict__ x, double* __restrict__ y, int64_t n, uint64_t seed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (fn) {
    case 0: y[i] = pgb_exp(x[i]); break;
    case 1: y[i] = pgb_log(x[i]); break;
    case 2: y[i] = pgb_sin(x[i]); break;
    case 3: y[i] = pgb_cos(x[i]); break;
    case 4: pgb_normal_pair(pgb_stream_block(seed, PGB_STREAM_Z, 0, 1, 5, (uint32_t)i), &y[2 * i], &y[2 * i + 1]); break;
    case 20: {  // x holds pairs (a, d): y[i] = 1 if div_inv(a, d) == a / d bitwise (or both NaN)
      double a = x[2 * i], d = x[2 * i + 1];
      InvDiv iv = make_invdiv(d);
      double q = div_inv(a, iv), t = __ddiv_rn(a, d);
      y[i] = (__double_as_longlong(q) == __double_as_longlong(t) || (q != q && t != t)) ? 1.0 : 0.0;
      break;
    }
    case 21: {  // self-generated operands: x[i] is a seed; 4096 pairs per thread; y[i] = number of mismatches
      uint64_t st = (uint64_t)x[i] * 0x9E3779B97F4A7C15ull + 0x1234567ull + (uint64_t)i;
      int bad = 0;
      for (int k = 0; k < 4096; ++k) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        uint64_t ma = st;
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        uint64_t md = st;
        // mantissas: random, or special patterns (all ones, few bits) ; exponents within the fast range
        if ((k & 7) == 1) md |= 0x000fffffffffff00ull;
        if ((k & 7) == 2) md &= 0xfff00000000000ffull;
        if ((k & 7) == 3) ma |= 0x000ffffffffffff0ull;
        int ea = 1023 + (int)((ma >> 52) % 600) - 300, ed = 1023 + (int)((md >> 52) % 600) - 300;
        double a = __longlong_as_double((long long)((ma & 0x000fffffffffffffull) | ((uint64_t)ea << 52)));
        double d = __longlong_as_double((long long)((md & 0x000fffffffffffffull) | ((uint64_t)ed << 52)));
        if ((k & 7) == 4) {  // numerators next to exact multiples of d (rounding boundary cases)
          double qq = __longlong_as_double((long long)((ma & 0x000fffffffffffffull) | (1023ull << 52)));
          a = __dmul_rn(qq, d);
          a = __longlong_as_double(__double_as_longlong(a) + (long long)(k % 3) - 1);
        }
        InvDiv iv = make_invdiv(d);
        double q = div_inv(a, iv), t = __ddiv_rn(a, d);
        if (__double_as_longlong(q) != __double_as_longlong(t)) ++bad;
      }
      y[i] = (double)bad;
      break;
    }
    default: break;
  }
}

// K-wide branch-free versions (pgb_vmath.cuh): thread i handles elements 4i..4i+3
__global__ void k_test_vmath(int fn, const double* __restrict__ x, double* __restrict__ y, int64_t n, uint64_t seed) {
  int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  double a[4], b[4], c[4];
  for (int k = 0; k < 4; ++k) a[k] = (fn != 14 && i4 + k < n) ? x[i4 + k] : 1.0;
  if (fn == 10) vexp_wide<4>(a, b);
  else if (fn == 11) vlog_pos_normal_wide<4>(a, b);
  else if (fn == 12) vsincos_wide<4>(a, b, c);
  else if (fn == 13) vsincos_wide<4>(a, c, b);
  else {
    pgb_u32x4 ctr[4];
    for (int k = 0; k < 4; ++k) {
      ctr[k].v[0] = (uint32_t)(i4 + k); ctr[k].v[1] = 5; ctr[k].v[2] = 1; ctr[k].v[3] = ((uint32_t)PGB_STREAM_Z << 24);
    }
    vphilox_wide<4>(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    vnormal_pairs_wide<4>(ctr, b, c);
    for (int k = 0; k < 4; ++k)
      if (i4 + k < n) { y[2 * (i4 + k)] = b[k]; y[2 * (i4 + k) + 1] = c[k]; }
    return;
  }
  for (int k = 0; k < 4; ++k)
    if (i4 + k < n) y[i4 + k] = b[k];
}

extern "C" int pgb_test_math(int32_t device, int32_t fn, const double* x, double* y, int64_t n) {
  if (!x || !y || n < 1) return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return set_err(nullptr, PGB_ERR_CUDA, "no CUDA device available (%s)", cudaGetErrorString(e));
  pgb_handle tmp;
  pgb_handle* h = &tmp;
  struct Cleanup {
    pgb_handle* h;
    ~Cleanup() { for (void* p : h->allocs) cudaFree(p); }
  } cleanup{h};
  CK(h, cudaSetDevice(device));
  const bool pairs = (fn == 4 || fn == 14);
  const int64_t nout = pairs ? 2 * n : n;
  const int64_t nin = pairs ? 1 : (fn == 20 ? 2 * n : n);
  double *dx, *dy;
  CK(h, dalloc(h, &dx, (size_t)nin));
  CK(h, dalloc(h, &dy, (size_t)nout));
  CK(h, cudaMemcpy(dx, x, (size_t)nin * 8, cudaMemcpyHostToDevice));
  uint64_t seed = pairs ? (uint64_t)x[0] : 0;
  if (fn >= 10 && fn < 20) k_test_vmath<<<(unsigned)((n + 1023) / 1024), 256>>>(fn, dx, dy, n, seed);
  else k_test_math<<<(unsigned)((n + 255) / 256), 256>>>(fn, dx, dy, n, seed);
  CK(h, cudaGetLastError());
  CK(h, cudaDeviceSynchronize());
  CK(h, cudaMemcpy(y, dy, (size_t)nout * 8, cudaMemcpyDeviceToHost));
  return PGB_OK;
}

__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int pgb_test_fp64_peak(int32_t device, double* tflops) {
  if (!tflops) return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return set_err(nullptr, PGB_ERR_CUDA, "no CUDA device available (%s)", cudaGetErrorString(e));
  pgb_handle tmp;
  pgb_handle* h = &tmp;
  struct Cleanup {
    pgb_handle* h;
    ~Cleanup() { for (void* p : h->allocs) cudaFree(p); }
  } cleanup{h};
  CK(h, cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(h, cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, iters = 20000;
  double* out;
  CK(h, dalloc(h, &out, (size_t)blocks * 256));
  cudaEvent_t e0, e1;
  CK(h, cudaEventCreate(&e0));
  CK(h, cudaEventCreate(&e1));
  k_fp64_peak<<<blocks, 256>>>(out, 1000);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(h, cudaEventRecord(e0));
    k_fp64_peak<<<blocks, 256>>>(out, iters);
    CK(h, cudaEventRecord(e1));
    CK(h, cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(h, cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return PGB_OK;
}

// ------------------------------------------------------------------------------------------------
// micro-benchmark of one elementwise phase in isolation (what a persistent phase can reach):
// variant 0: y = exp(x - m) with per-CTA tree sums (the weight phase), 1: plain copy, 2: y = x / s
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_micro(int variant, const double* __restrict__ x, double* __restrict__ y,
                                                  double* part, unsigned int* bar, int64_t N, int iters, int tpc) {
  __shared__ double wpart[8][2];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x, nwt = 8 * tpc;
  unsigned int target = 0;
  for (int it = 0; it < iters; ++it) {
    TreeAcc acc;
    acc.init();
    const double m = 0.25 + 1e-3 * it;
    for (int j = 0; j < tpc; ++j) {
      const int64_t base = ((int64_t)c * nwt + w * tpc + j) * 128 + lane * 4;
      double a[4], v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = (base + i < N) ? x[base + i] : 0.0;
      if (variant == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = __dsub_rn(a[i], m);
        vexp<4>(a, v);
      } else if (variant == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ddiv_rn(a[i], m);
      } else if (variant == 3) {  // Philox + Box-Muller for 4 outputs
        pgb_u32x4 ctr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ctr[i].v[0] = (uint32_t)(base + i); ctr[i].v[1] = it; ctr[i].v[2] = 1; ctr[i].v[3] = 7;
        }
        vphilox<4>(ctr, 123u, 456u);
        double z0[4], z1[4];
        vnormal_pairs<4>(ctr, z0, z1);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __dadd_rn(z0[i], __dadd_rn(z1[i], a[i]));
      } else if (variant == 4) {  // two sincos per element (known-basis F)
        double s1[4], c1[4], s2[4], c2[4];
        vsincos<4>(a, s1, c1);
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = __dmul_rn(a[i], 2.0);
        vsincos<4>(a, s2, c2);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __dadd_rn(c1[i], s2[i]);
      } else if (variant == 5) {  // classify + maps + warp segmented scan (scan-local without the division)
        double s1 = a[0], s2 = __dadd_rn(s1, a[1]), s3 = __dadd_rn(s2, a[2]), s4 = __dadd_rn(s3, a[3]);
        double inc = s4;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          double t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc = __dadd_rn(t, inc);
        }
        double P = __dadd_rn(0.25 + 1e-3 * j, __dsub_rn(inc, s4));
        double Pn = __shfl_down_sync(0xffffffffu, P, 1);
        double qt[5] = {P, __dadd_rn(P, s1), __dadd_rn(P, s2), __dadd_rn(P, s3), Pn};
        pgb_agg agg = pgb_classify4(a, qt);
        pgb_agg total;
        pgb_agg ex = warp_excl_scan_agg(agg, &total);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = a[i] + (double)(ex.f.c0 + total.f.c1 + ex.cnt);
      } else if (variant == 6) {  // sampling-grid counts (5 per thread) on a synthetic table
        __shared__ pgb_utable ut;
        if (it == 0 && j == 0) {
          if (tid == 0) pgb_utable_build(&ut, N - 1, 0.37);
          __syncthreads();
        }
        int row = -1;
        long long r = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) r += pgb_utable_count_le_hint(&ut, __dmul_rn(a[i], 1e-6 * (double)(base + i)), &row);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = a[i] + (double)r;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = a[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (base + i < N) y[base + i] = v[i];
      acc.push(warp_tree_sum(tree4(v)));
    }
    if (lane == 0) wpart[w][0] = acc.result();
    __syncthreads();
    if (tid == 0) part[c] = tree8(wpart, 0);
    grid_barrier(bar, target);
  }
}

extern "C" int pgb_test_microbench(int32_t device, int32_t variant, int64_t N, int32_t iters, double* us_per_iter) {
  pgb_handle tmp;
  pgb_handle* h = &tmp;
  struct Cleanup {
    pgb_handle* h;
    ~Cleanup() { for (void* p : h->allocs) cudaFree(p); }
  } cleanup{h};
  CK(h, cudaSetDevice(device));
  double *x, *y, *part;
  unsigned int* bar;
  CK(h, dalloc(h, &x, (size_t)N));
  CK(h, dalloc(h, &y, (size_t)N));
  CK(h, dalloc(h, &part, 1024));
  CK(h, dalloc(h, &bar, 4));
  int tpc = 1;
  const int nt = (int)((N + 1023) / 1024);
  while ((nt + tpc - 1) / tpc > 256) tpc *= 2;
  const int grid = (nt + tpc - 1) / tpc;
  cudaEvent_t e0, e1;
  CK(h, cudaEventCreate(&e0));
  CK(h, cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(h, cudaMemset(bar, 0, 4));
    void* args[] = {(void*)&variant, (void*)&x, (void*)&y, (void*)&part, (void*)&bar, (void*)&N, (void*)&iters, (void*)&tpc};
    CK(h, cudaEventRecord(e0));
    CK(h, cudaLaunchCooperativeKernel((const void*)k_micro, dim3(grid), dim3(256), args, 0, 0));
    CK(h, cudaEventRecord(e1));
    CK(h, cudaEventSynchronize(e1));
  }
  float ms = 0.f;
  CK(h, cudaEventElapsedTime(&ms, e0, e1));
  *us_per_iter = 1e3 * ms / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return PGB_OK;
}
