// pgb_capi.cu — host side of libpgopt_b200.so: the C ABI of include/pgopt_b200.h, device memory
// management and the launch sequences shared by every filter path.  No torch types, no CPU compute path: every
// entry point that computes needs a CUDA device and fails with PGB_ERR_CUDA otherwise.
#include "pgb_internal.h"
#include "pgb_trace_f32.cuh"

using namespace pgb;

static thread_local std::string g_last_error;
int g_pgb_force_fallback = 0;

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

int pgb_set_err(pgb_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (h) h->err = buf;
  return code;
}

int pgb_require_device(void) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return set_err(nullptr, PGB_ERR_CUDA, "no CUDA device available (%s); pgopt_b200 has no CPU fallback",
                   cudaGetErrorString(e));
  return PGB_OK;
}

extern "C" const char* pgb_version(void) { return "pgopt_b200 0.2.0 (sm_100a)"; }
extern "C" const char* pgb_last_error(const pgb_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int pgb_validate_cfg(const pgb_config* c) {
  if (!c) return set_err(nullptr, PGB_ERR_INVALID, "config is NULL");
  if (c->n_x < 1 || c->n_x > 4) return set_err(nullptr, PGB_ERR_INVALID, "n_x must be 1..4 (got %d)", c->n_x);
  if (c->n_u < 1 || c->n_u > 4) return set_err(nullptr, PGB_ERR_INVALID, "n_u must be 1..4 (got %d)", c->n_u);
  if (c->n_y < 1 || c->n_y > 4) return set_err(nullptr, PGB_ERR_INVALID, "n_y must be 1..4 (got %d)", c->n_y);
  if (c->precision != PGB_PRECISION_F64 && c->precision != PGB_PRECISION_F32)
    return set_err(nullptr, PGB_ERR_INVALID, "precision must be PGB_PRECISION_F64 or PGB_PRECISION_F32 (got %d)", c->precision);
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

void pgb_fill_model(FilterDev& fd, const pgb_config* c) {
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
  int rc = pgb_validate_cfg(cfg);
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
  pgb_fill_model(fd, cfg);
  fd.N = N; fd.T = T; fd.nc = nc;
  fd.f32 = (cfg->precision == PGB_PRECISION_F32) ? 1 : 0;
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
    const double hist_bytes = (double)nc * (double)T * nx * (double)N * (fd.f32 ? 4.0 : 8.0);
    const bool store = cfg->x_history == PGB_X_HISTORY_STORE ||
                       (cfg->x_history == PGB_X_HISTORY_AUTO && hist_bytes <= 1024.0 * 1024.0 * 1024.0);
    fd.x_rows = store ? T : 2;
    if (T < 2) fd.x_rows = T;
  }
  // F32 mode: the same buffer holds FP32 values (half the FP64 element count, rounded up)
  CKC(dalloc(h, &fd.xh, fd.f32 ? ((size_t)nc * fd.x_rows * nx * N + 1) / 2 : (size_t)nc * fd.x_rows * nx * N, false));
  CKC(dalloc(h, &fd.lineage, (size_t)nc * T));
  CKC(dalloc(h, &h->d_replay_noise, (size_t)nc * T * nx));
  CKC(dalloc(h, &fd.ah, (size_t)nc * T * N, false));
  CKC(cudaMemsetAsync(fd.ah, 0xff, (size_t)nc * T * N * sizeof(int32_t), h->stream));
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
  CKC(pgb_alloc_scanws(h, &fd.main, N, nc, &h->d_counters_main));
  CKC(pgb_alloc_scanws(h, &fd.side, N, nc, &h->d_counters_side));
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
    CKC(dalloc(h, &pd.cand, (size_t)2 * nc));
    CKC(dalloc(h, &pd.sm_slots, (size_t)1024));
    CKC(dalloc(h, &pd.cta_key, (size_t)4096));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, cfg->device));
    h->num_sms = prop.multiProcessorCount;
  }
  h->have_state.assign(nc, 0);
  h->h_u.assign((size_t)nu * T, 0.0);
  CKC(cudaStreamSynchronize(h->stream));  // every zero-fill above has landed before the handle is handed out
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
  h->md.Vinv = nullptr;
  h->have_prior = true;
  return PGB_OK;
}

extern "C" int pgb_set_prior_dense(pgb_handle* h, const double* V, const double* Lambda_Q, double ell_Q) {
  NEED(h);
  if (!V || !Lambda_Q) return set_err(h, PGB_ERR_INVALID, "V/Lambda_Q NULL");
  const size_t np = (size_t)h->cfg.n_phi, pp = np * np;
  std::vector<double> inv(pp), work(pp + np);
  if (pgb_host_spd_inverse(V, (int)np, inv.data(), work.data()))
    return set_err(h, PGB_ERR_NOT_POSDEF, "V is not positive definite");
  if (!h->d_Vinv) CK(h, dalloc(h, &h->d_Vinv, pp, false));
  CK(h, cudaMemcpyAsync(h->d_Vinv, inv.data(), pp * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->d_LambdaQ, Lambda_Q, (size_t)h->cfg.n_x * h->cfg.n_x * 8, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->md.ell_Q = ell_Q;
  h->md.Vinv = h->d_Vinv;
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
  // the stream id (chain_base + chain) occupies 24 bits of the Philox counter (pgb_math.h: pgb_stream_block)
  if ((uint64_t)chain_base + (uint64_t)h->cfg.n_chains > (1ull << 24))
    return set_err(h, PGB_ERR_INVALID, "chain_base + n_chains = %llu exceeds the 2^24 Philox stream ids",
                   (unsigned long long)chain_base + (unsigned long long)h->cfg.n_chains);
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

// What every filter path does first: remember the (A, Q) this sweep runs with (sample fields :204-205), publish the
// sweep counter to the kernels, reset the running max of the log-weights.
int pgb_filter_prologue(pgb_handle* h, int* first) {
  FilterDev& fd = h->fd;
  cudaStream_t s = h->stream;
  const int nc = fd.nc, nx = fd.n_x;
  *first = (h->sweep_index == 1);
  fd.sweep = (uint32_t)h->sweep_index;
  h->md.sweep = fd.sweep;
  const size_t nA = (size_t)nc * nx * fd.n_phi * 8, nQ = (size_t)nc * nx * nx * 8;
  CK(h, cudaMemcpyAsync(h->d_A_used, fd.A, nA, cudaMemcpyDeviceToDevice, s));
  CK(h, cudaMemcpyAsync(h->d_Q_used, h->d_Q, nQ, cudaMemcpyDeviceToDevice, s));
  CK(h, cudaMemsetAsync(fd.maxkey, 0, (size_t)nc * 8, s));
  return PGB_OK;
}

// ... and last: backward trace of the ancestor path (:213-218; replayed when the particle history is not kept)
// and the sufficient statistics (:221-225).
template <int NX>
static int trace_and_stats_nx(pgb_handle* h, int first, bool trace_done) {
  FilterDev& fd = h->fd;
  MniwDev& md = h->md;
  cudaStream_t s = h->stream;
  const int64_t T = fd.T;
  const int nc = fd.nc;
  if (!trace_done && fd.f32) {
    if (fd.x_rows == T) {
      LAUNCH(h, PGB_PROF_TAIL, (k_trace_f32<NX><<<(nc + 31) / 32, 32, 0, s>>>(fd, h->d_star_idx, h->d_star_out)));
    } else {
      LAUNCH(h, PGB_PROF_TAIL, (k_trace_replay_f32<NX><<<nc, 128, (size_t)fd.n_phi * 4, s>>>(fd, h->d_star_idx, h->d_star_out,
                                                                                           h->d_replay_noise, first)));
    }
  } else if (!trace_done) {
    if (fd.x_rows == T) {
      LAUNCH(h, PGB_PROF_TAIL, (k_trace<NX><<<(nc + 31) / 32, 32, 0, s>>>(fd, h->d_star_idx, h->d_star_out)));
    } else {
      LAUNCH(h, PGB_PROF_TAIL, (k_trace_replay<NX><<<nc, TR_TPB, (size_t)fd.n_phi * 8, s>>>(fd, h->d_star_idx, h->d_star_out,
                                                                                          h->d_replay_noise, first)));
    }
  }
  LAUNCH(h, PGB_PROF_TAIL, (k_stats_z<NX><<<dim3((unsigned)((T - 1 + 127) / 128), nc), 128, 0, s>>>(fd, md)));
  const int64_t nent = (int64_t)NX * NX + (int64_t)NX * fd.n_phi + (int64_t)fd.n_phi * fd.n_phi;
  LAUNCH(h, PGB_PROF_TAIL, (k_stats<<<dim3((unsigned)nent, nc), TPB, 0, s>>>(md)));
  CK(h, cudaGetLastError());
  return PGB_OK;
}
int pgb_launch_trace_and_stats(pgb_handle* h, int first, bool small_trace_done) {
  switch (h->cfg.n_x) {
    case 1: return trace_and_stats_nx<1>(h, first, small_trace_done);
    case 2: return trace_and_stats_nx<2>(h, first, small_trace_done);
    case 3: return trace_and_stats_nx<3>(h, first, small_trace_done);
    default: return trace_and_stats_nx<4>(h, first, small_trace_done);
  }
}

static int run_filter(pgb_handle* h) {
  int mode = h->persistent;
  int rc;
  if (h->fd.f32) {  // the F32 precision mode exists in the chunk-stationary sweep kernel only
    rc = pgb_run_filter_v2(h);
    if (rc >= 0) return rc;
    return set_err(h, PGB_ERR_INVALID, "PGB_PRECISION_F32: this shape (n_chains = %d, N = %lld) does not fit the sweep kernel", h->fd.nc,
                   (long long)h->fd.N);
  }
  if (mode < 0 || mode == 3) {
    // the warp kernel up to N = 32, the one-CTA-per-chain kernel up to N = 1024 (automatic and mode 3)
    rc = pgb_run_filter_small(h, 1024);
    if (rc >= 0) return rc;
    mode = (mode == 3) ? 1 : 4;
  }
  if (mode == 4) {  // a configuration that does not fit falls through to the tile kernels
    rc = pgb_run_filter_v2(h);
    if (rc >= 0) return rc;
    mode = 1;
  }
  if (mode >= 1) {
    rc = pgb_run_filter_tile_persistent(h);
    if (rc >= 0) return rc;
  }
  return pgb_run_filter_tile_multikernel(h);
}

static int run_mniw(pgb_handle* h) {
  if (!h->have_prior) return set_err(h, PGB_ERR_STATE, "pgb_set_prior must be called before the MNIW draw");
  const size_t sm = mniw_smem_bytes(h->md.n_phi);
  if (sm) CK(h, cudaFuncSetAttribute(k_mniw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  LAUNCH(h, PGB_PROF_MNIW, (k_mniw<<<h->fd.nc, MNIW_TPB, sm, h->stream>>>(h->md, sm ? 1 : 0)));
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
  const bool f32 = h->fd.f32 != 0;
  if (x_m1) {
    tmp.resize((size_t)nx * N);
    const size_t row = (h->fd.x_rows == T) ? (size_t)(T - 1) : (size_t)((T - 1) & 1);
    const size_t off = ((size_t)chain * h->fd.x_rows + row) * nx * N;
    if (f32)
      CK(h, cudaMemcpyAsync(tmp.data(), reinterpret_cast<const float*>(h->fd.xh) + off, (size_t)nx * N * 4, cudaMemcpyDeviceToHost, s));
    else
      CK(h, cudaMemcpyAsync(tmp.data(), h->fd.xh + off, (size_t)nx * N * 8, cudaMemcpyDeviceToHost, s));
  }
  CK(h, cudaStreamSynchronize(s));
  if (x_m1) {
    const float* tf = reinterpret_cast<const float*>(tmp.data());
    for (int64_t n = 0; n < N; ++n)
      for (int d = 0; d < nx; ++d) x_m1[d + nx * n] = f32 ? (double)tf[(size_t)d * N + n] : tmp[(size_t)d * N + n];
  }
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
  // particle_Gibbs.jl:154 always starts at k = 1 (bootstrap filter on the first iteration, also with a supplied x_prim)
  if (h->sweep_index != 1)
    return set_err(h, PGB_ERR_STATE, "pgb_particle_gibbs starts a sampler at sweep 1: call pgb_set_state(..., sweep_index = 1) "
                                     "first (handle is at sweep %lld)", (long long)h->sweep_index);
  const int nc = h->cfg.n_chains, nx = h->cfg.n_x, np = h->cfg.n_phi;
  const int64_t N = h->cfg.N;
  const int64_t K_total = K_b + 1 + (K - 1) * (k_d + 1);  // particle_Gibbs.jl:116
  std::vector<double> tmp;
  int64_t cur = 0;
  for (int64_t k = 1; k <= K_total; ++k) {
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

// The sampler with the K kept samples of every chain left ON THE DEVICE (slot chain*K + k of `out`), copied
// device-to-device at the reference's store condition (:203): nothing but the positive-definiteness flag crosses PCIe.
extern "C" int pgb_particle_gibbs_dev(pgb_handle* h, int64_t K, int64_t K_b, int64_t k_d, pgb_samples* out) {
  NEED(h);
  int rc = check_ready(h);
  if (rc) return rc;
  if (!out) return set_err(h, PGB_ERR_INVALID, "sample set is NULL");
  if (!h->fd.philox) return set_err(h, PGB_ERR_INVALID, "pgb_particle_gibbs_dev needs Philox streams");
  if (!h->have_prior) return set_err(h, PGB_ERR_STATE, "pgb_set_prior missing");
  if (K < 1 || K_b < 0 || k_d < 0) return set_err(h, PGB_ERR_INVALID, "bad K/K_b/k_d");
  if (h->sweep_index != 1)
    return set_err(h, PGB_ERR_STATE, "pgb_particle_gibbs_dev starts a sampler at sweep 1: call pgb_set_state(..., sweep_index = 1) "
                                     "first (handle is at sweep %lld)", (long long)h->sweep_index);
  const int nc = h->cfg.n_chains;
  if (out->K < (int64_t)nc * K) return set_err(h, PGB_ERR_INVALID, "sample set holds %lld slots, n_chains*K = %lld needed",
                                                 (long long)out->K, (long long)nc * K);
  const int64_t K_total = K_b + 1 + (K - 1) * (k_d + 1);  // particle_Gibbs.jl:116
  int64_t cur = 0;
  for (int64_t k = 1; k <= K_total; ++k) {
    if ((rc = run_filter(h))) return rc;
    if (K_b < k && ((k - (K_b + 1)) % (k_d + 1) == 0) && cur < K) {  // :203
      for (int c = 0; c < nc; ++c)
        if ((rc = pgb_samples_store_async(out, (int64_t)c * K + cur, h, c))) return rc;
      ++cur;
    }
    if ((rc = run_mniw(h))) return rc;
    if ((k & 63) == 0 && (rc = check_posdef(h))) return rc;
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
    const bool f32 = h->fd.f32 != 0;
    if (f32)
      CK(h, cudaMemcpy(tx.data(), reinterpret_cast<const float*>(h->fd.xh) + (size_t)chain * T * nx * N, tx.size() * 4, cudaMemcpyDeviceToHost));
    else
      CK(h, cudaMemcpy(tx.data(), h->fd.xh + (size_t)chain * T * nx * N, tx.size() * 8, cudaMemcpyDeviceToHost));
    const float* tf = reinterpret_cast<const float*>(tx.data());
    for (int64_t t = 0; t < T; ++t)
      for (int64_t n = 0; n < N; ++n)
        for (int d = 0; d < nx; ++d) {
          const size_t i = ((size_t)t * nx + d) * N + n;
          x[(t * N + n) * nx + d] = f32 ? (double)tf[i] : tx[i];
        }
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

static int mniw_sample_impl(int32_t device, int32_t n_x, int32_t n_phi, const double* Phi, const double* Psi,
                            const double* Sigma, const double* V_diag, bool dense, const double* Lambda_Q, double ell_Q,
                            int64_t T, const double* bartlett, const double* Xmn, uint64_t seed, uint32_t chain,
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
  CK(h, dalloc(h, &dV, dense ? pp : (size_t)n_phi));
  CK(h, dalloc(h, &dL, xx));
  CK(h, dalloc(h, &md.Phi, xx)); CK(h, dalloc(h, &md.Psi, xp)); CK(h, dalloc(h, &md.Sigma, pp));
  CK(h, dalloc(h, &md.Ls, pp)); CK(h, dalloc(h, &md.Sinv, pp)); CK(h, dalloc(h, &md.Lv, pp));
  CK(h, dalloc(h, &md.Mpost, xp)); CK(h, dalloc(h, &md.Xn, xp)); CK(h, dalloc(h, &md.LX, xp));
  CK(h, dalloc(h, &md.A, xp)); CK(h, dalloc(h, &dQ, xx)); CK(h, dalloc(h, &md.Lq, 16)); CK(h, dalloc(h, &md.c0, 1));
  CK(h, dalloc(h, &md.status, 1));
  md.Q = dQ; md.Lambda_Q = dL;
  if (dense) {  // V_diag points at the dense V: I / V on the host, as pgb_set_prior_dense
    std::vector<double> inv(pp), work(pp + (size_t)n_phi);
    if (pgb_host_spd_inverse(V_diag, n_phi, inv.data(), work.data()))
      return set_err(nullptr, PGB_ERR_NOT_POSDEF, "V is not positive definite");
    CK(h, cudaMemcpy(dV, inv.data(), pp * 8, cudaMemcpyHostToDevice));
    md.Vinv = dV;
  } else {
    CK(h, cudaMemcpy(dV, V_diag, (size_t)n_phi * 8, cudaMemcpyHostToDevice));
    md.V_diag = dV;
  }
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
  const size_t sm = mniw_smem_bytes(n_phi);
  if (sm) CK(h, cudaFuncSetAttribute(k_mniw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_mniw<<<1, MNIW_TPB, sm>>>(md, sm ? 1 : 0);
  CK(h, cudaGetLastError());
  CK(h, cudaDeviceSynchronize());
  int st = 0;
  CK(h, cudaMemcpy(&st, md.status, 4, cudaMemcpyDeviceToHost));
  if (st & PGB_STATUS_NOT_POSDEF) return set_err(nullptr, PGB_ERR_NOT_POSDEF, "matrix not positive definite");
  CK(h, cudaMemcpy(A_out, md.A, xp * 8, cudaMemcpyDeviceToHost));
  CK(h, cudaMemcpy(Q_out, dQ, xx * 8, cudaMemcpyDeviceToHost));
  return PGB_OK;
}

extern "C" int pgb_mniw_sample(int32_t device, int32_t n_x, int32_t n_phi, const double* Phi, const double* Psi,
                               const double* Sigma, const double* V_diag, const double* Lambda_Q, double ell_Q, int64_t T,
                               const double* bartlett, const double* Xmn, uint64_t seed, uint32_t chain,
                               uint32_t sweep_index, double* A_out, double* Q_out) {
  return mniw_sample_impl(device, n_x, n_phi, Phi, Psi, Sigma, V_diag, false, Lambda_Q, ell_Q, T, bartlett, Xmn, seed,
                          chain, sweep_index, A_out, Q_out);
}

extern "C" int pgb_mniw_sample_dense(int32_t device, int32_t n_x, int32_t n_phi, const double* Phi, const double* Psi,
                                     const double* Sigma, const double* V, const double* Lambda_Q, double ell_Q, int64_t T,
                                     const double* bartlett, const double* Xmn, uint64_t seed, uint32_t chain,
                                     uint32_t sweep_index, double* A_out, double* Q_out) {
  return mniw_sample_impl(device, n_x, n_phi, Phi, Psi, Sigma, V, true, Lambda_Q, ell_Q, T, bartlett, Xmn, seed, chain,
                          sweep_index, A_out, Q_out);
}

extern "C" int pgb_set_mode(pgb_handle* h, int32_t persistent) {
  NEED(h);
  h->persistent = persistent < 0 ? -1 : (persistent > 4 ? 4 : persistent);
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
extern "C" void pgb_test_force_fallback(int32_t on) { g_pgb_force_fallback = on; }

__global__ void k_test_math(int fn, const double* __restrict__ x, double* __restrict__ y, int64_t n, uint64_t seed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (fn) {
    case 0: y[i] = pgb_exp(x[i]); break;
    case 1: y[i] = pgb_log(x[i]); break;
    case 2: y[i] = pgb_sin(x[i]); break;
    case 3: y[i] = pgb_cos(x[i]); break;
    case 4: pgb_normal_pair(pgb_stream_block(seed, PGB_STREAM_Z, 0, 1, 5, (uint32_t)i), &y[2 * i], &y[2 * i + 1]); break;
    // 30..35: the FP32 functions of pgb_mathf.h (F32 precision mode) through FP64 in/out: x is rounded to FP32, the FP32
    // result widened — bit identity of the device build with the host build (oracle) is checked on these
    case 30: y[i] = (double)pgf_exp((float)x[i]); break;
    case 31: y[i] = (double)pgf_log_pos_normal((float)x[i]); break;
    case 32: y[i] = (double)pgf_sin((float)x[i]); break;
    case 33: y[i] = (double)pgf_cos((float)x[i]); break;
    case 34: y[i] = pgf_exp_wide((float)x[i]); break;
    case 35: {  // x[0] = seed (read by the host wrapper); 4 FP32 normals of particle i, purpose Z, sweep 1, t 5
      float z[4];
      pgf_stream_normals(seed, PGB_STREAM_Z, 0, 1, 5, (uint32_t)i, 4, z);
      for (int d = 0; d < 4; ++d) y[4 * i + d] = (double)z[d];
      break;
    }
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
  if (int rcd = pgb_require_device()) return rcd;
  TmpHandle tmp;
  pgb_handle* h = &tmp;
  CK(h, cudaSetDevice(device));
  const bool pairs = (fn == 4 || fn == 14 || fn == 35);
  const int64_t nout = fn == 35 ? 4 * n : (pairs ? 2 * n : n);
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
  if (int rcd = pgb_require_device()) return rcd;
  TmpHandle tmp;
  pgb_handle* h = &tmp;
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
