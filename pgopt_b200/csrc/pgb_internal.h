// pgb_internal.h — what the translation units of libpgopt_b200.so share: the handle, error / launch helpers and
// the entry points each unit exports to the C-ABI unit (pgb_capi.cu).  The library is built from several .cu files
// compiled in parallel (pgopt_b200/build.py); every kernel is still compiled for sm_100a only.
#pragma once
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
#include "pgb_persist.cuh"
#include "pgb_v2_types.h"


struct pgb_handle {
  pgb_config cfg;
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;
  std::string err;
  pgb::FilterDev fd;
  pgb::MniwDev md;
  bool have_data = false, have_meas = false, have_init = false, have_prior = false, have_rng = false;
  std::vector<char> have_state;
  int64_t sweep_index = 1;
  // device
  double *d_u = nullptr, *d_y = nullptr, *d_Q = nullptr, *d_Vdiag = nullptr, *d_LambdaQ = nullptr;
  // warp-per-chain kernel, Philox mode: the normals / uniforms of a whole sweep drawn ahead of the time loop (pgb_smalln.cu)
  double *d_Zpre = nullptr, *d_Upre = nullptr, *d_Apre = nullptr;
  double* d_Vinv = nullptr;  // I / V of a dense prior covariance (pgb_set_prior_dense), allocated on first use
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
  // filter kernel: -1 automatic (warp / one-CTA kernels for N <= 1024 with stored history, else the sweep kernel),
  // 4 the chunk-stationary sweep kernel (pgb_v2.cuh), 1 the tile persistent kernel (pgb_persist.cuh), 0 one launch per
  // phase, 3 the warp / one-CTA kernels
  int persistent = -1;
  pgb::PersistDev pd;
  pgb::V2Dev v2 = {};
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

// K posterior samples (PG_sample fields, PGopt.jl:23-29) resident on one device, plus the scenario buffers the rollout
// writes (x_vec_0, v_vec, e_vec, X, Y of optimal_control_Ipopt.jl:47-125) so that sampler -> rollout -> screening never
// leaves the GPU.  One arena allocation per buffer for the life of the set; `hh` is its allocation list / error sink.
struct pgb_samples {
  pgb_handle hh;
  int64_t K = 0;        // capacity
  int64_t K_active = 0; // slots in use (= number of scenarios of a rollout); = K unless a multi-GPU shard is ragged
  double *A = nullptr, *Q = nullptr, *w = nullptr, *x = nullptr, *um = nullptr;  // [K][...] layouts of pgb_rollout
  double* xT = nullptr; // [K][n_x] x_m1[:, star] (pgb_samples_draw_xT) or null
  bool have_xT = false;
  bool have_particles = false;  // w_m1 / x_m1 hold data (set, store or unpack with particle blocks): the star can be drawn
  // scenario buffers of the last rollout (allocated for horizon H_cap on first use)
  int64_t H = 0, H_cap = 0;
  double *x0 = nullptr, *v = nullptr, *e = nullptr, *X = nullptr, *Y = nullptr, *ui = nullptr, *hmax = nullptr;
  double *ylo = nullptr, *yhi = nullptr;
  int* status = nullptr;
  bool have_scen = false;
  double C[16] = {0}, Le[16] = {0};
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  float last_ms[3] = {0.f, 0.f, 0.f};  // device time of the last rollout kernel, screening kernel, input upload
  // pgb_dynamics_dev scratch (allocated on first use): stacked state in, next state and Jacobian blocks out
  double *dyn_x = nullptr, *dyn_xn = nullptr, *dyn_jx = nullptr, *dyn_ju = nullptr;
  float dyn_ms = 0.f;
};
// pgb_rollout.cu: device-to-device copy of the chain's current PG_sample into slot k (ordered on h->stream, no sync)
int pgb_samples_store_async(pgb_samples* s, int64_t k, pgb_handle* h, int chain);

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

int pgb_set_err(pgb_handle* h, int code, const char* fmt, ...);
#define set_err pgb_set_err

#define CK(h, call)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return set_err(h, PGB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define NEED(h) \
  if (!h) return set_err(nullptr, PGB_ERR_INVALID, "handle is NULL"); \
  CK(h, cudaSetDevice(h->cfg.device))

// Device allocation owned by the handle (freed in pgb_destroy).  Zero-filling is ordered on the handle's stream
// (h->stream is created before the first allocation; the temporary handles of the stand-alone entry points use
// the legacy stream for everything).
template <typename T>
static cudaError_t dalloc(pgb_handle* h, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (e != cudaSuccess) return e;
  if (h) h->allocs.push_back(q);
  if (zero) {
    e = cudaMemsetAsync(q, 0, count * sizeof(T), h ? h->stream : 0);
    if (e != cudaSuccess) return e;
  }
  *p = (T*)q;
  return cudaSuccess;
}

// temporary handle of the stand-alone entry points: an allocation list / error sink that frees on scope exit
struct TmpHandle : pgb_handle {
  ~TmpHandle() {
    for (void* p : allocs) cudaFree(p);
  }
};

int pgb_require_device(void);  // PGB_OK or PGB_ERR_CUDA ("no CUDA device available ...; no CPU fallback")
int pgb_validate_cfg(const pgb_config* c);
void pgb_fill_model(pgb::FilterDev& fd, const pgb_config* c);
extern int g_pgb_force_fallback;

// pgb_tile.cu: the 1024-particle tile kernels (persistent cooperative kernel, one-launch-per-phase path)
cudaError_t pgb_alloc_scanws(pgb_handle* h, pgb::ScanWS* ws, int64_t N, int nc, long long** counters);
int pgb_run_filter_tile_persistent(pgb_handle* h);  // -1: configuration does not fit
int pgb_run_filter_tile_multikernel(pgb_handle* h);
// pgb_smalln.cu: one warp / one CTA per chain (N <= 1024, stored history); -1: does not qualify
int pgb_run_filter_small(pgb_handle* h, int64_t n_max);
// pgb_v2.cu: the chunk-stationary sweep kernel (FP64 exact-order scan, FP32 fixed-point CDF); -1: does not fit
int pgb_run_filter_v2(pgb_handle* h);
int pgb_v2_alloc(pgb_handle* h);
// pgb_capi.cu: what every filter path runs before / after its kernels
int pgb_filter_prologue(pgb_handle* h, int* first);     // snapshot of (A, Q), sweep counters, max key reset
int pgb_launch_trace_and_stats(pgb_handle* h, int first, bool small_trace_done);
