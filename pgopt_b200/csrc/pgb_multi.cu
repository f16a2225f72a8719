// pgb_multi.cu — the chain- and scenario-sharded paths over several GPUs of one node, inside the C ABI
// (SURVEY.md section 8b `pgb_multi_*`, section 8e; VERDICT r1 "missing" item 4).
//
// One process drives n_dev devices: an ordinary pgb_handle with the device's share of the chains and a device-resident
// pgb_samples set per device, one host thread per device while the samplers / rollouts run (they are independent: no
// collective on the data path), and ONE NCCL all-gather per array at the end — straight out of and into device buffers,
// no host staging:
//   independent chains    : the kept samples stay where their chain ran; x_T = x_m1[:, star] is drawn on the owning GPU
//                           (optimal_control_Ipopt.jl:58-59) and only the packed (A, Q, x_T, u_m1) records travel;
//   independent scenarios : every device rolls out the scenarios of its own samples (global Philox stream ids, so the
//                           result does not depend on the device count); x_vec_0 / X / Y are all-gathered.
// NCCL is bound at run time (dlopen("libnccl.so.2"): the copy the process already loaded, e.g. torch's, else the system
// library) so that libpgopt_b200.so itself has no link-time dependency on it; communicators come from ncclCommInitAll.
#include <dlfcn.h>
#include <nccl.h>

#include <thread>

#include "pgb_internal.h"

using namespace pgb;

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    api.err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return &api;
  }
#define SYM(field, name)                                                         \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));       \
  if (!api.field) api.err = std::string("NCCL symbol missing: ") + name
  SYM(CommInitAll, "ncclCommInitAll");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GetVersion, "ncclGetVersion");
#undef SYM
  return &api;
}

}  // namespace

struct pgb_multi_dev {
  int device = 0;
  int c0 = 0, nc = 0;             // chains [c0, c0 + nc) of the global numbering live here
  int64_t g0 = 0, gn = 0;         // global samples / scenarios [g0, g0 + gn) are resident in S
  pgb_handle* h = nullptr;
  pgb_samples* S = nullptr;       // the device's own kept samples, capacity `cap` (created by the first sampler run)
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;  // gather stream
  cudaStream_t stream2 = nullptr; // device-to-host stream (the owning device's block of the outputs, concurrent with the gather)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // gathered arrays, [n_dev][cap][...]: packed records, scenario outputs
  double* g_rec = nullptr;
  double *g_x0 = nullptr, *g_X = nullptr, *g_Y = nullptr;
  double* rec = nullptr;          // [cap][rec] this device's packed records (send buffer)
  int64_t gH = 0;
  int rc = 0;                     // result of the last per-device call
  std::string err;
};

struct pgb_multi {
  pgb_config cfg;
  int n_dev = 0;
  std::vector<pgb_multi_dev> dev;
  int64_t K = 0, cap = 0;         // samples per chain of the last sampler run; slots per device (max chains * K)
  uint64_t roll_seed = 0;
  bool have_samples = false, have_scen = false;
  std::string err;
  float gather_ms = 0.f, rollout_gather_ms = 0.f;
  bool scen_gather_pending = false;  // the all-gather of the last rollout's x_vec_0 / X / Y may still be in flight
  double rollout_kernel_ms_max = 0.0;
};

static int merr(pgb_multi* m, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (m) m->err = buf;
  return pgb_set_err(nullptr, code, "%s", buf);
}

// run fn(d) on every device concurrently (one host thread each); first failure wins
template <typename F>
static int for_each_device(pgb_multi* m, F fn) {
  std::vector<std::thread> th;
  for (int d = 0; d < m->n_dev; ++d)
    th.emplace_back([m, d, &fn]() {
      pgb_multi_dev& D = m->dev[d];
      cudaSetDevice(D.device);
      D.rc = fn(D);
      if (D.rc) {
        const char* e = pgb_last_error(nullptr);  // thread-local message of this worker
        D.err = e ? e : "";
      }
    });
  for (auto& t : th) t.join();
  for (int d = 0; d < m->n_dev; ++d)
    if (m->dev[d].rc) return merr(m, m->dev[d].rc, "device %d: %s", m->dev[d].device, m->dev[d].err.c_str());
  return PGB_OK;
}

// pgb_multi_rollout returns when the host arrays are complete; the NVLink all-gather of the scenario arrays finishes
// behind it.  Everything that reads the gathered copies or overwrites the gather's sources waits here first.
static void wait_scenario_gather(pgb_multi* m) {
  if (!m->scen_gather_pending) return;
  for (auto& D : m->dev) {
    cudaSetDevice(D.device);
    cudaStreamSynchronize(D.stream);
  }
  cudaSetDevice(m->dev[0].device);
  cudaEventElapsedTime(&m->rollout_gather_ms, m->dev[0].ev0, m->dev[0].ev1);
  m->scen_gather_pending = false;
}

extern "C" int64_t pgb_packed_record_doubles(int32_t n_x, int32_t n_phi, int32_t n_u) {
  return (int64_t)n_x * n_phi + (int64_t)n_x * n_x + n_x + n_u;
}

extern "C" const char* pgb_multi_last_error(const pgb_multi* m) { return m ? m->err.c_str() : pgb_last_error(nullptr); }

extern "C" void pgb_multi_destroy(pgb_multi* m) {
  if (!m) return;
  wait_scenario_gather(m);
  NcclApi* nc = nccl_api();
  for (auto& D : m->dev) {
    cudaSetDevice(D.device);
    if (D.stream) cudaStreamSynchronize(D.stream);
    if (D.comm && nc->CommDestroy) nc->CommDestroy(D.comm);
    if (D.S) pgb_samples_destroy(D.S);
    if (D.h) pgb_destroy(D.h);
    for (double* p : {D.g_rec, D.g_x0, D.g_X, D.g_Y, D.rec})
      if (p) cudaFree(p);
    if (D.ev0) cudaEventDestroy(D.ev0);
    if (D.ev1) cudaEventDestroy(D.ev1);
    if (D.stream) cudaStreamDestroy(D.stream);
    if (D.stream2) cudaStreamDestroy(D.stream2);
  }
  delete m;
}

extern "C" int pgb_multi_create(pgb_multi** out, const pgb_config* cfg, const int32_t* devices, int32_t n_devices) {
  if (!out) return merr(nullptr, PGB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = pgb_validate_cfg(cfg);
  if (rc) return rc;
  if (!devices || n_devices < 1) return merr(nullptr, PGB_ERR_INVALID, "need a device list");
  if (cfg->n_chains < n_devices)
    return merr(nullptr, PGB_ERR_INVALID, "n_chains = %d is the TOTAL number of chains and must be >= the %d devices", cfg->n_chains, n_devices);
  if ((rc = pgb_require_device())) return rc;
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  for (int i = 0; i < n_devices; ++i) {
    if (devices[i] < 0 || devices[i] >= ndev) return merr(nullptr, PGB_ERR_INVALID, "bad device ordinal %d (%d visible)", devices[i], ndev);
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) return merr(nullptr, PGB_ERR_INVALID, "device %d listed twice", devices[i]);
  }
  NcclApi* nc = nccl_api();
  if (n_devices > 1 && !nc->err.empty()) return merr(nullptr, PGB_ERR_CUDA, "NCCL unavailable: %s", nc->err.c_str());
  pgb_multi* m = new pgb_multi();
  m->cfg = *cfg;
  m->n_dev = n_devices;
  m->dev.resize(n_devices);
  const int total = cfg->n_chains;
  for (int d = 0; d < n_devices; ++d) {  // contiguous, balanced blocks of chains (the shard_range of pgopt_b200/multi.py)
    pgb_multi_dev& D = m->dev[d];
    D.device = devices[d];
    D.c0 = (int)((int64_t)total * d / n_devices);
    D.nc = (int)((int64_t)total * (d + 1) / n_devices) - D.c0;
  }
  rc = for_each_device(m, [m](pgb_multi_dev& D) -> int {
    pgb_config c = m->cfg;
    c.device = D.device;
    c.n_chains = D.nc;
    int r = pgb_create(&D.h, &c);
    if (r) return r;
    if (cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&D.stream2, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&D.ev0) != cudaSuccess ||
        cudaEventCreate(&D.ev1) != cudaSuccess)
      return pgb_set_err(nullptr, PGB_ERR_CUDA, "stream / event creation failed");
    return PGB_OK;
  });
  if (rc) {
    pgb_multi_destroy(m);
    return rc;
  }
  if (n_devices > 1) {
    std::vector<ncclComm_t> comms(n_devices);
    std::vector<int> devs(devices, devices + n_devices);
    ncclResult_t r = nc->CommInitAll(comms.data(), n_devices, devs.data());
    if (r != ncclSuccess) {
      rc = merr(nullptr, PGB_ERR_CUDA, "ncclCommInitAll failed: %s", nc->GetErrorString(r));
      pgb_multi_destroy(m);
      return rc;
    }
    for (int d = 0; d < n_devices; ++d) m->dev[d].comm = comms[d];
  }
  *out = m;
  return PGB_OK;
}

#define NEEDM(m) \
  if (!m) return merr(nullptr, PGB_ERR_INVALID, "multi handle is NULL")

extern "C" int pgb_multi_n_devices(const pgb_multi* m) { return m ? m->n_dev : 0; }

extern "C" int pgb_multi_set_data(pgb_multi* m, const double* u, const double* y) {
  NEEDM(m);
  return for_each_device(m, [=](pgb_multi_dev& D) { return pgb_set_data(D.h, u, y); });
}
extern "C" int pgb_multi_set_measurement(pgb_multi* m, const double* C, const double* R) {
  NEEDM(m);
  return for_each_device(m, [=](pgb_multi_dev& D) { return pgb_set_measurement(D.h, C, R); });
}
extern "C" int pgb_multi_set_init_dist(pgb_multi* m, const double* mean, const double* chol_lower) {
  NEEDM(m);
  return for_each_device(m, [=](pgb_multi_dev& D) { return pgb_set_init_dist(D.h, mean, chol_lower); });
}
extern "C" int pgb_multi_set_prior(pgb_multi* m, const double* V_diag, const double* Lambda_Q, double ell_Q) {
  NEEDM(m);
  return for_each_device(m, [=](pgb_multi_dev& D) { return pgb_set_prior(D.h, V_diag, Lambda_Q, ell_Q); });
}

extern "C" int pgb_multi_set_prior_dense(pgb_multi* m, const double* V, const double* Lambda_Q, double ell_Q) {
  NEEDM(m);
  return for_each_device(m, [=](pgb_multi_dev& D) { return pgb_set_prior_dense(D.h, V, Lambda_Q, ell_Q); });
}
// chain: global chain index, or -1 = the same state for every chain
extern "C" int pgb_multi_set_state(pgb_multi* m, int32_t chain, const double* A, const double* Q, const double* x_prim,
                                   int64_t sweep_index) {
  NEEDM(m);
  if (chain < -1 || chain >= m->cfg.n_chains) return merr(m, PGB_ERR_INVALID, "bad chain %d", chain);
  return for_each_device(m, [=](pgb_multi_dev& D) -> int {
    for (int c = 0; c < D.nc; ++c) {
      if (chain >= 0 && chain != D.c0 + c) continue;
      int r = pgb_set_state(D.h, c, A, Q, x_prim, sweep_index);
      if (r) return r;
    }
    return PGB_OK;
  });
}
// global chain c draws from Philox stream id chain_base + c whatever the device count
extern "C" int pgb_multi_set_rng_philox(pgb_multi* m, uint64_t seed, uint32_t chain_base) {
  NEEDM(m);
  return for_each_device(m, [=](pgb_multi_dev& D) { return pgb_set_rng_philox(D.h, seed, chain_base + (uint32_t)D.c0); });
}

// record s of this device <- (A, Q, x_T, u_m1) of slot s, contiguous
static __global__ void k_pack_records(int64_t n, int nA, int nQ, int nx, int nu, const double* __restrict__ A,
                                      const double* __restrict__ Q, const double* __restrict__ xT,
                                      const double* __restrict__ um, double* __restrict__ rec) {
  const int R = nA + nQ + nx + nu;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * R; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / R;
    const int j = (int)(i - s * R);
    double v;
    if (j < nA) v = A[s * nA + j];
    else if (j < nA + nQ) v = Q[s * nQ + (j - nA)];
    else if (j < nA + nQ + nx) v = xT[s * nx + (j - nA - nQ)];
    else v = um[s * nu + (j - nA - nQ - nx)];
    rec[i] = v;
  }
}

// one all-gather of `count` doubles per device from send[d] into recv[d] ([n_dev][count]) on every device's gather stream
static int all_gather(pgb_multi* m, const std::vector<const double*>& send, const std::vector<double*>& recv, size_t count) {
  NcclApi* nc = nccl_api();
  if (m->n_dev == 1) {
    pgb_multi_dev& D = m->dev[0];
    cudaSetDevice(D.device);
    if (cudaMemcpyAsync(recv[0], send[0], count * 8, cudaMemcpyDeviceToDevice, D.stream) != cudaSuccess)
      return merr(m, PGB_ERR_CUDA, "device copy failed");
    return PGB_OK;
  }
  ncclResult_t r = nc->GroupStart();
  for (int d = 0; d < m->n_dev && r == ncclSuccess; ++d)
    r = nc->AllGather(send[d], recv[d], count, ncclDouble, m->dev[d].comm, m->dev[d].stream);
  ncclResult_t r2 = nc->GroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess)
    return merr(m, PGB_ERR_CUDA, "ncclAllGather failed: %s", nc->GetErrorString(r != ncclSuccess ? r : r2));
  return PGB_OK;
}

// The sampler on every chain of every device (particle_Gibbs.jl:114-237), the kept samples left on the device that
// produced them; then star / x_T on the owning GPU (roll_seed, stream id = global sample index c*K + k) and ONE NCCL
// all-gather of the packed records.  packed (host, may be NULL): [n_chains*K][pgb_packed_record_doubles] in global order.
extern "C" int pgb_multi_particle_gibbs(pgb_multi* m, int64_t K, int64_t K_b, int64_t k_d, uint64_t roll_seed, double* packed) {
  NEEDM(m);
  wait_scenario_gather(m);
  if (K < 1) return merr(m, PGB_ERR_INVALID, "bad K");
  int maxnc = 0;
  for (auto& D : m->dev) maxnc = std::max(maxnc, D.nc);
  const int64_t cap = (int64_t)maxnc * K;
  if ((uint64_t)m->cfg.n_chains * (uint64_t)K > (1ull << 24)) return merr(m, PGB_ERR_INVALID, "n_chains*K exceeds the 2^24 Philox stream ids");
  const int nx = m->cfg.n_x, np = m->cfg.n_phi, nu = m->cfg.n_u;
  const int64_t R = pgb_packed_record_doubles(nx, np, nu);
  if (m->cap != cap) {  // (re)size the per-device sets and gather buffers
    for (auto& D : m->dev) {
      cudaSetDevice(D.device);
      if (D.S) pgb_samples_destroy(D.S);
      D.S = nullptr;
      for (double** p : {&D.g_rec, &D.rec, &D.g_x0, &D.g_X, &D.g_Y}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
      }
      D.gH = 0;
    }
    m->cap = cap;
  }
  m->K = K;
  m->roll_seed = roll_seed;
  m->have_samples = false;
  m->have_scen = false;
  int rc = for_each_device(m, [=](pgb_multi_dev& D) -> int {
    if (!D.S) {
      pgb_config c = m->cfg;
      c.device = D.device;
      c.n_chains = 1;
      int r = pgb_samples_create(&D.S, &c, cap);
      if (r) return r;
      if (cudaMalloc(&D.rec, (size_t)cap * R * 8) != cudaSuccess || cudaMalloc(&D.g_rec, (size_t)m->n_dev * cap * R * 8) != cudaSuccess)
        return pgb_set_err(nullptr, PGB_ERR_CUDA, "cudaMalloc of the gather buffers failed");
      cudaMemset(D.rec, 0, (size_t)cap * R * 8);
    }
    D.S->K_active = D.gn = (int64_t)D.nc * K;
    D.g0 = (int64_t)D.c0 * K;
    int r = pgb_particle_gibbs_dev(D.h, K, K_b, k_d, D.S);
    if (r) return r;
    // star ~ w_m1 on the owning GPU: the draw of optimal_control_Ipopt.jl:58-59 with the scenario's own stream
    r = pgb_samples_draw_xT(D.S, roll_seed, (uint32_t)((int64_t)D.c0 * K), nullptr);
    if (r) return r;
    const int64_t n = D.S->K_active;
    k_pack_records<<<(unsigned)std::min<int64_t>((n * R + 255) / 256, 4096), 256, 0, D.stream>>>(n, nx * np, nx * nx, nx, nu, D.S->A, D.S->Q,
                                                                                             D.S->xT, D.S->um, D.rec);
    if (cudaGetLastError() != cudaSuccess) return pgb_set_err(nullptr, PGB_ERR_CUDA, "k_pack_records launch failed");
    if (cudaStreamSynchronize(D.stream) != cudaSuccess) return pgb_set_err(nullptr, PGB_ERR_CUDA, "pack failed");
    return PGB_OK;
  });
  if (rc) return rc;
  // the gather: device buffers in, device buffers out, timed on device 0's gather stream
  std::vector<const double*> send;
  std::vector<double*> recv;
  for (auto& D : m->dev) { send.push_back(D.rec); recv.push_back(D.g_rec); }
  cudaSetDevice(m->dev[0].device);
  cudaEventRecord(m->dev[0].ev0, m->dev[0].stream);
  if ((rc = all_gather(m, send, recv, (size_t)cap * R))) return rc;
  cudaSetDevice(m->dev[0].device);
  cudaEventRecord(m->dev[0].ev1, m->dev[0].stream);
  for (auto& D : m->dev) {
    cudaSetDevice(D.device);
    if (cudaStreamSynchronize(D.stream) != cudaSuccess) return merr(m, PGB_ERR_CUDA, "gather failed on device %d: %s", D.device, cudaGetErrorString(cudaGetLastError()));
  }
  cudaSetDevice(m->dev[0].device);
  cudaEventElapsedTime(&m->gather_ms, m->dev[0].ev0, m->dev[0].ev1);
  m->have_samples = true;
  if (packed) return pgb_multi_get_packed(m, 0, packed);
  return PGB_OK;
}

// the gathered records as device `which` of the list holds them (every device holds all of them) -> host, compacted
extern "C" int pgb_multi_get_packed(pgb_multi* m, int32_t which, double* packed) {
  NEEDM(m);
  if (!m->have_samples || !packed) return merr(m, PGB_ERR_STATE, "no gathered samples / NULL output");
  if (which < 0 || which >= m->n_dev) return merr(m, PGB_ERR_INVALID, "bad device index %d", which);
  const int64_t R = pgb_packed_record_doubles(m->cfg.n_x, m->cfg.n_phi, m->cfg.n_u);
  pgb_multi_dev& W = m->dev[which];
  cudaSetDevice(W.device);
  for (int d = 0; d < m->n_dev; ++d) {
    const pgb_multi_dev& D = m->dev[d];
    const size_t n = (size_t)D.gn * R;
    if (cudaMemcpy(packed + (size_t)D.g0 * R, W.g_rec + (size_t)d * m->cap * R, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
      return merr(m, PGB_ERR_CUDA, "copy of the gathered records failed");
  }
  return PGB_OK;
}

// Scenario rollout of ALL kept samples (scenario g = global sample index c*K + k, Philox stream id g), every device
// rolling out its own samples in place, then one all-gather each of x_vec_0 / X / Y.  Host outputs (may be NULL) in
// global order: x0[n_chains*K][n_x], X[..][H+1][n_x], Y[..][H][n_y].  seed must be the roll_seed of the sampler call for
// the x_T already drawn to be the star of this rollout (it is checked).
extern "C" int pgb_multi_rollout(pgb_multi* m, int64_t H, const double* C, const double* Le, const double* u_init, uint64_t seed,
                                 double* x0, double* X, double* Y) {
  NEEDM(m);
  if (!m->have_samples) return merr(m, PGB_ERR_STATE, "pgb_multi_particle_gibbs (or pgb_multi_set_samples) must run first");
  if (seed != m->roll_seed) return merr(m, PGB_ERR_INVALID, "seed differs from the roll_seed the x_T were drawn with");
  wait_scenario_gather(m);  // the previous gather reads the buffers this rollout overwrites
  const int nx = m->cfg.n_x, ny = m->cfg.n_y;
  const int64_t cap = m->cap;
  int rc = for_each_device(m, [=](pgb_multi_dev& D) -> int {
    if (D.gH != H) {
      for (double** p : {&D.g_x0, &D.g_X, &D.g_Y}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
      }
      if (cudaMalloc(&D.g_x0, (size_t)m->n_dev * cap * nx * 8) != cudaSuccess ||
          cudaMalloc(&D.g_X, (size_t)m->n_dev * cap * (H + 1) * nx * 8) != cudaSuccess ||
          cudaMalloc(&D.g_Y, (size_t)m->n_dev * cap * H * ny * 8) != cudaSuccess)
        return pgb_set_err(nullptr, PGB_ERR_CUDA, "cudaMalloc of the scenario gather buffers failed");
      D.gH = H;
    }
    int r = pgb_rollout_dev(D.S, H, C, Le, u_init, seed, (uint32_t)D.g0, 0, nullptr, nullptr, nullptr);
    if (r) return r;
    // host outputs: every device sends ITS OWN block over its own PCIe link (stream2) as soon as ITS kernel is done —
    // not after the slowest device — and concurrently with the NVLink gather below
    const size_t n = (size_t)D.gn, o = (size_t)D.g0;
    cudaError_t e = cudaSuccess;
    if (x0 && e == cudaSuccess) e = cudaMemcpyAsync(x0 + o * nx, D.S->x0, n * nx * 8, cudaMemcpyDeviceToHost, D.stream2);
    if (X && e == cudaSuccess) e = cudaMemcpyAsync(X + o * (H + 1) * nx, D.S->X, n * (H + 1) * nx * 8, cudaMemcpyDeviceToHost, D.stream2);
    if (Y && e == cudaSuccess) e = cudaMemcpyAsync(Y + o * H * ny, D.S->Y, n * H * ny * 8, cudaMemcpyDeviceToHost, D.stream2);
    if (e != cudaSuccess) return pgb_set_err(nullptr, PGB_ERR_CUDA, "D2H failed: %s", cudaGetErrorString(e));
    return PGB_OK;
  });
  if (rc) return rc;
  m->rollout_kernel_ms_max = 0.0;
  for (auto& D : m->dev) m->rollout_kernel_ms_max = std::max(m->rollout_kernel_ms_max, (double)D.S->last_ms[0]);
  cudaSetDevice(m->dev[0].device);
  cudaEventRecord(m->dev[0].ev0, m->dev[0].stream);
  struct Arr { double* pgb_multi_dev::*g; double* pgb_samples::*s; size_t per; };
  const Arr arrs[3] = {{&pgb_multi_dev::g_x0, &pgb_samples::x0, (size_t)nx},
                       {&pgb_multi_dev::g_X, &pgb_samples::X, (size_t)(H + 1) * nx},
                       {&pgb_multi_dev::g_Y, &pgb_samples::Y, (size_t)H * ny}};
  for (const Arr& a : arrs) {
    std::vector<const double*> send;
    std::vector<double*> recv;
    for (auto& D : m->dev) { send.push_back(D.S->*(a.s)); recv.push_back(D.*(a.g)); }
    if ((rc = all_gather(m, send, recv, (size_t)cap * a.per))) return rc;
  }
  cudaSetDevice(m->dev[0].device);
  cudaEventRecord(m->dev[0].ev1, m->dev[0].stream);
  m->scen_gather_pending = true;  // the gather completes behind the call (wait_scenario_gather)
  for (auto& D : m->dev) {
    cudaSetDevice(D.device);
    if (cudaStreamSynchronize(D.stream2) != cudaSuccess)
      return merr(m, PGB_ERR_CUDA, "D2H of the scenario arrays failed on device %d", D.device);
  }
  m->have_scen = true;
  return PGB_OK;
}

// discrete_dynamics[!] + Jacobians of the stacked scenario model (optimal_control_Altro.jl:79-107, :51) on the sharded
// resident samples: device d evaluates its own scenarios [g0, g0 + gn) from / into its block of the host arrays
// (x_vec, x_next: [K][n_x]; dfdx: [K][n_x*n_x]; dfdu: [K][n_x*n_u]) — independent scenarios, no collective.  t >= 0 adds
// the process noise v_vec[:, t+1, k] of the last pgb_multi_rollout.  last_ms: the slowest device's kernel.
extern "C" int pgb_multi_dynamics(pgb_multi* m, const double* x_vec, const double* u, int64_t t, int32_t phi_variant,
                                  double* x_next, double* dfdx, double* dfdu, double* kernel_ms_max) {
  NEEDM(m);
  if (!m->have_samples) return merr(m, PGB_ERR_STATE, "no samples are resident (pgb_multi_particle_gibbs / pgb_multi_set_samples first)");
  if (!x_vec || !u || !x_next) return merr(m, PGB_ERR_INVALID, "bad arguments");
  const size_t nx = (size_t)m->cfg.n_x, nu = (size_t)m->cfg.n_u;
  int rc = for_each_device(m, [=](pgb_multi_dev& D) -> int {
    if (!D.S || D.gn == 0) return PGB_OK;
    const size_t o = (size_t)D.g0;
    int r = pgb_dynamics_dev(D.S, x_vec + o * nx, u, t, phi_variant, x_next + o * nx, dfdx ? dfdx + o * nx * nx : nullptr,
                             dfdu ? dfdu + o * nx * nu : nullptr);
    if (r) pgb_set_err(nullptr, r, "%s", pgb_samples_last_error(D.S));
    return r;
  });
  if (rc) return rc;
  if (kernel_ms_max) {
    double mx = 0.0;
    for (auto& D : m->dev)
      if (D.S && D.gn) {
        double v = 0.0;
        pgb_dynamics_last_ms(D.S, &v);
        mx = std::max(mx, v);
      }
    *kernel_ms_max = mx;
  }
  return PGB_OK;
}

// the gathered scenario arrays as device `which_device` of the list holds them (every device holds all) -> host
extern "C" int pgb_multi_get_scenarios(pgb_multi* m, int32_t which_device, double* x0, double* X, double* Y) {
  NEEDM(m);
  if (!m->have_scen) return merr(m, PGB_ERR_STATE, "no rollout has run");
  wait_scenario_gather(m);
  if (which_device < 0 || which_device >= m->n_dev) return merr(m, PGB_ERR_INVALID, "bad device index %d", which_device);
  const int nx = m->cfg.n_x, ny = m->cfg.n_y;
  const int64_t cap = m->cap, H = m->dev[0].gH;
  pgb_multi_dev& W = m->dev[which_device];
  cudaSetDevice(W.device);
  for (int d = 0; d < m->n_dev; ++d) {
    const pgb_multi_dev& D = m->dev[d];
    const size_t n = (size_t)D.gn, o = (size_t)D.g0, g = (size_t)d * cap;
    if (x0 && cudaMemcpy(x0 + o * nx, W.g_x0 + g * nx, n * nx * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return merr(m, PGB_ERR_CUDA, "D2H failed");
    if (X && cudaMemcpy(X + o * (H + 1) * nx, W.g_X + g * (H + 1) * nx, n * (H + 1) * nx * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return merr(m, PGB_ERR_CUDA, "D2H failed");
    if (Y && cudaMemcpy(Y + o * H * ny, W.g_Y + g * H * ny, n * H * ny * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return merr(m, PGB_ERR_CUDA, "D2H failed");
  }
  return PGB_OK;
}

// Scenario-sharded rollout of K host-supplied samples (the configs[4] shape: the samples do not come from chains of this
// handle): sample g goes to device floor(g * n_dev / K) (contiguous blocks), stays resident there, and pgb_multi_rollout
// then behaves as above.  K_total scenarios; layouts of pgb_rollout.
extern "C" int pgb_multi_set_samples(pgb_multi* m, int64_t K_total, const double* A, const double* Q, const double* w_m1,
                                     const double* x_m1, const double* u_m1, uint64_t roll_seed) {
  NEEDM(m);
  wait_scenario_gather(m);
  if (K_total < m->n_dev || !A || !Q || !w_m1 || !x_m1 || !u_m1) return merr(m, PGB_ERR_INVALID, "bad arguments");
  if ((uint64_t)K_total > (1ull << 24)) return merr(m, PGB_ERR_INVALID, "K exceeds the 2^24 Philox stream ids");
  const int nx = m->cfg.n_x, np = m->cfg.n_phi, nu = m->cfg.n_u;
  const int64_t N = m->cfg.N;
  int64_t cap = 0;
  for (int d = 0; d < m->n_dev; ++d) {
    pgb_multi_dev& D = m->dev[d];
    const int64_t lo = K_total * d / m->n_dev, hi = K_total * (d + 1) / m->n_dev;
    D.g0 = lo;
    D.gn = hi - lo;
    cap = std::max<int64_t>(cap, hi - lo);
  }
  for (auto& D : m->dev) {
    cudaSetDevice(D.device);
    if (D.S) pgb_samples_destroy(D.S);
    D.S = nullptr;
    for (double** p : {&D.g_rec, &D.rec, &D.g_x0, &D.g_X, &D.g_Y}) {
      if (*p) cudaFree(*p);
      *p = nullptr;
    }
    D.gH = 0;
  }
  m->cap = cap;
  m->K = 0;  // not chain samples
  m->roll_seed = roll_seed;
  m->have_scen = false;
  int rc = for_each_device(m, [=](pgb_multi_dev& D) -> int {
    pgb_config c = m->cfg;
    c.device = D.device;
    c.n_chains = 1;
    int r = pgb_samples_create(&D.S, &c, cap);
    if (r) return r;
    D.S->K_active = D.gn;
    const size_t o = (size_t)D.g0;
    r = pgb_samples_set(D.S, 0, D.gn, A + o * nx * np, Q + o * nx * nx, w_m1 + o * N, x_m1 + o * nx * N, u_m1 + o * nu);
    if (r) return r;
    return pgb_samples_draw_xT(D.S, roll_seed, (uint32_t)D.g0, nullptr);
  });
  if (rc) return rc;
  m->have_samples = true;
  return PGB_OK;
}

// device-side time (CUDA events, ms): NCCL gather of the packed samples, slowest device's rollout kernel, NCCL gather
// of the scenario arrays
extern "C" int pgb_multi_last_ms(pgb_multi* m, double* sample_gather_ms, double* rollout_kernel_ms_max, double* scenario_gather_ms) {
  NEEDM(m);
  wait_scenario_gather(m);
  if (sample_gather_ms) *sample_gather_ms = m->gather_ms;
  if (rollout_kernel_ms_max) *rollout_kernel_ms_max = m->rollout_kernel_ms_max;
  if (scenario_gather_ms) *scenario_gather_ms = m->rollout_gather_ms;
  return PGB_OK;
}

// per-device engine access for the callers that want the reference's PG_sample records of a chain (w_m1, x_m1 stay where
// the chain ran and are NOT gathered: 1 MB per sample at N = 2^16): sample k of global chain c -> host
extern "C" int pgb_multi_get_sample(pgb_multi* m, int32_t chain, int64_t k, double* A, double* Q, double* w_m1, double* x_m1,
                                    double* u_m1) {
  NEEDM(m);
  if (!m->have_samples || m->K < 1) return merr(m, PGB_ERR_STATE, "no chain samples (pgb_multi_particle_gibbs has not run)");
  for (auto& D : m->dev)
    if (chain >= D.c0 && chain < D.c0 + D.nc) {
      if (k < 0 || k >= m->K) return merr(m, PGB_ERR_INVALID, "bad sample index");
      return pgb_samples_get(D.S, (int64_t)(chain - D.c0) * m->K + k, 1, A, Q, w_m1, x_m1, u_m1);
    }
  return merr(m, PGB_ERR_INVALID, "bad chain %d", chain);
}
