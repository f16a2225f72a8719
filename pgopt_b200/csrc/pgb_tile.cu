// pgb_tile.cu — launch sequences of the 1024-particle tile kernels: the persistent cooperative sweep kernel
// (pgb_persist.cuh), the one-launch-per-phase path (pgb_filter.cuh / pgb_scan.cuh) and the stand-alone
// systematic_resampling entry point.
#include "pgb_internal.h"

using namespace pgb;

cudaError_t pgb_alloc_scanws(pgb_handle* h, ScanWS* ws, int64_t N, int nc, long long** counters) {
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

// Launch sequence of one exact scan + consumer on stream s.
template <int NX, int MODE>
static void launch_scan(pgb_handle* h, const ScanWS& ws, const double* v, const pgb_utable* utab, int64_t utab_stride,
                        int64_t utab_index, int32_t* out_idx, int64_t out_stride, int64_t t, int first,
                        unsigned long long* maxkey_reset) {
  const FilterDev& fd = h->fd;
  cudaStream_t s = h->stream;
  dim3 gt(ws.nt, fd.nc), g1(1, fd.nc);
  LAUNCH(h, PGB_PROF_SCAN_LOCAL, (k_scan_local<<<gt, TPB, 0, s>>>(ws, v, fd.status)));
  LAUNCH(h, PGB_PROF_RESOLVE, (k_resolve<<<g1, TPB, 0, s>>>(ws, g_pgb_force_fallback & 1)));
  LAUNCH(h, PGB_PROF_EXPAND, (k_expand<NX, MODE><<<gt, TPB, 0, s>>>(fd, ws, v, utab, utab_stride, utab_index, out_idx,
                                                                    out_stride, t, first, 0)));
  LAUNCH(h, PGB_PROF_SEQ, (k_seq_exact<<<g1, 32, 0, s>>>(ws, v, fd.status, maxkey_reset)));
  LAUNCH(h, PGB_PROF_EXPAND_FB, (k_expand<NX, MODE><<<gt, TPB, 0, s>>>(fd, ws, v, utab, utab_stride, utab_index, out_idx,
                                                                       out_stride, t, first, 1)));
}

template <int NX>
static int run_filter_nx(pgb_handle* h) {
  FilterDev& fd = h->fd;
  cudaStream_t s = h->stream;
  const int64_t N = fd.N, T = fd.T;
  const int nc = fd.nc;
  int first = 0;
  int rc = pgb_filter_prologue(h, &first);
  if (rc) return rc;
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
  CK(h, cudaGetLastError());
  return pgb_launch_trace_and_stats(h, first, false);
}
int pgb_run_filter_tile_multikernel(pgb_handle* h) {
  switch (h->cfg.n_x) {
    case 1: return run_filter_nx<1>(h);
    case 2: return run_filter_nx<2>(h);
    case 3: return run_filter_nx<3>(h);
    default: return run_filter_nx<4>(h);
  }
}

// One cooperative launch for the whole time loop.  Returns -1 when the configuration does not fit
// (too many chains / tiles for the co-resident grid); the caller then uses the multi-kernel path.
template <int NX>
static int run_filter_persistent_nx(pgb_handle* h) {
  FilterDev& fd = h->fd;
  PersistDev& pd = h->pd;
  cudaStream_t s = h->stream;
  const int64_t T = fd.T;
  const int nc = fd.nc;
  const size_t smA = sizeof(PersistShared) + (size_t)NX * fd.n_phi * 8;
  const void* kfn = (const void*)k_sweep_persistent<NX>;
  int occ = 0;
  CK(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA));
  CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_persistent<NX>, TPB, smA));
  const int max_ctas = occ * h->num_sms;
  if (max_ctas > 4096) return -1;
  int nbmax = max_ctas / nc;
  if (nbmax < 1) return -1;
  int tpc = 1;
  // finest blocks the per-chain tables allow, dealt over every co-resident CTA slot
  while ((fd.nt + tpc - 1) / tpc > P_MAXNB) tpc *= 2;
  pd.tpc = tpc;
  pd.nb = (fd.nt + tpc - 1) / tpc;
  pd.npc = pd.nb < nbmax ? pd.nb : nbmax;
  if (((pd.nb + pd.npc - 1) / pd.npc) * tpc > P_MAXTPC) return -1;
  pd.force_fallback = g_pgb_force_fallback;
  int first = 0;
  int rc = pgb_filter_prologue(h, &first);
  if (rc) return rc;
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
  CK(h, cudaGetLastError());
  return pgb_launch_trace_and_stats(h, first, false);
}
int pgb_run_filter_tile_persistent(pgb_handle* h) {
  switch (h->cfg.n_x) {
    case 1: return run_filter_persistent_nx<1>(h);
    case 2: return run_filter_persistent_nx<2>(h);
    case 3: return run_filter_persistent_nx<3>(h);
    default: return run_filter_persistent_nx<4>(h);
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone helpers
// ------------------------------------------------------------------------------------------------
static float g_resampling_ms = 0.f;  // device time of the last pgb_systematic_resampling (sum + exact scan + search)
__global__ void k_build_one_utab(pgb_utable* ut, int64_t M, double rnd) { pgb_utable_build(ut, M, rnd); }

extern "C" int pgb_systematic_resampling(int32_t device, const double* W, int64_t n_w, int64_t n_draw, double rnd,
                                         int64_t* idx, int32_t* status) {
  if (!W || !idx || n_w < 1 || n_draw < 1) return set_err(nullptr, PGB_ERR_INVALID, "bad arguments");
  if (n_w >= (1ll << 31) - TILE) return set_err(nullptr, PGB_ERR_INVALID, "n_w too large");
  if (int rc = pgb_require_device()) return rc;
  TmpHandle tmp;  // only used as an allocation list / error sink
  pgb_handle* h = &tmp;
  h->cfg.device = device;
  CK(h, cudaSetDevice(device));
  FilterDev fd;
  memset(&fd, 0, sizeof(fd));
  fd.nc = 1; fd.N = n_w; fd.T = 1; fd.n_x = 1; fd.n_y = 1; fd.n_u = 1;
  ScanWS ws;
  CK(h, pgb_alloc_scanws(h, &ws, n_w, 1, nullptr));
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
  cudaEvent_t ev[2];
  CK(h, cudaEventCreate(&ev[0]));
  CK(h, cudaEventCreate(&ev[1]));
  CK(h, cudaEventRecord(ev[0], s));
  k_build_one_utab<<<1, 1, 0, s>>>(d_ut, n_draw, rnd);
  k_tile_sums<<<dim3(ws.nt, 1), TPB, 0, s>>>(ws, d_W);
  launch_scan<1, MODE_INDEX>(h, ws, d_W, d_ut, 1, 0, d_idx, n_draw, 0, 0, nullptr);
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(ev[1], s));
  CK(h, cudaDeviceSynchronize());
  cudaEventElapsedTime(&g_resampling_ms, ev[0], ev[1]);
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
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

extern "C" int pgb_resampling_last_ms(double* kernel_ms) {
  if (kernel_ms) *kernel_ms = g_resampling_ms;
  return PGB_OK;
}
