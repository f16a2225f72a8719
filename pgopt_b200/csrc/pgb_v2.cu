// pgb_v2.cu — launch sequence of the chunk-stationary sweep kernel (pgb_v2.cuh): one cooperative launch per sweep,
// one 896-thread CTA per SM, the chain's particles dealt to the CTAs in contiguous tile-aligned ranges.
#include "pgb_internal.h"
#include "pgb_v2.cuh"

using namespace pgb;

int pgb_v2_alloc(pgb_handle* h) {
  V2Dev& vd = h->v2;
  if (vd.bar) return PGB_OK;
  const int nc = h->fd.nc;
  const int64_t nt = h->fd.nt;
  CK(h, dalloc(h, &vd.bar, 4));
  CK(h, dalloc(h, &vd.pe, (size_t)nc * nt));
  CK(h, dalloc(h, &vd.pw, (size_t)nc * nt));
  CK(h, dalloc(h, &vd.pa, (size_t)nc * nt));
  CK(h, dalloc(h, &vd.pa2, (size_t)nc * nt));
  CK(h, dalloc(h, &vd.cta_agg, (size_t)2 * nc * V2_MAXNPC));
  CK(h, dalloc(h, &vd.recs, (size_t)2 * nc * V2_MAXNPC * V2_RCTA));
  CK(h, dalloc(h, &vd.cand, (size_t)2 * nc));
  CK(h, dalloc(h, &vd.Wn, (size_t)nc * nt * TILE));
  CK(h, dalloc(h, &vd.qin_fb, (size_t)nc * nt * TPB));
  vd.star_idx = h->d_star_idx;
#ifdef PGB_EXPERIMENT
  CK(h, dalloc(h, &vd.prof, (size_t)1024 * 16));
#endif
  return PGB_OK;
}

#ifdef PGB_EXPERIMENT
// experiment builds only (never in the shipped library): average cycles thread 0 spent per section of the last sweep
extern "C" int pgb_experiment_phase_cycles(pgb_handle* h, double* avg /*[16]*/) {
  NEED(h);
  CK(h, cudaStreamSynchronize(h->stream));
  const int g = h->fd.nc * h->v2.npc;
  std::vector<long long> v((size_t)1024 * 16);
  CK(h, cudaMemcpy(v.data(), h->v2.prof, v.size() * 8, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 16; ++i) {
    double s = 0;
    for (int c = 0; c < g && c < 1024; ++c) s += (double)v[(size_t)c * 16 + i];
    avg[i] = g > 0 ? s / g : 0.0;
  }
  return PGB_OK;
}
#endif

template <int NX, int BASIS, bool FIRST, int CA, typename RT>
static int launch_v2(pgb_handle* h, size_t smem) {
  const void* kfn = (const void*)k_sweep_v2<NX, BASIS, FIRST, CA, RT>;
  CK(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_v2<NX, BASIS, FIRST, CA, RT>, V2_NT, smem));
  if (occ < 1) return -1;
  LaunchScope ls(h, PGB_PROF_EXPAND);
  void* args[] = {(void*)&h->fd, (void*)&h->v2};
  CK(h, cudaLaunchCooperativeKernel(kfn, dim3((unsigned)(h->fd.nc * h->v2.npc)), dim3(V2_NT), args, smem, h->stream));
  return PGB_OK;
}

template <int NX, int BASIS, int CA = 0>
static int launch_v2_first(pgb_handle* h, size_t smem, int first) {
#ifdef PGB_V2_ONLY_MAIN  // experiment builds: compile the headline kernels only (known / 5x5x5 Hilbert basis)
  if constexpr (BASIS == 2) {
    return -1;
  } else
#endif
  {
    if (h->fd.f32)
      return first ? launch_v2<NX, BASIS, true, CA, float>(h, smem) : launch_v2<NX, BASIS, false, CA, float>(h, smem);
    return first ? launch_v2<NX, BASIS, true, CA, double>(h, smem) : launch_v2<NX, BASIS, false, CA, double>(h, smem);
  }
}

int pgb_run_filter_v2(pgb_handle* h) {
  FilterDev& fd = h->fd;
  V2Dev& vd = h->v2;
  cudaStream_t s = h->stream;
  const int nc = fd.nc, nt = fd.nt;
  const int64_t T = fd.T;
  if (nc > h->num_sms) return -1;
  int npc = h->num_sms / nc;
  if (npc > nt) npc = nt;
  if (npc > V2_MAXNPC) npc = V2_MAXNPC;
  if (npc < 1) return -1;
  int rc = pgb_v2_alloc(h);
  if (rc) return rc;
  vd.npc = npc;
  vd.tiles_lo = nt / npc;
  vd.n_hi = nt % npc;
  vd.mch = V2_CPT * (vd.tiles_lo + (vd.n_hi > 0 ? 1 : 0));
  vd.force_fallback = g_pgb_force_fallback;
  const int n_A = fd.n_x * fd.n_phi;
  vd.stationary = 0;
  if (vd.mch <= V2_STAT_MAXCH && v2_smem_layout(vd.mch, n_A, 1).total <= 227 * 1024) vd.stationary = 1;
  const size_t smem = v2_smem_layout(vd.mch, n_A, vd.stationary).total;
  if (smem > 227 * 1024) return -1;
  if (!vd.stationary && !vd.gE) {
    CK(h, dalloc(h, &vd.gE, (size_t)nc * nt * TILE));
    CK(h, dalloc(h, &vd.gP, (size_t)nc * nt * TILE));
    CK(h, dalloc(h, &vd.gTE, (size_t)nc * nt * TPB));
  }
  int first = 0;
  rc = pgb_filter_prologue(h, &first);
  if (rc) return rc;
  CK(h, cudaMemsetAsync(vd.bar, 0, 4, s));
  CK(h, cudaMemsetAsync(vd.cand, 0, (size_t)2 * nc * 4, s));
  CK(h, cudaMemsetAsync(fd.main.flag, 0, (size_t)2 * nc * 4, s));
  CK(h, cudaMemsetAsync(fd.side.flag, 0, (size_t)2 * nc * 4, s));
  LAUNCH(h, PGB_PROF_INIT, (k_build_utabs<<<dim3((unsigned)((T + 127) / 128), nc), 128, 0, s>>>(fd, first)));
  const bool h355 = fd.basis == PGB_BASIS_HILBERT_GP && fd.n_x == 2 && fd.n_z == 3 && fd.hg_count[0] == 5 &&
                    fd.hg_count[1] == 5 && fd.hg_count[2] == 5;
  if (fd.basis == PGB_BASIS_KNOWN_VB) rc = launch_v2_first<2, 0>(h, smem, first);
  else if (h355 && nc == 1) {  // single chain: A as constant-bank operands of the contraction
    if (fd.f32) {
      float* stage = reinterpret_cast<float*>(vd.Wn);  // scratch of the fix-up path, idle between sweeps
      k_round_to_f32<<<(n_A + 255) / 256, 256, 0, s>>>(fd.A, stage, n_A);
      CK(h, cudaMemcpyToSymbolAsync(c_v2Af, stage, (size_t)n_A * 4, 0, cudaMemcpyDeviceToDevice, s));
    } else {
      CK(h, cudaMemcpyToSymbolAsync(c_v2A, fd.A, (size_t)n_A * 8, 0, cudaMemcpyDeviceToDevice, s));
    }
    rc = launch_v2_first<2, 1, 1>(h, smem, first);
  } else if (h355) rc = launch_v2_first<2, 1, 0>(h, smem, first);
  else {
    switch (fd.n_x) {
      case 1: rc = launch_v2_first<1, 2>(h, smem, first); break;
      case 2: rc = launch_v2_first<2, 2>(h, smem, first); break;
      case 3: rc = launch_v2_first<3, 2>(h, smem, first); break;
      default: rc = launch_v2_first<4, 2>(h, smem, first); break;
    }
  }
  if (rc) return rc;
  CK(h, cudaGetLastError());
  return pgb_launch_trace_and_stats(h, first, false);
}
