// pgb_smalln.cu — launch sequence of the warp-per-chain / CTA-per-chain sweep kernels (pgb_small.cuh).
#include "pgb_internal.h"
#include "pgb_small.cuh"

using namespace pgb;

// Small particle counts with a stored history: one warp per chain (N <= 32) or one CTA per chain (N <= 1024), the
// whole time loop in one ordinary launch (pgb_small.cuh).  Returns -1 when the configuration does not qualify.
template <int NX>
static int run_filter_small_nx(pgb_handle* h, int64_t n_max) {
  FilterDev& fd = h->fd;
  cudaStream_t s = h->stream;
  const int64_t T = fd.T;
  const int nc = fd.nc;
  if (fd.N > n_max || fd.N > CTA_MAXN || fd.x_rows != T) return -1;
  const bool cta = fd.N > SMALL_MAXN;  // mid-size N: one CTA per chain (k_sweep_cta), thread = particle
  const int np_thr = (int)((fd.N + 31) / 32) * 32;
  const size_t smem = cta ? ((size_t)((NX * fd.n_phi + 1) & ~1) + (size_t)(NX + 4) * np_thr + 4) * 8 + sizeof(CtaShared)
                          : ((size_t)((NX * fd.n_phi + 1) & ~1) + (size_t)NX * 32 + 64) * 8;
  if (smem > 200 * 1024) return -1;
  int first = 0;
  int rc = pgb_filter_prologue(h, &first);
  if (rc) return rc;
  if (cta) {
    CK(h, cudaFuncSetAttribute(k_sweep_cta<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(h, PGB_PROF_EXPAND, (k_sweep_cta<NX><<<nc, np_thr, smem, s>>>(fd, h->pd.star_idx, first)));
    LAUNCH(h, PGB_PROF_TAIL, (k_trace<NX><<<(nc + 31) / 32, 32, 0, s>>>(fd, h->d_star_idx, h->d_star_out)));
  } else {
    // Philox mode: the noise of the whole sweep is drawn ahead of the dependent time loop (bounded: 512 MB)
    const size_t zbytes = (size_t)nc * T * fd.N * NX * 8;
    const bool predraw = fd.philox && zbytes <= ((size_t)512 << 20);
    if (predraw) {
      if (!h->d_Zpre) {
        CK(h, dalloc(h, &h->d_Zpre, (size_t)nc * T * fd.N * NX, false));
        CK(h, dalloc(h, &h->d_Upre, (size_t)nc * T, false));
        CK(h, dalloc(h, &h->d_Apre, (size_t)nc * T, false));
      }
      LAUNCH(h, PGB_PROF_INIT, (k_predraw_small<NX><<<dim3((unsigned)((T + 3) / 4), nc), 128, 0, s>>>(fd, h->d_Zpre, h->d_Upre,
                                                                                                  h->d_Apre, first)));
    }
    CK(h, cudaFuncSetAttribute(k_sweep_small<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(h, PGB_PROF_EXPAND, (k_sweep_small<NX><<<nc, 32, smem, s>>>(fd, h->pd.star_idx, first, predraw ? h->d_Zpre : nullptr,
                                                                         predraw ? h->d_Upre : nullptr,
                                                                         predraw ? h->d_Apre : nullptr)));
    LAUNCH(h, PGB_PROF_TAIL, (k_trace_small<NX><<<nc, 32, 0, s>>>(fd, h->d_star_idx, h->d_star_out)));
  }
  CK(h, cudaGetLastError());
  return pgb_launch_trace_and_stats(h, first, true);
}
int pgb_run_filter_small(pgb_handle* h, int64_t n_max) {
  switch (h->cfg.n_x) {
    case 1: return run_filter_small_nx<1>(h, n_max);
    case 2: return run_filter_small_nx<2>(h, n_max);
    case 3: return run_filter_small_nx<3>(h, n_max);
    default: return run_filter_small_nx<4>(h, n_max);
  }
}
