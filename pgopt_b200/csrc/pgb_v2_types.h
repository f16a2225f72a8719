// pgb_v2_types.h — constants and the device-side parameter block of the chunk-stationary sweep kernel (pgb_v2.cuh);
// split out so that only pgb_v2.cu compiles the kernel itself.
#pragma once
#include "pgb_scan.cuh"

namespace pgb {

constexpr int V2_NW = 28;              // warps per CTA
constexpr int V2_NT = V2_NW * 32;      // 896 threads, one CTA per SM
constexpr int V2_CHUNK = 128;          // particles per warp work unit (aligned subtree of the summation tree)
constexpr int V2_CPT = TILE / V2_CHUNK;  // 8 chunks per 1024-particle tile
constexpr int V2_MAXNPC = 160;         // CTAs per chain (<= SMs)
constexpr int V2_MAXREC = 160;         // complex lanes resolved per scan (more -> serial fallback)
constexpr int V2_RCTA = 128;           // complex-lane records kept per CTA and scan (a shared-memory pool; more -> fallback)
constexpr int V2_CAP = 8192;           // output slots staged per piece
constexpr int V2_STAT_MAXCH = 56;      // chunks per CTA up to which the per-CTA arrays fit shared memory

struct V2Dev {
  int npc;          // CTAs per chain
  int tiles_lo;     // CTA rank r owns tiles_lo + (r < n_hi) consecutive tiles
  int n_hi;
  int mch;          // max chunks per CTA (= 8 * (tiles_lo + (n_hi > 0)))
  int stationary;   // per-CTA arrays (E, P, TE) in shared memory (1) or in the global buffers below (0)
  int force_fallback;
  unsigned int* bar;
  double *pe, *pw, *pa, *pa2;  // [nc][nt] tile sums
  pgb_agg* cta_agg;   // [2][nc][npc]  (main, side)
  ScanRec* recs;      // [2][nc][npc][V2_RCTA]
  int* cand;          // [nc][2] ancestor draw: {number of certain hits, index}
  int32_t* star_idx;  // [nc]
  double *gE, *gP;    // [nc][nt*TILE]   non-stationary mode
  pgb_agg* gTE;       // [nc][nt*TPB]
  double* Wn;         // [nc][nt*TILE]   normalised weights dumped for the serial fallback walk
  double* qin_fb;     // [nc][nt*TPB]    exact incoming value per lane written by the fallback walk
  long long* prof;    // -DPGB_EXPERIMENT builds only: [grid][16] cycles per section (thread 0 of every CTA)
};


}  // namespace pgb
