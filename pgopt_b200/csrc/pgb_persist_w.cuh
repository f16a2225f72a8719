// pgb_persist_w.cuh — persistent CPF-AS kernel, warp-tile formulation (default path).
//
// Same phases and barriers as pgb_persist.cuh (A prep | B scan-local | C resolve + expansion | D weights), but the
// unit of work inside a CTA is a WARP-TILE: 128 consecutive particles handled by one warp (lane = 4 consecutive
// particles).  Warp w of CTA c owns the contiguous, aligned block of tpc warp-tiles
// [(c*8 + w)*tpc, (c*8 + w + 1)*tpc).  Scans, seam checks, tree sums and the by-output expansion are done with
// shuffles and __syncwarp only; the eight warps of a CTA meet at two or three __syncthreads per PHASE (CTA-level
// combination of the warp-tile aggregates) instead of ~16 per 1024-element tile, so warps drift apart and hide
// each other's memory / FP64 / instruction-fetch latency.  All arithmetic and every summation order is unchanged
// (balanced tree: lane 4 -> warp 128 -> warp's tpc tiles -> CTA's 8 warps -> grid), hence bit-identical results.
#pragma once
#include "pgb_persist.cuh"

namespace pgb {

constexpr int W_MAXNB = 640;               // max CTAs (= blocks) per chain
constexpr int W_MAXTPC = 16;              // warp-tiles per warp
constexpr int W_MAXWT = 8 * W_MAXTPC;     // warp-tiles per CTA
constexpr int W_STAGE = 64;               // complex threads staged per CTA and scan
constexpr int W_CAP = 256;                // output slots staged per warp and chunk
constexpr int WT = 128;                   // particles per warp-tile

struct WShared {
  double sm[NWARP + 1];
  pgb_agg smagg[NWARP + 1];
  int sI, any_flag, stage_cnt, pad_;
  double wpart[NWARP][2];
  double wtsum_w[W_MAXWT];
  double wtsum_a2[W_MAXWT];
  double wtpfx[W_MAXWT + 1];
  pgb_agg wtagg[W_MAXWT];
  pgb_agg wtex[W_MAXWT + 1];
  pgb_agg cex[W_MAXNB + 1];
  double pfx[W_MAXNB + 1];
  pgb_map rf[P_MAXREC];
  double rW[P_MAXREC][4];
  double seeds[P_MAXREC + 1];
  ScanRec stage[W_STAGE];
  int stage_wt[W_STAGE];
  int s_anc[NWARP][W_CAP];
  pgb_utable ut;
};

__device__ __forceinline__ double tree8(const double (*p)[2], int j) {
  return __dadd_rn(__dadd_rn(__dadd_rn(p[0][j], p[1][j]), __dadd_rn(p[2][j], p[3][j])),
                   __dadd_rn(__dadd_rn(p[4][j], p[5][j]), __dadd_rn(p[6][j], p[7][j])));
}

// exclusive approximate scan of p[0..nb) into pfx[0..nb] (identical in every CTA)
PGB_DEV_HELPER void w_approx_prefix(const double* p, int nb, double* pfx, double* sm) {
  double carry = 0.0;
  for (int base = 0; base < nb; base += TPB) {
    int i = base + threadIdx.x;
    double v = (i < nb) ? p[i] : 0.0;
    double tot;
    double ex = block_excl_scan_approx(v, sm, &tot);
    if (i < nb) pfx[i] = __dadd_rn(carry, ex);
    carry = __dadd_rn(carry, tot);
  }
  if (threadIdx.x == 0) pfx[nb] = carry;
  __syncthreads();
}

// warp 0: wtpfx[i] = start + sum_{j<i} wtsum[j] (sequential order per lane block, approximate use only),
// wtpfx[nwt] = end (the next CTA's start, so that both sides of the CTA seam see the same number)
__device__ __forceinline__ void w_wt_prefix(const double* wtsum, int nwt, double start, double end, double* wtpfx) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) return;
  const int per = (nwt + 31) / 32;
  double loc[4], s = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int i = lane * per + j;
    loc[j] = s;
    if (j < per && i < nwt) s = __dadd_rn(s, wtsum[i]);
  }
  double inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = __dadd_rn(t, inc);
  }
  double ex = __dadd_rn(start, __dsub_rn(inc, s));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int i = lane * per + j;
    if (j < per && i < nwt) wtpfx[i] = __dadd_rn(ex, loc[j]);
  }
  if (lane == 0) wtpfx[nwt] = end;
}

// warp 0: exclusive segmented composition scan of the CTA's warp-tile aggregates; returns the CTA total (lane-uniform)
__device__ __forceinline__ pgb_agg w_wt_agg_scan(const pgb_agg* wtagg, int nwt, pgb_agg* wtex) {
  const int lane = threadIdx.x & 31;
  const int per = (nwt + 31) / 32;
  pgb_agg loc[4], s = pgb_agg_identity();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int i = lane * per + j;
    loc[j] = s;
    if (j < per && i < nwt) s = pgb_agg_combine(s, wtagg[i]);
  }
  pgb_agg inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    pgb_agg t = shfl_up_agg(inc, o);
    if (lane >= o) inc = pgb_agg_combine(t, inc);
  }
  pgb_agg ex = shfl_up_agg(inc, 1);
  if (lane == 0) ex = pgb_agg_identity();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int i = lane * per + j;
    if (j < per && i < nwt) wtex[i] = pgb_agg_combine(ex, loc[j]);
  }
  pgb_agg tot;
  tot.f.c0 = __shfl_sync(0xffffffffu, (long long)inc.f.c0, 31);
  tot.f.c1 = __shfl_sync(0xffffffffu, (long long)inc.f.c1, 31);
  tot.cnt = __shfl_sync(0xffffffffu, inc.cnt, 31);
  tot.key = __shfl_sync(0xffffffffu, inc.key, 31);
  if (lane == 0) wtex[nwt] = tot;
  return tot;
}

PGB_DEV_HELPER pgb_agg warp_excl_scan_agg(pgb_agg v, pgb_agg* total) {
  const int lane = threadIdx.x & 31;
  pgb_agg inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    pgb_agg t = shfl_up_agg(inc, o);
    if (lane >= o) inc = pgb_agg_combine(t, inc);
  }
  pgb_agg ex = shfl_up_agg(inc, 1);
  if (lane == 0) ex = pgb_agg_identity();
  total->f.c0 = __shfl_sync(0xffffffffu, (long long)inc.f.c0, 31);
  total->f.c1 = __shfl_sync(0xffffffffu, (long long)inc.f.c1, 31);
  total->cnt = __shfl_sync(0xffffffffu, inc.cnt, 31);
  total->key = __shfl_sync(0xffffffffu, inc.key, 31);
  return ex;
}

// ---- scan-local for the warp's tiles: W = v/S, prediction, maps, warp-level composition; CTA-level combination.
// Contains __syncthreads: must be called by all threads of the CTA.
PGB_DEV_HELPER void w_scan_local(const ScanWS& ws, const double* v, double S, const double* wtsum,
                                             const double* pfx, int chain, int c, int tpc, WShared& sh, int* status,
                                             int* flagp) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int nwt = 8 * tpc;
  const int64_t N = ws.N;
  const double* vc = v + (int64_t)chain * N;
  const InvDiv ivS = make_invdiv(S);
  w_wt_prefix(wtsum, nwt, pfx[c], pfx[c + 1], sh.wtpfx);
  __syncthreads();
  for (int j = 0; j < tpc; ++j) {
    const int idx = w * tpc + j;
    const int64_t g = (int64_t)c * nwt + idx;  // global warp-tile id
    const int64_t base = g * WT + lane * IPT;
    double W[4];
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      W[i] = (base + i < N) ? div_inv(vc[base + i], ivS) : 0.0;
      if (!(W[i] >= 0.0 && W[i] <= 1.7976931348623157e308)) bad = true;
    }
    if (bad) atomicOr(&status[chain], 4);
    if (base < (int64_t)ws.nt * TILE) {
      double2* wn = reinterpret_cast<double2*>(ws.Wn + (int64_t)chain * ws.nt * TILE + base);
      wn[0] = make_double2(W[0], W[1]);
      wn[1] = make_double2(W[2], W[3]);
    }
    double s1 = W[0], s2 = __dadd_rn(s1, W[1]), s3 = __dadd_rn(s2, W[2]), s4 = __dadd_rn(s3, W[3]);
    double inc = s4;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc = __dadd_rn(t, inc);
    }
    // approximate quantities only: a multiplication by the reciprocal is enough (and identical on both sides of a seam)
    const double P = __dadd_rn(__dmul_rn(sh.wtpfx[idx], ivS.y), __dsub_rn(inc, s4));
    double Pn = __shfl_down_sync(0xffffffffu, P, 1);
    if (lane == 31) Pn = __dmul_rn(sh.wtpfx[idx + 1], ivS.y);
    double qt[5] = {P, __dadd_rn(P, s1), __dadd_rn(P, s2), __dadd_rn(P, s3), Pn};
    pgb_agg agg = bad ? pgb_agg_complex() : pgb_classify4(W, qt);
    pgb_agg total;
    pgb_agg exagg = warp_excl_scan_agg(agg, &total);
    if (base < (int64_t)ws.nt * TILE) ws.thr_ex[(int64_t)chain * ws.nt * TPB + base / IPT] = exagg;  // warp-tile-local
    if (lane == 0) sh.wtagg[idx] = total;
    if (agg.cnt) {
      int slot = atomicAdd(&sh.stage_cnt, 1);
      if (slot < W_STAGE) {
        ScanRec r;
        r.ex = exagg;
        r.W[0] = W[0]; r.W[1] = W[1]; r.W[2] = W[2]; r.W[3] = W[3];
        r.pad_ = 0.0;
        sh.stage[slot] = r;
        sh.stage_wt[slot] = idx;
      } else {
        atomicOr(flagp, 1);
      }
    }
  }
  __syncthreads();
  if (w == 0) {
    pgb_agg tot = w_wt_agg_scan(sh.wtagg, nwt, sh.wtex);
    if (lane == 0) ws.tile_agg[(int64_t)chain * ws.nt + c] = tot;  // CTA aggregate
  }
  __syncthreads();
  {
    int cnt = sh.stage_cnt;
    if (cnt > W_STAGE) cnt = W_STAGE;
    if (tid < cnt) {
      ScanRec r = sh.stage[tid];
      r.ex = pgb_agg_combine(sh.wtex[sh.stage_wt[tid]], r.ex);  // CTA-local exclusive aggregate
      if (r.ex.cnt < tpc * RMAX) ws.recs[((int64_t)chain * ws.nt + (int64_t)c * tpc) * RMAX + r.ex.cnt] = r;
      else atomicOr(flagp, 1);
    }
  }
  __syncthreads();
  if (tid == 0) sh.stage_cnt = 0;
}

// Resolve (every CTA, redundantly) — block-wide, same as p_resolve but on WShared.
PGB_DEV_HELPER void w_resolve(const ScanWS& ws, int chain, int nb, int tpc, WShared& sh, int force_fallback,
                                          int* flagp) {
  const int tid = threadIdx.x;
  const pgb_agg* cagg = ws.tile_agg + (int64_t)chain * ws.nt;
  pgb_agg carry = pgb_agg_identity();
  for (int base = 0; base < nb; base += TPB) {
    int i = base + tid;
    pgb_agg v = (i < nb) ? cagg[i] : pgb_agg_identity();
    pgb_agg tot;
    pgb_agg ex = block_excl_scan_agg(v, sh.smagg, &tot);
    if (i < nb) sh.cex[i] = pgb_agg_combine(carry, ex);
    carry = pgb_agg_combine(carry, tot);
  }
  if (tid == 0) sh.cex[nb] = carry;
  __syncthreads();
  const int C = carry.cnt;
  if (C > P_MAXREC || force_fallback) {
    if (tid == 0) {
      atomicOr(flagp, 1);
      sh.seeds[0] = 0.0;
    }
    __syncthreads();
    return;
  }
  for (int b = tid; b < nb; b += TPB) {
    int cnt = cagg[b].cnt;
    if (cnt > tpc * RMAX) cnt = tpc * RMAX;
    pgb_agg te = sh.cex[b];
    for (int i = 0; i < cnt; ++i) {
      const ScanRec* r = &ws.recs[((int64_t)chain * ws.nt + (int64_t)b * tpc) * RMAX + i];
      int g = te.cnt + i;
      if (g < P_MAXREC) {
        pgb_agg e = pgb_agg_combine(te, r->ex);
        sh.rf[g] = e.f;
        sh.rW[g][0] = r->W[0]; sh.rW[g][1] = r->W[1]; sh.rW[g][2] = r->W[2]; sh.rW[g][3] = r->W[3];
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    double q = 0.0;
    sh.seeds[0] = q;
    for (int g = 0; g < C; ++g) {
      double qb = pgb_from_mant(pgb_segkey(q), pgb_map_apply(sh.rf[g], pgb_mant(q)));
      qb = __dadd_rn(qb, sh.rW[g][0]);
      qb = __dadd_rn(qb, sh.rW[g][1]);
      qb = __dadd_rn(qb, sh.rW[g][2]);
      qb = __dadd_rn(qb, sh.rW[g][3]);
      q = qb;
      sh.seeds[g + 1] = q;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ double w_qin(const WShared& sh, pgb_agg ex) {
  int cc = ex.cnt < 0 ? 0 : (ex.cnt > P_MAXREC ? P_MAXREC : ex.cnt);
  return pgb_qin_from(ex, sh.seeds[cc]);
}

// exact prefixes of one warp-tile (warp-level, no CTA sync).  Returns false (warp-uniform) if the seam check failed.
PGB_DEV_HELPER bool w_exact_prefixes(const ScanWS& ws, int chain, int c, int idx, int nb, int tpc,
                                                 bool fallback_pass, const WShared& sh, double W[4], double q[5],
                                                 int* flagp) {
  const int lane = threadIdx.x & 31;
  const int nwt = 8 * tpc;
  const int64_t g = (int64_t)c * nwt + idx;
  const int64_t base = g * WT + lane * IPT;
  const int64_t cap = (int64_t)ws.nt * TILE;
  if (base < cap) {
    const double2* wn = reinterpret_cast<const double2*>(ws.Wn + (int64_t)chain * cap + base);
    double2 a = wn[0], b = wn[1];
    W[0] = a.x; W[1] = a.y; W[2] = b.x; W[3] = b.y;
  } else {
    W[0] = W[1] = W[2] = W[3] = 0.0;
  }
  double qin;
  if (fallback_pass) {
    qin = (base < cap) ? ws.qin_fb[(int64_t)chain * ws.nt * TPB + base / IPT] : 0.0;
  } else {
    pgb_agg te = (base < cap) ? ws.thr_ex[(int64_t)chain * ws.nt * TPB + base / IPT] : pgb_agg_identity();
    qin = w_qin(sh, pgb_agg_combine(pgb_agg_combine(sh.cex[c], sh.wtex[idx]), te));
  }
  q[0] = qin;
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i + 1] = __dadd_rn(q[i], W[i]);
  if (fallback_pass) return true;
  double nxt = __shfl_down_sync(0xffffffffu, qin, 1);
  bool have = (base + IPT < ws.N);
  if (lane == 31 && have) {
    if (idx + 1 < nwt) nxt = w_qin(sh, pgb_agg_combine(sh.cex[c], sh.wtex[idx + 1]));
    else if (c + 1 < nb) nxt = w_qin(sh, sh.cex[c + 1]);
    else have = false;
  }
  int mism = have && (__double_as_longlong(nxt) != __double_as_longlong(q[4]));
  if (mism) atomicOr(flagp, 1);
  return !__any_sync(0xffffffffu, mism);
}

// by-output expansion of one warp-tile (warp-level staging in shared memory)
template <int NX>
PGB_DEV_HELPER void w_consume_by_output(const FilterDev& fd, const EmitCtx<NX>& ec, int chain, int64_t base,
                                                    const double q[5], WShared& sh, int64_t t, int first, int* stat,
                                                    double& lmax, HeavyEnt* heavy, int* heavy_cnt) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t N = fd.N;
  const pgb_utable* ut = &sh.ut;
  const int64_t M = ut->M;
  int64_t rr[5];
  int64_t zero_upto = 0;
  if (base < N) {
    int row = -1;
    rr[0] = pgb_utable_count_le_hint(ut, q[0], &row);
    if (base == 0) {
      if (rr[0] > 0) atomicOr(stat, 2);
      zero_upto = rr[0];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t n = base + i;
      if (n < N) {
        int64_t r = pgb_utable_count_le_hint(ut, q[i + 1], &row);
        if (r < rr[i]) r = rr[i];
        if (n == N - 1 && r < M) {
          atomicOr(stat, 1);
          r = M;
        }
        rr[i + 1] = r;
      } else {
        rr[i + 1] = rr[i];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 5; ++i) rr[i] = M;
  }
  const int64_t jfrom0 = (base == 0) ? 0 : rr[0];
  const int64_t J0 = __shfl_sync(0xffffffffu, (long long)jfrom0, 0);
  const int64_t J1 = __shfl_sync(0xffffffffu, (long long)rr[4], 31);
  bool is_heavy[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t lo = (i == 0) ? jfrom0 : rr[i], hi = rr[i + 1];
    is_heavy[i] = false;
    if (heavy && hi - lo > P_HEAVY && base + i < N) {
      int slot = atomicAdd(heavy_cnt, 1);
      if (slot < P_HEAVY_MAX) {
        HeavyEnt he;
        he.n = base + i; he.jfrom = lo; he.r = hi; he.zero_upto = (i == 0) ? zero_upto : 0;
        heavy[slot] = he;
        is_heavy[i] = true;
      }
    }
  }
  int* anc = sh.s_anc[w];
  for (int64_t cs = J0; cs < J1; cs += W_CAP) {
    const int64_t ce = (cs + W_CAP < J1) ? cs + W_CAP : J1;  // chunk covers draws (cs, ce]
    const int cnt = (int)(ce - cs);
    for (int o = lane; o < cnt; o += 32) anc[o] = -1;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (is_heavy[i]) continue;
      int64_t lo = (i == 0) ? jfrom0 : rr[i], hi = rr[i + 1];
      if (lo < cs) lo = cs;
      if (hi > ce) hi = ce;
      const int64_t zu = (i == 0) ? zero_upto : 0;
      for (int64_t j = lo + 1; j <= hi; ++j) anc[j - cs - 1] = (j <= zu) ? 0 : (int)(base + i) + 1;
    }
    __syncwarp();
    for (int o0 = 0; o0 < cnt; o0 += 128) {
      int64_t jj[4];
      int32_t a0[4];
      bool act[4];
      double F[4][NX];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int o = o0 + k * 32 + lane;
        const int v = (o < cnt) ? anc[o] : -1;
        act[k] = v >= 0;
        const int64_t n = (v <= 0) ? 0 : (int64_t)v - 1;
        a0[k] = (v == 0) ? -1 : (int32_t)n;
        jj[k] = cs + o;
#pragma unroll
        for (int d = 0; d < NX; ++d) F[k][d] = act[k] ? ec.Fb[(int64_t)d * N + n] : 0.0;
      }
      if (act[0] | act[1] | act[2] | act[3]) emit_offspring4<NX>(fd, ec, chain, t, jj, a0, act, F, stat, lmax);
    }
    __syncwarp();
  }
  if (!first && base == 0) {
    double x[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      x[d] = fd.xprim[(int64_t)chain * NX * fd.T + d + NX * t];
      ec.xt[(int64_t)d * N + (N - 1)] = x[d];
    }
    double lw = log_weight<NX>(fd, x, ec.yt);
    ec.lwb[N - 1] = lw;
    lmax = fmax(lmax, lw);
  }
}

// single-index consumer of one thread's four elements (ancestor of the conditioned trajectory, star draw)
PGB_DEV_HELPER void w_consume_index(int64_t N, int64_t base, const double q[5], const pgb_utable* ut,
                                                int32_t* out_idx, int* stat) {
  const int64_t M = ut->M;
  if (base >= N) return;
  int64_t rprev = pgb_utable_count_le(ut, q[0]);
  if (base == 0 && rprev > 0) atomicOr(stat, 2);
  int64_t jfrom = (base == 0) ? 0 : rprev;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t n = base + i;
    if (n >= N) break;
    int64_t r = pgb_utable_count_le(ut, q[i + 1]);
    if (r < rprev) r = rprev;
    if (n == N - 1 && r < M) {
      atomicOr(stat, 1);
      r = M;
    }
    const int64_t zero_upto = (base == 0 && i == 0) ? rprev : 0;
    for (int64_t j = jfrom + 1; j <= r; ++j) out_idx[j - 1] = (j <= zero_upto) ? -1 : (int32_t)n;
    rprev = r;
    jfrom = r;
  }
}

__device__ __forceinline__ void warp_max_to_global(double v, unsigned long long* dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) atomicMax(dst, ordered_key(v));
}

// ------------------------------------------------------------------------------------------------------------
// The kernel is split into one __noinline__ function per phase: each gets its own register allocation (the
// monolithic version spilled ~2 KB per thread and ran ~2.7x slower than the sum of its parts measured in
// isolation), and the rare paths (fix-up, exact ancestor scan, heavy particles) stay out of the hot code.
// The parameter structs are __grid_constant__, so the phase functions read them from the constant bank through
// a reference instead of a per-thread local copy.
// ------------------------------------------------------------------------------------------------------------
struct WCtx {
  const FilterDev* fd;
  const PersistDev* pd;
  WShared* sh;
  const double* A_sm;
  int chain, c, nb, tpc, first;
};

__device__ __forceinline__ int64_t w_base(const WCtx& x, int j) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  return ((int64_t)x.c * (8 * x.tpc) + w * x.tpc + j) * WT + lane * IPT;
}

// t = 0: initial particles (:160-165) and their log-weights
template <int NX>
__device__ __noinline__ void w_phase_init(const WCtx& cx) {
  const FilterDev& fd = *cx.fd;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chain = cx.chain, tpc = cx.tpc;
  const int64_t N = fd.N, T = fd.T;
  double* x0 = x_row(fd, chain, 0);
  const double* xp = fd.xprim + (int64_t)chain * NX * T;
  double lmax = -INFINITY;
  for (int j = 0; j < tpc; ++j) {
    const int64_t b0 = ((int64_t)cx.c * (8 * tpc) + w * tpc + j) * WT;
    for (int i = 0; i < IPT; ++i) {
      int64_t n = b0 + (int64_t)i * 32 + lane;  // coalesced: ownership is irrelevant here
      if (n >= N) continue;
      double x[NX];
      if (n == N - 1) {
#pragma unroll
        for (int d = 0; d < NX; ++d) x[d] = xp[d];
      } else {
        double z[NX];
        draw_normals<NX>(fd, chain, PGB_STREAM_Z0, 0, n, fd.Z0 ? fd.Z0 + (int64_t)chain * NX * (N - 1) : nullptr, z);
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          double acc = 0.0;
#pragma unroll
          for (int e = 0; e <= d; ++e) acc = __fma_rn(fd.x0_chol[d + NX * e], z[e], acc);
          x[d] = __dadd_rn(fd.x0_mean[d], acc);
        }
      }
#pragma unroll
      for (int d = 0; d < NX; ++d) x0[(int64_t)d * N + n] = x[d];
      double lw = log_weight<NX>(fd, x, fd.y);
      fd.lwbuf[(int64_t)chain * N + n] = lw;
      lmax = fmax(lmax, lw);
      if (!(lw == lw)) atomicOr(&fd.status[chain], 4);
    }
  }
  warp_max_to_global(lmax, &fd.maxkey[chain]);
}

// e = exp(lw - max) (:198) and CTA partial sums
__device__ __noinline__ void w_phase_weights(const WCtx& cx) {
  const FilterDev& fd = *cx.fd;
  WShared& sh = *cx.sh;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int chain = cx.chain, tpc = cx.tpc;
  const int64_t N = fd.N;
  const double m = ordered_unkey(fd.maxkey[chain]);
  const double* lw = fd.lwbuf + (int64_t)chain * N;
  double* e = fd.ebuf + (int64_t)chain * N;
  TreeAcc acc;
  acc.init();
  for (int j = 0; j < tpc; ++j) {
    const int64_t base = w_base(cx, j);
    double a[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (base + i < N) ? __dsub_rn(lw[base + i], m) : 0.0;
    vexp<4>(a, v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (base + i < N) e[base + i] = v[i];
      else v[i] = 0.0;
    }
    acc.push(warp_tree_sum(tree4(v)));
  }
  if (lane == 0) sh.wpart[w][0] = acc.result();
  __syncthreads();
  if (tid == 0) cx.pd->pe[(int64_t)chain * cx.nb + cx.c] = tree8(sh.wpart, 0);
  __syncthreads();
}

// phase A: w_{t-1} = e/S1 (:199), F = A*phi(x_{t-1}, u_{t-1}) (:175,:179), waN = w .* pdf (:179)
template <int NX>
__device__ __noinline__ void w_phase_A(const WCtx& cx, int64_t t, bool last_only, bool anc) {
  const FilterDev& fd = *cx.fd;
  const PersistDev& pd = *cx.pd;
  WShared& sh = *cx.sh;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int chain = cx.chain, c = cx.c, nb = cx.nb, tpc = cx.tpc;
  const int64_t N = fd.N, T = fd.T;
  const double S1 = block_tree_sum_array(pd.pe + (int64_t)chain * nb, nb, sh.sm);
  const InvDiv ivS1 = make_invdiv(S1);
  if (c == 0 && tid == 0) {
    fd.maxkey[chain] = 0ull;  // its readers finished before the last barrier
    pd.heavy_cnt[chain] = 0;
    pd.cand[2 * chain] = 0;
  }
  const double* e = fd.ebuf + (int64_t)chain * N;
  double* wgt = fd.wbuf + (int64_t)chain * N;
  double* wh = fd.whist ? fd.whist + ((int64_t)chain * T + (t - 1)) * N : nullptr;
  const double* xp = x_row(fd, chain, t - 1);
  double* Fb = fd.Fbuf + (int64_t)chain * NX * N;
  double* wa = fd.waN + (int64_t)chain * N;
  const double* up = fd.u + (int64_t)fd.n_u * (t - 1);
  double Lq[NX * NX], xref[NX];
  InvDiv ivL[NX];
  double c0 = 0.0;
  if (anc) {
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Lq[i] = fd.Lq[chain * 16 + i];
#pragma unroll
    for (int d = 0; d < NX; ++d) ivL[d] = make_invdiv(Lq[d + NX * d]);
#pragma unroll
    for (int d = 0; d < NX; ++d) xref[d] = fd.xprim[(int64_t)chain * NX * T + d + NX * t];
    c0 = fd.c0[chain];
  }
  TreeAcc accw, acca;
  accw.init();
  acca.init();
  for (int j = 0; j < tpc; ++j) {
    const int64_t base = w_base(cx, j);
    double wv[4], av[4];
    bool live[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      live[i] = (base + i < N);
      wv[i] = live[i] ? div_inv(e[base + i], ivS1) : 0.0;
      av[i] = 0.0;
      if (live[i]) {
        wgt[base + i] = wv[i];
        if (wh) wh[base + i] = wv[i];
      }
    }
    if (!last_only) {
      double x[4][NX], F[4][NX];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < NX; ++d) x[i][d] = live[i] ? xp[(int64_t)d * N + base + i] : 0.0;
      eval_F4<NX>(fd, cx.A_sm, x, up, F);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (live[i]) {
#pragma unroll
          for (int d = 0; d < NX; ++d) Fb[(int64_t)d * N + base + i] = F[i][d];
        }
      if (anc) {
        // pdf(MvNormal(x'_t, Q), F) = exp(c0 - |L^-1 (F - x'_t)|^2 / 2)    (:178-179)
        double ea[4], pv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double r[NX], sq = 0.0;
#pragma unroll
          for (int d = 0; d < NX; ++d) {
            double acc = __dsub_rn(F[i][d], xref[d]);
#pragma unroll
            for (int e2 = 0; e2 < d; ++e2) acc = __fma_rn(-Lq[d + NX * e2], r[e2], acc);
            r[d] = div_inv(acc, ivL[d]);
            sq = __fma_rn(r[d], r[d], sq);
          }
          ea[i] = __dsub_rn(c0, __dmul_rn(sq, 0.5));
        }
        vexp<4>(ea, pv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          av[i] = live[i] ? __dmul_rn(wv[i], pv[i]) : 0.0;
          if (live[i]) wa[base + i] = av[i];
        }
      }
    }
    double sw = warp_tree_sum(tree4(wv));
    if (lane == 0) sh.wtsum_w[w * tpc + j] = sw;
    accw.push(sw);
    if (anc) acca.push(warp_tree_sum(tree4(av)));
  }
  if (lane == 0) {
    sh.wpart[w][0] = accw.result();
    sh.wpart[w][1] = anc ? acca.result() : 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    pd.pw[(int64_t)chain * nb + c] = tree8(sh.wpart, 0);
    if (anc) pd.pa[(int64_t)chain * nb + c] = tree8(sh.wpart, 1);
  }
}

// phase B: main scan-local; ancestor weights normalised (:180).  Returns S2 = sum(w) (the resampler's normaliser, :19)
__device__ __noinline__ double w_phase_B(const WCtx& cx, bool anc, int* fmain) {
  const FilterDev& fd = *cx.fd;
  const PersistDev& pd = *cx.pd;
  WShared& sh = *cx.sh;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int chain = cx.chain, c = cx.c, nb = cx.nb, tpc = cx.tpc;
  const int64_t N = fd.N;
  const double* pw = pd.pw + (int64_t)chain * nb;
  const double S2 = block_tree_sum_array(pw, nb, sh.sm);
  w_approx_prefix(pw, nb, sh.pfx, sh.sm);
  w_scan_local(fd.main, fd.wbuf, S2, sh.wtsum_w, sh.pfx, chain, c, tpc, sh, fd.status, fmain);
  if (anc) {
    const double S3 = block_tree_sum_array(pd.pa + (int64_t)chain * nb, nb, sh.sm);
    const InvDiv ivS3 = make_invdiv(S3);
    double* wa = fd.waN + (int64_t)chain * N;
    TreeAcc acc;
    acc.init();
    for (int j = 0; j < tpc; ++j) {
      const int64_t base = w_base(cx, j);
      double v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = 0.0;
        if (base + i < N) {
          v[i] = div_inv(wa[base + i], ivS3);
          wa[base + i] = v[i];
        }
      }
      double s2 = warp_tree_sum(tree4(v));
      if (lane == 0) sh.wtsum_a2[w * tpc + j] = s2;
      acc.push(s2);
    }
    if (lane == 0) sh.wpart[w][0] = acc.result();
    __syncthreads();
    if (tid == 0) pd.pa2[(int64_t)chain * nb + c] = tree8(sh.wpart, 0);
  }
  return S2;
}

// phase C (main): resolve, exact prefixes, offspring expansion + propagation + log-weights (:172-175, :194)
template <int NX>
__device__ __noinline__ void w_phase_C_main(const WCtx& cx, int64_t t, bool last_only, int* fmain) {
  const FilterDev& fd = *cx.fd;
  const PersistDev& pd = *cx.pd;
  WShared& sh = *cx.sh;
  const int tid = threadIdx.x, w = tid >> 5;
  const int chain = cx.chain, c = cx.c, nb = cx.nb, tpc = cx.tpc, first = cx.first;
  const int64_t N = fd.N, T = fd.T;
  w_resolve(fd.main, chain, nb, tpc, sh, pd.force_fallback & 1, fmain);
  if (tid == 0) {
    sh.sI = *fmain & 1;
    if (c == 0) fd.main.counters[2 * chain + 1] += 1;
  }
  const pgb_utable* utg = last_only ? &fd.utab_star[chain] : &fd.utab_res[(int64_t)chain * T + t];
  {
    const int* src = reinterpret_cast<const int*>(utg);
    int* dst = reinterpret_cast<int*>(&sh.ut);
    for (int i = tid; i < (int)(sizeof(pgb_utable) / 4); i += TPB) dst[i] = src[i];
  }
  __syncthreads();
  const bool skip = sh.sI != 0;
  double lmax = -INFINITY;
  EmitCtx<NX> ec;
  if (!last_only) emit_ctx_init<NX>(ec, fd, chain, t);
  for (int j = 0; j < tpc && !skip; ++j) {
    const int idx = w * tpc + j;
    const int64_t base = w_base(cx, j);
    double W[4], q[5];
    if (!w_exact_prefixes(fd.main, chain, c, idx, nb, tpc, false, sh, W, q, fmain)) continue;
    if (last_only)
      w_consume_index(N, base, q, &sh.ut, pd.star_idx + chain, &fd.main.pend[chain]);
    else
      w_consume_by_output<NX>(fd, ec, chain, base, q, sh, t, first, &fd.main.pend[chain], lmax,
                              pd.heavy + (int64_t)chain * P_HEAVY_MAX, pd.heavy_cnt + chain);
  }
  if (!last_only) warp_max_to_global(lmax, &fd.maxkey[chain]);
}

// phase C (side): certified approximate shortcut for the ancestor index (:181)
__device__ __noinline__ void w_phase_C_side(const WCtx& cx, int64_t t) {
  const FilterDev& fd = *cx.fd;
  const PersistDev& pd = *cx.pd;
  WShared& sh = *cx.sh;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int chain = cx.chain, c = cx.c, nb = cx.nb, tpc = cx.tpc, nwt = 8 * tpc;
  const int64_t N = fd.N, T = fd.T;
  const double* pa2 = pd.pa2 + (int64_t)chain * nb;
  const double S4 = block_tree_sum_array(pa2, nb, sh.sm);
  w_approx_prefix(pa2, nb, sh.pfx, sh.sm);
  if (w == 0) w_wt_prefix(sh.wtsum_a2, nwt, sh.pfx[c], sh.pfx[c + 1], sh.wtpfx);
  __syncthreads();
  const double ua = fd.utab_anc[(int64_t)chain * T + t].u0[0];
  const double delta = 6.0 * ((double)N + 4096.0) * 0x1p-53;
  const double* vc = fd.waN + (int64_t)chain * N;
  const InvDiv ivS4 = make_invdiv(S4);
  for (int j = 0; j < tpc; ++j) {
    const int idx = w * tpc + j;
    const int64_t base = w_base(cx, j);
    double W[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) W[i] = (base + i < N) ? div_inv(vc[base + i], ivS4) : 0.0;
    double s1 = W[0], s2 = __dadd_rn(s1, W[1]), s3 = __dadd_rn(s2, W[2]), s4 = __dadd_rn(s3, W[3]);
    double inc = s4;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double tt = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc = __dadd_rn(tt, inc);
    }
    const double P = __dadd_rn(__dmul_rn(sh.wtpfx[idx], ivS4.y), __dsub_rn(inc, s4));
    double qt[5] = {P, __dadd_rn(P, s1), __dadd_rn(P, s2), __dadd_rn(P, s3), __dadd_rn(P, s4)};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (base + i < N && (qt[i] + delta < ua) && (ua <= qt[i + 1] - delta)) {
        atomicAdd(&pd.cand[2 * chain], 1);
        pd.cand[2 * chain + 1] = (int)(base + i);
      }
  }
}

// rare: exact scan of the ancestor weights when the shortcut is undecided (two grid barriers inside)
__device__ __noinline__ void w_side_exact(const WCtx& cx, int64_t t, bool amb_me, int* fside, unsigned int& bar_target) {
  const FilterDev& fd = *cx.fd;
  const PersistDev& pd = *cx.pd;
  WShared& sh = *cx.sh;
  const int tid = threadIdx.x, w = tid >> 5;
  const int chain = cx.chain, c = cx.c, nb = cx.nb, tpc = cx.tpc;
  const int64_t N = fd.N, T = fd.T;
  const double* pa2 = pd.pa2 + (int64_t)chain * nb;
  const double S4 = block_tree_sum_array(pa2, nb, sh.sm);
  w_approx_prefix(pa2, nb, sh.pfx, sh.sm);
  if (amb_me) w_scan_local(fd.side, fd.waN, S4, sh.wtsum_a2, sh.pfx, chain, c, tpc, sh, fd.status, fside);
  grid_barrier(pd.bar, bar_target);
  if (amb_me) {
    w_resolve(fd.side, chain, nb, tpc, sh, pd.force_fallback & 1, fside);
    if (tid == 0) {
      sh.sI = *fside & 1;
      if (c == 0) fd.side.counters[2 * chain + 1] += 1;
    }
    __syncthreads();
    const bool skip = sh.sI != 0;
    const pgb_utable* ut = &fd.utab_anc[(int64_t)chain * T + t];
    int32_t* out = fd.ah + ((int64_t)chain * T + t) * N + (N - 1);
    for (int j = 0; j < tpc && !skip; ++j) {
      double W[4], q[5];
      if (!w_exact_prefixes(fd.side, chain, c, w * tpc + j, nb, tpc, false, sh, W, q, fside)) continue;
      w_consume_index(N, w_base(cx, j), q, ut, out, &fd.side.pend[chain]);
    }
  }
}

// rare: cooperative expansion of heavy particles, then one extra grid barrier
template <int NX>
__device__ __noinline__ void w_heavy(const WCtx& cx, int64_t t, int* fmain, unsigned int& bar_target) {
  const FilterDev& fd = *cx.fd;
  const PersistDev& pd = *cx.pd;
  const int chain = cx.chain;
  const int hc = pd.heavy_cnt[chain];
  double lmax = -INFINITY;
  if (hc > 0 && !(*fmain & 1))
    p_heavy<NX>(fd, chain, cx.c, cx.nb, t, pd.heavy + (int64_t)chain * P_HEAVY_MAX, hc, &fd.main.pend[chain], lmax);
  warp_max_to_global(lmax, &fd.maxkey[chain]);
  grid_barrier(pd.bar, bar_target);
}

// rare: a seam check failed somewhere -> serial exact walk + redo of the consumers (grid-uniform)
template <int NX>
__device__ __noinline__ void w_fixup(const WCtx& cx, int64_t t, bool last_only, bool anc, double S2, int* fmain,
                                     int* fside, unsigned int& bar_target) {
  const FilterDev& fd = *cx.fd;
  const PersistDev& pd = *cx.pd;
  WShared& sh = *cx.sh;
  const int tid = threadIdx.x, w = tid >> 5;
  const int chain = cx.chain, c = cx.c, nb = cx.nb, tpc = cx.tpc, first = cx.first;
  const int64_t N = fd.N, T = fd.T;
  const int fm = *fmain & 1, fs = anc ? (*fside & 1) : 0;
  const double S4 = (anc && fs) ? block_tree_sum_array(pd.pa2 + (int64_t)chain * nb, nb, sh.sm) : 1.0;
  if (c == 0) {
    if (fm) {
      if (tid == 0) {
        fd.main.pend[chain] = 0;
        if (!last_only) fd.maxkey[chain] = 0ull;
      }
      p_seq_exact(fd.main, fd.wbuf, S2, chain);
    }
    if (fs) {
      if (tid == 0) fd.side.pend[chain] = 0;
      p_seq_exact(fd.side, fd.waN, S4, chain);
    }
  }
  grid_barrier(pd.bar, bar_target);
  double lmax = -INFINITY;
  EmitCtx<NX> ec;
  if (!last_only) emit_ctx_init<NX>(ec, fd, chain, t);
  if (fm) {
    const pgb_utable* utg = last_only ? &fd.utab_star[chain] : &fd.utab_res[(int64_t)chain * T + t];
    const int* src = reinterpret_cast<const int*>(utg);
    int* dst = reinterpret_cast<int*>(&sh.ut);
    for (int i = tid; i < (int)(sizeof(pgb_utable) / 4); i += TPB) dst[i] = src[i];
  }
  __syncthreads();
  for (int j = 0; j < tpc; ++j) {
    const int idx = w * tpc + j;
    const int64_t base = w_base(cx, j);
    double W[4], q[5];
    if (fm) {
      w_exact_prefixes(fd.main, chain, c, idx, nb, tpc, true, sh, W, q, fmain);
      if (last_only)
        w_consume_index(N, base, q, &sh.ut, pd.star_idx + chain, &fd.status[chain]);
      else
        w_consume_by_output<NX>(fd, ec, chain, base, q, sh, t, first, &fd.status[chain], lmax, nullptr, nullptr);
    }
    if (fs) {
      w_exact_prefixes(fd.side, chain, c, idx, nb, tpc, true, sh, W, q, fside);
      w_consume_index(N, base, q, &fd.utab_anc[(int64_t)chain * T + t], fd.ah + ((int64_t)chain * T + t) * N + (N - 1),
                      &fd.status[chain]);
    }
  }
  if (fm && !last_only) warp_max_to_global(lmax, &fd.maxkey[chain]);
  grid_barrier(pd.bar, bar_target);
  if (!last_only) {
    if (fm) w_phase_weights(cx);
    grid_barrier(pd.bar, bar_target);
  }
}

template <int NX>
__global__ void __launch_bounds__(TPB, PGB_PERSIST_MINBLOCKS)
k_sweep_w(const __grid_constant__ FilterDev fd, const __grid_constant__ PersistDev pd, int first) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WShared& sh = *reinterpret_cast<WShared*>(smem_raw);
  double* A_sm = reinterpret_cast<double*>(smem_raw + sizeof(WShared));
  const int tid = threadIdx.x;
  const int nb = pd.nb, tpc = pd.tpc;
  const int chain = blockIdx.x / nb, c = blockIdx.x % nb;
  const int64_t N = fd.N, T = fd.T;
  unsigned int bar_target = 0;
  long long tacc[P_NPROF];
#pragma unroll
  for (int i = 0; i < P_NPROF; ++i) tacc[i] = 0;
  long long tlast = clock64();
#define W_TICK(idx) do { if (tid == 0) { long long now_ = clock64(); tacc[idx] += now_ - tlast; tlast = now_; } } while (0)
  {
    const double* Ag = fd.A + (int64_t)chain * NX * fd.n_phi;
    for (int i = tid; i < NX * fd.n_phi; i += TPB) A_sm[i] = Ag[i];
    if (tid == 0) sh.stage_cnt = 0;
  }
  __syncthreads();
  WCtx cx;
  cx.fd = &fd; cx.pd = &pd; cx.sh = &sh; cx.A_sm = A_sm;
  cx.chain = chain; cx.c = c; cx.nb = nb; cx.tpc = tpc; cx.first = first;

  w_phase_init<NX>(cx);
  grid_barrier(pd.bar, bar_target);
  w_phase_weights(cx);
  grid_barrier(pd.bar, bar_target);
  W_TICK(0);

  for (int64_t t = 1; t <= T; ++t) {
    const bool last_only = (t == T);  // finalise w_T and draw the trajectory index only
    const bool anc = !first && !last_only;
    // flags are double-buffered by step parity so that re-arming never races with the grid-uniform fix-up test
    int* fmain = &fd.main.flag[chain + fd.nc * (int)(t & 1)];
    int* fside = &fd.side.flag[chain + fd.nc * (int)(t & 1)];
    w_phase_A<NX>(cx, t, last_only, anc);
    W_TICK(1);
    grid_barrier(pd.bar, bar_target);  // B1
    W_TICK(2);
    const double S2 = w_phase_B(cx, anc, fmain);
    W_TICK(3);
    grid_barrier(pd.bar, bar_target);  // B2
    W_TICK(5);
    w_phase_C_main<NX>(cx, t, last_only, fmain);
    W_TICK(7);
    if (anc) w_phase_C_side(cx, t);
    W_TICK(8);
    grid_barrier(pd.bar, bar_target);  // B3
    W_TICK(9);
    // ---------------------------------------------------------------- phase D: ancestor index, weights (optimistic)
    if (anc) {
      if (tid == 0) {
        int amb = 0;
        for (int ch = 0; ch < fd.nc; ++ch) amb |= (pd.cand[2 * ch] == 0) || pd.force_fallback;
        sh.any_flag = amb;
      }
      __syncthreads();
      const bool amb_any = sh.any_flag != 0;
      const bool amb_me = (pd.cand[2 * chain] == 0) || pd.force_fallback;
      if (!amb_me && c == 0 && tid == 0) fd.ah[((int64_t)chain * T + t) * N + (N - 1)] = pd.cand[2 * chain + 1];
      if (amb_any) w_side_exact(cx, t, amb_me, fside, bar_target);
    }
    W_TICK(10);
    if (!last_only) {
      if (tid == 0) {
        int any = 0;
        for (int ch = 0; ch < fd.nc; ++ch) any |= pd.heavy_cnt[ch];
        sh.any_flag = any;
      }
      __syncthreads();
      if (sh.any_flag) w_heavy<NX>(cx, t, fmain, bar_target);
      w_phase_weights(cx);
    }
    if (c == 0 && tid == 0) {  // re-arm the other parity's flags (nobody touches them between B1(t) and B1(t+1))
      fd.main.flag[chain + fd.nc * (int)((t + 1) & 1)] = 0;
      fd.side.flag[chain + fd.nc * (int)((t + 1) & 1)] = 0;
    }
    W_TICK(11);
    grid_barrier(pd.bar, bar_target);  // B4 (all consumers done; this step's flags are final)
    W_TICK(12);
    {
      if (tid == 0) {
        int any = 0;
        for (int ch = 0; ch < fd.nc; ++ch)
          any |= (fd.main.flag[ch + fd.nc * (int)(t & 1)] | fd.side.flag[ch + fd.nc * (int)(t & 1)]) & 1;
        sh.any_flag = any;
      }
      __syncthreads();
      if (sh.any_flag) w_fixup<NX>(cx, t, last_only, anc, S2, fmain, fside, bar_target);
      if (c == 0 && tid == 0) {  // merge the status bits raised by consumer passes that were not redone
        int p = fd.main.pend[chain] | fd.side.pend[chain];
        if (p) atomicOr(&fd.status[chain], p);
        fd.main.pend[chain] = 0;
        fd.side.pend[chain] = 0;
      }
    }
    W_TICK(13);
  }
  if (pd.prof && tid == 0) {
#pragma unroll
    for (int i = 0; i < P_NPROF; ++i) pd.prof[(int64_t)blockIdx.x * P_NPROF + i] = tacc[i];
  }
#undef W_TICK
}

}  // namespace pgb
