// pgb_persist.cuh — the whole CPF-AS time loop (Julia/src/particle_Gibbs.jl:160-200) as ONE persistent
// cooperative kernel: T-1 steps x 4 phases separated by grid barriers, instead of 13 launches per step.
//
// Work decomposition: a chain's particles are cut into nb "blocks" of `tpc` (power of two) consecutive
// 1024-element tiles, block c = tiles [c*tpc, (c+1)*tpc).  Per-block partial sums are the tree over that aligned
// range, so the grid-wide sums are still the balanced binary tree of the oracle.  Chain `ch` owns npc consecutive
// CTAs (npc <= nb, NOT a power of two: every co-resident slot of the GPU is used); a CTA processes the blocks
// o, o+npc, o+2npc, ... where o is its rank in an SM-aware order (all first-resident CTAs of the SMs, then the
// second-resident ones), so the blocks-per-SM counts differ by at most one (7/6 tiles per SM at N = 2^20 on 148 SMs
// instead of 8/4 with one power-of-two block per CTA).
//
// Per step t (k > 1):                                                    grid barriers: 4
//   A  w = e/S1, F = A*phi(x_{t-1}), waN = w*N(x'_t;F,Q); CTA sums of w, waN          | B1
//   B  main: W = w/S2, binade prediction, integer maps, segmented composition          |
//      side: waN ./= S3, CTA sums                                                      | B2
//   C  main: resolve (every CTA, redundantly) -> exact prefixes -> offspring expansion,|
//            x_t = F[a] + L z, log-weights, running max                                |
//      side: W' = waN/S4, prediction, maps, composition                                | B3
//   D  side: resolve -> a[t,N];  e = exp(lw - max), CTA sums                           | B4
// The ancestor-sampling chain ("side") lags the main chain by one phase and rides on the same barriers.
// A failed seam check (mispredicted binade) raises a per-chain flag; the fix-up (serial exact walk by
// one warp, redo of the consumer) runs between B4 and the next step, uniformly for the whole grid.
#pragma once
#include "pgb_filter.cuh"

namespace pgb {

constexpr int P_MAXNB = 1024;     // max blocks per chain
constexpr int P_MAXTPC = 64;      // max tiles per CTA (all of its blocks together)
constexpr int P_MAXREC = 160;     // complex threads resolved per scan (more -> serial fallback)
#ifndef PGB_P_HEAVY
#define PGB_P_HEAVY 512
#endif
constexpr int P_HEAVY = PGB_P_HEAVY;  // offspring count above which a particle goes to the cooperative list
constexpr int P_HEAVY_MAX = 1 << 15;
constexpr int P_NPROF = 16;
constexpr int P_CAP = 2048;        // output slots staged per chunk in shared memory

struct HeavyEnt {
  int64_t n;          // ancestor (0-based)
  int64_t jfrom, r;   // its draws are (jfrom, r]
  int64_t zero_upto;  // draws <= zero_upto get index 0 (only the very first element, u_1 <= 0)
};

struct PersistDev {
  int nb, tpc;              // blocks per chain, tiles per block
  int npc;                  // CTAs per chain
  int* sm_slots;            // [1024] residency counter per SM id (zeroed before launch)
  int* cta_key;             // [grid] (residency slot << 12) | SM id
  unsigned int* bar;        // [1] grid barrier counter (zeroed before launch)
  double *pe, *pw, *pa, *pa2;  // [nc][nb] CTA partial sums
  int32_t* star_idx;        // [nc]
  int force_fallback;
  // particles with more than P_HEAVY offspring are expanded by the whole chain's CTAs (phase D)
  struct HeavyEnt* heavy;   // [nc][P_HEAVY_MAX]
  int* heavy_cnt;           // [nc]
  long long* prof;          // [grid][P_NPROF] per-CTA cycle counts per section (thread 0), or null
  int* cand;                // [nc][2] ancestor draw: {number of certain hits, index} (approximate-scan shortcut)
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the grid (co-resident: cooperative launch).  `target` is this CTA's running count.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (ld_acquire_u32(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// tree over the CTA's tpc tile sums (tpc power of two): binary-counter stack, uniform across threads
struct TreeAcc {
  double st[8];
  int sp, i;
  __device__ __forceinline__ void init() { sp = 0; i = 0; }
  __device__ __forceinline__ void push(double x) {
    int merges = __ffs(i + 1) - 1;
    for (int m = 0; m < merges; ++m) x = __dadd_rn(st[--sp], x);
    st[sp++] = x;
    ++i;
  }
  __device__ __forceinline__ double result() const { return st[0]; }
};

struct PersistShared {
  double sm[NWARP + 1];
  pgb_agg smagg[NWARP + 1];
  double sP[TPB];
  double sQ[TPB];
  int sI;
  double tsum_w[P_MAXTPC];   // tile sums of w      (phase A -> B)
  double tsum_a2[P_MAXTPC];  // tile sums of waN'   (phase B -> C)
  pgb_agg cex[P_MAXNB + 1];  // exclusive scan of the chain's CTA aggregates
  double pfx[P_MAXNB + 1];   // approximate exclusive prefix of CTA partial sums
  pgb_map rf[P_MAXREC];
  double rW[P_MAXREC][4];
  double seeds[P_MAXREC + 1];
  int any_flag;
  long long sJ[2];           // output range of the current tile
  int s_anc[P_CAP];          // ancestor (+1; 0 = index 0; -1 = skip) of each output slot of the current chunk
  pgb_utable ut;             // this step's sampling grid
};

// exclusive approximate scan of p[0..nb) into sh.pfx[0..nb] (every CTA computes identical values)
__device__ __forceinline__ void approx_prefix(const double* p, int nb, PersistShared& sh) {
  const int per = (nb + TPB - 1) / TPB;  // <= P_MAXNB / TPB consecutive entries per thread
  const int i0 = threadIdx.x * per;
  double v[P_MAXNB / TPB];
  double run = 0.0;
#pragma unroll
  for (int j = 0; j < P_MAXNB / TPB; ++j) {
    v[j] = (j < per && i0 + j < nb) ? p[i0 + j] : 0.0;
    run = __dadd_rn(run, v[j]);
  }
  double tot;
  double ex = block_excl_scan_approx(run, sh.sm, &tot);
#pragma unroll
  for (int j = 0; j < P_MAXNB / TPB; ++j) {
    if (j < per && i0 + j < nb) sh.pfx[i0 + j] = ex;
    ex = __dadd_rn(ex, v[j]);
  }
  if (threadIdx.x == 0) sh.pfx[nb] = tot;
  __syncthreads();
}

// Phase "scan local" for the CTA's tiles: v = values (already the numerators), S the normaliser.
__device__ __forceinline__ void p_scan_local(const ScanWS& ws, const double* v, double S, const double* tsum_sm,
                                             int chain, int c, int nb, int tpc, PersistShared& sh, int* status,
                                             int* flagp) {
  const int tid = threadIdx.x;
  const int64_t N = ws.N;
  const double* vc = v + (int64_t)chain * N;
  const InvDiv ivS = make_invdiv(S);
  pgb_agg carry = pgb_agg_identity();
  double praw = sh.pfx[c];
  for (int k = 0; k < tpc; ++k) {
    const int tile = c * tpc + k;
    if (tile >= ws.nt) break;
    int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
    double W[4];
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      W[i] = (base + i < N) ? div_inv(vc[base + i], ivS) : 0.0;
      if (!(W[i] >= 0.0 && W[i] <= 1.7976931348623157e308)) bad = true;
    }
    if (bad) atomicOr(&status[chain], 4);
    {
      double2* wn = reinterpret_cast<double2*>(ws.Wn + ((int64_t)chain * ws.nt + tile) * TILE + (int64_t)tid * IPT);
      wn[0] = make_double2(W[0], W[1]);
      wn[1] = make_double2(W[2], W[3]);
    }
    double s1 = W[0], s2 = __dadd_rn(s1, W[1]), s3 = __dadd_rn(s2, W[2]), s4 = __dadd_rn(s3, W[3]);
    double tot;
    double ex = block_excl_scan_approx(s4, sh.sm, &tot);
    const bool last_tile = (k == tpc - 1) || (tile + 1 >= ws.nt);
    double pnext_raw = last_tile ? sh.pfx[c + 1] : __dadd_rn(praw, tsum_sm[k]);
    double Pc = __dmul_rn(praw, ivS.y), Pn = __dmul_rn(pnext_raw, ivS.y);  // approximate quantities only
    double P = __dadd_rn(Pc, ex);
    __syncthreads();
    sh.sP[tid] = P;
    __syncthreads();
    double qt[5];
    qt[0] = P;
    qt[1] = __dadd_rn(P, s1);
    qt[2] = __dadd_rn(P, s2);
    qt[3] = __dadd_rn(P, s3);
    qt[4] = (tid + 1 < TPB) ? sh.sP[tid + 1] : Pn;
    pgb_agg agg = bad ? pgb_agg_complex() : pgb_classify4(W, qt);
    pgb_agg total;
    pgb_agg exagg = pgb_agg_combine(carry, block_excl_scan_agg(agg, sh.smagg, &total));
    ws.thr_ex[((int64_t)chain * ws.nt + tile) * TPB + tid] = exagg;  // CTA-local exclusive aggregate
    if (agg.cnt) {
      if (exagg.cnt < tpc * RMAX) {
        ScanRec r;
        r.ex = exagg;
        r.W[0] = W[0]; r.W[1] = W[1]; r.W[2] = W[2]; r.W[3] = W[3];
        r.pad_ = 0.0;
        ws.recs[((int64_t)chain * ws.nt + (int64_t)c * tpc) * RMAX + exagg.cnt] = r;
      } else {
        atomicOr(flagp, 1);
      }
    }
    carry = pgb_agg_combine(carry, total);
    praw = pnext_raw;
  }
  if (tid == 0) ws.tile_agg[(int64_t)chain * ws.nt + c] = carry;  // CTA aggregate (slot c of tile_agg)
}

// Ancestor index of the conditioned trajectory (particle_Gibbs.jl:181) without the exact machinery: the draw is a
// single comparison q_{n-1} < u <= q_n, so an ordinary (approximate) scan decides it whenever u is farther than a
// rigorous error bound `delta` from both neighbouring prefixes: |q~_n - q_n(sequential)| <= delta implies that an
// element with q~_{n-1} + delta < u <= q~_n - delta IS the sequential answer (q is monotone, the answer unique).
// delta bounds both the sequential rounding error (< n 2^-53 sum) and that of the blocked scan.  No certain
// element (probability ~ 2 delta / W ~ 1e-3 per step) -> the caller runs the exact scan for this step.
__device__ __forceinline__ void p_side_approx(const ScanWS& ws, const double* v, double S, const double* tsum_sm,
                                              int chain, int c, int tpc, PersistShared& sh, double u, double delta,
                                              int* cand) {
  const int tid = threadIdx.x;
  const int64_t N = ws.N;
  const double* vc = v + (int64_t)chain * N;
  double praw = sh.pfx[c];
  const InvDiv ivS = make_invdiv(S);
  for (int k = 0; k < tpc; ++k) {
    const int tile = c * tpc + k;
    if (tile >= ws.nt) break;
    int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
    double W[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) W[i] = (base + i < N) ? div_inv(vc[base + i], ivS) : 0.0;
    double s1 = W[0], s2 = __dadd_rn(s1, W[1]), s3 = __dadd_rn(s2, W[2]), s4 = __dadd_rn(s3, W[3]);
    double tot;
    double ex = block_excl_scan_approx(s4, sh.sm, &tot);
    double P = __dadd_rn(__dmul_rn(praw, ivS.y), ex);
    double qt[5] = {P, __dadd_rn(P, s1), __dadd_rn(P, s2), __dadd_rn(P, s3), __dadd_rn(P, s4)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (base + i < N && (qt[i] + delta < u) && (u <= qt[i + 1] - delta)) {
        atomicAdd(&cand[0], 1);
        cand[1] = (int)(base + i);
      }
    }
    praw = __dadd_rn(praw, tsum_sm[k]);
  }
}

// Resolve (every CTA redundantly): exclusive scan of CTA aggregates, gather records, serial chain.
__device__ __forceinline__ void p_resolve(const ScanWS& ws, int chain, int nb, int tpc, PersistShared& sh,
                                          int force_fallback, int* flagp) {
  const int tid = threadIdx.x;
  const pgb_agg* cagg = ws.tile_agg + (int64_t)chain * ws.nt;
  pgb_agg carry;
  const int per = (nb + TPB - 1) / TPB;
  const int i0 = tid * per;
  pgb_agg v[P_MAXNB / TPB], exs[P_MAXNB / TPB];
  {
    pgb_agg run = pgb_agg_identity();
#pragma unroll
    for (int j = 0; j < P_MAXNB / TPB; ++j) {
      v[j] = (j < per && i0 + j < nb) ? cagg[i0 + j] : pgb_agg_identity();
      run = pgb_agg_combine(run, v[j]);
    }
    pgb_agg ex = block_excl_scan_agg(run, sh.smagg, &carry);
#pragma unroll
    for (int j = 0; j < P_MAXNB / TPB; ++j) {
      exs[j] = ex;
      if (j < per && i0 + j < nb) sh.cex[i0 + j] = ex;
      ex = pgb_agg_combine(ex, v[j]);
    }
  }
  if (tid == 0) sh.cex[nb] = carry;
  const int C = carry.cnt;
  if (C > P_MAXREC || force_fallback) {
    if (tid == 0) atomicOr(flagp, 1);
    if (tid == 0) sh.seeds[0] = 0.0;
    __syncthreads();
    return;
  }
  // gather the complex threads' records of this thread's blocks (the aggregates are still in registers)
#pragma unroll
  for (int j = 0; j < P_MAXNB / TPB; ++j) {
    int cnt = v[j].cnt;
    if (cnt > tpc * RMAX) cnt = tpc * RMAX;
    const pgb_agg te = exs[j];
    for (int i = 0; i < cnt; ++i) {
      const ScanRec* r = &ws.recs[((int64_t)chain * ws.nt + (int64_t)(i0 + j) * tpc) * RMAX + i];
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

// exact prefixes for one tile of this CTA.  Returns false (CTA-uniform) when the consumer must be skipped.
__device__ __forceinline__ bool p_exact_prefixes(const ScanWS& ws, const double* v, double S, int chain, int c, int k,
                                                 int nb, int tpc, bool fallback_pass, PersistShared& sh, double W[4],
                                                 double q[5], int* flagp) {
  const int tid = threadIdx.x;
  const int tile = c * tpc + k;
  int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
  {
    const double2* wn = reinterpret_cast<const double2*>(ws.Wn + ((int64_t)chain * ws.nt + tile) * TILE + (int64_t)tid * IPT);
    double2 a = wn[0], b = wn[1];
    W[0] = a.x; W[1] = a.y; W[2] = b.x; W[3] = b.y;
  }
  const int64_t gthr = ((int64_t)chain * ws.nt + tile) * TPB + tid;
  double qin;
  if (fallback_pass) {
    qin = ws.qin_fb[gthr];
  } else {
    pgb_agg ex = pgb_agg_combine(sh.cex[c], ws.thr_ex[gthr]);
    int cc = ex.cnt < 0 ? 0 : (ex.cnt > P_MAXREC ? P_MAXREC : ex.cnt);
    qin = pgb_qin_from(ex, sh.seeds[cc]);
  }
  q[0] = qin;
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i + 1] = __dadd_rn(q[i], W[i]);
  if (!fallback_pass) {
    __syncthreads();
    sh.sQ[tid] = qin;
    __syncthreads();
    double nxt = 0.0;
    bool have = (base + IPT < ws.N);
    if (have) {
      if (tid + 1 < TPB) {
        nxt = sh.sQ[tid + 1];
      } else if (tile + 1 < ws.nt) {
        const bool same_cta = (k + 1 < tpc);
        pgb_agg ex = pgb_agg_combine(sh.cex[same_cta ? c : c + 1], ws.thr_ex[gthr + 1]);
        int cc = ex.cnt < 0 ? 0 : (ex.cnt > P_MAXREC ? P_MAXREC : ex.cnt);
        nxt = pgb_qin_from(ex, sh.seeds[cc]);
      } else {
        have = false;
      }
    }
    int mism = have && (__double_as_longlong(nxt) != __double_as_longlong(q[4]));
    if (mism) atomicOr(flagp, 1);
    if (__syncthreads_or(mism)) return false;
  }
  return true;
}

// Propagation of one resampled particle: a[t,j], x_t^j = F[a_j] + L z_j (:175|:188), log-weight (:194).
template <int NX>
struct EmitCtx {
  double Lq[NX * NX];
  const double* Fb;
  double* xt;
  int32_t* at;
  double* lwb;
  const double* yt;
  const double* zin;
  InvDiv ivR;
};
template <int NX>
__device__ __forceinline__ void emit_ctx_init(EmitCtx<NX>& ec, const FilterDev& fd, int chain, int64_t t) {
  const int64_t N = fd.N;
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) ec.Lq[i] = fd.Lq[chain * 16 + i];
  ec.Fb = fd.Fbuf + (int64_t)chain * NX * N;
  ec.xt = x_row(fd, chain, t);
  ec.at = fd.ah + ((int64_t)chain * fd.T + t) * N;
  ec.lwb = fd.lwbuf + (int64_t)chain * N;
  ec.yt = fd.y + (int64_t)fd.n_y * t;
  ec.zin = fd.Z ? fd.Z + ((int64_t)chain * fd.T + t) * NX * N : nullptr;
  ec.ivR = make_invdiv(fd.R[0]);
}
template <int NX>
__device__ __forceinline__ void emit_offspring(const FilterDev& fd, const EmitCtx<NX>& ec, int chain, int64_t t,
                                               int64_t jj, int32_t a0, const double F[NX], int* stat, double& lmax) {
  const int64_t N = fd.N;
  ec.at[jj] = a0;
  double z[NX], x[NX];
  draw_normals<NX>(fd, chain, PGB_STREAM_Z, (uint32_t)t, jj, ec.zin, z);
#pragma unroll
  for (int d = 0; d < NX; ++d) {
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e <= d; ++e) acc = __fma_rn(ec.Lq[d + NX * e], z[e], acc);
    x[d] = __dadd_rn(F[d], acc);
    ec.xt[(int64_t)d * N + jj] = x[d];
  }
  double lw = log_weight<NX>(fd, x, ec.yt);
  ec.lwb[jj] = lw;
  lmax = fmax(lmax, lw);
  if (!(lw == lw)) atomicOr(stat, 4);
}

// Four output slots at once (independent Philox / Box-Muller / log-weight chains interleaved).  Same values as
// emit_offspring for every slot.
template <int NX>
PGB_DEV_HELPER void emit_offspring4(const FilterDev& fd, const EmitCtx<NX>& ec, int chain, int64_t t,
                                                const int64_t (&jj)[4], const int32_t (&a0)[4], const bool (&act)[4],
                                                const double (&F)[4][NX], int* stat, double& lmax) {
  constexpr int NP = (NX + 1) / 2;
  const int64_t N = fd.N;
  double z[4][2 * NP];
  if (fd.philox) {
    pgb_u32x4 ctr[4 * NP];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        ctr[k * NP + p].v[0] = (uint32_t)(jj[k] * NP + p);
        ctr[k * NP + p].v[1] = (uint32_t)t;
        ctr[k * NP + p].v[2] = fd.sweep;
        ctr[k * NP + p].v[3] = ((uint32_t)PGB_STREAM_Z << 24) | ((fd.chain_base + chain) & 0x00ffffffu);
      }
    vphilox<4 * NP>(ctr, (uint32_t)fd.seed, (uint32_t)(fd.seed >> 32));
    double z0[4 * NP], z1[4 * NP];
    vnormal_pairs<4 * NP>(ctr, z0, z1);
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        z[k][2 * p] = z0[k * NP + p];
        z[k][2 * p + 1] = z1[k * NP + p];
      }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int d = 0; d < NX; ++d) z[k][d] = act[k] ? ec.zin[jj[k] * NX + d] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double x[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      double acc = 0.0;
#pragma unroll
      for (int e = 0; e <= d; ++e) acc = __fma_rn(ec.Lq[d + NX * e], z[k][e], acc);
      x[d] = __dadd_rn(F[k][d], acc);
    }
    double lw = log_weight_iv<NX>(fd, x, ec.yt, ec.ivR);
    if (act[k]) {
      ec.at[jj[k]] = a0[k];
#pragma unroll
      for (int d = 0; d < NX; ++d) ec.xt[(int64_t)d * N + jj[k]] = x[d];
      ec.lwb[jj[k]] = lw;
      lmax = fmax(lmax, lw);
      if (!(lw == lw)) atomicOr(stat, 4);
    }
  }
}

// Consumer for one tile (same arithmetic as k_expand).  heavy != nullptr: particles with more than
// P_HEAVY offspring are deferred to the chain-wide cooperative list instead of being expanded by one thread.
template <int NX, int MODE>
__device__ __forceinline__ void p_consume(const FilterDev& fd, int chain, int tile, const double q[5],
                                          const pgb_utable* ut, int32_t* out_idx, int64_t t, int first, int* stat,
                                          double& lmax, HeavyEnt* heavy, int* heavy_cnt) {
  const int tid = threadIdx.x;
  const int64_t N = fd.N;
  const int64_t M = ut->M;
  int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
  EmitCtx<NX> ec;
  if (MODE == MODE_PROPAGATE) emit_ctx_init<NX>(ec, fd, chain, t);
  if (base < N) {
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
      if (MODE == MODE_INDEX) {
        for (int64_t j = jfrom + 1; j <= r; ++j) out_idx[j - 1] = (j <= zero_upto) ? -1 : (int32_t)n;
      } else if (r > jfrom) {
        bool deferred = false;
        if (heavy && r - jfrom > P_HEAVY) {
          int slot = atomicAdd(heavy_cnt, 1);
          if (slot < P_HEAVY_MAX) {
            HeavyEnt he;
            he.n = n; he.jfrom = jfrom; he.r = r; he.zero_upto = zero_upto;
            heavy[slot] = he;
            deferred = true;
          }
        }
        if (!deferred) {
          double F[NX];
#pragma unroll
          for (int d = 0; d < NX; ++d) F[d] = ec.Fb[(int64_t)d * N + n];
          for (int64_t j = jfrom + 1; j <= r; ++j)
            emit_offspring<NX>(fd, ec, chain, t, j - 1, (j <= zero_upto) ? -1 : (int32_t)n, F, stat, lmax);
        }
      }
      rprev = r;
      jfrom = r;
    }
  }
  if (MODE == MODE_PROPAGATE && !first && tile == 0 && tid == 0) {
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

// Cooperative expansion of the deferred heavy particles by all nb CTAs of the chain.
template <int NX>
__device__ __forceinline__ void p_heavy(const FilterDev& fd, int chain, int c, int nb, int64_t t, const HeavyEnt* heavy,
                                        int cnt, int* stat, double& lmax) {
  if (cnt > P_HEAVY_MAX) cnt = P_HEAVY_MAX;
  EmitCtx<NX> ec;
  emit_ctx_init<NX>(ec, fd, chain, t);
  const int64_t N = fd.N;
  const int64_t stride = (int64_t)nb * TPB;
  for (int e = 0; e < cnt; ++e) {
    const HeavyEnt he = heavy[e];
    double F[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) F[d] = ec.Fb[(int64_t)d * N + he.n];
    // rotate the starting CTA with e so that many medium-heavy particles spread over the chain's CTAs
    int64_t off = ((int64_t)((c + e) % nb)) * TPB + threadIdx.x;
    for (int64_t j = he.jfrom + 1 + off; j <= he.r; j += stride)
      emit_offspring<NX>(fd, ec, chain, t, j - 1, (j <= he.zero_upto) ? -1 : (int32_t)he.n, F, stat, lmax);
  }
}

// By-output expansion of one tile (normal pass, MODE_PROPAGATE): the tile's elements own a contiguous range of
// output slots; each thread first *scatters* the ancestor index of its few offspring into shared memory, then
// every thread *propagates* one output slot at a time — no SIMT divergence in the expensive part (Philox,
// Box-Muller, log-weight) and fully coalesced stores of a[t,:], x_t, lw.
template <int NX>
__device__ __forceinline__ void p_consume_by_output(const FilterDev& fd, const EmitCtx<NX>& ec, int chain, int tile,
                                                    const double q[5], PersistShared& sh, int64_t t, int first,
                                                    int* stat, double& lmax, HeavyEnt* heavy, int* heavy_cnt) {
  const int tid = threadIdx.x;
  const int64_t N = fd.N;
  const pgb_utable* ut = &sh.ut;
  const int64_t M = ut->M;
  const int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
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
  __syncthreads();
  if (tid == 0) sh.sJ[0] = jfrom0;
  if (tid == TPB - 1) sh.sJ[1] = rr[4];
  __syncthreads();
  const int64_t J0 = sh.sJ[0], J1 = sh.sJ[1];
  // heavy particles go to the chain-wide cooperative list (once)
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
  for (int64_t cs = J0; cs < J1; cs += P_CAP) {
    const int64_t ce = (cs + P_CAP < J1) ? cs + P_CAP : J1;  // chunk covers draws (cs, ce]
    for (int o = tid; o < (int)(ce - cs); o += TPB) sh.s_anc[o] = -1;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (is_heavy[i]) continue;
      int64_t lo = (i == 0) ? jfrom0 : rr[i], hi = rr[i + 1];
      if (lo < cs) lo = cs;
      if (hi > ce) hi = ce;
      const int64_t zu = (i == 0) ? zero_upto : 0;
      for (int64_t j = lo + 1; j <= hi; ++j) sh.s_anc[j - cs - 1] = (j <= zu) ? 0 : (int)(base + i) + 1;
    }
    __syncthreads();
    const int cnt = (int)(ce - cs);
    for (int o0 = 0; o0 < cnt; o0 += 4 * TPB) {
      int64_t jj[4];
      int32_t a0[4];
      bool act[4];
      double F[4][NX];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int o = o0 + k * TPB + tid;
        const int v = (o < cnt) ? sh.s_anc[o] : -1;
        act[k] = v >= 0;
        const int64_t n = (v <= 0) ? 0 : (int64_t)v - 1;
        a0[k] = (v == 0) ? -1 : (int32_t)n;
        jj[k] = cs + o;
#pragma unroll
        for (int d = 0; d < NX; ++d) F[k][d] = act[k] ? ec.Fb[(int64_t)d * N + n] : 0.0;
      }
      if (act[0] | act[1] | act[2] | act[3]) emit_offspring4<NX>(fd, ec, chain, t, jj, a0, act, F, stat, lmax);
    }
    __syncthreads();
  }
  if (!first && tile == 0 && tid == 0) {
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

// serial exact walk by warp 0 of the chain's first CTA (fix-up path)
__device__ __forceinline__ void p_seq_exact(const ScanWS& ws, const double* v, double S, int chain) {
  const int lane = threadIdx.x;
  if (lane >= 32) return;
  const double* vc = v + (int64_t)chain * ws.N;
  double* qin = ws.qin_fb + (int64_t)chain * ws.nt * TPB;
  const int64_t nthr = (int64_t)ws.nt * TPB;
  double q = 0.0;
  for (int64_t t0 = 0; t0 < nthr; t0 += 32) {
    int64_t base = (t0 + lane) * IPT;
    double W[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) W[i] = (base + i < ws.N) ? __ddiv_rn(vc[base + i], S) : 0.0;
    for (int l = 0; l < 32; ++l) {
      if (lane == l) qin[t0 + l] = q;
#pragma unroll
      for (int i = 0; i < 4; ++i) q = __dadd_rn(q, __shfl_sync(0xffffffffu, W[i], l));
    }
  }
  if (lane == 0) ws.counters[2 * chain + 0] += 1;
}

#ifndef PGB_PERSIST_MINBLOCKS
#define PGB_PERSIST_MINBLOCKS 2
#endif
template <int NX>
__global__ void __launch_bounds__(TPB, PGB_PERSIST_MINBLOCKS)
k_sweep_persistent(const __grid_constant__ FilterDev fd, const __grid_constant__ PersistDev pd, int first) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PersistShared& sh = *reinterpret_cast<PersistShared*>(smem_raw);
  double* A_sm = reinterpret_cast<double*>(smem_raw + sizeof(PersistShared));
  const int tid = threadIdx.x;
  const int nb = pd.nb, tpc = pd.tpc, npc = pd.npc;
  const int chain = blockIdx.x / npc;
  const int64_t N = fd.N, T = fd.T;
  const int nt = fd.nt;
  unsigned int bar_target = 0;
  // SM-aware rank of this CTA inside its chain: first-resident CTAs of all SMs come first, so that the chain's
  // blocks (dealt round-robin over ranks) spread evenly over SMs rather than over CTAs
  if (tid == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int slot = atomicAdd(&pd.sm_slots[smid & 1023], 1);
    pd.cta_key[blockIdx.x] = (slot << 12) | (int)(smid & 1023);
  }
  grid_barrier(pd.bar, bar_target);
  int o = 0;
  {
    const int mykey = __ldcg(&pd.cta_key[blockIdx.x]);
    for (int pass = 0; pass < (npc + TPB - 1) / TPB; ++pass) {
      const int i = pass * TPB + tid;
      o += __syncthreads_count(i < npc && __ldcg(&pd.cta_key[chain * npc + i]) < mykey);
    }
  }
  const bool lead = (o == 0);  // one designated CTA per chain for the scalar bookkeeping
#define P_FOR_BLOCKS for (int vi = 0, c = o; c < nb; ++vi, c += npc)
  // section timers (thread 0 of every CTA; SM clock) exist only in -DPGB_EXPERIMENT builds, never in the shipped library
#ifdef PGB_EXPERIMENT
  long long tacc[P_NPROF];
#pragma unroll
  for (int i = 0; i < P_NPROF; ++i) tacc[i] = 0;
  long long tlast = clock64();
#define P_TICK(idx) do { if (tid == 0) { long long now_ = clock64(); tacc[idx] += now_ - tlast; tlast = now_; } } while (0)
#else
#define P_TICK(idx) do { } while (0)
#endif
  double* pe = pd.pe + (int64_t)chain * nb;
  double* pw = pd.pw + (int64_t)chain * nb;
  double* pa = pd.pa + (int64_t)chain * nb;
  double* pa2 = pd.pa2 + (int64_t)chain * nb;
  {
    const double* Ag = fd.A + (int64_t)chain * NX * fd.n_phi;
    for (int i = tid; i < NX * fd.n_phi; i += TPB) A_sm[i] = Ag[i];
  }
  __syncthreads();

  // ---------------------------------------------------------------- t = 0: initial particles (:160-165)
  {
    double* x0 = x_row(fd, chain, 0);
    const double* xp = fd.xprim + (int64_t)chain * NX * T;
    double lmax = -INFINITY;
    P_FOR_BLOCKS for (int k = 0; k < tpc; ++k) {
      const int tile = c * tpc + k;
      if (tile >= nt) break;
      for (int i = 0; i < IPT; ++i) {
        int64_t n = (int64_t)tile * TILE + (int64_t)i * TPB + tid;  // coalesced: ownership is irrelevant here
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
    block_max_to_global(lmax, &fd.maxkey[chain], sh.sm);
  }
  grid_barrier(pd.bar, bar_target);

  // weight update e = exp(lw - max) (:198) and CTA partial sums, for the CTA's tiles
  auto weights_phase = [&]() {
    const double m = ordered_unkey(fd.maxkey[chain]);
    const double* lw = fd.lwbuf + (int64_t)chain * N;
    double* e = fd.ebuf + (int64_t)chain * N;
    P_FOR_BLOCKS {
      TreeAcc acc;
      acc.init();
      for (int k = 0; k < tpc; ++k) {
        const int tile = c * tpc + k;
        int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
        double a[4], v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = (tile < nt && base + i < N) ? __dsub_rn(lw[base + i], m) : 0.0;
        vexp<4>(a, v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (tile < nt && base + i < N) e[base + i] = v[i];
          else v[i] = 0.0;
        }
        acc.push(block_tree_sum(tree4(v), sh.sm));
      }
      if (tid == 0) pe[c] = acc.result();
    }
  };
  weights_phase();
  grid_barrier(pd.bar, bar_target);

  P_TICK(0);
  for (int64_t t = 1; t <= T; ++t) {
    const bool last_only = (t == T);  // finalise w_T and draw the trajectory index only
    const bool anc = !first && !last_only;
    // flags are double-buffered by step parity so that resetting never races with the uniform fix-up test
    int* fmain = &fd.main.flag[chain + fd.nc * (int)(t & 1)];
    int* fside = &fd.side.flag[chain + fd.nc * (int)(t & 1)];
    // ================================================================ phase A: w_{t-1}, F, waN
    {
      const double S1 = block_tree_sum_array(pe, nb, sh.sm);
      const InvDiv ivS1 = make_invdiv(S1);
      if (lead && tid == 0) {
        fd.maxkey[chain] = 0ull;  // its readers finished before the last barrier
        pd.heavy_cnt[chain] = 0;
        pd.cand[2 * chain] = 0;
      }
      const double* e = fd.ebuf + (int64_t)chain * N;
      double* w = fd.wbuf + (int64_t)chain * N;
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
      P_FOR_BLOCKS {
      TreeAcc accw, acca;
      accw.init();
      acca.init();
      for (int k = 0; k < tpc; ++k) {
        const int tile = c * tpc + k;
        int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
        double wv[4], av[4];
        bool live[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int64_t n = base + i;
          live[i] = (tile < nt && n < N);
          wv[i] = live[i] ? div_inv(e[n], ivS1) : 0.0;
          av[i] = 0.0;
          if (live[i]) {
            w[n] = wv[i];
            if (wh) wh[n] = wv[i];
          }
        }
        if (!last_only) {
          double x[4][NX], F[4][NX];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int d = 0; d < NX; ++d) x[i][d] = live[i] ? xp[(int64_t)d * N + base + i] : 0.0;
          eval_F4<NX>(fd, A_sm, x, up, F);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (live[i]) {
#pragma unroll
              for (int d = 0; d < NX; ++d) Fb[(int64_t)d * N + base + i] = F[i][d];
            }
          if (anc) {
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
        double sw = block_tree_sum(tree4(wv), sh.sm);
        if (tid == 0) sh.tsum_w[vi * tpc + k] = sw;
        accw.push(sw);
        if (anc) acca.push(block_tree_sum(tree4(av), sh.sm));
      }
      if (tid == 0) {
        pw[c] = accw.result();
        if (anc) pa[c] = acca.result();
      }
      }
    }
    P_TICK(1);
    grid_barrier(pd.bar, bar_target);  // B1
    P_TICK(2);
    // ================================================================ phase B: main scan-local, side normalise
    const double S2 = block_tree_sum_array(pw, nb, sh.sm);
    {
      approx_prefix(pw, nb, sh);
      P_FOR_BLOCKS p_scan_local(fd.main, fd.wbuf, S2, sh.tsum_w + vi * tpc, chain, c, nb, tpc, sh, fd.status, fmain);
      P_TICK(3);
      if (anc) {
        const double S3 = block_tree_sum_array(pa, nb, sh.sm);
        const InvDiv ivS3 = make_invdiv(S3);
        double* wa = fd.waN + (int64_t)chain * N;
        P_FOR_BLOCKS {
          TreeAcc acc;
          acc.init();
          for (int k = 0; k < tpc; ++k) {
            const int tile = c * tpc + k;
            int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
            double v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[i] = 0.0;
              if (tile < nt && base + i < N) {
                v[i] = div_inv(wa[base + i], ivS3);
                wa[base + i] = v[i];
              }
            }
            double s2 = block_tree_sum(tree4(v), sh.sm);
            if (tid == 0) sh.tsum_a2[vi * tpc + k] = s2;
            acc.push(s2);
          }
          if (tid == 0) pa2[c] = acc.result();
        }
      }
      P_TICK(4);
    }
    grid_barrier(pd.bar, bar_target);  // B2
    P_TICK(5);
    // ================================================================ phase C: main resolve + consume, side scan-local
    {
      p_resolve(fd.main, chain, nb, tpc, sh, pd.force_fallback & 1, fmain);
      P_TICK(6);
      if (tid == 0) {
        sh.sI = *fmain & 1;
        if (lead) fd.main.counters[2 * chain + 1] += 1;
      }
      __syncthreads();
      const bool skip = sh.sI != 0;
      double lmax = -INFINITY;
      const pgb_utable* ut = last_only ? &fd.utab_star[chain] : &fd.utab_res[(int64_t)chain * T + t];
      {  // stage this step's sampling grid in shared memory (every thread searches it five times)
        const int* src = reinterpret_cast<const int*>(ut);
        int* dst = reinterpret_cast<int*>(&sh.ut);
        for (int i = tid; i < (int)(sizeof(pgb_utable) / 4); i += TPB) dst[i] = src[i];
        __syncthreads();
      }
      EmitCtx<NX> ec;
      if (!last_only) emit_ctx_init<NX>(ec, fd, chain, t);
      P_FOR_BLOCKS for (int k = 0; k < tpc && !skip; ++k) {
        const int tile = c * tpc + k;
        if (tile >= nt) break;
        double W[4], q[5];
        if (!p_exact_prefixes(fd.main, fd.wbuf, S2, chain, c, k, nb, tpc, false, sh, W, q, fmain)) continue;
        if (last_only)
          p_consume<NX, MODE_INDEX>(fd, chain, tile, q, &sh.ut, pd.star_idx + chain, t, first, &fd.main.pend[chain], lmax,
                                    nullptr, nullptr);
        else
          p_consume_by_output<NX>(fd, ec, chain, tile, q, sh, t, first, &fd.main.pend[chain], lmax,
                                  pd.heavy + (int64_t)chain * P_HEAVY_MAX, pd.heavy_cnt + chain);
      }
      if (!last_only) block_max_to_global(lmax, &fd.maxkey[chain], sh.sm);
      P_TICK(7);
      if (anc) {
        const double S4 = block_tree_sum_array(pa2, nb, sh.sm);
        approx_prefix(pa2, nb, sh);
        const double ua = fd.utab_anc[(int64_t)chain * T + t].u0[0];  // one draw: u = rand()
        const double delta = 6.0 * ((double)N + 4096.0) * 0x1p-53;
        P_FOR_BLOCKS p_side_approx(fd.side, fd.waN, S4, sh.tsum_a2 + vi * tpc, chain, c, tpc, sh, ua, delta, pd.cand + 2 * chain);
      }
      P_TICK(8);
    }
    grid_barrier(pd.bar, bar_target);  // B3
    P_TICK(9);
    // ================================================================ phase D: side resolve + find, weights (optimistic)
    {
      if (anc) {
        // approximate shortcut decided?  (uniform over the grid: every CTA reads every chain's counter)
        if (tid == 0) {
          int amb = 0;
          for (int ch = 0; ch < fd.nc; ++ch) amb |= (pd.cand[2 * ch] == 0) || pd.force_fallback;
          sh.any_flag = amb;
        }
        __syncthreads();
        const bool amb_any = sh.any_flag != 0;
        const bool amb_me = (pd.cand[2 * chain] == 0) || pd.force_fallback;
        if (!amb_me && lead && tid == 0) fd.ah[((int64_t)chain * T + t) * N + (N - 1)] = pd.cand[2 * chain + 1];
        if (amb_any) {
          // rare: run the exact scan of the ancestor weights for the chains that need it (two extra barriers)
          const double S4 = block_tree_sum_array(pa2, nb, sh.sm);
          approx_prefix(pa2, nb, sh);
          if (amb_me) P_FOR_BLOCKS p_scan_local(fd.side, fd.waN, S4, sh.tsum_a2 + vi * tpc, chain, c, nb, tpc, sh, fd.status, fside);
          grid_barrier(pd.bar, bar_target);
          if (amb_me) {
            p_resolve(fd.side, chain, nb, tpc, sh, pd.force_fallback & 1, fside);
            if (tid == 0) {
              sh.sI = *fside & 1;
              if (lead) fd.side.counters[2 * chain + 1] += 1;
            }
            __syncthreads();
            const bool skip = sh.sI != 0;
            double dummy = 0.0;
            const pgb_utable* ut = &fd.utab_anc[(int64_t)chain * T + t];
            int32_t* out = fd.ah + ((int64_t)chain * T + t) * N + (N - 1);
            P_FOR_BLOCKS for (int k = 0; k < tpc && !skip; ++k) {
              const int tile = c * tpc + k;
              if (tile >= nt) break;
              double W[4], q[5];
              if (!p_exact_prefixes(fd.side, fd.waN, S4, chain, c, k, nb, tpc, false, sh, W, q, fside)) continue;
              p_consume<NX, MODE_INDEX>(fd, chain, tile, q, ut, out, t, first, &fd.side.pend[chain], dummy, nullptr, nullptr);
            }
          }
        }
      }
      P_TICK(10);
      // degenerate weights: a few particles own most offspring -> expand them with the whole chain, then one
      // extra (grid-uniform) barrier before the weight update can see every log-weight
      if (!last_only) {
        if (tid == 0) {
          int any = 0;
          for (int ch = 0; ch < fd.nc; ++ch) any |= pd.heavy_cnt[ch];
          sh.any_flag = any;
        }
        __syncthreads();
        if (sh.any_flag) {
          const int hc = pd.heavy_cnt[chain];
          double lmax = -INFINITY;
          if (hc > 0 && !(*fmain & 1))
            p_heavy<NX>(fd, chain, o, npc, t, pd.heavy + (int64_t)chain * P_HEAVY_MAX, hc, &fd.main.pend[chain], lmax);
          block_max_to_global(lmax, &fd.maxkey[chain], sh.sm);
          grid_barrier(pd.bar, bar_target);
        }
      }
      if (!last_only) weights_phase();
      // re-arm the other parity's flags for the next step (nobody touches them between B1(t) and B1(t+1))
      if (lead && tid == 0) {
        fd.main.flag[chain + fd.nc * (int)((t + 1) & 1)] = 0;
        fd.side.flag[chain + fd.nc * (int)((t + 1) & 1)] = 0;
      }
    }
    P_TICK(11);
    grid_barrier(pd.bar, bar_target);  // B4 (all consumers done; this step's flags are final)
    P_TICK(12);
    // ---------------------------------------------------------------- fix-up: rare, uniform across the grid
    {
      if (tid == 0) {
        int any = 0;
        for (int ch = 0; ch < fd.nc; ++ch)
          any |= (fd.main.flag[ch + fd.nc * (int)(t & 1)] | fd.side.flag[ch + fd.nc * (int)(t & 1)]) & 1;
        sh.any_flag = any;
      }
      __syncthreads();
      if (sh.any_flag) {
        const int fm = *fmain & 1, fs = anc ? (*fside & 1) : 0;
        const double S4 = (anc && fs) ? block_tree_sum_array(pa2, nb, sh.sm) : 1.0;
        if (lead) {
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
        P_FOR_BLOCKS for (int k = 0; k < tpc; ++k) {
          const int tile = c * tpc + k;
          if (tile >= nt) break;
          double W[4], q[5];
          if (fm) {
            p_exact_prefixes(fd.main, fd.wbuf, S2, chain, c, k, nb, tpc, true, sh, W, q, fmain);
            const pgb_utable* ut = last_only ? &fd.utab_star[chain] : &fd.utab_res[(int64_t)chain * T + t];
            if (last_only)
              p_consume<NX, MODE_INDEX>(fd, chain, tile, q, ut, pd.star_idx + chain, t, first, &fd.status[chain], lmax,
                                        nullptr, nullptr);
            else
              p_consume<NX, MODE_PROPAGATE>(fd, chain, tile, q, ut, nullptr, t, first, &fd.status[chain], lmax, nullptr,
                                            nullptr);
          }
          if (fs) {
            double dummy = 0.0;
            p_exact_prefixes(fd.side, fd.waN, S4, chain, c, k, nb, tpc, true, sh, W, q, fside);
            p_consume<NX, MODE_INDEX>(fd, chain, tile, q, &fd.utab_anc[(int64_t)chain * T + t],
                                      fd.ah + ((int64_t)chain * T + t) * N + (N - 1), t, first, &fd.status[chain], dummy,
                                      nullptr, nullptr);
          }
        }
        if (fm && !last_only) block_max_to_global(lmax, &fd.maxkey[chain], sh.sm);
        grid_barrier(pd.bar, bar_target);
        if (!last_only) {  // the optimistic weight update used log-weights that have just been redone
          if (fm) weights_phase();
          grid_barrier(pd.bar, bar_target);
        }
      }
      // merge the status bits raised by consumer passes that were not redone
      if (lead && tid == 0) {
        int p = fd.main.pend[chain] | fd.side.pend[chain];
        if (p) atomicOr(&fd.status[chain], p);
        fd.main.pend[chain] = 0;
        fd.side.pend[chain] = 0;
      }
    }
    P_TICK(13);
  }
#ifdef PGB_EXPERIMENT
  if (pd.prof && tid == 0) {
#pragma unroll
    for (int i = 0; i < P_NPROF; ++i) pd.prof[(int64_t)blockIdx.x * P_NPROF + i] = tacc[i];
  }
#endif
#undef P_TICK
#undef P_FOR_BLOCKS
}

}  // namespace pgb
