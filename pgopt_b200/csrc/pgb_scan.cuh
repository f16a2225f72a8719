// pgb_scan.cuh — the exact sequential-order prefix scan (systematic_resampling's q = q + W[n],
// Julia/src/particle_Gibbs.jl:23-32) as a parallel pipeline.  See pgb_exact.h for the idea.
//
//   k_scan_local   per tile: W = v/S, approximate prefix -> binade prediction -> per-thread maps,
//                  segmented composition scan inside the tile, records for "complex" threads
//   k_resolve      one CTA per chain: scan over tile aggregates, short serial chain over the
//                  complex threads with real FP adds -> seeds
//   k_expand       per tile: map-derived incoming value per thread, true FP walk of its 4 elements,
//                  seam check, then the consumer (offspring expansion + propagation, or index only)
//   k_seq_exact    serial exact walk, runs only when a seam check failed (flag), then k_expand
//                  is repeated in fallback mode.
#pragma once
#include "pgb_device.cuh"

namespace pgb {

struct __align__(16) ScanRec {
  pgb_agg ex;    // exclusive aggregate of the complex thread (tile-local in recs, global in crec)
  double W[4];   // its four normalised weights
  double pad_;
};

// Workspace of one scan pipeline; every array has a leading chain dimension.
struct ScanWS {
  int64_t N;          // elements per chain
  int nt;             // tiles per chain
  double* S;          // [nc]            normaliser: tree sum of v
  double* tsum;       // [nc][nt]        tile tree sums of v
  double* tile_prefix;// [nc][nt+1]      approximate exclusive prefix of tsum
  pgb_agg* thr_ex;    // [nc][nt*TPB]    tile-local exclusive aggregate per thread
  pgb_agg* tile_agg;  // [nc][nt]
  pgb_agg* tile_ex;   // [nc][nt+1]      exclusive scan of tile_agg (entry nt = total)
  ScanRec* recs;      // [nc][nt][RMAX]
  ScanRec* crec;      // [nc][CMAX]      compacted records in global order
  double* seeds;      // [nc][CMAX+1]
  double* qin_fb;     // [nc][nt*TPB]    exact incoming values written by the serial fallback
  double* Wn;         // [nc][nt*TILE]   normalised weights W = v/S kept between scan-local and consumer (persistent path)
  unsigned int* ticket; // [nc]
  int* flag;          // [nc]  bit0: redo this scan with the serial exact walk
  int* pend;          // [nc]  status bits raised by the normal consumer pass (merged or discarded)
  long long* counters;// [nc][2]  {fallbacks, scans}
};

// Called by every thread of the LAST tile-CTA of a producer kernel: total (tree) + approximate
// exclusive prefix of the tile sums.
__device__ __forceinline__ void finalize_tile_sums(const double* __restrict__ tsum, int nt, double* S_out,
                                                   double* __restrict__ tile_prefix, double* sm) {
  double S = block_tree_sum_array(tsum, nt, sm);
  if (threadIdx.x == 0) *S_out = S;
  double carry = 0.0;
  for (int base = 0; base < nt; base += TPB) {
    int i = base + threadIdx.x;
    double v = (i < nt) ? tsum[i] : 0.0;
    double tot;
    double ex = block_excl_scan_approx(v, sm, &tot);
    if (i < nt) tile_prefix[i] = __dadd_rn(carry, ex);
    carry = __dadd_rn(carry, tot);
  }
  if (threadIdx.x == 0) tile_prefix[nt] = carry;
}

// Stand-alone tile-sum kernel (used by the stand-alone resampling entry point and the star draw).
static __global__ void __launch_bounds__(TPB) k_tile_sums(ScanWS ws, const double* __restrict__ v) {
  __shared__ double sm[NWARP + 1];
  __shared__ int sflag;
  const int chain = blockIdx.y, tile = blockIdx.x;
  const double* vc = v + (int64_t)chain * ws.N;
  int64_t base = (int64_t)tile * TILE + (int64_t)threadIdx.x * IPT;
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = (base + i < ws.N) ? vc[base + i] : 0.0;
  double s = block_tree_sum(tree4(x), sm);
  if (threadIdx.x == 0) ws.tsum[(int64_t)chain * ws.nt + tile] = s;
  if (last_block_arrives(&ws.ticket[chain], gridDim.x, &sflag)) {
    finalize_tile_sums(ws.tsum + (int64_t)chain * ws.nt, ws.nt, &ws.S[chain],
                       ws.tile_prefix + (int64_t)chain * (ws.nt + 1), sm);
    if (threadIdx.x == 0) ws.flag[chain] = 0;
  }
}

static __global__ void __launch_bounds__(TPB) k_scan_local(ScanWS ws, const double* __restrict__ v, int* __restrict__ status) {
  __shared__ double sm[NWARP + 1];
  __shared__ pgb_agg smagg[NWARP + 1];
  __shared__ double sP[TPB];
  const int chain = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const double* vc = v + (int64_t)chain * ws.N;
  const double S = ws.S[chain];
  int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
  double W[4];
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    W[i] = (base + i < ws.N) ? __ddiv_rn(vc[base + i], S) : 0.0;
    if (!(W[i] >= 0.0 && W[i] <= 1.7976931348623157e308)) bad = true;
  }
  if (bad) atomicOr(&status[chain], 4 /*PGB_STATUS_NONFINITE*/);
  double s1 = W[0], s2 = __dadd_rn(s1, W[1]), s3 = __dadd_rn(s2, W[2]), s4 = __dadd_rn(s3, W[3]);
  double tot;
  double ex = block_excl_scan_approx(s4, sm, &tot);
  const double* tp = ws.tile_prefix + (int64_t)chain * (ws.nt + 1);
  double Pc = __ddiv_rn(tp[tile], S), Pn = __ddiv_rn(tp[tile + 1], S);
  double P = __dadd_rn(Pc, ex);
  sP[tid] = P;
  __syncthreads();
  double qt[5];
  qt[0] = P;
  qt[1] = __dadd_rn(P, s1);
  qt[2] = __dadd_rn(P, s2);
  qt[3] = __dadd_rn(P, s3);
  qt[4] = (tid + 1 < TPB) ? sP[tid + 1] : Pn;
  pgb_agg agg = bad ? pgb_agg_complex() : pgb_classify4(W, qt);
  pgb_agg total;
  pgb_agg exagg = block_excl_scan_agg(agg, smagg, &total);
  ws.thr_ex[((int64_t)chain * ws.nt + tile) * TPB + tid] = exagg;
  if (agg.cnt) {
    if (exagg.cnt < RMAX) {
      ScanRec r;
      r.ex = exagg;
      r.W[0] = W[0]; r.W[1] = W[1]; r.W[2] = W[2]; r.W[3] = W[3];
      r.pad_ = 0.0;
      ws.recs[((int64_t)chain * ws.nt + tile) * RMAX + exagg.cnt] = r;
    } else {
      atomicOr(&ws.flag[chain], 1);
    }
  }
  if (tid == 0) ws.tile_agg[(int64_t)chain * ws.nt + tile] = total;
}

constexpr int RESOLVE_SMEM_RECS = 512;

static __global__ void __launch_bounds__(TPB) k_resolve(ScanWS ws, int force_fallback) {
  __shared__ pgb_agg smagg[NWARP + 1];
  __shared__ pgb_map s_f[RESOLVE_SMEM_RECS];
  __shared__ double s_W[RESOLVE_SMEM_RECS][4];
  const int chain = blockIdx.y, tid = threadIdx.x;
  const int nt = ws.nt;
  const pgb_agg* tagg = ws.tile_agg + (int64_t)chain * nt;
  pgb_agg* tex = ws.tile_ex + (int64_t)chain * (nt + 1);
  pgb_agg carry = pgb_agg_identity();
  for (int base = 0; base < nt; base += TPB) {
    int i = base + tid;
    pgb_agg v = (i < nt) ? tagg[i] : pgb_agg_identity();
    pgb_agg tot;
    pgb_agg ex = block_excl_scan_agg(v, smagg, &tot);
    if (i < nt) tex[i] = pgb_agg_combine(carry, ex);
    carry = pgb_agg_combine(carry, tot);
  }
  if (tid == 0) {
    tex[nt] = carry;
    ws.counters[2 * chain + 1] += 1;
  }
  const int C = carry.cnt;
  if (C > CMAX || force_fallback) {
    if (tid == 0) atomicOr(&ws.flag[chain], 1);
    return;
  }
  __syncthreads();  // tile_ex written by this CTA is visible to it
  ScanRec* crec = ws.crec + (int64_t)chain * CMAX;
  for (int tile = tid; tile < nt; tile += TPB) {
    int c = tagg[tile].cnt;
    if (c > RMAX) c = RMAX;
    pgb_agg te = tex[tile];
    for (int i = 0; i < c; ++i) {
      ScanRec r = ws.recs[((int64_t)chain * nt + tile) * RMAX + i];
      int g = te.cnt + i;
      if (g < CMAX) {
        r.ex = pgb_agg_combine(te, r.ex);
        crec[g] = r;
        if (g < RESOLVE_SMEM_RECS) {
          s_f[g] = r.ex.f;
          s_W[g][0] = r.W[0]; s_W[g][1] = r.W[1]; s_W[g][2] = r.W[2]; s_W[g][3] = r.W[3];
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    double* seeds = ws.seeds + (int64_t)chain * (CMAX + 1);
    double q = 0.0;
    seeds[0] = q;
    for (int g = 0; g < C; ++g) {
      pgb_map f;
      double w0, w1, w2, w3;
      if (g < RESOLVE_SMEM_RECS) {
        f = s_f[g];
        w0 = s_W[g][0]; w1 = s_W[g][1]; w2 = s_W[g][2]; w3 = s_W[g][3];
      } else {
        f = crec[g].ex.f;
        w0 = crec[g].W[0]; w1 = crec[g].W[1]; w2 = crec[g].W[2]; w3 = crec[g].W[3];
      }
      double qb = pgb_from_mant(pgb_segkey(q), pgb_map_apply(f, pgb_mant(q)));
      qb = __dadd_rn(qb, w0);
      qb = __dadd_rn(qb, w1);
      qb = __dadd_rn(qb, w2);
      qb = __dadd_rn(qb, w3);
      q = qb;
      seeds[g + 1] = q;
    }
  }
}

// Serial exact walk (one warp per chain).  Runs only if the chain's flag is raised.
// Normal case (flag clear): merge the consumer's pending status bits and return.  Flag raised:
// discard what the failed pass produced (pending status, running max) and walk serially.
static __global__ void __launch_bounds__(32) k_seq_exact(ScanWS ws, const double* __restrict__ v, int* __restrict__ status,
                                                  unsigned long long* __restrict__ maxkey_reset) {
  const int chain = blockIdx.y, lane = threadIdx.x;
  if ((ws.flag[chain] & 1) == 0) {
    if (lane == 0) {
      int p = ws.pend[chain];
      if (p) atomicOr(&status[chain], p);
      ws.pend[chain] = 0;
    }
    return;
  }
  if (lane == 0) {
    ws.pend[chain] = 0;
    if (maxkey_reset) maxkey_reset[chain] = 0ull;
  }
  const double* vc = v + (int64_t)chain * ws.N;
  const double S = ws.S[chain];
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

// Per-thread exact prefixes q[0..4] (q[0] = value before the thread's first element).
// Returns false (uniformly for the CTA) if this launch has nothing to do for the chain: the normal
// pass skips flagged chains, the fallback pass skips unflagged ones, and a CTA whose own seam
// check failed skips its consumer (the fallback pass redoes it).  sQ: TPB doubles, sI: 1 int.
__device__ __forceinline__ bool exact_prefixes(const ScanWS& ws, const double* __restrict__ v, int chain, int tile,
                                               bool fallback_pass, double W[4], double q[5], double* sQ, int* sI) {
  const int tid = threadIdx.x;
  if (tid == 0) *sI = ws.flag[chain] & 1;
  __syncthreads();
  const int flag = *sI;
  if (fallback_pass != (flag != 0)) return false;
  const double* vc = v + (int64_t)chain * ws.N;
  const double S = ws.S[chain];
  int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
#pragma unroll
  for (int i = 0; i < 4; ++i) W[i] = (base + i < ws.N) ? __ddiv_rn(vc[base + i], S) : 0.0;
  const int64_t gthr = ((int64_t)chain * ws.nt + tile) * TPB + tid;
  double qin;
  if (fallback_pass) {
    qin = ws.qin_fb[gthr];
  } else {
    const pgb_agg* tex = ws.tile_ex + (int64_t)chain * (ws.nt + 1);
    const double* seeds = ws.seeds + (int64_t)chain * (CMAX + 1);
    pgb_agg ex = pgb_agg_combine(tex[tile], ws.thr_ex[gthr]);
    int c = ex.cnt < 0 ? 0 : (ex.cnt > CMAX ? CMAX : ex.cnt);
    qin = pgb_qin_from(ex, seeds[c]);
  }
  q[0] = qin;
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i + 1] = __dadd_rn(q[i], W[i]);
  if (!fallback_pass) {
    // seam check: my outgoing value must equal my right neighbour's incoming value, bitwise
    sQ[tid] = qin;
    __syncthreads();
    double nxt;
    bool have = (base + IPT < ws.N);  // seams past the last real element cannot matter
    if (!have) {
      nxt = 0.0;
    } else if (tid + 1 < TPB) {
      nxt = sQ[tid + 1];
    } else if (tile + 1 < ws.nt) {
      const pgb_agg* tex = ws.tile_ex + (int64_t)chain * (ws.nt + 1);
      const double* seeds = ws.seeds + (int64_t)chain * (CMAX + 1);
      pgb_agg ex = pgb_agg_combine(tex[tile + 1], ws.thr_ex[gthr + 1]);
      int c = ex.cnt < 0 ? 0 : (ex.cnt > CMAX ? CMAX : ex.cnt);
      nxt = pgb_qin_from(ex, seeds[c]);
    } else {
      have = false;
      nxt = 0.0;
    }
    int mism = have && (__double_as_longlong(nxt) != __double_as_longlong(q[4]));
    if (mism) atomicOr(&ws.flag[chain], 1);
    if (__syncthreads_or(mism)) return false;
  }
  return true;
}

}  // namespace pgb
