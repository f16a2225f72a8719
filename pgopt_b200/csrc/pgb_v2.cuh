// pgb_v2.cuh — the CPF-AS time loop (Julia/src/particle_Gibbs.jl:160-200) as ONE persistent cooperative kernel,
// second generation: "chunk-stationary".
//
// What changed against the tile kernel of round 1 (pgb_persist.cuh), and why (ncu of that kernel: issue
// slots 30 % busy; stalls: barrier 2.7, no-instruction 2.3, long scoreboard 2.0, wait 2.7 cycles per issue):
//   * One 896-thread CTA per SM (28 warps, <= 72 registers) instead of 2 x 256 threads x 128 registers: latency is
//     hidden by 7 warps per scheduler running ONE element at a time, not by 4-wide interleaved math, so the hot
//     code of a phase is a few KB (the 32 KB instruction cache holds it) and the kernel is specialised at compile time
//     on the basis family and on first / conditional sweep.
//   * A CTA owns a CONTIGUOUS, tile-aligned range of the chain's particles for the whole launch; between the grid
//     barriers of a step its e -> w -> W, pdf -> waN -> waN', the per-lane scan aggregates and offspring counts stay in
//     shared memory ("stationary": 22 B per particle, 7 tiles per SM at N = 2^20), so the only global traffic of a
//     step is lw, x (read), F (write + gather), a, x, lw (write) — and nothing rolls through L2 into DRAM.
//   * The unit of work is a 128-particle CHUNK handled by one warp with shuffles only; a phase has one or two
//     __syncthreads in total instead of ~10 per 1024-particle tile.  Heavy per-particle arithmetic (basis functions,
//     exp, Philox / Box-Muller) uses the strided mapping lane = particle mod 32 (coalesced scalar accesses); the
//     sum / scan parts use the mapping lane = 4 consecutive particles, which is the leaf order of the oracle's
//     balanced summation tree.
//   * Offspring are expanded BY OUTPUT over the whole CTA: ancestors are scattered into a CTA-wide staging buffer,
//     then every thread propagates output slots — balanced inside the SM whatever the weight distribution.
// The arithmetic — every sum tree, every division, the exact sequential-order scan with its seam check and serial
// fallback, the certified shortcut of the ancestor draw — is that of pgb_persist.cuh / the oracle, bit for bit.
//
// Per step t (conditional sweep):                                            grid barriers: 4
//   PA  e = exp(lw - max) (:198), F = A*phi(x_{t-1}), pdf = N(x'_t; F, Q); tile sums of e          | B1
//   PB  w = e/S1 (:199), waN = w*pdf (:179); tile sums of w, waN                                    | B2
//   PC  W = w/S2 (:19), binade prediction, integer maps, composition scan per chunk / CTA;
//       ancestor draw of slot N (:181) by the certified approximate scan of the raw waN             | B3
//   PD  resolve -> exact prefixes, seam check -> offspring counts -> by-output expansion:
//       a[t,:], x_t = F[a] + L z (:175|:188), lw (:194), max                                        | B0
//   (rare, grid-uniform: waN ./= S3 (:180) + exact scan of the ancestor weights when the shortcut is undecided;
//    serial exact walk after a failed seam check)
#pragma once
#include "pgb_real.cuh"
#include "pgb_v2_types.h"

namespace pgb {

// fixed part of the shared memory; the dynamic arrays follow (see v2_smem_layout)
struct V2Shared {
  double red[2][8], redq[2][8], redlo[2][8], redhi[2][8];
  double redm[32];
  pgb_agg cex[V2_MAXNPC + 1];   // exclusive scan of the chain's CTA aggregates
  pgb_map rf[V2_MAXREC];
  double rW[V2_MAXREC][4];
  double seeds[V2_MAXREC + 1];
  pgb_utable ut;
  double tp_lo, tp_hi;          // approximate raw prefix of the tile sums at the CTA's first / past-the-last tile
  double tq_lo;                 // the same for the second array of v2_tiles_collective2 (first tile only)
  double redqlo[2][8];
  long long J0, J1;
  int flag_sh, any_sh, nrec, pad_;
};
struct V2PoolRec {  // a complex lane's record before its position in the CTA's list is known
  ScanRec rec;
  int cc, k, pad0_, pad1_;
};
static_assert(sizeof(V2PoolRec) * V2_RCTA <= (size_t)V2_CAP * 4, "record pool must fit the staging buffer");

struct V2Layout {
  size_t chs, chpfx, chagg, chex, A, sanc, E, P, TE, total;
};
__host__ __device__ inline V2Layout v2_smem_layout(int mch, int n_A, int stationary) {
  V2Layout L;
  size_t o = (sizeof(V2Shared) + 15) & ~(size_t)15;
  L.chs = o;   o += (size_t)4 * mch * 8;
  L.chpfx = o; o += (size_t)2 * (mch + 1) * 8;
  L.chagg = o; o += (size_t)mch * sizeof(pgb_agg);
  L.chex = o;  o += (size_t)(mch + 1) * sizeof(pgb_agg);
  o = (o + 15) & ~(size_t)15;
  L.A = o;     o += (size_t)((n_A + 1) & ~1) * 8;
  L.sanc = o;  o += (size_t)V2_CAP * 4;  // >= V2_RCTA * sizeof(PoolRec): the record pool of scan-local aliases the staging buffer
  L.E = o;
  if (stationary) o += (size_t)mch * V2_CHUNK * 8;
  L.P = o;
  if (stationary) o += (size_t)mch * V2_CHUNK * 8;
  L.TE = o;
  if (stationary) o += (size_t)mch * 32 * sizeof(pgb_agg);
  L.total = o;
  return L;
}

// Grid barrier of the co-resident CTAs: bar.sync makes the CTA's writes happen-before thread 0's RELEASE increment
// (cumulativity), the ACQUIRE spin + bar.sync make the other CTAs' writes visible to every thread afterwards.
__device__ __forceinline__ void v2_grid_barrier(unsigned int* counter, unsigned int& target) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    while (ld_acquire_u32(counter) < target) {
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------- CTA collectives
// Balanced tree over PER consecutive entries of p starting at base (PER a power of two, entries >= n read as +0.0),
// fully unrolled: no run-time stack.  *lo / *hi accumulate (left to right) the entries with index < at_lo / < at_hi.
template <int PER>
__device__ __forceinline__ double v2_tree_load(const double* __restrict__ p, int base, int n, int at_lo, int at_hi,
                                               double& lo, double& hi) {
  if (PER == 1) {
    const double v = (base < n) ? __ldcg(p + base) : 0.0;
    if (base < at_lo) lo = __dadd_rn(lo, v);
    if (base < at_hi) hi = __dadd_rn(hi, v);
    return v;
  }
  const double a = v2_tree_load<(PER > 1 ? PER / 2 : 1)>(p, base, n, at_lo, at_hi, lo, hi);
  const double b = v2_tree_load<(PER > 1 ? PER / 2 : 1)>(p, base + PER / 2, n, at_lo, at_hi, lo, hi);
  return __dadd_rn(a, b);
}
// CTA collective over the tile sums p[0..n), n <= 4096: returns the balanced-tree sum (the oracle's `sum`) to every
// thread and, when at_hi >= 0, leaves in sh.tp_lo / sh.tp_hi an APPROXIMATE raw prefix of p at tile indices at_lo / at_hi
// (a fixed function of (p, index), so neighbouring CTAs agree bitwise on their common boundary).  One __syncthreads;
// the scratch is double-buffered by a parity all threads advance together.
__device__ __forceinline__ double v2_tiles_collective(const double* __restrict__ p, int n, int at_lo, int at_hi,
                                                      V2Shared& sh, int& par) {
  const int tid = threadIdx.x;
  if (tid < 256) {
    double lo = 0.0, hi = 0.0, v;
    if (n <= 256) v = v2_tree_load<1>(p, tid, n, at_lo, at_hi, lo, hi);
    else if (n <= 512) v = v2_tree_load<2>(p, tid * 2, n, at_lo, at_hi, lo, hi);
    else if (n <= 1024) v = v2_tree_load<4>(p, tid * 4, n, at_lo, at_hi, lo, hi);
    else if (n <= 2048) v = v2_tree_load<8>(p, tid * 8, n, at_lo, at_hi, lo, hi);
    else v = v2_tree_load<16>(p, tid * 16, n, at_lo, at_hi, lo, hi);
    v = warp_tree_sum(v);
    if (at_hi >= 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = __dadd_rn(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = __dadd_rn(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
    }
    if ((tid & 31) == 0) {
      sh.red[par][tid >> 5] = v;
      sh.redlo[par][tid >> 5] = lo;
      sh.redhi[par][tid >> 5] = hi;
    }
  }
  __syncthreads();
  const double* s = sh.red[par];
  if (at_hi >= 0 && tid == 0) {
    double lo = 0.0, hi = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      lo = __dadd_rn(lo, sh.redlo[par][i]);
      hi = __dadd_rn(hi, sh.redhi[par][i]);
    }
    sh.tp_lo = lo;
    sh.tp_hi = hi;
  }
  par ^= 1;
  return __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[1]), __dadd_rn(s[2], s[3])),
                   __dadd_rn(__dadd_rn(s[4], s[5]), __dadd_rn(s[6], s[7])));
}
// The same for two arrays at once (S2 and S3 of PC): prefixes are taken of p only.  *Sq = tree sum of q.
__device__ __forceinline__ double v2_tiles_collective2(const double* __restrict__ p, const double* __restrict__ q, int n,
                                                       int at_lo, int at_hi, V2Shared& sh, int& par, double* Sq) {
  const int tid = threadIdx.x;
  if (tid < 256) {
    double lo = 0.0, hi = 0.0, d0 = 0.0, d1 = 0.0, v, w;  // d0: raw prefix of q at at_lo
    if (n <= 256) { v = v2_tree_load<1>(p, tid, n, at_lo, at_hi, lo, hi); w = v2_tree_load<1>(q, tid, n, at_lo, 0, d0, d1); }
    else if (n <= 512) { v = v2_tree_load<2>(p, tid * 2, n, at_lo, at_hi, lo, hi); w = v2_tree_load<2>(q, tid * 2, n, at_lo, 0, d0, d1); }
    else if (n <= 1024) { v = v2_tree_load<4>(p, tid * 4, n, at_lo, at_hi, lo, hi); w = v2_tree_load<4>(q, tid * 4, n, at_lo, 0, d0, d1); }
    else if (n <= 2048) { v = v2_tree_load<8>(p, tid * 8, n, at_lo, at_hi, lo, hi); w = v2_tree_load<8>(q, tid * 8, n, at_lo, 0, d0, d1); }
    else { v = v2_tree_load<16>(p, tid * 16, n, at_lo, at_hi, lo, hi); w = v2_tree_load<16>(q, tid * 16, n, at_lo, 0, d0, d1); }
    v = warp_tree_sum(v);
    w = warp_tree_sum(w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = __dadd_rn(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = __dadd_rn(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      d0 = __dadd_rn(d0, __shfl_xor_sync(0xffffffffu, d0, o));
    }
    if ((tid & 31) == 0) {
      sh.red[par][tid >> 5] = v;
      sh.redq[par][tid >> 5] = w;
      sh.redlo[par][tid >> 5] = lo;
      sh.redhi[par][tid >> 5] = hi;
      sh.redqlo[par][tid >> 5] = d0;
    }
  }
  __syncthreads();
  const double* s = sh.red[par];
  const double* sq = sh.redq[par];
  if (tid == 0) {
    double lo = 0.0, hi = 0.0, ql = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      lo = __dadd_rn(lo, sh.redlo[par][i]);
      hi = __dadd_rn(hi, sh.redhi[par][i]);
      ql = __dadd_rn(ql, sh.redqlo[par][i]);
    }
    sh.tp_lo = lo;
    sh.tp_hi = hi;
    sh.tq_lo = ql;
  }
  par ^= 1;
  *Sq = __dadd_rn(__dadd_rn(__dadd_rn(sq[0], sq[1]), __dadd_rn(sq[2], sq[3])),
                  __dadd_rn(__dadd_rn(sq[4], sq[5]), __dadd_rn(sq[6], sq[7])));
  return __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[1]), __dadd_rn(s[2], s[3])),
                   __dadd_rn(__dadd_rn(s[4], s[5]), __dadd_rn(s[6], s[7])));
}
// tree over the 8 chunk sums of one tile
__device__ __forceinline__ double v2_tile_from_chunks(const double* c) {
  return __dadd_rn(__dadd_rn(__dadd_rn(c[0], c[1]), __dadd_rn(c[2], c[3])),
                   __dadd_rn(__dadd_rn(c[4], c[5]), __dadd_rn(c[6], c[7])));
}

__device__ __forceinline__ pgb_agg v2_shfl_agg(pgb_agg a, int src) {
  pgb_agg r;
  r.f.c0 = __shfl_sync(0xffffffffu, (long long)a.f.c0, src);
  r.f.c1 = __shfl_sync(0xffffffffu, (long long)a.f.c1, src);
  r.cnt = __shfl_sync(0xffffffffu, a.cnt, src);
  r.key = __shfl_sync(0xffffffffu, a.key, src);
  return r;
}
// exclusive composition scan over the lanes of a warp; *total = combination of all 32
__device__ __forceinline__ pgb_agg v2_warp_excl_scan_agg(pgb_agg v, pgb_agg* total) {
  const int lane = threadIdx.x & 31;
  pgb_agg inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    pgb_agg t = shfl_up_agg(inc, o);
    if (lane >= o) inc = pgb_agg_combine(t, inc);
  }
  *total = v2_shfl_agg(inc, 31);
  pgb_agg ex = shfl_up_agg(inc, 1);
  if (lane == 0) ex = pgb_agg_identity();
  return ex;
}
__device__ __forceinline__ double v2_warp_incl_scan(double v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = __dadd_rn(t, v);
  }
  return v;
}

// ---------------------------------------------------------------------------------------------------- model pieces
// The model matrix A of a single-chain handle in constant memory (copied device-to-device before the launch): the
// 250 fused multiply-adds of the 5x5x5 Hilbert basis then take A as a constant-bank OPERAND instead of 125 shared-memory
// loads per particle (ncu of the first version: the contraction line stalled on the short scoreboard, i.e. on LDS).
constexpr int V2_CA_MAX = 2560;
static __constant__ double c_v2A[V2_CA_MAX];
static __constant__ float c_v2Af[V2_CA_MAX];  // the same rounded to FP32 (F32 precision mode)
template <typename RT> __device__ __forceinline__ const RT* v2_const_A();
template <> __device__ __forceinline__ const double* v2_const_A<double>() { return c_v2A; }
template <> __device__ __forceinline__ const float* v2_const_A<float>() { return c_v2Af; }
// FP64 -> FP32 staging of A for the constant bank
static __global__ void k_round_to_f32(const double* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

// ---------------------------------------------------------------------------------------------------- the exact scan
// One exact-order scan over the CTA's particles.  v: stationary array holding the numerators on entry of `local`
// (overwritten with W = v/S); TE: per-lane exclusive aggregates (later: offspring counts); which = 0 main, 1 side.
struct V2Scan {
  double* v;        // CTA-local values (shared or global), index = particle - p0
  pgb_agg* TE;      // CTA-local per-lane slots, index = chunk*32 + lane
  double* chsum;    // chunk sums of the numerators (raw), [nch]
  double* chpfx;    // exclusive raw prefix of chsum, [nch+1]
  pgb_agg* chagg;   // [nch]
  pgb_agg* chex;    // [nch+1]
  V2PoolRec* srec;  // [V2_RCTA] pool (aliases the staging buffer)
  pgb_agg* cta_agg; // global [npc] of this chain / scan
  ScanRec* recs;    // global [npc][V2_RCTA]
  int* flagp;       // global flag of this chain / scan / step parity
  int* status;      // global status word of the chain
  int64_t N;
  int p0, nch, r, npc, nt;
};

static __device__ __noinline__ pgb_agg v2_classify4_slow(const double (&W)[4], const double (&qt)[5]) {
  return pgb_classify4(W, qt);
}

// PC: W = v/S, prediction, maps, chunk-level composition scan.  Ends with the CTA aggregate and the CTA's complex-lane
// records in global memory.  Contains __syncthreads.
__device__ __forceinline__ void v2_scan_local(const V2Scan& sc, double S, V2Shared& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const InvDiv ivS = make_invdiv(S);
  for (int cc = warp; cc < sc.nch; cc += V2_NW) {
    const int loc = cc * V2_CHUNK + lane * 4;
    const int n0 = sc.p0 + loc;
    double W[4];
    bool bad = false;
    {
      const double2 a = *reinterpret_cast<const double2*>(sc.v + loc), b = *reinterpret_cast<const double2*>(sc.v + loc + 2);
      const double v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        W[i] = (n0 + i < sc.N) ? div_inv(v[i], ivS) : 0.0;
        if (!(W[i] >= 0.0 && W[i] <= 1.7976931348623157e308)) bad = true;
      }
      *reinterpret_cast<double2*>(sc.v + loc) = make_double2(W[0], W[1]);
      *reinterpret_cast<double2*>(sc.v + loc + 2) = make_double2(W[2], W[3]);
    }
    if (bad) atomicOr(sc.status, 4);
    const double s1 = W[0], s2 = __dadd_rn(s1, W[1]), s3 = __dadd_rn(s2, W[2]), s4 = __dadd_rn(s3, W[3]);
    const double inc = v2_warp_incl_scan(s4);
    const double Pc = __dmul_rn(__dadd_rn(sh.tp_lo, sc.chpfx[cc]), ivS.y);  // approximate quantities only
    const double P = __dadd_rn(Pc, __dsub_rn(inc, s4));
    double nxt = __shfl_down_sync(0xffffffffu, P, 1);
    if (lane == 31)
      nxt = (cc + 1 < sc.nch) ? __dmul_rn(__dadd_rn(sh.tp_lo, sc.chpfx[cc + 1]), ivS.y) : __dmul_rn(sh.tp_hi, ivS.y);
    const double qt[5] = {P, __dadd_rn(P, s1), __dadd_rn(P, s2), __dadd_rn(P, s3), nxt};
    // Common case (all but ~log2 N lanes of a step): the lane's four additions are predicted to stay inside ONE binade,
    // away from its ends.  Its composed map is then obtained by simply performing the four additions on the binade's
    // first even and first odd grid point (a map depends on its argument through the parity only): 8 real FP adds
    // instead of four shift-and-round constructions and three compositions.  Everything else takes pgb_classify4.  As
    // always the prediction is only a prediction: the seam check in PD proves or refutes it.
    pgb_agg agg;
    const int key = pgb_segkey(qt[0]);
    if (!bad && key > 1 && key == pgb_segkey(qt[4]) && !pgb_near_pow2(qt[0]) && !pgb_near_pow2(qt[4])) {
      const double Qe = __longlong_as_double((long long)key << 52);            // 2^(key-1023): mantissa 2^52, even
      const double Qo = __longlong_as_double(((long long)key << 52) + 1);      // next grid point: odd
      const double re = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(Qe, W[0]), W[1]), W[2]), W[3]);
      const double ro = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(Qo, W[0]), W[1]), W[2]), W[3]);
      agg.f.c0 = __double_as_longlong(re) - __double_as_longlong(Qe);
      agg.f.c1 = __double_as_longlong(ro) - __double_as_longlong(Qo);
      agg.cnt = 0;
      agg.key = (agg.f.c0 | agg.f.c1) ? key : 0;
    } else {
      agg = bad ? pgb_agg_complex() : v2_classify4_slow(W, qt);
    }
    pgb_agg total;
    const pgb_agg ex = v2_warp_excl_scan_agg(agg, &total);
    sc.TE[cc * 32 + lane] = ex;
    if (lane == 0) sc.chagg[cc] = total;
    if (agg.cnt) {  // complex lane (a handful per step): stash its record; its place in the list is known after the scan
      const int slot = atomicAdd(&sh.nrec, 1);
      if (slot < V2_RCTA) {
        V2PoolRec pr;
        pr.rec.ex = ex;
        pr.rec.W[0] = W[0]; pr.rec.W[1] = W[1]; pr.rec.W[2] = W[2]; pr.rec.W[3] = W[3];
        pr.rec.pad_ = 0.0;
        pr.cc = cc; pr.k = ex.cnt; pr.pad0_ = 0; pr.pad1_ = 0;
        sc.srec[slot] = pr;
      } else {
        atomicOr(sc.flagp, 1);
      }
    }
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the chunk aggregates, CTA aggregate
    pgb_agg carry = pgb_agg_identity();
    for (int base = 0; base < sc.nch; base += 32) {
      const int cc = base + lane;
      const pgb_agg a = (cc < sc.nch) ? sc.chagg[cc] : pgb_agg_identity();
      pgb_agg tot;
      pgb_agg ex = pgb_agg_combine(carry, v2_warp_excl_scan_agg(a, &tot));
      if (cc < sc.nch) sc.chex[cc] = ex;
      carry = pgb_agg_combine(carry, tot);
    }
    if (lane == 0) {
      sc.chex[sc.nch] = carry;
      sc.cta_agg[sc.r] = carry;
    }
  }
  __syncthreads();
  {  // records -> the CTA's list in global memory, in scan order, with CTA-local exclusive aggregates
    const int nrec = sh.nrec < V2_RCTA ? sh.nrec : V2_RCTA;
    if (tid < nrec) {
      const V2PoolRec pr = sc.srec[tid];
      const pgb_agg cex_ = sc.chex[pr.cc];
      const int g = cex_.cnt + pr.k;
      if (g < V2_RCTA) {
        ScanRec rr = pr.rec;
        rr.ex = pgb_agg_combine(cex_, rr.ex);
        sc.recs[(int64_t)sc.r * V2_RCTA + g] = rr;
      } else {
        atomicOr(sc.flagp, 1);
      }
    }
  }
}

// PD, first part (every CTA of the chain, redundantly): exclusive scan of the CTA aggregates, gather of the records,
// serial chain over the complex lanes with real FP adds -> seeds.  Returns false when the scan must take the serial
// fallback (too many records / forced).  Contains __syncthreads.
__device__ __forceinline__ bool v2_resolve(const V2Scan& sc, V2Shared& sh, int force_fallback) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npc = sc.npc;
  if (tid < npc) {  // one CTA aggregate per thread, then the scan runs out of shared memory
    const pgb_agg* g = sc.cta_agg + tid;
    pgb_agg a;
    a.f.c0 = __ldcg(&g->f.c0); a.f.c1 = __ldcg(&g->f.c1); a.cnt = __ldcg(&g->cnt); a.key = __ldcg(&g->key);
    sh.cex[tid] = a;
  }
  __syncthreads();
  if (warp == 0) {  // in-place exclusive scan: lane l owns entries [l*PL, (l+1)*PL)
    constexpr int PL = (V2_MAXNPC + 31) / 32;
    pgb_agg v[PL], run = pgb_agg_identity();
#pragma unroll
    for (int j = 0; j < PL; ++j) {
      const int i = lane * PL + j;
      v[j] = (i < npc) ? sh.cex[i] : pgb_agg_identity();
      run = pgb_agg_combine(run, v[j]);
    }
    pgb_agg tot;
    pgb_agg ex = v2_warp_excl_scan_agg(run, &tot);
#pragma unroll
    for (int j = 0; j < PL; ++j) {
      const int i = lane * PL + j;
      if (i < npc) sh.cex[i] = ex;
      ex = pgb_agg_combine(ex, v[j]);
    }
    if (lane == 0) sh.cex[npc] = tot;
  }
  __syncthreads();
  const int C = sh.cex[npc].cnt;
  if (C > V2_MAXREC || force_fallback) {
    if (tid == 0) {
      atomicOr(sc.flagp, 1);
      sh.seeds[0] = 0.0;
    }
    __syncthreads();
    return false;
  }
  if (tid < C) {  // one record per thread: find its CTA (cex[i].cnt <= g < cex[i+1].cnt), load, rebase
    int lo = 0, hi = npc - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sh.cex[mid].cnt <= tid) lo = mid; else hi = mid - 1;
    }
    const pgb_agg te = sh.cex[lo];
    const int k = tid - te.cnt;
    if (k < V2_RCTA) {
      const ScanRec* rp = sc.recs + (int64_t)lo * V2_RCTA + k;
      pgb_agg rex;
      rex.f.c0 = __ldcg(&rp->ex.f.c0); rex.f.c1 = __ldcg(&rp->ex.f.c1); rex.cnt = __ldcg(&rp->ex.cnt); rex.key = __ldcg(&rp->ex.key);
      sh.rf[tid] = pgb_agg_combine(te, rex).f;
#pragma unroll
      for (int i = 0; i < 4; ++i) sh.rW[tid][i] = __ldcg(&rp->W[i]);
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
  return true;
}

__device__ __forceinline__ double v2_qin(const V2Shared& sh, pgb_agg ex) {
  const int c = ex.cnt < 0 ? 0 : (ex.cnt > V2_MAXREC ? V2_MAXREC : ex.cnt);
  return pgb_qin_from(ex, sh.seeds[c]);
}

// Exact prefixes q[0..4] of one lane of chunk cc (q[0] = value entering the lane) + the bitwise seam check against the
// right neighbour.  fallback_pass: incoming values come from the serial walk (no check).
__device__ __forceinline__ void v2_exact_prefixes(const V2Scan& sc, const V2Shared& sh, int cc, bool fallback_pass,
                                                  const double* __restrict__ qin_fb, double (&q)[5]) {
  const int lane = threadIdx.x & 31;
  const int loc = cc * V2_CHUNK + lane * 4;
  const int n0 = sc.p0 + loc;
  double W[4];
  {
    const double2 a = *reinterpret_cast<const double2*>(sc.v + loc), b = *reinterpret_cast<const double2*>(sc.v + loc + 2);
    W[0] = a.x; W[1] = a.y; W[2] = b.x; W[3] = b.y;
  }
  double qin;
  if (fallback_pass) {
    qin = qin_fb[n0 >> 2];
  } else {
    qin = v2_qin(sh, pgb_agg_combine(pgb_agg_combine(sh.cex[sc.r], sc.chex[cc]), sc.TE[cc * 32 + lane]));
  }
  q[0] = qin;
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i + 1] = __dadd_rn(q[i], W[i]);
  if (!fallback_pass) {
    double nxt = __shfl_down_sync(0xffffffffu, qin, 1);
    bool have = (n0 + 4 < sc.N);  // seams past the last real element cannot matter
    if (lane == 31 && have) {
      if (cc + 1 < sc.nch) nxt = v2_qin(sh, pgb_agg_combine(sh.cex[sc.r], sc.chex[cc + 1]));
      else if (sc.r + 1 < sc.npc) nxt = v2_qin(sh, sh.cex[sc.r + 1]);
      else have = false;
    }
    if (have && __double_as_longlong(nxt) != __double_as_longlong(q[4])) atomicOr(sc.flagp, 1);
  }
}

// Offspring counts of one lane: rr[k] = #{ draws j : u_j <= q[k] } (non-decreasing), with the reference's edge cases:
// draws before every prefix get index 0 (first lane of the chain), q_N < u_j raises the overrun bit.
__device__ __forceinline__ void v2_counts(const pgb_utable* ut, const double (&q)[5], int n0, int64_t N, int* stat, int (&rr)[5],
                                          int& zero_upto) {
  const int M = (int)ut->M;
  zero_upto = 0;
  if (n0 >= N) {
#pragma unroll
    for (int i = 0; i < 5; ++i) rr[i] = M;
    return;
  }
  int row = pgb_utable_row_of(ut, q[0]);
  int r0 = (int)pgb_utable_count_le_hint(ut, q[0], &row);
  if (n0 == 0) {
    if (r0 > 0) atomicOr(stat, 2);
    zero_upto = r0;
    r0 = 0;
  }
  rr[0] = r0;
  int prev = (n0 == 0) ? zero_upto : r0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (n0 + i < N) {
      int rv = (int)pgb_utable_count_le_hint(ut, q[i + 1], &row);
      if (rv < prev) rv = prev;
      if (n0 + i == N - 1 && rv < M) {
        atomicOr(stat, 1);
        rv = M;
      }
      rr[i + 1] = rv;
      prev = rv;
    } else {
      rr[i + 1] = prev;
    }
  }
}

// serial exact walk by warp 0 of the chain's lead CTA over the dumped normalised weights (fix-up path): q_n = q_{n-1} + W[n]
// strictly left to right (:28).  The 128 weights of a 32-thread group are staged in shared memory (double-buffered, the next
// group's loads in flight meanwhile) and every lane walks them redundantly with broadcast reads — one dependent add per
// element and nothing else on the chain (the first version broadcast each weight with two shuffles from its owner lane and
// loaded a group only after the previous one was finished: 35 cycles per element, 19 ms per walk at N = 2^20; now ~5 ms).
// Lane l keeps the value that enters thread t0 + l of the tile layout.
__device__ __forceinline__ void v2_seq_exact(const double* __restrict__ Wn, double* __restrict__ qin_fb, int64_t N, int nt,
                                             long long* counters, double* __restrict__ sbuf /* [2][128] shared */) {
  const int lane = threadIdx.x;
  if (lane >= 32) return;
  const int64_t nthr = (int64_t)nt * TPB;
  double q = 0.0;
  double Wc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) Wc[i] = ((int64_t)lane * 4 + i < N) ? __ldcg(Wn + (int64_t)lane * 4 + i) : 0.0;
  for (int64_t t0 = 0; t0 < nthr; t0 += 32) {
    double* buf = sbuf + ((t0 >> 5) & 1) * 128;
    *reinterpret_cast<double2*>(buf + lane * 4) = make_double2(Wc[0], Wc[1]);
    *reinterpret_cast<double2*>(buf + lane * 4 + 2) = make_double2(Wc[2], Wc[3]);
    if (t0 + 32 < nthr) {
      const int64_t base = (t0 + 32 + lane) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) Wc[i] = (base + i < N) ? __ldcg(Wn + base + i) : 0.0;
    }
    __syncwarp();
    double mine = 0.0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {
      mine = (lane == l) ? q : mine;
      const double2 a = *reinterpret_cast<const double2*>(buf + l * 4), b = *reinterpret_cast<const double2*>(buf + l * 4 + 2);
      q = __dadd_rn(q, a.x);
      q = __dadd_rn(q, a.y);
      q = __dadd_rn(q, b.x);
      q = __dadd_rn(q, b.y);
    }
    qin_fb[t0 + lane] = mine;
  }
  if (lane == 0) counters[0] += 1;
}

// ---------------------------------------------------------------------------------------------------- the kernel
template <int NX, int BASIS, bool FIRST, int CA, typename RT>
__global__ void __launch_bounds__(V2_NT, 1)
k_sweep_v2(const __grid_constant__ FilterDev fd, const __grid_constant__ V2Dev vd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V2Shared& sh = *reinterpret_cast<V2Shared*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npc = vd.npc;
  const int chain = blockIdx.x / npc, r = blockIdx.x - chain * npc;
  const int ntile = vd.tiles_lo + (r < vd.n_hi ? 1 : 0);
  const int t0tile = r * vd.tiles_lo + (r < vd.n_hi ? r : vd.n_hi);
  const int nch = ntile * V2_CPT;
  const int p0 = t0tile * TILE;
  const int64_t N = fd.N, T = fd.T;
  const int nt = fd.nt;
  const bool lead = (r == 0);
  constexpr bool ANC = !FIRST;
  const V2Layout L = v2_smem_layout(vd.mch, NX * fd.n_phi, vd.stationary);
  double* chs = reinterpret_cast<double*>(smem_raw + L.chs);        // [4][mch]: chunk sums of e, w, waN, waN'
  double* chpfx = reinterpret_cast<double*>(smem_raw + L.chpfx);    // [2][mch+1]
  pgb_agg* chagg = reinterpret_cast<pgb_agg*>(smem_raw + L.chagg);
  pgb_agg* chex = reinterpret_cast<pgb_agg*>(smem_raw + L.chex);
  RT* A_sm = reinterpret_cast<RT*>(smem_raw + L.A);  // the chain's A as RT (F32 mode: rounded once)
  int* s_anc = reinterpret_cast<int*>(smem_raw + L.sanc);
  const int64_t NP = (int64_t)nt * TILE;
  double* E = vd.stationary ? reinterpret_cast<double*>(smem_raw + L.E) : vd.gE + (int64_t)chain * NP + p0;
  double* Pv = vd.stationary ? reinterpret_cast<double*>(smem_raw + L.P) : vd.gP + (int64_t)chain * NP + p0;
  pgb_agg* TE = vd.stationary ? reinterpret_cast<pgb_agg*>(smem_raw + L.TE) : vd.gTE + ((int64_t)chain * NP + p0) / 4;
  const int mch = vd.mch;
  unsigned int bar_target = 0;
  int par = 0;
#ifdef PGB_EXPERIMENT
  long long tacc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) tacc[i] = 0;
  long long tlast = clock64();
#define V2_TICK(idx) do { if (tid == 0) { long long now_ = clock64(); tacc[idx] += now_ - tlast; tlast = now_; } } while (0)
#else
#define V2_TICK(idx) do { } while (0)
#endif
  // tile sums (replicating them over several L2 slices for the 148 simultaneous readers was tried: no change)
  double* pe = vd.pe + (int64_t)chain * nt;
  double* pw = vd.pw + (int64_t)chain * nt;
  double* pa = vd.pa + (int64_t)chain * nt;
  double* pa2 = vd.pa2 + (int64_t)chain * nt;
  RT* lwg = reinterpret_cast<RT*>(fd.lwbuf) + (int64_t)chain * N;
  const uint32_t stream_z = ((uint32_t)PGB_STREAM_Z << 24) | ((fd.chain_base + chain) & 0x00ffffffu);
  for (int i = tid; i < NX * fd.n_phi; i += V2_NT) A_sm[i] = (RT)fd.A[(int64_t)chain * NX * fd.n_phi + i];
  RT Lq[NX * NX];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) Lq[i] = (RT)fd.Lq[chain * 16 + i];
  RDiv<RT> ivR;
  ivR.set(fd.R[0]);

  V2Scan sm_;  // main scan
  sm_.v = E; sm_.TE = TE; sm_.chsum = chs + mch; sm_.chpfx = chpfx; sm_.chagg = chagg; sm_.chex = chex;
  sm_.srec = reinterpret_cast<V2PoolRec*>(s_anc);
  sm_.cta_agg = vd.cta_agg + (int64_t)chain * npc;
  sm_.recs = vd.recs + (int64_t)chain * npc * V2_RCTA;
  sm_.status = &fd.status[chain];
  sm_.N = N; sm_.p0 = p0; sm_.nch = nch; sm_.r = r; sm_.npc = npc; sm_.nt = nt;
  V2Scan ss_ = sm_;  // side scan (ancestor weights)
  ss_.v = Pv; ss_.chsum = chs + 3 * mch; ss_.chpfx = chpfx + (mch + 1);
  ss_.cta_agg = vd.cta_agg + ((int64_t)fd.nc + chain) * npc;
  ss_.recs = vd.recs + ((int64_t)fd.nc + chain) * npc * V2_RCTA;

  // tile sums of one of the chunk-sum rows -> global; row = 0..3.  Called by all threads after a __syncthreads.
  auto publish_tiles = [&](int row, double* dst) {
    if (tid < ntile) dst[t0tile + tid] = v2_tile_from_chunks(chs + row * mch + tid * V2_CPT);
  };
  // exclusive raw prefix of a chunk-sum row (one warp)
  auto chunk_prefix = [&](int row, double* out) {
    double carry = 0.0;
    for (int base = 0; base < nch; base += 32) {
      const int cc = base + lane;
      const double v = (cc < nch) ? chs[row * mch + cc] : 0.0;
      const double inc = v2_warp_incl_scan(v);
      if (cc < nch) out[cc] = __dadd_rn(carry, __dsub_rn(inc, v));
      carry = __dadd_rn(carry, __shfl_sync(0xffffffffu, inc, 31));
    }
    if (lane == 0) out[nch] = carry;
  };
  auto block_max = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) sh.redm[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double mm = sh.redm[0];
      for (int i = 1; i < V2_NW; ++i) mm = fmax(mm, sh.redm[i]);
      atomicMax(&fd.maxkey[chain], ordered_key(mm));
    }
  };

  // ------------------------------------------------------------------ t = 0: initial particles (:160-165)
  {
    RT* x0 = r_x_row<RT>(fd, chain, 0);
    const double* xp = fd.xprim + (int64_t)chain * NX * T;
    const double* z0in = fd.Z0 ? fd.Z0 + (int64_t)chain * NX * (N - 1) : nullptr;
    const uint32_t stream_z0 = ((uint32_t)PGB_STREAM_Z0 << 24) | ((fd.chain_base + chain) & 0x00ffffffu);
    RT lmax = (RT)-INFINITY;
    const int pend = p0 + nch * V2_CHUNK;
    for (int n = p0 + tid; n < pend && n < N; n += V2_NT) {
      RT x[NX];
      if (n == N - 1) {
#pragma unroll
        for (int d = 0; d < NX; ++d) x[d] = (RT)xp[d];
      } else {
        RT z[NX];
        r_normals<NX>(fd, stream_z0, 0, n, z0in, z);
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          RT acc = (RT)0.0;
#pragma unroll
          for (int e = 0; e <= d; ++e) acc = rfma((RT)fd.x0_chol[d + NX * e], z[e], acc);
          x[d] = radd((RT)fd.x0_mean[d], acc);
        }
      }
#pragma unroll
      for (int d = 0; d < NX; ++d) x0[(int64_t)d * N + n] = x[d];
      const RT lw = r_log_weight<NX>(fd, x, fd.y, ivR);
      lwg[n] = lw;
      lmax = fmax(lmax, lw);
      if (!(lw == lw)) atomicOr(&fd.status[chain], 4);
    }
    block_max((double)lmax);
  }
  v2_grid_barrier(vd.bar, bar_target);

  for (int64_t t = 1; t <= T; ++t) {
    const bool last_only = (t == T);  // finalise w_T and draw the trajectory index only
    const bool anc = ANC && !last_only;
    int* fmain = &fd.main.flag[chain + fd.nc * (int)(t & 1)];
    int* fside = &fd.side.flag[chain + fd.nc * (int)(t & 1)];
    sm_.flagp = fmain;
    ss_.flagp = fside;
    // ================================================================ PA: e, F, pdf
    {
      const RT m = (RT)ordered_unkey(__ldcg(&fd.maxkey[chain]));  // the maximum of RT values: the conversion is exact
      const RT* xp = r_x_row<RT>(fd, chain, t - 1);
      RT* Fb = reinterpret_cast<RT*>(fd.Fbuf) + (int64_t)chain * NX * N;
      const double* up = fd.u + (int64_t)fd.n_u * (t - 1);
      RT tku[BASIS == 1 ? 5 : (BASIS == 2 ? PGB_MAX_Z_ * MAXC : 1)];
      if (!last_only) r_input_sines<BASIS, RT>(fd, up, tku);  // the same for every particle of this step
      RT xref[NX];
      RDiv<RT> ivL[NX];
      RT c0 = (RT)0.0;
      if (anc) {
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          xref[d] = (RT)fd.xprim[(int64_t)chain * NX * T + d + NX * t];
          ivL[d].set(fd.Lq[chain * 16 + d + NX * d]);
        }
        c0 = (RT)fd.c0[chain];
      }
      // strided mapping: coalesced scalar accesses, one particle at a time; the inputs of the NEXT particle are
      // loaded before the current one is processed (the loop is latency-bound on these L2 reads otherwise)
      RT lw_n = (RT)0.0, x_n[NX];
#pragma unroll
      for (int d = 0; d < NX; ++d) x_n[d] = (RT)0.0;
      {
        const int n = p0 + warp * V2_CHUNK + lane;
        if (warp < nch && n < N) {
          lw_n = __ldcg(lwg + n);
          if (!last_only) {
#pragma unroll
            for (int d = 0; d < NX; ++d) x_n[d] = __ldcg(xp + (int64_t)d * N + n);
          }
        }
      }
      for (int cc = warp; cc < nch; cc += V2_NW) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int loc = cc * V2_CHUNK + k * 32 + lane;
          const int n = p0 + loc;
          const RT lw_c = lw_n;
          RT x[NX];
#pragma unroll
          for (int d = 0; d < NX; ++d) x[d] = x_n[d];
          {  // prefetch
            const int cn = (k == 3) ? cc + V2_NW : cc, kn = (k + 1) & 3;
            const int nn = p0 + cn * V2_CHUNK + kn * 32 + lane;
            if (cn < nch && nn < N) {
              lw_n = __ldcg(lwg + nn);
              if (!last_only) {
#pragma unroll
                for (int d = 0; d < NX; ++d) x_n[d] = __ldcg(xp + (int64_t)d * N + nn);
              }
            }
          }
          double e = 0.0, pdf = 0.0;
          if (n < N) {
            e = (double)rexp(rsub(lw_c, m));  // F32 mode: the FP32 exponential, widened exactly
            if (!last_only) {
              RT F[NX];
              r_eval_F<NX, BASIS, CA, RT>(fd, A_sm, v2_const_A<RT>(), x, up, tku, F);
#pragma unroll
              for (int d = 0; d < NX; ++d) Fb[(int64_t)d * N + n] = F[d];
              if (anc) {  // pdf(MvNormal(x'_t, Q), F) = exp(c0 - |L^-1 (F - x'_t)|^2 / 2)    (:178-179)
                RT rv[NX], sq = (RT)0.0;
#pragma unroll
                for (int d = 0; d < NX; ++d) {
                  RT acc = rsub(F[d], xref[d]);
#pragma unroll
                  for (int e2 = 0; e2 < d; ++e2) acc = rfma(-Lq[d + NX * e2], rv[e2], acc);
                  rv[d] = ivL[d](acc);
                  sq = rfma(rv[d], rv[d], sq);
                }
                pdf = rexp_wide(rsub(c0, rmul(sq, (RT)0.5)));
              }
            }
          }
          E[loc] = e;
          if (anc) Pv[loc] = pdf;
        }
        __syncwarp();
        {  // chunk sum of e in the tree's leaf order
          const int loc = cc * V2_CHUNK + lane * 4;
          const double2 a = *reinterpret_cast<const double2*>(E + loc), b = *reinterpret_cast<const double2*>(E + loc + 2);
          const double v[4] = {a.x, a.y, b.x, b.y};
          const double s = warp_tree_sum(tree4(v));
          if (lane == 0) chs[cc] = s;
        }
      }
      __syncthreads();
      publish_tiles(0, pe);
    }
    V2_TICK(1);
    v2_grid_barrier(vd.bar, bar_target);  // B1
    V2_TICK(2);
    // ================================================================ PB: w = e/S1, waN = w*pdf
    {
      const double S1 = v2_tiles_collective(pe, nt, 0, -1, sh, par);
      V2_TICK(13);
      const InvDiv ivS1 = make_invdiv(S1);
      if (tid == 0) sh.nrec = 0;  // record pool of this step's scan-local (PC)
      if (lead && tid == 0) {
        fd.maxkey[chain] = 0ull;  // every CTA has read it (PA) and nobody raises it before PD
        vd.cand[2 * chain] = 0;
      }
      double* wh = fd.whist ? fd.whist + ((int64_t)chain * T + (t - 1)) * N : nullptr;
      double* wout = last_only ? fd.wbuf + (int64_t)chain * N : nullptr;
      for (int cc = warp; cc < nch; cc += V2_NW) {
        const int loc = cc * V2_CHUNK + lane * 4;
        const int n0 = p0 + loc;
        const double2 a = *reinterpret_cast<const double2*>(E + loc), b = *reinterpret_cast<const double2*>(E + loc + 2);
        const double e[4] = {a.x, a.y, b.x, b.y};
        double w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = (n0 + i < N) ? div_inv(e[i], ivS1) : 0.0;
        *reinterpret_cast<double2*>(E + loc) = make_double2(w[0], w[1]);
        *reinterpret_cast<double2*>(E + loc + 2) = make_double2(w[2], w[3]);
        if (wh || wout) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (n0 + i < N) {
              if (wh) wh[n0 + i] = w[i];
              if (wout) wout[n0 + i] = w[i];
            }
        }
        const double sw = warp_tree_sum(tree4(w));
        if (lane == 0) chs[mch + cc] = sw;
        if (anc) {
          const double2 c = *reinterpret_cast<const double2*>(Pv + loc), d2 = *reinterpret_cast<const double2*>(Pv + loc + 2);
          const double pd[4] = {c.x, c.y, d2.x, d2.y};
          double av[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = (n0 + i < N) ? __dmul_rn(w[i], pd[i]) : 0.0;
          *reinterpret_cast<double2*>(Pv + loc) = make_double2(av[0], av[1]);
          *reinterpret_cast<double2*>(Pv + loc + 2) = make_double2(av[2], av[3]);
          const double sa = warp_tree_sum(tree4(av));
          if (lane == 0) chs[2 * mch + cc] = sa;
        }
      }
      __syncthreads();
      publish_tiles(1, pw);
      if (anc) publish_tiles(2, pa);
      if (warp == V2_NW - 1) chunk_prefix(1, chpfx);
      if (anc && warp == V2_NW - 2) chunk_prefix(2, chpfx + (mch + 1));  // raw prefix of the waN chunk sums: ancestor shortcut (PC)
    }
    V2_TICK(3);
    v2_grid_barrier(vd.bar, bar_target);  // B2
    V2_TICK(4);
    // ================================================================ PC: main scan-local, side normalise
    double S3 = 1.0;
    const double S2 = anc ? v2_tiles_collective2(pw, pa, nt, t0tile, t0tile + ntile, sh, par, &S3)  // + sh.tp_lo / tp_hi
                          : v2_tiles_collective(pw, nt, t0tile, t0tile + ntile, sh, par);
    {
      __syncthreads();  // sh.tp_lo / tp_hi (written by thread 0 after the collective's barrier)
      V2_TICK(14);
      v2_scan_local(sm_, S2, sh);
      V2_TICK(15);
      if (anc) {
        // Ancestor draw of slot N (:181) by the CERTIFIED shortcut, straight from the un-normalised waN: the reference
        // normalises twice (waN' = waN/S3 at :180, W' = waN'/S4 inside systematic_resampling :19) and walks q += W'[n];
        // S3 and S4 are balanced-tree sums, so |S4 - 1| <= 44 eps and W'_n = (waN_n/S3)(1 +- 46 eps), and the sequential
        // walk adds at most n eps: |q_n - sum_{i<=n} waN_i/S3| <= (N + 46) eps.  The approximate prefix below (tile /
        // chunk / lane prefixes of waN times fl(1/S3)) is within ~100 eps of that sum, so delta = 6 (N + 4096) eps bounds
        // their distance with a wide margin; a draw farther than delta from both neighbouring prefixes is decided here,
        // the others (about 1e-3 of the steps) take the exact scan after B0 — which is also where waN', its sums and S4
        // are computed now: the common path no longer touches them.
        const double inv3 = __ddiv_rn(1.0, S3);
        const double tqa = sh.tq_lo;
        const double ua = fd.utab_anc[(int64_t)chain * T + t].u0[0];  // one draw: u = rand()
        const double delta = 6.0 * ((double)N + 4096.0) * 0x1p-53;
        const double* cpa = chpfx + (mch + 1);
        for (int cc = warp; cc < nch; cc += V2_NW) {
          const int loc = cc * V2_CHUNK + lane * 4;
          const int n0 = p0 + loc;
          const double2 c = *reinterpret_cast<const double2*>(Pv + loc), d2 = *reinterpret_cast<const double2*>(Pv + loc + 2);
          const double av[4] = {c.x, c.y, d2.x, d2.y};
          double W[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) W[i] = (n0 + i < N) ? __dmul_rn(av[i], inv3) : 0.0;
          const double s1 = W[0], s2 = __dadd_rn(s1, W[1]), s3 = __dadd_rn(s2, W[2]), s4 = __dadd_rn(s3, W[3]);
          const double inc = v2_warp_incl_scan(s4);
          const double P = __dadd_rn(__dmul_rn(__dadd_rn(tqa, cpa[cc]), inv3), __dsub_rn(inc, s4));
          const double qt[5] = {P, __dadd_rn(P, s1), __dadd_rn(P, s2), __dadd_rn(P, s3), __dadd_rn(P, s4)};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (n0 + i < N && (qt[i] + delta < ua) && (ua <= qt[i + 1] - delta)) {
              atomicAdd(&vd.cand[2 * chain], 1);
              vd.cand[2 * chain + 1] = n0 + i;
            }
          }
        }
      }
    }
    V2_TICK(5);
    v2_grid_barrier(vd.bar, bar_target);  // B3
    V2_TICK(6);
    // By-output expansion of the offspring counts the lanes left in their TE slots (rr[0..4], zero_upto): the chain's
    // output slots [J0, J1) of this CTA in pieces of V2_CAP — ancestors scattered into the staging buffer, then every
    // thread propagates output slots (:175 | :188, :194).  Used by the regular PD phase and by the serial-fallback re-walk
    // (whose first version expanded by input, one lane walking all offspring of its particle: with the degenerate weights
    // that cause a record overflow in the first place — one particle holding most of the mass — a single thread emitted
    // ~N outputs, 450 ms per fallback at N = 2^20; chains whose streams hit two such steps ran 4 x slower).
    auto expand_pieces = [&](const int J0, const int J1, int* pend, double& lmax) {
      RT* xt = r_x_row<RT>(fd, chain, t);
      int32_t* at = fd.ah + ((int64_t)chain * T + t) * N;
      const RT* Fb = reinterpret_cast<const RT*>(fd.Fbuf) + (int64_t)chain * NX * N;
      const double* yt = fd.y + (int64_t)fd.n_y * t;
      const double* zin = fd.Z ? fd.Z + ((int64_t)chain * T + t) * NX * N : nullptr;
      for (int ps = J0; ps < J1; ps += V2_CAP) {
        const int pe_ = (J1 - ps < V2_CAP) ? J1 : ps + V2_CAP;
        // (2) scatter: slot jj in [rr[k], rr[k+1]) belongs to particle n0 + k
        for (int cc = warp; cc < nch; cc += V2_NW) {
          const int* slot = reinterpret_cast<const int*>(&TE[cc * 32 + lane]);
          const int n0 = p0 + cc * V2_CHUNK + lane * 4;
          const int zu = slot[5];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int lo = slot[i], hi = slot[i + 1];
            if (lo < ps) lo = ps;
            if (hi > pe_) hi = pe_;
            const int anc_id = n0 + i;
            // short ranges by the lane itself, long ones by the whole warp
            int len = hi - lo;
            if (len < 0) len = 0;
            const int mine = len < 8 ? len : 0;
            for (int j = 0; j < mine; ++j) s_anc[lo + j - ps] = (lo + j < zu) ? -1 : anc_id;
            unsigned big = __ballot_sync(0xffffffffu, len >= 8);
            while (big) {
              const int src = __ffs(big) - 1;
              big &= big - 1;
              const int blo = __shfl_sync(0xffffffffu, lo, src), bhi = __shfl_sync(0xffffffffu, hi, src);
              const int bid = __shfl_sync(0xffffffffu, anc_id, src), bzu = __shfl_sync(0xffffffffu, zu, src);
              for (int j = blo + lane; j < bhi; j += 32) s_anc[j - ps] = (j < bzu) ? -1 : bid;
            }
          }
        }
        __syncthreads();
        V2_TICK(0);
        // (3) propagate: every thread takes output slots (:175 | :188, :194)
        const int cnt = pe_ - ps;
        int a_n = (tid < cnt) ? s_anc[tid] : 0;
        RT F_n[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) F_n[d] = (tid < cnt) ? __ldcg(Fb + (int64_t)d * N + (a_n < 0 ? 0 : a_n)) : (RT)0.0;
#pragma unroll 1
        for (int o = tid; o < cnt; o += V2_NT) {
          const int jj = ps + o;
          const int a0 = a_n;
          RT F[NX], z[NX], x[NX];
#pragma unroll
          for (int d = 0; d < NX; ++d) F[d] = F_n[d];
          if (o + V2_NT < cnt) {  // the gathered F of the next slot is in flight while this one is propagated
            a_n = s_anc[o + V2_NT];
#pragma unroll
            for (int d = 0; d < NX; ++d) F_n[d] = __ldcg(Fb + (int64_t)d * N + (a_n < 0 ? 0 : a_n));
          }
          r_normals<NX>(fd, stream_z, (uint32_t)t, jj, zin, z);
#pragma unroll
          for (int d = 0; d < NX; ++d) {
            RT acc = (RT)0.0;
#pragma unroll
            for (int e = 0; e <= d; ++e) acc = rfma(Lq[d + NX * e], z[e], acc);
            x[d] = radd(F[d], acc);
            xt[(int64_t)d * N + jj] = x[d];
          }
          at[jj] = a0;
          const RT lw = r_log_weight<NX>(fd, x, yt, ivR);
          lwg[jj] = lw;
          lmax = fmax(lmax, (double)lw);
          if (!(lw == lw)) atomicOr(pend, 4);
        }
        __syncthreads();
      }
    };
    // ================================================================ PD: resolve, exact prefixes, expansion; side shortcut
    {
      {  // the step's sampling grid -> shared memory; issued first so that its L2 latency hides behind the resolve
        const pgb_utable* utg = last_only ? &fd.utab_star[chain] : &fd.utab_res[(int64_t)chain * T + t];
        const int* src = reinterpret_cast<const int*>(utg);
        int* dst = reinterpret_cast<int*>(&sh.ut);
        for (int i = tid; i < (int)(sizeof(pgb_utable) / 4); i += V2_NT) dst[i] = src[i];
      }
      const bool ok = v2_resolve(sm_, sh, vd.force_fallback & 1);
      if (tid == 0) {
        sh.flag_sh = __ldcg(fmain) & 1;
        if (lead) fd.main.counters[2 * chain + 1] += 1;
      }
      __syncthreads();
      const bool skip = !ok || sh.flag_sh != 0;
      V2_TICK(12);
      int* pend = &fd.main.pend[chain];
      double lmax = -INFINITY;
      if (!skip) {
        // (1) exact prefixes, seam check, offspring counts (kept in the lane's TE slot)
        for (int cc = warp; cc < nch; cc += V2_NW) {
          double q[5];
          v2_exact_prefixes(sm_, sh, cc, false, nullptr, q);
          int rr[5], zu;
          const int n0 = p0 + cc * V2_CHUNK + lane * 4;
          v2_counts(&sh.ut, q, n0, N, pend, rr, zu);
          int* slot = reinterpret_cast<int*>(&TE[cc * 32 + lane]);
#pragma unroll
          for (int i = 0; i < 5; ++i) slot[i] = rr[i];
          slot[5] = zu;
          if (cc == 0 && lane == 0) sh.J0 = rr[0];
          if (cc == nch - 1 && lane == 31) sh.J1 = rr[4];
        }
        __syncthreads();
        V2_TICK(11);
        const int J0 = (int)sh.J0, J1 = (int)sh.J1;
        if (last_only) {
          // one draw: the lane whose range contains it writes the index
          for (int cc = warp; cc < nch; cc += V2_NW) {
            const int* slot = reinterpret_cast<const int*>(&TE[cc * 32 + lane]);
            const int n0 = p0 + cc * V2_CHUNK + lane * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (n0 + i < N && slot[i + 1] > slot[i]) vd.star_idx[chain] = (slot[i] < slot[5]) ? -1 : n0 + i;
          }
        } else {
          RT* xt = r_x_row<RT>(fd, chain, t);
          int32_t* at = fd.ah + ((int64_t)chain * T + t) * N;
          const RT* Fb = reinterpret_cast<const RT*>(fd.Fbuf) + (int64_t)chain * NX * N;
          const double* yt = fd.y + (int64_t)fd.n_y * t;
          const double* zin = fd.Z ? fd.Z + ((int64_t)chain * T + t) * NX * N : nullptr;
          expand_pieces(J0, J1, pend, lmax);
          if (ANC && lead && tid == 0) {  // the conditioned trajectory keeps slot N-1 (:160)
            RT x[NX];
#pragma unroll
            for (int d = 0; d < NX; ++d) {
              x[d] = (RT)fd.xprim[(int64_t)chain * NX * T + d + NX * t];
              xt[(int64_t)d * N + (N - 1)] = x[d];
            }
            const RT lw = r_log_weight<NX>(fd, x, yt, ivR);
            lwg[N - 1] = lw;
            lmax = fmax(lmax, (double)lw);
          }
        }
      }
      if (!last_only) block_max(lmax);
      V2_TICK(7);
    }
    V2_TICK(8);
    v2_grid_barrier(vd.bar, bar_target);  // B0
    V2_TICK(9);
    // ================================================================ rare, grid-uniform sections
    {
      // Every word this section reads is requested up front (one L2 round trip instead of four or five dependent ones:
      // the grid-uniform conditions, this chain's candidate, the lead thread's pending status bits).
      const int tpar = (int)(t & 1);
      int amb = 0, anyf0 = 0;
      for (int ch = lane; ch < fd.nc; ch += 32) {
        const int cz = anc ? __ldcg(&vd.cand[2 * ch]) : 1;
        const int f1 = __ldcg(&fd.main.flag[ch + fd.nc * tpar]), f2 = __ldcg(&fd.side.flag[ch + fd.nc * tpar]);
        amb |= (cz == 0) ? 1 : 0;
        anyf0 |= (f1 | f2) & 1;
      }
      const int my_cnt = anc ? __ldcg(&vd.cand[2 * chain]) : 1;
      const int my_idx = (anc && lead && tid == 0) ? __ldcg(&vd.cand[2 * chain + 1]) : 0;
      int pend0 = 0;
      if (lead && tid == 0) pend0 = __ldcg(&fd.main.pend[chain]) | __ldcg(&fd.side.pend[chain]);
      bool rare_ran = false;  // a rare path ran: flags / pending bits may have changed since the prefetch
      // (a) ancestor draw undecided by the shortcut (about 1e-3 of the steps), or forced: exact scan of waN'
      if (anc) {
        const bool amb_any = __any_sync(0xffffffffu, amb) || (vd.force_fallback != 0);
        const bool amb_me = (my_cnt == 0) || (vd.force_fallback != 0);
        int32_t* aout = fd.ah + ((int64_t)chain * T + t) * N + (N - 1);
        if (!amb_me && lead && tid == 0) *aout = my_idx;
        rare_ran = amb_any;
        if (amb_any) {
          // waN' = waN / S3 (:180) with its chunk / tile sums and chunk prefixes: only this exact scan needs them
          if (amb_me) {
            const InvDiv ivS3 = make_invdiv(S3);
            for (int cc = warp; cc < nch; cc += V2_NW) {
              const int loc = cc * V2_CHUNK + lane * 4;
              const int n0 = p0 + loc;
              const double2 c = *reinterpret_cast<const double2*>(Pv + loc), d2 = *reinterpret_cast<const double2*>(Pv + loc + 2);
              const double av[4] = {c.x, c.y, d2.x, d2.y};
              double v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = (n0 + i < N) ? div_inv(av[i], ivS3) : 0.0;
              *reinterpret_cast<double2*>(Pv + loc) = make_double2(v[0], v[1]);
              *reinterpret_cast<double2*>(Pv + loc + 2) = make_double2(v[2], v[3]);
              const double sv = warp_tree_sum(tree4(v));
              if (lane == 0) chs[3 * mch + cc] = sv;
            }
          }
          __syncthreads();
          if (amb_me) {
            publish_tiles(3, pa2);
            if (warp == V2_NW - 1) chunk_prefix(3, chpfx + (mch + 1));
          }
          v2_grid_barrier(vd.bar, bar_target);  // every CTA of the chain has published its tile sums of waN'
          const double S4 = v2_tiles_collective(pa2, nt, t0tile, t0tile + ntile, sh, par);
          if (tid == 0) sh.nrec = 0;
          __syncthreads();
          if (amb_me) v2_scan_local(ss_, S4, sh);
          v2_grid_barrier(vd.bar, bar_target);
          if (amb_me) {
            const bool ok = v2_resolve(ss_, sh, vd.force_fallback & 1);
            if (tid == 0) {
              sh.flag_sh = __ldcg(fside) & 1;
              if (lead) fd.side.counters[2 * chain + 1] += 1;
            }
            {
              const int* src = reinterpret_cast<const int*>(&fd.utab_anc[(int64_t)chain * T + t]);
              int* dst = reinterpret_cast<int*>(&sh.ut);
              for (int i = tid; i < (int)(sizeof(pgb_utable) / 4); i += V2_NT) dst[i] = src[i];
            }
            __syncthreads();
            if (ok && sh.flag_sh == 0) {
              for (int cc = warp; cc < nch; cc += V2_NW) {
                double q[5];
                v2_exact_prefixes(ss_, sh, cc, false, nullptr, q);
                int rr[5], zu;
                const int n0 = p0 + cc * V2_CHUNK + lane * 4;
                v2_counts(&sh.ut, q, n0, N, &fd.side.pend[chain], rr, zu);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (n0 + i < N && rr[i + 1] > rr[i]) *aout = (rr[i] < zu) ? -1 : n0 + i;
              }
            }
          }
          v2_grid_barrier(vd.bar, bar_target);
        }
      }
      // (b) a failed seam check / record overflow / forced fallback: serial exact walk, consumers redone
      int anyf = anyf0;
      if (rare_ran) {  // the exact side scan may have raised a flag since
        anyf = 0;
        for (int ch = lane; ch < fd.nc; ch += 32)
          anyf |= (__ldcg(&fd.main.flag[ch + fd.nc * tpar]) | __ldcg(&fd.side.flag[ch + fd.nc * tpar])) & 1;
      }
      if (__any_sync(0xffffffffu, anyf)) {
        rare_ran = true;
        const int fm = __ldcg(fmain) & 1, fs = anc ? (__ldcg(fside) & 1) : 0;
        double* Wn = vd.Wn + (int64_t)chain * NP;
        double* qfb = vd.qin_fb + (int64_t)chain * nt * TPB;
        for (int pass = 0; pass < 2; ++pass) {  // pass 0: main weights, pass 1: ancestor weights
          const bool mine = pass == 0 ? fm != 0 : fs != 0;
          const V2Scan& sc = pass == 0 ? sm_ : ss_;
          if (mine) {  // dump this CTA's normalised weights (the stationary array still holds them)
            // (a side flag can only have been raised by the exact side scan above, which left W' = waN'/S4 in Pv)
            for (int i = tid; i < nch * V2_CHUNK; i += V2_NT) Wn[p0 + i] = sc.v[i];
          }
          v2_grid_barrier(vd.bar, bar_target);
          if (mine && lead) {
            if (tid == 0) {
              (pass == 0 ? fd.main.pend : fd.side.pend)[chain] = 0;
              if (pass == 0 && !last_only) fd.maxkey[chain] = 0ull;
            }
            v2_seq_exact(Wn, qfb, N, nt, (pass == 0 ? fd.main.counters : fd.side.counters) + 2 * chain,
                         reinterpret_cast<double*>(s_anc));  // the staging buffer is idle here
          }
          v2_grid_barrier(vd.bar, bar_target);
          if (mine) {
            // the stationary array must hold W for the re-walk: reload what was dumped
            for (int i = tid; i < nch * V2_CHUNK; i += V2_NT) sc.v[i] = __ldcg(Wn + p0 + i);
            const pgb_utable* utg = pass == 1 ? &fd.utab_anc[(int64_t)chain * T + t]
                                              : (last_only ? &fd.utab_star[chain] : &fd.utab_res[(int64_t)chain * T + t]);
            {
              const int* src = reinterpret_cast<const int*>(utg);
              int* dst = reinterpret_cast<int*>(&sh.ut);
              for (int i = tid; i < (int)(sizeof(pgb_utable) / 4); i += V2_NT) dst[i] = src[i];
            }
            __syncthreads();
            double lmax = -INFINITY;
            const bool index_only = (pass == 1) || last_only;
            int32_t* iout = pass == 1 ? fd.ah + ((int64_t)chain * T + t) * N + (N - 1) : vd.star_idx + chain;
            for (int cc = warp; cc < nch; cc += V2_NW) {
              double q[5];
              v2_exact_prefixes(sc, sh, cc, true, qfb, q);
              int rr[5], zu;
              const int n0 = p0 + cc * V2_CHUNK + lane * 4;
              v2_counts(&sh.ut, q, n0, N, &fd.status[chain], rr, zu);
              if (index_only) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (n0 + i < N && rr[i + 1] > rr[i]) *iout = (rr[i] < zu) ? -1 : n0 + i;
              } else {  // counts into the lane's TE slot, expanded by output below (balanced over the CTA)
                int* slot = reinterpret_cast<int*>(&TE[cc * 32 + lane]);
#pragma unroll
                for (int i = 0; i < 5; ++i) slot[i] = rr[i];
                slot[5] = zu;
                if (cc == 0 && lane == 0) sh.J0 = rr[0];
                if (cc == nch - 1 && lane == 31) sh.J1 = rr[4];
              }
            }
            if (!index_only) {
              __syncthreads();
              expand_pieces((int)sh.J0, (int)sh.J1, &fd.status[chain], lmax);
            }
            if (pass == 0 && !last_only) {
              if (ANC && lead && tid == 0) {  // the conditioned trajectory keeps slot N-1 (:160)
                const double* yt = fd.y + (int64_t)fd.n_y * t;
                RT* xt = r_x_row<RT>(fd, chain, t);
                RT x[NX];
#pragma unroll
                for (int d = 0; d < NX; ++d) {
                  x[d] = (RT)fd.xprim[(int64_t)chain * NX * T + d + NX * t];
                  xt[(int64_t)d * N + (N - 1)] = x[d];
                }
                const RT lw = r_log_weight<NX>(fd, x, yt, ivR);
                lwg[N - 1] = lw;
                lmax = fmax(lmax, (double)lw);
              }
              block_max(lmax);
            }
          }
          v2_grid_barrier(vd.bar, bar_target);
        }
      }
      // merge the status bits raised by consumer passes that were not redone; re-arm the other parity's flags
      if (lead && tid == 0) {
        const int p = rare_ran ? (__ldcg(&fd.main.pend[chain]) | __ldcg(&fd.side.pend[chain])) : pend0;
        if (p) atomicOr(&fd.status[chain], p);
        if (p || rare_ran) {
          fd.main.pend[chain] = 0;
          fd.side.pend[chain] = 0;
        }
        fd.main.flag[chain + fd.nc * (int)((t + 1) & 1)] = 0;
        fd.side.flag[chain + fd.nc * (int)((t + 1) & 1)] = 0;
      }
    }
    V2_TICK(10);
  }
#ifdef PGB_EXPERIMENT
  if (vd.prof && tid == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) vd.prof[(int64_t)blockIdx.x * 16 + i] = tacc[i];
  }
#endif
#undef V2_TICK
}

}  // namespace pgb
