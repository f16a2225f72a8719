// pgb_device.cuh — block-level building blocks shared by all kernels (sm_100a).
//
// Summation contract (matches oracle/pg_oracle.c orc_pairwise_sum): every "sum(v)" of the reference
// is a balanced binary tree over the index range padded with +0.0 to a power of two.  A thread
// owns 4 consecutive elements (two tree levels), a warp butterfly adds five levels, the eight warp
// totals of a 256-thread CTA three more: one CTA = one aligned 1024-element subtree ("tile").
// Tile sums are reduced by the same tree again.  FP add is commutative, so the xor-butterfly
// (both lanes compute a+b and b+a) yields the same bits as the recursive definition.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pgb_exact.h"

namespace pgb {

constexpr int PGB_MAX_Z_ = 8;     // = PGB_MAX_Z of the public header
#ifndef PGB_TPB
#define PGB_TPB 256
#endif
constexpr int TPB = PGB_TPB;      // threads per CTA
constexpr int IPT = 4;            // consecutive elements per thread
constexpr int TILE = TPB * IPT;   // elements per CTA tile (aligned power of two)
constexpr int NWARP = TPB / 32;
constexpr int RMAX = 32;          // complex-thread records kept per tile
constexpr int CMAX = 4096;        // complex threads resolved per scan (more -> serial fallback)

__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Balanced-tree sum of one value per thread over the CTA; every thread gets the result.
// `sm` must hold NWARP doubles.  Contains two __syncthreads().
__device__ __forceinline__ double block_tree_sum(double v, double* sm) {
  v = warp_tree_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  double r = (l < NWARP) ? sm[l] : 0.0;
#pragma unroll
  for (int o = 1; o < NWARP; o <<= 1) r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, o));
  return __shfl_sync(0xffffffffu, r, 0);
}

// pairwise tree over `per` (power of two) consecutive entries p[base .. base+per), entries >= n read as 0
__device__ __forceinline__ double local_tree(const double* __restrict__ p, int64_t base, int per, int64_t n) {
  double st[16];
  int sp = 0;
  for (int i = 0; i < per; ++i) {
    double x = (base + i < n) ? p[base + i] : 0.0;
    int merges = __ffs(i + 1) - 1;
    for (int m = 0; m < merges; ++m) x = __dadd_rn(st[--sp], x);
    st[sp++] = x;
  }
  return st[0];
}

// Balanced tree over an array of n partial sums (n <= TPB * 2^15), all threads get the result.
__device__ __forceinline__ double block_tree_sum_array(const double* __restrict__ p, int64_t n, double* sm) {
  int64_t P = TPB;
  while (P < n) P <<= 1;
  const int per = (int)(P / TPB);
  double v = (per == 1) ? ((threadIdx.x < n) ? p[threadIdx.x] : 0.0)
                        : local_tree(p, (int64_t)threadIdx.x * per, per, n);
  return block_tree_sum(v, sm);
}

// sum of a thread's four consecutive elements in tree order: (a+b) + (c+d)
__device__ __forceinline__ double tree4(const double v[4]) {
  return __dadd_rn(__dadd_rn(v[0], v[1]), __dadd_rn(v[2], v[3]));
}

// Deterministic (but otherwise arbitrary-order) exclusive FP scan across the CTA; used only for the
// *approximate* prefix that predicts binades.  sm: NWARP + 1 doubles.  *total = CTA sum (same order).
// The NWARP warp totals are scanned once by warp 0 (not redundantly by every thread).
__device__ __forceinline__ double block_excl_scan_approx(double v, double* sm, double* total) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (l >= o) inc = __dadd_rn(t, inc);
  }
  __syncthreads();
  if (l == 31) sm[w] = inc;
  __syncthreads();
  if (w == 0) {
    double x = (l < NWARP) ? sm[l] : 0.0;
    double xi = x;
#pragma unroll
    for (int o = 1; o < NWARP; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, xi, o);
      if (l >= o) xi = __dadd_rn(t, xi);
    }
    double ex = __shfl_up_sync(0xffffffffu, xi, 1);
    if (l == 0) ex = 0.0;
    if (l < NWARP) sm[l] = ex;
    if (l == NWARP - 1) sm[NWARP] = xi;
  }
  __syncthreads();
  double ex = __shfl_up_sync(0xffffffffu, inc, 1);
  if (l == 0) ex = 0.0;
  *total = sm[NWARP];
  return __dadd_rn(sm[w], ex);
}

__device__ __forceinline__ pgb_agg shfl_up_agg(pgb_agg a, int o) {
  pgb_agg r;
  r.f.c0 = __shfl_up_sync(0xffffffffu, (long long)a.f.c0, o);
  r.f.c1 = __shfl_up_sync(0xffffffffu, (long long)a.f.c1, o);
  r.cnt = __shfl_up_sync(0xffffffffu, a.cnt, o);
  r.key = __shfl_up_sync(0xffffffffu, a.key, o);
  return r;
}

// Exclusive segmented composition scan of one pgb_agg per thread across the CTA.
// sm: NWARP + 1 pgb_agg.  *total = combination of all TPB elements.
__device__ __forceinline__ pgb_agg block_excl_scan_agg(pgb_agg v, pgb_agg* sm, pgb_agg* total) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  pgb_agg inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    pgb_agg t = shfl_up_agg(inc, o);
    if (l >= o) inc = pgb_agg_combine(t, inc);
  }
  __syncthreads();
  if (l == 31) sm[w] = inc;
  __syncthreads();
  if (w == 0) {
    pgb_agg xi = (l < NWARP) ? sm[l] : pgb_agg_identity();
#pragma unroll
    for (int o = 1; o < NWARP; o <<= 1) {
      pgb_agg t = shfl_up_agg(xi, o);
      if (l >= o) xi = pgb_agg_combine(t, xi);
    }
    pgb_agg ex = shfl_up_agg(xi, 1);
    if (l == 0) ex = pgb_agg_identity();
    if (l < NWARP) sm[l] = ex;
    if (l == NWARP - 1) sm[NWARP] = xi;
  }
  __syncthreads();
  pgb_agg ex = shfl_up_agg(inc, 1);
  if (l == 0) ex = pgb_agg_identity();
  *total = sm[NWARP];
  return pgb_agg_combine(sm[w], ex);
}

// ---------------------------------------------------------------------------------------------
// Correctly rounded division by a divisor that is the same for many numerators (weight sums, R,
// Cholesky diagonal): the reciprocal is computed once with a true division, each quotient then costs
// one multiply and four fused multiply-adds instead of the ~30-instruction __ddiv_rn sequence.
//   q0 = RN(a y), r0 = RN(a - d q0), q1 = RN(q0 + r0 y)   -> q1 is a faithful quotient
//   r1 = a - d q1 (exact), q = RN(q1 + r1 y)              -> correctly rounded (Markstein 1990: y = RN(1/d)
//                                                            and q1 faithful imply q = RN(a/d))
// Operands outside a wide safe range (so that no intermediate can under/overflow), zeros, infinities and
// NaNs take the true division.  tests/test_gpu_parity.py::test_invariant_division_is_correctly_rounded
// compares against __ddiv_rn bit for bit on random and adversarial operands.
// ---------------------------------------------------------------------------------------------
struct InvDiv {
  double d, y;
  bool ok;
};
__device__ __forceinline__ InvDiv make_invdiv(double d) {
  InvDiv iv;
  iv.d = d;
  iv.y = __ddiv_rn(1.0, d);
  const double ad = fabs(d);
  iv.ok = (ad >= 0x1p-500) && (ad <= 0x1p500);
  return iv;
}
static __device__ __noinline__ double ddiv_slow(double a, double d) { return __ddiv_rn(a, d); }
__device__ __forceinline__ double div_inv(double a, const InvDiv& iv) {
  const double aa = fabs(a);
  if (iv.ok && aa >= 0x1p-500 && aa <= 0x1p500) {
    double q0 = __dmul_rn(a, iv.y);
    double r0 = __fma_rn(-iv.d, q0, a);
    double q1 = __fma_rn(r0, iv.y, q0);
    double r1 = __fma_rn(-iv.d, q1, a);
    return __fma_rn(r1, iv.y, q1);
  }
  return ddiv_slow(a, iv.d);
}

// order-preserving map double -> uint64 (for atomicMax over doubles)
__device__ __forceinline__ unsigned long long ordered_key(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_unkey(unsigned long long k) {
  unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// "last CTA done" ticket: returns true in exactly one CTA (the last to arrive) of the given group.
// Callers must have written their partial results before calling; the winner may read all of them.
__device__ __forceinline__ bool last_block_arrives(unsigned int* ticket, unsigned int nblocks, int* sm_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int prev = atomicAdd(ticket, 1u);
    *sm_flag = (prev == nblocks - 1);
    if (prev == nblocks - 1) *ticket = 0;  // re-arm for the next launch
  }
  __syncthreads();
  bool last = (*sm_flag != 0);
  if (last) __threadfence();
  return last;
}

}  // namespace pgb
