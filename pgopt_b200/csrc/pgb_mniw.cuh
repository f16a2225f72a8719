// pgb_mniw.cuh — sufficient statistics (Julia/src/particle_Gibbs.jl:221-225) and the MNIW parameter
// draw (particle_Gibbs.jl:56-81) on the device.
//
// Bit-exactness vs the oracle: every dense-algebra step is done in "right-looking" form, where each
// matrix entry receives its fma updates in the same (increasing / decreasing pivot) order as the
// oracle's sequential loops, so the parallel factorisations and triangular solves produce the same
// bits as oracle/pg_oracle.c.  Sums over time use the balanced-tree order of pgb_device.cuh.
#pragma once
#include "pgb_filter.cuh"

namespace pgb {

constexpr int MNIW_TPB = 1024;

struct MniwDev {
  int nc, n_x, n_phi;
  int64_t Tm1;
  double ell_Q;
  const double* V_diag;    // [n_phi]
  const double* Vinv;      // [n_phi*n_phi] I / V of a dense V (pgb_set_prior_dense), or null: V is diag(V_diag)
  const double* Lambda_Q;  // [n_x*n_x]
  double* zeta;   // [nc][n_x][Tm1]
  double* zb;     // [nc][n_phi][Tm1]
  double* Phi;    // [nc][n_x*n_x]
  double* Psi;    // [nc][n_x*n_phi]
  double* Sigma;  // [nc][n_phi*n_phi]
  double* Ls;     // [nc][n_phi*n_phi] workspaces
  double* Sinv;   // [nc][n_phi*n_phi]
  double* Lv;     // [nc][n_phi*n_phi]
  double* Mpost;  // [nc][n_x*n_phi]
  double* Xn;     // [nc][n_x*n_phi]
  double* LX;     // [nc][n_x*n_phi]
  double* A;      // [nc][n_x*n_phi]  chain state (in/out)
  double* Q;      // [nc][n_x*n_x]
  double* Lq;     // [nc][16]
  double* c0;     // [nc]
  int* status;    // [nc]
  int philox;
  uint64_t seed;
  uint32_t chain_base, sweep;
  const double* bartlett;  // injected [nc][n_x*n_x]
  const double* Xmn;       // injected [nc][n_x*n_phi]
};

// phi(x', u) for every t < T-1 and zeta = x'[:, 2:T]
template <int NX>
__global__ void __launch_bounds__(128) k_stats_z(FilterDev fd, MniwDev md) {
  const int chain = blockIdx.y;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t Tm1 = md.Tm1;
  if (t >= Tm1) return;
  const double* xp = fd.xprim + (int64_t)chain * NX * fd.T;
  const double* u = fd.u + (int64_t)fd.n_u * t;
  double* zeta = md.zeta + (int64_t)chain * NX * Tm1;
  double* zb = md.zb + (int64_t)chain * fd.n_phi * Tm1;
  double x[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) {
    x[d] = xp[d + NX * t];
    zeta[(int64_t)d * Tm1 + t] = xp[d + NX * (t + 1)];
  }
  if (fd.basis == 0) {
    zb[0 * Tm1 + t] = __dmul_rn(0.1, x[0]);
    zb[1 * Tm1 + t] = __dmul_rn(0.1, x[NX > 1 ? 1 : 0]);
    zb[2 * Tm1 + t] = u[0];
    zb[3 * Tm1 + t] = __dmul_rn(__dmul_rn(0.01, pgb_cos(__dmul_rn(3.0, x[0]))), x[NX > 1 ? 1 : 0]);
    zb[4 * Tm1 + t] = __dmul_rn(__dmul_rn(0.1, pgb_sin(__dmul_rn(2.0, x[NX > 1 ? 1 : 0]))), u[0]);
    return;
  }
  const int nz = fd.n_z;
  double tk[PGB_MAX_Z_][MAXC];
  for (int k = 0; k < nz; ++k) {
    double zk = (k < fd.n_u) ? u[k] : x[k - fd.n_u];
    double zl = __dadd_rn(zk, fd.hg_L[k]);
    for (int j = 0; j < fd.hg_count[k]; ++j)
      tk[k][j] = __dmul_rn(fd.hg_lsi[k], pgb_sin(__dmul_rn(fd.hg_pij2l[k][j], zl)));
  }
  int dig[PGB_MAX_Z_];
  for (int k = 0; k < nz; ++k) dig[k] = 0;
  for (int i = 0; i < fd.n_phi; ++i) {
    double p = 1.0;
    for (int k = 0; k < nz; ++k) p = __dmul_rn(p, tk[k][dig[k]]);
    zb[(int64_t)i * Tm1 + t] = p;
    int k = nz - 1;
    while (k >= 0 && ++dig[k] == fd.hg_count[k]) dig[k--] = 0;
  }
}

// One CTA per output entry: Phi = zeta zeta', Psi = zeta z', Sigma = z z' as tree sums over t.
static __global__ void __launch_bounds__(TPB) k_stats(MniwDev md) {
  __shared__ double sm[NWARP + 1];
  const int chain = blockIdx.y;
  const int nx = md.n_x, np = md.n_phi;
  const int64_t Tm1 = md.Tm1;
  const double* zeta = md.zeta + (int64_t)chain * nx * Tm1;
  const double* zb = md.zb + (int64_t)chain * np * Tm1;
  int64_t e = blockIdx.x;
  const double *pa, *pb;
  double* out;
  if (e < nx * nx) {
    int a = (int)(e % nx), b = (int)(e / nx);
    pa = zeta + (int64_t)a * Tm1; pb = zeta + (int64_t)b * Tm1;
    out = md.Phi + (int64_t)chain * nx * nx + e;
  } else if (e < nx * nx + (int64_t)nx * np) {
    e -= nx * nx;
    int a = (int)(e % nx), i = (int)(e / nx);
    pa = zeta + (int64_t)a * Tm1; pb = zb + (int64_t)i * Tm1;
    out = md.Psi + (int64_t)chain * nx * np + e;
  } else {
    e -= nx * nx + (int64_t)nx * np;
    int i = (int)(e % np), j = (int)(e / np);
    pa = zb + (int64_t)i * Tm1; pb = zb + (int64_t)j * Tm1;
    out = md.Sigma + (int64_t)chain * np * np + e;
  }
  int64_t P = TPB;
  while (P < Tm1) P <<= 1;
  const int per = (int)(P / TPB);
  const int64_t base = (int64_t)threadIdx.x * per;
  double st[16];
  int sp = 0;
  for (int i = 0; i < per; ++i) {
    double x = (base + i < Tm1) ? __dmul_rn(pa[base + i], pb[base + i]) : 0.0;
    int merges = __ffs(i + 1) - 1;
    for (int m = 0; m < merges; ++m) x = __dadd_rn(st[--sp], x);
    st[sp++] = x;
  }
  double s = block_tree_sum(st[0], sm);
  if (threadIdx.x == 0) *out = s;
}

// ---- small serial helpers (n <= 4), executed by one thread; same loops as the oracle's ----
__device__ inline int small_chol(double* A, int n) {
  for (int k = 0; k < n; ++k) {
    double d = A[k + n * k];
    if (!(d > 0.0)) return 1;
    d = __dsqrt_rn(d);
    A[k + n * k] = d;
    for (int i = k + 1; i < n; ++i) A[i + n * k] = __ddiv_rn(A[i + n * k], d);
    for (int j = k + 1; j < n; ++j)
      for (int i = j; i < n; ++i) A[i + n * j] = __fma_rn(-A[i + n * k], A[j + n * k], A[i + n * j]);
  }
  for (int j = 1; j < n; ++j)
    for (int i = 0; i < j; ++i) A[i + n * j] = 0.0;
  return 0;
}
__device__ inline void small_inverse_from_chol(const double* L, int n, double* inv) {
  double col[4];
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) col[i] = (i == j) ? 1.0 : 0.0;
    for (int d = 0; d < n; ++d) {
      double acc = col[d];
      for (int e = 0; e < d; ++e) acc = __fma_rn(-L[d + n * e], col[e], acc);
      col[d] = __ddiv_rn(acc, L[d + n * d]);
    }
    for (int d = n - 1; d >= 0; --d) {
      double acc = col[d];
      for (int e = n - 1; e > d; --e) acc = __fma_rn(-L[e + n * d], col[e], acc);
      col[d] = __ddiv_rn(acc, L[d + n * d]);
    }
    for (int i = 0; i < n; ++i) inv[i + n * j] = col[i];
  }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i) inv[i + n * j] = inv[j + n * i];
}
// Lq = chol(Q), c0 = -(n log 2pi + logdet Q)/2   (MvNormal(zeros, Q), particle_Gibbs.jl:157,178)
__device__ inline int chain_params_from_Q(const double* Q, int n, double* Lq16, double* c0) {
  double L[16];
  for (int i = 0; i < n * n; ++i) L[i] = Q[i];
  if (small_chol(L, n)) return 1;
  double logdet = 0.0;
  for (int d = 0; d < n; ++d) logdet = __dadd_rn(logdet, pgb_log(L[d + n * d]));
  logdet = __dmul_rn(2.0, logdet);
  *c0 = __ddiv_rn(-__dadd_rn(__dmul_rn((double)n, 1.8378770664093453), logdet), 2.0);
  for (int i = 0; i < 16; ++i) Lq16[i] = (i < n * n) ? L[i] : 0.0;
  return 0;
}
static __global__ void k_chain_params(int chain, int n_x, const double* Q, double* Lq, double* c0, int* status) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  if (chain_params_from_Q(Q + (int64_t)chain * n_x * n_x, n_x, Lq + chain * 16, c0 + chain))
    atomicOr(&status[chain], 16);
}

// Marsaglia-Tsang chi-square from the MNIW Philox stream (same consumption as the oracle's mt_chi2)
__device__ inline double mt_chi2_dev(uint64_t seed, uint32_t cid, uint32_t sweep, uint32_t which, double df) {
  double a = __ddiv_rn(df, 2.0);
  double d = __dsub_rn(a, __ddiv_rn(1.0, 3.0));
  double c = __ddiv_rn(1.0, __dsqrt_rn(__dmul_rn(9.0, d)));
  for (uint32_t attempt = 0;; ++attempt) {
    double x, unused;
    pgb_normal_pair(pgb_stream_block(seed, PGB_STREAM_MNIW, cid, sweep, which, 2 * attempt), &x, &unused);
    pgb_u32x4 b = pgb_stream_block(seed, PGB_STREAM_MNIW, cid, sweep, which, 2 * attempt + 1);
    double uu = pgb_u53(b.v[0], b.v[1]);
    double v = __dadd_rn(1.0, __dmul_rn(c, x));
    if (v <= 0.0) continue;
    v = __dmul_rn(__dmul_rn(v, v), v);
    // log(u) < 0.5*x*x + d - d*v + d*log(v)   (left-assoc, as written in the oracle)
    double rhs = __dadd_rn(__dsub_rn(__dadd_rn(__dmul_rn(__dmul_rn(0.5, x), x), d), __dmul_rn(d, v)),
                           __dmul_rn(d, pgb_log(v)));
    if (pgb_log(uu) < rhs) return __dmul_rn(2.0, __dmul_rn(d, v));
  }
}

// Right-looking Cholesky of an n x n column-major matrix (global or shared memory) by one CTA, in panels of CHOL_PB
// columns.  Every entry receives the same fma updates in the same (increasing pivot) order as the oracle's unblocked loops
// — a blocked schedule only changes WHEN an update is applied, never the chain an entry goes through — so the bits are
// the oracle's.  The first version synchronised the whole CTA three times per pivot (375 barriers per factorisation at
// n = 125, ~1 100 cycles each with the dependent sqrt / division in between: most of a 1.0 ms kernel).  Now warp 0
// factors a panel on its own (rows split over the lanes, __syncwarp between the dependent steps) and the CTA meets twice
// per panel for the trailing update, in which every thread applies the panel's CHOL_PB updates to its entries in order.
constexpr int CHOL_PB = 8;
__device__ inline int cta_cholesky(double* A, int n, int* s_fail) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int lx = tid & 31, wy = tid >> 5, nw = nth >> 5;
  __syncthreads();
  for (int k0 = 0; k0 < n; k0 += CHOL_PB) {
    const int kb = min(CHOL_PB, n - k0);
    if (wy == 0) {  // panel: columns k0 .. k0+kb-1, rows k0 .. n-1
      for (int k = k0; k < k0 + kb; ++k) {
        const double d = A[k + (int64_t)n * k];
        if (!(d > 0.0)) {  // warp-uniform
          if (lx == 0) *s_fail = 1;
          break;
        }
        const double sd = __dsqrt_rn(d);
        __syncwarp();
        if (lx == 0) A[k + (int64_t)n * k] = sd;
        for (int i = k + 1 + lx; i < n; i += 32) A[i + (int64_t)n * k] = __ddiv_rn(A[i + (int64_t)n * k], sd);
        __syncwarp();
        for (int j = k + 1; j < k0 + kb; ++j) {  // the panel's own trailing columns
          const double ljk = A[j + (int64_t)n * k];
          for (int i = j + lx; i < n; i += 32)
            A[i + (int64_t)n * j] = __fma_rn(-A[i + (int64_t)n * k], ljk, A[i + (int64_t)n * j]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (*s_fail) return 1;
    // trailing update: entry (i, j), i >= j >= k0 + kb, gets the panel's kb updates in pivot order
    for (int j = k0 + kb + wy; j < n; j += nw) {
      double lj[CHOL_PB];
#pragma unroll
      for (int q = 0; q < CHOL_PB; ++q) lj[q] = q < kb ? A[j + (int64_t)n * (k0 + q)] : 0.0;
      for (int i = j + lx; i < n; i += 32) {
        double acc = A[i + (int64_t)n * j];
#pragma unroll
        for (int q = 0; q < CHOL_PB; ++q)
          if (q < kb) acc = __fma_rn(-A[i + (int64_t)n * (k0 + q)], lj[q], acc);
        A[i + (int64_t)n * j] = acc;
      }
    }
    __syncthreads();
  }
  for (int j = 1 + wy; j < n; j += nw)
    for (int i = lx; i < j; i += 32) A[i + (int64_t)n * j] = 0.0;
  __syncthreads();
  return 0;
}
// Index of L[i, j] (i >= j): full column-major n x n, or the lower triangle packed column by column (shared-memory path:
// the n x n right-hand side block and the packed factor together fit 227 KB up to n = 135).  row_step(j): the offset from
// L[i, j] to L[i, j+1].
struct LIdxFull {
  int n;
  __device__ __forceinline__ int64_t operator()(int i, int j) const { return i + (int64_t)n * j; }
  __device__ __forceinline__ int row_step(int) const { return n; }
};
struct LIdxPacked {
  int n;
  __device__ __forceinline__ int64_t operator()(int i, int j) const { return (int64_t)j * n - ((int64_t)j * (j - 1)) / 2 + (i - j); }
  __device__ __forceinline__ int row_step(int j) const { return n - j - 1; }
};
// Solve (L L') X = B in place for two blocks of right-hand sides at once: the n1 columns of B1 (post_mean, :67) and the
// n2 columns of B2, which holds the IDENTITY on entry (left_cov = I / Sigbar, :68).  Right-hand sides are independent:
// ONE THREAD PER COLUMN walks the oracle's two substitutions on its own (row e divides by L[e,e], then updates the rows
// below / above in the oracle's order; L[., e] is a broadcast read), SOLVE_CPW columns per warp, no barrier inside.  The
// first version met with the whole CTA twice per row (1 000 barriers for the two solves); a warp-per-column version was
// instruction-bound on the one SM a chain's draw runs on (a 35-instruction exact division per row and column executed by
// 32 lanes for one useful result: 40 % of the kernel's instructions).  Column j of the identity is zero above row j and
// stays +0 through the forward sweep (0 / piv = +0, fma(-l, +0, +0) = +0), so its forward sweep starts at row j — the
// same bits with half the work.
constexpr int SOLVE_CPW = 16, SOLVE_UB = 8;
template <class IDX>
__device__ inline void cta_chol_solve2(const double* L, IDX ix, int n, double* B1, int n1, double* B2, int n2) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int lx = tid & 31, wy = tid >> 5, nw = nth >> 5;
  __syncthreads();
  const int ntot = n1 + n2;
  for (int cb = wy * SOLVE_CPW; cb < ntot; cb += nw * SOLVE_CPW) {  // warp-uniform
    const int c = cb + lx;
    double* col = nullptr;
    int start = n;
    if (lx < SOLVE_CPW) {
      if (c < n1) { col = B1 + (int64_t)n * c; start = 0; }
      else if (c < ntot) { col = B2 + (int64_t)n * (c - n1); start = c - n1; }
    }
    int first = start;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    for (int e = first; e < n; ++e) {  // forward: L y = b
      const double* Le = L + ix(e, e) - e;   // column e of the factor: Le[d] = L[d, e]
      const double piv = Le[e];
      if (e >= start) {
        const double be = __ddiv_rn(col[e], piv);
        col[e] = be;
        int d = e + 1;
        for (; d + SOLVE_UB <= n; d += SOLVE_UB) {  // loads first, then the stores: nothing waits on a possible alias
          double l[SOLVE_UB], v[SOLVE_UB];
#pragma unroll
          for (int q = 0; q < SOLVE_UB; ++q) { l[q] = Le[d + q]; v[q] = col[d + q]; }
#pragma unroll
          for (int q = 0; q < SOLVE_UB; ++q) col[d + q] = __fma_rn(-l[q], be, v[q]);
        }
        for (; d < n; ++d) col[d] = __fma_rn(-Le[d], be, col[d]);
      }
    }
    if (col) {
      for (int e = n - 1; e >= 0; --e) {  // backward: L' x = y
        const double* Lr = L + ix(e, 0);     // row e of the factor, walked along the columns
        const double be = __ddiv_rn(col[e], L[ix(e, e)]);
        col[e] = be;
        int64_t off = 0;
        int d = 0;
        for (; d + SOLVE_UB <= e; d += SOLVE_UB) {
          double l[SOLVE_UB], v[SOLVE_UB];
#pragma unroll
          for (int q = 0; q < SOLVE_UB; ++q) {
            l[q] = Lr[off];
            v[q] = col[d + q];
            off += ix.row_step(d + q);
          }
#pragma unroll
          for (int q = 0; q < SOLVE_UB; ++q) col[d + q] = __fma_rn(-l[q], be, v[q]);
        }
        for (; d < e; ++d) {
          col[d] = __fma_rn(-Lr[off], be, col[d]);
          off += ix.row_step(d);
        }
      }
    }
  }
  __syncthreads();
}

// bytes of dynamic shared memory the on-chip path of k_mniw needs (0: does not fit, the global-memory path runs)
__host__ __device__ inline size_t mniw_smem_bytes(int np) {
  const size_t b = ((size_t)np * np + (size_t)np * (np + 1) / 2 + (size_t)4 * np) * 8;  // work block, packed factor, n_x <= 4 right-hand sides
  return b <= (size_t)220 * 1024 ? b : 0;
}

// MNIW_sample + refresh of the chain parameters used by the next sweep.  One CTA per chain.
// n_phi <= 135 (both reference examples: 5 and 125): the n_phi x n_phi work matrix (Sigbar -> its factor -> the identity ->
// left_cov -> its factor) and the packed factor of Sigbar live in SHARED memory for the whole kernel — the two
// factorisations and the n_phi-right-hand-side solve are ~1 500 CTA barriers with a dependent memory round trip between
// each pair, which cost 2 ms against global memory and ~0.2 ms on chip.  Same operations in the same
// order on every entry, hence the same bits; larger bases take the global-memory path.
static __global__ void __launch_bounds__(MNIW_TPB) k_mniw(MniwDev md, int use_smem) {
  extern __shared__ double mn_sm[];
  __shared__ int s_fail;
  __shared__ double s_small[8][16];  // covM, Lc, S, B, LB, Wm, Qn, Lq
  const int chain = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int nx = md.n_x, np = md.n_phi;
  const int64_t pp = (int64_t)np * np;
  double* Ls = use_smem ? mn_sm : md.Ls + chain * pp;       // on chip the three stages reuse one n_phi x n_phi block
  double* Sinv = use_smem ? mn_sm : md.Sinv + chain * pp;
  double* Lv = use_smem ? mn_sm : md.Lv + chain * pp;
  double* Lp = mn_sm + pp;                                   // packed factor of Sigbar (on-chip path)
  double* Mp = md.Mpost + (int64_t)chain * nx * np;
  double* Xn = md.Xn + (int64_t)chain * nx * np;
  double* LX = use_smem ? Lp + (int64_t)np * (np + 1) / 2 : md.LX + (int64_t)chain * nx * np;  // [np x nx] / [nx x np] scratch
  const double* Phi = md.Phi + (int64_t)chain * nx * nx;
  const double* Psi = md.Psi + (int64_t)chain * nx * np;
  const double* Sigma = md.Sigma + chain * pp;
  double* Aout = md.A + (int64_t)chain * nx * np;
  double* Qout = md.Q + (int64_t)chain * nx * nx;
  const uint32_t cid = md.chain_base + chain;
  if (tid == 0) s_fail = 0;
  // Sigbar = Sigma + I/V (:64)
  for (int64_t e = tid; e < pp; e += nth) {
    int i = (int)(e % np), j = (int)(e / np);
    double v = Sigma[e];
    if (md.Vinv) v = __dadd_rn(v, md.Vinv[e]);
    else if (i == j) v = __dadd_rn(v, __ddiv_rn(1.0, md.V_diag[i]));
    Ls[e] = v;
  }
  if (cta_cholesky(Ls, np, &s_fail)) {
    if (tid == 0) atomicOr(&md.status[chain], 16);
    return;
  }
  // post_mean = Psibar / Sigbar (:67): RHS r = row r of Psi, stored as column r of a np x nx buffer (LX as scratch,
  // transposed into Mp afterwards); left_cov = I / Sigbar (:68): np right-hand sides holding the identity.  One joint
  // solve: the right-hand sides are independent and the warps own whole columns.
  for (int64_t e = tid; e < (int64_t)np * nx; e += nth) {
    int i = (int)(e % np), d = (int)(e / np);
    LX[e] = Psi[d + nx * i];
  }
  if (use_smem) {  // keep the factor, packed, before the block is reused for the inverse
    for (int64_t e = tid; e < pp; e += nth) {
      int i = (int)(e % np), j = (int)(e / np);
      if (i >= j) Lp[LIdxPacked{np}(i, j)] = Ls[e];
    }
    __syncthreads();
  }
  for (int64_t e = tid; e < pp; e += nth) Sinv[e] = ((e % np) == (e / np)) ? 1.0 : 0.0;
  if (use_smem) cta_chol_solve2(Lp, LIdxPacked{np}, np, LX, nx, Sinv, np);
  else cta_chol_solve2(Ls, LIdxFull{np}, np, LX, nx, Sinv, np);
  for (int64_t e = tid; e < (int64_t)np * nx; e += nth) {
    int i = (int)(e % np), d = (int)(e / np);
    Mp[d + nx * i] = LX[e];
  }
  __syncthreads();
  // cov_M = (Lambda_Q + Phibar) - M Psi' (:65), symmetrised (:66) -> s_small[1]
  if (tid < nx * nx) {
    int a = tid % nx, b = tid / nx;
    double acc = 0.0;
    for (int i = 0; i < np; ++i) acc = __fma_rn(Mp[a + nx * i], Psi[b + nx * i], acc);
    s_small[0][a + nx * b] = __dsub_rn(__dadd_rn(md.Lambda_Q[a + nx * b], Phi[a + nx * b]), acc);
  }
  __syncthreads();
  if (tid < nx * nx) {
    int a = tid % nx, b = tid / nx;
    s_small[1][a + nx * b] = __dmul_rn(0.5, __dadd_rn(s_small[0][a + nx * b], s_small[0][b + nx * a]));
  }
  // the inverse: lower triangle mirrored (LAPACK potri convention, as the oracle)
  for (int64_t e = tid; e < pp; e += nth) {
    int i = (int)(e % np), j = (int)(e / np);
    if (i < j) Sinv[e] = Sinv[j + (int64_t)np * i];
  }
  __syncthreads();
  // symmetrise (:69) and factor: chol(left_cov_sym)
  if (!use_smem) {
    for (int64_t e = tid; e < pp; e += nth) {
      int i = (int)(e % np), j = (int)(e / np);
      Lv[e] = __dmul_rn(0.5, __dadd_rn(Sinv[i + (int64_t)np * j], Sinv[j + (int64_t)np * i]));
    }
  }  // on chip Lv aliases Sinv, which the mirroring above made exactly symmetric: 0.5 * (a + a) == a
  if (cta_cholesky(Lv, np, &s_fail)) {
    if (tid == 0) atomicOr(&md.status[chain], 16);
    return;
  }
  // Q ~ InverseWishart(ell_Q + T n_x, cov_M_sym) (:72-73) — n_x x n_x, one thread
  if (tid == 0) {
    double* Lc = s_small[1];
    double* S = s_small[2];
    double* B = s_small[3];
    double* LB = s_small[4];
    double* Wm = s_small[5];
    double* Qn = s_small[6];
    double* Lq = s_small[7];
    int fail = 0;
    double df = __dadd_rn(md.ell_Q, __dmul_rn((double)md.Tm1, (double)nx));
    fail |= small_chol(Lc, nx);
    if (!fail) {
      small_inverse_from_chol(Lc, nx, S);
      fail |= small_chol(S, nx);
    }
    if (!fail) {
      for (int i = 0; i < nx * nx; ++i) B[i] = 0.0;
      if (md.philox) {
        for (int i = 0; i < nx; ++i)
          B[i + nx * i] = __dsqrt_rn(mt_chi2_dev(md.seed, cid, md.sweep, (uint32_t)i, __dsub_rn(df, (double)i)));
        int e = 0;
        for (int c = 0; c < nx; ++c)
          for (int r = c + 1; r < nx; ++r, ++e) {
            double z0, z1;
            pgb_normal_pair(pgb_stream_block(md.seed, PGB_STREAM_MNIW, cid, md.sweep, 200, (uint32_t)e), &z0, &z1);
            B[r + nx * c] = z0;
          }
      } else {
        const double* bi = md.bartlett + (int64_t)chain * nx * nx;
        for (int c = 0; c < nx; ++c)
          for (int r = c; r < nx; ++r) B[r + nx * c] = bi[r + nx * c];
      }
      for (int r = 0; r < nx; ++r)
        for (int c = 0; c < nx; ++c) {
          double acc = 0.0;
          for (int k = 0; k < nx; ++k) acc = __fma_rn(S[r + nx * k], B[k + nx * c], acc);
          LB[r + nx * c] = acc;
        }
      for (int r = 0; r < nx; ++r)
        for (int c = 0; c < nx; ++c) {
          double acc = 0.0;
          for (int k = 0; k < nx; ++k) acc = __fma_rn(LB[r + nx * k], LB[c + nx * k], acc);
          Wm[r + nx * c] = acc;
        }
      fail |= small_chol(Wm, nx);
    }
    if (!fail) {
      small_inverse_from_chol(Wm, nx, Qn);
      for (int i = 0; i < nx * nx; ++i) Lq[i] = Qn[i];
      fail |= small_chol(Lq, nx);
    }
    if (fail) s_fail = 1;
  }
  // matrix-normal normals X (n_x x n_phi, column-major) (:77)
  if (md.philox) {
    const int tot = nx * np;
    for (int p = tid; p < (tot + 1) / 2; p += nth) {
      double z0, z1;
      pgb_normal_pair(pgb_stream_block(md.seed, PGB_STREAM_MNIW, cid, md.sweep, 300, (uint32_t)p), &z0, &z1);
      Xn[2 * p] = z0;
      if (2 * p + 1 < tot) Xn[2 * p + 1] = z1;
    }
  } else {
    const double* xi = md.Xmn + (int64_t)chain * nx * np;
    for (int e = tid; e < nx * np; e += nth) Xn[e] = xi[e];
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) atomicOr(&md.status[chain], 16);
    return;
  }
  // A = M + chol(Q).L * X * chol(left_cov_sym).U (:76-77)
  const double* Lq = s_small[7];
  for (int e = tid; e < nx * np; e += nth) {
    int r = e % nx, i = e / nx;
    double acc = 0.0;
    for (int k = 0; k <= r; ++k) acc = __fma_rn(Lq[r + nx * k], Xn[k + nx * i], acc);
    LX[e] = acc;
  }
  __syncthreads();
  for (int e = tid; e < nx * np; e += nth) {
    int r = e % nx, i = e / nx;
    double acc = 0.0;
    for (int k = 0; k <= i; ++k) acc = __fma_rn(LX[r + nx * k], Lv[i + (int64_t)np * k], acc);
    Aout[e] = __dadd_rn(Mp[e], acc);
  }
  if (tid == 0) {
    for (int i = 0; i < nx * nx; ++i) Qout[i] = s_small[6][i];
    if (chain_params_from_Q(s_small[6], nx, md.Lq + chain * 16, md.c0 + chain)) atomicOr(&md.status[chain], 16);
  }
}

}  // namespace pgb
