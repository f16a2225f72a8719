// pgb_rollout.cuh — scenario sampling + forward rollout for solve_PG_OCP
// (Julia/src/optimal_control_Ipopt.jl:47-64 x_vec_0, :67-74 v_vec, :77-83 e_vec, :114-125 X_init/Y_init).
//
// One thread per scenario k keeps the evaluation order of the oracle (bit-exact), and the model
// matrices are read through a scenario-fastest transposed copy At[i][d][k] so that the K-wide loads
// of A_k coalesce.  Streams: scenario k uses Philox chain id k_offset + k, purpose ROLL with
// t = 0 (star uniform), 1 (x0 noise), 2 (process noise, j = step), 3 (measurement noise, j = step).
#pragma once
#include "pgb_filter.cuh"

namespace pgb {

struct RolloutDev {
  int64_t K, H, N;
  const double* At;    // [n_phi][n_x][K]   transposed A
  const double* Q;     // [K][n_x*n_x]
  const double* w_m1;  // [K][N]
  const double* x_m1;  // [K][n_x*N]  column-major n_x x N per scenario
  const double* u_m1;  // [K][n_u]
  const double* u_init;// [n_u*H]
  double Le[16];       // lower factor of the measurement noise, n_y x n_y
  uint64_t seed;
  uint32_t k_offset;
  double *x0, *v, *e, *X, *Y;  // [K][n_x], [K][H][n_x], [K][H][n_y], [K][H+1][n_x], [K][H][n_y]
  int* status;         // [1]
};

__global__ void k_transpose_A(const double* __restrict__ A, double* __restrict__ At, int64_t K, int m) {
  // A: [K][m] -> At: [m][K]
  __shared__ double tile[32][33];
  int64_t k0 = (int64_t)blockIdx.x * 32;
  int i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int64_t k = k0 + r;
    int i = i0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < K && i < m) ? A[k * m + i] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int i = i0 + r;
    int64_t k = k0 + threadIdx.x;
    if (k < K && i < m) At[(int64_t)i * K + k] = tile[threadIdx.x][r];
  }
}

// A_k * phi(x,u) with A read from the transposed copy (stride K between consecutive entries)
template <int NX>
__device__ __forceinline__ void eval_F_strided(const FilterDev& fd, const double* __restrict__ At, int64_t K,
                                               const double x[NX], const double* __restrict__ u, double F[NX]) {
  double acc[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) acc[d] = 0.0;
  if (fd.basis == 0) {
    double phi[5];
    phi[0] = __dmul_rn(0.1, x[0]);
    phi[1] = __dmul_rn(0.1, x[NX > 1 ? 1 : 0]);
    phi[2] = u[0];
    phi[3] = __dmul_rn(__dmul_rn(0.01, pgb_cos(__dmul_rn(3.0, x[0]))), x[NX > 1 ? 1 : 0]);
    phi[4] = __dmul_rn(__dmul_rn(0.1, pgb_sin(__dmul_rn(2.0, x[NX > 1 ? 1 : 0]))), u[0]);
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int d = 0; d < NX; ++d) acc[d] = __fma_rn(At[(int64_t)(i * NX + d) * K], phi[i], acc[d]);
  } else {
    const int nz = fd.n_z;
    double tk[PGB_MAX_Z_][MAXC];
    for (int k = 0; k < nz; ++k) {
      double zk = (k < fd.n_u) ? u[k] : x[k - fd.n_u];
      double zl = __dadd_rn(zk, fd.hg_L[k]);
      for (int j = 0; j < fd.hg_count[k]; ++j)
        tk[k][j] = __dmul_rn(fd.hg_lsi[k], pgb_sin(__dmul_rn(fd.hg_pij2l[k][j], zl)));
    }
    int dig[PGB_MAX_Z_];
    double pp[PGB_MAX_Z_];
    for (int k = 0; k < nz; ++k) dig[k] = 0;
    int from = 0;
    for (int i = 0; i < fd.n_phi; ++i) {
      for (int k = from; k < nz; ++k) pp[k] = __dmul_rn(k ? pp[k - 1] : 1.0, tk[k][dig[k]]);
      double p = pp[nz - 1];
#pragma unroll
      for (int d = 0; d < NX; ++d) acc[d] = __fma_rn(At[(int64_t)(i * NX + d) * K], p, acc[d]);
      int k = nz - 1;
      while (k >= 0 && ++dig[k] == fd.hg_count[k]) dig[k--] = 0;
      from = k < 0 ? 0 : k;
    }
  }
#pragma unroll
  for (int d = 0; d < NX; ++d) F[d] = acc[d];
}

template <int NX>
__global__ void __launch_bounds__(128) k_rollout(FilterDev fd, RolloutDev rd) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= rd.K) return;
  const int64_t K = rd.K, H = rd.H, N = rd.N;
  const int ny = fd.n_y, nu = fd.n_u;
  const uint32_t cid = rd.k_offset + (uint32_t)k;
  // chol(Q_k)
  double Lq[16];
  for (int i = 0; i < NX * NX; ++i) Lq[i] = rd.Q[k * NX * NX + i];
  if (small_chol(Lq, NX)) {
    atomicOr(rd.status, 16);
    return;
  }
  // star = systematic_resampling(w_m1, 1)  (:58): W = w/sum(w) (tree), first n with q_n >= u
  const double* w = rd.w_m1 + k * N;
  double st[40];
  int sp = 0;
  int64_t P = 1;
  while (P < N) P <<= 1;
  for (int64_t i = 0; i < P; ++i) {
    double xv = (i < N) ? w[i] : 0.0;
    int merges = __ffsll((long long)(i + 1)) - 1;
    for (int m = 0; m < merges; ++m) xv = __dadd_rn(st[--sp], xv);
    st[sp++] = xv;
  }
  const double S = st[0];
  pgb_u32x4 b = pgb_stream_block(rd.seed, PGB_STREAM_ROLL, cid, 0, 0, 0);
  const double us = pgb_u53_halfopen(b.v[0], b.v[1]);
  int64_t star = 0;
  {
    double q = 0.0;
    int64_t n = 0;
    while (q < us) {
      n = n + 1;
      if (n > N) { atomicOr(rd.status, 1); n = N; break; }
      q = __dadd_rn(q, __ddiv_rn(w[n - 1], S));
    }
    if (n == 0) { atomicOr(rd.status, 2); n = 1; }
    star = n - 1;
  }
  double x[NX], F[NX], z[4];
  for (int d = 0; d < NX; ++d) x[d] = rd.x_m1[(k * N + star) * NX + d];
  const double* At = rd.At + k;
  // x0 = f(x_m1, u_m1) + rand(MvNormal(0, Q))  (:62)
  eval_F_strided<NX>(fd, At, K, x, rd.u_m1 + k * nu, F);
  philox_normals<NX>(rd.seed, cid, 0, PGB_STREAM_ROLL, 1, 0, z);
  for (int d = 0; d < NX; ++d) {
    double acc = 0.0;
    for (int e2 = 0; e2 <= d; ++e2) acc = __fma_rn(Lq[d + NX * e2], z[e2], acc);
    x[d] = __dadd_rn(F[d], acc);
    if (rd.x0) rd.x0[k * NX + d] = x[d];
    if (rd.X) rd.X[(k * (H + 1)) * NX + d] = x[d];
  }
  for (int64_t t = 0; t < H; ++t) {
    double vv[NX], ee[4];
    philox_normals<NX>(rd.seed, cid, 0, PGB_STREAM_ROLL, 2, t, z);  // v_vec[:, t, k]  (:72)
    for (int d = 0; d < NX; ++d) {
      double acc = 0.0;
      for (int e2 = 0; e2 <= d; ++e2) acc = __fma_rn(Lq[d + NX * e2], z[e2], acc);
      vv[d] = acc;
      if (rd.v) rd.v[(k * H + t) * NX + d] = acc;
    }
    {  // e_vec[:, t, k]  (:81)
      const int npairs = (ny + 1) / 2;
      for (int p = 0; p < npairs; ++p) {
        double z0, z1;
        pgb_normal_pair(pgb_stream_block(rd.seed, PGB_STREAM_ROLL, cid, 0, 3, (uint32_t)(t * npairs + p)), &z0, &z1);
        z[2 * p] = z0;
        if (2 * p + 1 < ny) z[2 * p + 1] = z1;
      }
      for (int o = 0; o < ny; ++o) {
        double acc = 0.0;
        for (int e2 = 0; e2 <= o; ++e2) acc = __fma_rn(rd.Le[o + ny * e2], z[e2], acc);
        ee[o] = acc;
        if (rd.e) rd.e[(k * H + t) * ny + o] = acc;
      }
    }
    // Y[t] = g(X[t]) + e (:122), X[t+1] = f(X[t], u_t) + v (:121)
    for (int o = 0; o < ny; ++o) {
      double g = 0.0;
      for (int d = 0; d < NX; ++d) g = __fma_rn(fd.C[o + ny * d], x[d], g);
      if (rd.Y) rd.Y[(k * H + t) * ny + o] = __dadd_rn(g, ee[o]);
    }
    eval_F_strided<NX>(fd, At, K, x, rd.u_init + (int64_t)nu * t, F);
    for (int d = 0; d < NX; ++d) {
      x[d] = __dadd_rn(F[d], vv[d]);
      if (rd.X) rd.X[(k * (H + 1) + t + 1) * NX + d] = x[d];
    }
  }
}

}  // namespace pgb
