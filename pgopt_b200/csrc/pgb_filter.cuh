// pgb_filter.cuh — the CPF-AS time loop of particle_Gibbs (Julia/src/particle_Gibbs.jl:160-200) as
// per-step kernels.  Data layout (per chain): x history SoA xh[t][d][n], Int32 0-based ancestors
// ah[t][n], rolling weight buffers.  Particle slot N-1 is the conditioned trajectory (:160).
//
// Per step t >= 1 (k > 1 branch, :170-181):
//   k_prep     w_{t-1} = e/S1 (:199), F_i = A*phi(x_{t-1}^i,u_{t-1}) for all i (:175,:179),
//              waN_i = w_i * N(x'_t; F_i, Q) (:179); tile sums of w and waN
//   k_norm     waN ./= sum(waN) (:180); tile sums again (the resampler renormalises, :19)
//   scan pipeline on w   -> offspring expansion: a[t,j], x_t^j = F_{a_j} + L z_j (:172-175),
//              log-weights (:194) and their running max
//   scan pipeline on waN -> a[t,N] (:181)
//   k_weights  e = exp(lw - max) (:198) and tile sums
#pragma once
#include "pgb_scan.cuh"
#include "pgb_vmath.cuh"

namespace pgb {

constexpr int MAXC = 8;  // max basis functions per augmented dimension (Hilbert basis)

struct FilterDev {
  int64_t N, T;
  int nc, n_x, n_u, n_y, n_phi, basis, nt, n_z;
  int f32;           // PGB_PRECISION_F32: xh / Fbuf / lwbuf hold FP32 values (pgb_real.cuh); everything else FP64
  int hg_count[PGB_MAX_Z_];
  double hg_L[PGB_MAX_Z_], hg_lsi[PGB_MAX_Z_];
  double hg_pij2l[PGB_MAX_Z_][MAXC];
  double C[16], R[16], Rchol[16];  // n_y x n_x, n_y x n_y (column-major)
  double x0_mean[4], x0_chol[16];
  const double* u;   // [n_u*T]
  const double* y;   // [n_y*T]
  double* A;         // [nc][n_x*n_phi]
  double* Lq;        // [nc][16]  lower Cholesky of Q, column-major n_x x n_x
  double* c0;        // [nc]      -(n_x log 2pi + logdet Q)/2
  double* xprim;     // [nc][n_x*T]
  double* xh;        // [nc][x_rows][n_x][N]: particle states, row xrow(t) holds x_t
  int64_t x_rows;    // T: the whole history is kept; 2: rolling (replay mode, see k_trace_replay)
  int32_t* lineage;  // [nc][T] traced ancestor path b_t (replay mode)
  int32_t* ah;       // [nc][T][N]
  double *wbuf, *ebuf, *lwbuf, *Fbuf, *waN;  // [nc][N] ([nc][n_x][N] for Fbuf)
  double* whist;     // [nc][T][N] or null
  unsigned long long* maxkey;  // [nc]
  int* status;       // [nc]
  double* S1;        // [nc] sum of e
  double* S3;        // [nc] first sum of waN
  double* tsum_e;    // [nc][nt]
  double* tsum_a;    // [nc][nt]
  unsigned int* ticket_e; // [nc]
  unsigned int* ticket_a; // [nc]
  // random streams
  int philox;
  uint64_t seed;
  uint32_t chain_base;
  uint32_t sweep;    // k of the running sweep
  const double *Z0, *Z, *Ures, *Uanc, *Ustar;  // injected: [nc][n_x*(N-1)], [nc][T][n_x*N], [nc][T], [nc][T], [nc]
  pgb_utable *utab_res, *utab_anc, *utab_star; // [nc][T], [nc][T], [nc]
  ScanWS main, side;
};

// row of fd.xh that holds the particles of step t (0-based)
__device__ __forceinline__ double* x_row(const FilterDev& fd, int chain, int64_t t) {
  const int64_t r = (fd.x_rows == fd.T) ? t : (t & 1);
  return fd.xh + ((int64_t)chain * fd.x_rows + r) * fd.n_x * fd.N;
}

// ------------------------------------------------------------------------------------------
// model pieces (device restatement; the oracle has its own independent copy of the formulas)
// ------------------------------------------------------------------------------------------
// F = A * phi for the Hilbert-space basis, given the scaled sines tk[k][j] of every augmented dimension:
// phi_i = ((1*t_0)*t_1)*... with dimension 0 slowest, F_d = fma(A[d,i], phi_i, F_d) in increasing i.
template <int NX>
__device__ __forceinline__ void hilbert_apply(const FilterDev& fd, const double* __restrict__ A_sm,
                                              const double (*tk)[MAXC], double F[NX]) {
  const int nz = fd.n_z;
  int dig[PGB_MAX_Z_];
  double pp[PGB_MAX_Z_];
  double acc[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) acc[d] = 0.0;
  for (int k = 0; k < nz; ++k) dig[k] = 0;
  int from = 0;
  for (int i = 0; i < fd.n_phi; ++i) {
    for (int k = from; k < nz; ++k) pp[k] = __dmul_rn(k ? pp[k - 1] : 1.0, tk[k][dig[k]]);
    double p = pp[nz - 1];
#pragma unroll
    for (int d = 0; d < NX; ++d) acc[d] = __fma_rn(A_sm[d + NX * i], p, acc[d]);
    int k = nz - 1;
    while (k >= 0 && ++dig[k] == fd.hg_count[k]) dig[k--] = 0;
    from = k < 0 ? 0 : k;
  }
#pragma unroll
  for (int d = 0; d < NX; ++d) F[d] = acc[d];
}

template <int NX>
__device__ __forceinline__ void eval_F(const FilterDev& fd, const double* __restrict__ A_sm, const double x[NX],
                                       const double* __restrict__ u, double F[NX]) {
  if (fd.basis == 0) {
    // examples/PG_OCP_known_basis_functions_Altro.jl:34
    double phi[5];
    phi[0] = __dmul_rn(0.1, x[0]);
    phi[1] = __dmul_rn(0.1, x[NX > 1 ? 1 : 0]);
    phi[2] = u[0];
    phi[3] = __dmul_rn(__dmul_rn(0.01, pgb_cos(__dmul_rn(3.0, x[0]))), x[NX > 1 ? 1 : 0]);
    phi[4] = __dmul_rn(__dmul_rn(0.1, pgb_sin(__dmul_rn(2.0, x[NX > 1 ? 1 : 0]))), u[0]);
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) acc = __fma_rn(A_sm[d + NX * i], phi[i], acc);
      F[d] = acc;
    }
    return;
  }
  // Hilbert-space basis (examples/PG_OCP_generic_basis_functions_Altro.jl:85-99): per dimension
  // only count[k] distinct sines exist; phi_i = ((1*t_0)*t_1)*...  with dimension 0 slowest.
  const int nz = fd.n_z;
  double tk[PGB_MAX_Z_][MAXC];
  for (int k = 0; k < nz; ++k) {
    double zk = (k < fd.n_u) ? u[k] : x[k - fd.n_u];
    double zl = __dadd_rn(zk, fd.hg_L[k]);
    for (int j = 0; j < fd.hg_count[k]; ++j)
      tk[k][j] = __dmul_rn(fd.hg_lsi[k], pgb_sin(__dmul_rn(fd.hg_pij2l[k][j], zl)));
  }
  hilbert_apply<NX>(fd, A_sm, tk, F);
}

// Hilbert-space basis with 5 x 5 x 5 functions over z = [u; x1; x2] (the reference's generic example), two particles
// at once: the 15 sines of each particle are evaluated 5-wide, the 125 tensor-product terms are fully unrolled with
// the per-dimension factors in registers, and every A entry read from shared memory feeds both particles.  Same
// products and the same fma order as eval_F (i increasing), hence the same bits.
template <int NX>
__device__ __noinline__ void eval_F_hilbert355_pair(const FilterDev& fd, const double* __restrict__ A_sm,
                                                       const double xa[NX], const double xb[NX],
                                                       const double* __restrict__ u, double Fa[NX], double Fb2[NX]) {
  double t[2][3][5];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double zk = (k < fd.n_u) ? u[k] : (p == 0 ? xa[(k - fd.n_u) < NX ? (k - fd.n_u) : 0] : xb[(k - fd.n_u) < NX ? (k - fd.n_u) : 0]);
      const double zl = __dadd_rn(zk, fd.hg_L[k]);
      double arg[5], sn[5], cs[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) arg[j] = __dmul_rn(fd.hg_pij2l[k][j], zl);
      vsincos_wide<5>(arg, sn, cs);
#pragma unroll
      for (int j = 0; j < 5; ++j) t[p][k][j] = __dmul_rn(fd.hg_lsi[k], sn[j]);
    }
  double acc[2][NX];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int d = 0; d < NX; ++d) acc[p][d] = 0.0;
#pragma unroll
  for (int j0 = 0; j0 < 5; ++j0)
#pragma unroll
    for (int j1 = 0; j1 < 5; ++j1) {
      const double pa = __dmul_rn(t[0][0][j0], t[0][1][j1]);
      const double pb = __dmul_rn(t[1][0][j0], t[1][1][j1]);
#pragma unroll
      for (int j2 = 0; j2 < 5; ++j2) {
        const int i = (j0 * 5 + j1) * 5 + j2;
        const double qa = __dmul_rn(pa, t[0][2][j2]);
        const double qb = __dmul_rn(pb, t[1][2][j2]);
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          const double a = A_sm[d + NX * i];
          acc[0][d] = __fma_rn(a, qa, acc[0][d]);
          acc[1][d] = __fma_rn(a, qb, acc[1][d]);
        }
      }
    }
#pragma unroll
  for (int d = 0; d < NX; ++d) {
    Fa[d] = acc[0][d];
    Fb2[d] = acc[1][d];
  }
}

// The same specialisation for ONE particle (warp-per-chain kernel, pgb_small.cuh): branch-free 5-wide sines, the 125
// tensor-product terms fully unrolled, factors in registers.  Same products, same fma order, same bits as eval_F.
template <int NX>
__device__ __forceinline__ void eval_F_hilbert355_one(const FilterDev& fd, const double* __restrict__ A_sm,
                                                      const double x[NX], const double* __restrict__ u, double F[NX]) {
  double t[3][5];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double zk = (k < fd.n_u) ? u[k] : x[(k - fd.n_u) < NX ? (k - fd.n_u) : 0];
    const double zl = __dadd_rn(zk, fd.hg_L[k]);
    double arg[5], sn[5], cs[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) arg[j] = __dmul_rn(fd.hg_pij2l[k][j], zl);
    vsincos_wide<5>(arg, sn, cs);
#pragma unroll
    for (int j = 0; j < 5; ++j) t[k][j] = __dmul_rn(fd.hg_lsi[k], sn[j]);
  }
  double acc[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) acc[d] = 0.0;
#pragma unroll
  for (int j0 = 0; j0 < 5; ++j0)
#pragma unroll
    for (int j1 = 0; j1 < 5; ++j1) {
      const double pa = __dmul_rn(t[0][j0], t[1][j1]);
#pragma unroll
      for (int j2 = 0; j2 < 5; ++j2) {
        const int i = (j0 * 5 + j1) * 5 + j2;
        const double qa = __dmul_rn(pa, t[2][j2]);
#pragma unroll
        for (int d = 0; d < NX; ++d) acc[d] = __fma_rn(A_sm[d + NX * i], qa, acc[d]);
      }
    }
#pragma unroll
  for (int d = 0; d < NX; ++d) F[d] = acc[d];
}

// F for the thread's four particles at once (known basis: the four sin/cos chains are interleaved)
template <int NX>
__device__ __forceinline__ void eval_F4(const FilterDev& fd, const double* __restrict__ A_sm, const double (&x)[4][NX],
                                        const double* __restrict__ u, double (&F)[4][NX]) {
  if (fd.basis == 0) {
    double a3[4], a2[4], s3[4], c3[4], s2[4], c2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a3[k] = __dmul_rn(3.0, x[k][0]);
      a2[k] = __dmul_rn(2.0, x[k][NX > 1 ? 1 : 0]);
    }
    vsincos<4>(a3, s3, c3);
    vsincos<4>(a2, s2, c2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double phi[5];
      phi[0] = __dmul_rn(0.1, x[k][0]);
      phi[1] = __dmul_rn(0.1, x[k][NX > 1 ? 1 : 0]);
      phi[2] = u[0];
      phi[3] = __dmul_rn(__dmul_rn(0.01, c3[k]), x[k][NX > 1 ? 1 : 0]);
      phi[4] = __dmul_rn(__dmul_rn(0.1, s2[k]), u[0]);
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) acc = __fma_rn(A_sm[d + NX * i], phi[i], acc);
        F[k][d] = acc;
      }
    }
    return;
  }
  if (NX == 2 && fd.n_z == 3 && fd.hg_count[0] == 5 && fd.hg_count[1] == 5 && fd.hg_count[2] == 5) {
#pragma unroll 1
    for (int k = 0; k < 4; k += 2) eval_F_hilbert355_pair<NX>(fd, A_sm, x[k], x[k + 1], u, F[k], F[k + 1]);
    return;
  }
#pragma unroll 1
  for (int k = 0; k < 4; ++k) eval_F<NX>(fd, A_sm, x[k], u, F[k]);
}

// same as log_weight below with the division by the scalar R done through a precomputed reciprocal (n_y == 1)
template <int NX>
__device__ __forceinline__ double log_weight(const FilterDev& fd, const double x[NX], const double* __restrict__ y);
template <int NX>
__device__ __forceinline__ double log_weight_iv(const FilterDev& fd, const double x[NX], const double* __restrict__ y,
                                                const InvDiv& ivR) {
  if (fd.n_y != 1) return log_weight<NX>(fd, x, y);
  double g = 0.0;
#pragma unroll
  for (int d = 0; d < NX; ++d) g = __fma_rn(fd.C[d], x[d], g);
  double dd = __dsub_rn(g, y[0]);
  return div_inv(__dmul_rn(-__dmul_rn(dd, dd), 0.5), ivR);
}

template <int NX>
__device__ __forceinline__ double log_weight(const FilterDev& fd, const double x[NX], const double* __restrict__ y) {
  const int ny = fd.n_y;
  if (ny == 1) {
    double g = 0.0;
#pragma unroll
    for (int d = 0; d < NX; ++d) g = __fma_rn(fd.C[d], x[d], g);
    double dd = __dsub_rn(g, y[0]);
    return __ddiv_rn(__dmul_rn(-__dmul_rn(dd, dd), 0.5), fd.R[0]);  // :194  -(d)^2 / 2 / R  (x/2 == x*0.5 exactly)
  }
  double r[4];
  for (int o = 0; o < ny; ++o) {
    double g = 0.0;
#pragma unroll
    for (int d = 0; d < NX; ++d) g = __fma_rn(fd.C[o + ny * d], x[d], g);
    r[o] = __dsub_rn(y[o], g);
  }
  double s = 0.0;
  for (int o = 0; o < ny; ++o) {  // forward substitution with chol(R), then |.|^2   (:196)
    double acc = r[o];
    for (int e = 0; e < o; ++e) acc = __fma_rn(-fd.Rchol[o + ny * e], r[e], acc);
    r[o] = __ddiv_rn(acc, fd.Rchol[o + ny * o]);
    s = __fma_rn(r[o], r[o], s);
  }
  return __dmul_rn(-s, 0.5);
}

// standard normals for output slot j (all NX dims): dims (2p, 2p+1) come from Philox block j*NP + p
template <int NX>
__device__ __forceinline__ void philox_normals(uint64_t seed, uint32_t cid, uint32_t sweep, uint32_t purpose,
                                               uint32_t t, int64_t j, double z[NX]) {
  constexpr int NP = (NX + 1) / 2;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    double z0, z1;
    pgb_normal_pair(pgb_stream_block(seed, purpose, cid, sweep, t, (uint32_t)(j * NP + p)), &z0, &z1);
    z[2 * p] = z0;
    if (2 * p + 1 < NX) z[2 * p + 1] = z1;
  }
}
template <int NX>
__device__ __forceinline__ void draw_normals(const FilterDev& fd, int chain, uint32_t purpose, uint32_t t, int64_t j,
                                             const double* __restrict__ inj, double z[NX]) {
  if (fd.philox) {
    philox_normals<NX>(fd.seed, fd.chain_base + chain, fd.sweep, purpose, t, j, z);
  } else {
#pragma unroll
    for (int d = 0; d < NX; ++d) z[d] = inj[j * NX + d];
  }
}

__device__ __forceinline__ void block_max_to_global(double v, unsigned long long* dst, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = sm[0];
    for (int i = 1; i < NWARP; ++i) m = fmax(m, sm[i]);
    atomicMax(dst, ordered_key(m));
  }
}

// ------------------------------------------------------------------------------------------
// t = 0: initial particles (:160-165) and their log-weights
// ------------------------------------------------------------------------------------------
template <int NX>
__global__ void __launch_bounds__(TPB) k_init(FilterDev fd, int first) {
  __shared__ double sm[NWARP + 1];
  const int chain = blockIdx.y;
  const int64_t N = fd.N;
  double* x0 = x_row(fd, chain, 0);  // step 0
  const double* xp = fd.xprim + (int64_t)chain * NX * fd.T;
  double lmax = -INFINITY;
  int64_t n = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (n < N) {
    double x[NX];
    if (n == N - 1) {
#pragma unroll
      for (int d = 0; d < NX; ++d) x[d] = xp[d];  // x_pf[:, end, 1] = x_prim[:, 1]
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
    lmax = lw;
    if (!(lw == lw)) atomicOr(&fd.status[chain], 4);
  }
  block_max_to_global(lmax, &fd.maxkey[chain], sm);
}

// ------------------------------------------------------------------------------------------
// e = exp(lw - max), tile sums; last CTA: S1
// ------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(TPB) k_weights(FilterDev fd) {
  __shared__ double sm[NWARP + 1];
  __shared__ int sflag;
  const int chain = blockIdx.y, tile = blockIdx.x;
  const int64_t N = fd.N;
  const double m = ordered_unkey(fd.maxkey[chain]);
  const double* lw = fd.lwbuf + (int64_t)chain * N;
  double* e = fd.ebuf + (int64_t)chain * N;
  int64_t base = (int64_t)tile * TILE + (int64_t)threadIdx.x * IPT;
  double v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < N) {
      v[i] = pgb_exp(__dsub_rn(lw[base + i], m));
      e[base + i] = v[i];
    } else {
      v[i] = 0.0;
    }
  }
  double s = block_tree_sum(tree4(v), sm);
  if (threadIdx.x == 0) fd.tsum_e[(int64_t)chain * fd.nt + tile] = s;
  if (last_block_arrives(&fd.ticket_e[chain], gridDim.x, &sflag)) {
    double S = block_tree_sum_array(fd.tsum_e + (int64_t)chain * fd.nt, fd.nt, sm);
    if (threadIdx.x == 0) fd.S1[chain] = S;
  }
}

// ------------------------------------------------------------------------------------------
// k_prep: w = e/S1, F = A*phi(x_{t-1}, u_{t-1}), ancestor weights; tile sums of w (and waN).
// Also used after the last step (t == T) to finalise w_T only (no F / waN).
// ------------------------------------------------------------------------------------------
template <int NX>
__global__ void __launch_bounds__(TPB) k_prep(FilterDev fd, int64_t t, int first) {
  extern __shared__ double A_sm[];
  __shared__ double sm[NWARP + 1];
  __shared__ int sflag;
  const int chain = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const int64_t N = fd.N, T = fd.T;
  const bool last_only = (t >= T);  // finalise w_T
  const bool anc = !first && !last_only;
  if (!last_only) {
    const double* Ag = fd.A + (int64_t)chain * NX * fd.n_phi;
    for (int i = tid; i < NX * fd.n_phi; i += TPB) A_sm[i] = Ag[i];
  }
  __syncthreads();
  const double S1 = fd.S1[chain];
  const double* e = fd.ebuf + (int64_t)chain * N;
  double* w = fd.wbuf + (int64_t)chain * N;
  double* wh = fd.whist ? fd.whist + ((int64_t)chain * T + (t - 1)) * N : nullptr;
  const double* xp = x_row(fd, chain, t - 1);
  double* Fb = fd.Fbuf + (int64_t)chain * NX * N;
  double* wa = fd.waN + (int64_t)chain * N;
  const double* up = fd.u + (int64_t)fd.n_u * (t - 1);
  double Lq[NX * NX], xref[NX];
  double c0 = 0.0;
  if (anc) {
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Lq[i] = fd.Lq[chain * 16 + i];
#pragma unroll
    for (int d = 0; d < NX; ++d) xref[d] = fd.xprim[(int64_t)chain * NX * T + d + NX * t];
    c0 = fd.c0[chain];
  }
  int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
  double wv[4], av[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t n = base + i;
    wv[i] = 0.0;
    av[i] = 0.0;
    if (n < N) {
      double wi = __ddiv_rn(e[n], S1);  // :199
      wv[i] = wi;
      w[n] = wi;
      if (wh) wh[n] = wi;
      if (!last_only) {
        double x[NX], F[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) x[d] = xp[(int64_t)d * N + n];
        eval_F<NX>(fd, A_sm, x, up, F);
#pragma unroll
        for (int d = 0; d < NX; ++d) Fb[(int64_t)d * N + n] = F[d];
        if (anc) {
          // pdf(MvNormal(x'_t, Q), F) = exp(c0 - |L^-1 (F - x'_t)|^2 / 2)    (:178-179)
          double r[NX], sq = 0.0;
#pragma unroll
          for (int d = 0; d < NX; ++d) {
            double acc = __dsub_rn(F[d], xref[d]);
#pragma unroll
            for (int e2 = 0; e2 < d; ++e2) acc = __fma_rn(-Lq[d + NX * e2], r[e2], acc);
            r[d] = __ddiv_rn(acc, Lq[d + NX * d]);
            sq = __fma_rn(r[d], r[d], sq);
          }
          double a = __dmul_rn(wi, pgb_exp(__dsub_rn(c0, __ddiv_rn(sq, 2.0))));
          av[i] = a;
          wa[n] = a;
        }
      }
    }
  }
  double sw = block_tree_sum(tree4(wv), sm);
  if (tid == 0) fd.main.tsum[(int64_t)chain * fd.nt + tile] = sw;
  if (anc) {
    double sa = block_tree_sum(tree4(av), sm);
    if (tid == 0) fd.tsum_a[(int64_t)chain * fd.nt + tile] = sa;
  }
  if (last_block_arrives(&fd.main.ticket[chain], gridDim.x, &sflag)) {
    finalize_tile_sums(fd.main.tsum + (int64_t)chain * fd.nt, fd.nt, &fd.main.S[chain],
                       fd.main.tile_prefix + (int64_t)chain * (fd.nt + 1), sm);
    if (anc) {
      double S3 = block_tree_sum_array(fd.tsum_a + (int64_t)chain * fd.nt, fd.nt, sm);
      if (tid == 0) fd.S3[chain] = S3;
    }
    if (tid == 0) {
      fd.main.flag[chain] = 0;
      fd.maxkey[chain] = 0ull;  // all weights of step t-1 have been consumed by now
    }
  }
}

// waN ./= sum(waN) (:180) in place, then tile sums for the resampler's own normalisation (:19)
static __global__ void __launch_bounds__(TPB) k_norm(FilterDev fd) {
  __shared__ double sm[NWARP + 1];
  __shared__ int sflag;
  const int chain = blockIdx.y, tile = blockIdx.x;
  const int64_t N = fd.N;
  const double S3 = fd.S3[chain];
  double* wa = fd.waN + (int64_t)chain * N;
  int64_t base = (int64_t)tile * TILE + (int64_t)threadIdx.x * IPT;
  double v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = 0.0;
    if (base + i < N) {
      v[i] = __ddiv_rn(wa[base + i], S3);
      wa[base + i] = v[i];
    }
  }
  double s = block_tree_sum(tree4(v), sm);
  if (threadIdx.x == 0) fd.side.tsum[(int64_t)chain * fd.nt + tile] = s;
  if (last_block_arrives(&fd.side.ticket[chain], gridDim.x, &sflag)) {
    finalize_tile_sums(fd.side.tsum + (int64_t)chain * fd.nt, fd.nt, &fd.side.S[chain],
                       fd.side.tile_prefix + (int64_t)chain * (fd.nt + 1), sm);
    if (threadIdx.x == 0) fd.side.flag[chain] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// Consumers of the exact prefixes.
//   MODE_PROPAGATE: M draws from N weights -> a[t,j], x_t^j = F_{a_j} + L z_j, lw_j, running max
//   MODE_INDEX:     M draws -> 0-based indices into out_idx[0..M)  (ancestor of slot N, star,
//                   stand-alone systematic_resampling)
// Output j (1-based draw) belongs to element n iff q_{n-1} < u_j <= q_n  (the `while q < u` rule).
// ------------------------------------------------------------------------------------------
enum { MODE_PROPAGATE = 0, MODE_INDEX = 1 };

template <int NX, int MODE>
__global__ void __launch_bounds__(TPB) k_expand(FilterDev fd, ScanWS ws, const double* __restrict__ v,
                                                const pgb_utable* __restrict__ utab_base, int64_t utab_stride,
                                                int64_t utab_index, int32_t* __restrict__ out_idx,
                                                int64_t out_stride, int64_t t, int first, int fallback_pass) {
  __shared__ double sQ[TPB];
  __shared__ double sm[NWARP + 1];
  __shared__ int sI;
  const int chain = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const int64_t N = ws.N;
  double W[4], q[5];
  if (!exact_prefixes(ws, v, chain, tile, fallback_pass != 0, W, q, sQ, &sI)) return;
  int* stat = fallback_pass ? &fd.status[chain] : &ws.pend[chain];
  const pgb_utable* ut = utab_base + (int64_t)chain * utab_stride + utab_index;
  const int64_t M = ut->M;
  int64_t base = (int64_t)tile * TILE + (int64_t)tid * IPT;
  double lmax = -INFINITY;

  double Lq[NX * NX];
  const double* Fb = nullptr;
  double* xt = nullptr;
  int32_t* at = nullptr;
  double* lwb = nullptr;
  const double* yt = nullptr;
  const double* zin = nullptr;
  if (MODE == MODE_PROPAGATE) {
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Lq[i] = fd.Lq[chain * 16 + i];
    Fb = fd.Fbuf + (int64_t)chain * NX * fd.N;
    xt = x_row(fd, chain, t);
    at = fd.ah + ((int64_t)chain * fd.T + t) * fd.N;
    lwb = fd.lwbuf + (int64_t)chain * fd.N;
    yt = fd.y + (int64_t)fd.n_y * t;
    zin = fd.Z ? fd.Z + ((int64_t)chain * fd.T + t) * NX * fd.N : nullptr;
  }

  if (base < N) {
    int64_t rprev = pgb_utable_count_le(ut, q[0]);
    if (base == 0 && rprev > 0) atomicOr(stat, 2 /*INDEX_ZERO*/);
    // draws (0, rprev] of the very first thread precede every q_n: the reference returns index 0
    int64_t jfrom = (base == 0) ? 0 : rprev;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t n = base + i;
      if (n >= N) break;
      int64_t r = pgb_utable_count_le(ut, q[i + 1]);
      if (r < rprev) r = rprev;
      if (n == N - 1 && r < M) {  // q_N < u_j: the reference raises BoundsError (:28)
        atomicOr(stat, 1 /*RESAMPLE_OVERRUN*/);
        r = M;
      }
      double F[NX];
      if (MODE == MODE_PROPAGATE && r > jfrom) {
#pragma unroll
        for (int d = 0; d < NX; ++d) F[d] = Fb[(int64_t)d * fd.N + n];
      }
      for (int64_t j = jfrom + 1; j <= r; ++j) {
        // draws j <= rprev of the first thread have index 0 (stored as -1), ancestor clamped to 0
        int32_t a0 = (base == 0 && i == 0 && j <= rprev) ? -1 : (int32_t)n;
        const int64_t jj = j - 1;
        if (MODE == MODE_INDEX) {
          out_idx[(int64_t)chain * out_stride + jj] = a0;
        } else {
          at[jj] = a0;
          double z[NX], x[NX];
          draw_normals<NX>(fd, chain, PGB_STREAM_Z, (uint32_t)t, jj, zin, z);
#pragma unroll
          for (int d = 0; d < NX; ++d) {
            double acc = 0.0;
#pragma unroll
            for (int e = 0; e <= d; ++e) acc = __fma_rn(Lq[d + NX * e], z[e], acc);
            x[d] = __dadd_rn(F[d], acc);  // :175 | :188
            xt[(int64_t)d * fd.N + jj] = x[d];
          }
          double lw = log_weight<NX>(fd, x, yt);
          lwb[jj] = lw;
          lmax = fmax(lmax, lw);
          if (!(lw == lw)) atomicOr(stat, 4);
        }
      }
      rprev = r;
      jfrom = r;
    }
  }
  if (MODE == MODE_PROPAGATE) {
    if (!first && tile == 0 && tid == 0) {
      // conditioned trajectory keeps slot N-1: x_pf[:, N, t] = x_prim[:, t]   (:160)
      double x[NX];
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        x[d] = fd.xprim[(int64_t)chain * NX * fd.T + d + NX * t];
        xt[(int64_t)d * fd.N + (fd.N - 1)] = x[d];
      }
      double lw = log_weight<NX>(fd, x, yt);
      lwb[fd.N - 1] = lw;
      lmax = fmax(lmax, lw);
    }
    block_max_to_global(lmax, &fd.maxkey[chain], sm);
  }
}

// ------------------------------------------------------------------------------------------
// sampling-grid tables for every step of the sweep (independent of the weights)
// ------------------------------------------------------------------------------------------
static __global__ void k_build_utabs(FilterDev fd, int first) {
  const int chain = blockIdx.y;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= fd.T) return;
  const uint32_t cid = fd.chain_base + chain;
  if (t >= 1) {
    double ur, ua;
    if (fd.philox) {
      pgb_u32x4 b = pgb_stream_block(fd.seed, PGB_STREAM_URES, cid, fd.sweep, (uint32_t)t, 0);
      ur = pgb_u53_halfopen(b.v[0], b.v[1]);
      b = pgb_stream_block(fd.seed, PGB_STREAM_UANC, cid, fd.sweep, (uint32_t)t, 0);
      ua = pgb_u53_halfopen(b.v[0], b.v[1]);
    } else {
      ur = fd.Ures[(int64_t)chain * fd.T + t];
      ua = fd.Uanc[(int64_t)chain * fd.T + t];
    }
    pgb_utable_build(&fd.utab_res[(int64_t)chain * fd.T + t], first ? fd.N : fd.N - 1, ur);
    pgb_utable_build(&fd.utab_anc[(int64_t)chain * fd.T + t], 1, ua);
  } else {
    double us;
    if (fd.philox) {
      pgb_u32x4 b = pgb_stream_block(fd.seed, PGB_STREAM_USTAR, cid, fd.sweep, 0, 0);
      us = pgb_u53_halfopen(b.v[0], b.v[1]);
    } else {
      us = fd.Ustar[chain];
    }
    pgb_utable_build(&fd.utab_star[chain], 1, us);
  }
}

// ------------------------------------------------------------------------------------------
// backward trace of the ancestor path (:213-218) and the MNIW regressors (:221-222)
// ------------------------------------------------------------------------------------------
template <int NX>
__global__ void k_trace(FilterDev fd, const int32_t* __restrict__ star_idx, int64_t* __restrict__ star_out) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= fd.nc) return;
  const int64_t N = fd.N, T = fd.T;
  int64_t star = star_idx[chain];
  if (star < 0) star = 0;
  star_out[chain] = star + 1;
  double* xp = fd.xprim + (int64_t)chain * NX * T;
  const double* xh = fd.xh + (int64_t)chain * T * NX * N;
  const int32_t* ah = fd.ah + (int64_t)chain * T * N;
  for (int d = 0; d < NX; ++d) xp[d + NX * (T - 1)] = xh[((T - 1) * NX + d) * N + star];
  for (int64_t t = T - 2; t >= 0; --t) {
    star = ah[(t + 1) * N + star];
    if (star < 0) star = 0;
    for (int d = 0; d < NX; ++d) xp[d + NX * t] = xh[(t * NX + d) * N + star];
  }
}

// Replay mode (x_rows == 2): the particle history is not kept.  The conditioned trajectory of the next sweep
// (:213-218) is the state along ONE ancestor lineage, so it is recomputed: thread 0 walks the lineage backwards through
// the stored ancestor indices, the noise L z(t, b_t) of every step is drawn in parallel (it does not depend on the
// state), then the CTA re-runs the T propagation steps of that single lineage with exactly the arithmetic of the
// filter (same sines, same product order, F_d = fma(A[d,i], phi_i, F_d) in increasing i, x = F + L z), which reproduces
// the stored values bit for bit.  One 128-thread CTA per chain; dynamic shared memory: phi[n_phi].
constexpr int TR_TPB = 128;
template <int NX>
__global__ void __launch_bounds__(TR_TPB) k_trace_replay(FilterDev fd, const int32_t* __restrict__ star_idx,
                                                         int64_t* __restrict__ star_out, double* __restrict__ noise,
                                                         int first) {
  extern __shared__ double phi_sm[];
  __shared__ double tk[PGB_MAX_Z_][MAXC];
  __shared__ double xs[4];
  const int chain = blockIdx.x, tid = threadIdx.x;
  const int64_t N = fd.N, T = fd.T;
  double* xp = fd.xprim + (int64_t)chain * NX * T;
  const int32_t* ah = fd.ah + (int64_t)chain * T * N;
  int32_t* lin = fd.lineage + (int64_t)chain * T;
  double* nz_ = noise + (int64_t)chain * T * NX;
  if (tid == 0) {
    int64_t star = star_idx[chain];
    if (star < 0) star = 0;
    star_out[chain] = star + 1;
    lin[T - 1] = (int32_t)star;
    for (int64_t t = T - 2; t >= 0; --t) {
      star = ah[(t + 1) * N + star];
      if (star < 0) star = 0;
      lin[t] = (int32_t)star;
    }
  }
  __syncthreads();
  // noise of every step along the lineage: t = 0 initial draw (:161), t >= 1 process noise (:175|:188)
  {
    double Lq[NX * NX];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Lq[i] = fd.Lq[chain * 16 + i];
    for (int64_t t = tid; t < T; t += TR_TPB) {
      const int64_t b = lin[t];
      double z[NX];
      if (t == 0) {
        if (b == N - 1) {
#pragma unroll
          for (int d = 0; d < NX; ++d) z[d] = 0.0;  // pinned slot: no draw
        } else {
          draw_normals<NX>(fd, chain, PGB_STREAM_Z0, 0, b, fd.Z0 ? fd.Z0 + (int64_t)chain * NX * (N - 1) : nullptr, z);
        }
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          double acc = 0.0;
#pragma unroll
          for (int e = 0; e <= d; ++e) acc = __fma_rn(fd.x0_chol[d + NX * e], z[e], acc);
          nz_[t * NX + d] = acc;
        }
      } else {
        draw_normals<NX>(fd, chain, PGB_STREAM_Z, (uint32_t)t, b, fd.Z ? fd.Z + ((int64_t)chain * T + t) * NX * N : nullptr, z);
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          double acc = 0.0;
#pragma unroll
          for (int e = 0; e <= d; ++e) acc = __fma_rn(Lq[d + NX * e], z[e], acc);
          nz_[t * NX + d] = acc;
        }
      }
    }
  }
  __syncthreads();
  const double* A = fd.A + (int64_t)chain * NX * fd.n_phi;
  const int nz = fd.n_z, n_phi = fd.n_phi;
  int ns = 0;
  if (fd.basis != 0)
    for (int k = 0; k < nz; ++k) ns += fd.hg_count[k];
  // the sine this thread evaluates (Hilbert basis): s -> (dimension sk, frequency sj)
  int sk = 0, sj = tid;
  if (tid < ns)
    while (sj >= fd.hg_count[sk]) sj -= fd.hg_count[sk++];
  for (int64_t t = 0; t < T; ++t) {
    const int64_t b = lin[t];
    const bool pinned = (b == N - 1) && (t == 0 || !first);  // slot N holds the previous sweep's trajectory
    if (pinned) {
      if (tid < NX) xs[tid] = xp[tid + NX * t];
    } else if (t == 0) {
      if (tid < NX) xs[tid] = __dadd_rn(fd.x0_mean[tid], nz_[tid]);
    } else {
      const double* u = fd.u + (int64_t)fd.n_u * (t - 1);
      if (fd.basis == 0) {
        if (tid == 0) {
          double x[NX], F[NX];
#pragma unroll
          for (int d = 0; d < NX; ++d) x[d] = xs[d];
          eval_F<NX>(fd, A, x, u, F);
#pragma unroll
          for (int d = 0; d < NX; ++d) xs[d] = __dadd_rn(F[d], nz_[t * NX + d]);
        }
      } else {
        if (tid < ns) {
          const double zk = (sk < fd.n_u) ? u[sk] : xs[sk - fd.n_u];
          const double arg[1] = {__dmul_rn(fd.hg_pij2l[sk][sj], __dadd_rn(zk, fd.hg_L[sk]))};
          double sn[1], cs[1];
          vsincos_wide<1>(arg, sn, cs);
          tk[sk][sj] = __dmul_rn(fd.hg_lsi[sk], sn[0]);
        }
        __syncthreads();
        for (int i = tid; i < n_phi; i += TR_TPB) {  // phi_i = ((1*t_0)*t_1)*..., dimension 0 slowest
          int dg[PGB_MAX_Z_];
          int rem = i;
#pragma unroll
          for (int k = PGB_MAX_Z_ - 1; k >= 0; --k)
            if (k < nz) {
              dg[k] = rem % fd.hg_count[k];
              rem /= fd.hg_count[k];
            }
          double p = 1.0;
#pragma unroll
          for (int k = 0; k < PGB_MAX_Z_; ++k)
            if (k < nz) p = __dmul_rn(p, tk[k][dg[k]]);
          phi_sm[i] = p;
        }
        __syncthreads();
        if (tid < NX) {
          double acc = 0.0;
          for (int i = 0; i < n_phi; ++i) acc = __fma_rn(A[tid + NX * i], phi_sm[i], acc);
          xs[tid] = __dadd_rn(acc, nz_[t * NX + tid]);
        }
      }
    }
    __syncthreads();
    if (tid < NX) xp[tid + NX * t] = xs[tid];
  }
}

}  // namespace pgb
