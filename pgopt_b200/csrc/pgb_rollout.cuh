// pgb_rollout.cuh — scenario sampling + forward rollout for solve_PG_OCP
// (Julia/src/optimal_control_Ipopt.jl:47-64 x_vec_0, :67-74 v_vec, :77-83 e_vec, :114-125 X_init/Y_init).
//
// Two kernels, identical bits: k_rollout_warp (below, default: one WARP per scenario) and k_rollout (one 128-thread CTA
// per scenario; kept for n_phi > 640, very long horizons and as a cross-check).  Description of k_rollout:
// One 128-thread CTA per scenario k.  The scenario's model matrix A_k (n_x x n_phi, 20 KB at n_phi = 625) is read
// from HBM exactly once into REGISTERS: thread j keeps the c = ceil(n_phi/128) consecutive columns [j*c, (j+1)*c).
// Every step of the horizon is then
//   sines of the Hilbert basis (one per thread) -> phi_i = prod_k t_k[dig_k(i)] for the thread's columns (consecutive
//   columns share all digits but the last: one prefix product + one multiplication per column) ->
//   per-thread partial sums fma(A[d,i], phi_i, .) -> balanced tree over the 128 threads (shuffles + 4 warp sums)
// which is bit for bit the blocked product the oracle defines for the rollout (oracle/pg_oracle.c eval_f_blocked).
// All process / measurement noise of the horizon is drawn up front (it does not depend on the state), one step per
// thread.  Streams: scenario k uses Philox chain id k_offset + k, purpose ROLL with t = 0 (star uniform),
// 1 (x0 noise), 2 (process noise, j = step), 3 (measurement noise, j = step).
// HBM traffic per scenario: A_k once + w_m1 + one column of x_m1 + outputs — the algorithmic minimum.
#pragma once
#include "pgb_filter.cuh"

namespace pgb {

constexpr int R_TPB = 128;  // threads per scenario (= number of interleaved partial sums of the blocked product)
constexpr int R_WSTAGE = 1024;  // weights of the last filtering step staged in shared memory (the rest stay in L2)

struct RolloutDev {
  int64_t K, H, N;
  int64_t n_models;    // 0: scenario k reads model k; > 0: scenario k reads model k mod n_models (test_prediction)
  const double* A;     // [K][n_x*n_phi]   column-major n_x x n_phi per scenario
  const double* Q;     // [K][n_x*n_x]
  const double* w_m1;  // [K][N]
  const double* x_m1;  // [K][n_x*N]  column-major n_x x N per scenario
  const double* u_m1;  // [K][n_u]
  const double* u_init;// [n_u*H]
  double Le[16];       // lower factor of the measurement noise, n_y x n_y
  uint64_t seed;
  uint32_t k_offset;
  double *x0, *v, *e, *X, *Y;  // [K][n_x], [K][H][n_x], [K][H][n_y], [K][H+1][n_x], [K][H][n_y]
  // supplied scenarios (optimal_control_Ipopt.jl:48,67,77: the draws are skipped when the kwargs are given): same
  // layouts as x0 / v / e, null = draw.  May alias the outputs (the noise of a previous rollout kept on the device).
  const double *x0_in, *v_in, *e_in;
  const double* x_star;  // [K][n_x] or null: x_m1[:, star] drawn beforehand (k_draw_xT); w_m1 / x_m1 are not read then
  int* status;         // [1]
};

struct RolloutShared {
  double tk[PGB_MAX_Z_][MAXC];  // scaled sines of the current state, per input dimension
  double red[R_TPB / 32][4];    // warp partial sums
  double x[4];                  // current state
  double x0n[4];                // noise of the initial state
  double phi_known[5];
  int sine_dim[PGB_MAX_Z_ * MAXC], sine_j[PGB_MAX_Z_ * MAXC];  // which sine thread s evaluates
  int n_sines;
};

// Mixed-radix digits of basis-function index i (last dimension fastest): dimensions 0..nz-2 in 4-bit fields from
// bit 0 (at most 7 of them), the last dimension's digit in bits 28..31.
__device__ __forceinline__ unsigned int rollout_digits(const FilterDev& fd, int i) {
  const int nz = fd.n_z;
  int rem = i;
  unsigned int g = (unsigned int)(rem % fd.hg_count[nz - 1]) << 28;
  rem /= fd.hg_count[nz - 1];
  for (int kk = nz - 2; kk >= 0; --kk) {
    g |= (unsigned int)(rem % fd.hg_count[kk]) << (4 * kk);
    rem /= fd.hg_count[kk];
  }
  return g;
}

// 2*NP standard normals of the ROLL stream (chain cid, sweep 0, time index t, blocks j0 .. j0+NP-1): the same values
// as pgb_normal_pair(pgb_stream_block(...)), through the branch-free vector routines.
template <int NP>
__device__ __forceinline__ void rollout_normals(uint64_t seed, uint32_t cid, uint32_t t, uint32_t j0, double (&z)[4]) {
  pgb_u32x4 ctr[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    ctr[p].v[0] = j0 + p;
    ctr[p].v[1] = t;
    ctr[p].v[2] = 0u;
    ctr[p].v[3] = ((uint32_t)PGB_STREAM_ROLL << 24) | (cid & 0x00ffffffu);
  }
  vphilox_wide<NP>(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  double z0[NP], z1[NP];
  vnormal_pairs_wide<NP>(ctr, z0, z1);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    z[2 * p] = z0[p];
    z[2 * p + 1] = z1[p];
  }
}

// Run index of a thread: lane bits reversed, so that the shuffle distances 16, 8, 4, 2, 1 pair runs j and j^1, then
// j^2, ... — the balanced binary tree over j = 0..127 with adjacent runs added first, as the oracle defines it.
__device__ __forceinline__ int rollout_run(int tid) { return (tid & ~31) | (int)(__brev((unsigned int)(tid & 31)) >> 27); }

// Role index of a thread: the warps of a CTA are rotated by the block index, so that the warp doing the serial work
// (sines, state update, start-up) does not land on the same SM sub-partition (warp id mod 4) for every resident CTA.
__device__ __forceinline__ int rollout_role(int tid) { return (tid + 32 * (int)(blockIdx.x & 3u)) & (R_TPB - 1); }

__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

// Tree sum over the warp's 32 runs of NX values per thread with 6 (instead of 20) shuffles for NX = 4: the first
// levels exchange only the half of the values the partner keeps.  On return v[0] of lane 8*d (NX > 2) or 16*d
// (NX = 2) holds the warp's sum of component d.
template <int NX>
__device__ __forceinline__ double warp_tree_multi(const double (&v)[NX], int lane) {
  if (NX == 1) {
    double b = v[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = __dadd_rn(b, shfl_xor_d(b, o));
    return b;
  } else if (NX == 2) {
    const bool hi = lane & 16;
    double b = __dadd_rn(hi ? v[1] : v[0], shfl_xor_d(hi ? v[0] : v[1], 16));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) b = __dadd_rn(b, shfl_xor_d(b, o));
    return b;
  } else {
    const double v3 = (NX > 3) ? v[NX > 3 ? 3 : 0] : 0.0;
    const bool hi = lane & 16;
    const double a0 = __dadd_rn(hi ? v[2] : v[0], shfl_xor_d(hi ? v[0] : v[2], 16));
    const double a1 = __dadd_rn(hi ? v3 : v[1], shfl_xor_d(hi ? v[1] : v3, 16));
    const bool hi2 = lane & 8;
    double b = __dadd_rn(hi2 ? a1 : a0, shfl_xor_d(hi2 ? a0 : a1, 8));
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) b = __dadd_rn(b, shfl_xor_d(b, o));
    return b;
  }
}

// F = A_k * phi(x, u) in the blocked order; component d of the result is returned to thread d (< NX) only.
// Areg: the thread's columns of A_k.  Contains two __syncthreads(); the caller must place one more before the
// next call (sh.red / sh.tk / sh.x are reused).
template <int NX, int NI>
__device__ __forceinline__ void rollout_eval_F(const FilterDev& fd, RolloutShared& sh, const double (&Areg)[NI][NX],
                                               const double* __restrict__ Ag, int c, const unsigned int (&dig)[NI],
                                               const double* __restrict__ u, double& Fd) {
  const int tid = threadIdx.x, run = rollout_run(tid), vt = rollout_role(tid);
  double part[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) part[d] = 0.0;
  if (fd.basis == 0) {
    // examples/PG_OCP_known_basis_functions_Altro.jl:34 (five functions; thread i < 5 owns column i)
    if (run < 5) {
      const double x0 = sh.x[0], x1 = sh.x[NX > 1 ? 1 : 0];
      double p;
      if (run == 0) p = __dmul_rn(0.1, x0);
      else if (run == 1) p = __dmul_rn(0.1, x1);
      else if (run == 2) p = u[0];
      else if (run == 3) p = __dmul_rn(__dmul_rn(0.01, pgb_cos(__dmul_rn(3.0, x0))), x1);
      else p = __dmul_rn(__dmul_rn(0.1, pgb_sin(__dmul_rn(2.0, x1))), u[0]);
#pragma unroll
      for (int d = 0; d < NX; ++d) part[d] = __fma_rn(Areg[0][d], p, 0.0);
    }
  } else {
    // Hilbert-space basis (examples/PG_OCP_generic_basis_functions_Altro.jl:85-99): one scaled sine per thread
    if (vt < sh.n_sines) {
      const int k = sh.sine_dim[vt], j = sh.sine_j[vt];
      const double zk = (k < fd.n_u) ? u[k] : sh.x[k - fd.n_u];
      const double zl = __dadd_rn(zk, fd.hg_L[k]);
      // branch-free sine (the scalar routine's quadrant switch would serialise the warp's lanes); same bits
      const double arg[1] = {__dmul_rn(fd.hg_pij2l[k][j], zl)};
      double sn[1], cs[1];
      vsincos_wide<1>(arg, sn, cs);
      sh.tk[k][j] = __dmul_rn(fd.hg_lsi[k], sn[0]);
    }
    __syncthreads();
    const int nz = fd.n_z, n_phi = fd.n_phi;
    const double* tl = sh.tk[nz - 1];
    // product over the dimensions 0..nz-2 (digits 4 bits each, from bit 0), in the oracle's left-to-right order
    auto prefix = [&](unsigned int g) {
      double p = 1.0;
#pragma unroll
      for (int k = 0; k < PGB_MAX_Z_ - 1; ++k)
        if (k < nz - 1) p = __dmul_rn(p, sh.tk[k][(g >> (4 * k)) & 15u]);
      return p;
    };
    const int i0 = run * c;
    double pre = prefix(dig[0]);
#pragma unroll
    for (int m = 0; m < NI; ++m) {
      // consecutive columns share every digit but the last one until the last dimension wraps; the vote keeps the
      // recomputation behind a real (warp-uniform) branch instead of predicated-off instructions
      const bool live = (m < c && i0 + m < n_phi);
      if (m > 0) {
        const bool wrap = live && (dig[m] & 0x0fffffffu) != (dig[m - 1] & 0x0fffffffu);
        if (__any_sync(0xffffffffu, wrap)) {
          const double pn = prefix(dig[m]);
          if (wrap) pre = pn;
        }
      }
      if (live) {
        const double p = __dmul_rn(pre, tl[dig[m] >> 28]);
#pragma unroll
        for (int d = 0; d < NX; ++d) part[d] = __fma_rn(Areg[m][d], p, part[d]);
      }
    }
    // columns beyond the register-resident ones (c > NI): A from global memory (L2), digits on the fly
    for (int m = NI; m < c && i0 + m < n_phi; ++m) {
      const unsigned int g = rollout_digits(fd, i0 + m);
      const double p = __dmul_rn(prefix(g), tl[g >> 28]);
#pragma unroll
      for (int d = 0; d < NX; ++d) part[d] = __fma_rn(Ag[(int64_t)NX * (i0 + m) + d], p, part[d]);
    }
  }
  // balanced tree over the 128 partial sums: inside the warp by shuffles, then (w0 + w1) + (w2 + w3)
  const int w = tid >> 5, l = tid & 31;
  const double tot = warp_tree_multi<NX>(part, l);
  constexpr int LSTEP = (NX == 1) ? 32 : (NX == 2 ? 16 : 8);  // lane LSTEP*d holds component d
  if ((l % LSTEP) == 0 && l / LSTEP < NX) sh.red[w][l / LSTEP] = tot;
  __syncthreads();
  if (vt < NX) Fd = __dadd_rn(__dadd_rn(sh.red[0][vt], sh.red[1][vt]), __dadd_rn(sh.red[2][vt], sh.red[3][vt]));
}

#ifndef PGB_ROLLOUT_MINBLOCKS
#define PGB_ROLLOUT_MINBLOCKS 4
#endif
template <int NX, int NI>
__global__ void __launch_bounds__(R_TPB, PGB_ROLLOUT_MINBLOCKS) k_rollout(const __grid_constant__ FilterDev fd,
                                                   const __grid_constant__ RolloutDev rd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RolloutShared& sh = *reinterpret_cast<RolloutShared*>(smem_raw);
  double* vs = reinterpret_cast<double*>(smem_raw + sizeof(RolloutShared));  // [H][NX] process noise
  const int tid = threadIdx.x, vt = rollout_role(tid);
  const int64_t k = blockIdx.x;
  const int64_t K = rd.K, H = rd.H, N = rd.N;
  (void)K;
  const int ny = fd.n_y, nu = fd.n_u, n_phi = fd.n_phi, nz = fd.n_z;
  double* es = vs + H * NX;  // [H][ny] measurement noise
  double* us = es + H * ny;  // [H][nu] planned inputs u_init (read once, not once per step from L2)
  double* ws = us + H * nu;  // [min(N, R_WSTAGE)] staged weights w_m1 (coalesced load; thread 0 walks them serially)
  const uint32_t cid = rd.k_offset + (uint32_t)k;
  const int64_t km = rd.n_models > 0 ? k % rd.n_models : k;  // model (PG sample) this scenario simulates
  const double* Ag = rd.A + km * (int64_t)NX * n_phi;

  // ---- A_k -> registers: thread j owns the c consecutive columns [j*c, (j+1)*c)
  const int c = (n_phi + R_TPB - 1) / R_TPB;
  double Areg[NI][NX];
#pragma unroll
  for (int m = 0; m < NI; ++m) {
    const int i = rollout_run(tid) * c + m;
#pragma unroll
    for (int d = 0; d < NX; ++d) Areg[m][d] = (m < c && i < n_phi) ? Ag[(int64_t)NX * i + d] : 0.0;
  }
  unsigned int dig[NI];
#pragma unroll
  for (int m = 0; m < NI; ++m) dig[m] = (fd.basis != 0 && m < c && rollout_run(tid) * c + m < n_phi) ? rollout_digits(fd, rollout_run(tid) * c + m) : 0u;
  if (vt == 0) {
    int s = 0;
    if (fd.basis != 0)
      for (int kk = 0; kk < nz; ++kk)
        for (int j = 0; j < fd.hg_count[kk]; ++j) {
          sh.sine_dim[s] = kk;
          sh.sine_j[s] = j;
          ++s;
        }
    sh.n_sines = s;
  }
  for (int64_t i = tid; i < H * nu; i += R_TPB) us[i] = rd.u_init[i];
  const int64_t nstage = N < R_WSTAGE ? N : R_WSTAGE;
  if (!rd.x0_in && !rd.x_star)  // the weights are only read by the star draw
    for (int64_t i = tid; i < nstage; i += R_TPB) ws[i] = rd.w_m1[km * N + i];
  // ---- chol(Q_k) (every thread; n_x <= 4)
  double Lq[NX * NX];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) Lq[i] = rd.Q[km * NX * NX + i];
  const int notpd = small_chol(Lq, NX) && !(rd.x0_in && rd.v_in);  // Q is only used for the draws (:55, :71)
  if (notpd) {
    if (vt == 0) atomicOr(rd.status, 16);
    return;
  }
  // x0 noise (:62): drawn by thread 0, added by threads d < NX after the first evaluation of f
  if (vt == 0 && !rd.x0_in) {
    double z[4];
    rollout_normals<(NX + 1) / 2>(rd.seed, cid, 1, 0u, z);
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      double acc = 0.0;
#pragma unroll
      for (int e2 = 0; e2 <= d; ++e2) acc = __fma_rn(Lq[d + NX * e2], z[e2], acc);
      sh.x0n[d] = acc;
    }
  }
  // ---- noise of the whole horizon, one step per thread:  v_vec[:, t, k] (:72), e_vec[:, t, k] (:81)
  for (int64_t t = vt; t < H; t += R_TPB) {
    double z[4];
    if (rd.v_in) {  // supplied v_vec (:67)
      for (int d = 0; d < NX; ++d) vs[t * NX + d] = rd.v_in[(k * H + t) * NX + d];
    } else {
      rollout_normals<(NX + 1) / 2>(rd.seed, cid, 2, (uint32_t)(t * ((NX + 1) / 2)), z);
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
        for (int e2 = 0; e2 <= d; ++e2) acc = __fma_rn(Lq[d + NX * e2], z[e2], acc);
        vs[t * NX + d] = acc;
        if (rd.v) rd.v[(k * H + t) * NX + d] = acc;
      }
    }
    if (rd.e_in) {  // supplied e_vec (:77)
      for (int o = 0; o < ny; ++o) es[t * ny + o] = rd.e_in[(k * H + t) * ny + o];
    } else {
      const int npairs = (ny + 1) / 2;
      if (npairs == 1) rollout_normals<1>(rd.seed, cid, 3, (uint32_t)t, z);
      else rollout_normals<2>(rd.seed, cid, 3, (uint32_t)(t * 2), z);
      for (int o = 0; o < ny; ++o) {
        double acc = 0.0;
        for (int e2 = 0; e2 <= o; ++e2) acc = __fma_rn(rd.Le[o + ny * e2], z[e2], acc);
        es[t * ny + o] = acc;
        if (rd.e) rd.e[(k * H + t) * ny + o] = acc;
      }
    }
  }
  // ---- star = systematic_resampling(w_m1, 1) (:58): W = w/sum(w) (tree), first n with q_n >= u   (thread 0)
  __syncthreads();
  if (vt == 0 && !rd.x0_in && rd.x_star) {  // x_m1[:, star] drawn beforehand where the sample lives (k_draw_xT)
    for (int d = 0; d < NX; ++d) sh.x[d] = rd.x_star[km * NX + d];
  } else if (vt == 0 && !rd.x0_in) {
    const double* wg = rd.w_m1 + km * N;
#define R_W(i) ((i) < nstage ? ws[(i)] : wg[(i)])
    double st[40];
    int sp = 0;
    int64_t P = 1;
    while (P < N) P <<= 1;
    for (int64_t i = 0; i < P; ++i) {
      double xv = (i < N) ? R_W(i) : 0.0;
      int merges = __ffsll((long long)(i + 1)) - 1;
      for (int m = 0; m < merges; ++m) xv = __dadd_rn(st[--sp], xv);
      st[sp++] = xv;
    }
    const double S = st[0];
    pgb_u32x4 b = pgb_stream_block(rd.seed, PGB_STREAM_ROLL, cid, 0, 0, 0);
    const double us = pgb_u53_halfopen(b.v[0], b.v[1]);
    double q = 0.0;
    int64_t n = 0;
    while (q < us) {
      n = n + 1;
      if (n > N) { atomicOr(rd.status, 1); n = N; break; }
      q = __dadd_rn(q, __ddiv_rn(R_W(n - 1), S));
    }
#undef R_W
    if (n == 0) { atomicOr(rd.status, 2); n = 1; }
    const int64_t star = n - 1;
    for (int d = 0; d < NX; ++d) sh.x[d] = rd.x_m1[(km * N + star) * NX + d];
  }
  __syncthreads();
  // ---- x0 = f(x_m1, u_m1) + rand(MvNormal(0, Q))  (:62)
  double Fd = 0.0;
  if (!rd.x0_in) rollout_eval_F<NX, NI>(fd, sh, Areg, Ag, c, dig, rd.u_m1 + km * nu, Fd);  // block-uniform
  if (vt < NX) {
    const double xv = rd.x0_in ? rd.x0_in[k * NX + vt] : __dadd_rn(Fd, sh.x0n[vt]);  // supplied x_vec_0 (:48)
    sh.x[vt] = xv;
    if (rd.x0 && !rd.x0_in) rd.x0[k * NX + vt] = xv;
    if (rd.X) rd.X[(k * (H + 1)) * NX + vt] = xv;
  }
  __syncthreads();
  // ---- horizon: Y[t] = g(X[t]) + e (:122), X[t+1] = f(X[t], u_t) + v (:121)
  for (int64_t t = 0; t < H; ++t) {
    if (vt == 32 && rd.Y) {  // a lane of the next warp: the first role warp is busy with the sines
      for (int o = 0; o < ny; ++o) {
        double g = 0.0;
#pragma unroll
        for (int d = 0; d < NX; ++d) g = __fma_rn(fd.C[o + ny * d], sh.x[d], g);
        rd.Y[(k * H + t) * ny + o] = __dadd_rn(g, es[t * ny + o]);
      }
    }
    rollout_eval_F<NX, NI>(fd, sh, Areg, Ag, c, dig, us + (int64_t)nu * t, Fd);
    if (vt < NX) {
      const double xv = __dadd_rn(Fd, vs[t * NX + vt]);
      sh.x[vt] = xv;
      if (rd.X) rd.X[(k * (H + 1) + t + 1) * NX + vt] = xv;
    }
    __syncthreads();
  }
}


// ==========================================================================================================
// k_rollout_warp — one WARP per scenario (default whenever A_k fits the registers of 32 lanes).
//
// The CTA-per-scenario kernel above spends most of its issue slots on idle lanes (20 sines on 128 threads), CTA
// barriers (3 per step) and per-scenario start-up that nothing overlaps.  Here lane l owns the oracle's runs
// 4l .. 4l+3 (4*c consecutive columns of A_k, register-resident for the whole horizon), so that
//   * the first two levels of the oracle's balanced tree over the 128 runs are in-thread additions,
//   * the remaining five levels are xor-butterflies at lane distance 1, 2, 4, 8, 16 (adjacent runs first), with the
//     n_x components split over the lanes at the first levels (6 shuffles for n_x = 4),
//   * the trajectory X[:, 0..H] lives in the warp's shared-memory slice (the lanes holding component d after the tree
//     add the noise and store it; the sine lanes read their input from there) and nothing but __syncwarp() separates
//     the steps: the warps of a CTA run different scenarios completely independently, one hides the other's
//     latency-bound sines and start-up; x_vec_0 / X / Y go to global memory once per scenario, after the horizon,
//   * RS of a lane's four runs of A_k sit in lane-private shared-memory slots instead of registers (168 instead of
//     255 registers: 12 instead of 8 warps per SM),
//   * when a run is exactly one row of the last tensor dimension (c == n_phi_dims[end], e.g. 5^4 = 625 functions on
//     128 runs) every column of a run shares its prefix product and column m has last digit m.
// Same arithmetic, same order, same bits as the kernel above and as oracle/pg_oracle.c eval_f_blocked.
// ==========================================================================================================
constexpr int RW_WPB = 4;  // warps (independent scenarios) per CTA
constexpr int RW_RPL = 4;  // runs per lane: 32 lanes x 4 = the 128 runs of the blocked product
constexpr int RW_MAXNI = 5;

struct __align__(16) RolloutWShared {
  unsigned int dig[RW_RPL * RW_MAXNI * 32];  // packed digits of column m of run rr of lane l at [(rr*NI + m)*32 + l]
  int sine_dim[PGB_MAX_Z_ * MAXC], sine_j[PGB_MAX_Z_ * MAXC];
  int n_sines;
};

// Tree over the warp's lanes in the order lane^1, lane^2, lane^4, ...; component d of the result is valid in lane
// rw_src<NX>(d) (and in every lane that agrees with it in the low bits).
template <int NX>
__device__ __forceinline__ double rw_tree(const double (&v)[NX], int lane) {
  if (NX == 1) {
    double b = v[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) b = __dadd_rn(b, shfl_xor_d(b, o));
    return b;
  } else if (NX == 2) {
    const bool hi = lane & 1;
    double b = __dadd_rn(hi ? v[1] : v[0], shfl_xor_d(hi ? v[0] : v[1], 1));
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) b = __dadd_rn(b, shfl_xor_d(b, o));
    return b;
  } else {
    const double v3 = (NX > 3) ? v[NX > 3 ? 3 : 0] : 0.0;
    const bool hi = lane & 1;
    const double a0 = __dadd_rn(hi ? v[2] : v[0], shfl_xor_d(hi ? v[0] : v[2], 1));
    const double a1 = __dadd_rn(hi ? v3 : v[1], shfl_xor_d(hi ? v[1] : v3, 1));
    const bool hi2 = lane & 2;
    double b = __dadd_rn(hi2 ? a1 : a0, shfl_xor_d(hi2 ? a0 : a1, 2));
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) b = __dadd_rn(b, shfl_xor_d(b, o));
    return b;
  }
}
template <int NX>
__device__ __forceinline__ int rw_src(int d) {
  return NX == 1 ? 0 : (NX == 2 ? d : (((d & 1) << 1) | (d >> 1)));
}
// component of the tree result that lane holds (inverse of rw_src on the low lane bits; may be >= NX for NX = 3)
template <int NX>
__device__ __forceinline__ int rw_comp(int lane) {
  return NX == 1 ? 0 : (NX == 2 ? (lane & 1) : (((lane & 1) << 1) | ((lane >> 1) & 1)));
}

// constants of the one sine a lane evaluates per step (n_sines <= 32): dimension, slot in tk, L_k, pi j / (2 L_k), 1/sqrt(L_k)
struct RwSine {
  int k, off;
  double L, pij2l, lsi;
};

// Column m of run rr, component d of the lane's part of A_k: runs 0 .. RW_RPL-RS-1 sit in registers, the last RS
// runs in the warp's shared-memory slice (lane-private slots [(rs*NI + m)*NX + d][lane]: conflict-free, no
// synchronisation needed — a lane only reads what it wrote).
template <int NX, int NI, int RS>
__device__ __forceinline__ double rw_A(const double (&Areg)[RW_RPL - RS > 0 ? RW_RPL - RS : 1][NI][NX],
                                       const double* __restrict__ As, int rr, int m, int d, int lane) {
  constexpr int NR = RW_RPL - RS;
  if (rr < NR) return Areg[rr < NR ? rr : 0][m][d];
  return As[(((rr - NR) * NI + m) * NX + d) * 32 + lane];
}

// F = A_k * phi(x, u) in the blocked order; every lane returns all NX components.
template <int NX, int NI, int RS>
__device__ __forceinline__ void rw_eval_F(const FilterDev& fd, const RolloutWShared& sh, double* __restrict__ tkw,
                                          const double (&Areg)[RW_RPL - RS > 0 ? RW_RPL - RS : 1][NI][NX],
                                          const double* __restrict__ As, const int (&ncols)[RW_RPL], int c,
                                          int kl, const RwSine& ls, const double* __restrict__ x,
                                          const double* __restrict__ u, double& Fb) {
  const int lane = threadIdx.x & 31;
  double P[NX];
  if (fd.basis == 0) {
    // examples/PG_OCP_known_basis_functions_Altro.jl:34 — five functions, run j = column j (c = 1): lane 0 owns
    // columns 0..3, lane 1 column 4.  Every lane evaluates the five values (warp-uniform, no divergence).
    const double x0 = x[0], x1 = x[NX > 1 ? 1 : 0], u0 = u[0];
    const double a3[1] = {__dmul_rn(3.0, x0)}, a2[1] = {__dmul_rn(2.0, x1)};
    double s3[1], c3[1], s2[1], c2[1];
    vsincos_wide<1>(a3, s3, c3);
    vsincos_wide<1>(a2, s2, c2);
    double ph[5];
    ph[0] = __dmul_rn(0.1, x0);
    ph[1] = __dmul_rn(0.1, x1);
    ph[2] = u0;
    ph[3] = __dmul_rn(__dmul_rn(0.01, c3[0]), x1);
    ph[4] = __dmul_rn(__dmul_rn(0.1, s2[0]), u0);
    double part[RW_RPL][NX];
#pragma unroll
    for (int rr = 0; rr < RW_RPL; ++rr) {
      const double p = (lane == 0) ? ph[rr] : ph[4];  // lane 1: only run 0 (= column 4) is live
#pragma unroll
      for (int d = 0; d < NX; ++d) part[rr][d] = (ncols[rr] > 0) ? __fma_rn(rw_A<NX, NI, RS>(Areg, As, rr, 0, d, lane), p, 0.0) : 0.0;
    }
#pragma unroll
    for (int d = 0; d < NX; ++d)
      P[d] = __dadd_rn(__dadd_rn(part[0][d], part[1][d]), __dadd_rn(part[2][d], part[3][d]));
  } else {
    // Hilbert-space basis (examples/PG_OCP_generic_basis_functions_Altro.jl:85-99): one scaled sine per lane
    if (sh.n_sines <= 32) {  // the lane's sine is the same for every step and scenario: its constants sit in registers
      if (lane < sh.n_sines) {
        const double zk = (ls.k < fd.n_u) ? u[ls.k] : x[ls.k - fd.n_u];
        const double arg[1] = {__dmul_rn(ls.pij2l, __dadd_rn(zk, ls.L))};
        double sn[1], cs[1];
        vsincos_wide<1>(arg, sn, cs);
        tkw[ls.off] = __dmul_rn(ls.lsi, sn[0]);
      }
    } else {
      for (int s = lane; s < sh.n_sines; s += 32) {
        const int k = sh.sine_dim[s], j = sh.sine_j[s];
        const double zk = (k < fd.n_u) ? u[k] : x[k - fd.n_u];
        const double zl = __dadd_rn(zk, fd.hg_L[k]);
        const double arg[1] = {__dmul_rn(fd.hg_pij2l[k][j], zl)};
        double sn[1], cs[1];
        vsincos_wide<1>(arg, sn, cs);
        tkw[k * MAXC + j] = __dmul_rn(fd.hg_lsi[k], sn[0]);
      }
    }
    __syncwarp();
    const int nz = fd.n_z;
    const double* tl = tkw + (nz - 1) * MAXC;
    // product over the dimensions 0..nz-2 (digits 4 bits each, from bit 0), in the oracle's left-to-right order
    auto prefix = [&](unsigned int g) {
      double p = 1.0;
#pragma unroll
      for (int k = 0; k < PGB_MAX_Z_ - 1; ++k)
        if (k < nz - 1) p = __dmul_rn(p, tkw[k * MAXC + ((g >> (4 * k)) & 15u)]);
      return p;
    };
    double part[RW_RPL][NX];
#pragma unroll
    for (int rr = 0; rr < RW_RPL; ++rr)
#pragma unroll
      for (int d = 0; d < NX; ++d) part[rr][d] = 0.0;
    if (kl >= 0) {
      // aligned runs: a run is one full row of dimension kl (the last one with more than one function; at most one
      // single-function dimension follows it) -> one prefix per run, column m has digit m in dimension kl.
      // Columns beyond n_phi carry A = 0 and digit 0: fma(0, p, acc) leaves acc unchanged (p is a finite product of
      // sines unless the state itself is already non-finite), so there is no per-column predicate.
      unsigned int g[RW_RPL];
      double pre[RW_RPL];
#pragma unroll
      for (int rr = 0; rr < RW_RPL; ++rr) {
        g[rr] = sh.dig[(rr * NI) * 32 + lane];
        pre[rr] = 1.0;
      }
      const double* tkk = tkw;
      for (int k = 0; k < kl; ++k) {  // the four runs' prefix products, interleaved (left to right over the dimensions)
#pragma unroll
        for (int rr = 0; rr < RW_RPL; ++rr) {
          pre[rr] = __dmul_rn(pre[rr], tkk[g[rr] & 15u]);
          g[rr] >>= 4;
        }
        tkk += MAXC;
      }
      double tlm[NI];
#pragma unroll
      for (int m = 0; m < NI; ++m) tlm[m] = (m < c) ? tkk[m] : 0.0;
      if (kl < nz - 1) {
        const double tr = tl[0];
#pragma unroll
        for (int rr = 0; rr < RW_RPL; ++rr)
#pragma unroll
          for (int m = 0; m < NI; ++m) {
            const double p = __dmul_rn(__dmul_rn(pre[rr], tlm[m]), tr);
#pragma unroll
            for (int d = 0; d < NX; ++d) part[rr][d] = __fma_rn(rw_A<NX, NI, RS>(Areg, As, rr, m, d, lane), p, part[rr][d]);
          }
      } else {
#pragma unroll
        for (int rr = 0; rr < RW_RPL; ++rr)
#pragma unroll
          for (int m = 0; m < NI; ++m) {
            const double p = __dmul_rn(pre[rr], tlm[m]);
#pragma unroll
            for (int d = 0; d < NX; ++d) part[rr][d] = __fma_rn(rw_A<NX, NI, RS>(Areg, As, rr, m, d, lane), p, part[rr][d]);
          }
      }
    } else {
#pragma unroll
      for (int rr = 0; rr < RW_RPL; ++rr) {
        unsigned int gprev = 0u;
        double pre = 1.0;
#pragma unroll
        for (int m = 0; m < NI; ++m) {
          if (m < c) {  // warp-uniform
            const unsigned int g = sh.dig[(rr * NI + m) * 32 + lane];
            const bool live = m < ncols[rr];
            const bool need = live && (m == 0 || (g & 0x0fffffffu) != (gprev & 0x0fffffffu));
            if (__any_sync(0xffffffffu, need)) {
              const double pn = prefix(g);
              if (need) pre = pn;
            }
            if (live) {
              const double p = __dmul_rn(pre, tl[g >> 28]);
#pragma unroll
              for (int d = 0; d < NX; ++d) part[rr][d] = __fma_rn(rw_A<NX, NI, RS>(Areg, As, rr, m, d, lane), p, part[rr][d]);
            }
            gprev = g;
          }
        }
      }
    }
#pragma unroll
    for (int d = 0; d < NX; ++d)
      P[d] = __dadd_rn(__dadd_rn(part[0][d], part[1][d]), __dadd_rn(part[2][d], part[3][d]));
    __syncwarp();  // tkw is rewritten by the next evaluation
  }
  Fb = rw_tree<NX>(P, lane);
}

#ifndef PGB_ROLLOUT_W_MINBLOCKS
#define PGB_ROLLOUT_W_MINBLOCKS 2
#endif
// star = systematic_resampling(w_m1, 1) (optimal_control_Ipopt.jl:58) by one warp: S = balanced-tree sum of w, then the
// first n with q_n >= u, q strictly sequential (particle_Gibbs.jl:26-29); every lane walks the same path.  Returns the
// 0-based index; raises the reference's edge cases in *status (overrun: BoundsError, index 0).
__device__ __forceinline__ int64_t rw_draw_star(const double* __restrict__ wg, int64_t N, uint64_t seed, uint32_t cid,
                                                int* status) {
  const int lane = threadIdx.x & 31;
  int64_t P = 1;
  while (P < N) P <<= 1;
  double S;
  if (P >= 32) {
    const int64_t per = P / 32, base = (int64_t)lane * per;
    double st[40];
    int sp = 0;
    for (int64_t i = 0; i < per; ++i) {
      double xv = (base + i < N) ? wg[base + i] : 0.0;
      const int merges = __ffsll((long long)(i + 1)) - 1;
      for (int m = 0; m < merges; ++m) xv = __dadd_rn(st[--sp], xv);
      st[sp++] = xv;
    }
    S = st[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) S = __dadd_rn(S, shfl_xor_d(S, o));
  } else {
    S = (lane < N) ? wg[lane] : 0.0;
    for (int o = 1; o < (int)P; o <<= 1) S = __dadd_rn(S, shfl_xor_d(S, o));
    S = __shfl_sync(0xffffffffu, S, 0);
  }
  const pgb_u32x4 b = pgb_stream_block(seed, PGB_STREAM_ROLL, cid, 0, 0, 0);
  const double ust = pgb_u53_halfopen(b.v[0], b.v[1]);
  double q = 0.0;
  int64_t n = 0;
  bool over = false;
  for (int64_t base = 0;; base += 32) {  // every lane walks the same (uniform) path
    if (!(q < ust)) break;
    if (base >= N) { over = true; break; }
    const double Wl = (base + lane < N) ? __ddiv_rn(wg[base + lane], S) : 0.0;
    const int lim = (N - base < 32) ? (int)(N - base) : 32;
    for (int l = 0; l < lim; ++l) {
      if (!(q < ust)) break;
      ++n;
      q = __dadd_rn(q, __shfl_sync(0xffffffffu, Wl, l));
    }
  }
  if (over) {
    if (lane == 0) atomicOr(status, 1);
    n = N;
  }
  if (n == 0) {
    if (lane == 0) atomicOr(status, 2);
    n = 1;
  }
  return n - 1;
}

// x_T[k] = x_m1[:, star_k] for K samples, one warp each — the draw of optimal_control_Ipopt.jl:58-59 done where the
// sample lives, so that only (A, Q, x_T) has to travel between GPUs.  Scenario ids as in the rollout.
static __global__ void __launch_bounds__(128) k_draw_xT(int64_t K, int64_t N, int nx, const double* __restrict__ w_m1,
                                                        const double* __restrict__ x_m1, uint64_t seed, uint32_t k_offset,
                                                        double* __restrict__ xT, int* status) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (k >= K) return;
  const int64_t star = rw_draw_star(w_m1 + k * N, N, seed, k_offset + (uint32_t)k, status);
  if (lane < nx) xT[k * nx + lane] = x_m1[(k * N + star) * nx + lane];
}

template <int NX, int NI, int RS>
__global__ void __launch_bounds__(RW_WPB * 32, RS >= 2 ? 3 : PGB_ROLLOUT_W_MINBLOCKS)
k_rollout_warp(const __grid_constant__ FilterDev fd, const __grid_constant__ RolloutDev rd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RolloutWShared& sh = *reinterpret_cast<RolloutWShared*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t K = rd.K, H = rd.H, N = rd.N;
  const int ny = fd.n_y, nu = fd.n_u, n_phi = fd.n_phi, nz = fd.n_z;
  double* us = reinterpret_cast<double*>(smem_raw + sizeof(RolloutWShared));  // [H][nu] planned inputs (CTA-wide)
  const int64_t per_warp = PGB_MAX_Z_ * MAXC + H * (NX + ny) + RS * NI * NX * 32 + (H + 2) * NX;
  double* tkw = us + H * nu + (int64_t)warp * per_warp;  // [n_z][MAXC] scaled sines of the warp's scenario
  double* vs = tkw + PGB_MAX_Z_ * MAXC;                  // [H][NX] process noise
  double* es = vs + H * NX;                              // [H][ny] measurement noise
  double* As = es + H * ny;                              // [RS][NI][NX][32] the lanes' last RS runs of A_k
  double* xs = As + RS * NI * NX * 32;                   // [H+1][NX] trajectory X[:, t] (outputs are written at the end)
  double* xin = xs + (H + 1) * NX;                       // [NX] x_m1[:, star]

  // ---- model tables shared by every scenario (CTA-wide, once)
  const int c = (n_phi + 128 - 1) / 128;
  for (int i = tid; i < RW_RPL * NI * 32; i += RW_WPB * 32) {
    const int l = i & 31, rm = i >> 5, rr = rm / NI, m = rm % NI;
    const int col = (RW_RPL * l + rr) * c + m;
    sh.dig[i] = (fd.basis != 0 && m < c && col < n_phi) ? rollout_digits(fd, col) : 0u;
  }
  if (tid == 0) {
    int s = 0;
    if (fd.basis != 0)
      for (int kk = 0; kk < nz; ++kk)
        for (int j = 0; j < fd.hg_count[kk]; ++j) {
          sh.sine_dim[s] = kk;
          sh.sine_j[s] = j;
          ++s;
        }
    sh.n_sines = s;
  }
  for (int64_t i = tid; i < H * nu; i += RW_WPB * 32) us[i] = rd.u_init[i];
  __syncthreads();
  int ncols[RW_RPL];
#pragma unroll
  for (int rr = 0; rr < RW_RPL; ++rr) {
    const int first = (RW_RPL * lane + rr) * c;
    int n = n_phi - first;
    ncols[rr] = n < 0 ? 0 : (n > c ? c : n);
  }
  // aligned runs (see rw_eval_F): c columns per run = one full row of the last dimension that has more than one function
  int kl = -1;
  if (fd.basis != 0) {
    int k2 = nz - 1;
    while (k2 > 0 && fd.hg_count[k2] == 1) --k2;
    if (c == fd.hg_count[k2] && nz - 1 - k2 <= 1) kl = k2;
  }

  RwSine ls;
  ls.k = 0; ls.off = 0; ls.L = 0.0; ls.pij2l = 0.0; ls.lsi = 0.0;
  if (fd.basis != 0 && lane < sh.n_sines) {
    ls.k = sh.sine_dim[lane];
    ls.off = ls.k * MAXC + sh.sine_j[lane];
    ls.L = fd.hg_L[ls.k];
    ls.pij2l = fd.hg_pij2l[ls.k][sh.sine_j[lane]];
    ls.lsi = fd.hg_lsi[ls.k];
  }

  for (int64_t k = (int64_t)blockIdx.x * RW_WPB + warp; k < K; k += (int64_t)gridDim.x * RW_WPB) {
    const uint32_t cid = rd.k_offset + (uint32_t)k;
    const int64_t km = rd.n_models > 0 ? k % rd.n_models : k;  // model (PG sample) this scenario simulates
    const double* Ag = rd.A + km * (int64_t)NX * n_phi;
    // ---- A_k -> registers (read from HBM exactly once)
    constexpr int NR = RW_RPL - RS;
    double Areg[NR > 0 ? NR : 1][NI][NX];
#pragma unroll
    for (int rr = 0; rr < RW_RPL; ++rr)
#pragma unroll
      for (int m = 0; m < NI; ++m) {
        const int64_t col = (int64_t)(RW_RPL * lane + rr) * c + m;
        double av[NX];
        if (NX % 2 == 0) {
#pragma unroll
          for (int d = 0; d < NX; d += 2) {
            double2 a = make_double2(0.0, 0.0);
            if (m < ncols[rr]) a = __ldg(reinterpret_cast<const double2*>(Ag + NX * col + d));
            av[d] = a.x;
            av[d + 1 < NX ? d + 1 : d] = a.y;
          }
        } else {
#pragma unroll
          for (int d = 0; d < NX; ++d) av[d] = (m < ncols[rr]) ? __ldg(Ag + NX * col + d) : 0.0;
        }
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          if (rr < NR) Areg[rr < NR ? rr : 0][m][d] = av[d];
          else As[(((rr - NR) * NI + m) * NX + d) * 32 + lane] = av[d];
        }
      }
    // ---- chol(Q_k) (every lane; n_x <= 4)
    double Lq[NX * NX];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Lq[i] = rd.Q[km * NX * NX + i];
    if (small_chol(Lq, NX) && !(rd.x0_in && rd.v_in)) {  // warp-uniform; Q is only used for the draws (:55, :71)
      if (lane == 0) atomicOr(rd.status, 16);
      continue;
    }
    // ---- noise of the whole horizon, one step per lane:  v_vec[:, t, k] (:72), e_vec[:, t, k] (:81)
    for (int64_t t = lane; t < H; t += 32) {
      double z[4];
      if (rd.v_in) {  // supplied v_vec (:67)
#pragma unroll
        for (int d = 0; d < NX; ++d) vs[t * NX + d] = rd.v_in[(k * H + t) * NX + d];
      } else {
        rollout_normals<(NX + 1) / 2>(rd.seed, cid, 2, (uint32_t)(t * ((NX + 1) / 2)), z);
#pragma unroll
        for (int d = 0; d < NX; ++d) {
          double acc = 0.0;
#pragma unroll
          for (int e2 = 0; e2 <= d; ++e2) acc = __fma_rn(Lq[d + NX * e2], z[e2], acc);
          vs[t * NX + d] = acc;
          if (rd.v) rd.v[(k * H + t) * NX + d] = acc;
        }
      }
      if (rd.e_in) {  // supplied e_vec (:77)
        for (int o = 0; o < ny; ++o) es[t * ny + o] = rd.e_in[(k * H + t) * ny + o];
      } else {
        const int npairs = (ny + 1) / 2;
        if (npairs == 1) rollout_normals<1>(rd.seed, cid, 3, (uint32_t)t, z);
        else rollout_normals<2>(rd.seed, cid, 3, (uint32_t)(t * 2), z);
        for (int o = 0; o < ny; ++o) {
          double acc = 0.0;
          for (int e2 = 0; e2 <= o; ++e2) acc = __fma_rn(rd.Le[o + ny * e2], z[e2], acc);
          es[t * ny + o] = acc;
          if (rd.e) rd.e[(k * H + t) * ny + o] = acc;
        }
      }
    }
    // ---- x0 noise (:62), every lane redundantly
    double x0n[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) x0n[d] = 0.0;
    if (!rd.x0_in) {
      double z[4];
      rollout_normals<(NX + 1) / 2>(rd.seed, cid, 1, 0u, z);
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int e2 = 0; e2 <= d; ++e2) acc = __fma_rn(Lq[d + NX * e2], z[e2], acc);
        x0n[d] = acc;
      }
    }
    // ---- star = systematic_resampling(w_m1, 1) (:58), or the x_T the owning GPU already drew (packed samples)
    if (!rd.x0_in) {
      if (rd.x_star) {
        if (lane < NX) xin[lane] = rd.x_star[km * NX + lane];
      } else {
        const int64_t star = rw_draw_star(rd.w_m1 + km * N, N, rd.seed, cid, rd.status);
        if (lane < NX) xin[lane] = rd.x_m1[(km * N + star) * NX + lane];
      }
    }
    __syncwarp();
    // ---- x0 = f(x_m1, u_m1) + rand(MvNormal(0, Q))  (:62).  The tree leaves component rw_comp(lane) of F in every
    // lane; the lanes add that component's noise and the first holders store X[:, t] in shared memory, from where
    // the sine lanes read their input of the next step.  Global outputs are written once, after the horizon.
    const int comp = rw_comp<NX>(lane);
    const int compc = comp < NX ? comp : 0;
    const bool holder = lane < (NX == 1 ? 1 : (NX == 2 ? 2 : 4)) && comp < NX;
    double Fb;
    if (rd.x0_in) {  // supplied x_vec_0 (:48)
      if (lane < NX) xs[lane] = rd.x0_in[k * NX + lane];
    } else {
      rw_eval_F<NX, NI, RS>(fd, sh, tkw, Areg, As, ncols, c, kl, ls, xin, rd.u_m1 + km * nu, Fb);
      double nz0 = x0n[0];
#pragma unroll
      for (int d = 1; d < NX; ++d) nz0 = (compc == d) ? x0n[d] : nz0;
      if (holder) xs[compc] = __dadd_rn(Fb, nz0);
    }
    __syncwarp();
    // ---- horizon: X[t+1] = f(X[t], u_t) + v (:121)
    for (int64_t t = 0; t < H; ++t) {
      rw_eval_F<NX, NI, RS>(fd, sh, tkw, Areg, As, ncols, c, kl, ls, xs + t * NX, us + (int64_t)nu * t, Fb);
      if (holder) xs[(t + 1) * NX + compc] = __dadd_rn(Fb, vs[t * NX + compc]);
      __syncwarp();
    }
    // ---- outputs: x_vec_0, X (coalesced copy of the trajectory), Y[t] = g(X[t]) + e (:122), one step per lane
    if (rd.x0 && !rd.x0_in && lane < NX) rd.x0[k * NX + lane] = xs[lane];
    if (rd.X)
      for (int64_t i = lane; i < (H + 1) * NX; i += 32) rd.X[k * (H + 1) * NX + i] = xs[i];
    if (rd.Y)
      for (int64_t t = lane; t < H; t += 32)
        for (int o = 0; o < ny; ++o) {
          double g = 0.0;
#pragma unroll
          for (int d = 0; d < NX; ++d) g = __fma_rn(fd.C[o + ny * d], xs[t * NX + d], g);
          rd.Y[(k * H + t) * ny + o] = __dadd_rn(g, es[t * ny + o]);
        }
    __syncwarp();  // vs / es are rewritten by the next scenario
  }
}

// ---------------------------------------------------------------------------------------------------------
// mean_rmse of test_prediction (Julia/src/particle_Gibbs.jl:311): sum over the column-major (n_y, T_test, K*k_n) array
// of (y_sim - y_test)^2 as the balanced tree the oracle defines (tile sums, then the tree over the tiles).
static __global__ void __launch_bounds__(TPB) k_sqerr_tiles(const double* __restrict__ Y, const double* __restrict__ y_test,
                                                      int64_t M, int64_t per, double* __restrict__ tile_sums) {
  __shared__ double sm[NWARP + 1];
  const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * IPT;
  double v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t n = base + i;
    double d = 0.0;
    if (n < M) d = __dsub_rn(Y[n], y_test[n % per]);
    v[i] = __dmul_rn(d, d);
  }
  const double s = block_tree_sum(tree4(v), sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = s;
}
static __global__ void __launch_bounds__(TPB) k_tree_final(const double* __restrict__ p, int64_t n, double* __restrict__ out) {
  __shared__ double sm[NWARP + 1];
  const double s = block_tree_sum_array(p, n, sm);
  if (threadIdx.x == 0) *out = s;
}

}  // namespace pgb
