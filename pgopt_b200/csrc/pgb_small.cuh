// pgb_small.cuh — the CPF-AS time loop (Julia/src/particle_Gibbs.jl:160-200) for SMALL particle counts (N <= 32, the
// reference's own examples use N = 30): one WARP per chain, lane n = particle n, the whole T loop in one launch.
//
// The tile kernels (pgb_persist.cuh) are built for throughput at N >> 10^4: four grid barriers, block scans and a
// dozen dependent global-memory round trips per step cost ~50 us per time step however few particles there are
// (N = 30, T = 2000: 10 sweeps/s on a B200, slower than one host core).  At N <= 32 nothing needs a barrier:
//   * sum(v) is the balanced tree over the index range padded to 32 = the xor-butterfly over the lanes,
//   * the strictly sequential recurrence q = q + W[n] (:28) is simply executed — every lane walks the N additions
//     itself (broadcast reads from shared memory) and counts how many prefixes lie below ITS draw u_i, where u_i is
//     the repeated sum u = u + 1/N (:31) taken i times: idx[i] = #{n >= 0 : q_n < u_i}, exactly the while loop of
//     :26-29 because both sequences are non-decreasing (weights with n_draw <= 32 only: u needs at most 31 steps),
//   * the gather F[a_j] goes through shared memory, everything else (basis functions, ancestor weights, normals,
//     log-weights) is per-lane arithmetic with the same device functions as the other kernels.
// Same definitions, same order, same bits as the oracle and as the tile kernels (tests run all of them).  The ancestor
// history a[t, :], the particle history x_pf (always stored: N*T is tiny here) and the final weights are left in
// the handle's buffers exactly as the tile kernels leave them, so trace / statistics / MNIW run unchanged.
#pragma once
#include "pgb_filter.cuh"

namespace pgb {

constexpr int SMALL_MAXN = 32;

struct SmallRes {
  int idx;       // 1-based index (0: the reference's index 0; N on overrun)
  int status;    // PGB_STATUS bits raised by this draw
};

// systematic_resampling(v, n_draw) (:17-34) for the warp: v = this lane's un-normalised weight (0 for lanes >= N),
// rnd = rand() (:21).  Lane i < n_draw receives idx[i]; Ws is a 32-double shared scratch.
__device__ __forceinline__ SmallRes small_resample(double v, int N, int n_draw, double rnd, double* Ws, int lane) {
  SmallRes r;
  r.status = 0;
  const double S = warp_tree_sum(v);                      // :19
  const double W = (lane < N) ? __ddiv_rn(v, S) : 0.0;
  if (lane < N && !(W >= 0.0 && W <= 1.7976931348623157e308)) r.status |= 4;
  __syncwarp();
  Ws[lane] = W;
  __syncwarp();
  const double inc = __ddiv_rn(1.0, (double)n_draw);
  double u = __dmul_rn(inc, rnd);                         // :21
  // the draws' grid u_i (:31, the increment taken `lane` times) and the prefix sums q_n (:28) are two independent
  // sequential chains advanced by one unrolled loop; the prefixes stay in registers and are counted against the
  // lane's final u afterwards (32 independent comparisons).  Beyond N the loop adds +0.0, which leaves q = q_N.
  double qs[SMALL_MAXN];
  double q = 0.0;                                         // :23
  const int nu = (lane < n_draw - 1) ? lane : n_draw - 1;  // increments of u this lane takes
#pragma unroll
  for (int n = 0; n < SMALL_MAXN; ++n) {
    qs[n] = q;                                            // q_n (= q_N for n >= N: lanes >= N stored W = 0)
    q = __dadd_rn(q, Ws[n]);
    if (n < nu) u = __dadd_rn(u, inc);
  }
  // all 32 entries are compared: the entries n >= N equal q_N, so a count above N means q_N < u (the overrun case)
  int cnt = 0;
#pragma unroll
  for (int n = 0; n < SMALL_MAXN; ++n) cnt += (qs[n] < u) ? 1 : 0;
  if (cnt > N) cnt = N;
  if (lane < n_draw) {
    if (cnt == N && q < u) r.status |= 1;                 // q_N < u_i: BoundsError in the reference
    if (cnt == 0) r.status |= 2;
  }
  r.idx = cnt;
  return r;
}

// The state-independent draws of a sweep — the process-noise normals z[t, n, :] (Philox + Box-Muller: a third of the
// instructions of a step) and the two uniforms per step — taken off the dependent chain: k_predraw_small fills
// Zp [nc][T][N][NX], Up / Ap [nc][T] with the very values k_sweep_small would generate in place (same counters, same
// routines), thousands of threads wide, and the time loop reads them back like injected streams.
// The two draws of a conditional step — a[t, 1:N-1] = systematic_resampling(w, N-1) (:172) and
// a[t, N] = systematic_resampling(waN, 1)[1] (:181) — are independent of each other: their tree sums, divisions and
// sequential prefix walks are interleaved in one pass (three dependent chains side by side instead of two walks one after
// the other).  Same operations per draw as small_resample; inc_r = 1 / n_draw is loop-invariant and passed in.
struct SmallPair {
  int idx_r, status_r;  // resampling: lane i < n_draw receives idx[i]
  int idx_a, status_a;  // ancestor of slot N (every lane holds the count; lane 0 is the draw)
};
__device__ __forceinline__ SmallPair small_resample_pair(double vr, double va, int N, int n_draw, double inc_r, double rnd_r,
                                                         double rnd_a, double* Wr_s, double* Wa_s, int lane) {
  SmallPair r;
  r.status_r = r.status_a = 0;
  double sr = vr, sa = va;                                 // :19, both draws
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    sr = __dadd_rn(sr, __shfl_xor_sync(0xffffffffu, sr, o));
    sa = __dadd_rn(sa, __shfl_xor_sync(0xffffffffu, sa, o));
  }
  const double Wr = (lane < N) ? __ddiv_rn(vr, sr) : 0.0;
  const double Wa = (lane < N) ? __ddiv_rn(va, sa) : 0.0;
  if (lane < N && !(Wr >= 0.0 && Wr <= 1.7976931348623157e308)) r.status_r |= 4;
  if (lane < N && !(Wa >= 0.0 && Wa <= 1.7976931348623157e308)) r.status_a |= 4;
  __syncwarp();
  Wr_s[lane] = Wr;
  Wa_s[lane] = Wa;
  __syncwarp();
  double ur = __dmul_rn(inc_r, rnd_r);                     // :21
  const double ua = __dmul_rn(1.0, rnd_a);                 // n_draw = 1: 1 / 1 = 1
  double qs[SMALL_MAXN];
  double qr = 0.0, qa = 0.0;                               // :23
  int cnt_a = 0;
  const int nu = (lane < n_draw - 1) ? lane : n_draw - 1;
#pragma unroll
  for (int n = 0; n < SMALL_MAXN; ++n) {
    qs[n] = qr;
    cnt_a += (qa < ua) ? 1 : 0;                            // the single draw's u never moves: counted on the fly
    qr = __dadd_rn(qr, Wr_s[n]);
    qa = __dadd_rn(qa, Wa_s[n]);
    if (n < nu) ur = __dadd_rn(ur, inc_r);
  }
  int cnt = 0;
#pragma unroll
  for (int n = 0; n < SMALL_MAXN; ++n) cnt += (qs[n] < ur) ? 1 : 0;
  if (cnt > N) cnt = N;
  if (lane < n_draw) {
    if (cnt == N && qr < ur) r.status_r |= 1;
    if (cnt == 0) r.status_r |= 2;
  }
  r.idx_r = cnt;
  if (cnt_a > N) cnt_a = N;
  if (lane < 1) {
    if (cnt_a == N && qa < ua) r.status_a |= 1;
    if (cnt_a == 0) r.status_a |= 2;
  }
  r.idx_a = cnt_a;
  return r;
}

template <int NX>
__global__ void __launch_bounds__(128) k_predraw_small(const __grid_constant__ FilterDev fd, double* __restrict__ Zp,
                                                       double* __restrict__ Up, double* __restrict__ Ap, int first) {
  const int chain = blockIdx.y, lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) + 1;
  const int N = (int)fd.N;
  if (t >= fd.T) return;
  const uint32_t cid = fd.chain_base + chain;
  if (lane < N) {
    constexpr int NP = (NX + 1) / 2;
    pgb_u32x4 ctr[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      ctr[p].v[0] = (uint32_t)(lane * NP + p);
      ctr[p].v[1] = (uint32_t)t;
      ctr[p].v[2] = fd.sweep;
      ctr[p].v[3] = ((uint32_t)PGB_STREAM_Z << 24) | (cid & 0x00ffffffu);
    }
    vphilox_wide<NP>(ctr, (uint32_t)fd.seed, (uint32_t)(fd.seed >> 32));
    double z0[NP], z1[NP];
    vnormal_pairs_wide<NP>(ctr, z0, z1);
    double* zo = Zp + (((int64_t)chain * fd.T + t) * N + lane) * NX;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      zo[2 * p] = z0[p];
      if (2 * p + 1 < NX) zo[2 * p + 1] = z1[p];
    }
  }
  if (lane == 0) {
    pgb_u32x4 b = pgb_stream_block(fd.seed, PGB_STREAM_URES, cid, fd.sweep, (uint32_t)t, 0);
    Up[(int64_t)chain * fd.T + t] = pgb_u53_halfopen(b.v[0], b.v[1]);
    if (!first) {
      b = pgb_stream_block(fd.seed, PGB_STREAM_UANC, cid, fd.sweep, (uint32_t)t, 0);
      Ap[(int64_t)chain * fd.T + t] = pgb_u53_halfopen(b.v[0], b.v[1]);
    }
  }
}

template <int NX>
__global__ void __launch_bounds__(32, 1) k_sweep_small(const __grid_constant__ FilterDev fd, int32_t* __restrict__ star_idx,
                                                    int first, const double* __restrict__ Zp,
                                                    const double* __restrict__ Up, const double* __restrict__ Ap) {
  extern __shared__ __align__(16) double sm_small[];
  double* A_sm = sm_small;                                   // [NX * n_phi]
  double* Fs = A_sm + ((NX * fd.n_phi + 1) & ~1);             // [NX][32]  F of every particle (gather source)
  double* Ws = Fs + NX * 32;                                  // [32]      normalised weights of the running draw
  double* Ws2 = Ws + 32;                                      // [32]      ... of the ancestor draw (conditional sweeps)
  const int chain = blockIdx.x, lane = threadIdx.x;
  const int N = (int)fd.N;
  const int64_t T = fd.T;
  const bool live = lane < N;
  const uint32_t cid = fd.chain_base + chain;
  const bool hil355 = fd.basis != 0 && NX == 2 && fd.n_z == 3 && fd.n_u == 1 && fd.hg_count[0] == 5 &&
                      fd.hg_count[1] == 5 && fd.hg_count[2] == 5;  // the reference's generic example
  int stat = 0;
  {
    const double* Ag = fd.A + (int64_t)chain * NX * fd.n_phi;
    for (int i = lane; i < NX * fd.n_phi; i += 32) A_sm[i] = Ag[i];
  }
  double Lq[NX * NX];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) Lq[i] = fd.Lq[chain * 16 + i];
  const double c0 = fd.c0[chain];
  const double* xprim = fd.xprim + (int64_t)chain * NX * T;
  double* wb = fd.wbuf + (int64_t)chain * N;
  __syncwarp();

  // ---------------------------------------------------------------- t = 0: initial particles (:160-165)
  double x[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) x[d] = 0.0;
  double lw = -INFINITY;
  if (live) {
    if (lane == N - 1) {
#pragma unroll
      for (int d = 0; d < NX; ++d) x[d] = xprim[d];
    } else {
      double z[NX];
      draw_normals<NX>(fd, chain, PGB_STREAM_Z0, 0, lane, fd.Z0 ? fd.Z0 + (int64_t)chain * NX * (N - 1) : nullptr, z);
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e <= d; ++e) acc = __fma_rn(fd.x0_chol[d + NX * e], z[e], acc);
        x[d] = __dadd_rn(fd.x0_mean[d], acc);
      }
    }
    double* x0 = x_row(fd, chain, 0);
#pragma unroll
    for (int d = 0; d < NX; ++d) x0[(int64_t)d * N + lane] = x[d];
    lw = log_weight<NX>(fd, x, fd.y);
    if (!(lw == lw)) stat |= 4;
  }

  // the per-step scalars (u_{t-1}, y_t, x'_t, the uniforms) are read sequentially in t: lane l < 5 pulls the cache line
  // of stream l 16 steps ahead into L1 so that no step of this latency-bound loop waits for L2 / HBM (base and stride
  // per lane fixed here: the first version re-derived five addresses in every step, 14 % of the loop's stall samples)
  const double* pf_base = nullptr;
  int64_t pf_stride = 0;
  {
    const double* ures = Up ? Up : (fd.philox ? nullptr : fd.Ures);
    const double* uanc = Ap ? Ap : (fd.philox ? nullptr : fd.Uanc);
    if (lane == 0) { pf_base = fd.u + (int64_t)fd.n_u * 15; pf_stride = fd.n_u; }
    else if (lane == 1) { pf_base = fd.y + (int64_t)fd.n_y * 16; pf_stride = fd.n_y; }
    else if (lane == 2) { pf_base = xprim + NX * 16; pf_stride = NX; }
    else if (lane == 3 && ures) { pf_base = ures + (int64_t)chain * T + 16; pf_stride = 1; }
    else if (lane == 4 && uanc) { pf_base = uanc + (int64_t)chain * T + 16; pf_stride = 1; }
  }
  const double inc_draw = __ddiv_rn(1.0, (double)(first ? N : N - 1));   // 1 / n_draw of the resampling draw (:21, :31)
  for (int64_t t = 1; t <= T; ++t) {
    if (pf_base && t + 16 < T) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf_base + t * pf_stride));
    // pre-drawn noise of this step: requested now, consumed after the resampling (the whole step hides the latency)
    double zpre[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) zpre[d] = 0.0;
    if (Zp && live && t < T) {
      const double* zin = Zp + (((int64_t)chain * T + t) * N + lane) * NX;
#pragma unroll
      for (int d = 0; d < NX; ++d) zpre[d] = zin[d];
    }
    // ---- weights of step t-1 (:198-199): e = exp(lw - max), w = e / sum(e)
    double m = lw;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    double e;
    {
      const double ea[1] = {live ? __dsub_rn(lw, m) : 0.0};
      double ev[1];
      vexp_wide<1>(ea, ev);  // branch-free, bit-identical to pgb_exp (the lanes' arguments fall in different ranges)
      e = live ? ev[0] : 0.0;
    }
    const double S1 = warp_tree_sum(e);
    const double w = live ? __ddiv_rn(e, S1) : 0.0;
    if (live) {
      wb[lane] = w;
      if (fd.whist) fd.whist[((int64_t)chain * T + (t - 1)) * N + lane] = w;
    }
    if (t == T) {
      // ---- star = systematic_resampling(w[T, :], 1)[1] (:213)
      double us;
      if (fd.philox) {
        const pgb_u32x4 b = pgb_stream_block(fd.seed, PGB_STREAM_USTAR, cid, fd.sweep, 0, 0);
        us = pgb_u53_halfopen(b.v[0], b.v[1]);
      } else {
        us = fd.Ustar[chain];
      }
      const SmallRes rs = small_resample(w, N, 1, us, Ws, lane);
      stat |= __shfl_sync(0xffffffffu, rs.status, 0) | (rs.status & 4);
      if (lane == 0) star_idx[chain] = rs.idx - 1;
      break;
    }
    const bool anc = !first;
    const double* up = fd.u + (int64_t)fd.n_u * (t - 1);
    double ur, ua = 0.0;
    if (Up) {
      ur = Up[(int64_t)chain * T + t];
      if (anc) ua = Ap[(int64_t)chain * T + t];
    } else if (fd.philox) {
      pgb_u32x4 b = pgb_stream_block(fd.seed, PGB_STREAM_URES, cid, fd.sweep, (uint32_t)t, 0);
      ur = pgb_u53_halfopen(b.v[0], b.v[1]);
      if (anc) {
        b = pgb_stream_block(fd.seed, PGB_STREAM_UANC, cid, fd.sweep, (uint32_t)t, 0);
        ua = pgb_u53_halfopen(b.v[0], b.v[1]);
      }
    } else {
      ur = fd.Ures[(int64_t)chain * T + t];
      if (anc) ua = fd.Uanc[(int64_t)chain * T + t];
    }
    // ---- F = A * phi(x_{t-1}, u_{t-1}) once per particle (:175 and :179 evaluate the same function)
    double F[NX];
    if (hil355) {
      eval_F_hilbert355_one<NX>(fd, A_sm, x, up, F);
    } else if (fd.basis == 0) {
      // examples/PG_OCP_known_basis_functions_Altro.jl:34 with the two trigonometric chains interleaved, branch-free
      const double x1 = x[NX > 1 ? 1 : 0], u0 = up[0];
      const double ang[2] = {__dmul_rn(3.0, x[0]), __dmul_rn(2.0, x1)};
      double sn[2], cs[2];
      vsincos_wide<2>(ang, sn, cs);
      double phi[5];
      phi[0] = __dmul_rn(0.1, x[0]);
      phi[1] = __dmul_rn(0.1, x1);
      phi[2] = u0;
      phi[3] = __dmul_rn(__dmul_rn(0.01, cs[0]), x1);
      phi[4] = __dmul_rn(__dmul_rn(0.1, sn[1]), u0);
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) acc = __fma_rn(A_sm[d + NX * i], phi[i], acc);
        F[d] = acc;
      }
    } else {
      eval_F<NX>(fd, A_sm, x, up, F);
    }
    __syncwarp();
#pragma unroll
    for (int d = 0; d < NX; ++d) Fs[d * 32 + lane] = F[d];
    // ---- ancestor of the conditioned trajectory (:178-181)
    int a_last = 0;
    double wa_n = 0.0;  // normalised ancestor weight of this lane (:180)
    if (anc) {
      double r[NX], sq = 0.0;
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = __dsub_rn(F[d], xprim[d + NX * t]);
#pragma unroll
        for (int e2 = 0; e2 < d; ++e2) acc = __fma_rn(-Lq[d + NX * e2], r[e2], acc);
        r[d] = __ddiv_rn(acc, Lq[d + NX * d]);
        sq = __fma_rn(r[d], r[d], sq);
      }
      const double pa[1] = {__dsub_rn(c0, __dmul_rn(sq, 0.5))};
      double pv[1];
      vexp_wide<1>(pa, pv);
      double wa = live ? __dmul_rn(w, pv[0]) : 0.0;
      const double S3 = warp_tree_sum(wa);                 // :180
      wa = live ? __ddiv_rn(wa, S3) : 0.0;
      wa_n = wa;
    }
    // ---- a[t, 1:n_draw] = systematic_resampling(w[t-1, :], n_draw) (:172 | :185) and, in a conditional sweep,
    //      a[t, N] = systematic_resampling(waN, 1)[1] (:181): the two draws walk side by side
    const int n_draw = first ? N : N - 1;
    SmallRes rr;
    if (anc) {
      const SmallPair pr = small_resample_pair(w, wa_n, N, n_draw, inc_draw, ur, ua, Ws, Ws2, lane);
      rr.idx = pr.idx_r;
      rr.status = pr.status_r;
      a_last = __shfl_sync(0xffffffffu, pr.idx_a, 0);
      stat |= __shfl_sync(0xffffffffu, pr.status_a, 0) | (pr.status_a & 4);
    } else {
      rr = small_resample(w, N, n_draw, ur, Ws, lane);
    }
    stat |= rr.status;
    // ---- x_t = f(x_{t-1}[a]) + L z (:175 | :188); slot N stays the conditioned trajectory (:160)
    int32_t* at = fd.ah + ((int64_t)chain * T + t) * N;
    if (lane < n_draw) {
      const int aj = rr.idx < 1 ? 0 : rr.idx - 1;
      double z[NX];
      if (Zp) {
#pragma unroll
        for (int d = 0; d < NX; ++d) z[d] = zpre[d];
      } else if (fd.philox) {  // the branch-free Philox / Box-Muller routines (same bits as pgb_normal_pair)
        constexpr int NP = (NX + 1) / 2;
        pgb_u32x4 ctr[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          ctr[p].v[0] = (uint32_t)(lane * NP + p);
          ctr[p].v[1] = (uint32_t)t;
          ctr[p].v[2] = fd.sweep;
          ctr[p].v[3] = ((uint32_t)PGB_STREAM_Z << 24) | (cid & 0x00ffffffu);
        }
        vphilox_wide<NP>(ctr, (uint32_t)fd.seed, (uint32_t)(fd.seed >> 32));
        double z0[NP], z1[NP];
        vnormal_pairs_wide<NP>(ctr, z0, z1);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          z[2 * p] = z0[p];
          if (2 * p + 1 < NX) z[2 * p + 1] = z1[p];
        }
      } else {
        const double* zin = fd.Z + ((int64_t)chain * T + t) * NX * N;
#pragma unroll
        for (int d = 0; d < NX; ++d) z[d] = zin[lane * NX + d];
      }
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int e2 = 0; e2 <= d; ++e2) acc = __fma_rn(Lq[d + NX * e2], z[e2], acc);
        x[d] = __dadd_rn(Fs[d * 32 + aj], acc);
      }
      at[lane] = rr.idx - 1;
    } else if (live) {  // lane == N-1, conditional sweep
#pragma unroll
      for (int d = 0; d < NX; ++d) x[d] = xprim[d + NX * t];
      at[lane] = a_last - 1;
    }
    if (live) {
      double* xt = x_row(fd, chain, t);
#pragma unroll
      for (int d = 0; d < NX; ++d) xt[(int64_t)d * N + lane] = x[d];
      lw = log_weight<NX>(fd, x, fd.y + (int64_t)fd.n_y * t);  // :193-197
      if (!(lw == lw)) stat |= 4;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) stat |= __shfl_xor_sync(0xffffffffu, stat, o);
  if (lane == 0 && stat) atomicOr(&fd.status[chain], stat);
}

// Backward trace of the ancestor path (:213-218) for N <= 32, one warp per chain.  k_trace walks the T dependent
// loads a[t+1, star] with one thread (2.6 ms at T = 2000, a quarter of a small-N sweep); here lane n loads a[t, n] and
// x_pf[:, n, t] of a whole row — addresses that do not depend on the path, so the loads of many steps are in flight
// at once — and the hop itself is a shuffle.
template <int NX>
__global__ void __launch_bounds__(32) k_trace_small(FilterDev fd, const int32_t* __restrict__ star_idx,
                                                    int64_t* __restrict__ star_out) {
  const int chain = blockIdx.x, lane = threadIdx.x;
  const int N = (int)fd.N;
  const int64_t T = fd.T;
  int b = star_idx[chain];
  if (b < 0) b = 0;
  if (lane == 0) star_out[chain] = (int64_t)b + 1;
  double* xp = fd.xprim + (int64_t)chain * NX * T;
  const double* xh = fd.xh + (int64_t)chain * T * NX * N;
  const int32_t* ah = fd.ah + (int64_t)chain * T * N;
  const int ln = lane < N ? lane : 0;
#pragma unroll 8
  for (int64_t t = T - 1; t >= 0; --t) {
    double xv[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) xv[d] = __ldg(xh + (t * NX + d) * N + ln);
    const int a = (t > 0) ? __ldg(ah + t * N + ln) : 0;
#pragma unroll
    for (int d = 0; d < NX; ++d) {
      const double xb = __shfl_sync(0xffffffffu, xv[d], b);
      if (lane == 0) xp[d + NX * t] = xb;
    }
    b = __shfl_sync(0xffffffffu, a, b);  // star = a[t, star]  (:216)
    if (b < 0) b = 0;
  }
}

// ==========================================================================================================
// k_sweep_cta — the same idea for mid-size particle counts (32 < N <= 1024): one CTA per chain, thread n = particle n,
// the whole T loop in one ordinary launch.  sum(v) = warp butterfly, then the butterfly over the warp totals (aligned
// 32-leaf subtrees of the oracle's balanced tree); the sequential prefix q_n = q_{n-1} + W[n] (:28) is walked ONCE by
// thread 0 into shared memory while every other thread builds its own u_i = u + 1/N + ... (:31, i additions) — the two
// sequential chains of the reference run side by side — and idx[i] = #{n : q_n < u_i} is a binary search in the
// monotone prefix array.  Everything else is the per-particle arithmetic of k_sweep_small.
// ==========================================================================================================
constexpr int CTA_MAXN = 1024;

struct CtaShared {
  double red[2][32];   // warp partials of the block reductions (double-buffered)
  int bci[4];          // broadcast: index / status of a single draw
};

__device__ __forceinline__ double cta_tree_sum(double v, CtaShared& cs, int W, int& flip) {
  v = warp_tree_sum(v);
  if (W == 1) return v;
  double* r = cs.red[flip];
  flip ^= 1;
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  return warp_tree_sum(lane < W ? r[lane] : 0.0);
}
__device__ __forceinline__ double cta_max(double v, CtaShared& cs, int W, int& flip) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (W == 1) return v;
  double* r = cs.red[flip];
  flip ^= 1;
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double p = lane < W ? r[lane] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p = fmax(p, __shfl_xor_sync(0xffffffffu, p, o));
  return p;
}

// q_n = q_{n-1} + W[n] (:28) for n = 0 .. Np-1 by the calling thread, strictly left to right; qs[n] = q_n, qs[Np] = q_Np
// (= q_N: W[n] = 0 beyond N).  The weights are loaded eight at a time so that the loads do not wait for the stores.
__device__ __forceinline__ void cta_prefix_walk(const double* __restrict__ Wv, double* __restrict__ qv, int Np) {
  double q = 0.0;  // :23
  for (int n = 0; n < Np; n += 8) {
    double wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = Wv[n + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      qv[n + j] = q;
      q = __dadd_rn(q, wv[j]);
    }
  }
  qv[Np] = q;
}

// systematic_resampling(v, n_draw) (:17-34) for the CTA; thread i < n_draw receives idx[i].
__device__ __forceinline__ SmallRes cta_resample(double v, int N, int n_draw, double rnd, double* Ws, double* qs,
                                                 CtaShared& cs, int W, int& flip) {
  const int tid = threadIdx.x;
  SmallRes r;
  r.status = 0;
  const double S = cta_tree_sum(v, cs, W, flip);           // :19
  const double Wn = (tid < N) ? __ddiv_rn(v, S) : 0.0;
  if (tid < N && !(Wn >= 0.0 && Wn <= 1.7976931348623157e308)) r.status |= 4;
  __syncthreads();                                         // readers of the previous draw's Ws / qs are done
  Ws[tid] = Wn;
  __syncthreads();
  const double inc = __ddiv_rn(1.0, (double)n_draw);
  double u = __dmul_rn(inc, rnd);                          // :21
  if (tid == 0) {
    cta_prefix_walk(Ws, qs, (int)blockDim.x);
  } else {
    const int nu = (tid < n_draw - 1) ? tid : n_draw - 1;  // :31, the increment taken tid times
    for (int j = 0; j < nu; ++j) u = __dadd_rn(u, inc);
  }
  __syncthreads();
  // idx = #{n in 0..N-1 : q_n < u} (q_0 = 0): first position of the monotone prefix array that is >= u
  int lo = 0, hi = N;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (qs[mid] < u) lo = mid + 1;
    else hi = mid;
  }
  if (tid < n_draw) {
    if (lo == N && qs[N] < u) r.status |= 1;
    if (lo == 0) r.status |= 2;
  }
  r.idx = lo;
  return r;
}

// The two resampling calls of a conditional step — a[t, 1:N-1] = systematic_resampling(w, N-1) (:172) and
// a[t, N] = systematic_resampling(waN, 1)[1] (:181) — are independent of each other, so their sequential prefix walks
// run concurrently: thread 0 walks the weights, thread 32 (another warp) the ancestor weights, everybody else builds
// u_i meanwhile.  Thread i < n_draw receives idx[i] of the first call in .idx; *a_last = the index of the second call
// (all threads).  Same arithmetic per call as cta_resample.
__device__ __forceinline__ SmallRes cta_resample_pair(double v, int N, int n_draw, double rnd, double va, double rnda,
                                                      int* a_last, double* Ws, double* qs, double* Ws2, double* qs2,
                                                      CtaShared& cs, int W, int& flip) {
  const int tid = threadIdx.x, Np = (int)blockDim.x;
  SmallRes r;
  r.status = 0;
  const double S = cta_tree_sum(v, cs, W, flip);            // :19 (first call)
  const double Sa = cta_tree_sum(va, cs, W, flip);          // :19 (second call)
  const double Wn = (tid < N) ? __ddiv_rn(v, S) : 0.0;
  const double Wa = (tid < N) ? __ddiv_rn(va, Sa) : 0.0;
  if (tid < N && !(Wn >= 0.0 && Wn <= 1.7976931348623157e308)) r.status |= 4;
  if (tid < N && !(Wa >= 0.0 && Wa <= 1.7976931348623157e308)) r.status |= 4;
  __syncthreads();
  Ws[tid] = Wn;
  Ws2[tid] = Wa;
  __syncthreads();
  const double inc = __ddiv_rn(1.0, (double)n_draw);
  double u = __dmul_rn(inc, rnd);                           // :21
  if (tid == 0) cta_prefix_walk(Ws, qs, Np);
  if (tid == 32) cta_prefix_walk(Ws2, qs2, Np);
  if (tid != 0) {
    const int nu = (tid < n_draw - 1) ? tid : n_draw - 1;   // :31, the increment taken tid times
    for (int j = 0; j < nu; ++j) u = __dadd_rn(u, inc);
  }
  __syncthreads();
  int lo = 0, hi = N;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (qs[mid] < u) lo = mid + 1;
    else hi = mid;
  }
  if (tid < n_draw) {
    if (lo == N && qs[N] < u) r.status |= 1;
    if (lo == 0) r.status |= 2;
  }
  r.idx = lo;
  if (tid == 0) {  // the single draw of the second call: u = (1/1) * rnda
    const double ua = __dmul_rn(1.0, rnda);
    int l2 = 0, h2 = N;
    while (l2 < h2) {
      const int mid = (l2 + h2) >> 1;
      if (qs2[mid] < ua) l2 = mid + 1;
      else h2 = mid;
    }
    if (l2 == N && qs2[N] < ua) r.status |= 1;
    if (l2 == 0) r.status |= 2;
    cs.bci[0] = l2;
  }
  __syncthreads();
  *a_last = cs.bci[0];
  return r;
}

template <int NX>
__global__ void __launch_bounds__(CTA_MAXN, 1) k_sweep_cta(const __grid_constant__ FilterDev fd,
                                                           int32_t* __restrict__ star_idx, int first) {
  extern __shared__ __align__(16) double sm_cta[];
  const int tid = threadIdx.x;
  const int W = blockDim.x >> 5, Np = blockDim.x;
  double* A_sm = sm_cta;                                      // [NX * n_phi]
  double* Fs = A_sm + ((NX * fd.n_phi + 1) & ~1);             // [NX][Np]
  double* Ws = Fs + NX * Np;                                  // [Np]
  double* qs = Ws + Np;                                       // [Np + 2]
  double* Ws2 = qs + Np + 2;                                  // [Np]      the ancestor draw's weights / prefixes
  double* qs2 = Ws2 + Np;                                     // [Np + 2]
  CtaShared& cs = *reinterpret_cast<CtaShared*>(qs2 + Np + 2);
  const int chain = blockIdx.x;
  const int N = (int)fd.N;
  const int64_t T = fd.T;
  const bool live = tid < N;
  const uint32_t cid = fd.chain_base + chain;
  const bool hil355 = fd.basis != 0 && NX == 2 && fd.n_z == 3 && fd.n_u == 1 && fd.hg_count[0] == 5 &&
                      fd.hg_count[1] == 5 && fd.hg_count[2] == 5;
  int stat = 0, flip = 0;
  {
    const double* Ag = fd.A + (int64_t)chain * NX * fd.n_phi;
    for (int i = tid; i < NX * fd.n_phi; i += Np) A_sm[i] = Ag[i];
  }
  double Lq[NX * NX];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) Lq[i] = fd.Lq[chain * 16 + i];
  const double c0 = fd.c0[chain];
  const double* xprim = fd.xprim + (int64_t)chain * NX * T;
  double* wb = fd.wbuf + (int64_t)chain * N;
  __syncthreads();

  // ---------------------------------------------------------------- t = 0 (:160-165)
  double x[NX];
#pragma unroll
  for (int d = 0; d < NX; ++d) x[d] = 0.0;
  double lw = -INFINITY;
  if (live) {
    if (tid == N - 1) {
#pragma unroll
      for (int d = 0; d < NX; ++d) x[d] = xprim[d];
    } else {
      double z[NX];
      draw_normals<NX>(fd, chain, PGB_STREAM_Z0, 0, tid, fd.Z0 ? fd.Z0 + (int64_t)chain * NX * (N - 1) : nullptr, z);
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e <= d; ++e) acc = __fma_rn(fd.x0_chol[d + NX * e], z[e], acc);
        x[d] = __dadd_rn(fd.x0_mean[d], acc);
      }
    }
    double* x0 = x_row(fd, chain, 0);
#pragma unroll
    for (int d = 0; d < NX; ++d) x0[(int64_t)d * N + tid] = x[d];
    lw = log_weight<NX>(fd, x, fd.y);
    if (!(lw == lw)) stat |= 4;
  }

  for (int64_t t = 1; t <= T; ++t) {
    // ---- weights of step t-1 (:198-199)
    const double m = cta_max(lw, cs, W, flip);
    double e;
    {
      const double ea[1] = {live ? __dsub_rn(lw, m) : 0.0};
      double ev[1];
      vexp_wide<1>(ea, ev);
      e = live ? ev[0] : 0.0;
    }
    const double S1 = cta_tree_sum(e, cs, W, flip);
    const double w = live ? __ddiv_rn(e, S1) : 0.0;
    if (live) {
      wb[tid] = w;
      if (fd.whist) fd.whist[((int64_t)chain * T + (t - 1)) * N + tid] = w;
    }
    if (t == T) {  // star = systematic_resampling(w[T, :], 1)[1] (:213)
      double us;
      if (fd.philox) {
        const pgb_u32x4 b = pgb_stream_block(fd.seed, PGB_STREAM_USTAR, cid, fd.sweep, 0, 0);
        us = pgb_u53_halfopen(b.v[0], b.v[1]);
      } else {
        us = fd.Ustar[chain];
      }
      const SmallRes rs = cta_resample(w, N, 1, us, Ws, qs, cs, W, flip);
      stat |= (rs.status & 4);
      if (tid == 0) {
        stat |= rs.status;
        star_idx[chain] = rs.idx - 1;
      }
      break;
    }
    const bool anc = !first;
    const double* up = fd.u + (int64_t)fd.n_u * (t - 1);
    double ur, ua = 0.0;
    if (fd.philox) {
      pgb_u32x4 b = pgb_stream_block(fd.seed, PGB_STREAM_URES, cid, fd.sweep, (uint32_t)t, 0);
      ur = pgb_u53_halfopen(b.v[0], b.v[1]);
      if (anc) {
        b = pgb_stream_block(fd.seed, PGB_STREAM_UANC, cid, fd.sweep, (uint32_t)t, 0);
        ua = pgb_u53_halfopen(b.v[0], b.v[1]);
      }
    } else {
      ur = fd.Ures[(int64_t)chain * T + t];
      if (anc) ua = fd.Uanc[(int64_t)chain * T + t];
    }
    // ---- F = A * phi(x_{t-1}, u_{t-1})
    double F[NX];
    if (hil355) {
      eval_F_hilbert355_one<NX>(fd, A_sm, x, up, F);
    } else if (fd.basis == 0) {
      const double x1 = x[NX > 1 ? 1 : 0], u0 = up[0];
      const double ang[2] = {__dmul_rn(3.0, x[0]), __dmul_rn(2.0, x1)};
      double sn[2], csn[2];
      vsincos_wide<2>(ang, sn, csn);
      double phi[5];
      phi[0] = __dmul_rn(0.1, x[0]);
      phi[1] = __dmul_rn(0.1, x1);
      phi[2] = u0;
      phi[3] = __dmul_rn(__dmul_rn(0.01, csn[0]), x1);
      phi[4] = __dmul_rn(__dmul_rn(0.1, sn[1]), u0);
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) acc = __fma_rn(A_sm[d + NX * i], phi[i], acc);
        F[d] = acc;
      }
    } else {
      eval_F<NX>(fd, A_sm, x, up, F);
    }
    __syncthreads();  // the previous step's gathers from Fs are done
#pragma unroll
    for (int d = 0; d < NX; ++d) Fs[d * Np + tid] = F[d];
    // ---- ancestor of the conditioned trajectory (:178-181)
    int a_last = 0;
    double wan = 0.0;  // waN ./ sum(waN) (:180)
    if (anc) {
      double r[NX], sq = 0.0;
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = __dsub_rn(F[d], xprim[d + NX * t]);
#pragma unroll
        for (int e2 = 0; e2 < d; ++e2) acc = __fma_rn(-Lq[d + NX * e2], r[e2], acc);
        r[d] = __ddiv_rn(acc, Lq[d + NX * d]);
        sq = __fma_rn(r[d], r[d], sq);
      }
      const double pa[1] = {__dsub_rn(c0, __dmul_rn(sq, 0.5))};
      double pv[1];
      vexp_wide<1>(pa, pv);
      double wa = live ? __dmul_rn(w, pv[0]) : 0.0;
      const double S3 = cta_tree_sum(wa, cs, W, flip);      // :180
      wan = live ? __ddiv_rn(wa, S3) : 0.0;
    }
    // ---- a[t, 1:n_draw] = systematic_resampling(w[t-1, :], n_draw) (:172 | :185) and, for a conditional sweep,
    //      a[t, N] = systematic_resampling(waN, 1)[1] (:181): independent draws, their prefix walks run concurrently
    const int n_draw = first ? N : N - 1;
    SmallRes rr;
    if (anc) rr = cta_resample_pair(w, N, n_draw, ur, wan, ua, &a_last, Ws, qs, Ws2, qs2, cs, W, flip);
    else rr = cta_resample(w, N, n_draw, ur, Ws, qs, cs, W, flip);
    stat |= rr.status;
    // ---- x_t = f(x_{t-1}[a]) + L z (:175 | :188)
    int32_t* at = fd.ah + ((int64_t)chain * T + t) * N;
    if (tid < n_draw) {
      const int aj = rr.idx < 1 ? 0 : rr.idx - 1;
      double z[NX];
      if (fd.philox) {
        constexpr int NP = (NX + 1) / 2;
        pgb_u32x4 ctr[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          ctr[p].v[0] = (uint32_t)(tid * NP + p);
          ctr[p].v[1] = (uint32_t)t;
          ctr[p].v[2] = fd.sweep;
          ctr[p].v[3] = ((uint32_t)PGB_STREAM_Z << 24) | (cid & 0x00ffffffu);
        }
        vphilox_wide<NP>(ctr, (uint32_t)fd.seed, (uint32_t)(fd.seed >> 32));
        double z0[NP], z1[NP];
        vnormal_pairs_wide<NP>(ctr, z0, z1);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          z[2 * p] = z0[p];
          if (2 * p + 1 < NX) z[2 * p + 1] = z1[p];
        }
      } else {
        const double* zin = fd.Z + ((int64_t)chain * T + t) * NX * N;
#pragma unroll
        for (int d = 0; d < NX; ++d) z[d] = zin[tid * NX + d];
      }
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int e2 = 0; e2 <= d; ++e2) acc = __fma_rn(Lq[d + NX * e2], z[e2], acc);
        x[d] = __dadd_rn(Fs[d * Np + aj], acc);
      }
      at[tid] = rr.idx - 1;
    } else if (live) {  // tid == N-1, conditional sweep (:160)
#pragma unroll
      for (int d = 0; d < NX; ++d) x[d] = xprim[d + NX * t];
      at[tid] = a_last - 1;
    }
    if (live) {
      double* xt = x_row(fd, chain, t);
#pragma unroll
      for (int d = 0; d < NX; ++d) xt[(int64_t)d * N + tid] = x[d];
      lw = log_weight<NX>(fd, x, fd.y + (int64_t)fd.n_y * t);  // :193-197
      if (!(lw == lw)) stat |= 4;
    }
  }
  if (stat) atomicOr(&fd.status[chain], stat);
}

}  // namespace pgb
