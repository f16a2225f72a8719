// pgb_trace_f32.cuh — backward trace of the ancestor path (:213-218) in the F32 precision mode: the particle history
// holds FP32 values; x_prim (FP64 storage) receives them widened.  Stored history: a plain walk.  Replay mode: the
// lineage is recomputed with the F32 arithmetic of the sweep kernel (pgb_real.cuh = oracle eval_f_f32), bit for bit.
#pragma once
#include "pgb_filter.cuh"
#include "pgb_mathf.h"

namespace pgb {

template <int NX>
__global__ void k_trace_f32(FilterDev fd, const int32_t* __restrict__ star_idx, int64_t* __restrict__ star_out) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= fd.nc) return;
  const int64_t N = fd.N, T = fd.T;
  int64_t star = star_idx[chain];
  if (star < 0) star = 0;
  star_out[chain] = star + 1;
  double* xp = fd.xprim + (int64_t)chain * NX * T;
  const float* xh = reinterpret_cast<const float*>(fd.xh) + (int64_t)chain * T * NX * N;
  const int32_t* ah = fd.ah + (int64_t)chain * T * N;
  for (int d = 0; d < NX; ++d) xp[d + NX * (T - 1)] = (double)xh[((T - 1) * NX + d) * N + star];
  for (int64_t t = T - 2; t >= 0; --t) {
    star = ah[(t + 1) * N + star];
    if (star < 0) star = 0;
    for (int d = 0; d < NX; ++d) xp[d + NX * t] = (double)xh[(t * NX + d) * N + star];
  }
}

// one 128-thread CTA per chain; dynamic shared memory: phi[n_phi] floats; noise: [nc][T][NX] doubles holding FP32 values
template <int NX>
__global__ void __launch_bounds__(128) k_trace_replay_f32(FilterDev fd, const int32_t* __restrict__ star_idx,
                                                          int64_t* __restrict__ star_out, double* __restrict__ noise,
                                                          int first) {
  extern __shared__ float phi_smf[];
  __shared__ float tk[PGB_MAX_Z_][MAXC];
  __shared__ float xs[4];
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
  for (int64_t t = tid; t < T; t += 128) {  // L z along the lineage (state-independent): all steps in parallel
    const int64_t b = lin[t];
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    const bool pinned0 = (t == 0 && b == N - 1);
    if (!pinned0) {
      if (fd.philox) {
        pgf_stream_normals(fd.seed, t == 0 ? PGB_STREAM_Z0 : PGB_STREAM_Z, fd.chain_base + chain, fd.sweep, (uint32_t)t, (uint32_t)b, NX, z);
      } else {
        const double* inj = t == 0 ? fd.Z0 + (int64_t)chain * NX * (N - 1) : fd.Z + ((int64_t)chain * T + t) * NX * N;
        for (int d = 0; d < NX; ++d) z[d] = (float)inj[b * NX + d];
      }
    }
    for (int d = 0; d < NX; ++d) {
      float acc = 0.0f;
      for (int e = 0; e <= d; ++e)
        acc = __fmaf_rn(t == 0 ? (float)fd.x0_chol[d + NX * e] : (float)fd.Lq[chain * 16 + d + NX * e], z[e], acc);
      nz_[t * NX + d] = (double)acc;
    }
  }
  __syncthreads();
  const double* A = fd.A + (int64_t)chain * NX * fd.n_phi;
  const int nz = fd.n_z, n_phi = fd.n_phi;
  int ns = 0;
  if (fd.basis != 0)
    for (int k = 0; k < nz; ++k) ns += fd.hg_count[k];
  int sk = 0, sj = tid;
  if (tid < ns)
    while (sj >= fd.hg_count[sk]) sj -= fd.hg_count[sk++];
  for (int64_t t = 0; t < T; ++t) {
    const int64_t b = lin[t];
    const bool pinned = (b == N - 1) && (t == 0 || !first);  // slot N holds the previous sweep's trajectory
    if (pinned) {
      if (tid < NX) xs[tid] = (float)xp[tid + NX * t];
    } else if (t == 0) {
      if (tid < NX) xs[tid] = __fadd_rn((float)fd.x0_mean[tid], (float)nz_[tid]);
    } else {
      const double* u = fd.u + (int64_t)fd.n_u * (t - 1);
      if (fd.basis == 0) {
        if (tid == 0) {
          const float x0 = xs[0], x1 = xs[NX > 1 ? 1 : 0], u0 = (float)u[0];
          float phi[5];
          phi[0] = __fmul_rn(0.1f, x0);
          phi[1] = __fmul_rn(0.1f, x1);
          phi[2] = u0;
          phi[3] = __fmul_rn(__fmul_rn(0.01f, pgf_cos(__fmul_rn(3.0f, x0))), x1);
          phi[4] = __fmul_rn(__fmul_rn(0.1f, pgf_sin(__fmul_rn(2.0f, x1))), u0);
          float F[NX];
          for (int d = 0; d < NX; ++d) {
            float acc = 0.0f;
            for (int i = 0; i < 5; ++i) acc = __fmaf_rn((float)A[d + NX * i], phi[i], acc);
            F[d] = acc;
          }
          for (int d = 0; d < NX; ++d) xs[d] = __fadd_rn(F[d], (float)nz_[t * NX + d]);
        }
      } else {
        if (tid < ns) {
          const float zk = (sk < fd.n_u) ? (float)u[sk] : xs[sk - fd.n_u];
          const float arg = __fmul_rn((float)fd.hg_pij2l[sk][sj], __fadd_rn(zk, (float)fd.hg_L[sk]));
          tk[sk][sj] = __fmul_rn((float)fd.hg_lsi[sk], pgf_sin(arg));
        }
        __syncthreads();
        for (int i = tid; i < n_phi; i += 128) {  // phi_i = ((1*t_0)*t_1)*..., dimension 0 slowest
          int dg[PGB_MAX_Z_];
          int rem = i;
#pragma unroll
          for (int k = PGB_MAX_Z_ - 1; k >= 0; --k)
            if (k < nz) {
              dg[k] = rem % fd.hg_count[k];
              rem /= fd.hg_count[k];
            }
          float p = 1.0f;
#pragma unroll
          for (int k = 0; k < PGB_MAX_Z_; ++k)
            if (k < nz) p = __fmul_rn(p, tk[k][dg[k]]);
          phi_smf[i] = p;
        }
        __syncthreads();
        if (tid < NX) {
          float acc = 0.0f;
          for (int i = 0; i < n_phi; ++i) acc = __fmaf_rn((float)A[tid + NX * i], phi_smf[i], acc);
          xs[tid] = __fadd_rn(acc, (float)nz_[t * NX + tid]);
        }
      }
    }
    __syncthreads();
    if (tid < NX) xp[tid + NX * t] = (double)xs[tid];
  }
}

}  // namespace pgb
