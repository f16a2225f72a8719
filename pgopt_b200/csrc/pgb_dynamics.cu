// pgb_dynamics.cu — one step of the stacked K-scenario model and its Jacobians on device-resident samples
// (SURVEY.md section 8f row 1): RobotDynamics.discrete_dynamics[!] of PGS_model_obj (Julia/src/optimal_control_Altro.jl:79-107)
//   x_vec_n[k] = PG_samples[k].A * phi(x_vec[k], u) + v_vec[:, t+1, k]
// and what RobotDynamics.@autodiff (:51) differentiates out of it, d x_n[k] / d x[k] = A_k dphi/dx (block diagonal) and
// d x_n[k] / d u = A_k dphi/du — the inner loop of every Altro forward / backward pass, a K-way Julia loop allocating
// Array{Any} in the reference (examples/PG_OCP_generic_basis_functions_Altro.jl:204-217).
//
// Kernel: one warp per scenario, persistent grid.  The products A_k * (.) use the order the rollout defines (128 runs of
// c = ceil(n_phi/128) consecutive columns, fma in increasing i, balanced tree over the runs; oracle: blocked_dot), lane l
// owning runs 4l .. 4l+3, so x_next equals the rollout's X[t+1] bit for bit.  The 1 + n_z contractions (phi and its n_z
// partial derivatives) are accumulated in ONE pass over the lane's columns: A_k comes from HBM once (20 KB per scenario at
// n_phi = 625).  HBM traffic: 8 (n_x n_phi + 2 n_x + n_x n_z) bytes per scenario.
#include "pgb_internal.h"

using namespace pgb;

namespace {

struct DynDev {
  int64_t K, H, t;      // t < 0: no noise term
  int variant;          // 0: phi_sampling rounding, 1: phi_opt rounding (arg = (pi j (z + L)) / (2L))
  const double* A;      // [K][n_x*n_phi]
  const double* x;      // [K][n_x]
  const double* v;      // [K][H][n_x] or null
  double u[4];
  double *xn, *jx, *ju; // [K][n_x], [K][n_x*n_x], [K][n_x*n_u] (column-major blocks); jx / ju may be null
};

constexpr int DY_WPB = 4;
constexpr int DY_MAXC = 8;  // columns per run up to which the digit decode is hoisted (n_phi <= 1024)
constexpr int DY_MAXQ = 1 + PGB_MAX_Z_;  // phi and its partial derivatives

template <int NX>
__global__ void __launch_bounds__(DY_WPB * 32) k_dynamics(const __grid_constant__ FilterDev fd, const __grid_constant__ DynDev dd) {
  __shared__ double tks[DY_WPB][PGB_MAX_Z_ * MAXC], dks[DY_WPB][PGB_MAX_Z_ * MAXC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nz = fd.n_z, nu = fd.n_u, n_phi = fd.n_phi;
  const int c = (n_phi + 127) / 128;
  double* tk = tks[warp];
  double* dk = dks[warp];
  // mixed-radix digits of the lane's columns (4 runs x c columns), 4 bits per augmented dimension, decoded ONCE: the
  // first version decoded them per column, per pass and per scenario — ten integer divisions around five multiplications
  // (2.0 ms at K = 2*10^4, n_phi = 625; now the decode is off the scenario loop)
  __shared__ unsigned dgs[DY_WPB][4 * DY_MAXC][32];  // [run * DY_MAXC + column][lane]: conflict-free, lane-private
  const bool packed_ok = (c <= DY_MAXC) && fd.basis != 0;
  if (packed_ok) {
    for (int r = 0; r < 4; ++r)
      for (int m = 0; m < c; ++m) {
        unsigned pk = 0;
        const int i = (4 * lane + r) * c + m;
        if (i < n_phi) {
          int rem = i;
          for (int kk = nz - 1; kk >= 0; --kk) {
            pk |= (unsigned)(rem % fd.hg_count[kk]) << (4 * kk);
            rem /= fd.hg_count[kk];
          }
        }
        dgs[warp][r * DY_MAXC + m][lane] = pk;
      }
  }
  for (int64_t k = (int64_t)blockIdx.x * DY_WPB + warp; k < dd.K; k += (int64_t)gridDim.x * DY_WPB) {
    const double* Ak = dd.A + k * (int64_t)NX * n_phi;
    double x[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) x[d] = dd.x[k * NX + d];
    double out[NX];
    if (fd.basis == 0) {
      // known basis (examples/PG_OCP_known_basis_functions_Altro.jl:34), n_phi = 5: lane q < 4 evaluates quantity q
      // (0: phi, 1: d/du, 2: d/dx1, 3: d/dx2) — five columns, run j = column j, tree over the 128 (mostly empty) runs
      const double u0 = dd.u[0];
      const double c3 = pgb_cos(__dmul_rn(3.0, x[0])), s3 = pgb_sin(__dmul_rn(3.0, x[0]));
      const double s2 = pgb_sin(__dmul_rn(2.0, x[NX > 1 ? 1 : 0])), c2 = pgb_cos(__dmul_rn(2.0, x[NX > 1 ? 1 : 0]));
      const double x1 = x[NX > 1 ? 1 : 0];
      double col[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      if (lane == 0) {
        col[0] = __dmul_rn(0.1, x[0]); col[1] = __dmul_rn(0.1, x1); col[2] = u0;
        col[3] = __dmul_rn(__dmul_rn(0.01, c3), x1); col[4] = __dmul_rn(__dmul_rn(0.1, s2), u0);
      } else if (lane == 1) {
        col[2] = 1.0; col[4] = __dmul_rn(0.1, s2);
      } else if (lane == 2) {
        col[0] = 0.1; col[3] = __dmul_rn(__dmul_rn(0.01, -__dmul_rn(s3, 3.0)), x1);
      } else if (lane == 3) {
        col[1] = 0.1; col[3] = __dmul_rn(0.01, c3); col[4] = __dmul_rn(__dmul_rn(0.1, __dmul_rn(c2, 2.0)), u0);
      }
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        // run j holds the single product fma(A[d,j], col_j, 0); the tree over runs 0..127: ((p0+p1)+(p2+p3)) + ((p4+0)+(0+0)) ...
        double p[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) p[j] = __fma_rn(Ak[d + NX * j], col[j], 0.0);
        const double lo = __dadd_rn(__dadd_rn(p[0], p[1]), __dadd_rn(p[2], p[3]));
        const double hi = __dadd_rn(__dadd_rn(p[4], 0.0), 0.0);
        double s = __dadd_rn(lo, hi);
#pragma unroll
        for (int lvl = 0; lvl < 4; ++lvl) s = __dadd_rn(s, 0.0);  // the remaining tree levels add empty (+0.0) subtrees
        out[d] = s;
      }
      if (lane == 0) {
#pragma unroll
        for (int d = 0; d < NX; ++d)
          dd.xn[k * NX + d] = dd.v ? __dadd_rn(out[d], dd.v[(k * dd.H + dd.t) * NX + d]) : out[d];
      } else if (lane == 1 && dd.ju) {
#pragma unroll
        for (int d = 0; d < NX; ++d) dd.ju[k * NX * nu + d] = out[d];
      } else if ((lane == 2 || lane == 3) && dd.jx) {
#pragma unroll
        for (int d = 0; d < NX; ++d) dd.jx[k * NX * NX + d + NX * (lane - 2)] = out[d];
      }
      continue;
    }
    // ---- Hilbert basis: scaled sines and their derivatives, one (dimension, frequency) pair per lane
    {
      int ns = 0;
      for (int kk = 0; kk < nz; ++kk) ns += fd.hg_count[kk];
      for (int s = lane; s < ns; s += 32) {
        int sk = 0, sj = s;
        while (sj >= fd.hg_count[sk]) sj -= fd.hg_count[sk++];
        const double zk = (sk < nu) ? dd.u[sk] : x[(sk - nu) < NX ? (sk - nu) : 0];
        const double zl = __dadd_rn(zk, fd.hg_L[sk]);
        const double w = fd.hg_pij2l[sk][sj];
        double arg;
        if (dd.variant) {  // phi_opt: (pi*j*(z + L)) / (2L)
          const double pij = __dmul_rn(3.141592653589793, (double)(sj + 1));
          arg = __ddiv_rn(__dmul_rn(pij, zl), __dmul_rn(2.0, fd.hg_L[sk]));
        } else {
          arg = __dmul_rn(w, zl);
        }
        double sn, cs;
        pgb_sincos(arg, &sn, &cs);
        tk[sk * MAXC + sj] = __dmul_rn(fd.hg_lsi[sk], sn);
        dk[sk * MAXC + sj] = __dmul_rn(fd.hg_lsi[sk], __dmul_rn(cs, w));
      }
    }
    __syncwarp();
    // ONE pass over the lane's columns: A_k is read once (from HBM) and feeds phi and all n_z partial derivatives — the
    // first version made 1 + n_z passes and re-read A_k through L2 each time (L1 cannot hold 32 warps x 20 KB): 5 % of the
    // HBM roofline.  q = 0: phi; q >= 1: d phi / d z_{q-1}; every product left to right with the differentiated factor
    // in its place (the oracle's order).
    const bool want_j = (dd.ju != nullptr) || (dd.jx != nullptr);
    const int nq = want_j ? 1 + nz : 1;
    double part[4][DY_MAXQ * NX];
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
      double acc[DY_MAXQ][NX];
#pragma unroll
      for (int q = 0; q < DY_MAXQ; ++q)
#pragma unroll
        for (int d = 0; d < NX; ++d) acc[q][d] = 0.0;
      const int i0 = (4 * lane + r) * c;
#pragma unroll 1
      for (int m = 0; m < c; ++m) {
        const int i = i0 + m;
        if (i >= n_phi) break;
        unsigned pk;
        if (packed_ok) {
          pk = dgs[warp][r * DY_MAXC + m][lane];
        } else {
          pk = 0;
          int rem = i;
          for (int kk = nz - 1; kk >= 0; --kk) {
            pk |= (unsigned)(rem % fd.hg_count[kk]) << (4 * kk);
            rem /= fd.hg_count[kk];
          }
        }
        double a[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) a[d] = __ldg(Ak + (int64_t)NX * i + d);
        double tv[PGB_MAX_Z_], dv[PGB_MAX_Z_];
#pragma unroll
        for (int kk = 0; kk < PGB_MAX_Z_; ++kk)
          if (kk < nz) {
            const int o = kk * MAXC + ((pk >> (4 * kk)) & 15u);
            tv[kk] = tk[o];
            dv[kk] = want_j ? dk[o] : 0.0;
          }
#pragma unroll
        for (int q = 0; q < DY_MAXQ; ++q)
          if (q < nq) {
            double p = 1.0;
#pragma unroll
            for (int kk = 0; kk < PGB_MAX_Z_; ++kk)
              if (kk < nz) p = __dmul_rn(p, (kk == q - 1) ? dv[kk] : tv[kk]);
#pragma unroll
            for (int d = 0; d < NX; ++d) acc[q][d] = __fma_rn(a[d], p, acc[q][d]);
          }
      }
#pragma unroll
      for (int q = 0; q < DY_MAXQ; ++q)
#pragma unroll
        for (int d = 0; d < NX; ++d) part[r][q * NX + d] = acc[q][d];
    }
#pragma unroll 1
    for (int q = 0; q < nq; ++q) {
      const int mm = q - 1;
      if (q >= 1 && ((mm < nu && !dd.ju) || (mm >= nu && !dd.jx))) continue;  // warp-uniform
#pragma unroll
      for (int d = 0; d < NX; ++d) {
        double s = __dadd_rn(__dadd_rn(part[0][q * NX + d], part[1][q * NX + d]), __dadd_rn(part[2][q * NX + d], part[3][q * NX + d]));
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        out[d] = s;
      }
      if (lane == 0) {
        if (q == 0) {
#pragma unroll
          for (int d = 0; d < NX; ++d)
            dd.xn[k * NX + d] = dd.v ? __dadd_rn(out[d], dd.v[(k * dd.H + dd.t) * NX + d]) : out[d];
        } else if (mm < nu) {
#pragma unroll
          for (int d = 0; d < NX; ++d) dd.ju[k * NX * nu + d + NX * mm] = out[d];
        } else {
#pragma unroll
          for (int d = 0; d < NX; ++d) dd.jx[k * NX * NX + d + NX * (mm - nu)] = out[d];
        }
      }
    }
    __syncwarp();  // tk / dk are rewritten by the next scenario
  }
}

}  // namespace

// x_vec: host [K][n_x] (the reference's stacking: scenario k occupies rows n_x*(k-1)+1 .. n_x*k); u: [n_u];
// t >= 0 adds v_vec[:, t+1, k] of the set's resident process noise (the last pgb_rollout_dev's), t < 0 adds nothing.
extern "C" int pgb_dynamics_dev(pgb_samples* s, const double* x_vec, const double* u, int64_t t, int32_t phi_variant,
                                double* x_next, double* dfdx, double* dfdu) {
  if (!s) return set_err(nullptr, PGB_ERR_INVALID, "sample set is NULL");
  pgb_handle* h = &s->hh;
  CK(h, cudaSetDevice(h->cfg.device));
  if (!x_vec || !u || !x_next) return set_err(h, PGB_ERR_INVALID, "bad arguments");
  if (t >= 0 && (!s->have_scen || t >= s->H))
    return set_err(h, PGB_ERR_STATE, "t = %lld needs the process noise of a previous pgb_rollout_dev with H > t", (long long)t);
  const int nx = h->cfg.n_x, nu = h->cfg.n_u;
  const int64_t K = s->K_active;
  cudaStream_t st = h->stream;
  if (!s->dyn_x) {
    CK(h, dalloc(h, &s->dyn_x, (size_t)s->K * nx, false));
    CK(h, dalloc(h, &s->dyn_xn, (size_t)s->K * nx, false));
    CK(h, dalloc(h, &s->dyn_jx, (size_t)s->K * nx * nx, false));
    CK(h, dalloc(h, &s->dyn_ju, (size_t)s->K * nx * nu, false));
  }
  FilterDev fd;
  memset(&fd, 0, sizeof(fd));
  pgb_fill_model(fd, &h->cfg);
  DynDev dd;
  memset(&dd, 0, sizeof(dd));
  dd.K = K; dd.H = s->H; dd.t = t; dd.variant = phi_variant ? 1 : 0;
  dd.A = s->A; dd.x = s->dyn_x; dd.v = t >= 0 ? s->v : nullptr;
  for (int i = 0; i < nu; ++i) dd.u[i] = u[i];
  dd.xn = s->dyn_xn; dd.jx = dfdx ? s->dyn_jx : nullptr; dd.ju = dfdu ? s->dyn_ju : nullptr;
  CK(h, cudaMemcpyAsync(s->dyn_x, x_vec, (size_t)K * nx * 8, cudaMemcpyHostToDevice, st));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device);
  const int64_t want = (K + DY_WPB - 1) / DY_WPB;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sms * 8);
  CK(h, cudaEventRecord(s->ev[0], st));
  switch (nx) {
    case 1: k_dynamics<1><<<grid, DY_WPB * 32, 0, st>>>(fd, dd); break;
    case 2: k_dynamics<2><<<grid, DY_WPB * 32, 0, st>>>(fd, dd); break;
    case 3: k_dynamics<3><<<grid, DY_WPB * 32, 0, st>>>(fd, dd); break;
    default: k_dynamics<4><<<grid, DY_WPB * 32, 0, st>>>(fd, dd); break;
  }
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(s->ev[1], st));
  CK(h, cudaMemcpyAsync(x_next, s->dyn_xn, (size_t)K * nx * 8, cudaMemcpyDeviceToHost, st));
  if (dfdx) CK(h, cudaMemcpyAsync(dfdx, s->dyn_jx, (size_t)K * nx * nx * 8, cudaMemcpyDeviceToHost, st));
  if (dfdu) CK(h, cudaMemcpyAsync(dfdu, s->dyn_ju, (size_t)K * nx * nu * 8, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&s->dyn_ms, s->ev[0], s->ev[1]);
  return PGB_OK;
}

extern "C" int pgb_dynamics_last_ms(pgb_samples* s, double* kernel_ms) {
  if (!s) return set_err(nullptr, PGB_ERR_INVALID, "sample set is NULL");
  if (kernel_ms) *kernel_ms = s->dyn_ms;
  return PGB_OK;
}
