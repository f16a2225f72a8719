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


// ---------------------------------------------------------------------------------------------------------------------
// k_dynamics_bulk — the HBM-bound form of the same step (default for the Hilbert basis with even n_x and n_z = 5 | 3).
// A scenario is 8 n_x n_phi bytes of A_k (20 KB at n_phi = 625) against ~650 FP64 warp instructions: the roof is HBM, and
// k_dynamics' per-lane strided __ldg stream (lane stride 640 B, one dependent L1 round trip per column) reached 14 % of
// it.  Here A_k is staged by the copy engine: ONE cp.async.bulk per scenario (the whole row of A, contiguous) into one of
// the warp's two shared-memory stages, completion on the warp's mbarrier; the copy of the warp's NEXT scenario is in
// flight while the current one is contracted, no thread ever waits on a global load of A.  (One copy per lane into padded
// slots was measured first: 32 copies of 640 B per scenario cost ~70 cycles each in the copy engine, 1.06 ms at K = 10^5.)
// Bank conflicts without padding: lane l owns the contiguous columns [4 l c, 4 (l+1) c) — a lane stride of 160 words at
// c = 5, n_x = 4, i.e. every lane on the same banks.  The runs of a lane are independent fma chains combined by a
// commutative tree, (r0 + r1) + (r2 + r3), so lane l walks them in the order s, s^1, s^2, s^3 with s = l & 3 (pairs stay
// adjacent) and reads the two 16-byte halves of a column in the order h, h^1 with h = (l >> 2) & 1: the eight lanes of a
// quarter warp then touch eight different bank groups (c odd), same operands, same additions, same bits.
// Arithmetic (bit-identical to k_dynamics / orc_dynamics): left-to-right products with the differentiated factor in
// place — the leading factors (dimensions 0 .. n_z-2) are shared between the columns of a row of the last dimension and
// recomputed only when the mixed-radix index carries (6 multiplications + 6 n_x FMAs per column instead of 30 + 6 n_x);
// the 128-run tree: runs 4l .. 4l+3 in-thread, then the xor butterfly with the (1 + n_z) n_x sums halved between the two
// partners at every level while their count is even (27 double shuffles instead of 120 at n_x = 4, n_z = 5).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// number of sums a lane still holds after the halving levels of tree_split, and the levels that halve
__host__ __device__ constexpr int split_final(int n) { int l = 0; while (l < 5 && n % 2 == 0) { n /= 2; ++l; } return n; }
__host__ __device__ constexpr int split_levels(int n) { int l = 0; while (l < 5 && n % 2 == 0) { n /= 2; ++l; } return l; }

// s = s + shfl_xor(s, 1), 2, 4, 8, 16 for N sums at once; while N is even the two partners of a level each keep one half
// (the lane with the level's bit set keeps the upper half): same additions, same operands, half the traffic per level.
// Afterwards v[0 .. split_final(N)-1] are the complete sums number base + i (base = split_base(lane)).
template <int N, int LVL>
__device__ __forceinline__ void tree_split(double* v, int lane) {
  if constexpr (LVL < 5) {
    constexpr int o = 1 << LVL;
    if constexpr (N % 2 == 0) {
      constexpr int Hh = N / 2;
      const bool up = (lane >> LVL) & 1;
#pragma unroll
      for (int i = 0; i < Hh; ++i) {
        const double keep = up ? v[i + Hh] : v[i];
        const double send = up ? v[i] : v[i + Hh];
        v[i] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, o));
      }
      tree_split<Hh, LVL + 1>(v, lane);
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = __dadd_rn(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
      tree_split<N, LVL + 1>(v, lane);
    }
  }
}
template <int N>
__device__ __forceinline__ int split_base(int lane) {
  int base = 0, n = N;
#pragma unroll
  for (int l = 0; l < 5; ++l)
    if (n % 2 == 0) {
      n /= 2;
      if ((lane >> l) & 1) base += n;
    } else {
      break;
    }
  return base;
}

__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// CT = 5: the aligned shapes — c = 5 columns per run, n_phi a multiple of 5 and a run = one row of the last dimension
// (TAIL = 1: its count is 5) or of the last but one when the last has a single function (TAIL = 2; BASELINE.json
// configs[4], n_phi_dims = [1 5 5 5 5]): the factors of the leading n_z - TAIL dimensions are computed once per run and
// the five columns are unrolled; CT = 0 (TAIL = 1): any shape, run-time c, mixed-radix increment per column.
template <int NX, int NZ, bool JAC, int WPB, int CT, int TAIL>
__global__ void __launch_bounds__(WPB * 32, 1) k_dynamics_bulk(const __grid_constant__ FilterDev fd,
                                                               const __grid_constant__ DynDev dd, int stage_b) {
  static_assert(NX % 2 == 0 && NZ >= 2, "even n_x (16-byte columns), at least two augmented dimensions");
  extern __shared__ __align__(128) unsigned char dsm[];
  constexpr int NQ = JAC ? 1 + NZ : 1;
  constexpr int VC = NQ * NX;
  constexpr int NF = split_final(VC);
  constexpr int WRITERS = 1 << split_levels(VC);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_phi = fd.n_phi, nu = fd.n_u;
  const int c = CT ? CT : (n_phi + 127) / 128;
  const uint32_t sbase = smem_u32(dsm);
  const uint32_t bar = sbase + 8 * warp;                                            // [WPB] mbarriers
  const uint32_t tk = sbase + 128 + (uint32_t)warp * 2 * NZ * MAXC * 8;             // scaled sines of this scenario
  const uint32_t dk = tk + NZ * MAXC * 8;                                           // and their derivatives
  const uint32_t stage = sbase + 128 + (((uint32_t)WPB * 2 * NZ * MAXC * 8 + 127u) & ~127u) + (uint32_t)warp * stage_b;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // the lane's columns are the contiguous range [4 l c, 4 (l+1) c), walked run by run in the order s, s^1, s^2, s^3;
  // pk0[j]: mixed-radix digits (4 bits per dimension) of the first column of the j-th run walked
  const int i0 = 4 * lane * c;
  const int rot = lane & 3, hsel = (NX == 4) ? (lane >> 2) & 1 : 0;
  int cnt[NZ];
#pragma unroll
  for (int kk = 0; kk < NZ; ++kk) cnt[kk] = fd.hg_count[kk];
  unsigned pk0[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int rem = i0 + (j ^ rot) * c;
    if (rem >= n_phi) rem = 0;
    unsigned pk = 0;
#pragma unroll
    for (int kk = NZ - 1; kk >= 0; --kk) {
      pk |= (unsigned)(rem % cnt[kk]) << (4 * kk);
      rem /= cnt[kk];
    }
    pk0[j] = pk;
  }
  // the (dimension, frequency) pair whose sine this lane evaluates (ns <= 32 is checked by the launcher)
  int ns = 0;
#pragma unroll
  for (int kk = 0; kk < NZ; ++kk) ns += cnt[kk];
  int sk = 0, sj = lane;
  if (lane < ns) {
    while (sj >= fd.hg_count[sk]) sj -= fd.hg_count[sk++];
  }
  const double sL = fd.hg_L[sk], sw = fd.hg_pij2l[sk][lane < ns ? sj : 0], slsi = fd.hg_lsi[sk];
  const int base = split_base<VC>(lane);
  const int64_t kstride = (int64_t)gridDim.x * WPB;
  const uint32_t row_b = (uint32_t)NX * n_phi * 8;
  const uint32_t my = stage + (uint32_t)i0 * NX * 8;
  uint32_t phase = 0;
  for (int64_t k = (int64_t)blockIdx.x * WPB + warp; k < dd.K; k += kstride, phase ^= 1u) {
    // the whole row A_k by one bulk copy; the previous scenario's reads of the stage ended with __syncwarp
    if (lane == 0) {
      mbar_expect_tx(bar, row_b);
      bulk_g2s(stage, reinterpret_cast<const unsigned char*>(dd.A) + (size_t)k * row_b, row_b, bar);
    }
    double x[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) x[d] = dd.x[k * NX + d];
    // the noise this lane adds when it writes x_next (loaded now, used after the tree)
    double nv[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const int vi = base + i;
      nv[i] = (dd.v && lane < WRITERS && vi < NX) ? dd.v[(k * dd.H + dd.t) * NX + vi] : 0.0;
    }
    if (lane < ns) {
      double zk = dd.u[sk < nu ? sk : 0];
#pragma unroll
      for (int d = 0; d < NX; ++d) zk = (sk - nu == d) ? x[d] : zk;
      const double zl = __dadd_rn(zk, sL);
      double arg;
      if (dd.variant) {  // phi_opt: (pi*j*(z + L)) / (2L)
        const double pij = __dmul_rn(3.141592653589793, (double)(sj + 1));
        arg = __ddiv_rn(__dmul_rn(pij, zl), __dmul_rn(2.0, sL));
      } else {
        arg = __dmul_rn(sw, zl);
      }
      double sn, cs;
      pgb_sincos(arg, &sn, &cs);
      sts_f64(tk + (sk * MAXC + sj) * 8, __dmul_rn(slsi, sn));
      sts_f64(dk + (sk * MAXC + sj) * 8, __dmul_rn(slsi, __dmul_rn(cs, sw)));
    }
    __syncwarp();
    constexpr int NH = NZ - TAIL;  // leading dimensions, folded into head[] once per run
    double tl[CT ? CT : 1], dl[CT ? CT : 1];  // aligned shapes: the factors of the dimension that varies inside a run
    double tz = 1.0, dz = 0.0;                // TAIL = 2: the single function of the last dimension
    if constexpr (CT > 0) {
#pragma unroll
      for (int m = 0; m < CT; ++m) {
        tl[m] = lds_f64(tk + (NH * MAXC + m) * 8);
        dl[m] = JAC ? lds_f64(dk + (NH * MAXC + m) * 8) : 0.0;
      }
      if constexpr (TAIL == 2) {
        tz = lds_f64(tk + (NZ - 1) * MAXC * 8);
        dz = JAC ? lds_f64(dk + (NZ - 1) * MAXC * 8) : 0.0;
      }
    }
    mbar_wait(bar, phase);
    double head[JAC ? 1 + NH : 1];  // products over the leading dimensions: [0] plain, [1 + j] differentiated in dimension j
    double acc[VC], pa[VC], h0[VC];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = j ^ rot;  // the run walked now
      int dig[NZ];
#pragma unroll
      for (int kk = 0; kk < NZ; ++kk) dig[kk] = (int)((pk0[j] >> (4 * kk)) & 15u);
#pragma unroll
      for (int i = 0; i < VC; ++i) acc[i] = 0.0;
      auto heads = [&]() {
        double tv[NH], dv[NH], pre[NH];
#pragma unroll
        for (int kk = 0; kk < NH; ++kk) {
          tv[kk] = lds_f64(tk + (kk * MAXC + dig[kk]) * 8);
          dv[kk] = JAC ? lds_f64(dk + (kk * MAXC + dig[kk]) * 8) : 0.0;
        }
        pre[0] = tv[0];
#pragma unroll
        for (int kk = 1; kk < NH; ++kk) pre[kk] = __dmul_rn(pre[kk - 1], tv[kk]);
        head[0] = pre[NH - 1];
        if constexpr (JAC) {
#pragma unroll
          for (int jj = 0; jj < NH; ++jj) {
            double p = (jj == 0) ? dv[0] : __dmul_rn(pre[jj - 1], dv[jj]);
#pragma unroll
            for (int kk = jj + 1; kk < NH; ++kk) p = __dmul_rn(p, tv[kk]);
            head[1 + jj] = p;
          }
        }
      };
      // one column: a = A_k[:, col] (lanes with hsel read the upper half first: slots d and d^2 are exchanged, undone
      // once after the tree), p_q = head_q * (factors of the trailing dimensions, left to right), acc_q += a * p_q
      auto column = [&](int col, double t0_, double d0_, double t1_, double d1_) {
        double a[NX];
        if constexpr (NX == 4) {
          const double2 v0 = lds_v2f64(my + (uint32_t)col * 32 + 16 * hsel);
          const double2 v1 = lds_v2f64(my + (uint32_t)col * 32 + 16 * (hsel ^ 1));
          a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
        } else {
#pragma unroll
          for (int d = 0; d < NX; d += 2) {
            const double2 a2 = lds_v2f64(my + ((uint32_t)col * NX + d) * 8);
            a[d] = a2.x;
            a[d + 1] = a2.y;
          }
        }
        const double h0t = __dmul_rn(head[0], t0_);  // shared by phi and the derivative in the last dimension (TAIL = 2)
        {
          const double p = TAIL == 2 ? __dmul_rn(h0t, t1_) : h0t;
#pragma unroll
          for (int d = 0; d < NX; ++d) acc[d] = __fma_rn(a[d], p, acc[d]);
        }
        if constexpr (JAC) {
#pragma unroll
          for (int jj = 0; jj < NH; ++jj) {
            double p = __dmul_rn(head[1 + jj], t0_);
            if constexpr (TAIL == 2) p = __dmul_rn(p, t1_);
#pragma unroll
            for (int d = 0; d < NX; ++d) acc[(1 + jj) * NX + d] = __fma_rn(a[d], p, acc[(1 + jj) * NX + d]);
          }
          {
            double p = __dmul_rn(head[0], d0_);
            if constexpr (TAIL == 2) p = __dmul_rn(p, t1_);
#pragma unroll
            for (int d = 0; d < NX; ++d) acc[(1 + NH) * NX + d] = __fma_rn(a[d], p, acc[(1 + NH) * NX + d]);
          }
          if constexpr (TAIL == 2) {
            const double p = __dmul_rn(h0t, d1_);
#pragma unroll
            for (int d = 0; d < NX; ++d) acc[NZ * NX + d] = __fma_rn(a[d], p, acc[NZ * NX + d]);
          }
        }
      };
      if constexpr (CT > 0) {
        if (i0 + rr * CT < n_phi) {  // n_phi is a multiple of CT: a run is complete or empty
          heads();
#pragma unroll
          for (int m = 0; m < CT; ++m) column(rr * CT + m, tl[m], dl[m], tz, dz);
        }
      } else {
        bool need_head = true;
#pragma unroll 1
        for (int m = 0; m < c; ++m) {
          const int col = rr * c + m;
          if (i0 + col >= n_phi) break;
          if (need_head) heads();
          column(col, lds_f64(tk + ((NZ - 1) * MAXC + dig[NZ - 1]) * 8),
                 JAC ? lds_f64(dk + ((NZ - 1) * MAXC + dig[NZ - 1]) * 8) : 0.0, 1.0, 0.0);
          // next column: mixed-radix increment, last dimension fastest
          bool carry = true;
          need_head = false;
#pragma unroll
          for (int kk = NZ - 1; kk >= 0; --kk)
            if (carry) {
              if (++dig[kk] == cnt[kk]) {
                dig[kk] = 0;
                if (kk == NZ - 1) need_head = true;
              } else {
                carry = false;
              }
            }
        }
      }
      // (r0 + r1) + (r2 + r3): the lane meets the pairs in either order and either member first — the additions commute
      if (j == 0 || j == 2) {
#pragma unroll
        for (int i = 0; i < VC; ++i) pa[i] = acc[i];
      } else if (j == 1) {
#pragma unroll
        for (int i = 0; i < VC; ++i) h0[i] = __dadd_rn(pa[i], acc[i]);
      } else {
#pragma unroll
        for (int i = 0; i < VC; ++i) acc[i] = __dadd_rn(h0[i], __dadd_rn(pa[i], acc[i]));
      }
    }
    __syncwarp();  // every lane has read its part of the stage and tk / dk: the next iteration overwrites them
    if constexpr (NX == 4) {  // undo the exchange of the two halves
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const double lo = acc[q * NX + d], hi = acc[q * NX + d + 2];
          acc[q * NX + d] = hsel ? hi : lo;
          acc[q * NX + d + 2] = hsel ? lo : hi;
        }
      }
    }
    tree_split<VC, 0>(acc, lane);
    if (lane < WRITERS) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int vi = base + i, q = vi / NX, d = vi % NX, mm = q - 1;
        if (q == 0) {
          dd.xn[k * NX + d] = dd.v ? __dadd_rn(acc[i], nv[i]) : acc[i];
        } else if (mm < nu) {
          if (dd.ju) dd.ju[k * NX * nu + d + NX * mm] = acc[i];
        } else {
          if (dd.jx) dd.jx[k * NX * NX + d + NX * (mm - nu)] = acc[i];
        }
      }
    }
  }
}

static int g_dynamics_force = -1;  // pgb_test_dynamics_kernel: -1 automatic, 0 k_dynamics, 1 / 2 k_dynamics_bulk with fewer warps

constexpr int DB_WARPS_JAC = 10, DB_WARPS_STATE = 10;

template <int NX, int NZ, bool JAC, int WPB, int CT, int TAIL>
static cudaError_t launch_bulk_one(const FilterDev& fd, const DynDev& dd, int sms, cudaStream_t st) {
  const int stage_b = (NX * fd.n_phi * 8 + 127) & ~127;
  const size_t smem = 128 + (((size_t)WPB * 2 * NZ * MAXC * 8 + 127) & ~(size_t)127) + (size_t)WPB * stage_b;
  const int64_t want = (dd.K + WPB - 1) / WPB;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sms);
  cudaError_t e = cudaFuncSetAttribute(k_dynamics_bulk<NX, NZ, JAC, WPB, CT, TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_dynamics_bulk<NX, NZ, JAC, WPB, CT, TAIL><<<grid, WPB * 32, smem, st>>>(fd, dd, stage_b);
  return cudaGetLastError();
}

// warps per CTA: as many one-row stages as shared memory holds, at most WMAX (10 rows of 20 KB at n_phi = 625; 204 registers per thread)
template <int NX, int NZ, bool JAC, int CT, int TAIL>
static cudaError_t launch_bulk_w(const FilterDev& fd, const DynDev& dd, int sms, cudaStream_t st, int warps) {
  constexpr int WMAX = JAC ? DB_WARPS_JAC : DB_WARPS_STATE;
  if (warps >= WMAX) return launch_bulk_one<NX, NZ, JAC, WMAX, CT, TAIL>(fd, dd, sms, st);
  if (warps >= 8) return launch_bulk_one<NX, NZ, JAC, 8, CT, TAIL>(fd, dd, sms, st);
  if (warps >= 4) return launch_bulk_one<NX, NZ, JAC, 4, CT, TAIL>(fd, dd, sms, st);
  return launch_bulk_one<NX, NZ, JAC, 2, CT, TAIL>(fd, dd, sms, st);
}

template <int NX, int NZ>
static cudaError_t launch_bulk(const FilterDev& fd, const DynDev& dd, int sms, cudaStream_t st, bool jac, int warps) {
  const int c = (fd.n_phi + 127) / 128;
  // a run of c = 5 columns = one row of the dimension that varies fastest (the last, or the last but one when the last
  // has one function)
  int tail = 0;
  if (c == 5 && fd.n_phi % 5 == 0) {
    if (fd.hg_count[NZ - 1] == 5) tail = 1;
    else if (fd.hg_count[NZ - 1] == 1 && fd.hg_count[NZ - 2] == 5) tail = 2;
  }
  if constexpr (NZ == 5) {  // three dimensions of at most MAXC = 8 functions never reach c = 5
    if (jac) {
      if (tail == 1) return launch_bulk_w<NX, NZ, true, 5, 1>(fd, dd, sms, st, warps);
      if (tail == 2) return launch_bulk_w<NX, NZ, true, 5, 2>(fd, dd, sms, st, warps);
    } else {
      if (tail == 1) return launch_bulk_w<NX, NZ, false, 5, 1>(fd, dd, sms, st, warps);
      if (tail == 2) return launch_bulk_w<NX, NZ, false, 5, 2>(fd, dd, sms, st, warps);
    }
  }
  if (jac) return launch_bulk_w<NX, NZ, true, 0, 1>(fd, dd, sms, st, warps);
  return launch_bulk_w<NX, NZ, false, 0, 1>(fd, dd, sms, st, warps);
}

// 0: not eligible (k_dynamics runs), else the number of one-row stages (= warps per CTA) shared memory holds
static int bulk_warps(const FilterDev& fd, int nx, int force, bool jac) {
  if (force == 0 || fd.basis == 0 || (nx != 2 && nx != 4) || (fd.n_z != 3 && fd.n_z != 5)) return 0;
  int ns = 0;
  for (int kk = 0; kk < fd.n_z; ++kk) ns += fd.hg_count[kk];
  if (ns > 32) return 0;
  const size_t stage = (size_t)((nx * fd.n_phi * 8 + 127) & ~127);
  const size_t fixed = 128 + DB_WARPS_STATE * 2 * 8 * MAXC * 8 + 256;
  const size_t cap = 227 * 1024;
  if (fixed + 2 * stage > cap) return 0;
  int w = (int)std::min<size_t>((cap - fixed) / stage, DB_WARPS_STATE);
  if (force == 1) w = std::min(w, 4);
  if (force == 2) w = std::min(w, 2);
  if (force == 3 || (force < 0 && jac)) w = std::min(w, 8);
  return w;
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
  const bool jac = dd.jx || dd.ju;
  const int bw = bulk_warps(fd, nx, g_dynamics_force, jac);
  CK(h, cudaEventRecord(s->ev[0], st));
  if (bw) {  // A_k staged by cp.async.bulk, one stage per warp (Hilbert basis, even n_x, n_z = 3 | 5)
    cudaError_t e = cudaSuccess;
    if (nx == 4 && fd.n_z == 5) e = launch_bulk<4, 5>(fd, dd, sms, st, jac, bw);
    else if (nx == 4) e = launch_bulk<4, 3>(fd, dd, sms, st, jac, bw);
    else if (fd.n_z == 5) e = launch_bulk<2, 5>(fd, dd, sms, st, jac, bw);
    else e = launch_bulk<2, 3>(fd, dd, sms, st, jac, bw);
    CK(h, e);
  } else {
    switch (nx) {
      case 1: k_dynamics<1><<<grid, DY_WPB * 32, 0, st>>>(fd, dd); break;
      case 2: k_dynamics<2><<<grid, DY_WPB * 32, 0, st>>>(fd, dd); break;
      case 3: k_dynamics<3><<<grid, DY_WPB * 32, 0, st>>>(fd, dd); break;
      default: k_dynamics<4><<<grid, DY_WPB * 32, 0, st>>>(fd, dd); break;
    }
    CK(h, cudaGetLastError());
  }
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

extern "C" void pgb_test_dynamics_kernel(int32_t which) { g_dynamics_force = which; }
