// pgb_exact.h — primitives that let a *parallel* scan reproduce, bit for bit, the strictly
// sequential FP64 recurrences of the reference's systematic_resampling
// (Julia/src/particle_Gibbs.jl:23-32):
//
//     q = q + W[n]          (left-to-right rounded prefix sum of the normalised weights)
//     u = u + 1/N           (repeatedly incremented sampling grid)
//
// Idea.  While q stays inside one binade [2^e, 2^(e+1)) every partial sum is
// an integer multiple m of ulp_e, and fl(q + W) = (m + c_{m&1}) * ulp_e where the two
// increments (one per parity of m — parity only matters on exact round-half-even ties) depend
// on W and e alone.  The maps m -> m + c_{m&1} are closed under composition, so a run of
// additions inside one binade is an *associative* scan over 16-byte elements.  Binade
// crossings (about log2(N) of them) are not composable; they are handled by real FP adds in
// a short serial chain.  Which binade an element sees is predicted from an ordinary
// (approximate) FP scan; the prediction is never trusted: every thread re-walks its own four
// elements with true FP adds from its map-derived incoming value and the result is compared
// bitwise with its right neighbour's incoming value ("seam check").  q_in(0) = 0 plus all
// seams equal proves, by induction, that every q_n is the sequential one; any mismatch
// raises a flag and the step is redone by the exact serial walk.
//
// All functions are host+device so tests/host_model.cpp can exercise the same code on the CPU.
#ifndef PGB_EXACT_H
#define PGB_EXACT_H

#include "pgb_math.h"

#define PGB_TWO52 (1ll << 52)
#define PGB_TWO53 (1ll << 53)

// Segment key of a finite q >= 0: biased exponent, with subnormals/zero folded into the
// first normal binade (same ulp 2^-1074).  Mantissa integer m with q = m * 2^(key-1075).
PGB_HD int pgb_segkey(double q) {
  int be = (int)((pgb_d2u(q) >> 52) & 0x7ff);
  return be < 1 ? 1 : be;
}
PGB_HD int64_t pgb_mant(double q) {
  uint64_t u = pgb_d2u(q);
  uint64_t frac = u & 0x000fffffffffffffull;
  return (int64_t)(((u >> 52) & 0x7ff) ? (frac | (1ull << 52)) : frac);
}
// inverse of (segkey, mant); also correct for m == 2^53 (lands exactly on the next binade)
PGB_HD double pgb_from_mant(int key, int64_t m) {
  return pgb_u2d(((uint64_t)(key - 1) << 52) + (uint64_t)m);
}

typedef struct {
  int64_t c0, c1;  // increment applied when m is even / odd
} pgb_map;

PGB_HD pgb_map pgb_map_identity(void) {
  pgb_map f;
  f.c0 = 0;
  f.c1 = 0;
  return f;
}
PGB_HD int64_t pgb_map_apply(pgb_map f, int64_t m) { return m + ((m & 1) ? f.c1 : f.c0); }
// apply L first, then R
PGB_HD pgb_map pgb_map_compose(pgb_map L, pgb_map R) {
  pgb_map o;
  o.c0 = L.c0 + ((L.c0 & 1) ? R.c1 : R.c0);
  o.c1 = L.c1 + (((1 + L.c1) & 1) ? R.c1 : R.c0);
  return o;
}
// Map of "q -> fl(q + W)" for q in binade `key` (result assumed to stay in that binade).
// Returns 0 and leaves *f untouched if W lives on a coarser grid than q (the add certainly
// leaves the binade, so the caller mispredicted).  W must be finite and >= 0.
PGB_HD int pgb_make_map(int key, double W, pgb_map* f) {
  if (W == 0.0) {
    *f = pgb_map_identity();
    return 1;
  }
  int keyw = pgb_segkey(W);
  int64_t mw = pgb_mant(W);
  int sh = key - keyw;
  if (sh < 0) return 0;
  pgb_map o;
  if (sh == 0) {
    o.c0 = o.c1 = mw;
  } else if (sh >= 54) {
    o.c0 = o.c1 = 0;
  } else {
    int64_t k = mw >> sh;
    int64_t rnd = (mw >> (sh - 1)) & 1;
    int64_t sticky = (mw & ((1ll << (sh - 1)) - 1)) != 0;
    if (!rnd) {
      o.c0 = o.c1 = k;
    } else if (sticky) {
      o.c0 = o.c1 = k + 1;
    } else {  // exact tie: round half to even => depends on parity of m + k
      o.c0 = k + (k & 1);
      o.c1 = k + ((k + 1) & 1);
    }
  }
  *f = o;
  return 1;
}

// ---------------------------------------------------------------------------------------
// Scan element for the hierarchical (thread -> warp -> tile -> grid) segmented composition.
//   cnt : number of "complex" threads (threads whose 4 elements are not one clean same-binade
//         run; they break composability and are resolved with real FP adds in the chain)
//   f   : composition of the simple threads after the last complex one (or all of them)
//   key : binade that tail run lives in; 0 = wildcard (only zero weights so far), -1 = poison
// ---------------------------------------------------------------------------------------
typedef struct {
  pgb_map f;
  int32_t cnt;
  int32_t key;
} pgb_agg;

PGB_HD pgb_agg pgb_agg_identity(void) {
  pgb_agg a;
  a.f = pgb_map_identity();
  a.cnt = 0;
  a.key = 0;
  return a;
}
PGB_HD pgb_agg pgb_agg_complex(void) {
  pgb_agg a;
  a.f = pgb_map_identity();
  a.cnt = 1;
  a.key = 0;
  return a;
}
PGB_HD int32_t pgb_key_merge(int32_t a, int32_t b) {
  if (a == 0) return b;
  if (b == 0) return a;
  return (a == b) ? a : -1;
}
// associative: L covers the elements to the left of R
PGB_HD pgb_agg pgb_agg_combine(pgb_agg L, pgb_agg R) {
  pgb_agg o;
  if (R.cnt > 0) {
    o.f = R.f;
    o.key = R.key;
    o.cnt = L.cnt + R.cnt;
  } else {
    o.f = pgb_map_compose(L.f, R.f);
    o.key = pgb_key_merge(L.key, R.key);
    o.cnt = L.cnt;
  }
  return o;
}

// A predicted prefix this close (relative 2^-32) to a power of two cannot be trusted to tell
// which side of the binade boundary the *sequential* sum is on.
PGB_HD int pgb_near_pow2(double q) {
  uint64_t frac = pgb_d2u(q) & 0x000fffffffffffffull;
  return frac < (1ull << 20) || frac > ((1ull << 52) - (1ull << 20));
}
// lower of the two binades a near-boundary prediction could mean
PGB_HD int pgb_lower_candidate_key(double q) {
  int key = pgb_segkey(q);
  uint64_t frac = pgb_d2u(q) & 0x000fffffffffffffull;
  return (frac < (1ull << 20) && key > 1) ? key - 1 : key;
}

// Classify one thread's 4 consecutive elements.  qt[0] is the approximate prefix before the
// first element, qt[1..4] the approximate prefixes after each element (qt[4] must be the
// *neighbour's* qt[0] so that adjacent threads agree on the seam).  Returns a simple
// aggregate, or the complex marker when any element changes binade / is ambiguous.
// Elements whose addition is a no-op in every candidate binade (zero or negligible weight)
// constrain nothing and are skipped.
PGB_HD_HELPER pgb_agg pgb_classify4(const double W[4], const double qt[5]) {
  pgb_agg a = pgb_agg_identity();
  for (int i = 0; i < 4; ++i) {
    if (W[i] == 0.0) continue;  // identity in every binade
    pgb_map f;
    if (pgb_near_pow2(qt[i]) || pgb_near_pow2(qt[i + 1])) {
      int klo = pgb_lower_candidate_key(pgb_near_pow2(qt[i]) ? qt[i] : qt[i + 1]);
      if (pgb_make_map(klo, W[i], &f) && f.c0 == 0 && f.c1 == 0) continue;  // negligible either way
      return pgb_agg_complex();
    }
    int kin = pgb_segkey(qt[i]);
    int kout = pgb_segkey(qt[i + 1]);
    if (kin != kout || !pgb_make_map(kin, W[i], &f)) return pgb_agg_complex();
    if (f.c0 == 0 && f.c1 == 0) continue;  // negligible here, hence in every higher binade too
    if (a.key != 0 && a.key != kin) return pgb_agg_complex();
    a.key = kin;
    a.f = pgb_map_compose(a.f, f);
  }
  return a;
}

// Value entering a thread given its global exclusive aggregate and the chain seed selected by
// agg.cnt.  (A poisoned/mismatching key just produces a value the seam check rejects.)
PGB_HD double pgb_qin_from(pgb_agg ex, double seed) {
  return pgb_from_mant(pgb_segkey(seed), pgb_map_apply(ex.f, pgb_mant(seed)));
}

// ---------------------------------------------------------------------------------------
// The sampling grid u_1 = fl(fl(1/M)*rnd), u_{i+1} = fl(u_i + fl(1/M)) in closed form:
// inside one binade the increment is a constant integer number of ulps (after at most one
// tie-parity step), so u_i = (m0 + (i - i0) * c) * ulp.  <= ~70 rows for any M < 2^31.
// ---------------------------------------------------------------------------------------
#define PGB_UT_MAXROWS 96
typedef struct {
  int64_t i0[PGB_UT_MAXROWS];   // 1-based draw index of the row's first grid point
  int64_t m0[PGB_UT_MAXROWS];   // its mantissa integer
  int64_t c[PGB_UT_MAXROWS];    // ulps added per draw inside the row
  double u0[PGB_UT_MAXROWS];    // its value (= from_mant(key, m0))
  double invc[PGB_UT_MAXROWS];  // 1/c (quotient estimate, corrected exactly with integer arithmetic)
  int32_t key[PGB_UT_MAXROWS];
  int32_t nrows;
  int64_t M;                    // number of draws
  int32_t overflow;             // table did not fit (never for sane inputs)
  int32_t pad_;
} pgb_utable;

PGB_HD void pgb_utable_build(pgb_utable* T, int64_t M, double rnd) {
  double d = PGB_DIV(1.0, (double)M);
  double u = PGB_MUL(d, rnd);
  int64_t i = 1;
  int r = 0;
  T->M = M;
  T->overflow = 0;
  T->pad_ = 0;
  while (i <= M) {
    if (r >= PGB_UT_MAXROWS) {
      T->overflow = 1;
      break;
    }
    int key = pgb_segkey(u);
    int64_t m = pgb_mant(u);
    pgb_map f;
    int64_t len = 1;  // grid points in this row
    int64_t c = 0;
    if (pgb_make_map(key, d, &f)) {
      if (f.c0 != f.c1 && (m & 1)) {
        len = 1;  // odd start on a tie: one irregular step, afterwards m stays even
      } else {
        c = f.c0;
        if (c > 0) {
          len = (PGB_TWO53 - 1 - m) / c + 1;  // points with m + j*c < 2^53
        } else {
          len = M - i + 1;  // cannot happen for d = 1/M (d >> ulp(u)); keep it finite
        }
      }
    }
    if (len > M - i + 1) len = M - i + 1;
    T->i0[r] = i;
    T->m0[r] = m;
    T->c[r] = c;
    T->u0[r] = u;
    T->invc[r] = (c > 0) ? PGB_DIV(1.0, (double)c) : 0.0;
    T->key[r] = key;
    ++r;
    double last = pgb_from_mant(key, m + (len - 1) * c);
    u = PGB_ADD(last, d);
    i += len;
  }
  T->nrows = r;
}
// u_i, 1 <= i <= M
PGB_HD double pgb_utable_at(const pgb_utable* T, int64_t i) {
  int lo = 0, hi = T->nrows - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (T->i0[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return pgb_from_mant(T->key[lo], T->m0[lo] + (i - T->i0[lo]) * T->c[lo]);
}
// floor((mq - m0) / c) without a 64-bit integer division: double estimate, exact integer correction
PGB_HD int64_t pgb_floor_div_est(int64_t num, int64_t c, double invc) {
  int64_t j = (int64_t)((double)num * invc);
  if (j < 0) j = 0;
  while ((j + 1) * c <= num) ++j;
  while (j > 0 && j * c > num) --j;
  return j;
}
// #{ i in row `lo` of T : u_i <= q } given that row lo is the last row with u0 <= q
PGB_HD int64_t pgb_utable_count_in_row(const pgb_utable* T, int lo, double q) {
  int64_t i0 = T->i0[lo];
  int64_t len = ((lo + 1 < T->nrows) ? T->i0[lo + 1] : T->M + 1) - i0;
  int64_t j;
  if (T->c[lo] == 0) {
    j = 0;
  } else if (pgb_segkey(q) != T->key[lo]) {
    j = len - 1;  // q lies in a higher binade than the whole row
  } else {
    j = pgb_floor_div_est(pgb_mant(q) - T->m0[lo], T->c[lo], T->invc[lo]);
    if (j > len - 1) j = len - 1;
  }
  return i0 + j;
}
// #{ i in [1, M] : u_i <= q }   (q finite >= 0)
PGB_HD_HELPER int64_t pgb_utable_count_le(const pgb_utable* T, double q) {
  if (q != q) return T->M;  // NaN prefix: `while q < u` is false for every draw, all of them stop here
  if (T->nrows == 0 || !(q >= T->u0[0])) return 0;
  int lo = 0, hi = T->nrows - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (T->u0[mid] <= q) lo = mid; else hi = mid - 1;
  }
  return pgb_utable_count_in_row(T, lo, q);
}
// last row whose first grid point is <= q (0 when there is none): binary search, the starting hint of a lane's queries
PGB_HD int pgb_utable_row_of(const pgb_utable* T, double q) {
  int lo = 0, hi = T->nrows - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (T->u0[mid] <= q) lo = mid; else hi = mid - 1;
  }
  return lo;
}
// same, for a non-decreasing sequence of queries: *row is the row found for the previous query
PGB_HD_HELPER int64_t pgb_utable_count_le_hint(const pgb_utable* T, double q, int* row) {
  if (q != q) return T->M;
  if (T->nrows == 0 || !(q >= T->u0[0])) return 0;
  int lo = *row;
  if (lo < 0 || lo >= T->nrows || !(T->u0[lo] <= q)) lo = 0;
  while (lo + 1 < T->nrows && T->u0[lo + 1] <= q) ++lo;
  *row = lo;
  return pgb_utable_count_in_row(T, lo, q);
}

#endif  // PGB_EXACT_H
