// Selected eigenpairs of a symmetric tridiagonal matrix by bisection + inverse iteration, shared by the CUDA
// kernels (stebz.cu) and the host-side check library (host_check.cpp) so the numerics AND the per-cluster
// orchestration can be unit-tested on a CPU against LAPACK's dstebz / dstein.
//
// Replaces the tridiagonal part of pdsyevx('V','I','L', ..., il = 1, iu = n_vec, abstol = 2 safmin, orfac = 0)
// (reference src/solver_scalapack_select.f90:52-60), i.e. pdstebz + pdstein:
//   K8a  one eigenvalue per thread: bisection on the Sturm count of the LDL^T recurrence (dstebz's inner loop),
//        to relative width 2 ulp like dstebz with abstol = 2 safmin;
//   K8b  one cluster per warp: lane 0 factors T - lambda I with partial pivoting and runs the (inherently serial)
//        triangular solves, all lanes do the norms, the modified Gram-Schmidt sweep against the earlier vectors of
//        the cluster and the scaling, following dstein's iteration (random start, rhs scaled to
//        n |T|_1 max(eps, |u_nn|), stop two iterations after the growth test passes, at most five).
// The reference passes orfac = 0 (no reorthogonalisation at all); here vectors whose eigenvalues lie within
// `ortol` of each other ARE reorthogonalised, with ortol adapted to the spectrum so clusters stay small
// (stein_ortol below), which keeps degenerate eigenvalues (e.g. the shipped VCNT400 matrix) orthogonal.
#pragma once
#include <math.h>
#include <stdint.h>

#include "secular.cuh"  // EKB_HD

#ifdef __CUDACC__
#define EKB_UNROLL _Pragma("unroll")
#else
#define EKB_UNROLL
#endif

namespace ekb {

constexpr double TRI_SAFMIN = 2.2250738585072014e-308;  // dlamch('S')
constexpr double TRI_ULP = 2.220446049250313e-16;       // dlamch('P')
constexpr double TRI_EPS = 1.1102230246251565e-16;      // dlamch('E')
constexpr int STEIN_MAXITS = 5, STEIN_EXTRA = 2;        // dstein's MAXITS / EXTRA
constexpr int STEIN_MAX_CLUSTER = 512;                  // larger clusters are cut (like pdstein at process borders)

// Number of eigenvalues <= x (dstebz's count: pivots of T - x I with the pivmin safeguard).  e2[i] = e[i]^2.
EKB_HD long long sturm_count(long long n, const double* d, const double* e2, double x, double pivmin) {
  double q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  long long cnt = q <= 0.0 ? 1 : 0;
  for (long long i = 1; i < n; ++i) {
    q = d[i] - e2[i - 1] / q - x;
    if (fabs(q) < pivmin) q = -pivmin;
    if (q <= 0.0) ++cnt;
  }
  return cnt;
}

// e / q for the interleaved Sturm chains.  The IEEE double division of the toolchain carries a branch to a slow path
// (subnormal / huge operands); with a branch per division the compiler cannot interleave the K chains and K = 7 ran
// 1.8x SLOWER than plain bisection on the B200 (89 ms vs 50 ms at n = 8192).  On the device the quotient is therefore
// formed branch-free: MUFU reciprocal seed, three Newton steps, one residual correction (error < 1 ulp).  The pivmin
// safeguard keeps |q| >= safmin max(1, max e^2), so neither the reciprocal nor the quotient leaves the normal range.
// The host build keeps the IEEE division; the two can only disagree on a count when the shift is within rounding
// error of an eigenvalue.  Measured (profiles/r01_probe_select_inverse.jsonl): 58 ms at n = 8192, 0.28 s at n = 32768
// for all n eigenvalues -- per bit of the result no better than the plain kernel (50 ms / 0.345 s): SASS showed why --
// ptxas had re-serialised the 7 chains (one Newton sequence after the other on shared temporaries).  Fixed after
// the measurement by stage-major source + a minimum-blocks launch bound (SASS now interleaved; not re-timed yet).
EKB_HD double sturm_quot(double e, double q) {
#if defined(__CUDA_ARCH__) && !defined(EKB_STURM_IEEE_DIV)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(q));
  double t = fma(-q, r, 1.0);
  r = fma(r, t, r);
  t = fma(-q, r, 1.0);
  r = fma(r, t, r);
  t = fma(-q, r, 1.0);
  r = fma(r, t, r);
  const double v = e * r;
  const double rem = fma(-v, q, e);
  return fma(rem, r, v);
#else
  return e / q;
#endif
}

// v[k] = e / q[k] for K independent chains, written STAGE BY STAGE (all seeds, then every Newton step for all k, ...):
// ptxas keeps the source order inside the loop body, and with the chains written one after the other it emitted seven
// serial Newton sequences on shared temporaries (SASS of the first 7-chain kernel: no gain over plain bisection).
// Stage-major source gives K independent instructions between two dependent ones in the PTX; ptxas keeps that order
// only when the kernel carries a minimum-blocks launch bound (see bisect_kernel in stebz.cu).  Per chain the
// operations and their order are exactly those of sturm_quot.
template <int K>
EKB_HD void sturm_quot_multi(double e, const double* q, double* v) {
#if defined(__CUDA_ARCH__) && !defined(EKB_STURM_IEEE_DIV)
  double r[K], t[K];
  EKB_UNROLL
  for (int k = 0; k < K; ++k) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[k]) : "d"(q[k]));
  EKB_UNROLL
  for (int s = 0; s < 3; ++s) {
    EKB_UNROLL
    for (int k = 0; k < K; ++k) t[k] = fma(-q[k], r[k], 1.0);
    EKB_UNROLL
    for (int k = 0; k < K; ++k) r[k] = fma(r[k], t[k], r[k]);
  }
  EKB_UNROLL
  for (int k = 0; k < K; ++k) v[k] = e * r[k];
  EKB_UNROLL
  for (int k = 0; k < K; ++k) t[k] = fma(-v[k], q[k], e);
  EKB_UNROLL
  for (int k = 0; k < K; ++k) v[k] = fma(t[k], r[k], v[k]);
#else
  for (int k = 0; k < K; ++k) v[k] = e / q[k];
#endif
}

// K Sturm counts in one sweep: K independent recurrences interleaved so the divide latency of one hides behind the
// others (the single-chain kernel was latency-bound: ncu "wait" stalls 3.8 warps/issue, 16 % issue-active).  On the
// host each count is bit-identical to sturm_count at the same shift.
template <int K>
EKB_HD void sturm_count_multi(long long n, const double* __restrict__ d, const double* __restrict__ e2, const double* x,
                              double pivmin, long long* cnt) {
  double q[K], v[K];
  long long c[K];
  EKB_UNROLL
  for (int k = 0; k < K; ++k) {
    q[k] = d[0] - x[k];
    if (fabs(q[k]) < pivmin) q[k] = -pivmin;
    c[k] = q[k] <= 0.0 ? 1 : 0;
  }
  for (long long i = 1; i < n; ++i) {
    const double di = d[i], ei = e2[i - 1];
    sturm_quot_multi<K>(ei, q, v);
    EKB_UNROLL
    for (int k = 0; k < K; ++k) {
      double t = di - v[k] - x[k];
      if (fabs(t) < pivmin) t = -pivmin;
      q[k] = t;
      if (t <= 0.0) ++c[k];
    }
  }
  EKB_UNROLL
  for (int k = 0; k < K; ++k) cnt[k] = c[k];
}

// The j-th smallest eigenvalue (0-based) inside [gl, gu] (a Gershgorin interval widened like dstebz does), to dstebz's
// tolerance for abstol = 2 safmin.  Multi-section: K interior points per sweep cut the bracket by K + 1 (K = 1 is
// plain bisection); when the points no longer separate in floating point the sweep degenerates to a midpoint step.
template <int K>
EKB_HD double bisect_index_k(long long n, const double* __restrict__ d, const double* __restrict__ e2, long long j, double gl,
                             double gu, double pivmin, int* iters) {
  double lo = gl, hi = gu;
  int it = 0;
  for (; it < 128; ++it) {
    const double mid = 0.5 * (lo + hi);
    const double tol = fmax(2.0 * TRI_SAFMIN, fmax(pivmin, 2.0 * TRI_ULP * fmax(fabs(lo), fabs(hi))));
    if (hi - lo <= tol || mid <= lo || mid >= hi) break;
    double x[K];
    bool distinct = K > 1;
    if (K > 1) {
      const double h = (hi - lo) / (double)(K + 1);
      double prev = lo;
EKB_UNROLL
      for (int k = 0; k < K; ++k) {
        x[k] = lo + (double)(k + 1) * h;
        if (!(x[k] > prev)) distinct = false;
        prev = x[k];
      }
      if (!(prev < hi)) distinct = false;
    }
    if (!distinct) {
      if (sturm_count(n, d, e2, mid, pivmin) >= j + 1) hi = mid;
      else lo = mid;
      continue;
    }
    long long c[K];
    sturm_count_multi<K>(n, d, e2, x, pivmin, c);
    double nlo = x[K - 1], nhi = hi;
    bool found = false;
EKB_UNROLL
    for (int k = 0; k < K; ++k) {
      if (!found && c[k] >= j + 1) {
        found = true;
        nhi = x[k];
        nlo = k > 0 ? x[k - 1] : lo;
      }
    }
    lo = nlo;
    hi = nhi;
  }
  if (iters) *iters = it;
  return 0.5 * (lo + hi);
}

EKB_HD double bisect_index(long long n, const double* d, const double* e2, long long j, double gl, double gu,
                           double pivmin, int* iters) {
  return bisect_index_k<1>(n, d, e2, j, gl, gu, pivmin, iters);
}

// start vector of inverse iteration: counter-hash uniform in (-1, 1), identical on host and device
EKB_HD double stein_start(uint64_t j, uint64_t i) {
  uint64_t x = (j << 32) ^ i ^ 0x5DEECE66DULL;
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  x ^= x >> 31;
  return (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0 + 1.1102230246251565e-16;
}

// LU factorisation with partial pivoting of T - lambda I (the job dlagtf does for dstein).
//   U: a (diagonal), b (first superdiagonal), d2 (second superdiagonal); multipliers c; piv[i] = 1 when rows i and
//   i+1 were interchanged at step i.  All arrays have n entries.
EKB_HD void gt_factor(long long n, const double* __restrict__ d, const double* __restrict__ e, double lambda,
                      double* __restrict__ a, double* __restrict__ b, double* __restrict__ c, double* __restrict__ d2,
                      unsigned char* __restrict__ piv) {
  // The current row (diagonal ai, superdiagonal bi) is carried in registers and the untouched rows are read straight
  // from d and e, four at a time ahead of the dependent chain: the recurrence never waits on a load of something it
  // has just stored (the first version read a[i+1] / b[i+1] back from memory every step).
  constexpr int U = 4;
  double ai = d[0] - lambda;
  double bi = n > 1 ? e[0] : 0.0;
  long long i = 0;
  auto step = [&](long long row, double sub, double an0, double bn0, bool has_next2) {
    // sub = e[row] (element (row+1, row)); an0 / bn0 = original diagonal / superdiagonal of row + 1
    if (fabs(ai) >= fabs(sub)) {
      const double m = ai != 0.0 ? sub / ai : 0.0;
      c[row] = m;
      piv[row] = 0;
      d2[row] = 0.0;
      a[row] = ai;
      b[row] = bi;
      ai = an0 - m * bi;
      bi = bn0;
    } else {
      const double m = ai / sub;
      c[row] = m;
      piv[row] = 1;
      a[row] = sub;
      b[row] = an0;
      const double na = bi - m * an0;
      d2[row] = has_next2 ? bn0 : 0.0;
      bi = has_next2 ? -m * bn0 : bn0;
      ai = na;
    }
  };
  for (; i + U + 1 < n; i += U) {  // rows i .. i+U-1, all of which have a row + 2
    double sub[U], an0[U], bn0[U];
    EKB_UNROLL
    for (int u = 0; u < U; ++u) {
      sub[u] = e[i + u];
      an0[u] = d[i + u + 1] - lambda;
      bn0[u] = e[i + u + 1];
    }
    EKB_UNROLL
    for (int u = 0; u < U; ++u) step(i + u, sub[u], an0[u], bn0[u], true);
  }
  for (; i + 1 < n; ++i) {
    const bool has2 = i + 2 < n;
    step(i, e[i], d[i + 1] - lambda, has2 ? e[i + 1] : 0.0, has2);
  }
  a[n - 1] = ai;
  b[n - 1] = 0.0;
  c[n - 1] = 0.0;
  d2[n - 1] = 0.0;
  piv[n - 1] = 0;
}

// Reciprocal of a pivot of U, with pivots smaller than `pert` replaced by +-pert (dlagts with job = -1).  Applied to
// the whole diagonal once per factorisation (by all lanes of the team) so the back substitution multiplies instead of
// dividing on its critical path.
EKB_HD double gt_pivot_recip(double p, double pert) {
  if (fabs(p) < pert) p = p < 0.0 ? -pert : pert;
  return 1.0 / p;
}

// x <- (T - lambda I)^-1 x with the factors of gt_factor and ainv[i] = gt_pivot_recip(a[i]).  Both sweeps are serial
// recurrences (one FMA per row forward, two FMAs and a multiply backward); the operands of 8 rows are loaded before
// the dependent chain touches them, so the loads overlap it instead of adding an L2 round trip per row (the first
// version stalled 6.7 warps/issue on the scoreboard).  Components are clamped at 1e290 instead of overflowing.
EKB_HD void gt_solve(long long n, const double* __restrict__ ainv, const double* __restrict__ b,
                     const double* __restrict__ c, const double* __restrict__ d2, const unsigned char* __restrict__ piv,
                     double* __restrict__ x) {
  constexpr int U = 8;
  double cur = x[0];
  long long i = 0;
  for (; i + U < n; i += U) {
    double cc[U], xn[U];
    unsigned char pp[U];
EKB_UNROLL
    for (int u = 0; u < U; ++u) {
      cc[u] = c[i + u];
      pp[u] = piv[i + u];
      xn[u] = x[i + 1 + u];
    }
EKB_UNROLL
    for (int u = 0; u < U; ++u) {
      const double nxt = xn[u];
      const double out = pp[u] ? nxt : cur;
      const double carry = pp[u] ? cur - cc[u] * nxt : nxt - cc[u] * cur;
      x[i + u] = out;
      cur = carry;
    }
  }
  for (; i + 1 < n; ++i) {
    const double nxt = x[i + 1];
    const double out = piv[i] ? nxt : cur;
    const double carry = piv[i] ? cur - c[i] * nxt : nxt - c[i] * cur;
    x[i] = out;
    cur = carry;
  }
  x[n - 1] = cur;
  double x1 = 0.0, x2 = 0.0;  // x[i+1], x[i+2]
  i = n - 1;
  for (; i - (U - 1) >= 0; i -= U) {
    double bb[U], dd[U], aa[U], xx[U];
EKB_UNROLL
    for (int u = 0; u < U; ++u) {
      bb[u] = b[i - u];
      dd[u] = d2[i - u];
      aa[u] = ainv[i - u];
      xx[u] = x[i - u];
    }
EKB_UNROLL
    for (int u = 0; u < U; ++u) {
      double v = (xx[u] - bb[u] * x1 - dd[u] * x2) * aa[u];
      if (fabs(v) > 1e290) v = v < 0.0 ? -1e290 : 1e290;
      x[i - u] = v;
      x2 = x1;
      x1 = v;
    }
  }
  for (; i >= 0; --i) {
    double v = (x[i] - b[i] * x1 - d2[i] * x2) * ainv[i];
    if (fabs(v) > 1e290) v = v < 0.0 ? -1e290 : 1e290;
    x[i] = v;
    x2 = x1;
    x1 = v;
  }
}

// Reorthogonalisation threshold: eigenvalues closer than this belong to one cluster.  dstein uses 1e-3 |T|_1, which
// would chain a dense spectrum (n = 65536: spacing 1.5e-5 |T|) into one serial cluster; half the mean spacing of the
// requested eigenvalues keeps clusters at a few vectors while every non-clustered pair keeps a gap large enough for
// inner products of order eps |T| / gap << 1e-12 n.  Never below 64 eps |T|_1 (numerically coincident values).
inline double stein_ortol(long long k, const double* w, double onenrm) {
  double t = 1e-3 * onenrm;
  if (k > 1) {
    const double mean = (w[k - 1] - w[0]) / (double)(k - 1);
    if (0.5 * mean < t) t = 0.5 * mean;
  }
  const double floor_ = 64.0 * TRI_EPS * onenrm;
  return t > floor_ ? t : floor_;
}

// cluster start indices for w[lo..hi) (ascending); returns their number, starts[] gets count + 1 entries
inline long long stein_clusters(long long lo, long long hi, const double* w, double ortol, long long* starts) {
  long long nc = 0;
  long long cur = lo;
  if (hi <= lo) { starts[0] = lo; return 0; }
  starts[nc++] = lo;
  for (long long j = lo + 1; j < hi; ++j) {
    if (w[j] - w[j - 1] > ortol || j - cur >= STEIN_MAX_CLUSTER) {
      starts[nc++] = j;
      cur = j;
    }
  }
  starts[nc] = hi;
  return nc;
}

// Sharded solves: every rank cuts the clusters of ALL requested eigenvalues (same list everywhere) and processes the
// ones that intersect its column slab [col_lo, col_hi) completely, so a cluster that straddles a slab border is
// orthogonalised as a whole on both ranks (bit-identical: the per-cluster arithmetic has a fixed order) instead of
// being cut at the border the way pdstein cuts at process borders.  Columns outside the slab are scratch.
inline void stein_cluster_range(long long nc, const long long* starts, long long col_lo, long long col_hi,
                                long long* first, long long* count) {
  long long f = 0;
  while (f < nc && starts[f + 1] <= col_lo) ++f;
  long long l = f;
  while (l < nc && starts[l] < col_hi) ++l;
  *first = f;
  *count = col_hi > col_lo ? l - f : 0;
}

// Team abstraction: on the device a warp (lane 0 runs the serial recurrences, all lanes the vector work), on the
// host a single "lane".
struct HostTeam {
  EKB_HD int lane() const { return 0; }
  EKB_HD int width() const { return 1; }
  EKB_HD void sync() const {}
  EKB_HD double sum(double v) const { return v; }
  EKB_HD double max(double v) const { return v; }
  EKB_HD double bcast0(double v) const { return v; }
};

// Eigenvectors of one cluster w[j0..j1) into Z(:, j0 - zoff .. j1 - zoff).  ws: 4 n doubles + n bytes of team-private
// scratch.  *fail counts vectors whose growth test never passed (dstein's IFAIL).
template <class Team>
EKB_HD void stein_cluster(const Team& tm, long long n, const double* d, const double* e, const double* w, long long j0,
                          long long j1, double* Z, long long ldz, long long zoff, double* ws, double onenrm, int* fail) {
  double* a = ws;
  double* b = ws + n;
  double* c = ws + 2 * n;
  double* d2 = ws + 3 * n;
  unsigned char* piv = reinterpret_cast<unsigned char*>(ws + 4 * n);
  const int L = tm.lane(), W = tm.width();
  const double eps = TRI_EPS;
  if (!(onenrm > TRI_SAFMIN)) onenrm = TRI_SAFMIN;
  const double pert_floor = eps * onenrm > TRI_SAFMIN ? eps * onenrm : TRI_SAFMIN;
  const double dtpcrt = sqrt(0.1 / (double)n);
  double xjm = 0.0;
  for (long long j = j0; j < j1; ++j) {
    double* x = Z + (j - zoff) * ldz;
    // coincident eigenvalues get dstein's tiny separation so the factorisations differ
    double xj = w[j];
    if (j > j0) {
      const double pertol = 10.0 * fabs(eps * xj);
      if (xj - xjm < pertol) xj = xjm + pertol;
    }
    xjm = xj;
    if (n == 1) {
      if (L == 0) x[0] = 1.0;
      tm.sync();
      continue;
    }
    if (L == 0) gt_factor(n, d, e, xj, a, b, c, d2, piv);
    for (long long i = L; i < n; i += W) x[i] = stein_start((uint64_t)j, (uint64_t)i);
    tm.sync();
    // dstein scales the right-hand side to n |T|_1 max(eps, |u_nn|) and tests |x|_inf >= sqrt(0.1 / n); both are
    // written here relative to |T|_1 (as if T had unit norm) so matrices of norm 1e+-150 neither overflow nor fail
    const double unn = tm.bcast0(L == 0 ? fabs(a[n - 1]) : 0.0) / onenrm;
    for (long long i = L; i < n; i += W) a[i] = gt_pivot_recip(a[i], pert_floor);  // a <- reciprocal pivots, all lanes
    tm.sync();
    const double scl_target = (double)n * (eps > unn ? eps : unn);
    int nrmchk = 0, its = 0;
    bool ok = false;
    while (true) {
      ++its;
      if (its > STEIN_MAXITS) break;
      // scale the right-hand side to |x|_1 = n |T|_1 max(eps, |u_nn|)
      double s1 = 0.0;
      for (long long i = L; i < n; i += W) s1 += fabs(x[i]);
      s1 = tm.sum(s1);
      const double scl = s1 > 0.0 ? scl_target / s1 : 1.0;
      for (long long i = L; i < n; i += W) x[i] *= scl;
      tm.sync();
      if (L == 0) gt_solve(n, a, b, c, d2, piv, x);
      tm.sync();
      // modified Gram-Schmidt against the earlier vectors of the cluster
      for (long long q = j0; q < j; ++q) {
        const double* y = Z + (q - zoff) * ldz;
        double dot = 0.0;
        for (long long i = L; i < n; i += W) dot += y[i] * x[i];
        dot = tm.sum(dot);
        for (long long i = L; i < n; i += W) x[i] -= dot * y[i];
        tm.sync();
      }
      double mx = 0.0;
      for (long long i = L; i < n; i += W) mx = fmax(mx, fabs(x[i]));
      mx = tm.max(mx);
      if (!(mx * onenrm >= dtpcrt)) continue;  // not enough growth yet (also catches NaN)
      ++nrmchk;
      if (nrmchk < STEIN_EXTRA + 1) continue;
      ok = true;
      break;
    }
    if (!ok && L == 0 && fail) *fail += 1;
    // unit 2-norm, largest component positive
    double s2 = 0.0, mx = 0.0;
    for (long long i = L; i < n; i += W) mx = fmax(mx, fabs(x[i]));
    mx = tm.max(mx);
    const double imx = mx > 0.0 ? 1.0 / mx : 0.0;  // sum of squares of x / |x|_inf: no overflow for any scale of T
    double sgn = 0.0;
    for (long long i = L; i < n; i += W) {
      const double t = x[i] * imx;
      s2 += t * t;
      if (fabs(x[i]) == mx) sgn = x[i] < 0.0 ? -1.0 : 1.0;
    }
    s2 = tm.sum(s2);
    sgn = tm.sum(sgn) < 0.0 ? -1.0 : 1.0;
    const double inv = s2 > 0.0 ? sgn * imx / sqrt(s2) : 0.0;
    for (long long i = L; i < n; i += W) x[i] *= inv;
    tm.sync();
  }
}

}  // namespace ekb
