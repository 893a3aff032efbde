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

// The j-th smallest eigenvalue (0-based) inside [gl, gu] (a Gershgorin interval widened like dstebz does).
EKB_HD double bisect_index(long long n, const double* d, const double* e2, long long j, double gl, double gu,
                           double pivmin, int* iters) {
  double lo = gl, hi = gu;
  int it = 0;
  for (; it < 128; ++it) {
    const double mid = 0.5 * (lo + hi);
    const double tol = fmax(2.0 * TRI_SAFMIN, fmax(pivmin, 2.0 * TRI_ULP * fmax(fabs(lo), fabs(hi))));
    if (hi - lo <= tol || mid <= lo || mid >= hi) break;
    if (sturm_count(n, d, e2, mid, pivmin) >= j + 1) hi = mid;
    else lo = mid;
  }
  if (iters) *iters = it;
  return 0.5 * (lo + hi);
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
EKB_HD void gt_factor(long long n, const double* d, const double* e, double lambda, double* a, double* b, double* c,
                      double* d2, unsigned char* piv) {
  for (long long i = 0; i < n; ++i) {
    a[i] = d[i] - lambda;
    b[i] = i + 1 < n ? e[i] : 0.0;
    d2[i] = 0.0;
    c[i] = 0.0;
    piv[i] = 0;
  }
  for (long long i = 0; i + 1 < n; ++i) {
    const double sub = e[i];  // element (i+1, i)
    if (fabs(a[i]) >= fabs(sub)) {
      const double m = a[i] != 0.0 ? sub / a[i] : 0.0;
      c[i] = m;
      a[i + 1] -= m * b[i];
    } else {
      const double m = a[i] / sub;
      piv[i] = 1;
      c[i] = m;
      const double t = a[i + 1];
      a[i + 1] = b[i] - m * t;
      a[i] = sub;
      b[i] = t;
      if (i + 2 < n) {
        d2[i] = b[i + 1];
        b[i + 1] = -m * d2[i];
      }
    }
  }
}

// x <- (T - lambda I)^-1 x with the factors of gt_factor; pivots smaller than `pert` are replaced by +-pert and
// enlarged when the quotient would overflow (dlagts with job = -1).
EKB_HD void gt_solve(long long n, const double* a, const double* b, const double* c, const double* d2,
                     const unsigned char* piv, double* x, double pert) {
  for (long long i = 0; i + 1 < n; ++i) {
    if (piv[i]) {
      const double t = x[i];
      x[i] = x[i + 1];
      x[i + 1] = t - c[i] * x[i];
    } else {
      x[i + 1] -= c[i] * x[i];
    }
  }
  double x1 = 0.0, x2 = 0.0;  // x[i+1], x[i+2]
  for (long long i = n - 1; i >= 0; --i) {
    const double s = x[i] - b[i] * x1 - d2[i] * x2;
    double p = a[i];
    if (fabs(p) < pert) p = p < 0.0 ? -pert : pert;
    if (fabs(s) > fabs(p) * 1e290) p = (p < 0.0 ? -1.0 : 1.0) * fabs(s) * 1e-290;
    const double v = s / p;
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
      if (L == 0) gt_solve(n, a, b, c, d2, piv, x, pert_floor);
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
