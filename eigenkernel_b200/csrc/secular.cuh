// Secular-equation root finder shared by the CUDA kernel (stedc.cu) and the host-side check library
// (host_check.cpp), so the numerics can be unit-tested on a CPU against LAPACK's dlaed4.
// Replaces the dlaed4 step inside pdstedc (reference src/solver_scalapack_all.f90:96).
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define EKB_HD __host__ __device__ __forceinline__
#else
#define EKB_HD inline
#endif

namespace ekb {

constexpr double DC_EPS = 1.1102230246251565e-16;  // dlamch('E')

// Root j (0-based) of  1/rho + sum_i z_i^2 / (d_i - lambda) = 0,  d ascending and distinct, rho > 0.
// Returns the origin pole K and the offset tau (lambda = d[K] + tau) so that d_i - lambda can be formed
// as (d_i - d_K) - tau to high relative accuracy.  `iters` (optional) receives the iteration count.
EKB_HD void secular_root(int k, int j, const double* d, const double* z, double rho, int* K_out, double* tau_out,
                         int* iters) {
  const double rhoinv = 1.0 / rho;
  int nit = 0;
  int K;
  double tau;
  if (k == 1) {
    K = 0;
    tau = rho * z[0] * z[0];
  } else {
    const bool last = (j == k - 1);
    double lo, hi;
    if (!last) {
      const double del = d[j + 1] - d[j], half = 0.5 * del;
      // w at the midpoint, origin d_j
      double wmid = rhoinv, c = rhoinv;
      for (int i = 0; i < k; ++i) {
        const double t = z[i] * z[i] / ((d[i] - d[j]) - half);
        wmid += t;
        if (i != j && i != j + 1) c += t;
      }
      const double zj2 = z[j] * z[j], zj12 = z[j + 1] * z[j + 1];
      double a, b;
      if (wmid >= 0.0) {
        K = j; lo = 0.0; hi = half;
        a = c * del + zj2 + zj12; b = zj2 * del;
      } else {
        K = j + 1; lo = -half; hi = 0.0;
        a = -c * del + zj2 + zj12; b = -zj12 * del;
      }
      const double disc = sqrt(fabs(a * a - 4.0 * b * c));
      if (c == 0.0) tau = b / a;
      else if (a <= 0.0) tau = (a - disc) / (2.0 * c);
      else tau = 2.0 * b / (a + disc);
      if (!(tau > lo && tau < hi)) tau = 0.5 * (lo + hi);
    } else {
      K = k - 1;
      double zz = 0.0;
      for (int i = 0; i < k; ++i) zz += z[i] * z[i];
      lo = 0.0; hi = rho * zz;
      // one-pole guess: 1/rho - zK^2/tau ~ 0  -> tau ~ rho zK^2 (always inside the bracket)
      tau = rho * z[K] * z[K];
      if (!(tau > lo && tau < hi)) tau = 0.5 * hi;
    }
    const double dK = d[K];
    double wprev = 0.0;
    int slow = 0;
    for (int it = 0; it < 80; ++it) {
      nit = it + 1;
      // evaluate
      double psi = 0.0, phi = 0.0, dpsi = 0.0, dphi = 0.0, asum = 0.0;
      const int jsplit = last ? k - 1 : j;  // poles <= jsplit are below the root
      for (int i = 0; i < k; ++i) {
        const double dl = (d[i] - dK) - tau;
        const double q = z[i] / dl;
        const double t = z[i] * q;
        if (i <= jsplit) { psi += t; dpsi += q * q; } else { phi += t; dphi += q * q; }
        asum += fabs(t);
      }
      const double w = rhoinv + psi + phi;
      const double dw = dpsi + dphi;
      const double erretm = 8.0 * asum + rhoinv + fabs(tau) * dw;
      if (fabs(w) <= DC_EPS * erretm) break;
      if (w < 0.0) lo = fmax(lo, tau); else hi = fmin(hi, tau);
      if (!(hi - lo > 4.0 * DC_EPS * fmax(fabs(lo), fabs(hi)))) { tau = 0.5 * (lo + hi); break; }
      double eta;
      if (!last) {
        const double dj = (d[j] - dK) - tau, dj1 = (d[j + 1] - dK) - tau;
        const double c = w - dj * dpsi - dj1 * dphi;
        const double a = (dj + dj1) * w - dj * dj1 * dw;
        const double b = dj * dj1 * w;
        const double disc = sqrt(fabs(a * a - 4.0 * b * c));
        if (c == 0.0) eta = b / a;
        else if (a <= 0.0) eta = (a - disc) / (2.0 * c);
        else eta = 2.0 * b / (a + disc);
      } else {
        // exterior root: one-pole model through (w, dw) with the pole at d_K
        const double dl = -tau;  // d_K - lambda
        const double den = w - dw * dl;
        eta = dl + dw * dl * dl / den;
      }
      if (w * eta >= 0.0 || !isfinite(eta)) eta = -w / dw;  // Newton fallback
      double tn = tau + eta;
      if (it > 0 && fabs(w) > 0.5 * fabs(wprev)) ++slow; else slow = 0;
      if (!(tn > lo && tn < hi) || slow >= 2) { tn = 0.5 * (lo + hi); slow = 0; }
      wprev = w;
      tau = tn;
    }
  }
  *K_out = K;
  *tau_out = tau;
  if (iters) *iters = nit;
}

}  // namespace ekb
