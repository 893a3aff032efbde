// Host-side check library: compiles the __host__ __device__ numerics of the CUDA path for the CPU so the
// `-m "not gpu"` tests can exercise them (against LAPACK dlaed4) without a GPU.  Test infrastructure only;
// nothing on the solve path links this.
#include "layout.h"
#include "secular.cuh"
#include "tridiag.cuh"
#include <algorithm>
#include <cmath>
#include <atomic>
#include <thread>
#include <vector>

extern "C" int ekb200_host_secular(int k, const double* d, const double* z, double rho, double* lam, int* orig,
                                   double* tau, int* iters) {
  for (int j = 0; j < k; ++j) {
    int K, it;
    double t;
    ekb::secular_root(k, j, d, z, rho, &K, &t, &it);
    lam[j] = d[K] + t;
    orig[j] = K;
    tau[j] = t;
    iters[j] = it;
  }
  return 0;
}

// bounds: nranks + 1 entries
extern "C" int ekb200_host_slab_bounds(long long ncols, int nranks, int gran, long long* bounds) {
  std::vector<long long> b;
  ekb::slab_bounds(ncols, nranks, gran, b);
  for (int r = 0; r <= nranks; ++r) bounds[r] = b[r];
  return 0;
}

// the caller-side block-cyclic view of the eigenvector matrix (option "out_block"): clamp, numroc, local -> global column
extern "C" long long ekb200_host_block_clamp(long long rows, long long cols, long long want, int nprow, int npcol) {
  return ekb::reference_block_clamp(rows, cols, want, nprow, npcol);
}
extern "C" long long ekb200_host_numroc(long long n, long long nb, int iproc, int nprocs) {
  return ekb::numroc0(n, nb, iproc, nprocs);
}
extern "C" long long ekb200_host_cyclic_global_col(long long lc, long long nb, int nprocs, int r) {
  return ekb::cyclic_global_col0(lc, nb, nprocs, r);
}

extern "C" int ekb200_host_gemm_autosplit(long long m, long long n, long long k, int num_sms) {
  return ekb::gemm_autosplit_factor(m, n, k, num_sms);
}
extern "C" void ekb200_host_gemm_raster_tile(int pid, int gm, int gn, int raster, int* tm, int* tn) {
  ekb::gemm_raster_tile(pid, gm, gn, raster, tm, tn);
}

// ---- bisection + inverse iteration (tridiag.cuh), the host run of exactly the code the CUDA kernels execute
static void tri_setup(long long n, const double* d, const double* e, std::vector<double>& e2, double* gl, double* gu,
                      double* onenrm, double* pivmin) {
  e2.assign((size_t)n, 0.0);
  double lo = d[0], hi = d[0], nrm = 0.0, emax2 = 0.0;
  for (long long i = 0; i < n; ++i) {
    const double l = i > 0 ? std::fabs(e[i - 1]) : 0.0, r = i + 1 < n ? std::fabs(e[i]) : 0.0;
    lo = std::min(lo, d[i] - l - r);
    hi = std::max(hi, d[i] + l + r);
    nrm = std::max(nrm, std::fabs(d[i]) + l + r);
    if (i + 1 < n) {
      e2[i] = e[i] * e[i];
      emax2 = std::max(emax2, e2[i]);
    }
  }
  *pivmin = ekb::TRI_SAFMIN * std::max(1.0, emax2);
  const double tnorm = std::max(std::fabs(lo), std::fabs(hi));
  *gl = lo - 2.1 * tnorm * ekb::TRI_ULP * (double)n - 2.1 * *pivmin;
  *gu = hi + 2.1 * tnorm * ekb::TRI_ULP * (double)n + 2.1 * *pivmin;
  *onenrm = nrm;
}

// sections = 1 (plain bisection), 3 or 7 interior points per sweep: the variants the CUDA kernel instantiates
extern "C" int ekb200_host_stebz_k(long long n, const double* d, const double* e, int sections, double* w, int* max_iters) {
  std::vector<double> e2;
  double gl, gu, onenrm, pivmin;
  tri_setup(n, d, e, e2, &gl, &gu, &onenrm, &pivmin);
  int mx = 0;
  for (long long j = 0; j < n; ++j) {
    int it = 0;
    w[j] = sections == 7   ? ekb::bisect_index_k<7>(n, d, e2.data(), j, gl, gu, pivmin, &it)
           : sections == 3 ? ekb::bisect_index_k<3>(n, d, e2.data(), j, gl, gu, pivmin, &it)
                           : ekb::bisect_index_k<1>(n, d, e2.data(), j, gl, gu, pivmin, &it);
    mx = std::max(mx, it);
  }
  if (max_iters) *max_iters = mx;
  return 0;
}

extern "C" int ekb200_host_stebz(long long n, const double* d, const double* e, double* w, int* max_iters) {
  std::vector<double> e2;
  double gl, gu, onenrm, pivmin;
  tri_setup(n, d, e, e2, &gl, &gu, &onenrm, &pivmin);
  int mx = 0;
  for (long long j = 0; j < n; ++j) {
    int it = 0;
    w[j] = ekb::bisect_index(n, d, e2.data(), j, gl, gu, pivmin, &it);
    mx = std::max(mx, it);
  }
  if (max_iters) *max_iters = mx;
  return 0;
}

// eigenvectors of w[0..nev) (ascending, as produced by ekb200_host_stebz); returns the number of failed vectors,
// *nclusters / *max_cluster describe the clustering
// the same for a rank's column slab [col_lo, col_hi) of the nev requested vectors (Z: full n x nev buffer)
extern "C" int ekb200_host_stein_slab(long long n, const double* d, const double* e, long long nev, const double* w,
                                      long long col_lo, long long col_hi, double* Z, long long ldz) {
  std::vector<double> e2;
  double gl, gu, onenrm, pivmin;
  tri_setup(n, d, e, e2, &gl, &gu, &onenrm, &pivmin);
  const double ortol = ekb::stein_ortol(nev, w, onenrm);
  std::vector<long long> starts((size_t)nev + 2);
  const long long nc = ekb::stein_clusters(0, nev, w, ortol, starts.data());
  long long first = 0, count = 0;
  ekb::stein_cluster_range(nc, starts.data(), col_lo, col_hi, &first, &count);
  std::vector<double> ws((size_t)(4 * n + (n + 7) / 8 + 8));
  ekb::HostTeam tm;
  int fail = 0;
  for (long long c = first; c < first + count; ++c)
    ekb::stein_cluster(tm, n, d, e, w, starts[c], starts[c + 1], Z, ldz, 0, ws.data(), onenrm, &fail);
  return fail;
}

extern "C" int ekb200_host_stein(long long n, const double* d, const double* e, long long nev, const double* w, double* Z,
                                 long long ldz, long long* nclusters, long long* max_cluster) {
  std::vector<double> e2;
  double gl, gu, onenrm, pivmin;
  tri_setup(n, d, e, e2, &gl, &gu, &onenrm, &pivmin);
  const double ortol = ekb::stein_ortol(nev, w, onenrm);
  std::vector<long long> starts((size_t)nev + 2);
  const long long nc = ekb::stein_clusters(0, nev, w, ortol, starts.data());
  std::vector<double> ws((size_t)(4 * n + (n + 7) / 8 + 8));
  ekb::HostTeam tm;
  int fail = 0;
  long long mc = 0;
  for (long long c = 0; c < nc; ++c) {
    ekb::stein_cluster(tm, n, d, e, w, starts[c], starts[c + 1], Z, ldz, 0, ws.data(), onenrm, &fail);
    mc = std::max(mc, starts[c + 1] - starts[c]);
  }
  if (nclusters) *nclusters = nc;
  if (max_cluster) *max_cluster = mc;
  return fail;
}

// ---- the per-cluster procedure run the way a warp runs it: W host threads in lock step stand in for the lanes
// (lane 0 runs the serial recurrences, all lanes the strided vector work), with the butterfly reduction order of the
// __shfl_xor_sync sums.  Checks the SPMD orchestration of stein_cluster (sync placement, lane-0-only sections,
// strided loops) on the CPU.
namespace {
struct SpinBarrier {
  std::atomic<int> count{0}, gen{0};
  int n;
  explicit SpinBarrier(int n_) : n(n_) {}
  void wait() {
    const int g = gen.load();
    if (count.fetch_add(1) + 1 == n) {
      count.store(0);
      gen.fetch_add(1);
    } else {
      while (gen.load() == g) std::this_thread::yield();
    }
  }
};
struct LaneTeam {
  int lane_, width_;
  SpinBarrier* bar;
  double* slots;
  int lane() const { return lane_; }
  int width() const { return width_; }
  void sync() const { bar->wait(); }
  double sum(double v) const {
    for (int o = width_ / 2; o > 0; o >>= 1) {
      slots[lane_] = v;
      bar->wait();
      v += slots[lane_ ^ o];
      bar->wait();
    }
    return v;
  }
  double max(double v) const {
    for (int o = width_ / 2; o > 0; o >>= 1) {
      slots[lane_] = v;
      bar->wait();
      v = std::fmax(v, slots[lane_ ^ o]);
      bar->wait();
    }
    return v;
  }
  double bcast0(double v) const {
    if (lane_ == 0) slots[0] = v;
    bar->wait();
    const double r = slots[0];
    bar->wait();
    return r;
  }
};
}  // namespace

extern "C" int ekb200_host_stein_lanes(long long n, const double* d, const double* e, long long nev, const double* w,
                                       int lanes, double* Z, long long ldz) {
  std::vector<double> e2;
  double gl, gu, onenrm, pivmin;
  tri_setup(n, d, e, e2, &gl, &gu, &onenrm, &pivmin);
  const double ortol = ekb::stein_ortol(nev, w, onenrm);
  std::vector<long long> starts((size_t)nev + 2);
  const long long nc = ekb::stein_clusters(0, nev, w, ortol, starts.data());
  std::vector<double> ws((size_t)(4 * n + (n + 7) / 8 + 8));
  std::vector<double> slots((size_t)lanes);
  std::vector<int> fails((size_t)nc + 1, 0);
  SpinBarrier bar(lanes);
  std::vector<std::thread> th;
  for (int l = 0; l < lanes; ++l)
    th.emplace_back([&, l]() {
      LaneTeam tm{l, lanes, &bar, slots.data()};
      for (long long c = 0; c < nc; ++c)
        ekb::stein_cluster(tm, n, d, e, w, starts[c], starts[c + 1], Z, ldz, 0, ws.data(), onenrm, &fails[c]);
    });
  for (auto& t : th) t.join();
  int fail = 0;
  for (int f : fails) fail += f;
  return fail;
}
