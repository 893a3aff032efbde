// Selected eigenpairs of the tridiagonal matrix by bisection + inverse iteration (K8a / K8b of SURVEY.md 8(d)):
// the B200 twin of the pdstebz + pdstein half of pdsyevx (reference src/solver_scalapack_select.f90:52-60).
// Needs O(n k) memory instead of the three n x n workspaces of the divide-and-conquer path, which is what makes
// BASELINE.json's config 5 (n = 65536, lowest 6554 pairs) fit.  The numerics live in tridiag.cuh and are unit-tested
// on the host against dstebz / dstein.
//
//   bisect_kernel  one eigenvalue per thread; d and e^2 (1 MiB at n = 65536) are read by all threads of a warp at
//                  the same index (broadcast loads from L1/L2), the divide chain is the limiter: not byte-scored.
//   stein_kernel   one cluster per warp, warps fetch clusters from an atomic counter; lane 0 runs the serial
//                  factor / solve recurrences, all lanes the Gram-Schmidt sweeps and norms on the columns of Z.
#include <algorithm>

#include "common.cuh"
#include "tridiag.cuh"

namespace ekb {

struct WarpTeam {
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int width() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ double max(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ double bcast0(double v) const { return __shfl_sync(0xffffffffu, v, 0); }
};

// __launch_bounds__(128, 4): with no minimum-blocks hint ptxas schedules for the fewest registers (64) and re-serialises
// the K chains that the source interleaves (seven Newton sequences back to back on shared temporaries: measured no
// faster than plain bisection); with the hint it takes 88 registers and keeps K independent instructions between
// dependent ones (checked in SASS: cuobjdump -sass, MUFU.RCP64H of the 7 chains within 30 instructions).
template <int K>
__global__ void __launch_bounds__(128, 4) bisect_kernel(i64 n, i64 j_lo, i64 j_hi, const double* __restrict__ d,
                                                     const double* __restrict__ e2, double gl, double gu, double pivmin,
                                                     double* __restrict__ w) {
  const i64 j = j_lo + (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= j_hi) return;
  w[j] = bisect_index_k<K>(n, d, e2, j, gl, gu, pivmin, nullptr);
}

__global__ void __launch_bounds__(128) stein_kernel(i64 n, const double* __restrict__ d, const double* __restrict__ e,
                                                    const double* __restrict__ w, const i64* __restrict__ starts,
                                                    i64 nclusters, double* __restrict__ Z, i64 ldz, i64 zoff,
                                                    double* __restrict__ ws_all, i64 ws_stride, double onenrm,
                                                    int* __restrict__ fail, unsigned long long* __restrict__ next) {
  const int warp_in_grid = (int)(((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  double* ws = ws_all + (i64)warp_in_grid * ws_stride;
  WarpTeam tm;
  while (true) {
    unsigned long long c = 0;
    if (tm.lane() == 0) c = atomicAdd(next, 1ULL);
    c = __shfl_sync(0xffffffffu, c, 0);
    if ((i64)c >= nclusters) break;
    stein_cluster(tm, n, d, e, w, starts[c], starts[c + 1], Z, ldz, zoff, ws, onenrm, fail + c);
  }
}

size_t stebz_stein_workspace_bytes(i64 n, int num_sms) {
  const i64 warps = (i64)num_sms * 4 * 4;  // 4 CTAs of 4 warps per SM
  const i64 stride = round_up(4 * n + (n + 7) / 8 + 8, 8);
  return (size_t)warps * stride * sizeof(double);
}

// d, e: the tridiagonal matrix on the device (not modified).  w (device, n): ALL eigenvalues ascending.  Z: columns
// [col_lo, col_hi) of the eigenvector matrix of the nev lowest eigenvalues are written to Z(:, col_lo..col_hi) (the
// caller's slab; Z is the full n x nev buffer indexed by global column -- clusters of close eigenvalues that straddle
// the slab border are computed whole, which writes a few scratch columns outside the slab).  info > 0: number of vectors whose inverse iteration did not pass
// dstein's growth test.
int stebz_stein(Ctx* ctx, i64 n, const double* d, const double* e, double* w, i64 nev, i64 col_lo, i64 col_hi, double* Z,
                i64 ldz, void* work) {
  if (n <= 0) return 0;
  if (col_hi > nev) col_hi = nev;
  // Gershgorin interval, |T|_1 and pivmin on the host from a copy of d, e (2 n doubles), exactly as dstebz sets up
  std::vector<double> hd((size_t)n), he((size_t)(n > 1 ? n - 1 : 1), 0.0), he2((size_t)n, 0.0);
  EKB_CUDA(cudaMemcpyAsync(hd.data(), d, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (n > 1) EKB_CUDA(cudaMemcpyAsync(he.data(), e, (size_t)(n - 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  double gl = hd[0], gu = hd[0], onenrm = 0.0, emax2 = 0.0;
  for (i64 i = 0; i < n; ++i) {
    const double l = i > 0 ? fabs(he[i - 1]) : 0.0, r = i + 1 < n ? fabs(he[i]) : 0.0;
    gl = std::min(gl, hd[i] - l - r);
    gu = std::max(gu, hd[i] + l + r);
    onenrm = std::max(onenrm, fabs(hd[i]) + l + r);
    if (i + 1 < n) {
      he2[i] = he[i] * he[i];
      emax2 = std::max(emax2, he2[i]);
    }
  }
  const double pivmin = TRI_SAFMIN * std::max(1.0, emax2);
  const double tnorm = std::max(fabs(gl), fabs(gu));
  gl -= 2.1 * tnorm * TRI_ULP * (double)n + 2.1 * pivmin;  // dstebz's widening of the Gershgorin interval
  gu += 2.1 * tnorm * TRI_ULP * (double)n + 2.1 * pivmin;

  double* e2 = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&e2, (size_t)(n + 8) * 8));
  i64* d_starts = nullptr;
  int* d_fail = nullptr;
  unsigned long long* d_next = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    ctx_free(ctx, e2);
    if (d_starts) ctx_free(ctx, d_starts);
    if (d_fail) ctx_free(ctx, d_fail);
    if (d_next) ctx_free(ctx, d_next);
  };
  cudaError_t ce = cudaMemcpyAsync(e2, he2.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (ce != cudaSuccess) { cleanup(); EKB_CUDA(ce); }
  {
    // 7 interleaved branch-free chains per thread (3 bits per sweep).  Small problems use one-warp CTAs so that the
    // n / 32 warps spread over all SMs instead of filling a few.
    const int threads = n <= 16384 ? 32 : 128;
    prof_begin(ctx, PROF_STEBZ, (double)n * (double)n);  // Sturm steps per bisection sweep
    bisect_kernel<7><<<cdiv(n, threads), threads, 0, ctx->stream>>>(n, 0, n, d, e2, gl, gu, pivmin, w);
    EKB_COUNT_LAUNCH(ctx);
    prof_end(ctx);
    ce = cudaGetLastError();
    if (ce != cudaSuccess) { cleanup(); EKB_CUDA(ce); }
  }
  if (col_hi <= col_lo) { cleanup(); return 0; }
  // eigenvalues back to the host: enforce the ascending order bisection guarantees up to rounding of the count,
  // then cut the rank's column range into clusters
  std::vector<double> hw((size_t)n);
  ce = cudaMemcpyAsync(hw.data(), w, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) { cleanup(); EKB_CUDA(ce); }
  bool fixed = false;
  for (i64 j = 1; j < n; ++j)
    if (hw[j] < hw[j - 1]) { hw[j] = hw[j - 1]; fixed = true; }
  if (fixed) {
    ce = cudaMemcpyAsync(w, hw.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (ce != cudaSuccess) { cleanup(); EKB_CUDA(ce); }
  }
  const double ortol = stein_ortol(nev, hw.data(), onenrm);
  std::vector<i64> starts_all((size_t)nev + 2);
  const i64 nc_all = stein_clusters(0, nev, hw.data(), ortol, (long long*)starts_all.data());
  long long first = 0, nclusters = 0;
  stein_cluster_range(nc_all, (const long long*)starts_all.data(), col_lo, col_hi, &first, &nclusters);
  const i64* starts = starts_all.data() + first;  // the clusters that intersect this rank's slab, whole
  int rc = ctx_alloc(ctx, (void**)&d_starts, (size_t)(nclusters + 1) * sizeof(i64));
  if (!rc) rc = ctx_alloc(ctx, (void**)&d_fail, (size_t)(nclusters + 1) * sizeof(int));
  if (!rc) rc = ctx_alloc(ctx, (void**)&d_next, 64);
  if (rc) { cleanup(); return rc; }
  ce = cudaMemcpyAsync(d_starts, starts, (size_t)(nclusters + 1) * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(d_fail, 0, (size_t)(nclusters + 1) * sizeof(int), ctx->stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(d_next, 0, 64, ctx->stream);
  if (ce != cudaSuccess) { cleanup(); EKB_CUDA(ce); }
  {
    const i64 stride = round_up(4 * n + (n + 7) / 8 + 8, 8);
    const int ctas = ctx->num_sms * 4;
    prof_begin(ctx, PROF_STEIN, 7.0 * 5.0 * 8.0 * (double)n * (double)(col_hi - col_lo));  // bytes: ~7 passes over 5 arrays
    stein_kernel<<<ctas, 128, 0, ctx->stream>>>(n, d, e, w, d_starts, nclusters, Z, ldz, 0, (double*)work, stride, onenrm,
                                                d_fail, d_next); EKB_COUNT_LAUNCH(ctx);
    prof_end(ctx);
    ce = cudaGetLastError();
    if (ce != cudaSuccess) { cleanup(); EKB_CUDA(ce); }
  }
  std::vector<int> hfail((size_t)nclusters);
  ce = cudaMemcpyAsync(hfail.data(), d_fail, (size_t)nclusters * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) { cleanup(); EKB_CUDA(ce); }
  cleanup();
  int nfail = 0;
  for (int f : hfail) nfail += f;
  return nfail;
}

}  // namespace ekb
