// Acceptance metrics and IPRs on the device: the B200 twins of
//   eval_residual_norm_blacs   reference src/verifier.f90:75-204   (option -c)
//   eval_orthogonality_blacs   reference src/verifier.f90:233-330  (option -t)
//   get_ipratios               reference src/distribute_matrix.f90:18-78 (ipratios.dat, main.f90:131-143)
// Same sequence of operations as the reference (pdsymm / pdscal / pdsymm / pdnrm2; pdgemm, pdgemm, scaling;
// pdgemm + the n^2 pdelget loop), with the products on the DMMA GEMM engine and the reductions as HBM-stream
// kernels whose partial sums are combined in a fixed order (deterministic).  With P > 1 ranks the eigenvector
// columns are checked slab by slab (no exchange for the residual / IPR; the Gram matrix needs the all-gathered X)
// and the per-column results are combined with one small all-reduce.
#include "common.cuh"

namespace ekb {

// out[blockIdx.y * gridDim.x + blockIdx.x] = sum of squares of a chunk of column-major A (m x n)
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const double* __restrict__ A, i64 lda, i64 m, i64 n,
                                                            double* __restrict__ out) {
  __shared__ double red[8];
  double s = 0.0;
  for (i64 j = blockIdx.y; j < n; j += gridDim.y)
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
      const double x = A[j * lda + i];
      s += x * x;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += red[q];
    out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// out[0] = sum(in[0..cnt)) in a fixed order (one block)
__global__ void __launch_bounds__(256) sum_final_kernel(const double* __restrict__ in, int cnt, double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < cnt; i += 256) s += in[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// R(:, j) *= -w[j0 + j]   (pdscal loop, verifier.f90:151-154)
__global__ void scale_cols_neg_kernel(double* __restrict__ R, i64 ldr, i64 m, i64 n, const double* __restrict__ w) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (i64 j = blockIdx.y; j < n; j += gridDim.y) R[j * ldr + i] *= -w[j];
}

// One block per column: three per-column reductions selected by mode
//   0: out[j] = sqrt(sum_i R_ij^2)                      (pdnrm2)
//   1: out[j] = sum_i V_ij^4 / (sum_i V_ij S_ij)^2      (get_ipratios; S = V for the standard problem)
__global__ void __launch_bounds__(256) col_reduce_kernel(const double* __restrict__ V, i64 ldv, const double* __restrict__ S,
                                                         i64 lds, i64 m, int mode, double* __restrict__ out) {
  __shared__ double r1[8], r2[8];
  const i64 j = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (i64 i = threadIdx.x; i < m; i += blockDim.x) {
    const double x = V[j * ldv + i];
    if (mode == 0) {
      a += x * x;
    } else {
      const double x2 = x * x;
      a += x2 * x2;
      b += x * S[j * lds + i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = a; r2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0; b = 0.0;
    for (int q = 0; q < 8; ++q) { a += r1[q]; b += r2[q]; }
    out[j] = mode == 0 ? sqrt(a) : a / (b * b);
  }
}

// dg[j] = G(row0 + j, j): diagonal of the (global) Gram matrix held as the column slab G (k x kc, rows global)
__global__ void gram_diag_kernel(const double* __restrict__ G, i64 ldg, i64 row0, i64 kc, double* __restrict__ dg) {
  const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < kc) dg[j] = G[j * ldg + row0 + j];
}

// partial sums over the slab of (a) (G_ij / sqrt(d_i d_j))^2 for the off-diagonal elements (verifier.f90:310-325) and
// (b) (G_ij - delta_ij)^2 for ALL elements (|| X^T B X - I ||_F^2, the north-star metric: it also tests normalisation)
__global__ void __launch_bounds__(256) gram_offdiag_kernel(const double* __restrict__ G, i64 ldg, i64 k, i64 row0, i64 kc,
                                                           const double* __restrict__ dall, double* __restrict__ out,
                                                           double* __restrict__ out2) {
  __shared__ double red[8], red2[8];
  double s = 0.0, s2 = 0.0;
  for (i64 j = blockIdx.y; j < kc; j += gridDim.y) {
    const double dj = dall[row0 + j];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (i64)gridDim.x * blockDim.x) {
      const double g = G[j * ldg + i];
      if (i == row0 + j) {
        s2 += (g - 1.0) * (g - 1.0);
        continue;
      }
      s2 += g * g;
      const double x = g * (1.0 / sqrt(dall[i])) * (1.0 / sqrt(dj));
      s += x * x;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; red2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, t2 = 0.0;
    for (int q = 0; q < 8; ++q) { t += red[q]; t2 += red2[q]; }
    out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    out2[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t2;
  }
}

namespace {
struct Tmp {
  Ctx* ctx;
  std::vector<void*> ptrs;
  explicit Tmp(Ctx* c) : ctx(c) {}
  int get(double** p, size_t doubles) {
    int rc = ctx_alloc(ctx, (void**)p, (doubles > 0 ? doubles : 1) * sizeof(double));
    if (rc == 0) ptrs.push_back(*p);
    return rc;
  }
  ~Tmp() {
    cudaStreamSynchronize(ctx->stream);
    for (void* p : ptrs) ctx_free(ctx, p);
  }
};

constexpr int PGX = 32, PGY = 64;  // grid of the partial-sum kernels

int frob_sumsq(Ctx* ctx, const double* A, i64 lda, i64 m, i64 n, double* partial /* PGX*PGY */, double* out1) {
  sumsq_partial_kernel<<<dim3(PGX, PGY), 256, 0, ctx->stream>>>(A, lda, m, n, partial); EKB_COUNT_LAUNCH(ctx);
  sum_final_kernel<<<1, 256, 0, ctx->stream>>>(partial, PGX * PGY, out1); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

int d2h_sync(Ctx* ctx, void* host, const void* dev, size_t bytes) {
  EKB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
}  // namespace

// A, B: full symmetric n x n on the device (B = nullptr: standard problem).  X: this rank's eigenvector columns
// X(:, c0 : c0+kc) live at Xfull + c0 * ldx, where [c0, c0+kc) is the rank's slab of the first `ncheck` columns.
// w: all eigenvalues (device).  Outputs on the host, identical on every rank.
int eval_residual_norm(Ctx* ctx, i64 n, i64 ncheck, const double* A, i64 lda, const double* B, i64 ldb, const double* w,
                       const double* Xfull, i64 ldx, double* A_norm, double* res_ave, double* res_max) {
  StageTimer total(ctx, "eval_residual_norm_b200");
  std::vector<i64> zb;
  slab_bounds(ncheck, ctx->nranks, 128, zb);
  const i64 c0 = zb[ctx->rank], kc = zb[ctx->rank + 1] - c0;
  const double* X = Xfull + c0 * ldx;
  const i64 ldr = round_up(n, 8);
  Tmp tmp(ctx);
  double *R = nullptr, *partial = nullptr, *norms = nullptr, *scal = nullptr;
  EKB_TRY(tmp.get(&R, (size_t)ldr * (kc > 0 ? kc : 1)));
  EKB_TRY(tmp.get(&partial, PGX * PGY));
  EKB_TRY(tmp.get(&norms, (size_t)ncheck + 8));
  EKB_TRY(tmp.get(&scal, 8));
  EKB_TRY(frob_sumsq(ctx, A, lda, n, n, partial, scal));  // pdlange('F') of A
  EKB_CUDA(cudaMemsetAsync(norms, 0, ((size_t)ncheck + 8) * sizeof(double), ctx->stream));
  if (kc > 0) {
    GemmP g;
    g.m = (int)n; g.n = (int)kc; g.k = (int)n; g.alpha = 1.0;
    if (B) {  // Residual <- B * Eigenvectors
      StageTimer t(ctx, "eval_residual_norm_b200:pdsymm_B");
      g.A = B; g.lda = ldb; g.B = X; g.ldb = ldx; g.C = R; g.ldc = ldr; g.beta = 0.0;
      EKB_TRY(gemm(ctx, 0, g));
      t.stop();
    } else {  // Residual <- Eigenvectors
      EKB_TRY(copy_matrix(ctx, X, ldx, R, ldr, n, kc));
    }
    scale_cols_neg_kernel<<<dim3(cdiv(n, 256), (unsigned)(kc < 32768 ? kc : 32768)), 256, 0, ctx->stream>>>(R, ldr, n, kc,
                                                                                                          w + c0);
    EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    {  // Residual <- Residual + A * Eigenvectors
      StageTimer t(ctx, "eval_residual_norm_b200:pdsymm_R");
      g.A = A; g.lda = lda; g.B = X; g.ldb = ldx; g.C = R; g.ldc = ldr; g.beta = 1.0;
      EKB_TRY(gemm(ctx, 0, g));
      t.stop();
    }
    col_reduce_kernel<<<(unsigned)kc, 256, 0, ctx->stream>>>(R, ldr, nullptr, 0, n, 0, norms + c0); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
  }
  EKB_TRY(comm_allreduce_sum(ctx, norms, (size_t)ncheck));  // slabs are disjoint: the sum is a gather
  std::vector<double> hn((size_t)ncheck + 1);
  double an2 = 0.0;
  EKB_TRY(d2h_sync(ctx, hn.data(), norms, (size_t)ncheck * sizeof(double)));
  EKB_TRY(d2h_sync(ctx, &an2, scal, sizeof(double)));
  const double an = sqrt(an2);
  double sum = 0.0, mx = 0.0;
  for (i64 j = 0; j < ncheck; ++j) { sum += hn[j]; if (hn[j] > mx) mx = hn[j]; }
  if (A_norm) *A_norm = an;
  if (res_ave) *res_ave = ncheck > 0 ? sum / an / (double)ncheck : 0.0;
  if (res_max) *res_max = mx / an;
  total.stop();
  return 0;
}

// Xfull: n x (>= index2) eigenvectors, ALL columns index1..index2 (1-based, inclusive) valid on every rank.
int eval_orthogonality(Ctx* ctx, i64 n, i64 index1, i64 index2, const double* Xfull, i64 ldx, const double* B, i64 ldb,
                       double* orthogonality, double* gram_minus_identity) {
  StageTimer total(ctx, "eval_orthogonality_b200");
  const i64 k = index2 - index1 + 1;
  const double* V = Xfull + (index1 - 1) * ldx;
  std::vector<i64> zb;
  slab_bounds(k, ctx->nranks, 128, zb);
  const i64 c0 = zb[ctx->rank], kc = zb[ctx->rank + 1] - c0;
  const i64 ldn = round_up(n, 8), ldk = round_up(k, 8);
  Tmp tmp(ctx);
  double *BV = nullptr, *G = nullptr, *dall = nullptr, *partial = nullptr, *scal = nullptr;
  if (B) EKB_TRY(tmp.get(&BV, (size_t)ldn * (kc > 0 ? kc : 1)));
  EKB_TRY(tmp.get(&G, (size_t)ldk * (kc > 0 ? kc : 1)));
  EKB_TRY(tmp.get(&dall, (size_t)k + 8));
  EKB_TRY(tmp.get(&partial, 2 * PGX * PGY));
  EKB_TRY(tmp.get(&scal, 8));
  EKB_CUDA(cudaMemsetAsync(dall, 0, ((size_t)k + 8) * sizeof(double), ctx->stream));
  EKB_CUDA(cudaMemsetAsync(scal, 0, 8 * sizeof(double), ctx->stream));
  if (kc > 0) {
    GemmP g;
    g.alpha = 1.0; g.beta = 0.0;
    const double* R = V + c0 * ldx;  // right factor of the slab of the Gram matrix
    i64 ldrr = ldx;
    if (B) {  // BV <- B * V(:, slab)
      g.m = (int)n; g.n = (int)kc; g.k = (int)n; g.A = B; g.lda = ldb; g.B = V + c0 * ldx; g.ldb = ldx; g.C = BV; g.ldc = ldn;
      EKB_TRY(gemm(ctx, 0, g));
      R = BV;
      ldrr = ldn;
    }
    // InnerProducts(:, slab) <- V' * BV
    g.m = (int)k; g.n = (int)kc; g.k = (int)n; g.A = V; g.lda = ldx; g.B = R; g.ldb = ldrr; g.C = G; g.ldc = ldk;
    EKB_TRY(gemm(ctx, GEMM_TA, g));
    gram_diag_kernel<<<cdiv(kc, 256), 256, 0, ctx->stream>>>(G, ldk, c0, kc, dall + c0); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
  }
  EKB_TRY(comm_allreduce_sum(ctx, dall, (size_t)k));
  if (kc > 0) {
    gram_offdiag_kernel<<<dim3(PGX, PGY), 256, 0, ctx->stream>>>(G, ldk, k, c0, kc, dall, partial, partial + PGX * PGY);
    EKB_COUNT_LAUNCH(ctx);
    sum_final_kernel<<<1, 256, 0, ctx->stream>>>(partial, PGX * PGY, scal); EKB_COUNT_LAUNCH(ctx);
    sum_final_kernel<<<1, 256, 0, ctx->stream>>>(partial + PGX * PGY, PGX * PGY, scal + 1); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
  }
  EKB_TRY(comm_allreduce_sum(ctx, scal, 2));
  double s2[2] = {0.0, 0.0};
  EKB_TRY(d2h_sync(ctx, s2, scal, 2 * sizeof(double)));
  if (orthogonality) *orthogonality = sqrt(s2[0]);
  if (gram_minus_identity) *gram_minus_identity = sqrt(s2[1]);
  total.stop();
  return 0;
}

// ipr (host, nvec values): IPR of the first nvec eigenvectors; X slab convention as in eval_residual_norm.
int get_ipratios(Ctx* ctx, i64 n, i64 nvec, const double* Xfull, i64 ldx, const double* B, i64 ldb, double* ipr) {
  StageTimer total(ctx, "get_ipratios_b200");
  std::vector<i64> zb;
  slab_bounds(nvec, ctx->nranks, 128, zb);
  const i64 c0 = zb[ctx->rank], kc = zb[ctx->rank + 1] - c0;
  const double* X = Xfull + c0 * ldx;
  const i64 ldn = round_up(n, 8);
  Tmp tmp(ctx);
  double *SV = nullptr, *out = nullptr;
  if (B) EKB_TRY(tmp.get(&SV, (size_t)ldn * (kc > 0 ? kc : 1)));
  EKB_TRY(tmp.get(&out, (size_t)nvec + 8));
  EKB_CUDA(cudaMemsetAsync(out, 0, ((size_t)nvec + 8) * sizeof(double), ctx->stream));
  if (kc > 0) {
    const double* S = X;
    i64 lds = ldx;
    if (B) {  // SV <- S * V (pdgemm, distribute_matrix.f90:47-48)
      GemmP g;
      g.m = (int)n; g.n = (int)kc; g.k = (int)n; g.A = B; g.lda = ldb; g.B = X; g.ldb = ldx; g.C = SV; g.ldc = ldn;
      g.alpha = 1.0; g.beta = 0.0;
      EKB_TRY(gemm(ctx, 0, g));
      S = SV;
      lds = ldn;
    }
    col_reduce_kernel<<<(unsigned)kc, 256, 0, ctx->stream>>>(X, ldx, S, lds, n, 1, out + c0); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
  }
  EKB_TRY(comm_allreduce_sum(ctx, out, (size_t)nvec));
  EKB_TRY(d2h_sync(ctx, ipr, out, (size_t)nvec * sizeof(double)));
  total.stop();
  return 0;
}

}  // namespace ekb
