// Generalized <-> standard conversion: blocked Cholesky (PDPOTRF), reduction to standard form (PDSYGST)
// and the triangular back-substitution (PDTRTRS).  Reference call sites:
//   src/generalized_to_standard.f90:24   pdpotrf('L', dim, B, ...)            -> potrf_lower
//   src/generalized_to_standard.f90:37   pdsygst(1, 'L', dim, A, ..., B, ...) -> sygst_lower
//   src/generalized_to_standard.f90:103  pdtrtrs('L','T','N', dim, n_vec, B, ..., V, ...) -> trsm_lower(TRSM_LLT)
//
// B200 design: everything except 64x64 diagonal-block work is a GEMM on the DMMA engine.  The
// factorization and the triangular solves recurse by halves (multiples of 64) so the trailing updates
// are large GEMMs; diagonal blocks are factored and INVERTED in shared memory by one CTA, and a
// diagonal-block solve is a GEMM with the inverted block.
#include "common.cuh"

namespace ekb {

constexpr int NB = 64;  // diagonal block size of the triangular machinery
constexpr size_t LEAF_SMEM = 2 * NB * (NB + 1) * sizeof(double);

// One CTA factors an nb x nb (nb <= 64) diagonal block in shared memory, writes L back (lower part)
// and its inverse (lower triangular, zero padded to 64x64) to inv.  On a non-positive pivot the first
// failing global index (1-based) is recorded with atomicMin in *info (0 = none yet is stored as INT_MAX).
__global__ void __launch_bounds__(256) potrf_leaf_kernel(double* __restrict__ A, i64 lda, int nb, double* __restrict__ inv,
                                                         int* __restrict__ info, int global_offset) {
  extern __shared__ double dsm[];
  double(*s)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(dsm);
  double(*si)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(dsm + NB * (NB + 1));
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    int r = e % NB, c = e / NB;
    s[r][c] = (r < nb && c < nb && r >= c) ? A[(i64)c * lda + r] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    double d = s[j][j];
    __syncthreads();
    if (!(d > 0.0)) {
      if (tid == 0) atomicMin(info, global_offset + j + 1);
      d = 1.0;
    }
    double sd = sqrt(d);
    if (tid < NB) {
      if (tid == j) s[j][j] = sd;
      else if (tid > j) s[tid][j] /= sd;
    }
    __syncthreads();
    // trailing update of the lower triangle: s[r][c] -= s[r][j] * s[c][j], r >= c > j
    const int rem = nb - j - 1;
    for (int e = tid; e < rem * rem; e += blockDim.x) {
      int r = j + 1 + e % rem, c = j + 1 + e / rem;
      if (r >= c) s[r][c] -= s[r][j] * s[c][j];
    }
    __syncthreads();
  }
  // inverse by forward substitution, one column per thread
  if (tid < NB) {
    const int c = tid;
    for (int i = 0; i < NB; ++i) si[i][c] = 0.0;
    if (c < nb) {
      for (int i = c; i < nb; ++i) {
        double acc = (i == c) ? 1.0 : 0.0;
        for (int k = c; k < i; ++k) acc -= s[i][k] * si[k][c];
        si[i][c] = acc / s[i][i];
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    int r = e % NB, c = e / NB;
    if (r < nb && c < nb && r >= c) A[(i64)c * lda + r] = s[r][c];
    inv[e] = si[r][c];
  }
}

// Invert every 64x64 diagonal block of a given lower-triangular L (used when L did not come from potrf_lower).
__global__ void __launch_bounds__(64) trtri_blocks_kernel(const double* __restrict__ L, i64 ldl, i64 n,
                                                          double* __restrict__ invd) {
  extern __shared__ double dsm[];
  double(*s)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(dsm);
  double(*si)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(dsm + NB * (NB + 1));
  const int blk = blockIdx.x, tid = threadIdx.x;
  const i64 o = (i64)blk * NB;
  const int nb = (int)((n - o) < NB ? (n - o) : NB);
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    int r = e % NB, c = e / NB;
    s[r][c] = (r < nb && c < nb && r >= c) ? L[(o + c) * ldl + o + r] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  const int c = tid;
  for (int i = 0; i < NB; ++i) si[i][c] = 0.0;
  if (c < nb) {
    for (int i = c; i < nb; ++i) {
      double acc = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) acc -= s[i][k] * si[k][c];
      si[i][c] = acc / s[i][i];
    }
  }
  __syncthreads();
  double* inv = invd + (i64)blk * NB * NB;
  for (int e = tid; e < NB * NB; e += blockDim.x) inv[e] = si[e % NB][e / NB];
}

int trtri_diag_blocks(Ctx* ctx, i64 n, const double* L, i64 ldl, double* invd) {
  if (n <= 0) return 0;
  static bool attr_dev[64] = {};  // per device: the attribute belongs to the device's context
  bool& attr = attr_dev[ctx->device & 63];
  if (!attr) {
    EKB_CUDA(cudaFuncSetAttribute(trtri_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEAF_SMEM));
    attr = true;
  }
  trtri_blocks_kernel<<<cdiv(n, NB), 64, LEAF_SMEM, ctx->stream>>>(L, ldl, n, invd); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- recursive triangular solve
// tmp: workspace of at least max(m,n) * 64 doubles.
static int trsm_rec(Ctx* ctx, int kind, i64 m, i64 n, const double* L, i64 ldl, const double* invd, double* B,
                    i64 ldb, double* tmp) {
  const i64 ln = (kind == TRSM_RLT) ? n : m;  // order of L
  if (ln <= 0 || m <= 0 || n <= 0) return 0;
  if (ln <= NB) {
    GemmP p;
    p.alpha = 1.0;
    p.beta = 0.0;
    // IN PLACE (C = the operand being solved for): the leaf has a single tile in the solved dimension (ln <= 64 = one
    // column tile for RLT, one row tile otherwise) and runs through all of k = ln before its epilogue, so the CTA
    // that writes a tile of B is the only one that ever read it, and it has read all of it by then (beta = 0: the
    // epilogue does not read C).  Saves the copy launch every leaf used to need.
    if (kind == TRSM_RLT) {  // X = B * inv^T  (m x ln)
      p.m = (int)m; p.n = (int)ln; p.k = (int)ln;
      p.A = B; p.lda = ldb; p.B = invd; p.ldb = NB; p.C = B; p.ldc = ldb;
      EKB_TRY(gemm(ctx, GEMM_TB, p));
    } else {  // X = inv * B or inv^T * B   (ln x n)
      p.m = (int)ln; p.n = (int)n; p.k = (int)ln;
      p.A = invd; p.lda = NB; p.B = B; p.ldb = ldb; p.C = B; p.ldc = ldb;
      EKB_TRY(gemm(ctx, kind == TRSM_LLT ? GEMM_TA : 0, p));
    }
    return 0;
  }
  const i64 nblk = (ln + NB - 1) / NB;
  const i64 n1 = ((nblk + 1) / 2) * NB, n2 = ln - n1;
  const double* L11 = L;
  const double* L21 = L + n1;
  const double* L22 = L + n1 * ldl + n1;
  const double* inv2 = invd + (n1 / NB) * NB * NB;
  GemmP p;
  p.alpha = -1.0;
  p.beta = 1.0;
  if (kind == TRSM_RLT) {
    double* B1 = B;
    double* B2 = B + n1 * ldb;
    EKB_TRY(trsm_rec(ctx, kind, m, n1, L11, ldl, invd, B1, ldb, tmp));
    p.m = (int)m; p.n = (int)n2; p.k = (int)n1;  // B2 -= X1 * L21^T
    p.A = B1; p.lda = ldb; p.B = L21; p.ldb = ldl; p.C = B2; p.ldc = ldb;
    EKB_TRY(gemm(ctx, GEMM_TB, p));
    EKB_TRY(trsm_rec(ctx, kind, m, n2, L22, ldl, inv2, B2, ldb, tmp));
  } else if (kind == TRSM_LLN) {
    double* B1 = B;
    double* B2 = B + n1;
    EKB_TRY(trsm_rec(ctx, kind, n1, n, L11, ldl, invd, B1, ldb, tmp));
    p.m = (int)n2; p.n = (int)n; p.k = (int)n1;  // B2 -= L21 * X1
    p.A = L21; p.lda = ldl; p.B = B1; p.ldb = ldb; p.C = B2; p.ldc = ldb;
    EKB_TRY(gemm(ctx, 0, p));
    EKB_TRY(trsm_rec(ctx, kind, n2, n, L22, ldl, inv2, B2, ldb, tmp));
  } else {  // TRSM_LLT
    double* B1 = B;
    double* B2 = B + n1;
    EKB_TRY(trsm_rec(ctx, kind, n2, n, L22, ldl, inv2, B2, ldb, tmp));
    p.m = (int)n1; p.n = (int)n; p.k = (int)n2;  // B1 -= L21^T * X2
    p.A = L21; p.lda = ldl; p.B = B2; p.ldb = ldb; p.C = B1; p.ldc = ldb;
    EKB_TRY(gemm(ctx, GEMM_TA, p));
    EKB_TRY(trsm_rec(ctx, kind, n1, n, L11, ldl, invd, B1, ldb, tmp));
  }
  return 0;
}

int trsm_lower(Ctx* ctx, int kind, i64 m, i64 n, const double* L, i64 ldl, const double* invd, double* Bm, i64 ldb) {
  if (m <= 0 || n <= 0) return 0;
  double* tmp = nullptr;
  i64 mx = m > n ? m : n;
  EKB_TRY(ctx_alloc(ctx, (void**)&tmp, (size_t)round_up(mx, 2) * NB * sizeof(double)));
  int rc = trsm_rec(ctx, kind, m, n, L, ldl, invd, Bm, ldb, tmp);
  cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, tmp);
  return rc;
}

// ---------------------------------------------------------------- recursive Cholesky
static int potrf_rec(Ctx* ctx, i64 n, double* A, i64 lda, double* invd, i64 goff, double* tmp) {
  if (n <= 0) return 0;
  if (n <= NB) {
    static bool attr_dev[64] = {};  // per device: the attribute belongs to the device's context
    bool& attr = attr_dev[ctx->device & 63];
    if (!attr) {
      EKB_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEAF_SMEM));
      attr = true;
    }
    potrf_leaf_kernel<<<1, 256, LEAF_SMEM, ctx->stream>>>(A, lda, (int)n, invd, ctx->d_info, (int)goff); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    return 0;
  }
  const i64 nblk = (n + NB - 1) / NB;
  const i64 n1 = ((nblk + 1) / 2) * NB, n2 = n - n1;
  double* A21 = A + n1;
  double* A22 = A + n1 * lda + n1;
  EKB_TRY(potrf_rec(ctx, n1, A, lda, invd, goff, tmp));
  EKB_TRY(trsm_rec(ctx, TRSM_RLT, n2, n1, A, lda, invd, A21, lda, tmp));
  GemmP p;
  p.m = (int)n2; p.n = (int)n2; p.k = (int)n1;
  p.A = A21; p.lda = lda; p.B = A21; p.ldb = lda; p.C = A22; p.ldc = lda;
  p.alpha = -1.0; p.beta = 1.0;
  EKB_TRY(gemm(ctx, GEMM_TB, p, /*tri_keep=*/1));
  return potrf_rec(ctx, n2, A22, lda, invd + (n1 / NB) * NB * NB, goff + n1, tmp);
}

// Sharded recursion (P > 1 ranks, every rank holds the same A): at every level of at least POTRF_DIST_MIN
// columns the two GEMM-shaped steps are dealt to the ranks --
//   A21 <- A21 L11^-T   by ROW slabs of A21 (each rank solves for its rows), then an all-gather;
//   A22 -= A21 A21^T    by COLUMN slabs of the lower triangle (boundaries chosen for equal area), then an all-gather;
// -- while the small diagonal sub-problems stay replicated (deterministic kernels: identical bits on every rank).
// This is pdpotrf's panel broadcast + trailing update with the process grid folded onto the GPUs of the box.
constexpr i64 POTRF_DIST_MIN = 2048;

static int potrf_dist_rec(Ctx* ctx, i64 n, double* A, i64 lda, double* invd, i64 goff, double* tmp, double* pack) {
  if (ctx->nranks <= 1 || n < POTRF_DIST_MIN) return potrf_rec(ctx, n, A, lda, invd, goff, tmp);
  const int P = ctx->nranks, r = ctx->rank;
  const i64 nblk = (n + NB - 1) / NB;
  const i64 n1 = ((nblk + 1) / 2) * NB, n2 = n - n1;
  double* A21 = A + n1;
  double* A22 = A + n1 * lda + n1;
  EKB_TRY(potrf_dist_rec(ctx, n1, A, lda, invd, goff, tmp, pack));
  {  // row slabs of A21
    std::vector<i64> rb;
    slab_bounds(n2, P, NB, rb);
    const i64 chunk = rb[1] - rb[0];  // the largest slab (slab 0)
    const i64 r0 = rb[r], nr = rb[r + 1] - rb[r];
    if (nr > 0) {
      EKB_TRY(trsm_rec(ctx, TRSM_RLT, nr, n1, A, lda, invd, A21 + r0, lda, tmp));
      EKB_TRY(copy_matrix(ctx, A21 + r0, lda, pack + (size_t)r * chunk * n1, chunk, nr, n1));
    }
    EKB_TRY(comm_allgather(ctx, pack + (size_t)r * chunk * n1, pack, (size_t)chunk * n1));
    for (int q = 0; q < P; ++q) {
      const i64 q0 = rb[q], nq = rb[q + 1] - rb[q];
      if (q != r && nq > 0) EKB_TRY(copy_matrix(ctx, pack + (size_t)q * chunk * n1, chunk, A21 + q0, lda, nq, n1));
    }
  }
  {  // column slabs of the lower triangle of A22, equal areas: c_q = n2 (1 - sqrt(1 - q / P))
    std::vector<i64> cb(P + 1, 0);
    for (int q = 1; q < P; ++q) {
      i64 c = (i64)((double)n2 * (1.0 - sqrt(1.0 - (double)q / P)));
      c = c / NB * NB;
      cb[q] = c < cb[q - 1] ? cb[q - 1] : (c > n2 ? n2 : c);
    }
    cb[P] = n2;
    const i64 c0 = cb[r], nc = cb[r + 1] - cb[r];
    if (nc > 0) {
      GemmP p;
      p.m = (int)(n2 - c0); p.n = (int)nc; p.k = (int)n1;
      p.A = A21 + c0; p.lda = lda; p.B = A21 + c0; p.ldb = lda; p.C = A22 + c0 * lda + c0; p.ldc = lda;
      p.alpha = -1.0; p.beta = 1.0;
      EKB_TRY(gemm(ctx, GEMM_TB, p, /*tri_keep=*/1));
    }
    // whole columns of the enclosing matrix (row 0 .. ld): starting at row goff + n1 instead would run the last
    // slab past the end of the allocation by that many elements
    EKB_TRY(comm_allgather_cols(ctx, A22 - (goff + n1), lda, cb));
  }
  return potrf_dist_rec(ctx, n2, A22, lda, invd + (n1 / NB) * NB * NB, goff + n1, tmp, pack);
}

int potrf_lower(Ctx* ctx, i64 n, double* B, i64 ldb, double* invd) {
  if (n <= 0) return 0;
  double *tmp = nullptr, *pack = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&tmp, (size_t)round_up(n, 2) * NB * sizeof(double)));
  if (ctx->nranks > 1 && n >= POTRF_DIST_MIN) {
    const i64 h = (((n + NB - 1) / NB + 1) / 2) * NB;  // largest n1; n2 <= n1
    const size_t pack_doubles = (size_t)(h + (i64)NB * ctx->nranks) * h;
    int rc = ctx_alloc(ctx, (void**)&pack, pack_doubles * sizeof(double));
    if (rc) { ctx_free(ctx, tmp); return rc; }
  }
  *ctx->h_info = 0x7fffffff;
  EKB_CUDA(cudaMemcpyAsync(ctx->d_info, ctx->h_info, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  int rc = potrf_dist_rec(ctx, n, B, ldb, invd, 0, tmp, pack);
  if (rc == 0) {
    EKB_CUDA(cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EKB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_info != 0x7fffffff) rc = *ctx->h_info;
  }
  ctx_free(ctx, tmp);
  if (pack) ctx_free(ctx, pack);
  return rc;
}

// ---------------------------------------------------------------- reduction to standard form
// Leaf: A11 <- inv A11 inv^T for one diagonal block (nb <= 64); reads the lower triangle, writes BOTH triangles.
__global__ void __launch_bounds__(256) sygst_leaf_kernel(double* __restrict__ A, i64 lda, int nb, const double* __restrict__ inv) {
  extern __shared__ double dsm[];
  double(*sa)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(dsm);
  double(*si)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(dsm + NB * (NB + 1));
  double(*st)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(dsm + 2 * NB * (NB + 1));
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int r = e % NB, c = e / NB;
    if (r >= c) {
      const double x = (r < nb && c < nb) ? A[(i64)c * lda + r] : 0.0;
      sa[r][c] = x;
      sa[c][r] = x;
    }
    si[r][c] = inv[e];  // inv(r, c), lower triangular, zero padded
  }
  __syncthreads();
  // T = inv * A  (inv lower: k <= r)
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int r = e % NB, c = e / NB;
    double acc = 0.0;
    for (int k = 0; k <= r; ++k) acc += si[r][k] * sa[k][c];
    st[r][c] = acc;
  }
  __syncthreads();
  // R = T * inv^T : R(r, c) = sum_{k <= c} T(r, k) inv(c, k)
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int r = e % NB, c = e / NB;
    if (r < nb && c < nb) {
      double acc = 0.0;
      for (int k = 0; k <= c; ++k) acc += st[r][k] * si[c][k];
      A[(i64)c * lda + r] = acc;
    }
  }
}

// C (m x n) -= M
__global__ void sub_matrix_kernel(double* __restrict__ C, i64 ldc, const double* __restrict__ M, i64 ldm, i64 m, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (i64 j = blockIdx.y; j < n; j += gridDim.y) C[j * ldc + i] -= M[j * ldm + i];
}

// Recursive dsygst(itype = 1, 'L') on the LOWER triangle (n^3 FLOPs, all GEMM-shaped):
//   A11 <- sygst(A11);  A21 <- A21 L11^-T;  M = 1/2 L21 A11;  A21 -= M;
//   A22 -= A21 L21^T + L21 A21^T (lower);  A21 -= M;  A21 <- L22^-1 A21;  A22 <- sygst(A22).
// Diagonal leaves come back with both triangles; every finished A11 is mirrored to full before it is used as the
// symmetric factor of M.  Mws: scratch of at least ceil(n/2)^2 doubles (rounded to whole 64-blocks).
static int sygst_rec(Ctx* ctx, i64 n, double* A, i64 lda, const double* L, i64 ldl, const double* invd, double* tmp,
                     double* Mws) {
  if (n <= 0) return 0;
  if (n <= NB) {
    static bool attr_dev[64] = {};  // per device: the attribute belongs to the device's context
    bool& attr = attr_dev[ctx->device & 63];
    constexpr size_t smem = 3 * NB * (NB + 1) * sizeof(double);
    if (!attr) {
      EKB_CUDA(cudaFuncSetAttribute(sygst_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    sygst_leaf_kernel<<<1, 256, smem, ctx->stream>>>(A, lda, (int)n, invd); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    return 0;
  }
  const i64 nblk = (n + NB - 1) / NB;
  const i64 n1 = ((nblk + 1) / 2) * NB, n2 = n - n1;
  double* A21 = A + n1;
  double* A22 = A + n1 * lda + n1;
  const double* L21 = L + n1;
  const double* L22 = L + n1 * ldl + n1;
  const double* inv2 = invd + (n1 / NB) * NB * NB;
  EKB_TRY(sygst_rec(ctx, n1, A, lda, L, ldl, invd, tmp, Mws));
  if (n1 > NB) EKB_TRY(symmetrize_from_lower(ctx, A, lda, n1));
  EKB_TRY(trsm_rec(ctx, TRSM_RLT, n2, n1, L, ldl, invd, A21, lda, tmp));
  GemmP p;
  p.m = (int)n2; p.n = (int)n1; p.k = (int)n1;
  p.A = L21; p.lda = ldl; p.B = A; p.ldb = lda; p.C = Mws; p.ldc = n2;
  p.alpha = 0.5; p.beta = 0.0;
  EKB_TRY(gemm(ctx, 0, p));
  dim3 grid(cdiv(n2, 256), (unsigned)(n1 < 32768 ? n1 : 32768));
  sub_matrix_kernel<<<grid, 256, 0, ctx->stream>>>(A21, lda, Mws, n2, n2, n1); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  p.m = (int)n2; p.n = (int)n2; p.k = (int)n1; p.alpha = -1.0; p.beta = 1.0; p.C = A22; p.ldc = lda;
  p.A = A21; p.lda = lda; p.B = L21; p.ldb = ldl;
  EKB_TRY(gemm(ctx, GEMM_TB, p, /*tri_keep=*/1));
  p.A = L21; p.lda = ldl; p.B = A21; p.ldb = lda;
  EKB_TRY(gemm(ctx, GEMM_TB, p, /*tri_keep=*/1));
  sub_matrix_kernel<<<grid, 256, 0, ctx->stream>>>(A21, lda, Mws, n2, n2, n1); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  EKB_TRY(trsm_rec(ctx, TRSM_LLN, n2, n1, L22, ldl, inv2, A21, lda, tmp));
  return sygst_rec(ctx, n2, A22, lda, L22, ldl, inv2, tmp, Mws);
}

// A <- L^-1 A L^-T (pdsygst(1,'L')).  Reads the lower triangle of A; both triangles hold the result on exit.
int sygst_lower(Ctx* ctx, i64 n, double* A, i64 lda, const double* L, i64 ldl, const double* invd) {
  if (n <= 0) return 0;
  const i64 nblk = (n + NB - 1) / NB;
  const i64 h = ((nblk + 1) / 2) * NB;
  double *tmp = nullptr, *Mws = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&tmp, (size_t)round_up(n, 2) * NB * sizeof(double)));
  int rc = ctx_alloc(ctx, (void**)&Mws, (size_t)h * h * sizeof(double));
  if (!rc) rc = sygst_rec(ctx, n, A, lda, L, ldl, invd, tmp, Mws);
  if (!rc) rc = symmetrize_from_lower(ctx, A, lda, n);
  cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, tmp);
  if (Mws) ctx_free(ctx, Mws);
  return rc;
}

// ---------------------------------------------------------------- explicit-inverse reduction
// The ELPA-style workflow of the reference (src/solver_elpa_eigenexa.f90:110-150,189-190: cholesky ->
// invert_triangular -> hermitian_multiply -> pdtrmm, back-transformation by pdtrmm) with L = U^T:
//   X = L^-1 (trtri_lower);  A <- X A X^T (sygst_inverse);  Z <- X^T Z (trmm_lower_t).
// Everything is an engine GEMM; the zero triangle of X is skipped by cutting the k-range per block row / column.
constexpr i64 INV_BLK = 2048;

// X (full buffer, zero on entry) <- L^-1, lower triangular.  invd: the inverted 64x64 diagonal blocks of L.
// T: scratch of at least ceil(n/2)^2 doubles (rounded to whole 64-blocks).
static int trtri_rec(Ctx* ctx, i64 n, const double* L, i64 ldl, const double* invd, double* X, i64 ldx, double* T) {
  if (n <= 0) return 0;
  if (n <= NB) return copy_matrix(ctx, invd, NB, X, ldx, n, n);
  const i64 nblk = (n + NB - 1) / NB;
  const i64 n1 = ((nblk + 1) / 2) * NB, n2 = n - n1;
  const double* L21 = L + n1;
  const double* L22 = L + n1 * ldl + n1;
  const double* inv2 = invd + (n1 / NB) * NB * NB;
  double* X21 = X + n1;
  double* X22 = X + n1 * ldx + n1;
  EKB_TRY(trtri_rec(ctx, n1, L, ldl, invd, X, ldx, T));
  EKB_TRY(trtri_rec(ctx, n2, L22, ldl, inv2, X22, ldx, T));
  GemmP p;
  p.m = (int)n2; p.n = (int)n1; p.k = (int)n1;  // T = L21 X11
  p.A = L21; p.lda = ldl; p.B = X; p.ldb = ldx; p.C = T; p.ldc = n2;
  p.alpha = 1.0; p.beta = 0.0;
  EKB_TRY(gemm(ctx, 0, p));
  p.m = (int)n2; p.n = (int)n1; p.k = (int)n2;  // X21 = -X22 T
  p.A = X22; p.lda = ldx; p.B = T; p.ldb = n2; p.C = X21; p.ldc = ldx;
  p.alpha = -1.0; p.beta = 0.0;
  return gemm(ctx, 0, p);
}

int trtri_lower(Ctx* ctx, i64 n, const double* L, i64 ldl, const double* invd, double* X, i64 ldx) {
  if (n <= 0) return 0;
  const i64 h = (((n + NB - 1) / NB + 1) / 2) * NB;
  double* T = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&T, (size_t)h * h * sizeof(double)));
  int rc = set_zero(ctx, X, ldx, n, n);
  if (!rc) rc = trtri_rec(ctx, n, L, ldl, invd, X, ldx, T);
  cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, T);
  return rc;
}

// A <- X A X^T with X lower triangular (zero upper part).  A: full symmetric on entry, both triangles on exit.
// C: scratch n x n (ldc).  Block row I of C = X(I, 0:i1) A(0:i1, :); block column J of the lower triangle of the
// result = C(j0:n, 0:j1) X(J, 0:j1)^T  -- 4n^3/3 FLOPs instead of the 4n^3 of two full products.
int sygst_inverse(Ctx* ctx, i64 n, double* A, i64 lda, const double* X, i64 ldx, double* C, i64 ldc) {
  if (n <= 0) return 0;
  GemmP p;
  p.alpha = 1.0;
  p.beta = 0.0;
  for (i64 i0 = 0; i0 < n; i0 += INV_BLK) {
    const i64 i1 = i0 + INV_BLK < n ? i0 + INV_BLK : n;
    p.m = (int)(i1 - i0); p.n = (int)n; p.k = (int)i1;
    p.A = X + i0; p.lda = ldx; p.B = A; p.ldb = lda; p.C = C + i0; p.ldc = ldc;
    EKB_TRY(gemm(ctx, 0, p));
  }
  for (i64 j0 = 0; j0 < n; j0 += INV_BLK) {
    const i64 j1 = j0 + INV_BLK < n ? j0 + INV_BLK : n;
    p.m = (int)(n - j0); p.n = (int)(j1 - j0); p.k = (int)j1;
    p.A = C + j0; p.lda = ldc; p.B = X + j0; p.ldb = ldx; p.C = A + j0 * lda + j0; p.ldc = lda;
    EKB_TRY(gemm(ctx, GEMM_TB, p));
  }
  return symmetrize_from_lower(ctx, A, lda, n);
}

// Z (n x k) <- X^T Z with X lower triangular: block row I of the result = X(i0:n, I)^T Z(i0:n, :), which only needs
// rows >= i0 of the old Z, so the blocks are processed top-down through a (INV_BLK x k) buffer and stored in place.
int trmm_lower_t(Ctx* ctx, i64 n, i64 k, const double* X, i64 ldx, double* Z, i64 ldz) {
  if (n <= 0 || k <= 0) return 0;
  double* tmp = nullptr;
  const i64 rb = INV_BLK < n ? INV_BLK : n;
  EKB_TRY(ctx_alloc(ctx, (void**)&tmp, (size_t)round_up(rb, 8) * k * sizeof(double)));
  int rc = 0;
  GemmP p;
  p.alpha = 1.0;
  p.beta = 0.0;
  for (i64 i0 = 0; i0 < n && !rc; i0 += INV_BLK) {
    const i64 i1 = i0 + INV_BLK < n ? i0 + INV_BLK : n;
    p.m = (int)(i1 - i0); p.n = (int)k; p.k = (int)(n - i0);
    p.A = X + i0 * ldx + i0; p.lda = ldx; p.B = Z + i0; p.ldb = ldz; p.C = tmp; p.ldc = round_up(rb, 8);
    rc = gemm(ctx, GEMM_TA, p);
    if (!rc) rc = copy_matrix(ctx, tmp, round_up(rb, 8), Z + i0, ldz, i1 - i0, k);
  }
  cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, tmp);
  return rc;
}

}  // namespace ekb
