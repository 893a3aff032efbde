// Input construction and elementwise helpers (HBM-stream kernels).
// B200 twin of setup_distributed_matrix + distribute_global_sparse_matrix
// (reference src/distribute_matrix.f90:92-148, 401-422): zero-initialised dense matrix, COO entries
// scattered with symmetric mirroring; plus the deterministic synthetic generator of SURVEY.md §8(d).
#include "common.cuh"

namespace ekb {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// u(seed,i,j) in [-1,1), symmetric; identical bit pattern to oracle/lapack_twin.py:synthetic_u.
__global__ void fill_synth_kernel(double* __restrict__ A, i64 lda, i64 n, uint64_t seed, double offdiag_div,
                                  int diag_mode, double diag_value) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (i64 j = blockIdx.y; j < n; j += gridDim.y) {  // gridDim.y is capped at 32768 (n = 65536 exceeds the 65535 limit)
    uint64_t hi = i > j ? i : j, lo = i > j ? j : i;
    uint64_t z = splitmix64(seed ^ ((hi << 32) | lo));
    double u = (double)(z >> 11) * 0x1.0p-52 - 1.0;
    double v;
    if (i == j)
      v = diag_mode == 0 ? u + diag_value : diag_value;
    else
      v = u / offdiag_div;
    A[j * lda + i] = v;
  }
}

int fill_synthetic(Ctx* ctx, double* A, i64 lda, i64 n, uint64_t seed, double offdiag_div, int diag_mode,
                   double diag_value) {
  if (n <= 0) return 0;
  dim3 grid(cdiv(n, 256), (unsigned)(n < 32768 ? n : 32768));
  fill_synth_kernel<<<grid, 256, 0, ctx->stream>>>(A, lda, n, seed, offdiag_div, diag_mode, diag_value); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void set_zero_kernel(double* __restrict__ A, i64 lda, i64 m, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (i64 j = blockIdx.y; j < n; j += gridDim.y) A[j * lda + i] = 0.0;
}
int set_zero(Ctx* ctx, double* A, i64 lda, i64 m, i64 n) {
  if (m <= 0 || n <= 0) return 0;
  dim3 grid(cdiv(m, 256), (unsigned)(n < 32768 ? n : 32768));
  set_zero_kernel<<<grid, 256, 0, ctx->stream>>>(A, lda, m, n); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void copy_kernel(const double* __restrict__ A, i64 lda, double* __restrict__ B, i64 ldb, i64 m, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (i64 j = blockIdx.y; j < n; j += gridDim.y) B[j * ldb + i] = A[j * lda + i];
}
int copy_matrix(Ctx* ctx, const double* A, i64 lda, double* B, i64 ldb, i64 m, i64 n) {
  if (m <= 0 || n <= 0) return 0;
  dim3 grid(cdiv(m, 256), (unsigned)(n < 32768 ? n : 32768));
  copy_kernel<<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, m, n); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void identity_kernel(double* __restrict__ A, i64 lda, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (i64 j = blockIdx.y; j < n; j += gridDim.y) A[j * lda + i] = (i == j) ? 1.0 : 0.0;
}
int set_identity(Ctx* ctx, double* A, i64 lda, i64 n) {
  if (n <= 0) return 0;
  dim3 grid(cdiv(n, 256), (unsigned)(n < 32768 ? n : 32768));
  identity_kernel<<<grid, 256, 0, ctx->stream>>>(A, lda, n); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

// COO (1-based, Fortran suffix(2,nnz)) -> dense, mirrored when i != j.  The reference scatters with a sequential
// pdelset loop (distribute_matrix.f90:411-418): when an element occurs more than once -- repeated, or once as (i,j) and
// once as (j,i) -- the LAST entry of the file wins, for both triangles.  One thread per entry cannot just store: the
// result would depend on the order the threads happen to run in.  Three passes, deterministic:
//   1. every entry takes part in an atomicMax of (its index + 1) at the canonical position (max(i,j), min(i,j)) of
//      the zeroed matrix itself (positive doubles order like their bit patterns, so the matrix doubles as the index
//      table: no n x n workspace);
//   2. an entry whose index is the one that survived is the winner of its element (flag array);
//   3. the winners store their value at (i,j) and (j,i).
__global__ void coo_claim_kernel(double* __restrict__ A, i64 lda, i64 n, i64 nnz, const int32_t* __restrict__ ij) {
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  i64 i = ij[2 * t] - 1, j = ij[2 * t + 1] - 1;
  if (i < 0 || j < 0 || i >= n || j >= n) return;
  const i64 r = i > j ? i : j, c = i > j ? j : i;
  atomicMax(reinterpret_cast<unsigned long long*>(A + c * lda + r), (unsigned long long)__double_as_longlong((double)(t + 1)));
}
__global__ void coo_winner_kernel(const double* __restrict__ A, i64 lda, i64 n, i64 nnz, const int32_t* __restrict__ ij,
                                  unsigned char* __restrict__ win) {
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  i64 i = ij[2 * t] - 1, j = ij[2 * t + 1] - 1;
  unsigned char w = 0;
  if (!(i < 0 || j < 0 || i >= n || j >= n)) {
    const i64 r = i > j ? i : j, c = i > j ? j : i;
    w = A[c * lda + r] == (double)(t + 1);
  }
  win[t] = w;
}
__global__ void coo_store_kernel(double* __restrict__ A, i64 lda, i64 nnz, const int32_t* __restrict__ ij,
                                 const double* __restrict__ v, const unsigned char* __restrict__ win) {
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz || !win[t]) return;
  i64 i = ij[2 * t] - 1, j = ij[2 * t + 1] - 1;
  A[j * lda + i] = v[t];
  if (i != j) A[i * lda + j] = v[t];
}
// A must be zero on entry (ekb200_coo_to_dense zeroes it).
int coo_scatter(Ctx* ctx, double* A, i64 lda, i64 n, i64 nnz, const int32_t* d_ij, const double* d_v) {
  if (nnz <= 0) return 0;
  unsigned char* win = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&win, (size_t)nnz));
  const int blocks = cdiv(nnz, 256);
  coo_claim_kernel<<<blocks, 256, 0, ctx->stream>>>(A, lda, n, nnz, d_ij); EKB_COUNT_LAUNCH(ctx);
  coo_winner_kernel<<<blocks, 256, 0, ctx->stream>>>(A, lda, n, nnz, d_ij, win); EKB_COUNT_LAUNCH(ctx);
  coo_store_kernel<<<blocks, 256, 0, ctx->stream>>>(A, lda, nnz, d_ij, d_v, win); EKB_COUNT_LAUNCH(ctx);
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, win);
  if (ce != cudaSuccess) {
    ctx->last_cuda = ce;
    ctx->last_error = std::string("coo_scatter: ") + cudaGetErrorString(ce);
    return EKB_ERR_CUDA;
  }
  return 0;
}

// upper <- lower^T (32x32 smem transpose tiles)
__global__ void symmetrize_kernel(double* __restrict__ A, i64 lda, i64 n) {
  __shared__ double t[32][33];
  int bi = blockIdx.x, bj = blockIdx.y;
  if (bj > bi) return;
  i64 i = (i64)bi * 32 + threadIdx.x;
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    i64 j = (i64)bj * 32 + c;
    t[c][threadIdx.x] = (i < n && j < n) ? A[j * lda + i] : 0.0;
  }
  __syncthreads();
  // write A(j', i') = lower(i', j') for the mirrored tile
  i64 jj = (i64)bj * 32 + threadIdx.x;
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    i64 ii = (i64)bi * 32 + c;
    if (ii < n && jj < n && ii > jj) A[ii * lda + jj] = t[threadIdx.x][c];
  }
}
int symmetrize_from_lower(Ctx* ctx, double* A, i64 lda, i64 n) {
  if (n <= 0) return 0;
  dim3 grid(cdiv(n, 32), cdiv(n, 32));
  symmetrize_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(A, lda, n); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ekb
