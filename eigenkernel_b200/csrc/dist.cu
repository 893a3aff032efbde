// Multi-GPU plumbing: one process (one context) per B200, NCCL over NVLink 5 / NVSwitch for the exchanges.
// This is where the reference's process grid + block-cyclic distribution (src/processes.f90:17-65,
// src/distribute_matrix.f90:92-148) maps onto the GPUs of one box:
//   * the eigenvector matrix is distributed by contiguous COLUMN SLABS (a 1 x P grid with one block per rank):
//     the back-transformations (pdormtr), the top D&C merge products and the final pdtrtrs act on the columns
//     of Z independently, so they need no data-path collective at all;
//   * the reduction to standard form (pdsygst) runs as two column-slab triangular solves with an all-gather
//     after each (the second one on the transposed row slab: C = L^-1 (L^-1 A)^T by symmetry);
//   * the dense-to-band reduction shards the SYMM / SYR2K trailing work by block columns with one panel
//     broadcast and one all-gather per panel (sy2sb.cu);
//   * Cholesky, bulge chasing and the lower D&C levels are replicated (identical bits on every rank: the
//     kernels are deterministic -- no floating-point atomics anywhere in the library).
// NCCL is bound at run time with dlopen (libnccl.so.2: the copy PyTorch already loaded, else the system one),
// so single-GPU users never need it.  Only ncclBroadcast / ncclAllGather-style exchanges are used.
#include <dlfcn.h>

#include "common.cuh"

namespace ekb {

// ---- the handful of NCCL entry points we bind (ABI of nccl.h 2.x; types reduced to what we pass)
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { NCCL_INT8 = 0, NCCL_FLOAT64 = 8 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*CommAbort)(NcclComm) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;
static std::string g_nccl_err;

static int nccl_load() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    g_nccl_err = std::string("dlopen(libnccl.so.2): ") + dlerror();
    return EKB_ERR_INTERNAL;
  }
  NcclApi a;
  a.lib = lib;
#define BIND(field, sym)                                              \
  *(void**)(&a.field) = dlsym(lib, sym);                              \
  if (!a.field) {                                                     \
    g_nccl_err = std::string("dlsym(") + sym + ") failed";            \
    return EKB_ERR_INTERNAL;                                          \
  }
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(CommAbort, "ncclCommAbort");
  BIND(Broadcast, "ncclBroadcast");
  BIND(AllGather, "ncclAllGather");
  BIND(AllReduce, "ncclAllReduce");
  BIND(GroupStart, "ncclGroupStart");
  BIND(GroupEnd, "ncclGroupEnd");
  BIND(GetErrorString, "ncclGetErrorString");
  BIND(GetVersion, "ncclGetVersion");
#undef BIND
  g_nccl = a;
  return 0;
}

#define EKB_NCCL(call)                                                                   \
  do {                                                                                   \
    int _r = (call);                                                                     \
    if (_r != 0) {                                                                       \
      ctx->last_error = std::string(#call) + ": " + g_nccl.GetErrorString(_r);           \
      return EKB_ERR_COMM;                                                               \
    }                                                                                    \
  } while (0)

int comm_unique_id(void* id128, std::string* err) {
  int rc = nccl_load();
  if (rc) {
    if (err) *err = g_nccl_err;
    return rc;
  }
  NcclUniqueId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) {
    if (err) *err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
    return EKB_ERR_COMM;
  }
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int comm_init(Ctx* ctx, int nranks, int rank, const void* id128) {
  if (ctx->comm) return EKB_ERR_INTERNAL;
  if (nranks == 1) {
    ctx->nranks = 1;
    ctx->rank = 0;
    return 0;
  }
  if (nccl_load()) {
    ctx->last_error = g_nccl_err;
    return EKB_ERR_COMM;
  }
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm comm = nullptr;
  EKB_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return 0;
}

int comm_destroy(Ctx* ctx) {
  if (ctx->comm && g_nccl.lib) {
    // every collective of this context has been waited for by its caller, so nothing is in flight; Abort (unlike
    // Destroy) does not need the peers to arrive, which keeps a failing rank from hanging at exit
    g_nccl.CommAbort((NcclComm)ctx->comm);
  }
  ctx->comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
  return 0;
}

// Every rank owns columns [bounds[r], bounds[r+1]) of the column-major matrix M (ld, same on all ranks);
// afterwards every rank holds all of them.  One grouped set of broadcasts (slabs may be uneven).
int comm_allgather_cols(Ctx* ctx, double* M, i64 ld, const std::vector<i64>& bounds) {
  if (ctx->nranks <= 1) return 0;
  EKB_TRY(prof_begin(ctx, PROF_NCCL, 8.0 * (double)ld * (double)(bounds[ctx->nranks] - bounds[0])));
  EKB_NCCL(g_nccl.GroupStart());
  for (int r = 0; r < ctx->nranks; ++r) {
    const i64 c0 = bounds[r], nc = bounds[r + 1] - bounds[r];
    if (nc <= 0) continue;
    double* p = M + c0 * ld;
    EKB_NCCL(g_nccl.Broadcast(p, p, (size_t)nc * ld, NCCL_FLOAT64, r, (NcclComm)ctx->comm, ctx->stream));
  }
  EKB_NCCL(g_nccl.GroupEnd());
  ctx->collectives++;
  EKB_TRY(prof_end(ctx));
  return 0;
}

int comm_bcast(Ctx* ctx, void* buf, size_t bytes, int root) {
  if (ctx->nranks <= 1 || bytes == 0) return 0;
  EKB_TRY(prof_begin(ctx, PROF_NCCL, (double)bytes));
  EKB_NCCL(g_nccl.Broadcast(buf, buf, bytes, NCCL_INT8, root, (NcclComm)ctx->comm, ctx->stream));
  ctx->collectives++;
  EKB_TRY(prof_end(ctx));
  return 0;
}

// buf (count doubles) <- sum over ranks, identical bits on every rank
int comm_allreduce_sum(Ctx* ctx, double* buf, size_t count) {
  if (ctx->nranks <= 1 || count == 0) return 0;
  EKB_TRY(prof_begin(ctx, PROF_NCCL, 8.0 * (double)count));
  EKB_NCCL(g_nccl.AllReduce(buf, buf, count, NCCL_FLOAT64, /*ncclSum*/ 0, (NcclComm)ctx->comm, ctx->stream));
  ctx->collectives++;
  EKB_TRY(prof_end(ctx));
  return 0;
}

// recv (nranks * count doubles) <- concatenation over ranks of send (count doubles each)
int comm_allgather(Ctx* ctx, const double* send, double* recv, size_t count) {
  if (ctx->nranks <= 1) {
    if (send != recv) EKB_CUDA(cudaMemcpyAsync(recv, send, count * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  EKB_NCCL(g_nccl.AllGather(send, recv, count, NCCL_FLOAT64, (NcclComm)ctx->comm, ctx->stream));
  ctx->collectives++;
  return 0;
}

// ------------------------------------------------------------------------------------------ transpose
// B (n x m, ldb) = A (m x n, lda)^T, 32x32 tiles through shared memory (both sides coalesced).
__global__ void transpose_kernel(const double* __restrict__ A, i64 lda, i64 m, i64 n, double* __restrict__ B, i64 ldb) {
  __shared__ double t[32][33];
  const i64 i0 = (i64)blockIdx.x * 32, j0 = (i64)blockIdx.y * 32;
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    const i64 i = i0 + threadIdx.x, j = j0 + c;
    t[c][threadIdx.x] = (i < m && j < n) ? A[j * lda + i] : 0.0;
  }
  __syncthreads();
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    const i64 j = j0 + threadIdx.x, i = i0 + c;  // B(j, i) = A(i, j)
    if (i < m && j < n) B[i * ldb + j] = t[threadIdx.x][c];
  }
}

int transpose_matrix(Ctx* ctx, const double* A, i64 lda, i64 m, i64 n, double* B, i64 ldb) {
  if (m <= 0 || n <= 0) return 0;
  dim3 grid(cdiv(m, 32), cdiv(n, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(A, lda, m, n, B, ldb); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ sharded pdsygst
// A <- L^-1 A L^-T over the ranks of the context (reference: pdsygst(1,'L'), generalized_to_standard.f90:37).
// Every rank holds the full symmetric A and the full L on entry and the full symmetric result on exit.
//   1. Y(:, J_p) = L^-1 A(:, J_p)                 column-slab TRSM, n^3/P FLOPs
//   2. all-gather the column slabs of Y
//   3. S = Y(J_p, :)^T  (= (A L^-T)(:, J_p));  C(:, J_p) = L^-1 S     column-slab TRSM, n^3/P FLOPs
//   4. all-gather the column slabs of C
int sygst_dist(Ctx* ctx, i64 n, double* A, i64 lda, const double* L, i64 ldl, const double* invd) {
  if (ctx->nranks <= 1) return sygst_lower(ctx, n, A, lda, L, ldl, invd);
  std::vector<i64> bounds;
  slab_bounds(n, ctx->nranks, 64, bounds);
  const i64 c0 = bounds[ctx->rank], nc = bounds[ctx->rank + 1] - c0;
  const i64 lds = round_up(n, 8);
  double* S = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&S, (size_t)lds * (nc > 0 ? nc : 1) * sizeof(double)));
  int rc = 0;
  if (nc > 0) rc = trsm_lower(ctx, TRSM_LLN, n, nc, L, ldl, invd, A + c0 * lda, lda);
  if (!rc) rc = comm_allgather_cols(ctx, A, lda, bounds);
  if (!rc && nc > 0) rc = transpose_matrix(ctx, A + c0, lda, nc, n, S, lds);
  if (!rc && nc > 0) rc = trsm_lower(ctx, TRSM_LLN, n, nc, L, ldl, invd, S, lds);
  if (!rc && nc > 0) rc = copy_matrix(ctx, S, lds, A + c0 * lda, lda, n, nc);
  if (!rc) rc = comm_allgather_cols(ctx, A, lda, bounds);
  cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, S);
  return rc;
}

}  // namespace ekb
