// Flat C-ABI of libekb200 (declared in include/ekb200.h).
#include "../../include/ekb200.h"

#include <string.h>

#include <algorithm>

#include "common.cuh"

using namespace ekb;

struct ekb200_ctx {
  Ctx c;
  double* invd = nullptr;  // inverted 64x64 diagonal blocks of the current Cholesky factor
  i64 invd_n = 0;
  double merge_flops = 0.0;  // actual FLOPs of the D&C merge GEMMs of the last solve
};

namespace ekb {
int measure_fp64_peak(Ctx* ctx, double* dmma_tflops, double* dfma_tflops);
}

#define CHECK_CTX(h) \
  if (!(h)) return -1; \
  Ctx* ctx = &(h)->c; \
  cudaSetDevice(ctx->device);

static int ensure_invd(ekb200_ctx* h, i64 n) {
  Ctx* ctx = &h->c;
  if (h->invd_n >= n && h->invd) return 0;
  if (h->invd) ctx_free(ctx, h->invd);
  h->invd = nullptr;
  h->invd_n = 0;
  EKB_TRY(ctx_alloc(ctx, (void**)&h->invd, (size_t)cdiv(n, 64) * 64 * 64 * sizeof(double)));
  h->invd_n = n;
  return 0;
}

extern "C" {

int ekb200_version(void) { return 110; }

int ekb200_device_count(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return ndev;
}

int ekb200_create(ekb200_ctx** out, int device) {
  if (!out) return -1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return EKB_ERR_CUDA;
  if (device < 0 || device >= ndev) return -2;
  ekb200_ctx* h = new ekb200_ctx();
  Ctx* ctx = &h->c;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return EKB_ERR_CUDA; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return EKB_ERR_CUDA; }
  if (cudaMalloc((void**)&ctx->d_info, 64 * sizeof(int)) != cudaSuccess) { delete h; return EKB_ERR_NOMEM; }
  if (cudaMallocHost((void**)&ctx->h_info, 64 * sizeof(int)) != cudaSuccess) { delete h; return EKB_ERR_NOMEM; }
  *out = h;
  return 0;
}

int ekb200_destroy(ekb200_ctx* h) {
  if (!h) return 0;
  Ctx* ctx = &h->c;
  cudaSetDevice(ctx->device);
  if (ctx->comm == nullptr) cudaStreamSynchronize(ctx->stream);  // with a communicator a peer may never arrive
  comm_destroy(ctx);
  cudaStreamSynchronize(ctx->stream);
  for (void* p : ctx->allocs) cudaFree(p);
  ctx->allocs.clear();
  ctx->live.clear();
  ctx_trim(ctx);
  for (cudaEvent_t ev : ctx->prof_events) cudaEventDestroy(ev);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->d_info) cudaFree(ctx->d_info);
  if (ctx->h_info) cudaFreeHost(ctx->h_info);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  for (auto& e : ctx->aux_ev)
    if (e) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  delete h;
  return 0;
}

const char* ekb200_strerror(int info) {
  if (info == 0) return "ok";
  if (info < 0) return "illegal argument";
  if (info == EKB_ERR_CUDA) return "CUDA runtime failure";
  if (info == EKB_ERR_NOMEM) return "device memory allocation failed";
  if (info == EKB_ERR_INTERNAL) return "internal error";
  if (info == EKB_ERR_COMM) return "NCCL failure";
  return "numerical failure";
}

const char* ekb200_last_error(const ekb200_ctx* h) { return h ? h->c.last_error.c_str() : "null context"; }

int ekb200_set_option(ekb200_ctx* h, const char* key, int64_t value) {
  CHECK_CTX(h);
  if (!key) return -2;
  if (!strcmp(key, "band")) {
    if (value != 32 && value != 64) return -3;
    ctx->band = (int)value;
    return 0;
  }
  if (!strcmp(key, "select_method")) {  // -n solvers: 0 auto | 1 divide and conquer | 2 bisection + inverse iteration
    if (value < 0 || value > 2) return -3;
    ctx->select_method = (int)value;
    return 0;
  }
  if (!strcmp(key, "reduction")) {  // 0: blocked pdsygst-style reduction (default); 1: explicit inverse of L
    if (value != 0 && value != 1) return -3;
    ctx->reduction = (int)value;
    return 0;
  }
  if (!strcmp(key, "sb2st_variant")) {  // 1: register-resident lag-2 bulge chasing (default); 0: round-1 kernel
    if (value != 0 && value != 1) return -3;
    ctx->sb2st_variant = (int)value;
    return 0;
  }
  if (!strcmp(key, "sb2st_warps")) {  // warps per CTA of the register-resident bulge-chasing kernel
    if (value != 8 && value != 16) return -3;
    ctx->sb2st_warps = (int)value;
    return 0;
  }
  if (!strcmp(key, "sb2st_rwarp")) {  // dedicated reflector warp of the bulge-chasing kernel (tuning)
    if (value != 0 && value != 1) return -3;
    ctx->sb2st_rwarp = (int)value;
    return 0;
  }
  if (!strcmp(key, "sb2st_cps")) {  // cap on resident bulge-chasing CTAs per SM (tuning; 0 = no cap)
    if (value < 0 || value > 4) return -3;
    ctx->sb2st_cps = (int)value;
    return 0;
  }
  if (!strcmp(key, "out_block")) {  // block size NB of the caller's 1 x P block-cyclic eigenvector descriptor (0: slabs)
    if (value < 0) return -3;
    ctx->out_block = value;
    return 0;
  }
  if (!strcmp(key, "stedc_shard")) {  // shard the level-1 merges of the divide and conquer over the ranks (tuning)
    if (value != 0 && value != 1) return -3;
    ctx->stedc_shard = (int)value;
    return 0;
  }
  if (!strcmp(key, "gemm_autosplit")) {  // split-K factor of TMA-fed products chosen by the round-count model (tuning)
    if (value != 0 && value != 1) return -3;
    ctx->gemm_autosplit = (int)value;
    return 0;
  }
  if (!strcmp(key, "gemm_bulk")) {  // TMA-fed warp-specialised GEMM kernel for the big-tile products (tuning)
    if (value != 0 && value != 1) return -3;
    ctx->gemm_bulk = (int)value;
    return 0;
  }
  if (!strcmp(key, "panel_qr_variant")) {  // 1 shared-memory-resident panel QR (default) | 0 global-memory kernel
    if (value != 0 && value != 1) return -3;
    ctx->panel_qr_variant = (int)value;
    return 0;
  }
  if (!strcmp(key, "sy2sb_lookahead")) {  // panel look-ahead of the dense-to-band reduction (tuning; default 1)
    if (value != 0 && value != 1) return -3;
    ctx->sy2sb_lookahead = (int)value;
    return 0;
  }
  if (!strcmp(key, "q2_kc")) {
    if (value != 0 && value != 1004 && value != 1008 && value != 1012 && value != 1014 && (value < 64 || value > 128 || value % 16))
      return -3;
    ctx->q2_kc = (int)value;
    return 0;
  }
  if (!strcmp(key, "cache_device_memory")) {  // 0: release blocks to the driver at once; 1 (default): caching arena
    ctx->cache_enabled = value != 0;
    if (!ctx->cache_enabled) {
      cudaStreamSynchronize(ctx->stream);
      ctx_trim(ctx);
    }
    return 0;
  }
  if (!strcmp(key, "profile_gemm")) {
    ctx->profile_gemm = value != 0;
    ctx->prof_used = 0;
    ctx->prof_flops.clear();
    ctx->prof_family.clear();
    ctx->prof_stage.clear();
    ctx->prof_shape.clear();
    return 0;
  }
  return -2;
}

int ekb200_num_events(const ekb200_ctx* h) { return h ? (int)h->c.events.size() : 0; }
int ekb200_get_event(const ekb200_ctx* h, int i, const char** name, double* seconds, int* num_repeated) {
  if (!h) return -1;
  if (i < 0 || i >= (int)h->c.events.size()) return -2;
  const Event& e = h->c.events[i];
  if (name) *name = e.name.c_str();
  if (seconds) *seconds = e.seconds;
  if (num_repeated) *num_repeated = e.num_repeated;
  return 0;
}
int ekb200_clear_events(ekb200_ctx* h) {
  if (!h) return -1;
  h->c.events.clear();
  return 0;
}

int ekb200_dev_alloc(ekb200_ctx* h, int64_t bytes, void** p) {
  CHECK_CTX(h);
  if (bytes < 0) return -2;
  if (!p) return -3;
  return ctx_alloc(ctx, p, (size_t)bytes);
}
int ekb200_dev_free(ekb200_ctx* h, void* p) {
  CHECK_CTX(h);
  cudaStreamSynchronize(ctx->stream);
  return ctx_free(ctx, p);
}
int ekb200_h2d(ekb200_ctx* h, void* dst, const void* src, int64_t bytes) {
  CHECK_CTX(h);
  if (bytes < 0) return -4;
  EKB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int ekb200_d2h(ekb200_ctx* h, void* dst, const void* src, int64_t bytes) {
  CHECK_CTX(h);
  if (bytes < 0) return -4;
  EKB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int ekb200_h2d_matrix(ekb200_ctx* h, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t m, int64_t n) {
  CHECK_CTX(h);
  if (m < 0 || n < 0) return -6;
  if (m == 0 || n == 0) return 0;
  EKB_CUDA(cudaMemcpy2DAsync(dst, ldd * 8, src, lds * 8, m * 8, n, cudaMemcpyHostToDevice, ctx->stream));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int ekb200_d2h_matrix(ekb200_ctx* h, double* dst, int64_t ldh, const double* src, int64_t ldd, int64_t m, int64_t n) {
  CHECK_CTX(h);
  if (m < 0 || n < 0) return -6;
  if (m == 0 || n == 0) return 0;
  EKB_CUDA(cudaMemcpy2DAsync(dst, ldh * 8, src, ldd * 8, m * 8, n, cudaMemcpyDeviceToHost, ctx->stream));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int ekb200_sync(ekb200_ctx* h) {
  CHECK_CTX(h);
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int ekb200_coo_to_dense(ekb200_ctx* h, int64_t n, int64_t nnz, const int32_t* host_ij, const double* host_v,
                        double* dev_A, int64_t lda) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nnz < 0) return -3;
  if (lda < n) return -7;
  EKB_TRY(set_zero(ctx, dev_A, lda, n, n));
  if (nnz == 0) return 0;
  // "last duplicate wins", symmetric (distribute_matrix.f90:411-418 executes pdelset sequentially): resolved
  // deterministically on the device by coo_scatter (fill.cu)
  int32_t* d_ij = nullptr;
  double* d_v = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    ctx_free(ctx, d_ij);
    ctx_free(ctx, d_v);
  };
  int rc = ctx_alloc(ctx, (void**)&d_ij, (size_t)nnz * 2 * sizeof(int32_t));
  if (!rc) rc = ctx_alloc(ctx, (void**)&d_v, (size_t)nnz * sizeof(double));
  if (!rc) {
    cudaError_t ce = cudaMemcpyAsync(d_ij, host_ij, (size_t)nnz * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(d_v, host_v, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (ce != cudaSuccess) {
      ctx->last_cuda = ce;
      ctx->last_error = std::string("ekb200_coo_to_dense: ") + cudaGetErrorString(ce);
      rc = EKB_ERR_CUDA;
    }
  }
  if (!rc) rc = coo_scatter(ctx, dev_A, lda, n, nnz, d_ij, d_v);
  cleanup();
  return rc;
}

int ekb200_fill_synthetic(ekb200_ctx* h, int64_t n, uint64_t seed, double offdiag_div, int diag_mode,
                          double diag_value, double* dev_A, int64_t lda) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (lda < n) return -8;
  return fill_synthetic(ctx, dev_A, lda, n, seed, offdiag_div, diag_mode, diag_value);
}

int ekb200_dgemm(ekb200_ctx* h, char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
  CHECK_CTX(h);
  int flags = 0;
  if (transa == 'T' || transa == 't') flags |= GEMM_TA;
  else if (!(transa == 'N' || transa == 'n')) return -2;
  if (transb == 'T' || transb == 't') flags |= GEMM_TB;
  else if (!(transb == 'N' || transb == 'n')) return -3;
  if (m < 0) return -4;
  if (n < 0) return -5;
  if (k < 0) return -6;
  GemmP p;
  p.m = (int)m; p.n = (int)n; p.k = (int)k;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.alpha = alpha; p.beta = beta;
  return gemm(ctx, flags, p);
}

int ekb200_potrf(ekb200_ctx* h, int64_t n, double* B, int64_t ldb) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (ldb < n) return -4;
  if (n == 0) return 0;
  EKB_TRY(ensure_invd(h, n));
  return potrf_lower(ctx, n, B, ldb, h->invd);
}

int ekb200_sygst(ekb200_ctx* h, int64_t n, double* A, int64_t lda, const double* L, int64_t ldl) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (lda < n) return -4;
  if (ldl < n) return -6;
  if (n == 0) return 0;
  EKB_TRY(ensure_invd(h, n));
  EKB_TRY(trtri_diag_blocks(ctx, n, L, ldl, h->invd));
  EKB_TRY(sygst_dist(ctx, n, A, lda, L, ldl, h->invd));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int ekb200_trtrs_lt(ekb200_ctx* h, int64_t n, int64_t nrhs, const double* L, int64_t ldl, double* Z, int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nrhs < 0) return -3;
  if (ldl < n) return -5;
  if (ldz < n) return -7;
  if (n == 0 || nrhs == 0) return 0;
  EKB_TRY(ensure_invd(h, n));
  EKB_TRY(trtri_diag_blocks(ctx, n, L, ldl, h->invd));
  EKB_TRY(trsm_lower(ctx, TRSM_LLT, n, nrhs, L, ldl, h->invd, Z, ldz));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int ekb200_get_band(const ekb200_ctx* h) { return h ? h->c.band : 0; }
int ekb200_sy2sb_num_panels(const ekb200_ctx* h, int64_t n) { return h ? sy2sb_num_panels(n, h->c.band) : 0; }

int ekb200_sy2sb(ekb200_ctx* h, int64_t n, double* A, int64_t lda, double* AB, int64_t ldab, double* T1) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (lda < n) return -4;
  if (ldab < 2 * ctx->band) return -6;
  if (n == 0) return 0;
  double* work = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&work, sy2sb_workspace_doubles(n, ctx->band, ctx->num_sms) * sizeof(double)));
  int rc = sy2sb(ctx, n, ctx->band, A, lda, AB, ldab, T1, work);
  cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, work);
  if (rc == 0) EKB_CUDA(cudaGetLastError());
  return rc;
}

int ekb200_sb2st_max_tasks(const ekb200_ctx* h, int64_t n) { return h ? sb2st_max_tasks(n, h->c.band) : 0; }

int ekb200_sb2st(ekb200_ctx* h, int64_t n, double* AB, int64_t ldab, double* V2, int64_t ldv, double* TAU2,
                 int64_t ldtau, double* d, double* e) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (ldab < 2 * ctx->band) return -4;
  if (ldv < n) return -6;
  if (ldtau < sb2st_max_tasks(n, ctx->band)) return -8;
  if (n == 0) return 0;
  int* prog = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&prog, (size_t)(n + 8) * sizeof(int)));
  int rc = sb2st(ctx, n, ctx->band, AB, ldab, V2, ldv, TAU2, (int)ldtau, prog, d, e);
  cudaError_t ce = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, prog);
  if (rc == 0) EKB_CUDA(ce);
  return rc;
}

int ekb200_stedc(ekb200_ctx* h, int64_t n, double* d, double* e, double* w, double* Z, int64_t ldz, double* merge_flops) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (ldz < n) return -7;
  if (n == 0) return 0;
  void* work = nullptr;
  EKB_TRY(ctx_alloc(ctx, &work, stedc_workspace_bytes(n)));
  int rc = stedc(ctx, n, d, e, w, Z, ldz, work, merge_flops, 0, n);
  cudaError_t ce = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, work);
  if (rc == 0) EKB_CUDA(ce);
  return rc;
}

int ekb200_stebz_stein(ekb200_ctx* h, int64_t n, int64_t nev, const double* d, const double* e, double* w, double* Z,
                       int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nev < 0 || nev > n) return -3;
  if (n > 0 && (!d || (n > 1 && !e))) return -4;
  if (n > 0 && !w) return -6;
  if (nev > 0 && !Z) return -7;
  if (ldz < n) return -8;
  if (n == 0) return 0;
  void* work = nullptr;
  EKB_TRY(ctx_alloc(ctx, &work, stebz_stein_workspace_bytes(n, ctx->num_sms)));
  int rc = stebz_stein(ctx, n, d, e, w, nev, 0, nev, Z, ldz, work);
  cudaError_t ce = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, work);
  if (rc == 0) EKB_CUDA(ce);
  return rc;
}

int ekb200_apply_q2(ekb200_ctx* h, int64_t n, int64_t nrhs, const double* V2, int64_t ldv, const double* TAU2,
                    int64_t ldtau, double* Z, int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nrhs < 0) return -3;
  if (ldv < n) return -5;
  if (ldtau < sb2st_max_tasks(n, ctx->band)) return -7;
  if (ldz < n) return -9;
  if (n == 0 || nrhs == 0) return 0;
  return apply_q2(ctx, n, ctx->band, V2, ldv, TAU2, (int)ldtau, nrhs, Z, ldz);
}

int ekb200_apply_q1(ekb200_ctx* h, int64_t n, int64_t nrhs, double* A, int64_t lda, const double* T1, double* Z,
                    int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nrhs < 0) return -3;
  if (lda < n) return -5;
  if (ldz < n) return -8;
  if (n == 0 || nrhs == 0) return 0;
  double* work = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&work, apply_q1_workspace_doubles(n, ctx->band, nrhs) * sizeof(double)));
  int rc = apply_q1(ctx, n, ctx->band, A, lda, T1, nrhs, Z, ldz, work);
  cudaError_t ce = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, work);
  if (rc == 0) EKB_CUDA(ce);
  return rc;
}

int ekb200_host_alloc(ekb200_ctx* h, int64_t bytes, void** p) {
  CHECK_CTX(h);
  if (bytes < 0) return -2;
  if (!p) return -3;
  *p = nullptr;
  cudaError_t e = cudaHostAlloc(p, (size_t)(bytes > 0 ? bytes : 16), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    ctx->last_cuda = e;
    ctx->last_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return EKB_ERR_NOMEM;
  }
  return 0;
}
int ekb200_host_free(ekb200_ctx* h, void* p) {
  CHECK_CTX(h);
  if (p) cudaFreeHost(p);
  return 0;
}

int ekb200_syevd_dev(ekb200_ctx* h, int64_t n, int64_t nev, double* A, int64_t lda, double* w, double* Z,
                     int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nev < 0 || nev > n) return -3;
  if (lda < n) return -5;
  if (ldz < n) return -8;
  h->merge_flops = 0.0;
  return syevd_dev(ctx, n, nev, A, lda, w, Z, ldz, &h->merge_flops);
}

int ekb200_sygvd_dev(ekb200_ctx* h, int64_t n, int64_t nev, double* A, int64_t lda, double* B, int64_t ldb, double* w,
                     double* Z, int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nev < 0 || nev > n) return -3;
  if (lda < n) return -5;
  if (ldb < n) return -7;
  if (ldz < n) return -10;
  if (n == 0) return 0;
  EKB_TRY(ensure_invd(h, n));
  h->merge_flops = 0.0;
  return sygvd_dev(ctx, n, nev, A, lda, B, ldb, w, Z, ldz, h->invd, &h->merge_flops);
}

int ekb200_timer_start(ekb200_ctx* h) {
  CHECK_CTX(h);
  if (!ctx->ev0) EKB_CUDA(cudaEventCreate(&ctx->ev0));
  if (!ctx->ev1) EKB_CUDA(cudaEventCreate(&ctx->ev1));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  EKB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return 0;
}
int ekb200_timer_stop(ekb200_ctx* h, double* seconds) {
  CHECK_CTX(h);
  if (!seconds) return -2;
  if (!ctx->ev0 || !ctx->ev1) return -1;
  EKB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  EKB_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  EKB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *seconds = ms * 1e-3;
  return 0;
}

int ekb200_gemm_profile(ekb200_ctx* h, double* seconds, double* flops, int64_t* launches) {
  CHECK_CTX(h);
  if (!seconds) return -2;
  if (!flops) return -3;
  if (!launches) return -4;
  long long l = 0;
  int rc = gemm_profile_collect(ctx, seconds, flops, &l);
  *launches = l;
  return rc;
}
int ekb200_kernel_profile(ekb200_ctx* h, double* seconds, double* work, int64_t* launches) {
  CHECK_CTX(h);
  if (!seconds) return -2;
  if (!work) return -3;
  if (!launches) return -4;
  long long l[PROF_FAMILIES];
  int rc = profile_collect(ctx, seconds, work, l);
  for (int f = 0; f < PROF_FAMILIES; ++f) launches[f] = l[f];
  return rc;
}
int ekb200_profile_rows(const ekb200_ctx* h) { return h ? (int)h->c.prof_table.size() : 0; }
int ekb200_profile_row(const ekb200_ctx* h, int i, const char** stage, int* family, double* seconds, double* work,
                       int64_t* launches) {
  if (!h) return -1;
  if (i < 0 || i >= (int)h->c.prof_table.size()) return -2;
  const Ctx::ProfRow& r = h->c.prof_table[i];
  if (stage) *stage = r.stage.c_str();
  if (family) *family = r.family;
  if (seconds) *seconds = r.seconds;
  if (work) *work = r.work;
  if (launches) *launches = r.launches;
  return 0;
}
int64_t ekb200_num_launches(const ekb200_ctx* h) { return h ? h->c.launches : 0; }

double ekb200_last_merge_flops(const ekb200_ctx* h) { return h ? h->merge_flops : 0.0; }

// Shared body of the host-pointer front doors.  Exactly one of (A, cooA) describes A; B/cooB likewise
// (generalized iff hasB).  Host matrices: lower triangles referenced (the reference passes 'L' everywhere).
struct CooIn {
  int64_t nnz;
  const int32_t* ij;
  const double* v;
};
static int solve_host(ekb200_ctx* h, int64_t n, int64_t nev, const double* A, int64_t lda, const CooIn* cooA, bool hasB,
                      const double* B, int64_t ldb, const CooIn* cooB, double* w, double* Z, int64_t ldz) {
  Ctx* ctx = &h->c;
  if (n == 0 || nev == 0) return 0;
  const i64 ld = round_up(n, 8);
  double *dA = nullptr, *dB = nullptr, *dZ = nullptr, *dw = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    ctx_free(ctx, dA); ctx_free(ctx, dB); ctx_free(ctx, dZ); ctx_free(ctx, dw);
  };
  int rc = ctx_alloc(ctx, (void**)&dA, (size_t)ld * n * 8);
  if (!rc && hasB) rc = ctx_alloc(ctx, (void**)&dB, (size_t)ld * n * 8);
  if (!rc) rc = ctx_alloc(ctx, (void**)&dZ, (size_t)ld * nev * 8);
  if (!rc) rc = ctx_alloc(ctx, (void**)&dw, (size_t)(n + 8) * 8);
  if (rc) { cleanup(); return rc; }
  HostOverlap ov;
  {
    StageTimer t(ctx, hasB ? "solve_with_general_b200:setup_matrices" : "eigen_solver_b200:setup_matrices");
    cudaError_t ce = cudaSuccess;
    // Dense host matrices on P > 1 ranks: every rank holds the same (replicated) arrays, so each uploads only ITS
    // block of columns over PCIe (n^2 / P elements per matrix) and the blocks are all-gathered over NVLink -- the
    // counterpart of distribute_matrix.f90:92-148, where a process only ever touches its n^2 / P piece.
    std::vector<i64> ub;
    slab_bounds(n, ctx->nranks, 128, ub);
    const i64 u0 = ub[ctx->rank], uk = ub[ctx->rank + 1] - ub[ctx->rank];
    auto upload = [&](const double* H, i64 ldh, double* D) -> int {
      if (ctx->nranks == 1) {
        ce = cudaMemcpy2DAsync(D, ld * 8, H, ldh * 8, n * 8, n, cudaMemcpyHostToDevice, ctx->stream);
      } else {
        if (uk > 0)
          ce = cudaMemcpy2DAsync(D + u0 * ld, ld * 8, H + u0 * ldh, ldh * 8, n * 8, uk, cudaMemcpyHostToDevice, ctx->stream);
        if (ce == cudaSuccess) EKB_TRY(comm_allgather_cols(ctx, D, ld, ub));
      }
      if (ce != cudaSuccess) return 0;
      return symmetrize_from_lower(ctx, D, ld, n);
    };
    // One rank, generalized, dense host inputs: B goes first (the Cholesky factorization needs it), A follows on the
    // side stream and lands while B is being factored (0.16 s of PCIe time at n = 32768 hidden behind 0.5 s of potrf);
    // the solver waits for ov.a_ready before it first touches A.
    const bool a_beside = hasB && !cooA && !cooB && ctx->nranks == 1 && n >= 4096 && ctx_ensure_aux(ctx) == 0;
    if (a_beside) {
      rc = upload(B, ldb, dB);
      if (!rc && ce == cudaSuccess) ce = cudaEventRecord(ctx->aux_ev[1], ctx->stream);
      if (!rc && ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[1], 0);
      if (!rc && ce == cudaSuccess) {
        cudaStream_t main_stream = ctx->stream;
        ctx->stream = ctx->aux_stream;  // upload() and its symmetrize kernel go to the side stream
        rc = upload(A, lda, dA);
        if (!rc && ce == cudaSuccess) ce = cudaEventRecord(ctx->aux_ev[0], ctx->stream);
        ctx->stream = main_stream;
        ov.a_ready = ctx->aux_ev[0];
      }
    } else {
      if (cooA) rc = ekb200_coo_to_dense(h, n, cooA->nnz, cooA->ij, cooA->v, dA, ld);
      else rc = upload(A, lda, dA);
      if (!rc && ce == cudaSuccess && hasB) {
        if (cooB) rc = ekb200_coo_to_dense(h, n, cooB->nnz, cooB->ij, cooB->v, dB, ld);
        else rc = upload(B, ldb, dB);
      }
    }
    if (ce != cudaSuccess) {
      ctx->last_cuda = ce;
      ctx->last_error = std::string("h2d: ") + cudaGetErrorString(ce);
      rc = EKB_ERR_CUDA;
    }
    t.stop();
  }
  if (!rc) {
    h->merge_flops = 0.0;
    if (hasB) {
      rc = ensure_invd(h, n);
      if (!(ctx->nranks > 1 && ctx->out_block > 0)) {  // the slab goes to the caller's array chunk by chunk, beside the solve
        ov.host_Z = Z;
        ov.ld_host_Z = ldz;
      }
      if (!rc) rc = sygvd_dev(ctx, n, nev, dA, ld, dB, ld, dw, dZ, ld, h->invd, &h->merge_flops, &ov);
      if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);  // chunk downloads (and, on errors, a pending upload)
    } else {
      rc = syevd_dev(ctx, n, nev, dA, ld, dw, dZ, ld, &h->merge_flops);
    }
  }
  int warn = 0;
  if (rc > EKB_WARN_STEIN && rc < EKB_FAIL_STEDC) {  // unconverged inverse iteration: a warning, the results are complete
    warn = rc;
    rc = 0;
  }
  if (!rc) {
    StageTimer t(ctx, hasB ? "solve_with_general_b200:d2h" : "eigen_solver_b200:d2h");
    // multi-rank: every rank returns all of w and ITS column slab of the eigenvectors; the caller's Z is then
    // the LOCAL piece (n x nloc, ekb200_comm_slab), like blacs%Vectors(lld, loc_cols) of the reference
    std::vector<i64> zb;
    slab_bounds(nev, ctx->nranks, 128, zb);
    const i64 c0 = zb[ctx->rank], kc = zb[ctx->rank + 1] - zb[ctx->rank];
    cudaError_t ce = cudaMemcpyAsync(w, dw, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (ctx->nranks > 1 && ctx->out_block > 0) {
      // the caller's descriptor is block-cyclic with block size out_block (layout.h): every rank gets all columns over
      // NVLink (one all-gather of the slabs), then copies out the blocks r, r + P, ... it owns
      int rc2 = ce == cudaSuccess ? comm_allgather_cols(ctx, dZ, ld, zb) : 0;
      if (rc2) { t.stop(); cleanup(); return rc2; }
      const i64 nb = ctx->out_block;
      const i64 nloc = numroc0(nev, nb, ctx->rank, ctx->nranks);
      for (i64 lc = 0; lc < nloc && ce == cudaSuccess; lc += nb) {
        const i64 g = cyclic_global_col0(lc, nb, ctx->nranks, ctx->rank);
        const i64 wdt = std::min(nb, nloc - lc);
        ce = cudaMemcpy2DAsync(Z + lc * ldz, ldz * 8, dZ + g * ld, ld * 8, n * 8, wdt, cudaMemcpyDeviceToHost, ctx->stream);
      }
    } else if (ce == cudaSuccess && kc > 0 && !ov.z_downloaded) {
      ce = cudaMemcpy2DAsync(Z, ldz * 8, dZ + c0 * ld, ld * 8, n * 8, kc, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    t.stop();
    if (ce != cudaSuccess) {
      ctx->last_cuda = ce;
      ctx->last_error = std::string("d2h: ") + cudaGetErrorString(ce);
      rc = EKB_ERR_CUDA;
    }
  }
  cleanup();
  return rc ? rc : warn;
}

int ekb200_syevd(ekb200_ctx* h, int64_t n, int64_t nev, const double* A, int64_t lda, double* w, double* Z,
                 int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nev < 0 || nev > n) return -3;
  if (n > 0 && !A) return -4;
  if (lda < n) return -5;
  if (n > 0 && !w) return -6;
  if (nev > 0 && !Z) return -7;
  if (ldz < n) return -8;
  return solve_host(h, n, nev, A, lda, nullptr, false, nullptr, 0, nullptr, w, Z, ldz);
}

int ekb200_sygvd(ekb200_ctx* h, int64_t n, int64_t nev, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double* w, double* Z, int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nev < 0 || nev > n) return -3;
  if (n > 0 && !A) return -4;
  if (lda < n) return -5;
  if (n > 0 && !B) return -6;
  if (ldb < n) return -7;
  if (n > 0 && !w) return -8;
  if (nev > 0 && !Z) return -9;
  if (ldz < n) return -10;
  return solve_host(h, n, nev, A, lda, nullptr, true, B, ldb, nullptr, w, Z, ldz);
}

int ekb200_sygvd_coo(ekb200_ctx* h, int64_t n, int64_t nev, int64_t nnzA, const int32_t* ijA, const double* vA,
                     int64_t nnzB, const int32_t* ijB, const double* vB, double* w, double* Z, int64_t ldz) {
  CHECK_CTX(h);
  if (n < 0) return -2;
  if (nev < 0 || nev > n) return -3;
  if (nnzA < 0) return -4;
  if (nnzA > 0 && (!ijA || !vA)) return -5;
  if (nnzB < 0) return -7;
  if (nnzB > 0 && (!ijB || !vB)) return -8;
  if (n > 0 && !w) return -10;
  if (nev > 0 && !Z) return -11;
  if (ldz < n) return -12;
  CooIn a{nnzA, ijA, vA}, b{nnzB, ijB, vB};
  return solve_host(h, n, nev, nullptr, 0, &a, nnzB > 0, nullptr, 0, nnzB > 0 ? &b : nullptr, w, Z, ldz);
}

// ---- acceptance metrics and IPRs (verify.cu)
int ekb200_eval_residual_norm_dev(ekb200_ctx* h, int64_t n, int64_t ncheck, const double* A, int64_t lda, const double* B,
                                  int64_t ldb, const double* w, const double* X, int64_t ldx, double* A_norm,
                                  double* res_norm_ave, double* res_norm_max) {
  CHECK_CTX(h);
  if (n <= 0) return -2;
  if (ncheck <= 0 || ncheck > n) return -3;
  if (!A) return -4;
  if (lda < n) return -5;
  if (B && ldb < n) return -7;
  if (!w) return -8;
  if (!X) return -9;
  if (ldx < n) return -10;
  return eval_residual_norm(ctx, n, ncheck, A, lda, B, ldb, w, X, ldx, A_norm, res_norm_ave, res_norm_max);
}
int ekb200_eval_orthogonality_dev(ekb200_ctx* h, int64_t n, int64_t index1, int64_t index2, const double* X, int64_t ldx,
                                  const double* B, int64_t ldb, double* orthogonality) {
  CHECK_CTX(h);
  if (n <= 0) return -2;
  if (index1 < 1) return -3;
  if (index2 < index1 || index2 > n) return -4;
  if (!X) return -5;
  if (ldx < n) return -6;
  if (B && ldb < n) return -8;
  return eval_orthogonality(ctx, n, index1, index2, X, ldx, B, ldb, orthogonality);
}
int ekb200_eval_b_orthonormality_dev(ekb200_ctx* h, int64_t n, int64_t index1, int64_t index2, const double* X,
                                     int64_t ldx, const double* B, int64_t ldb, double* orthogonality,
                                     double* gram_minus_identity) {
  CHECK_CTX(h);
  if (n <= 0) return -2;
  if (index1 < 1) return -3;
  if (index2 < index1 || index2 > n) return -4;
  if (!X) return -5;
  if (ldx < n) return -6;
  if (B && ldb < n) return -8;
  return eval_orthogonality(ctx, n, index1, index2, X, ldx, B, ldb, orthogonality, gram_minus_identity);
}
int ekb200_get_ipratios_dev(ekb200_ctx* h, int64_t n, int64_t nvec, const double* X, int64_t ldx, const double* B,
                            int64_t ldb, double* ipratios) {
  CHECK_CTX(h);
  if (n <= 0) return -2;
  if (nvec <= 0 || nvec > n) return -3;
  if (!X) return -4;
  if (ldx < n) return -5;
  if (B && ldb < n) return -7;
  if (!ipratios) return -8;
  return get_ipratios(ctx, n, nvec, X, ldx, B, ldb, ipratios);
}

// Host-side eigenpairs + replicated COO matrices, as the reference routines receive them.  `what`: 0 residual,
// 1 orthogonality, 2 IPR.  X is the rank's LOCAL piece (n x nloc of the nvec eigenvector columns).
static int verify_host(ekb200_ctx* h, int what, int64_t n, int64_t nvec, int64_t a1, int64_t a2, int64_t nnzA,
                       const int32_t* ijA, const double* vA, int64_t nnzB, const int32_t* ijB, const double* vB,
                       const double* w, const double* X, int64_t ldx, double* o1, double* o2, double* o3) {
  Ctx* ctx = &h->c;
  const i64 ld = round_up(n, 8);
  double *dA = nullptr, *dB = nullptr, *dX = nullptr, *dw = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    ctx_free(ctx, dA); ctx_free(ctx, dB); ctx_free(ctx, dX); ctx_free(ctx, dw);
  };
  int rc = 0;
  if (what == 0) rc = ctx_alloc(ctx, (void**)&dA, (size_t)ld * n * 8);
  if (!rc && nnzB > 0) rc = ctx_alloc(ctx, (void**)&dB, (size_t)ld * n * 8);
  if (!rc) rc = ctx_alloc(ctx, (void**)&dX, (size_t)ld * nvec * 8);
  if (!rc && what == 0) rc = ctx_alloc(ctx, (void**)&dw, (size_t)(n + 8) * 8);
  if (!rc && what == 0) rc = ekb200_coo_to_dense(h, n, nnzA, ijA, vA, dA, ld);
  if (!rc && nnzB > 0) rc = ekb200_coo_to_dense(h, n, nnzB, ijB, vB, dB, ld);
  if (!rc) {
    std::vector<i64> zb;
    slab_bounds(nvec, ctx->nranks, 128, zb);
    const i64 c0 = zb[ctx->rank], kc = zb[ctx->rank + 1] - zb[ctx->rank];
    cudaError_t ce = cudaSuccess;
    const bool cyclic = ctx->nranks > 1 && ctx->out_block > 0;  // X is the block-cyclic local piece (option "out_block")
    if (cyclic) {
      rc = set_zero(ctx, dX, ld, ld, nvec);
      const i64 nb = ctx->out_block, nloc = numroc0(nvec, nb, ctx->rank, ctx->nranks);
      for (i64 lc = 0; lc < nloc && ce == cudaSuccess && !rc; lc += nb) {
        const i64 g = cyclic_global_col0(lc, nb, ctx->nranks, ctx->rank);
        ce = cudaMemcpy2DAsync(dX + g * ld, ld * 8, X + lc * ldx, ldx * 8, n * 8, std::min(nb, nloc - lc),
                               cudaMemcpyHostToDevice, ctx->stream);
      }
    } else if (kc > 0) {
      ce = cudaMemcpy2DAsync(dX + c0 * ld, ld * 8, X, ldx * 8, n * 8, kc, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (ce == cudaSuccess && what == 0)
      ce = cudaMemcpyAsync(dw, w, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (ce != cudaSuccess) {
      ctx->last_cuda = ce;
      ctx->last_error = std::string("h2d: ") + cudaGetErrorString(ce);
      rc = EKB_ERR_CUDA;
    }
    // the checked columns need not coincide with the rank's slab of the nvec computed ones: make X whole
    if (!rc && cyclic) rc = comm_allreduce_sum(ctx, dX, (size_t)ld * nvec);  // disjoint pieces: the sum is a gather
    else if (!rc && ctx->nranks > 1) rc = comm_allgather_cols(ctx, dX, ld, zb);
  }
  if (!rc) {
    if (what == 0) rc = eval_residual_norm(ctx, n, a1, dA, ld, dB, ld, dw, dX, ld, o1, o2, o3);
    else if (what == 1) rc = eval_orthogonality(ctx, n, a1, a2, dX, ld, dB, ld, o1);
    else rc = get_ipratios(ctx, n, nvec, dX, ld, dB, ld, o1);
  }
  cleanup();
  return rc;
}

int ekb200_eval_residual_norm(ekb200_ctx* h, int64_t n, int64_t nvec, int64_t ncheck, int64_t nnzA, const int32_t* ijA,
                              const double* vA, int64_t nnzB, const int32_t* ijB, const double* vB, const double* w,
                              const double* X, int64_t ldx, double* A_norm, double* res_norm_ave, double* res_norm_max) {
  CHECK_CTX(h);
  if (n <= 0) return -2;
  if (nvec <= 0 || nvec > n) return -3;
  if (ncheck <= 0 || ncheck > nvec) return -4;
  if (nnzA < 0 || (nnzA > 0 && (!ijA || !vA))) return -5;
  if (nnzB < 0 || (nnzB > 0 && (!ijB || !vB))) return -8;
  if (!w) return -11;
  if (!X) return -12;
  if (ldx < n) return -13;
  return verify_host(h, 0, n, nvec, ncheck, 0, nnzA, ijA, vA, nnzB, ijB, vB, w, X, ldx, A_norm, res_norm_ave,
                     res_norm_max);
}
int ekb200_eval_orthogonality(ekb200_ctx* h, int64_t n, int64_t nvec, int64_t index1, int64_t index2, int64_t nnzB,
                              const int32_t* ijB, const double* vB, const double* X, int64_t ldx,
                              double* orthogonality) {
  CHECK_CTX(h);
  if (n <= 0) return -2;
  if (nvec <= 0 || nvec > n) return -3;
  if (index1 < 1) return -4;
  if (index2 < index1 || index2 > nvec) return -5;
  if (nnzB < 0 || (nnzB > 0 && (!ijB || !vB))) return -6;
  if (!X) return -9;
  if (ldx < n) return -10;
  return verify_host(h, 1, n, nvec, index1, index2, 0, nullptr, nullptr, nnzB, ijB, vB, nullptr, X, ldx, orthogonality,
                     nullptr, nullptr);
}
int ekb200_get_ipratios(ekb200_ctx* h, int64_t n, int64_t nvec, int64_t nnzB, const int32_t* ijB, const double* vB,
                        const double* X, int64_t ldx, double* ipratios) {
  CHECK_CTX(h);
  if (n <= 0) return -2;
  if (nvec <= 0 || nvec > n) return -3;
  if (nnzB < 0 || (nnzB > 0 && (!ijB || !vB))) return -4;
  if (!X) return -7;
  if (ldx < n) return -8;
  if (!ipratios) return -9;
  return verify_host(h, 2, n, nvec, 0, 0, 0, nullptr, nullptr, nnzB, ijB, vB, nullptr, X, ldx, ipratios, nullptr, nullptr);
}

// ---- multi-GPU: one context per rank (dist.cu)
int ekb200_comm_unique_id(void* id128) {
  if (!id128) return -1;
  std::string err;
  return comm_unique_id(id128, &err);
}
int ekb200_comm_init(ekb200_ctx* h, int nranks, int rank, const void* id128) {
  CHECK_CTX(h);
  if (nranks < 1) return -2;
  if (rank < 0 || rank >= nranks) return -3;
  if (nranks > 1 && !id128) return -4;
  return comm_init(ctx, nranks, rank, id128);
}
int ekb200_comm_info(const ekb200_ctx* h, int* nranks, int* rank) {
  if (!h) return -1;
  if (nranks) *nranks = h->c.nranks;
  if (rank) *rank = h->c.rank;
  return 0;
}
int ekb200_comm_slab(const ekb200_ctx* h, int64_t ncols, int64_t* col0, int64_t* nloc) {
  if (!h) return -1;
  if (ncols < 0) return -2;
  std::vector<i64> zb;
  slab_bounds(ncols, h->c.nranks, 128, zb);
  if (col0) *col0 = zb[h->c.rank];
  if (nloc) *nloc = zb[h->c.rank + 1] - zb[h->c.rank];
  return 0;
}
int ekb200_comm_local_cols(const ekb200_ctx* h, int64_t ncols, int64_t* nloc) {
  if (!h) return -1;
  if (ncols < 0) return -2;
  if (!nloc) return -3;
  if (h->c.nranks > 1 && h->c.out_block > 0) {
    *nloc = numroc0(ncols, h->c.out_block, h->c.rank, h->c.nranks);
    return 0;
  }
  return ekb200_comm_slab(h, ncols, nullptr, nloc);
}
int ekb200_comm_allgather_slabs(ekb200_ctx* h, int64_t nrows, int64_t ncols, double* M, int64_t ld) {
  CHECK_CTX(h);
  if (nrows < 0) return -2;
  if (ncols < 0) return -3;
  if (ld < nrows) return -5;
  std::vector<i64> zb;
  slab_bounds(ncols, ctx->nranks, 128, zb);
  EKB_TRY(comm_allgather_cols(ctx, M, ld, zb));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int ekb200_comm_bcast(ekb200_ctx* h, void* dev_buf, int64_t bytes, int root) {
  CHECK_CTX(h);
  if (bytes < 0) return -3;
  if (root < 0 || root >= ctx->nranks) return -4;
  EKB_TRY(comm_bcast(ctx, dev_buf, (size_t)bytes, root));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int64_t ekb200_num_collectives(const ekb200_ctx* h) { return h ? h->c.collectives : 0; }

int ekb200_measure_fp64_peak(ekb200_ctx* h, double* dmma_tflops, double* dfma_tflops) {
  CHECK_CTX(h);
  if (!dmma_tflops) return -2;
  if (!dfma_tflops) return -3;
  return measure_fp64_peak(ctx, dmma_tflops, dfma_tflops);
}

}  // extern "C"
