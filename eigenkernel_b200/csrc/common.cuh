// Shared declarations of libekb200: context, error convention, stage timers, kernel launch helpers.
// B200-native (sm_100a) dense FP64 symmetric eigensolver behind EigenKernel's solver boundary
// (reference: src/solver_main.f90:52-99).  No cuSOLVER / cuBLAS / CPU fallback on the solve path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>

#include "layout.h"

namespace ekb {

typedef long long i64;

// LAPACK-style info codes of the C-ABI: 0 ok, <0 argument -i illegal, >0 numerical failure.
// Internal CUDA failures map to EKB_ERR_CUDA.
// Whole-solve drivers tell their positive codes apart by range: 1..n = order of the non-positive leading minor of the
// Cholesky step (info(pdpotrf)); EKB_WARN_STEIN + k = k eigenvectors did not converge in inverse iteration (a WARNING:
// the results were computed and returned, like pdsyevx's IFAIL); EKB_FAIL_STEDC + k = k leaf problems of the divide
// and conquer failed (info(pdstedc)).
enum { EKB_WARN_STEIN = 500000, EKB_FAIL_STEDC = 600000 };
enum { EKB_ERR_CUDA = 1000001, EKB_ERR_NOMEM = 1000002, EKB_ERR_INTERNAL = 1000003, EKB_ERR_COMM = 1000004 };

struct Event {
  std::string name;
  double seconds;
  int num_repeated;
};

// Device-memory arena: one cudaMalloc'd slab per request, freed with the context.
struct Ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t aux_stream = nullptr;  // high-priority side stream (panel look-ahead of sy2sb); created on first use
  cudaEvent_t aux_ev[2] = {nullptr, nullptr};
  std::vector<Event> events;          // timing table replayed through add_event by the caller
  std::vector<void*> allocs;          // every device allocation handed out and not yet released
  std::vector<std::pair<void*, size_t>> live;   // ... with its (rounded) size
  std::vector<std::pair<void*, size_t>> cache;  // released blocks kept for reuse (caching arena, context.cu)
  size_t cached_bytes = 0;
  bool cache_enabled = true;
  cudaError_t last_cuda = cudaSuccess;
  // scratch
  double* splitk_ws = nullptr;        // split-K partial sums
  size_t splitk_ws_bytes = 0;
  double* splitk_ws_aux = nullptr;    // the same for work issued on aux_stream (swapped in by OnAuxStream, sy2sb.cu)
  size_t splitk_ws_aux_bytes = 0;
  int* d_info = nullptr;              // device-side info word(s)
  int* h_info = nullptr;              // pinned mirror
  // options
  int band = 64;                      // b: half bandwidth of the two-stage reduction
  int select_method = 0;              // -n solvers: 0 auto (D&C unless its workspaces do not fit), 1 D&C, 2 bisection + inverse iteration
  int reduction = 0;                  // 0: blocked pdsygst-style reduction; 1: explicit inverse (ELPA-style)
  int sb2st_variant = 1;              // 1: register-resident lag-2 kernel (round 2); 0: the round-1 shared-memory kernel
  int sb2st_warps = 8;                // compute warps per CTA of the register-resident kernel (8 | 16)
  int sb2st_rwarp = 1;                // 1: one extra warp forms the reflectors beside the updates; 0: warp 0 does
  int sb2st_cps = 0;                  // cap on resident CTAs per SM (0 = what the occupancy calculator allows)
  long long out_block = 0;            // > 0: host entry points deliver the 1 x P block-cyclic piece with this block size (layout.h)
  int stedc_shard = 1;                // P > 1: the two level-1 merges of the D&C are sharded over the ranks (0: replicated)
  int gemm_autosplit = 1;             // 1: products on the TMA-fed kernel choose their split-K factor by the round-count model (gemm.cu)
  int gemm_bulk = 1;                  // 1: big-tile products run on the TMA-fed warp-specialised GEMM kernel (gemm.cu)
  int panel_qr_variant = 1;           // 1: panel QR with the panel resident in shared memory; 0: the round-1 global-memory kernel
  int sy2sb_lookahead = 0;            // 1: factor panel p+1 on the side stream while the rank-2b update of panel p runs (measured: a loss, see sy2sb.cu)
  int q2_kc = 0;                      // columns of Z per CTA in apply_q2 (0 = choose; 64|80|96|112|128)
  // stage timers
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string last_error;
  // instrumentation
  long long launches = 0;             // kernels launched by this context (bench.py: gpu_launches)
  bool profile_gemm = false;          // per-launch CUDA events around every engine GEMM (roofline evidence)
  std::vector<cudaEvent_t> prof_events;   // pairs (start, stop), one per profiled launch
  std::vector<double> prof_flops;         // algorithmic work of that launch (FLOPs, or bytes for HBM-bound families)
  std::vector<int> prof_family;           // kernel family of that launch (ProfFamily)
  std::vector<const char*> prof_stage;    // innermost StageTimer name active at that launch
  struct ProfShape { int m, n, k, flags, tri, splitk; };
  std::vector<ProfShape> prof_shape;      // engine GEMMs: the shape of that launch (EKB200_GEMM_TRACE dump, profile_collect)
  const char* cur_stage = "";
  size_t prof_used = 0;
  struct ProfRow { std::string stage; int family; double seconds, work; long long launches; };
  std::vector<ProfRow> prof_table;        // (stage, family) aggregation of the last profile_collect
  // multi-GPU (dist.cu): one context per rank, NCCL communicator bound at run time
  int nranks = 1, rank = 0;
  void* comm = nullptr;               // ncclComm_t
  long long collectives = 0;          // NCCL collectives issued by this context
};
#define EKB_COUNT_LAUNCH(c) ((c)->launches++)

#define EKB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ctx->last_cuda = _e;                                                                  \
      ctx->last_error = std::string(#call) + ": " + cudaGetErrorString(_e) + " [" + __FILE__ + ":" +  \
                        std::to_string(__LINE__) + ", stage " + ctx->cur_stage + "]";       \
      return EKB_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define EKB_TRY(call)             \
  do {                            \
    int _i = (call);              \
    if (_i != 0) return _i;       \
  } while (0)

int ctx_alloc(Ctx* ctx, void** p, size_t bytes);
int ctx_free(Ctx* ctx, void* p);
void ctx_trim(Ctx* ctx);  // return every cached block to the driver
void ctx_add_event(Ctx* ctx, const char* name, double seconds);

// Scoped stage timer on the context stream (CUDA events, device time).
struct StageTimer {
  Ctx* ctx;
  const char* name;
  cudaEvent_t a, b;
  const char* prev_stage;
  bool done = false;
  StageTimer(Ctx* c, const char* n);
  ~StageTimer();  // a timer abandoned by an early error return records nothing
  double stop();  // records, synchronises, adds the event; returns seconds
};

static inline int cdiv(i64 a, i64 b) { return (int)((a + b - 1) / b); }
static inline i64 round_up(i64 a, i64 b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------- GEMM engine (gemm.cu)
// C(m x n) = alpha * op(A)(m x k) * op(B)(k x n) + beta * C, column-major, FP64 on DMMA.8x8x4 tiles.
struct GemmP {
  int m, n, k;
  const double* A;
  const double* B;
  double* C;
  i64 lda, ldb, ldc;
  double alpha, beta;
};
enum GemmFlags {
  GEMM_TA = 1,      // op(A) = A^T
  GEMM_TB = 2,      // op(B) = B^T
  GEMM_SYMA = 4,    // A is symmetric, only elements with (col - row) < 128 are valid (NN only)
  GEMM_RASTER = 8,  // set by gemm() itself: L2-aware tile order (TMA-fed kernel, deep-k products)
};
// tri_keep < 0: full C.  Otherwise only elements with (col - row) < tri_keep are guaranteed to be
// computed (whole tiles above that region are skipped): 1 = lower triangle, 128 = lower + a 128 band.
int gemm(Ctx* ctx, int flags, const GemmP& p, int tri_keep = -1, int splitk = 1);
// Batched: `batch` is a DEVICE array of nb problems; (max_m, max_n) bound the grid.
// k_hint: typical k-depth of the batch (the descriptors live on the device): >= 1024 selects the deep pipeline geometry.
int gemm_batched(Ctx* ctx, int flags, const GemmP* d_batch, int nb, int max_m, int max_n, int k_hint = 0);
int gemm_profile_collect(Ctx* ctx, double* seconds, double* flops, long long* launches);
// Per-launch CUDA-event brackets in "profile_gemm" mode, tagged by kernel family (roofline evidence measured
// live inside bench.py).  prof_begin/prof_end are no-ops when profiling is off.
enum ProfFamily { PROF_GEMM = 0, PROF_PANEL_QR = 1, PROF_Q2_APPLY = 2, PROF_SB2ST = 3, PROF_GEMM_BATCHED = 4,
                  PROF_NCCL = 5, PROF_STEBZ = 6 /* Sturm steps */, PROF_STEIN = 7 /* bytes */, PROF_FAMILIES = 8 };
int prof_begin(Ctx* ctx, int family, double work);
int prof_end(Ctx* ctx);
int profile_collect(Ctx* ctx, double* seconds /*[PROF_FAMILIES]*/, double* work, long long* launches);

// ---------------------------------------------------------------- elementwise helpers (fill.cu)
int fill_synthetic(Ctx* ctx, double* A, i64 lda, i64 n, uint64_t seed, double offdiag_scale, int diag_mode,
                   double diag_value);
int set_zero(Ctx* ctx, double* A, i64 lda, i64 m, i64 n);
int copy_matrix(Ctx* ctx, const double* A, i64 lda, double* B, i64 ldb, i64 m, i64 n);
int set_identity(Ctx* ctx, double* A, i64 lda, i64 n);
int coo_scatter(Ctx* ctx, double* A, i64 lda, i64 n, i64 nnz, const int32_t* d_ij, const double* d_v);
int symmetrize_from_lower(Ctx* ctx, double* A, i64 lda, i64 n);

// ---------------------------------------------------------------- stages
int potrf_lower(Ctx* ctx, i64 n, double* B, i64 ldb, double* invd /* n x 64 inverted diagonal blocks */);
enum TrsmKind { TRSM_RLT = 0 /* X L^T = B */, TRSM_LLN = 1 /* L X = B */, TRSM_LLT = 2 /* L^T X = B */ };
int trsm_lower(Ctx* ctx, int kind, i64 m, i64 n, const double* L, i64 ldl, const double* invd, double* Bm, i64 ldb);
int trtri_diag_blocks(Ctx* ctx, i64 n, const double* L, i64 ldl, double* invd);
int sygst_lower(Ctx* ctx, i64 n, double* A, i64 lda, const double* L, i64 ldl, const double* invd);
// explicit-inverse reduction (option "reduction" = 1): X = L^-1, A <- X A X^T, Z <- X^T Z
int trtri_lower(Ctx* ctx, i64 n, const double* L, i64 ldl, const double* invd, double* X, i64 ldx);
int sygst_inverse(Ctx* ctx, i64 n, double* A, i64 lda, const double* X, i64 ldx, double* C, i64 ldc);
int trmm_lower_t(Ctx* ctx, i64 n, i64 k, const double* X, i64 ldx, double* Z, i64 ldz);

// two-stage tridiagonalization
size_t sy2sb_workspace_doubles(i64 n, int b, int num_sms);
int sy2sb_num_panels(i64 n, int b);
int sy2sb(Ctx* ctx, i64 n, int b, double* A, i64 lda, double* AB, i64 ldab, double* T1, double* work);
int sb2st_max_tasks(i64 n, int b);
int sb2st(Ctx* ctx, i64 n, int b, double* AB, i64 ldab, double* V2, i64 ldv, double* TAU2, int ldtau, int* prog,
          double* d, double* e);

// back-transformations (ormtr.cu)
int apply_q2(Ctx* ctx, i64 n, int b, const double* V2, i64 ldv, const double* TAU2, int ldtau, i64 k, double* Z,
             i64 ldz);
size_t apply_q1_workspace_doubles(i64 n, int b, i64 k);
int apply_q1(Ctx* ctx, i64 n, int b, double* A, i64 lda, const double* T1, i64 k, double* Z, i64 ldz, double* work);

// whole-solve drivers (solve.cu)
int syevd_dev(Ctx* ctx, i64 n, i64 nev, double* A, i64 lda, double* w, double* Z, i64 ldz, double* merge_flops);
// Transfers of the host-pointer entry points that hide behind compute (api.cu solve_host): A may still be arriving
// on the side stream while B is factored (a_ready: waited for before A's first use), and the eigenvectors of the
// generalized problem can leave in column chunks on the side stream while the next chunk is back-substituted.
struct HostOverlap {
  cudaEvent_t a_ready = nullptr;
  double* host_Z = nullptr;   // caller's array for the rank's slab (column 0 = first column of the slab)
  i64 ld_host_Z = 0;
  bool z_downloaded = false;  // out: the slab has been (queued to be) downloaded on ctx->aux_stream
};
int ctx_ensure_aux(Ctx* ctx);  // creates ctx->aux_stream and ctx->aux_ev on first use
int sygvd_dev(Ctx* ctx, i64 n, i64 nev, double* A, i64 lda, double* B, i64 ldb, double* w, double* Z, i64 ldz,
              double* invd, double* merge_flops, HostOverlap* ov = nullptr);

size_t stedc_workspace_bytes(i64 n);
// Only the eigenvector columns [col_lo, col_hi) (ascending-eigenvalue positions) of the TOP merge are formed
// (all of them for col_lo = 0, col_hi = n); the other columns of Z are left undefined.
int stedc(Ctx* ctx, i64 n, double* d, double* e, double* w, double* Z, i64 ldz, void* work, double* flops_out,
          i64 col_lo, i64 col_hi);
// bisection + inverse iteration (stebz.cu): all n eigenvalues into w, eigenvector columns [col_lo, col_hi) of the nev
// lowest into Z(:, col_lo..col_hi); returns the number of vectors that failed dstein's growth test
size_t stebz_stein_workspace_bytes(i64 n, int num_sms);
int stebz_stein(Ctx* ctx, i64 n, const double* d, const double* e, double* w, i64 nev, i64 col_lo, i64 col_hi, double* Z,
                i64 ldz, void* work);

// ---------------------------------------------------------------- multi-GPU (dist.cu)
int comm_unique_id(void* id128, std::string* err);
int comm_init(Ctx* ctx, int nranks, int rank, const void* id128);
int comm_destroy(Ctx* ctx);
int comm_allgather_cols(Ctx* ctx, double* M, i64 ld, const std::vector<i64>& bounds);
int comm_bcast(Ctx* ctx, void* buf, size_t bytes, int root);
int comm_allgather(Ctx* ctx, const double* send, double* recv, size_t count);
int comm_allreduce_sum(Ctx* ctx, double* buf, size_t count);
int sy2sb_dist(Ctx* ctx, i64 n, int b, double* A, i64 lda, double* AB, i64 ldab, double* T1);
int transpose_matrix(Ctx* ctx, const double* A, i64 lda, i64 m, i64 n, double* B, i64 ldb);
int sygst_dist(Ctx* ctx, i64 n, double* A, i64 lda, const double* L, i64 ldl, const double* invd);

// ---------------------------------------------------------------- acceptance metrics + IPR (verify.cu)
int eval_residual_norm(Ctx* ctx, i64 n, i64 ncheck, const double* A, i64 lda, const double* B, i64 ldb, const double* w,
                       const double* Xfull, i64 ldx, double* A_norm, double* res_ave, double* res_max);
int eval_orthogonality(Ctx* ctx, i64 n, i64 index1, i64 index2, const double* Xfull, i64 ldx, const double* B, i64 ldb,
                       double* orthogonality, double* gram_minus_identity = nullptr);
int get_ipratios(Ctx* ctx, i64 n, i64 nvec, const double* Xfull, i64 ldx, const double* B, i64 ldb, double* ipr);

}  // namespace ekb
