// Whole-solve drivers on device-resident matrices: the B200 twins of
//   eigen_solver_scalapack_all   (reference src/solver_scalapack_all.f90:19-124): pdsytrd -> gather d,e ->
//                                pdstedc -> pdormtr, here sy2sb -> sb2st -> stedc -> apply_q2 -> apply_q1;
//   solve_with_general_scalapack (src/solver_scalapack_all.f90:127-168): reduce_generalized ->
//                                eigen_solver_scalapack_all -> recovery_generalized;
//   the `-n` selecting workflows (src/solver_main.f90:59-75) as "D&C on all of T, back-transform nev columns".
// Every stage is timed with CUDA events on the context stream and recorded under the event names of
// SURVEY.md 8(b) so the caller can replay them through add_event (src/event_logger.f90:23-65).
#include <algorithm>

#include "common.cuh"

namespace ekb {

namespace {
// frees every registered pointer when it leaves scope (after draining the stream)
struct Scratch {
  Ctx* ctx;
  std::vector<void*> ptrs;
  explicit Scratch(Ctx* c) : ctx(c) {}
  int get(void** p, size_t bytes) {
    int rc = ctx_alloc(ctx, p, bytes);
    if (rc == 0) ptrs.push_back(*p);
    return rc;
  }
  void release(void* p) {
    for (auto& q : ptrs)
      if (q == p) {
        cudaStreamSynchronize(ctx->stream);
        ctx_free(ctx, p);
        q = nullptr;
      }
  }
  ~Scratch() {
    cudaStreamSynchronize(ctx->stream);
    for (void* p : ptrs)
      if (p) ctx_free(ctx, p);
  }
};
}  // namespace

// A (n x n, full symmetric, destroyed) -> w (n, ascending), Z (n x nev: eigenvectors of the nev lowest).
// With P > 1 ranks (dist.cu) every rank enters with the same A and leaves with all of w and with ITS column
// slab Z(:, c0 : c0+kc) (slab_bounds(nev, P, 128)); the other columns of Z are scratch.  Dense-to-band is
// sharded by block columns, bulge chasing and the lower D&C levels are replicated, the top D&C merge and
// both back-transformations act on the slab only (no data-path collective).
int syevd_dev(Ctx* ctx, i64 n, i64 nev, double* A, i64 lda, double* w, double* Z, i64 ldz, double* merge_flops) {
  if (n <= 0 || nev <= 0) return 0;
  int stein_fail = 0;  // eigenvectors whose inverse iteration did not converge: a warning, the solve goes on
  const int b = ctx->band;
  std::vector<i64> zb;
  slab_bounds(nev, ctx->nranks, 128, zb);
  const i64 c0 = zb[ctx->rank], kc = zb[ctx->rank + 1] - zb[ctx->rank];
  double* Zs = Z + c0 * ldz;  // this rank's slab
  Scratch sc(ctx);
  StageTimer total(ctx, "eigen_solver_b200");
  const i64 ldab = 2 * b;
  const int npan = sy2sb_num_panels(n, b);
  double *AB = nullptr, *T1 = nullptr, *d = nullptr, *e = nullptr;
  EKB_TRY(sc.get((void**)&AB, (size_t)ldab * n * sizeof(double)));
  EKB_TRY(sc.get((void**)&T1, (size_t)(npan > 0 ? npan : 1) * b * b * sizeof(double)));
  EKB_TRY(sc.get((void**)&d, (size_t)(n + 8) * sizeof(double)));
  EKB_TRY(sc.get((void**)&e, (size_t)(n + 8) * sizeof(double)));
  {
    StageTimer t(ctx, "eigen_solver_b200:sy2sb");
    double* work = nullptr;
    EKB_TRY(sc.get((void**)&work, sy2sb_workspace_doubles(n, b, ctx->num_sms) * sizeof(double)));
    int rc = ctx->nranks > 1 ? sy2sb_dist(ctx, n, b, A, lda, AB, ldab, T1) : sy2sb(ctx, n, b, A, lda, AB, ldab, T1, work);
    t.stop();
    sc.release(work);
    if (rc) return rc;
  }
  const i64 ldv = round_up(n, 8);
  const int ldtau = sb2st_max_tasks(n, b);
  double *V2 = nullptr, *TAU2 = nullptr;
  EKB_TRY(sc.get((void**)&V2, (size_t)ldv * n * sizeof(double)));
  EKB_TRY(sc.get((void**)&TAU2, (size_t)ldtau * n * sizeof(double)));
  {
    StageTimer t(ctx, "eigen_solver_b200:sb2st");
    int* prog = nullptr;
    EKB_TRY(sc.get((void**)&prog, (size_t)(n + 8) * sizeof(int)));
    EKB_CUDA(cudaMemsetAsync(e, 0, (size_t)(n + 8) * sizeof(double), ctx->stream));
    int rc = sb2st(ctx, n, b, AB, ldab, V2, ldv, TAU2, ldtau, prog, d, e);
    t.stop();
    sc.release(prog);
    if (rc) return rc;
  }
  sc.release(AB);
  // Tridiagonal eigenproblem.  All pairs: divide and conquer (pdstedc).  The -n solvers may instead follow pdsyevx
  // (bisection + inverse iteration, O(n k) memory): forced by option "select_method" = 2, and chosen automatically
  // when the n x n workspaces of the D&C path do not fit next to A and the bulge-chasing reflectors (n = 65536).
  bool use_stebz = false;
  if (nev < n) {
    if (ctx->select_method == 2) use_stebz = true;
    else if (ctx->select_method == 0) {
      // decided from the device's TOTAL memory and a size-only estimate of what is live (A, the bulge-chasing
      // reflectors, possibly the Cholesky factor, the eigenvector slab), never from the free memory of the moment:
      // every rank of a sharded solve must take the same branch (replicated stages have to stay bit-identical)
      size_t free_b = 0, total_b = 0;
      EKB_CUDA(cudaMemGetInfo(&free_b, &total_b));
      const double nn = (double)round_up(n, 8) * (double)n * sizeof(double);
      const double live = 3.0 * nn + (double)round_up(n, 8) * (double)nev * sizeof(double);
      const double need = (double)stedc_workspace_bytes(n) + nn;
      use_stebz = live + need > 0.9 * (double)total_b;
    }
  }
  if (use_stebz) {
    StageTimer t(ctx, "eigen_solver_b200:stebz_stein");
    void* work = nullptr;
    EKB_TRY(sc.get(&work, stebz_stein_workspace_bytes(n, ctx->num_sms)));
    int rc = stebz_stein(ctx, n, d, e, w, nev, c0, c0 + kc, Z, ldz, work);
    t.stop();
    sc.release(work);
    // > 0: eigenvectors that failed dstein's growth test (IFAIL of pdsyevx).  The reference only reports this
    // (solver_scalapack_select.f90:61-67) and carries on with what pdsyevx returned; so does this solver: the vectors
    // are back-transformed like the others and the caller gets EKB_WARN_STEIN + count.
    if (rc < 0 || rc >= 1000000) return rc;
    stein_fail = rc;
  } else {
    StageTimer t(ctx, "eigen_solver_b200:stedc");
    void* work = nullptr;
    EKB_TRY(sc.get(&work, stedc_workspace_bytes(n)));
    int rc;
    if (nev == n) {
      rc = stedc(ctx, n, d, e, w, Z, ldz, work, merge_flops, c0, c0 + kc);
    } else {
      double* ZT = nullptr;
      const i64 ldt = round_up(n, 8);
      EKB_TRY(sc.get((void**)&ZT, (size_t)ldt * n * sizeof(double)));
      rc = stedc(ctx, n, d, e, w, ZT, ldt, work, merge_flops, c0, c0 + kc);
      if (rc == 0) rc = copy_matrix(ctx, ZT + c0 * ldt, ldt, Zs, ldz, n, kc);
      sc.release(ZT);
    }
    t.stop();
    sc.release(work);
    if (rc > 0 && rc < 1000000) return EKB_FAIL_STEDC + rc;  // leaf problems that did not converge (info(pdstedc))
    if (rc) return rc;
  }
  {
    StageTimer t(ctx, "eigen_solver_b200:ormtr_sb2st");
    int rc = kc > 0 ? apply_q2(ctx, n, b, V2, ldv, TAU2, ldtau, kc, Zs, ldz) : 0;
    t.stop();
    if (rc) return rc;
  }
  sc.release(V2);
  sc.release(TAU2);
  {
    StageTimer t(ctx, "eigen_solver_b200:ormtr_sy2sb");
    double* work = nullptr;
    EKB_TRY(sc.get((void**)&work, apply_q1_workspace_doubles(n, b, kc > 0 ? kc : 1) * sizeof(double)));
    int rc = kc > 0 ? apply_q1(ctx, n, b, A, lda, T1, kc, Zs, ldz, work) : 0;
    t.stop();
    sc.release(work);
    if (rc) return rc;
  }
  total.stop();
  EKB_CUDA(cudaGetLastError());
  if (ctx->nranks > 1 && use_stebz) {  // every rank reports the same count (sum over the slabs)
    double* cnt = nullptr;
    EKB_TRY(sc.get((void**)&cnt, 8 * sizeof(double)));
    const double mine = (double)stein_fail;
    EKB_CUDA(cudaMemcpyAsync(cnt, &mine, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EKB_TRY(comm_allreduce_sum(ctx, cnt, 1));
    double tot = 0.0;
    EKB_CUDA(cudaMemcpyAsync(&tot, cnt, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    EKB_CUDA(cudaStreamSynchronize(ctx->stream));
    stein_fail = (int)tot;
  }
  return stein_fail > 0 ? EKB_WARN_STEIN + stein_fail : 0;
}

// A, B full symmetric (B SPD) on the device.  B <- L (lower), A destroyed, w ascending, Z^T B Z = I.
int sygvd_dev(Ctx* ctx, i64 n, i64 nev, double* A, i64 lda, double* B, i64 ldb, double* w, double* Z, i64 ldz,
              double* invd, double* merge_flops, HostOverlap* ov) {
  if (n <= 0 || nev <= 0) return 0;
  StageTimer total(ctx, "solve_with_general_b200");
  Scratch sc(ctx);
  double* X = nullptr;  // L^-1 of the explicit-inverse variant (option "reduction" = 1), kept for the recovery
  const i64 ldx = round_up(n, 8);
  {
    StageTimer red(ctx, "reduce_generalized_b200");
    {
      StageTimer t(ctx, "reduce_generalized_b200:potrf");
      int rc = potrf_lower(ctx, n, B, ldb, invd);
      t.stop();
      if (rc) return rc;  // > 0: order of the leading minor that is not positive definite (info(pdpotrf))
    }
    if (ov && ov->a_ready) EKB_CUDA(cudaStreamWaitEvent(ctx->stream, ov->a_ready, 0));  // A was uploaded beside potrf
    if (ctx->reduction == 1) {
      // ELPA-style workflow (reference src/solver_elpa_eigenexa.f90:110-150): invert the factor, two multiplies.
      // Replicated on every rank (deterministic kernels: identical bits); the recovery acts on the rank's slab.
      EKB_TRY(sc.get((void**)&X, (size_t)ldx * n * sizeof(double)));
      {
        StageTimer t(ctx, "reduce_generalized_b200:trtri");
        int rc = trtri_lower(ctx, n, B, ldb, invd, X, ldx);
        t.stop();
        if (rc) return rc;
      }
      {
        StageTimer t(ctx, "reduce_generalized_b200:multiply");
        double* C = nullptr;
        EKB_TRY(sc.get((void**)&C, (size_t)ldx * n * sizeof(double)));
        int rc = sygst_inverse(ctx, n, A, lda, X, ldx, C, ldx);
        t.stop();
        sc.release(C);
        if (rc) return rc;
      }
    } else {
      StageTimer t(ctx, "reduce_generalized_b200:sygst");
      int rc = sygst_dist(ctx, n, A, lda, B, ldb, invd);
      t.stop();
      if (rc) return rc;
    }
    red.stop();
  }
  int warn = syevd_dev(ctx, n, nev, A, lda, w, Z, ldz, merge_flops);
  if (warn != 0 && !(warn > EKB_WARN_STEIN && warn < EKB_FAIL_STEDC)) return warn;
  {
    StageTimer t(ctx, "recovery_generalized_b200");
    std::vector<i64> zb;
    slab_bounds(nev, ctx->nranks, 128, zb);
    const i64 c0 = zb[ctx->rank], kc = zb[ctx->rank + 1] - zb[ctx->rank];
    int rc = 0;
    if (kc > 0 && ov && ov->host_Z && !X && kc >= 2048 && ctx_ensure_aux(ctx) == 0) {
      // host entry point: back-substitute the slab in four column chunks and send each finished chunk to the caller's
      // array on the side stream while the next one is solved (columns are independent in L^T X = Z)
      const i64 step = round_up((kc + 3) / 4, 128);
      for (i64 q0 = 0; q0 < kc && !rc; q0 += step) {
        const i64 qc = std::min(step, kc - q0);
        double* Zq = Z + (c0 + q0) * ldz;
        rc = trsm_lower(ctx, TRSM_LLT, n, qc, B, ldb, invd, Zq, ldz);
        if (rc) break;
        EKB_CUDA(cudaEventRecord(ctx->aux_ev[1], ctx->stream));
        EKB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[1], 0));
        EKB_CUDA(cudaMemcpy2DAsync(ov->host_Z + q0 * ov->ld_host_Z, ov->ld_host_Z * 8, Zq, ldz * 8, n * 8, qc,
                                   cudaMemcpyDeviceToHost, ctx->aux_stream));
      }
      ov->z_downloaded = rc == 0;
    } else if (kc > 0) {
      rc = X ? trmm_lower_t(ctx, n, kc, X, ldx, Z + c0 * ldz, ldz)  // Z <- L^-T Z as a product (pdtrmm_EV)
             : trsm_lower(ctx, TRSM_LLT, n, kc, B, ldb, invd, Z + c0 * ldz, ldz);
    }
    t.stop();
    if (rc) return rc;
  }
  total.stop();
  EKB_CUDA(cudaGetLastError());
  return warn;
}

}  // namespace ekb
