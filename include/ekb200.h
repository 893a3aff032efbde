/* libekb200 -- B200-native (sm_100a) dense FP64 symmetric eigensolver: flat C-ABI.
 *
 * Drop-in boundary: EigenKernel's solver dispatch `eigen_solver` (reference src/solver_main.f90:22-100).
 * The new solver names b200 / b200_select / general_b200 / general_b200_select are served by module
 * ek_solver_b200_m (fortran/solver_b200.f90) which binds the entry points below with ISO_C_BINDING.
 * Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *  - every function returns a LAPACK-style `info`: 0 = ok, -i = i-th argument illegal, >0 = numerical
 *    failure (e.g. order of the non-positive leading minor in the Cholesky factorization);
 *    >= 1000001 = CUDA/internal failure (ekb200_strerror / ekb200_last_error give the text).
 *    The whole-solve entry points tell their positive codes apart by range: 1..n = info(pdpotrf);
 *    EKB200_WARN_STEIN + k = WARNING, k eigenvectors of a -n solve did not converge in inverse iteration, results
 *    were still computed and returned (pdsyevx's IFAIL report, solver_scalapack_select.f90:61-67);
 *    EKB200_FAIL_STEDC + k = k leaf problems of the divide and conquer failed (info(pdstedc)).
 *    The library never aborts and never prints; the Fortran wrapper turns info != 0 into
 *    `terminate(msg, info)` exactly as generalized_to_standard.f90:25-30 does.
 *  - matrices are column-major FP64; `ld*` are leading dimensions in elements; indices are int64.
 *  - "host" pointers may be pageable; "dev" pointers are device memory of the context's GPU.
 *  - a context is not thread-safe; different contexts are independent.
 */
#ifndef EKB200_H
#define EKB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ekb200_ctx ekb200_ctx;
#define EKB200_WARN_STEIN 500000
#define EKB200_FAIL_STEDC 600000

/* ---- context (replaces setup_distribution, src/processes.f90:17-36: one context = one GPU "grid") */
int ekb200_create(ekb200_ctx** ctx, int device);
int ekb200_destroy(ekb200_ctx* ctx);
const char* ekb200_strerror(int info);
const char* ekb200_last_error(const ekb200_ctx* ctx);
int ekb200_set_option(ekb200_ctx* ctx, const char* key, int64_t value); /* "band" = half bandwidth b (32|64); "profile_gemm" = 0|1;
                                                                           "cache_device_memory" = 1|0 (caching arena; 0 also trims);
                                                                           "select_method" = 0 auto | 1 D&C | 2 bisection + inverse
                                                                           iteration for the -n solvers;
                                                                           "reduction" = 0 blocked pdsygst-style | 1 explicit inverse of
                                                                           L (-s general_b200inv; solver_elpa_eigenexa.f90:110-150);
                                                                           "out_block" = NB of the caller's block-cyclic eigenvector
                                                                           descriptor (see ekb200_comm_local_cols; 0 = column slabs);
                                                                           tuning / experiments (defaults are the measured best):
                                                                           "sb2st_variant", "sb2st_warps", "sb2st_rwarp", "sb2st_cps",
                                                                           "panel_qr_variant", "sy2sb_lookahead", "gemm_bulk",
                                                                           "stedc_shard", "q2_kc", "gemm_autosplit" */
int ekb200_version(void);
int ekb200_device_count(void); /* visible CUDA devices (0 when there is none); a rank uses device = local rank */

/* ---- timing table (replaces add_event, src/event_logger.f90:23-65; seconds are CUDA-event times) */
int ekb200_num_events(const ekb200_ctx* ctx);
int ekb200_get_event(const ekb200_ctx* ctx, int i, const char** name, double* seconds, int* num_repeated);
int ekb200_clear_events(ekb200_ctx* ctx);

/* ---- device memory owned by the context (replaces the allocatable arrays of
 *      setup_distributed_matrix, src/distribute_matrix.f90:92-148) */
int ekb200_dev_alloc(ekb200_ctx* ctx, int64_t bytes, void** dev_ptr);
int ekb200_dev_free(ekb200_ctx* ctx, void* dev_ptr);
int ekb200_h2d(ekb200_ctx* ctx, void* dev_dst, const void* host_src, int64_t bytes);
int ekb200_d2h(ekb200_ctx* ctx, void* host_dst, const void* dev_src, int64_t bytes);
int ekb200_h2d_matrix(ekb200_ctx* ctx, double* dev_dst, int64_t ldd, const double* host_src, int64_t lds, int64_t m,
                      int64_t n);
int ekb200_d2h_matrix(ekb200_ctx* ctx, double* host_dst, int64_t ldh, const double* dev_src, int64_t ldd, int64_t m,
                      int64_t n);
int ekb200_sync(ekb200_ctx* ctx);

/* ---- input construction on the device
 * ekb200_coo_to_dense: distribute_global_sparse_matrix (src/distribute_matrix.f90:401-422): zero A, scatter
 *   the 1-based COO entries `ij` (Fortran suffix(2,nnz)) with symmetric mirroring.  Host COO in.
 * ekb200_fill_synthetic: counter-hash generator of SURVEY.md 8(d): off-diagonal u(seed,i,j)/offdiag_div,
 *   diagonal u+diag_value (diag_mode 0) or diag_value (diag_mode 1). */
int ekb200_coo_to_dense(ekb200_ctx* ctx, int64_t n, int64_t nnz, const int32_t* host_ij, const double* host_v,
                        double* dev_A, int64_t lda);
int ekb200_fill_synthetic(ekb200_ctx* ctx, int64_t n, uint64_t seed, double offdiag_div, int diag_mode,
                          double diag_value, double* dev_A, int64_t lda);

/* ---- stage-level entry points on device-resident matrices (hybrids, per-stage tests, ncu targets).
 * ekb200_dgemm: the DMMA GEMM engine; transa/transb are 'N' or 'T'.  (PBLAS pdgemm, distribute_matrix.f90:47) */
int ekb200_dgemm(ekb200_ctx* ctx, char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha,
                 const double* dev_A, int64_t lda, const double* dev_B, int64_t ldb, double beta, double* dev_C,
                 int64_t ldc);
/* ekb200_potrf: pdpotrf('L') (generalized_to_standard.f90:24). B <- L (lower; strict upper left untouched).
 *   info = i > 0: leading minor of order i not positive definite. */
int ekb200_potrf(ekb200_ctx* ctx, int64_t n, double* dev_B, int64_t ldb);
/* ekb200_sygst: pdsygst(1,'L') (generalized_to_standard.f90:37). A <- L^-1 A L^-T.  A must hold the full
 *   symmetric matrix on entry; on exit both triangles hold the (symmetric) result. L from ekb200_potrf. */
int ekb200_sygst(ekb200_ctx* ctx, int64_t n, double* dev_A, int64_t lda, const double* dev_L, int64_t ldl);
/* ekb200_trtrs_lt: pdtrtrs('L','T','N') (generalized_to_standard.f90:103). Z <- L^-T Z, Z is n x nrhs. */
int ekb200_trtrs_lt(ekb200_ctx* ctx, int64_t n, int64_t nrhs, const double* dev_L, int64_t ldl, double* dev_Z,
                    int64_t ldz);

/* ekb200_sy2sb: first half of pdsytrd('L') (solver_scalapack_all.f90:59) in its two-stage form: dense -> band.
 *   A (full symmetric on entry) is overwritten: the Householder panels V_p (explicit unit-lower-trapezoidal)
 *   stay below the band; dev_T1 receives the (b x b) compact-WY factors, one per panel
 *   (ekb200_sy2sb_num_panels of them); dev_AB (ldab >= 2b rows, n columns) receives the band in lower band
 *   storage AB(i-j, j) = A(i,j), rows b+1..ldab-1 zeroed (room for the bulges of ekb200_sb2st). */
int ekb200_sy2sb(ekb200_ctx* ctx, int64_t n, double* dev_A, int64_t lda, double* dev_AB, int64_t ldab, double* dev_T1);
int ekb200_sy2sb_num_panels(const ekb200_ctx* ctx, int64_t n);
/* ekb200_sb2st: second half of pdsytrd('L') + the d/e gather (solver_scalapack_all.f90:59,75-78): band ->
 *   tridiagonal by bulge chasing.  dev_AB is destroyed.  dev_d (n), dev_e (n-1) receive the tridiagonal.
 *   dev_V2 (ldv >= n rows, n columns): column s = concatenated Householder vectors of sweep s (task t at rows
 *   s+1+t*b.., leading 1 stored); dev_TAU2 (ldtau = ekb200_sb2st_max_tasks rows, n columns): TAU2(t, s). */
int ekb200_sb2st(ekb200_ctx* ctx, int64_t n, double* dev_AB, int64_t ldab, double* dev_V2, int64_t ldv,
                 double* dev_TAU2, int64_t ldtau, double* dev_d, double* dev_e);
int ekb200_sb2st_max_tasks(const ekb200_ctx* ctx, int64_t n);
int ekb200_get_band(const ekb200_ctx* ctx);

/* ekb200_stedc: pdstedc('I') (solver_scalapack_all.f90:96-98): eigen-decomposition of the symmetric tridiagonal
 *   (dev_d, dev_e), both destroyed.  dev_w (n) ascending eigenvalues, dev_Z (n x n) orthonormal eigenvectors.
 *   info > 0: that many leaf problems failed to converge.  merge_flops (host, may be NULL): actual FLOPs of the
 *   merge GEMMs after deflation. */
int ekb200_stedc(ekb200_ctx* ctx, int64_t n, double* dev_d, double* dev_e, double* dev_w, double* dev_Z, int64_t ldz,
                 double* merge_flops);

/* ekb200_stebz_stein: the tridiagonal half of pdsyevx('V','I','L', il = 1, iu = nev, abstol = 2 safmin)
 *   (solver_scalapack_select.f90:52-60), i.e. pdstebz + pdstein: bisection on the Sturm count, one eigenvalue per
 *   thread, then inverse iteration, one cluster of close eigenvalues per warp (reorthogonalised inside the cluster).
 *   dev_d (n), dev_e (n-1) are not modified.  dev_w (n) receives ALL eigenvalues ascending, dev_Z (n x nev) the
 *   eigenvectors of the nev lowest.  O(n nev) memory.  info > 0: that many eigenvectors failed to converge (IFAIL).
 *   The -n solvers use it instead of ekb200_stedc when option "select_method" = 2, or = 0 (auto) and the n x n
 *   workspaces of the divide-and-conquer path do not fit (n = 65536 on one B200). */
int ekb200_stebz_stein(ekb200_ctx* ctx, int64_t n, int64_t nev, const double* dev_d, const double* dev_e, double* dev_w,
                       double* dev_Z, int64_t ldz);

/* ekb200_apply_q2 / ekb200_apply_q1: the two halves of pdormtr('L','L','N') (solver_scalapack_all.f90:115-116):
 *   Z (n x nrhs) <- Q2 Z with the bulge-chasing reflectors of ekb200_sb2st, then Z <- Q1 Z with the panels
 *   (V below the band of dev_A, T in dev_T1) of ekb200_sy2sb.  dev_A is modified (explicit zeros). */
int ekb200_apply_q2(ekb200_ctx* ctx, int64_t n, int64_t nrhs, const double* dev_V2, int64_t ldv, const double* dev_TAU2,
                    int64_t ldtau, double* dev_Z, int64_t ldz);
int ekb200_apply_q1(ekb200_ctx* ctx, int64_t n, int64_t nrhs, double* dev_A, int64_t lda, const double* dev_T1,
                    double* dev_Z, int64_t ldz);

/* ---- whole-solve entry points.
 * nev = n: all eigenpairs (-s b200 / general_b200); nev < n: the nev lowest (-s b200_select /
 * general_b200_select, option -n; solver_main.f90:59-75).  w always receives all n eigenvalues ascending
 * (like `values` of pdsyevx, solver_scalapack_select.f90:45); Z is n x nev.
 *
 * Device-resident variants (inputs already in HBM; both triangles of the symmetric matrices must be filled):
 * ekb200_syevd_dev: eigen_solver_scalapack_all (solver_scalapack_all.f90:19-124).  dev_A destroyed.
 * ekb200_sygvd_dev: solve_with_general_scalapack (solver_scalapack_all.f90:127-168): dev_B <- L, dev_A
 *   destroyed, Z^T B Z = I.  info > 0 from the Cholesky step = order of the non-positive leading minor
 *   (info(pdpotrf), generalized_to_standard.f90:25-30). */
int ekb200_syevd_dev(ekb200_ctx* ctx, int64_t n, int64_t nev, double* dev_A, int64_t lda, double* dev_w, double* dev_Z,
                     int64_t ldz);
int ekb200_sygvd_dev(ekb200_ctx* ctx, int64_t n, int64_t nev, double* dev_A, int64_t lda, double* dev_B, int64_t ldb,
                     double* dev_w, double* dev_Z, int64_t ldz);
/* Host-pointer variants = what ek_solver_b200_m binds (1x1 BLACS grid: the local array IS the matrix).
 * Lower triangles of A and B are referenced ('L' everywhere in the reference); A and B are not modified;
 * host buffers may be pageable (pinned ones from ekb200_host_alloc transfer faster). */
int ekb200_syevd(ekb200_ctx* ctx, int64_t n, int64_t nev, const double* A, int64_t lda, double* w, double* Z,
                 int64_t ldz);
int ekb200_sygvd(ekb200_ctx* ctx, int64_t n, int64_t nev, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double* w, double* Z, int64_t ldz);
/* COO front door = setup_distributed_matrix + distribute_global_sparse_matrix + solve
 * (solver_scalapack_all.f90:141-144): ij is Fortran suffix(2,nnz), 1-based; nnzB = 0 selects the standard
 * problem. */
int ekb200_sygvd_coo(ekb200_ctx* ctx, int64_t n, int64_t nev, int64_t nnzA, const int32_t* ijA, const double* vA,
                     int64_t nnzB, const int32_t* ijB, const double* vB, double* w, double* Z, int64_t ldz);
/* actual FLOPs of the D&C merge products of the last solve (after deflation) */
double ekb200_last_merge_flops(const ekb200_ctx* ctx);
/* pinned host memory for fast transfers */
int ekb200_host_alloc(ekb200_ctx* ctx, int64_t bytes, void** host_ptr);
int ekb200_host_free(ekb200_ctx* ctx, void* host_ptr);

/* ---- instrumentation (bench.py): kernels launched so far by this context; with option "profile_gemm" = 1
 * every engine GEMM is bracketed by CUDA events and ekb200_gemm_profile returns (and resets) the summed device
 * seconds, algorithmic FLOPs and launch count of the DMMA GEMM kernel family. */
int64_t ekb200_num_launches(const ekb200_ctx* ctx);
/* CUDA-event stopwatch on the context's stream (the stream every kernel of the library is launched on) */
int ekb200_timer_start(ekb200_ctx* ctx);
int ekb200_timer_stop(ekb200_ctx* ctx, double* seconds);
int ekb200_gemm_profile(ekb200_ctx* ctx, double* seconds, double* flops, int64_t* launches);
/* the same for every profiled kernel family; arrays of EKB200_PROF_FAMILIES entries, indexed by
 * 0 GEMM engine (FLOPs) | 1 panel QR (FLOPs) | 2 Q2 apply (FLOPs) | 3 bulge chasing (effective bytes) |
 * 4 batched merge GEMMs (work not counted: 0) | 5 NCCL exchanges (bytes).  Resets all families. */
#define EKB200_PROF_FAMILIES 8
int ekb200_kernel_profile(ekb200_ctx* ctx, double* seconds, double* work, int64_t* launches);
/* (stage, family) breakdown of the last ekb200_kernel_profile / ekb200_gemm_profile call */
int ekb200_profile_rows(const ekb200_ctx* ctx);
int ekb200_profile_row(const ekb200_ctx* ctx, int i, const char** stage, int* family, double* seconds, double* work,
                       int64_t* launches);

/* ---- acceptance metrics and IPRs on the device.
 * ekb200_eval_residual_norm  = eval_residual_norm_blacs (src/verifier.f90:75-204, option -c):
 *     A_norm = ||A||_F, res_norm_ave / res_norm_max = mean / max over the first ncheck columns of
 *     ||A x_j - lambda_j B x_j||_2 / ||A||_F  (no division by ||x_j||, exactly as verifier.f90:198-199).
 * ekb200_eval_orthogonality  = eval_orthogonality_blacs (src/verifier.f90:233-330, option -t): Frobenius norm of the
 *     Gram matrix X(:, index1:index2)^T B X(:, index1:index2) scaled to unit diagonal with the diagonal zeroed
 *     (index1/index2 1-based, inclusive).
 * ekb200_get_ipratios        = get_ipratios (src/distribute_matrix.f90:18-78, ipratios.dat): sum_i x_ij^4 /
 *     (sum_i x_ij (B x)_ij)^2 for the first nvec columns, B = I for the standard problem.
 * Host variants take what the reference routines take: the replicated COO matrices (ij = Fortran suffix(2,nnz),
 * 1-based; nnzB = 0: standard problem) and the eigenpairs -- w(n) and X, the rank's LOCAL piece (n x nloc of the
 * nvec computed eigenvector columns; all of them on one rank).  *_dev variants work on device-resident full
 * symmetric A / B and on the n x nvec eigenvector buffer of the *_dev solvers (residual and IPR read the rank's own
 * slab of the checked columns, orthogonality needs all of index1..index2: call ekb200_comm_allgather_slabs first
 * when there is more than one rank).  Results are returned on the host, identical on every rank. */
int ekb200_eval_residual_norm(ekb200_ctx* ctx, int64_t n, int64_t nvec, int64_t ncheck, int64_t nnzA, const int32_t* ijA,
                              const double* vA, int64_t nnzB, const int32_t* ijB, const double* vB, const double* w,
                              const double* X, int64_t ldx, double* A_norm, double* res_norm_ave, double* res_norm_max);
int ekb200_eval_orthogonality(ekb200_ctx* ctx, int64_t n, int64_t nvec, int64_t index1, int64_t index2, int64_t nnzB,
                              const int32_t* ijB, const double* vB, const double* X, int64_t ldx, double* orthogonality);
int ekb200_get_ipratios(ekb200_ctx* ctx, int64_t n, int64_t nvec, int64_t nnzB, const int32_t* ijB, const double* vB,
                        const double* X, int64_t ldx, double* ipratios);
int ekb200_eval_residual_norm_dev(ekb200_ctx* ctx, int64_t n, int64_t ncheck, const double* dev_A, int64_t lda,
                                  const double* dev_B, int64_t ldb, const double* dev_w, const double* dev_X, int64_t ldx,
                                  double* A_norm, double* res_norm_ave, double* res_norm_max);
int ekb200_eval_orthogonality_dev(ekb200_ctx* ctx, int64_t n, int64_t index1, int64_t index2, const double* dev_X,
                                  int64_t ldx, const double* dev_B, int64_t ldb, double* orthogonality);
/* ekb200_eval_b_orthonormality_dev: the same Gram matrix, two numbers: `orthogonality` as above (verifier.f90:310-325)
 * and `gram_minus_identity` = || X^T B X - I ||_F including the diagonal, i.e. BASELINE.json's B-orthogonality
 * acceptance metric (the reference's own metric does not test normalisation).  Either output may be NULL. */
int ekb200_eval_b_orthonormality_dev(ekb200_ctx* ctx, int64_t n, int64_t index1, int64_t index2, const double* dev_X,
                                     int64_t ldx, const double* dev_B, int64_t ldb, double* orthogonality,
                                     double* gram_minus_identity);
int ekb200_get_ipratios_dev(ekb200_ctx* ctx, int64_t n, int64_t nvec, const double* dev_X, int64_t ldx,
                            const double* dev_B, int64_t ldb, double* ipratios);

/* ---- multi-GPU: ONE CONTEXT PER RANK, one rank per B200 (replaces the BLACS grid of src/processes.f90:17-65 and
 * the block-cyclic scatter of src/distribute_matrix.f90:92-148).  Rank 0 obtains a 128-byte id with
 * ekb200_comm_unique_id and hands it to the other ranks by whatever the host has (mpi_bcast in the Fortran app,
 * torch.distributed in the Python mirror); every rank then calls ekb200_comm_init.  After that the SAME solve entry
 * points run sharded: every rank passes the same (replicated) A and B -- exactly like the replicated COO the reference
 * hands to solve_with_general_scalapack (solver_scalapack_all.f90:127-132) -- and receives all of w and ITS column
 * slab of the eigenvectors, columns [col0, col0 + nloc) from ekb200_comm_slab(nev): a 1 x P process grid with one
 * column block per rank.  Host-pointer entry points: Z is the LOCAL piece (n x nloc, ldz), the analogue of
 * blacs%Vectors(lld, loc_cols); *_dev entry points: dev_Z is the full n x nev buffer on every rank (scratch outside
 * the slab), of which columns [col0, col0 + nloc) hold the result.  NCCL (bound with dlopen) carries
 * the all-gathers of the sharded pdsygst and the panel exchanges of the sharded dense-to-band reduction; the
 * back-transformations, the top D&C merge and pdtrtrs work on the slab with no data-path collective. */
int ekb200_comm_unique_id(void* id128);
int ekb200_comm_init(ekb200_ctx* ctx, int nranks, int rank, const void* id128);
int ekb200_comm_info(const ekb200_ctx* ctx, int* nranks, int* rank);
int ekb200_comm_slab(const ekb200_ctx* ctx, int64_t ncols, int64_t* col0, int64_t* nloc);
/* Rank-per-GPU callers whose eigenvector array was allocated by setup_distributed_matrix (distribute_matrix.f90:92-148):
 * that routine CLAMPS the block size to max(min(rows / nprow, cols / npcol), 1) (:114-120), i.e. floor(ncols / P) on
 * the 1 x P grid, which is not the slab width above unless 128 P divides ncols.  Option "out_block" = NB (the NB of
 * the caller's descriptor, desc(6)) makes the host-pointer entry points (ekb200_syevd / sygvd / sygvd_coo and the
 * ekb200_eval_* / ekb200_get_ipratios host variants) exchange the LOCAL PIECE OF THAT 1 x P BLOCK-CYCLIC DISTRIBUTION:
 * numroc(ncols, NB, rank, 0, P) columns, blocks rank, rank + P, ... (one NVLink all-gather inside the library).
 * ekb200_comm_local_cols returns that column count (the slab width when "out_block" is 0). */
int ekb200_comm_local_cols(const ekb200_ctx* ctx, int64_t ncols, int64_t* nloc);
/* every rank owns its slab of the columns of dev_M (nrows x ncols, ld): afterwards all ranks hold all columns */
int ekb200_comm_allgather_slabs(ekb200_ctx* ctx, int64_t nrows, int64_t ncols, double* dev_M, int64_t ld);
int ekb200_comm_bcast(ekb200_ctx* ctx, void* dev_buf, int64_t bytes, int root);
int64_t ekb200_num_collectives(const ekb200_ctx* ctx);

/* ---- measurement helper (roofline denominator; never on the solve path) */
int ekb200_measure_fp64_peak(ekb200_ctx* ctx, double* dmma_tflops, double* dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* EKB200_H */
