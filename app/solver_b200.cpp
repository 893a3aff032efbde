// ek_solver_b200_m -- C++ twin of fortran/solver_b200.f90: the B200 back-end behind eigen_solver
// (reference src/solver_main.f90:52-99), with the house signature of solve_with_general_scalapack
// (src/solver_scalapack_all.f90:127-168) and the error convention of src/generalized_to_standard.f90:25-30
// (`info(<routine>): N` on the master, then terminate).
//
// One library context per rank, created on first use and kept until b200_finalize so that the verifier and
// get_ipratios reuse the device arena of the solve.  With P ranks (--ngpu P) rank 0 draws the NCCL id, the shared
// board hands it over (the role of mpi_bcast), and the SAME entry points then run sharded: replicated inputs in,
// all eigenvalues plus the rank's column slab of the eigenvectors out (a 1 x P grid with one block column per rank).
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include "../include/ekb200.h"
#include "ek_app.hpp"
#include "launcher.hpp"

namespace ekapp {

static ekb200_ctx* s_ctx = nullptr;
// device-resident state of a synthetic run (kept for the verifier): full n x n_vec eigenvectors, eigenvalues
static double *s_dev_Z = nullptr, *s_dev_w = nullptr;
static int64_t s_dev_ldz = 0;
static void* s_host_vectors = nullptr;  // pinned local piece of the eigenvector matrix (blacs%Vectors)

static void check(int info, const char* routine) {
  if (info == 0) return;
  if (check_master()) {
    printf("info(%s): %d\n", routine, info);
    if (s_ctx && (info < 0 || info >= 1000000))
      printf("%s; %s\n", ekb200_strerror(info), ekb200_last_error(s_ctx));
    fflush(stdout);
  }
  terminate(std::string("solver_b200: ") + routine + " failed", info);
}

ekb200_ctx* b200_context(const ek_process_t& proc) {
  if (s_ctx) return s_ctx;
  const int n_dev = ekb200_device_count();
  if (n_dev < 1) terminate("solver_b200: no usable CUDA device (there is no CPU fallback)", 1);
  int info = ekb200_create(&s_ctx, proc.my_rank % n_dev);
  if (info != 0) {
    if (check_master()) printf("info(ekb200_create): %d\n", info);
    s_ctx = nullptr;
    terminate("solver_b200: no usable CUDA device (there is no CPU fallback)", info);
  }
  if (proc.n_procs > 1) {
    SharedBoard* board = shared_board();
    if (!board) terminate("solver_b200: rank launcher missing", 1);
    if (proc.my_rank == 0) {
      info = ekb200_comm_unique_id(board->nccl_id);
      board->error_code.store(info);
      board->id_ready.store(1);
    } else {
      wait_for_nccl_id();
      info = board->error_code.load();
    }
    if (info != 0) terminate("solver_b200: NCCL id could not be drawn", info);
    info = ekb200_comm_init(s_ctx, proc.n_procs, proc.my_rank, board->nccl_id);
    if (info != 0) terminate("solver_b200: NCCL communicator could not be created", info);
  }
  return s_ctx;
}

void b200_finalize() {
  if (!s_ctx) return;
  if (s_dev_Z) ekb200_dev_free(s_ctx, s_dev_Z);
  if (s_dev_w) ekb200_dev_free(s_ctx, s_dev_w);
  s_dev_Z = s_dev_w = nullptr;
  if (s_host_vectors) ekb200_host_free(s_ctx, s_host_vectors);
  s_host_vectors = nullptr;
  ekb200_destroy(s_ctx);
  s_ctx = nullptr;
}

// replays the library's CUDA-event timing table through add_event (src/event_logger.f90:23-65)
static void replay_events(ekb200_ctx* ctx) {
  const int n_ev = ekb200_num_events(ctx);
  for (int i = 0; i < n_ev; ++i) {
    const char* name = nullptr;
    double seconds = 0.0;
    int rep = 0;
    if (ekb200_get_event(ctx, i, &name, &seconds, &rep) != 0 || !name) continue;
    for (int r = 1; r < rep; ++r) add_event(name, 0.0, false);  // keep num_repeated
    add_event(name, seconds);
  }
  ekb200_clear_events(ctx);
}

// synthetic matrices of SURVEY 8(d): A = u(seed,i,j), a_ii += shift;  B = u(seed,i,j)/n off the diagonal, b_ii = 2
int b200_fill_synthetic_A(ekb200_ctx* ctx, const ek_matrix_info_t& info, double* dA, int64_t ld) {
  return ekb200_fill_synthetic(ctx, info.rows, info.seed, 1.0, 0, info.diag_shift, dA, ld);
}
int b200_fill_synthetic_B(ekb200_ctx* ctx, const ek_matrix_info_t& info, double* dB, int64_t ld) {
  return ekb200_fill_synthetic(ctx, info.rows, info.seed, (double)info.rows, 1, 2.0, dB, ld);
}
double* b200_device_vectors(int64_t* ld) {
  if (ld) *ld = s_dev_ldz;
  return s_dev_Z;
}
double* b200_device_values() { return s_dev_w; }

void solve_with_b200(const ek_argument_t& arg, int64_t n, const ek_process_t& proc, const ek_sparse_mat_t& matrix_A,
                     ek_eigenpairs_types_union_t& eigenpairs, const ek_sparse_mat_t* matrix_B) {
  const double time_start = wtime();
  if (proc.n_procs_row != 1)
    terminate("solver_b200: the process grid must be 1 x P (one process column per B200)", 1);
  const bool generalized = matrix_B != nullptr;
  const int64_t n_vec = is_b200_select_solver(arg.solver_type) ? arg.n_vec : n;
  ekb200_ctx* ctx = b200_context(proc);
  ekb200_clear_events(ctx);
  if (g_block_size == 32 || g_block_size == 64) check(ekb200_set_option(ctx, "band", g_block_size), "ekb200_set_option");
  // 0: blocked pdsygst-style reduction; 1: explicit-inverse (ELPA-style) reduction, SURVEY 8(f3)
  check(ekb200_set_option(ctx, "reduction", arg.solver_type == "general_b200inv" ? 1 : 0), "ekb200_set_option");

  eigenpairs.type_number = 2;
  ek_eigenpairs_blacs_t& ep = eigenpairs.blacs;
  ep.values.assign((size_t)n, 0.0);
  int64_t col0 = 0, nloc = n_vec;
  check(ekb200_comm_slab(ctx, n_vec, &col0, &nloc), "ekb200_comm_slab");
  ep.col0 = col0;
  ep.loc_cols = nloc;
  ep.lld = n > 1 ? n : 1;
  // descriptor of the n x n_vec eigenvector matrix on the 1 x P grid, one block column per rank: the block size
  // is the slab width of rank 0 (the widest); on one rank it is setup_distributed_matrix's clamped g_block_size
  int64_t nb = g_block_size < n ? g_block_size : n;
  if (nb < 1) nb = 1;
  if (proc.n_procs > 1) {
    // slab bounds are a pure function of (n_vec, P): ceil(n_vec / P) rounded up to 128-column granules
    int64_t chunk = (n_vec + proc.n_procs - 1) / proc.n_procs;
    chunk = (chunk + 127) / 128 * 128;
    nb = chunk < n_vec ? chunk : n_vec;
    if (nb < 1) nb = 1;
  }
  ep.desc[dtype_] = 1;
  ep.desc[context_] = proc.context;
  ep.desc[rows_] = n;
  ep.desc[cols_] = n_vec;
  ep.desc[block_row_] = nb;
  ep.desc[block_col_] = nb;
  ep.desc[rsrc_] = 0;
  ep.desc[csrc_] = 0;
  ep.desc[local_rows_] = ep.lld;
  void* host = nullptr;
  check(ekb200_host_alloc(ctx, (int64_t)sizeof(double) * ep.lld * (nloc > 0 ? nloc : 1), &host), "ekb200_host_alloc");
  ep.Vectors = (double*)host;
  if (s_host_vectors) ekb200_host_free(ctx, s_host_vectors);
  s_host_vectors = host;

  int info = 0;
  if (arg.matrix_A_info.synthetic) {
    // inputs are generated in HBM; the solve runs on the device-resident entry points and the eigenvectors stay
    // resident for the verifier
    const int64_t ld = (n + 7) / 8 * 8;
    double *dA = nullptr, *dB = nullptr;
    check(ekb200_dev_alloc(ctx, ld * n * 8, (void**)&dA), "ekb200_dev_alloc");
    if (generalized) check(ekb200_dev_alloc(ctx, ld * n * 8, (void**)&dB), "ekb200_dev_alloc");
    check(ekb200_dev_alloc(ctx, ld * n_vec * 8, (void**)&s_dev_Z), "ekb200_dev_alloc");
    check(ekb200_dev_alloc(ctx, (n + 8) * 8, (void**)&s_dev_w), "ekb200_dev_alloc");
    s_dev_ldz = ld;
    const double t0 = wtime();
    check(b200_fill_synthetic_A(ctx, arg.matrix_A_info, dA, ld), "ekb200_fill_synthetic");
    if (generalized) check(b200_fill_synthetic_B(ctx, arg.matrix_B_info, dB, ld), "ekb200_fill_synthetic");
    ekb200_sync(ctx);
    add_event(generalized ? "solve_with_general_b200:setup_matrices" : "eigen_solver_b200:setup_matrices", wtime() - t0);
    info = generalized ? ekb200_sygvd_dev(ctx, n, n_vec, dA, ld, dB, ld, s_dev_w, s_dev_Z, ld)
                       : ekb200_syevd_dev(ctx, n, n_vec, dA, ld, s_dev_w, s_dev_Z, ld);
    replay_events(ctx);
    if (info == 0) {
      const double t1 = wtime();
      check(ekb200_d2h(ctx, ep.values.data(), s_dev_w, n * 8), "ekb200_d2h");
      if (nloc > 0)
        check(ekb200_d2h_matrix(ctx, ep.Vectors, ep.lld, s_dev_Z + (size_t)col0 * ld, ld, n, nloc), "ekb200_d2h_matrix");
      add_event(generalized ? "solve_with_general_b200:d2h" : "eigen_solver_b200:d2h", wtime() - t1);
    }
    ekb200_dev_free(ctx, dA);
    if (dB) ekb200_dev_free(ctx, dB);
  } else {
    const int32_t ij_dummy[2] = {1, 1};
    const double v_dummy[1] = {0.0};
    info = ekb200_sygvd_coo(ctx, n, n_vec, matrix_A.num_non_zeros, matrix_A.suffix.data(), matrix_A.value.data(),
                            generalized ? matrix_B->num_non_zeros : 0, generalized ? matrix_B->suffix.data() : ij_dummy,
                            generalized ? matrix_B->value.data() : v_dummy, ep.values.data(), ep.Vectors, ep.lld);
    replay_events(ctx);
  }
  if (info > EKB200_WARN_STEIN && info < EKB200_FAIL_STEDC) {
    // inverse iteration left some eigenvectors unconverged: the reference only REPORTS pdsyevx's IFAIL and carries on
    // (solver_scalapack_select.f90:61-67); the library has returned all results
    if (check_master()) {
      printf("[Warning] eigen_solver_b200_select: inverse iteration did not converge for %d of %d requested eigenvectors\n",
             info - EKB200_WARN_STEIN, (int)n_vec);
      fflush(stdout);
    }
    info = 0;
  }
  if (info != 0) {
    // same reporting as generalized_to_standard.f90:25-30; the routine name follows the range of the code
    if (generalized && info > 0 && info <= n) check(info, "pdpotrf");
    if (info > EKB200_FAIL_STEDC && info < 1000000) check(info - EKB200_FAIL_STEDC, "pdstedc");
    check(info, generalized ? "ekb200_sygvd" : "ekb200_syevd");
  }
  add_event(generalized ? "solve_with_general_b200:wall" : "eigen_solver_b200:wall", wtime() - time_start);
}

}  // namespace ekapp
