// program eigbench (reference src/main.f90): read arguments and matrices, call eigen_solver, write
// eigenvalues.dat / eigenvector files / ipratios.dat, run the optional checks, write log.json.
// Same order of steps, same `main:*` events, same messages on stdout; mpi_init / mpi_bcast / mpi_finalize are
// replaced by the forked-rank launcher (processes.cpp), every rank reads the (small, replicated) inputs itself.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>

#include "ek_app.hpp"
#include "launcher.hpp"

using namespace ekapp;

static int run(int argc, char** argv) {
  ek_argument_t arg;
  ek_sparse_mat_t matrix_A, matrix_B;
  ek_eigenpairs_types_union_t eigenpairs;
  double A_norm = 0, rn_ave = 0, rn_max = 0, orthogonality = 0;
  std::vector<double> ipratios;
  ek_process_t proc;
  int ierr = 0;

  // mpi_init: the rank count comes from --ngpu (or EKB200_NGPU), ranks are forked before anything touches CUDA
  int ngpu = 1;
  if (const char* e = getenv("EKB200_NGPU")) ngpu = atoi(e) > 0 ? atoi(e) : 1;
  for (int i = 1; i + 1 < argc; ++i)
    if (!strcmp(argv[i], "--ngpu")) ngpu = atoi(argv[i + 1]);
  if (ngpu < 1 || ngpu > 8 || (ngpu & (ngpu - 1))) terminate("eigen_test: --ngpu must be 1, 2, 4 or 8", 1);
  if (ngpu == 1 && attach_external_ranks()) {
    ngpu = world_size();  // started by mpirun / srun / torchrun --no-python: one rank per B200, like the reference
    if (ngpu > 8 || (ngpu & (ngpu - 1))) terminate("eigen_test: the B200 solvers run on 1, 2, 4 or 8 ranks", 1);
  } else {
    launch_ranks(ngpu);
  }

  world_barrier();
  const double time_start = wtime();
  g_wtime_init = time_start;
  double time_start_part = time_start;

  read_command_argument(argc, argv, arg);
  arg.ngpu = ngpu;

  if (check_master()) {
    printf("---------- Eigen Test start ----------\n");
    printf("----- Configurations -----\n");
    print_command_argument(arg);
    printf("approximate required memory per process (Mbytes): %10.1f\n", required_memory(arg) / 1048576.0);
    printf("MPI processes: %d\n", world_size());
    printf("OpenMP threads per process (may be inaccurate): %u\n", std::thread::hardware_concurrency() / (unsigned)world_size());
    fflush(stdout);
  }

  double time_end = wtime();
  add_event("main:read_command_argument", time_end - time_start_part);
  time_start_part = time_end;

  validate_argument(arg);
  const int setting_g_block_size = g_block_size;  // fson_setting_add runs here, before --block-size is applied

  read_matrix_file(arg.matrix_A_filename, arg.matrix_A_info, matrix_A, ierr, arg.io_threads);
  if (ierr != 0) terminate("read_matrix_file " + arg.matrix_A_filename + " failed", ierr);
  if (arg.is_generalized_problem) {
    read_matrix_file(arg.matrix_B_filename, arg.matrix_B_info, matrix_B, ierr, arg.io_threads);
    if (ierr != 0) terminate("read_matrix_file " + arg.matrix_B_filename + " failed", ierr);
  }

  time_end = wtime();
  add_event("main:read_matrix_files", time_end - time_start_part);
  time_start_part = time_end;

  if (arg.is_dry_run) {
    if (check_master()) printf("\ndry run mode, exit\n");
    world_barrier();
    finalize_ranks();
    return 0;
  }

  time_end = wtime();
  add_event("main:bcast_sparse_matrices", time_end - time_start_part);  // nothing to broadcast: see the header
  time_start_part = time_end;

  if (check_master()) {
    printf("\n----- Solver Call -----\n");
    fflush(stdout);
  }
  eigen_solver(arg, matrix_A, eigenpairs, proc, arg.is_generalized_problem ? &matrix_B : nullptr);

  time_end = wtime();
  add_event("main:eigen_solver", time_end - time_start_part);
  time_start_part = time_end;

  // print eigenvalues, and eigenvectors if required
  if (check_master()) {
    FILE* f = fopen(arg.output_filename.c_str(), "w");
    if (!f) terminate("eigen_test: cannot open " + arg.output_filename, 1);
    for (int64_t j = 1; j <= arg.n_vec; ++j)
      fprintf(f, "%s %s\n", fortran_i(j, 8).c_str(), fortran_e(eigenpairs.blacs.values[(size_t)j - 1], 26, 16, 3).c_str());
    fclose(f);
  }
  if (arg.num_printed_vecs_ranges != 0) print_eigenvectors(arg, eigenpairs);

  time_end = wtime();
  add_event("main:print_eigenpairs", time_end - time_start_part);
  time_start_part = time_end;

  get_ipratios(arg, proc, eigenpairs, ipratios, arg.is_generalized_problem ? &matrix_B : nullptr);
  if (check_master()) {
    FILE* f = fopen(arg.ipratios_filename.c_str(), "w");
    if (!f) terminate("eigen_test: cannot open " + arg.ipratios_filename, 1);
    for (int64_t j = 1; j <= eigenpairs.blacs.desc[cols_]; ++j)
      fprintf(f, "%s %s\n", fortran_i(j, 8).c_str(), fortran_e(ipratios[(size_t)j - 1], 26, 16, 3).c_str());
    fclose(f);
  }

  time_end = wtime();
  add_event("main:compute_and_print_ipratios", time_end - time_start_part);
  time_start_part = time_end;

  if (arg.n_check_vec != 0) {
    if (check_master()) printf("\n----- Checker Call -----\n");
    eval_residual_norm(arg, matrix_A, eigenpairs, A_norm, rn_ave, rn_max, arg.is_generalized_problem ? &matrix_B : nullptr);
    if (check_master()) {
      printf("A norm: %s\n", fortran_e(A_norm, 15, 8, 0).c_str());
      printf("residual norm (average): %s\n", fortran_e(rn_ave, 15, 8, 0).c_str());
      printf("residual norm (max):     %s\n", fortran_e(rn_max, 15, 8, 0).c_str());
    }
  }

  time_end = wtime();
  add_event("main:eval_residual_norm", time_end - time_start_part);
  time_start_part = time_end;

  if (arg.ortho_check_index_start != 0) {
    eval_orthogonality(arg, eigenpairs, orthogonality, arg.is_generalized_problem ? &matrix_B : nullptr);
    if (check_master()) printf("orthogonality criterion: %s\n", fortran_e(orthogonality, 15, 8, 0).c_str());
  }

  time_end = wtime();
  add_event("main:eval_orthogonality", time_end - time_start_part);
  add_event("main", time_end - time_start);

  if (check_master()) {
    FILE* f = fopen(arg.log_filename.c_str(), "w");
    if (!f) terminate("eigen_test: cannot open " + arg.log_filename, 1);
    const int keep = g_block_size;
    g_block_size = setting_g_block_size;
    const std::string txt = log_json_text(arg, events());
    g_block_size = keep;
    fwrite(txt.data(), 1, txt.size(), f);
    fclose(f);
    fflush(stdout);
  }

  b200_finalize();
  world_barrier();
  finalize_ranks();
  return 0;
}

int main(int argc, char** argv) {
  try {
    return run(argc, argv);
  } catch (const Terminate& t) {
    // terminate (processes.f90:122-139): events first, then the message, then the whole job ends with the code
    if (check_master()) print_events();
    fprintf(stderr, t.code == 0 ? "[Info] %s\n" : "[Error] %s\n", t.message.c_str());
    fflush(stderr);
    abort_ranks();
    const int code = t.code == 0 ? 0 : ((t.code & 0xff) ? (t.code & 0xff) : 1);
    _Exit(code);
  }
}
