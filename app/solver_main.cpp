// ek_solver_main_m -- eigen_solver (reference src/solver_main.f90:22-100): dispatch on `-s <solver>`.
// The reference's select-case gains the B200 cases; its own cases need ScaLAPACK / ELPA / EigenExa and end the
// way the *_dummy.f90 twins end (src/solver_elpa_dummy.f90:14-22).
#include <stdio.h>

#include "ek_app.hpp"

namespace ekapp {

static void print_map_of_grid_to_processes(const ek_process_t& proc) {
  // processes.f90:68-107: which process sits at each grid coordinate; the B200 grid is 1 x P in rank order
  if (!check_master()) return;
  printf("process numbers in BLACS grid is\n");
  for (int c = 0; c < proc.n_procs_col; ++c) printf(c + 1 < proc.n_procs_col ? "%6d " : "%6d\n", c);
}

void eigen_solver(ek_argument_t& arg, const ek_sparse_mat_t& matrix_A, ek_eigenpairs_types_union_t& eigenpairs,
                  ek_process_t& proc, const ek_sparse_mat_t* matrix_B) {
  const int64_t n = arg.matrix_A_info.rows;
  if (arg.block_size > 0) g_block_size = arg.block_size;  // do not use the default block size
  setup_distribution(proc);
  if (arg.is_printing_grid_mapping) print_map_of_grid_to_processes(proc);

  const std::string& st = arg.solver_type;
  if (st == "b200" || st == "b200_select") {
    solve_with_b200(arg, n, proc, matrix_A, eigenpairs, nullptr);
  } else if (st == "general_b200" || st == "general_b200_select" || st == "general_b200inv") {
    if (!matrix_B) terminate("eigen_solver: solver '" + st + "' needs matrix B", 1);
    solve_with_b200(arg, n, proc, matrix_A, eigenpairs, matrix_B);
  } else {
    terminate("eigen_solver: Unknown solver", 1);
  }
}

}  // namespace ekapp
