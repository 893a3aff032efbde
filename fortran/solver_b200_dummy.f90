! Build-time stub selected when WITH_B200 is not set (pattern of src/solver_elpa_dummy.f90:14-22).
module ek_solver_b200_m
  use ek_distribute_matrix_m, only : ek_process_t
  use ek_eigenpairs_types_m, only : ek_eigenpairs_types_union_t
  use ek_matrix_io_m, only : ek_sparse_mat_t
  use ek_processes_m, only : terminate
  implicit none
  private
  public :: solve_with_b200, solve_with_general_b200
contains
  subroutine solve_with_b200(n, n_vec, proc, matrix_A, eigenpairs)
    integer, intent(in) :: n, n_vec
    type(ek_process_t), intent(in) :: proc
    type(ek_sparse_mat_t), intent(in) :: matrix_A
    type(ek_eigenpairs_types_union_t), intent(out) :: eigenpairs
    call terminate('solver_b200: B200 solvers are not supported in this build', 1)
  end subroutine solve_with_b200
  subroutine solve_with_general_b200(n, n_vec, proc, matrix_A, eigenpairs, matrix_B, reduction)
    integer, intent(in) :: n, n_vec
    type(ek_process_t), intent(in) :: proc
    type(ek_sparse_mat_t), intent(in) :: matrix_A, matrix_B
    type(ek_eigenpairs_types_union_t), intent(out) :: eigenpairs
    integer, intent(in), optional :: reduction
    call terminate('solver_b200: B200 solvers are not supported in this build', 1)
  end subroutine solve_with_general_b200
end module ek_solver_b200_m
