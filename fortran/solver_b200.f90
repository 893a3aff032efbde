! ek_solver_b200_m -- B200 (sm_100a) solvers for EigenKernel behind the solver boundary of
! src/solver_main.f90:52-99.  Thin ISO_C_BINDING glue over libekb200.so (include/ekb200.h); all arithmetic runs in
! hand-written CUDA.  House signature and error convention follow src/solver_scalapack_all.f90:127-168 and
! src/generalized_to_standard.f90:25-30; the dummy twin (solver_b200_dummy.f90) follows src/solver_elpa_dummy.f90.
!
! Deployment mode: one MPI rank (mpirun -np 1), GPUs selected by the library.  The BLACS grid is 1x1, so the
! local array of the type-2 eigenpairs IS the global matrix and main.f90's writers, get_ipratios and the verifier
! work unchanged.  NOTE: this file cannot be compiled in the development image (no Fortran toolchain); it is
! kept syntax-careful and uses only iso_c_binding scalars/arrays so that a maintainer can build it with
! `make WITH_B200=1` (see INTEGRATION.md).
module ek_solver_b200_m
  use, intrinsic :: iso_c_binding
  use ek_descriptor_parameters_m
  use ek_distribute_matrix_m, only : ek_process_t, setup_distributed_matrix
  use ek_eigenpairs_types_m, only : ek_eigenpairs_types_union_t
  use ek_event_logger_m, only : add_event
  use ek_matrix_io_m, only : ek_sparse_mat_t
  use ek_processes_m, only : check_master, terminate
  implicit none
  private
  public :: solve_with_b200, solve_with_general_b200

  interface
    integer(c_int) function ekb200_create(ctx, device) bind(C, name='ekb200_create')
      import :: c_ptr, c_int
      type(c_ptr), intent(out) :: ctx
      integer(c_int), value :: device
    end function ekb200_create
    integer(c_int) function ekb200_destroy(ctx) bind(C, name='ekb200_destroy')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function ekb200_destroy
    integer(c_int) function ekb200_sygvd_coo(ctx, n, nev, nnzA, ijA, vA, nnzB, ijB, vB, w, Z, ldz) &
         bind(C, name='ekb200_sygvd_coo')
      import :: c_ptr, c_int, c_int32_t, c_int64_t, c_double
      type(c_ptr), value :: ctx
      integer(c_int64_t), value :: n, nev, nnzA, nnzB, ldz
      integer(c_int32_t), intent(in) :: ijA(2, *), ijB(2, *)
      real(c_double), intent(in) :: vA(*), vB(*)
      real(c_double), intent(out) :: w(*), Z(ldz, *)
    end function ekb200_sygvd_coo
    integer(c_int) function ekb200_num_events(ctx) bind(C, name='ekb200_num_events')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function ekb200_num_events
    integer(c_int) function ekb200_get_event(ctx, i, name, seconds, num_repeated) bind(C, name='ekb200_get_event')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: i
      type(c_ptr), intent(out) :: name
      real(c_double), intent(out) :: seconds
      integer(c_int), intent(out) :: num_repeated
    end function ekb200_get_event
  end interface

contains

  ! Replays the library's CUDA-event timing table through add_event (src/event_logger.f90:23-65).
  subroutine replay_events(ctx)
    type(c_ptr), intent(in) :: ctx
    integer(c_int) :: i, n_ev, rep, ierr
    real(c_double) :: seconds
    type(c_ptr) :: cname
    character(kind=c_char), pointer :: chars(:)
    character(len=128) :: name
    integer :: k

    n_ev = ekb200_num_events(ctx)
    do i = 0, n_ev - 1
      ierr = ekb200_get_event(ctx, i, cname, seconds, rep)
      if (ierr /= 0) cycle
      call c_f_pointer(cname, chars, [128])
      name = ''
      do k = 1, 128
        if (chars(k) == c_null_char) exit
        name(k:k) = chars(k)
      end do
      call add_event(trim(name), seconds)
    end do
  end subroutine replay_events


  subroutine solve_b200_common(n, n_vec, proc, matrix_A, eigenpairs, matrix_B)
    integer, intent(in) :: n, n_vec
    type(ek_process_t), intent(in) :: proc
    type(ek_sparse_mat_t), intent(in) :: matrix_A
    type(ek_sparse_mat_t), intent(in), optional :: matrix_B
    type(ek_eigenpairs_types_union_t), intent(out) :: eigenpairs

    type(c_ptr) :: ctx
    integer(c_int) :: info
    integer(c_int64_t) :: nnzB
    integer(c_int32_t), allocatable :: ij_dummy(:, :)
    real(c_double), allocatable :: v_dummy(:)

    if (proc%n_procs_row /= 1 .or. proc%n_procs_col /= 1) then
      call terminate('solver_b200: run with one MPI rank (1x1 grid); the GPUs are driven by the library', 1)
    end if

    eigenpairs%type_number = 2
    allocate(eigenpairs%blacs%values(n))
    ! n x n_vec local array + live descriptor on the 1x1 grid (consumers call blacs_gridinfo on desc(context_))
    call setup_distributed_matrix('Eigenvectors', proc, n, n_vec, &
         eigenpairs%blacs%desc, eigenpairs%blacs%Vectors)

    info = ekb200_create(ctx, 0_c_int)
    if (info /= 0) then
      if (check_master()) print '("info(ekb200_create): ", i0)', info
      call terminate('solver_b200: no usable CUDA device (there is no CPU fallback)', info)
    end if

    if (present(matrix_B)) then
      info = ekb200_sygvd_coo(ctx, int(n, c_int64_t), int(n_vec, c_int64_t), &
           int(matrix_A%num_non_zeros, c_int64_t), matrix_A%suffix, matrix_A%value, &
           int(matrix_B%num_non_zeros, c_int64_t), matrix_B%suffix, matrix_B%value, &
           eigenpairs%blacs%values, eigenpairs%blacs%Vectors, &
           int(eigenpairs%blacs%desc(lld_), c_int64_t))
    else
      nnzB = 0
      allocate(ij_dummy(2, 1), v_dummy(1))
      info = ekb200_sygvd_coo(ctx, int(n, c_int64_t), int(n_vec, c_int64_t), &
           int(matrix_A%num_non_zeros, c_int64_t), matrix_A%suffix, matrix_A%value, &
           nnzB, ij_dummy, v_dummy, &
           eigenpairs%blacs%values, eigenpairs%blacs%Vectors, &
           int(eigenpairs%blacs%desc(lld_), c_int64_t))
    end if
    call replay_events(ctx)
    if (info /= 0) then
      ! same reporting as generalized_to_standard.f90:25-30
      if (check_master()) then
        if (present(matrix_B) .and. info > 0 .and. info <= n) then
          print '("info(pdpotrf): ", i0)', info
        else
          print '("info(ekb200_sygvd_coo): ", i0)', info
        end if
      end if
      call terminate('solver_b200: solve failed', info)
    end if
    info = ekb200_destroy(ctx)
  end subroutine solve_b200_common


  ! -s b200 / -s b200_select : standard problem (n_vec = n for all eigenpairs)
  subroutine solve_with_b200(n, n_vec, proc, matrix_A, eigenpairs)
    integer, intent(in) :: n, n_vec
    type(ek_process_t), intent(in) :: proc
    type(ek_sparse_mat_t), intent(in) :: matrix_A
    type(ek_eigenpairs_types_union_t), intent(out) :: eigenpairs
    call solve_b200_common(n, n_vec, proc, matrix_A, eigenpairs)
  end subroutine solve_with_b200


  ! -s general_b200 / -s general_b200_select : generalized problem
  subroutine solve_with_general_b200(n, n_vec, proc, matrix_A, eigenpairs, matrix_B)
    integer, intent(in) :: n, n_vec
    type(ek_process_t), intent(in) :: proc
    type(ek_sparse_mat_t), intent(in) :: matrix_A, matrix_B
    type(ek_eigenpairs_types_union_t), intent(out) :: eigenpairs
    call solve_b200_common(n, n_vec, proc, matrix_A, eigenpairs, matrix_B)
  end subroutine solve_with_general_b200
end module ek_solver_b200_m
