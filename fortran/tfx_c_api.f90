!===============================================================================================
! tfx_c_api -- ISO_C_BINDING interfaces of libtfx (include/tfx.h), one per C entry point.
!
! This file and its siblings in fortran/ are the reference-side binding: same-named replacement
! modules (sparse_matrix, wavelet_transform, wavelet_utils, lsqr_solver) whose bodies forward to the
! C ABI, so that problem_joint_gravmag.F90 / joint_inverse_problem.F90 / model.F90 compile unchanged.
! They are NOT compiled in the development image (no Fortran compiler there); INTEGRATION.md shows
! the Makefile edits. All logic lives behind the C ABI, these modules only marshal arguments.
!===============================================================================================
module tfx_c_api

  use, intrinsic :: iso_c_binding
  use mpi_tools, only: exit_MPI

  implicit none

  public

  interface
    function tfx_init(device) bind(C, name="tfx_init") result(rc)
      import :: c_int
      integer(c_int), value :: device
      integer(c_int) :: rc
    end function

    function tfx_finalize() bind(C, name="tfx_finalize") result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function

    function tfx_last_error() bind(C, name="tfx_last_error") result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function

    function tfx_comm_unique_id(id) bind(C, name="tfx_comm_unique_id") result(rc)
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id(128)
      integer(c_int) :: rc
    end function

    function tfx_comm_init(nranks, rank, id) bind(C, name="tfx_comm_init") result(rc)
      import :: c_int, c_char
      integer(c_int), value :: nranks, rank
      character(kind=c_char), intent(in) :: id(128)
      integer(c_int) :: rc
    end function

    function tfx_comm_allreduce_sum(buf, count) bind(C, name="tfx_comm_allreduce_sum") result(rc)
      import :: c_int, c_int64_t, c_double
      real(c_double), intent(inout) :: buf(*)
      integer(c_int64_t), value :: count
      integer(c_int) :: rc
    end function

    ! ---- sparse_matrix -------------------------------------------------------------------------
    function tfx_sparse_matrix_initialize(m, nl, ncolumns, nnz, myrank, nl_empty) &
        bind(C, name="tfx_sparse_matrix_initialize") result(rc)
      import :: c_int, c_int32_t, c_int64_t, c_ptr
      type(c_ptr), intent(out) :: m
      integer(c_int32_t), value :: nl, ncolumns, myrank, nl_empty
      integer(c_int64_t), value :: nnz
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_destroy(m) bind(C, name="tfx_sparse_matrix_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_reset(m) bind(C, name="tfx_sparse_matrix_reset") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: m
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_finalize(m, myrank) bind(C, name="tfx_sparse_matrix_finalize") result(rc)
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t), value :: myrank
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_add(m, val, column, myrank) bind(C, name="tfx_sparse_matrix_add") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), value :: val
      integer(c_int32_t), value :: column, myrank
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_add_row(m, nel_add, values, columns, myrank) &
        bind(C, name="tfx_sparse_matrix_add_row") result(rc)
      import :: c_int, c_int32_t, c_float, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t), value :: nel_add, myrank
      real(c_float), intent(in) :: values(*)
      integer(c_int32_t), intent(in) :: columns(*)
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_new_row(m, myrank) bind(C, name="tfx_sparse_matrix_new_row") result(rc)
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t), value :: myrank
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_add_empty_rows(m, nrows, myrank) &
        bind(C, name="tfx_sparse_matrix_add_empty_rows") result(rc)
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t), value :: nrows, myrank
      integer(c_int) :: rc
    end function

    ! The reference declares the products `pure` (sparse_matrix.f90:298,313,373,388); so are these.
    pure function tfx_sparse_matrix_mult_vector(m, x, b) bind(C, name="tfx_sparse_matrix_mult_vector") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
      integer(c_int) :: rc
    end function

    pure function tfx_sparse_matrix_add_mult_vector(m, x, b) &
        bind(C, name="tfx_sparse_matrix_add_mult_vector") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_part_mult_vector(m, nelements, x, ndata, b, line_start, param_shift, myrank) &
        bind(C, name="tfx_sparse_matrix_part_mult_vector") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t), value :: nelements, ndata, line_start, param_shift, myrank
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
      integer(c_int) :: rc
    end function

    pure function tfx_sparse_matrix_trans_mult_vector(m, x, b) &
        bind(C, name="tfx_sparse_matrix_trans_mult_vector") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
      integer(c_int) :: rc
    end function

    pure function tfx_sparse_matrix_add_trans_mult_vector(m, x, b) &
        bind(C, name="tfx_sparse_matrix_add_trans_mult_vector") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
      integer(c_int) :: rc
    end function

    function tfx_sparse_matrix_normalize_columns(m, column_norm) &
        bind(C, name="tfx_sparse_matrix_normalize_columns") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(out) :: column_norm(*)
      integer(c_int) :: rc
    end function

    pure function tfx_sparse_matrix_get_total_row_number(m) &
        bind(C, name="tfx_sparse_matrix_get_total_row_number") result(n)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t) :: n
    end function

    pure function tfx_sparse_matrix_get_current_row_number(m) &
        bind(C, name="tfx_sparse_matrix_get_current_row_number") result(n)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t) :: n
    end function

    pure function tfx_sparse_matrix_get_ncolumns(m) bind(C, name="tfx_sparse_matrix_get_ncolumns") result(n)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t) :: n
    end function

    pure function tfx_sparse_matrix_get_number_elements(m) &
        bind(C, name="tfx_sparse_matrix_get_number_elements") result(n)
      import :: c_int64_t, c_ptr
      type(c_ptr), value :: m
      integer(c_int64_t) :: n
    end function

    pure function tfx_sparse_matrix_get_nnz(m) bind(C, name="tfx_sparse_matrix_get_nnz") result(n)
      import :: c_int64_t, c_ptr
      type(c_ptr), value :: m
      integer(c_int64_t) :: n
    end function

    ! ---- wavelet_transform ---------------------------------------------------------------------
    function tfx_forward_wavelet(s, n1, n2, n3, wavelet_type) bind(C, name="tfx_forward_wavelet") result(rc)
      import :: c_int, c_int32_t, c_double
      real(c_double), intent(inout) :: s(*)
      integer(c_int32_t), value :: n1, n2, n3, wavelet_type
      integer(c_int) :: rc
    end function

    function tfx_inverse_wavelet(s, n1, n2, n3, wavelet_type) bind(C, name="tfx_inverse_wavelet") result(rc)
      import :: c_int, c_int32_t, c_double
      real(c_double), intent(inout) :: s(*)
      integer(c_int32_t), value :: n1, n2, n3, wavelet_type
      integer(c_int) :: rc
    end function

    ! ---- lsqr_solver ---------------------------------------------------------------------------
    function tfx_lsqr_solve(nlines, nelements, niter, rmin, gamma, matrix, u, x, myrank) &
        bind(C, name="tfx_lsqr_solve") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      integer(c_int32_t), value :: nlines, nelements, niter, myrank
      real(c_double), value :: rmin, gamma
      type(c_ptr), value :: matrix
      real(c_double), intent(inout) :: u(*), x(*)
      integer(c_int) :: rc
    end function

    function tfx_lsqr_solve_sensit(nlines, ncolumns, niter, rmin, gamma, target_misfit, matrix_sensit, matrix_cons, &
                                   u, x, solve_problem, nelements, nx, ny, nz, ncomponents, compression_type, &
                                   wavelet_domain, memory, myrank, nbproc) bind(C, name="tfx_lsqr_solve_sensit") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      integer(c_int32_t), value :: nlines, ncolumns, niter, nelements, nx, ny, nz, ncomponents
      integer(c_int32_t), value :: compression_type, wavelet_domain, myrank, nbproc
      real(c_double), value :: rmin, gamma, target_misfit
      type(c_ptr), value :: matrix_sensit, matrix_cons
      real(c_double), intent(inout) :: u(*), x(*)
      integer(c_int32_t), intent(in) :: solve_problem(2)
      real(c_double), intent(out) :: memory
      integer(c_int) :: rc
    end function
  end interface

contains

  !---------------------------------------------------------------------------------------------
  ! Fatal-abort convention of the reference: a failed C call ends the run through exit_MPI with the
  ! library's message (the text the reference itself would have printed).
  !---------------------------------------------------------------------------------------------
  subroutine tfx_check(rc, myrank)
    integer(c_int), intent(in) :: rc
    integer, intent(in) :: myrank
    character(kind=c_char), pointer :: cmsg(:)
    character(len=512) :: msg
    integer :: i

    if (rc == 0) return
    msg = ""
    call c_f_pointer(tfx_last_error(), cmsg, [512])
    do i = 1, 512
      if (cmsg(i) == c_null_char) exit
      msg(i:i) = cmsg(i)
    enddo
    call exit_MPI(trim(msg), myrank, int(rc))
  end subroutine tfx_check

  ! `pure` callers cannot stop the run; they record the code and the next non-pure call reports it.
  pure subroutine tfx_ignore(rc)
    integer(c_int), intent(in) :: rc
    if (rc /= 0) error stop "libtfx: product failed (see tfx_last_error)"
  end subroutine tfx_ignore

end module tfx_c_api
