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

  ! struct tfx_sensit_params (include/tfx.h)
  type, bind(C) :: tfx_sensit_params
    integer(c_int32_t) :: problem_type
    integer(c_int32_t) :: nx, ny, nz
    integer(c_int32_t) :: ndata
    integer(c_int32_t) :: ndata_components
    integer(c_int32_t) :: nmodel_components
    integer(c_int32_t) :: data_type
    integer(c_int32_t) :: compression_type
    real(c_double) :: compression_rate
    real(c_double) :: problem_weight
    real(c_double) :: mi, md, theta, intensity
    integer(c_int32_t) :: cell0, ncells_local
    integer(c_int32_t) :: param_shift
    integer(c_int32_t) :: ncolumns
  end type tfx_sensit_params

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

    function tfx_take_latched_error() bind(C, name="tfx_take_latched_error") result(rc)
      import :: c_int
      integer(c_int) :: rc
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

    ! The reference declares the products `pure` (sparse_matrix.f90:298,313,373,388). A pure FUNCTION may not have an
    ! intent(inout) dummy (F2008 C1276), a pure SUBROUTINE may: the void C variants (include/tfx.h) latch their return
    ! code inside the library and tfx_check() collects it at the next non-pure call.
    pure subroutine tfx_sparse_matrix_mult_vector_v(m, x, b) bind(C, name="tfx_sparse_matrix_mult_vector_v")
      import :: c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
    end subroutine

    pure subroutine tfx_sparse_matrix_add_mult_vector_v(m, x, b) bind(C, name="tfx_sparse_matrix_add_mult_vector_v")
      import :: c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
    end subroutine

    function tfx_sparse_matrix_part_mult_vector(m, nelements, x, ndata, b, line_start, param_shift, myrank) &
        bind(C, name="tfx_sparse_matrix_part_mult_vector") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: m
      integer(c_int32_t), value :: nelements, ndata, line_start, param_shift, myrank
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
      integer(c_int) :: rc
    end function

    pure subroutine tfx_sparse_matrix_trans_mult_vector_v(m, x, b) bind(C, name="tfx_sparse_matrix_trans_mult_vector_v")
      import :: c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
    end subroutine

    pure subroutine tfx_sparse_matrix_add_trans_mult_vector_v(m, x, b) bind(C, name="tfx_sparse_matrix_add_trans_mult_vector_v")
      import :: c_double, c_ptr
      type(c_ptr), value :: m
      real(c_double), intent(in) :: x(*)
      real(c_double), intent(inout) :: b(*)
    end subroutine

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
    ! ---- sensitivity_gravmag (csrc/sensit.cu, sensit_dist.cu, sensit_io.cu) ---------------------
    function tfx_device_mem_info(free_bytes, total_bytes) bind(C, name="tfx_device_mem_info") result(rc)
      import :: c_int, c_int64_t
      integer(c_int64_t), intent(out) :: free_bytes, total_bytes
      integer(c_int) :: rc
    end function

    function tfx_calculate_sensit(matrix_sensit, par, X1, X2, Y1, Y2, Z1, Z2, data_X, data_Y, data_Z, &
                                  column_weight_full, data_weight, sensit_nnz, comp_error, nnz_total) &
        bind(C, name="tfx_calculate_sensit") result(rc)
      import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr, tfx_sensit_params
      type(c_ptr), intent(out) :: matrix_sensit
      type(tfx_sensit_params), intent(in) :: par
      real(c_double), intent(in) :: X1(*), X2(*), Y1(*), Y2(*), Z1(*), Z2(*), data_X(*), data_Y(*), data_Z(*)
      real(c_double), intent(in) :: column_weight_full(*), data_weight(*)
      integer(c_int32_t), intent(out) :: sensit_nnz(*)
      real(c_double), intent(out) :: comp_error
      integer(c_int64_t), intent(out) :: nnz_total
      integer(c_int) :: rc
    end function

    ! Keeps the grid on the device for the following assembly calls that pass the same arrays (both problems of a joint
    ! inversion read the same grid_full, sensitivity_gravmag.F90:82-177).
    function tfx_grid_pin(ncells, X1, X2, Y1, Y2, Z1, Z2) bind(C, name="tfx_grid_pin") result(rc)
      import :: c_int, c_int32_t, c_double
      integer(c_int32_t), value :: ncells
      real(c_double), intent(in) :: X1(*), X2(*), Y1(*), Y2(*), Z1(*), Z2(*)
      integer(c_int) :: rc
    end function

    function tfx_grid_unpin() bind(C, name="tfx_grid_unpin") result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function

    function tfx_sensit_assemble_rows(rows, par, X1, X2, Y1, Y2, Z1, Z2, data_X, data_Y, data_Z, &
                                      column_weight_full, data_weight, myrank, nbproc, sensit_nnz, comp_error, nnz_total) &
        bind(C, name="tfx_sensit_assemble_rows") result(rc)
      import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr, tfx_sensit_params
      type(c_ptr), intent(out) :: rows
      type(tfx_sensit_params), intent(in) :: par
      real(c_double), intent(in) :: X1(*), X2(*), Y1(*), Y2(*), Z1(*), Z2(*), data_X(*), data_Y(*), data_Z(*)
      real(c_double), intent(in) :: column_weight_full(*), data_weight(*)
      integer(c_int32_t), value :: myrank, nbproc
      integer(c_int32_t), intent(out) :: sensit_nnz(*)
      real(c_double), intent(out) :: comp_error
      integer(c_int64_t), intent(out) :: nnz_total
      integer(c_int) :: rc
    end function

    function tfx_sensit_rows_apply_weights(rows, problem_weight, data_weight) &
        bind(C, name="tfx_sensit_rows_apply_weights") result(rc)
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: rows
      real(c_double), value :: problem_weight
      real(c_double), intent(in) :: data_weight(*)
      integer(c_int) :: rc
    end function

    function tfx_sensit_rows_destroy(rows) bind(C, name="tfx_sensit_rows_destroy") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: rows
      integer(c_int) :: rc
    end function

    function tfx_get_load_balancing_nelements(nelements_total, sensit_nnz, nbproc, nnz_at_cpu_new, nelements_at_cpu_new) &
        bind(C, name="tfx_get_load_balancing_nelements") result(rc)
      import :: c_int, c_int32_t, c_int64_t
      integer(c_int32_t), value :: nelements_total, nbproc
      integer(c_int32_t), intent(in) :: sensit_nnz(*)
      integer(c_int64_t), intent(out) :: nnz_at_cpu_new(*)
      integer(c_int32_t), intent(out) :: nelements_at_cpu_new(*)
      integer(c_int) :: rc
    end function

    function tfx_sensit_repartition_into(matrix_sensit, rows, problem_slot, nelements_at_cpu, myrank, nbproc) &
        bind(C, name="tfx_sensit_repartition_into") result(rc)
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr), value :: matrix_sensit, rows
      integer(c_int32_t), value :: problem_slot, myrank, nbproc
      integer(c_int32_t), intent(in) :: nelements_at_cpu(*)
      integer(c_int) :: rc
    end function

    function tfx_create_sensit_directory(dir) bind(C, name="tfx_create_sensit_directory") result(rc)
      import :: c_int, c_char
      character(kind=c_char), intent(in) :: dir(*)
      integer(c_int) :: rc
    end function

    function tfx_write_sensit_file(rows, dir) bind(C, name="tfx_write_sensit_file") result(rc)
      import :: c_int, c_char, c_ptr
      type(c_ptr), value :: rows
      character(kind=c_char), intent(in) :: dir(*)
      integer(c_int) :: rc
    end function

    function tfx_write_sensit_metadata(par, dir, nbproc, depth_weighting_type, comp_error, nnz_total, sensit_nnz) &
        bind(C, name="tfx_write_sensit_metadata") result(rc)
      import :: c_int, c_int32_t, c_int64_t, c_double, c_char, tfx_sensit_params
      type(tfx_sensit_params), intent(in) :: par
      character(kind=c_char), intent(in) :: dir(*)
      integer(c_int32_t), value :: nbproc, depth_weighting_type
      real(c_double), value :: comp_error
      integer(c_int64_t), value :: nnz_total
      integer(c_int32_t), intent(in) :: sensit_nnz(*)
      integer(c_int) :: rc
    end function

    function tfx_read_sensitivity_metadata(par, dir, depth_weighting_type, nbproc_sensit, comp_error, nnz_total) &
        bind(C, name="tfx_read_sensitivity_metadata") result(rc)
      import :: c_int, c_int32_t, c_int64_t, c_double, c_char, tfx_sensit_params
      type(tfx_sensit_params), intent(in) :: par
      character(kind=c_char), intent(in) :: dir(*)
      integer(c_int32_t), value :: depth_weighting_type
      integer(c_int32_t), intent(out) :: nbproc_sensit
      real(c_double), intent(out) :: comp_error
      integer(c_int64_t), intent(out) :: nnz_total
      integer(c_int) :: rc
    end function

    function tfx_read_sensit_nnz(par, dir, sensit_nnz) bind(C, name="tfx_read_sensit_nnz") result(rc)
      import :: c_int, c_int32_t, c_char, tfx_sensit_params
      type(tfx_sensit_params), intent(in) :: par
      character(kind=c_char), intent(in) :: dir(*)
      integer(c_int32_t), intent(out) :: sensit_nnz(*)
      integer(c_int) :: rc
    end function

    function tfx_write_depth_weight(par, dir, column_weight_full) bind(C, name="tfx_write_depth_weight") result(rc)
      import :: c_int, c_double, c_char, tfx_sensit_params
      type(tfx_sensit_params), intent(in) :: par
      character(kind=c_char), intent(in) :: dir(*)
      real(c_double), intent(in) :: column_weight_full(*)
      integer(c_int) :: rc
    end function

    function tfx_read_depth_weight(par, dir, column_weight_full) bind(C, name="tfx_read_depth_weight") result(rc)
      import :: c_int, c_double, c_char, tfx_sensit_params
      type(tfx_sensit_params), intent(in) :: par
      character(kind=c_char), intent(in) :: dir(*)
      real(c_double), intent(out) :: column_weight_full(*)
      integer(c_int) :: rc
    end function

    function tfx_read_sensitivity_kernel_into(matrix_sensit, par, dir, data_weight, depth_weighting_type, problem_slot, &
                                              myrank, nbproc, nelements_at_cpu, nnz_local) &
        bind(C, name="tfx_read_sensitivity_kernel_into") result(rc)
      import :: c_int, c_int32_t, c_int64_t, c_double, c_char, c_ptr, tfx_sensit_params
      type(c_ptr), value :: matrix_sensit
      type(tfx_sensit_params), intent(in) :: par
      character(kind=c_char), intent(in) :: dir(*)
      real(c_double), intent(in) :: data_weight(*)
      integer(c_int32_t), value :: depth_weighting_type, problem_slot, myrank, nbproc
      integer(c_int32_t), intent(in) :: nelements_at_cpu(*)
      integer(c_int64_t), intent(out) :: nnz_local
      integer(c_int) :: rc
    end function

    ! ---- t_model%calculate_data (model.F90:220-307), calculate_depth_weight (weights_gravmag.f90:46-199) ----
    function tfx_calculate_data(matrix_sensit, nelements, ncomponents, model_val, ndata, ndata_components, &
                                problem_weight, column_weight, data_weight, data_calc, compression_type, &
                                nx, ny, nz, line_start, param_shift, myrank, nbproc) &
        bind(C, name="tfx_calculate_data") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: matrix_sensit
      integer(c_int32_t), value :: nelements, ncomponents, ndata, ndata_components
      real(c_double), intent(in) :: model_val(*), column_weight(*), data_weight(*)
      real(c_double), value :: problem_weight
      real(c_double), intent(out) :: data_calc(*)
      integer(c_int32_t), value :: compression_type, nx, ny, nz, line_start, param_shift, myrank, nbproc
      integer(c_int) :: rc
    end function

    function tfx_calculate_depth_weight(depth_weighting_type, depth_weighting_power, depth_weighting_beta, Z0, &
                                        nelements_total, X1, X2, Y1, Y2, Z1, Z2, ndata, data_X, data_Y, data_Z, &
                                        nsmaller, nelements, column_weight, myrank, nbproc) &
        bind(C, name="tfx_calculate_depth_weight") result(rc)
      import :: c_int, c_int32_t, c_double
      integer(c_int32_t), value :: depth_weighting_type, nelements_total, ndata, nsmaller, nelements, myrank, nbproc
      real(c_double), value :: depth_weighting_power, depth_weighting_beta, Z0
      real(c_double), intent(in) :: X1(*), X2(*), Y1(*), Y2(*), Z1(*), Z2(*), data_X(*), data_Y(*), data_Z(*)
      real(c_double), intent(out) :: column_weight(*)
      integer(c_int) :: rc
    end function

    ! ---- constraint-matrix producers (csrc/cons.cu) ----
    function tfx_damping_add(matrix, nrows, b_RHS, alpha, problem_weight, norm_power, compression_type, nx, ny, nz, &
                             nelements, column_weight, model, model_ref, param_shift, wavelet_domain, local_weight, &
                             myrank, nbproc, cost) bind(C, name="tfx_damping_add") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: matrix
      integer(c_int32_t), value :: nrows, compression_type, nx, ny, nz, nelements, param_shift, wavelet_domain
      integer(c_int32_t), value :: myrank, nbproc
      real(c_double), intent(inout) :: b_RHS(*)
      real(c_double), value :: alpha, problem_weight, norm_power
      real(c_double), intent(in) :: column_weight(*), model(*), model_ref(*)
      type(c_ptr), value :: local_weight          ! c_loc(local_weight) or c_null_ptr when absent
      real(c_double), intent(out) :: cost
      integer(c_int) :: rc
    end function

    function tfx_damping_gradient_add(matrix, nrows, b_RHS, beta, problem_weight, nx, ny, nz, dX, dY, dZ, nelements, &
                                      val_full, column_weight, local_weight, param_shift, direction, myrank, nbproc, &
                                      cost) bind(C, name="tfx_damping_gradient_add") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: matrix
      integer(c_int32_t), value :: nrows, nx, ny, nz, nelements, param_shift, direction, myrank, nbproc
      real(c_double), intent(inout) :: b_RHS(*)
      real(c_double), value :: beta, problem_weight
      real(c_double), intent(in) :: dX(*), dY(*), dZ(*), val_full(*), column_weight(*), local_weight(*)
      real(c_double), intent(out) :: cost
      integer(c_int) :: rc
    end function

    function tfx_cross_gradient_calculate(matrix, nrows, b_RHS, nx, ny, nz, dX, dY, dZ, nparams_loc, model1, model2, &
                                          column_weight1, column_weight2, der_type, glob_weight, &
                                          keep_model_constant, myrank, nbproc, cost, cross_grad) &
        bind(C, name="tfx_cross_gradient_calculate") result(rc)
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr), value :: matrix
      integer(c_int32_t), value :: nrows, nx, ny, nz, nparams_loc, der_type, myrank, nbproc
      real(c_double), intent(inout) :: b_RHS(*)
      real(c_double), intent(in) :: dX(*), dY(*), dZ(*), model1(*), model2(*), column_weight1(*), column_weight2(*)
      real(c_double), value :: glob_weight
      integer(c_int32_t), intent(in) :: keep_model_constant(2)
      real(c_double), intent(out) :: cost(3)
      type(c_ptr), value :: cross_grad            ! c_loc(this%cross_grad) on rank 0, c_null_ptr elsewhere
      integer(c_int) :: rc
    end function

    function tfx_rescale_model(nelements, ncomponents, model, weight) bind(C, name="tfx_rescale_model") result(rc)
      import :: c_int, c_int32_t, c_double
      integer(c_int32_t), value :: nelements, ncomponents
      real(c_double), intent(inout) :: model(*)
      real(c_double), intent(in) :: weight(*)
      integer(c_int) :: rc
    end function

    function tfx_model_update(nelements, ncomponents, val, delta_model) bind(C, name="tfx_model_update") result(rc)
      import :: c_int, c_int32_t, c_double
      integer(c_int32_t), value :: nelements, ncomponents
      real(c_double), intent(inout) :: val(*)
      real(c_double), intent(in) :: delta_model(*)
      integer(c_int) :: rc
    end function

    function tfx_admm_iterate_admm_arrays(nelements, nlithos, xmin, xmax, x, z, u, x0) &
        bind(C, name="tfx_admm_iterate_admm_arrays") result(rc)
      import :: c_int, c_int32_t, c_double
      integer(c_int32_t), value :: nelements, nlithos
      real(c_double), intent(in) :: xmin(*), xmax(*), x(*)
      real(c_double), intent(inout) :: z(*), u(*)
      real(c_double), intent(out) :: x0(*)
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
    integer(c_int) :: code
    integer :: i

    ! A product called from a `pure` procedure cannot stop the run: its failure is latched in the library and
    ! reported here, by the next non-pure call (it comes first: it happened first).
    code = tfx_take_latched_error()
    if (code == 0) code = rc
    if (code == 0) return
    msg = ""
    call c_f_pointer(tfx_last_error(), cmsg, [512])
    do i = 1, 512
      if (cmsg(i) == c_null_char) exit
      msg(i:i) = cmsg(i)
    enddo
    call exit_MPI(trim(msg), myrank, int(code))
  end subroutine tfx_check

end module tfx_c_api
