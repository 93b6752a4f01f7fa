!===============================================================================================
! damping -- drop-in replacement of src/inversion/damping.F90 (type t_damping).
!
! Same public type and type-bound names / argument lists (damping.F90:31-62,67-68,97-98,239): the caller
! joint_inverse_problem.F90:452-470,509-527 compiles unchanged. add() forwards to tfx_damping_add, which appends
! the nx*ny*nz diagonal rows to matrix_cons on the GPU and fills b_RHS(row_beg:row_end) (csrc/cons.cu).
!===============================================================================================
module damping

  use, intrinsic :: iso_c_binding
  use global_typedefs
  use sparse_matrix
  use tfx_c_api

  implicit none

  private

  type, public :: t_damping
    private
    integer :: nelements, nelements_total
    real(kind=CUSTOM_REAL) :: alpha, problem_weight, norm_power, cost
    integer :: compression_type, nx, ny, nz
  contains
    private
    procedure, public, pass :: initialize => damping_initialize
    procedure, public, pass :: add => damping_add
    procedure, public, pass :: get_cost => damping_get_cost
  end type t_damping

contains

subroutine damping_initialize(this, nelements, alpha, problem_weight, norm_power, compression_type, nx, ny, nz)
  class(t_damping), intent(inout) :: this
  real(kind=CUSTOM_REAL), intent(in) :: alpha, problem_weight, norm_power
  integer, intent(in) :: nelements, compression_type, nx, ny, nz

  this%nelements = nelements
  this%alpha = alpha
  this%problem_weight = problem_weight
  this%norm_power = norm_power
  this%compression_type = compression_type
  this%nx = nx; this%ny = ny; this%nz = nz
  this%nelements_total = nx * ny * nz
  this%cost = 0.d0
end subroutine damping_initialize

subroutine damping_add(this, matrix, nrows, b_RHS, column_weight, model, model_ref, param_shift, &
                       WAVELET_DOMAIN, myrank, nbproc, local_weight)
  class(t_damping), intent(inout) :: this
  type(t_sparse_matrix), intent(inout) :: matrix
  integer, intent(in) :: nrows, param_shift, myrank, nbproc
  real(kind=CUSTOM_REAL), intent(inout) :: b_RHS(nrows)
  real(kind=CUSTOM_REAL), intent(in) :: column_weight(this%nelements), model(this%nelements), model_ref(this%nelements)
  logical, intent(in) :: WAVELET_DOMAIN
  real(kind=CUSTOM_REAL), optional, target, intent(in) :: local_weight(this%nelements)
  type(c_ptr) :: lw
  integer(c_int32_t) :: wd

  lw = c_null_ptr
  if (present(local_weight)) lw = c_loc(local_weight)
  wd = merge(1, 0, WAVELET_DOMAIN)
  call tfx_check(tfx_damping_add(matrix%c_handle(), nrows, b_RHS, this%alpha, this%problem_weight, this%norm_power, &
                                 this%compression_type, this%nx, this%ny, this%nz, this%nelements, column_weight, &
                                 model, model_ref, param_shift, wd, lw, myrank, nbproc, this%cost), myrank)
end subroutine damping_add

pure function damping_get_cost(this) result(res)
  class(t_damping), intent(in) :: this
  real(kind=CUSTOM_REAL) :: res
  res = this%cost
end function damping_get_cost

end module damping
