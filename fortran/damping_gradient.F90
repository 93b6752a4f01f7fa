!===============================================================================================
! damping_gradient -- drop-in replacement of src/inversion/damping_gradient.F90 (type t_damping_gradient).
!
! Same public names / argument lists (damping_gradient.F90:35-58,63,84,93-94); the caller
! joint_inverse_problem.F90:473-492 compiles unchanged. add() forwards to tfx_damping_gradient_add (csrc/cons.cu).
!===============================================================================================
module damping_gradient

  use, intrinsic :: iso_c_binding
  use global_typedefs
  use sparse_matrix
  use model
  use grid
  use tfx_c_api

  implicit none

  private

  type, public :: t_damping_gradient
    private
    integer :: nx, ny, nz, nelements, nelements_total
    real(kind=CUSTOM_REAL) :: beta, problem_weight, cost
  contains
    private
    procedure, public, pass :: initialize => damping_gradient_initialize
    procedure, public, pass :: add => damping_gradient_add
    procedure, public, pass :: get_cost => damping_gradient_get_cost
  end type t_damping_gradient

contains

subroutine damping_gradient_initialize(this, beta, problem_weight, nx, ny, nz, nelements)
  class(t_damping_gradient), intent(inout) :: this
  real(kind=CUSTOM_REAL), intent(in) :: beta, problem_weight
  integer, intent(in) :: nx, ny, nz, nelements

  this%beta = beta
  this%problem_weight = problem_weight
  this%nx = nx; this%ny = ny; this%nz = nz
  this%nelements = nelements
  this%nelements_total = nx * ny * nz
  this%cost = 0.d0
end subroutine damping_gradient_initialize

pure function damping_gradient_get_cost(this) result(res)
  class(t_damping_gradient), intent(in) :: this
  real(kind=CUSTOM_REAL) :: res
  res = this%cost
end function damping_gradient_get_cost

subroutine damping_gradient_add(this, model, grid, column_weight, local_weight, matrix, nrows, &
                                b_RHS, param_shift, direction, icomp, myrank, nbproc)
  class(t_damping_gradient), intent(inout) :: this
  type(t_model), intent(in) :: model
  type(t_grad_grid), intent(in) :: grid
  real(kind=CUSTOM_REAL), intent(in) :: column_weight(:), local_weight(:)
  integer, intent(in) :: param_shift, nrows, direction, icomp, myrank, nbproc
  type(t_sparse_matrix), intent(inout) :: matrix
  real(kind=CUSTOM_REAL), intent(inout) :: b_RHS(nrows)

  call tfx_check(tfx_damping_gradient_add(matrix%c_handle(), nrows, b_RHS, this%beta, this%problem_weight, &
                                          this%nx, this%ny, this%nz, grid%dX, grid%dY, grid%dZ, this%nelements, &
                                          model%val_full(:, icomp), column_weight, local_weight, param_shift, &
                                          direction, myrank, nbproc, this%cost), myrank)
end subroutine damping_gradient_add

end module damping_gradient
