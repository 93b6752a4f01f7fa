!===============================================================================================
! admm_method -- drop-in replacement of src/inversion/admm_method.F90 (type t_admm_method).
!
! Same public type, fields (nelements, z, u) and procedure names / argument lists (admm_method.F90:30-42,49,70);
! iterate_admm_arrays forwards to tfx_admm_iterate_admm_arrays (csrc/cons.cu).
!===============================================================================================
module admm_method

  use, intrinsic :: iso_c_binding
  use global_typedefs
  use mpi_tools, only: exit_MPI
  use tfx_c_api

  implicit none

  private

  type, public :: t_admm_method
    integer :: nelements
    real(kind=CUSTOM_REAL), allocatable :: z(:)
    real(kind=CUSTOM_REAL), allocatable :: u(:)
  contains
    private
    procedure, public, pass :: initialize => admm_method_initialize
    procedure, public, pass :: iterate_admm_arrays => admm_method_iterate_admm_arrays
  end type t_admm_method

contains

subroutine admm_method_initialize(this, nelements, myrank)
  class(t_admm_method), intent(inout) :: this
  integer, intent(in) :: nelements, myrank
  integer :: ierr

  this%nelements = nelements
  allocate(this%z(nelements), source=0._CUSTOM_REAL, stat=ierr)
  if (ierr == 0) allocate(this%u(nelements), source=0._CUSTOM_REAL, stat=ierr)
  if (ierr /= 0) call exit_MPI("Dynamic memory allocation error in admm_method_initialize!", myrank, ierr)
end subroutine admm_method_initialize

subroutine admm_method_iterate_admm_arrays(this, nlithos, xmin, xmax, x, x0)
  class(t_admm_method), intent(inout) :: this
  integer, intent(in) :: nlithos
  real(kind=CUSTOM_REAL), intent(in) :: xmin(nlithos, this%nelements), xmax(nlithos, this%nelements)
  real(kind=CUSTOM_REAL), intent(in) :: x(this%nelements)
  real(kind=CUSTOM_REAL), intent(out) :: x0(this%nelements)

  call tfx_check(tfx_admm_iterate_admm_arrays(this%nelements, nlithos, xmin, xmax, x, this%z, this%u, x0), 0)
end subroutine admm_method_iterate_admm_arrays

end module admm_method
