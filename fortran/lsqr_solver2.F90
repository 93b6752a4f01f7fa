!===============================================================================================
! lsqr_solver -- drop-in replacement of src/inversion/lsqr_solver2.F90.
! Same module name and public procedures (lsqr_solve_sensit :47-63, lsqr_solve :321-330,
! apply_soft_thresholding :478); the whole iteration loop runs on the GPU (csrc/lsqr.cu). The
! reductions the reference does with MPI_Allreduce (:214, :514) are NCCL all-reduces inside the
! library: call tfx_setup_comm once after MPI_Init (see INTEGRATION.md).
!===============================================================================================
module lsqr_solver

  use, intrinsic :: iso_c_binding
  use global_typedefs
  use mpi_tools, only: exit_MPI
  use sparse_matrix
  use tfx_c_api

  implicit none

  private

  public :: lsqr_solve, lsqr_solve_sensit
  public :: apply_soft_thresholding
  public :: tfx_setup_comm

contains

subroutine lsqr_solve_sensit(nlines, ncolumns, niter, rmin, gamma, target_misfit, &
                             matrix_sensit, matrix_cons, u, x, &
                             SOLVE_PROBLEM, nelements, nx, ny, nz, ncomponents, compression_type, WAVELET_DOMAIN, &
                             memory, myrank, nbproc)
  integer, intent(in) :: nlines, ncolumns, niter
  real(kind=CUSTOM_REAL), intent(in) :: rmin, gamma, target_misfit
  logical, intent(in) :: SOLVE_PROBLEM(2)
  integer, intent(in) :: nelements, nx, ny, nz, ncomponents, compression_type
  logical, intent(in) :: WAVELET_DOMAIN
  integer, intent(in) :: myrank, nbproc
  type(t_sparse_matrix), intent(in) :: matrix_sensit
  type(t_sparse_matrix), intent(in) :: matrix_cons
  real(kind=CUSTOM_REAL), intent(inout) :: u(nlines)
  real(kind=CUSTOM_REAL), intent(inout) :: x(ncolumns)
  real(kind=CUSTOM_REAL), intent(out) :: memory

  integer(c_int32_t) :: sp(2)

  sp = merge(1, 0, SOLVE_PROBLEM)
  call tfx_check(tfx_lsqr_solve_sensit(nlines, ncolumns, niter, real(rmin, c_double), real(gamma, c_double), &
                                       real(target_misfit, c_double), matrix_sensit%c_handle(), matrix_cons%c_handle(), &
                                       u, x, sp, nelements, nx, ny, nz, ncomponents, compression_type, &
                                       merge(1, 0, WAVELET_DOMAIN), memory, myrank, nbproc), myrank)
end subroutine lsqr_solve_sensit

subroutine lsqr_solve(nlines, nelements, niter, rmin, gamma, matrix, u, x, myrank)
  integer, intent(in) :: nlines, nelements, niter, myrank
  real(kind=CUSTOM_REAL), intent(in) :: rmin, gamma
  type(t_sparse_matrix), intent(in) :: matrix
  real(kind=CUSTOM_REAL), intent(inout) :: u(nlines)
  real(kind=CUSTOM_REAL), intent(inout) :: x(nelements)

  call tfx_check(tfx_lsqr_solve(nlines, nelements, niter, real(rmin, c_double), real(gamma, c_double), &
                                matrix%c_handle(), u, x, myrank), myrank)
end subroutine lsqr_solve

! Kept for callers outside the solver (same formula as reference :478-494; O(n) host loop).
pure subroutine apply_soft_thresholding(x, nelements, threshold)
  integer, intent(in) :: nelements
  real(kind=CUSTOM_REAL), intent(in) :: threshold
  real(kind=CUSTOM_REAL), intent(inout) :: x(nelements)

  where (abs(x) <= threshold)
    x = 0._CUSTOM_REAL
  elsewhere
    x = x - sign(threshold, x)
  end where
end subroutine apply_soft_thresholding

! One NCCL communicator over MPI_COMM_WORLD's ranks: rank 0 creates the id, MPI broadcasts it.
subroutine tfx_setup_comm(myrank, nbproc)
  use mpi
  integer, intent(in) :: myrank, nbproc
  character(kind=c_char) :: id(128)
  integer :: ierr

  call tfx_check(tfx_init(-1), myrank)
  if (nbproc == 1) return
  if (myrank == 0) call tfx_check(tfx_comm_unique_id(id), myrank)
  call MPI_Bcast(id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
  call tfx_check(tfx_comm_init(nbproc, myrank, id), myrank)
end subroutine tfx_setup_comm

end module lsqr_solver
