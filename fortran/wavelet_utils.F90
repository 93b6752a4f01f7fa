!===============================================================================================
! wavelet_utils -- drop-in replacement of src/inversion/wavelet_utils.F90 (apply_wavelet_transform,
! reference :37-72). With one rank the component volumes are transformed in place on the GPU; with
! several ranks the reference's gather-to-rank-0 / scatter pattern is kept (parallel_tools) and the
! transform of the gathered volume runs on rank 0's GPU.
!===============================================================================================
module wavelet_utils

  use global_typedefs
  use parallel_tools
  use wavelet_transform

  implicit none

  private

  public :: apply_wavelet_transform

contains

subroutine apply_wavelet_transform(nelements, nx, ny, nz, ncomponents, v, model_full, FWD, &
                                   compression_type, nproblems, SOLVE_PROBLEM, myrank, nbproc)
  integer, intent(in) :: nelements, nx, ny, nz, ncomponents
  logical, intent(in) :: FWD
  integer, intent(in) :: compression_type, nproblems
  logical, intent(in) :: SOLVE_PROBLEM(nproblems)
  integer, intent(in) :: myrank, nbproc
  real(kind=CUSTOM_REAL), intent(inout) :: model_full(nx * ny * nz)
  real(kind=CUSTOM_REAL), intent(inout) :: v(nelements, ncomponents, nproblems)

  integer :: ip, ic

  do ip = 1, nproblems
    if (.not. SOLVE_PROBLEM(ip)) cycle
    do ic = 1, ncomponents
      if (nbproc == 1) then
        call transform(v(:, ic, ip))
      else
        call get_full_array(v(:, ic, ip), nelements, model_full, .false., myrank, nbproc)
        if (myrank == 0) call transform(model_full)
        call scatter_full_array(nelements, model_full, v(:, ic, ip), myrank, nbproc)
      endif
    enddo
  enddo

contains

  subroutine transform(volume)
    real(kind=CUSTOM_REAL), intent(inout) :: volume(*)
    if (FWD) then
      call forward_wavelet(volume, nx, ny, nz, compression_type)
    else
      call inverse_wavelet(volume, nx, ny, nz, compression_type)
    endif
  end subroutine transform

end subroutine apply_wavelet_transform

end module wavelet_utils
