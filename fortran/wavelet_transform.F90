!===============================================================================================
! wavelet_transform -- drop-in replacement of src/utils/wavelet_transform.F90.
! Same public names (reference :23-30); the in-place 3-D lifting transforms run on the GPU.
! Callers pass 1-D arrays that are reinterpreted as s(n1,n2,n3) (sequence association), like the
! reference's explicit-shape dummies.
!===============================================================================================
module wavelet_transform

  use, intrinsic :: iso_c_binding
  use global_typedefs
  use tfx_c_api

  implicit none

  private

  public :: forward_wavelet, inverse_wavelet
  public :: Haar3D, iHaar3D, DaubD43D, iDaubD43D

contains

subroutine forward_wavelet(s, n1, n2, n3, wavelet_type)
  integer, intent(in) :: n1, n2, n3, wavelet_type
  real(kind=CUSTOM_REAL), intent(inout) :: s(n1, n2, n3)
  call tfx_check(tfx_forward_wavelet(s, n1, n2, n3, wavelet_type), 0)
end subroutine forward_wavelet

subroutine inverse_wavelet(s, n1, n2, n3, wavelet_type)
  integer, intent(in) :: n1, n2, n3, wavelet_type
  real(kind=CUSTOM_REAL), intent(inout) :: s(n1, n2, n3)
  call tfx_check(tfx_inverse_wavelet(s, n1, n2, n3, wavelet_type), 0)
end subroutine inverse_wavelet

subroutine Haar3D(s, n1, n2, n3)
  integer, intent(in) :: n1, n2, n3
  real(kind=CUSTOM_REAL), intent(inout) :: s(n1, n2, n3)
  call tfx_check(tfx_forward_wavelet(s, n1, n2, n3, 1), 0)
end subroutine Haar3D

subroutine iHaar3D(s, n1, n2, n3)
  integer, intent(in) :: n1, n2, n3
  real(kind=CUSTOM_REAL), intent(inout) :: s(n1, n2, n3)
  call tfx_check(tfx_inverse_wavelet(s, n1, n2, n3, 1), 0)
end subroutine iHaar3D

subroutine DaubD43D(s, n1, n2, n3)
  integer, intent(in) :: n1, n2, n3
  real(kind=CUSTOM_REAL), intent(inout) :: s(n1, n2, n3)
  call tfx_check(tfx_forward_wavelet(s, n1, n2, n3, 2), 0)
end subroutine DaubD43D

subroutine iDaubD43D(s, n1, n2, n3)
  integer, intent(in) :: n1, n2, n3
  real(kind=CUSTOM_REAL), intent(inout) :: s(n1, n2, n3)
  call tfx_check(tfx_inverse_wavelet(s, n1, n2, n3, 2), 0)
end subroutine iDaubD43D

end module wavelet_transform
