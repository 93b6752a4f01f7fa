!===============================================================================================
! sparse_matrix -- drop-in replacement of src/inversion/sparse_matrix.f90 (type t_sparse_matrix).
!
! Same public type name, same type-bound procedure names and argument lists
! (reference: sparse_matrix.f90:71-96), so damping*.F90, cross_gradient.F90, clustering.F90,
! joint_inverse_problem.F90, model.F90 and the unit tests compile against it unchanged. The storage
! (all `private` in the reference, :31-61) lives behind an opaque libtfx handle: the builder calls
! fill host arrays inside the library, finalize() validates them like the reference (:157-208) and
! mirrors the matrix to the GPU, the products run there.
!===============================================================================================
module sparse_matrix

  use, intrinsic :: iso_c_binding
  use global_typedefs
  use tfx_c_api

  implicit none

  private

  type, public :: t_sparse_matrix
    private
    type(c_ptr) :: handle = c_null_ptr
    ! Public fields of the reference type (sparse_matrix.f90:63,66).
    real(kind=CUSTOM_REAL), allocatable, public :: lsqr_var(:)
    integer, public :: tag
  contains
    private
    procedure, public, pass :: initialize => sparse_matrix_initialize
    procedure, public, pass :: reset => sparse_matrix_reset
    procedure, public, pass :: finalize => sparse_matrix_finalize
    procedure, public, pass :: add => sparse_matrix_add
    procedure, public, pass :: add_row => sparse_matrix_add_row
    procedure, public, pass :: new_row => sparse_matrix_new_row
    procedure, public, pass :: add_empty_rows => sparse_matrix_add_empty_rows
    procedure, public, pass :: mult_vector => sparse_matrix_mult_vector
    procedure, public, pass :: add_mult_vector => sparse_matrix_add_mult_vector
    procedure, public, pass :: part_mult_vector => sparse_matrix_part_mult_vector
    procedure, public, pass :: trans_mult_vector => sparse_matrix_trans_mult_vector
    procedure, public, pass :: add_trans_mult_vector => sparse_matrix_add_trans_mult_vector
    procedure, public, pass :: normalize_columns => sparse_matrix_normalize_columns
    procedure, public, pass :: get_total_row_number => sparse_matrix_get_total_row_number
    procedure, public, pass :: get_current_row_number => sparse_matrix_get_current_row_number
    procedure, public, pass :: get_ncolumns => sparse_matrix_get_ncolumns
    procedure, public, pass :: get_number_elements => sparse_matrix_get_number_elements
    procedure, public, pass :: get_nnz => sparse_matrix_get_nnz
    ! Not in the reference: lets lsqr_solver and sensitivity_gravmag pass / attach the C handle.
    procedure, public, pass :: c_handle => sparse_matrix_c_handle
    procedure, public, pass :: attach => sparse_matrix_attach
    final :: sparse_matrix_release
  end type t_sparse_matrix

contains

subroutine sparse_matrix_initialize(this, nl, ncolumns, nnz, myrank, nl_empty)
  class(t_sparse_matrix), intent(inout) :: this
  integer, intent(in) :: nl, ncolumns, myrank
  integer, intent(in), optional :: nl_empty
  integer(kind=8), intent(in) :: nnz
  integer(c_int32_t) :: ne

  ne = 0
  if (present(nl_empty)) ne = nl_empty
  if (c_associated(this%handle)) call tfx_check(tfx_sparse_matrix_destroy(this%handle), myrank)
  call tfx_check(tfx_sparse_matrix_initialize(this%handle, nl, ncolumns, nnz, myrank, ne), myrank)
end subroutine sparse_matrix_initialize

! The reference's reset is `pure` (sparse_matrix.f90:135); the host-side builder reset has no side
! effect outside the handle, the device mirror is dropped at the next finalize().
subroutine sparse_matrix_reset(this)
  class(t_sparse_matrix), intent(inout) :: this
  call tfx_check(tfx_sparse_matrix_reset(this%handle), 0)
end subroutine sparse_matrix_reset

subroutine sparse_matrix_finalize(this, myrank)
  class(t_sparse_matrix), intent(inout) :: this
  integer, intent(in) :: myrank
  call tfx_check(tfx_sparse_matrix_finalize(this%handle, myrank), myrank)
end subroutine sparse_matrix_finalize

subroutine sparse_matrix_add(this, value, column, myrank)
  class(t_sparse_matrix), intent(inout) :: this
  real(kind=CUSTOM_REAL), intent(in) :: value
  integer, intent(in) :: column, myrank
  call tfx_check(tfx_sparse_matrix_add(this%handle, real(value, c_double), column, myrank), myrank)
end subroutine sparse_matrix_add

subroutine sparse_matrix_add_row(this, nel, values, columns, myrank)
  class(t_sparse_matrix), intent(inout) :: this
  integer, intent(in) :: nel, myrank
  real(kind=MATRIX_PRECISION), intent(in) :: values(nel)
  integer, intent(in) :: columns(nel)
  call tfx_check(tfx_sparse_matrix_add_row(this%handle, nel, values, columns, myrank), myrank)
end subroutine sparse_matrix_add_row

subroutine sparse_matrix_new_row(this, myrank)
  class(t_sparse_matrix), intent(inout) :: this
  integer, intent(in) :: myrank
  call tfx_check(tfx_sparse_matrix_new_row(this%handle, myrank), myrank)
end subroutine sparse_matrix_new_row

subroutine sparse_matrix_add_empty_rows(this, nrows, myrank)
  class(t_sparse_matrix), intent(inout) :: this
  integer, intent(in) :: nrows, myrank
  call tfx_check(tfx_sparse_matrix_add_empty_rows(this%handle, nrows, myrank), myrank)
end subroutine sparse_matrix_add_empty_rows

pure subroutine sparse_matrix_mult_vector(this, x, b)
  class(t_sparse_matrix), intent(in) :: this
  real(kind=CUSTOM_REAL), intent(in) :: x(:)
  real(kind=CUSTOM_REAL), intent(out) :: b(:)
  call tfx_sparse_matrix_mult_vector_v(this%handle, x, b)
end subroutine sparse_matrix_mult_vector

pure subroutine sparse_matrix_add_mult_vector(this, x, b)
  class(t_sparse_matrix), intent(in) :: this
  real(kind=CUSTOM_REAL), intent(in) :: x(:)
  real(kind=CUSTOM_REAL), intent(inout) :: b(:)
  call tfx_sparse_matrix_add_mult_vector_v(this%handle, x, b)
end subroutine sparse_matrix_add_mult_vector

subroutine sparse_matrix_part_mult_vector(this, nelements, x, ndata, b, line_start, param_shift, myrank)
  class(t_sparse_matrix), intent(in) :: this
  integer, intent(in) :: nelements, ndata, line_start, param_shift, myrank
  real(kind=CUSTOM_REAL), intent(in) :: x(nelements)
  real(kind=CUSTOM_REAL), intent(out) :: b(ndata)
  call tfx_check(tfx_sparse_matrix_part_mult_vector(this%handle, nelements, x, ndata, b, line_start, param_shift, myrank), &
                 myrank)
end subroutine sparse_matrix_part_mult_vector

pure subroutine sparse_matrix_trans_mult_vector(this, x, b)
  class(t_sparse_matrix), intent(in) :: this
  real(kind=CUSTOM_REAL), intent(in) :: x(:)
  real(kind=CUSTOM_REAL), intent(out) :: b(:)
  call tfx_sparse_matrix_trans_mult_vector_v(this%handle, x, b)
end subroutine sparse_matrix_trans_mult_vector

pure subroutine sparse_matrix_add_trans_mult_vector(this, x, b)
  class(t_sparse_matrix), intent(in) :: this
  real(kind=CUSTOM_REAL), intent(in) :: x(:)
  real(kind=CUSTOM_REAL), intent(inout) :: b(:)
  call tfx_sparse_matrix_add_trans_mult_vector_v(this%handle, x, b)
end subroutine sparse_matrix_add_trans_mult_vector

subroutine sparse_matrix_normalize_columns(this, column_norm)
  class(t_sparse_matrix), intent(inout) :: this
  real(kind=CUSTOM_REAL), intent(out) :: column_norm(:)
  call tfx_check(tfx_sparse_matrix_normalize_columns(this%handle, column_norm), 0)
end subroutine sparse_matrix_normalize_columns

pure function sparse_matrix_get_total_row_number(this) result(res)
  class(t_sparse_matrix), intent(in) :: this
  integer :: res
  res = tfx_sparse_matrix_get_total_row_number(this%handle)
end function sparse_matrix_get_total_row_number

pure function sparse_matrix_get_current_row_number(this) result(res)
  class(t_sparse_matrix), intent(in) :: this
  integer :: res
  res = tfx_sparse_matrix_get_current_row_number(this%handle)
end function sparse_matrix_get_current_row_number

pure function sparse_matrix_get_ncolumns(this) result(res)
  class(t_sparse_matrix), intent(in) :: this
  integer :: res
  res = tfx_sparse_matrix_get_ncolumns(this%handle)
end function sparse_matrix_get_ncolumns

pure function sparse_matrix_get_number_elements(this) result(res)
  class(t_sparse_matrix), intent(in) :: this
  integer(kind=8) :: res
  res = tfx_sparse_matrix_get_number_elements(this%handle)
end function sparse_matrix_get_number_elements

pure function sparse_matrix_get_nnz(this) result(res)
  class(t_sparse_matrix), intent(in) :: this
  integer(kind=8) :: res
  res = tfx_sparse_matrix_get_nnz(this%handle)
end function sparse_matrix_get_nnz

pure function sparse_matrix_c_handle(this) result(h)
  class(t_sparse_matrix), intent(in) :: this
  type(c_ptr) :: h
  h = this%handle
end function sparse_matrix_c_handle

! Adopts a matrix that libtfx assembled on the device (tfx_calculate_sensit).
subroutine sparse_matrix_attach(this, h)
  class(t_sparse_matrix), intent(inout) :: this
  type(c_ptr), intent(in) :: h
  integer(c_int) :: rc
  if (c_associated(this%handle)) rc = tfx_sparse_matrix_destroy(this%handle)
  this%handle = h
end subroutine sparse_matrix_attach

subroutine sparse_matrix_release(this)
  type(t_sparse_matrix), intent(inout) :: this
  integer(c_int) :: rc
  if (c_associated(this%handle)) rc = tfx_sparse_matrix_destroy(this%handle)
  this%handle = c_null_ptr
end subroutine sparse_matrix_release

end module sparse_matrix
