!===============================================================================================
! sensitivity_gravmag -- drop-in replacement of src/forward/gravmag/sensitivity_gravmag.F90.
!
! Same module name and public procedures (reference :44-49), same argument lists, so
! problem_joint_gravmag.F90:186-248 compiles and runs unchanged:
!
!   calculate_and_write_sensit (:82)   rows of this rank's stations are evaluated, wavelet-compressed and
!                                      thresholded on the GPU (csrc/sensit.cu), written to the reference's
!                                      stream file sensit_<grav|magn>_<nbproc>_<rank> (+ _meta.txt, _nnz on
!                                      rank 0) AND kept in HBM for read_sensitivity_kernel.
!   calculate_new_partitioning (:573)  reference algorithm (get_load_balancing_nelements, :470-524) on rank 0
!                                      from the _nnz file(s), then MPI_Bcast like the reference.
!   read_sensitivity_kernel (:648)     sensit.readFromFiles = 0: the HBM-resident rows are re-partitioned to the
!                                      column slabs with one all-to-all over NVLink (csrc/sensit_dist.cu);
!                                      otherwise every rank scans the stream files (csrc/sensit_io.cu). Either
!                                      way the rows are APPENDED to sensit_matrix (one call per problem) and
!                                      sensit_matrix%finalize builds the device representations.
!   read_sensitivity_metadata (:974), write_depth_weight (:415), read_depth_weight (:888): file formats kept.
!
! Not compiled in the development image (no Fortran compiler there); all logic is behind the C ABI
! (include/tfx.h) and tested through it (tests/test_gpu_sensit.py, tests/test_sensit_files.py,
! tests/test_gpu_joint.py, tests/multi_rank_case.py).
!===============================================================================================
module sensitivity_gravmag

  use, intrinsic :: iso_c_binding
  use mpi
  use global_typedefs
  use mpi_tools, only: exit_MPI
  use parameters_gravmag
  use parameters_mag
  use parameters_grav
  use grid
  use data_gravmag
  use sparse_matrix
  use parallel_tools
  use tfx_c_api

  implicit none

  private

  public :: calculate_and_write_sensit
  public :: read_sensitivity_kernel
  public :: read_sensitivity_metadata
  public :: calculate_new_partitioning
  public :: write_depth_weight
  public :: read_depth_weight

  ! Row shards kept in HBM between calculate_and_write_sensit and read_sensitivity_kernel (per problem).
  type(c_ptr), save :: rows_in_hbm(2) = [c_null_ptr, c_null_ptr]

contains

!-----------------------------------------------------------------------------------------------
! The C copy of the parameters (struct tfx_sensit_params) and the problem type (select type, :126-137).
!-----------------------------------------------------------------------------------------------
subroutine fill_params(par, pc, problem_type)
  class(t_parameters_base), intent(in) :: par
  type(tfx_sensit_params), intent(out) :: pc
  integer, intent(out) :: problem_type

  pc%mi = 0.d0; pc%md = 0.d0; pc%theta = 0.d0; pc%intensity = 0.d0
  select type(par)
  class is (t_parameters_grav)
    problem_type = 1
  class is (t_parameters_mag)
    problem_type = 2
    pc%mi = par%mi; pc%md = par%md; pc%theta = par%theta; pc%intensity = par%intensity
  end select
  pc%problem_type = problem_type
  pc%nx = par%nx; pc%ny = par%ny; pc%nz = par%nz
  pc%ndata = par%ndata
  pc%ndata_components = par%ndata_components
  pc%nmodel_components = par%nmodel_components
  pc%data_type = par%data_type
  pc%compression_type = par%compression_type
  pc%compression_rate = par%compression_rate
  pc%problem_weight = 1.d0
  pc%cell0 = 0; pc%ncells_local = par%nx * par%ny * par%nz
  pc%param_shift = 0; pc%ncolumns = 0
end subroutine fill_params

function sensit_dir(par) result(dir)
  class(t_parameters_base), intent(in) :: par
  character(len=512) :: dir
  if (par%sensit_read /= 0) then
    dir = trim(par%sensit_path)//c_null_char
  else
    dir = trim(path_output)//"/SENSIT/"//c_null_char
  endif
end function sensit_dir

!===============================================================================================
subroutine calculate_and_write_sensit(par, grid_full, data, column_weight, memory, myrank, nbproc)
  class(t_parameters_base), intent(in) :: par
  type(t_grid), intent(in) :: grid_full
  type(t_data), intent(in) :: data
  real(kind=CUSTOM_REAL), intent(in) :: column_weight(par%nelements)
  integer, intent(in) :: myrank, nbproc
  real(kind=CUSTOM_REAL), intent(out) :: memory

  type(tfx_sensit_params) :: pc
  integer :: problem_type, nelements_total, ierr
  real(kind=CUSTOM_REAL), allocatable :: column_weight_full(:), unit_weight(:)
  integer(c_int32_t), allocatable :: sensit_nnz(:)
  real(c_double) :: comp_error
  integer(c_int64_t) :: nnz_total, free_b, total_b
  character(len=512) :: dir

  call fill_params(par, pc, problem_type)
  nelements_total = par%nx * par%ny * par%nz

  allocate(column_weight_full(nelements_total), source=0._CUSTOM_REAL, stat=ierr)
  allocate(unit_weight(par%ndata_components * par%ndata), source=1._CUSTOM_REAL, stat=ierr)
  allocate(sensit_nnz(nelements_total), source=0, stat=ierr)
  if (ierr /= 0) call exit_MPI("Dynamic memory allocation error in calculate_and_write_sensit!", myrank, ierr)

  ! Every rank needs the full column weight (reference :176).
  call get_full_array(column_weight, par%nelements, column_weight_full, .true., myrank, nbproc)

  if (c_associated(rows_in_hbm(problem_type))) call tfx_check(tfx_sensit_rows_destroy(rows_in_hbm(problem_type)), myrank)

  ! Rows of this rank's stations (even split of the data, reference :179-180), unit weights: the stream file
  ! holds the unweighted kernel, read_sensitivity_kernel applies problem and data weights (:837-843).
  call tfx_check(tfx_sensit_assemble_rows(rows_in_hbm(problem_type), pc, &
                                          grid_full%X1, grid_full%X2, grid_full%Y1, grid_full%Y2, grid_full%Z1, grid_full%Z2, &
                                          data%X, data%Y, data%Z, column_weight_full, unit_weight, myrank, nbproc, &
                                          sensit_nnz, comp_error, nnz_total), myrank)

  if (myrank == 0) print *, 'nnz_total = ', nnz_total
  if (myrank == 0) print *, 'COMPRESSION RATE = ', dble(nnz_total) / dble(nelements_total) / dble(par%ndata) &
                                                   / dble(par%nmodel_components) / dble(par%ndata_components)
  if (myrank == 0) print *, 'COMPRESSION ERROR, r = ', comp_error

  dir = trim(path_output)//"/SENSIT/"//c_null_char
  call tfx_check(tfx_create_sensit_directory(dir), myrank)
  call tfx_check(tfx_write_sensit_file(rows_in_hbm(problem_type), dir), myrank)
  if (myrank == 0) then
    call tfx_check(tfx_write_sensit_metadata(pc, dir, nbproc, par%depth_weighting_type, comp_error, nnz_total, &
                                             sensit_nnz), myrank)
  endif
  call MPI_Barrier(MPI_COMM_WORLD, ierr)

  call tfx_check(tfx_device_mem_info(free_b, total_b), myrank)
  memory = dble(total_b - free_b) / 1024.d0**3

  deallocate(column_weight_full, unit_weight, sensit_nnz)
  if (myrank == 0) print *, 'Finished calculating the sensitivity kernel.'
end subroutine calculate_and_write_sensit

!===============================================================================================
subroutine write_depth_weight(par, column_weight, myrank, nbproc)
  class(t_parameters_base), intent(in) :: par
  real(kind=CUSTOM_REAL), intent(in) :: column_weight(par%nelements)
  integer, intent(in) :: myrank, nbproc

  type(tfx_sensit_params) :: pc
  integer :: problem_type, nelements_total, ierr
  real(kind=CUSTOM_REAL), allocatable :: column_weight_full(:)
  character(len=512) :: dir

  call fill_params(par, pc, problem_type)
  nelements_total = par%nx * par%ny * par%nz
  if (myrank == 0) then
    allocate(column_weight_full(nelements_total), source=0._CUSTOM_REAL, stat=ierr)
  else
    allocate(column_weight_full(1), source=0._CUSTOM_REAL, stat=ierr)
  endif
  call get_full_array(column_weight, par%nelements, column_weight_full, .false., myrank, nbproc)
  if (myrank == 0) then
    dir = trim(path_output)//"/SENSIT/"//c_null_char
    call tfx_check(tfx_write_depth_weight(pc, dir, column_weight_full), myrank)
  endif
  deallocate(column_weight_full)
end subroutine write_depth_weight

!===============================================================================================
subroutine read_depth_weight(par, column_weight, myrank, nbproc)
  class(t_parameters_base), intent(in) :: par
  integer, intent(in) :: myrank, nbproc
  real(kind=CUSTOM_REAL), intent(out) :: column_weight(par%nelements)

  type(tfx_sensit_params) :: pc
  integer :: problem_type, nelements_total, ierr
  real(kind=CUSTOM_REAL), allocatable :: column_weight_full(:)

  call fill_params(par, pc, problem_type)
  nelements_total = par%nx * par%ny * par%nz
  if (myrank == 0) then
    allocate(column_weight_full(nelements_total), source=0._CUSTOM_REAL, stat=ierr)
    call tfx_check(tfx_read_depth_weight(pc, sensit_dir(par), column_weight_full), myrank)
  else
    allocate(column_weight_full(1), source=0._CUSTOM_REAL, stat=ierr)
  endif
  call scatter_full_array(par%nelements, column_weight_full, column_weight, myrank, nbproc)
  deallocate(column_weight_full)
end subroutine read_depth_weight

!===============================================================================================
subroutine calculate_new_partitioning(par, nnz, nelements_at_cpu, problem_type, myrank, nbproc)
  class(t_parameters_base), intent(in) :: par
  integer, intent(in) :: problem_type
  integer, intent(in) :: myrank, nbproc
  integer(kind=8), intent(out) :: nnz
  integer, intent(out) :: nelements_at_cpu(nbproc)

  type(tfx_sensit_params) :: pc
  integer(kind=8) :: nnz_at_cpu(nbproc)
  integer :: nelements_total, ierr, ptype_par
  integer(c_int32_t), allocatable :: sensit_nnz(:), sensit_nnz2(:)

  if (myrank == 0) then
    call fill_params(par, pc, ptype_par)
    nelements_total = par%nx * par%ny * par%nz
    allocate(sensit_nnz(nelements_total), source=0, stat=ierr)
    if (problem_type == 3) then
      ! Joint inversion: the total of both problems (reference :610-625).
      allocate(sensit_nnz2(nelements_total), source=0, stat=ierr)
      pc%problem_type = 1
      call tfx_check(tfx_read_sensit_nnz(pc, sensit_dir(par), sensit_nnz), myrank)
      pc%problem_type = 2
      call tfx_check(tfx_read_sensit_nnz(pc, sensit_dir(par), sensit_nnz2), myrank)
      sensit_nnz = sensit_nnz + sensit_nnz2
      deallocate(sensit_nnz2)
    else
      pc%problem_type = problem_type
      call tfx_check(tfx_read_sensit_nnz(pc, sensit_dir(par), sensit_nnz), myrank)
    endif
    call tfx_check(tfx_get_load_balancing_nelements(nelements_total, sensit_nnz, nbproc, nnz_at_cpu, nelements_at_cpu), &
                   myrank)
    deallocate(sensit_nnz)
  endif

  call MPI_Bcast(nnz_at_cpu, nbproc, MPI_INTEGER8, 0, MPI_COMM_WORLD, ierr)
  call MPI_Bcast(nelements_at_cpu, nbproc, MPI_INTEGER, 0, MPI_COMM_WORLD, ierr)

  nnz = nnz_at_cpu(myrank + 1)

  if (myrank == 0) then
    print *, "nelements_at_cpu =", nelements_at_cpu
    print *, "nnz_at_cpu =", nnz_at_cpu
  endif
end subroutine calculate_new_partitioning

!===============================================================================================
subroutine read_sensitivity_kernel(par, sensit_matrix, column_weight, problem_weight, data_weight, problem_type, &
                                   myrank, nbproc, nelements_at_cpu)
  class(t_parameters_base), intent(in) :: par
  real(kind=CUSTOM_REAL), intent(in) :: problem_weight
  real(kind=CUSTOM_REAL), intent(in) :: data_weight(par%ndata_components, par%ndata)
  integer, intent(in) :: problem_type
  integer, intent(in) :: myrank, nbproc
  integer, intent(in) :: nelements_at_cpu(nbproc)
  type(t_sparse_matrix), intent(inout) :: sensit_matrix
  real(kind=CUSTOM_REAL), intent(out) :: column_weight(par%nelements)

  type(tfx_sensit_params) :: pc
  integer :: ptype_par, ierr
  integer(c_int64_t) :: nnz_loc, nnz_total

  call fill_params(par, pc, ptype_par)
  pc%problem_weight = problem_weight

  if (par%sensit_read == 0 .and. c_associated(rows_in_hbm(problem_type))) then
    ! The kernel of this run is still in HBM: weights in real(4) (:837-843), then one all-to-all to the slabs.
    call tfx_check(tfx_sensit_rows_apply_weights(rows_in_hbm(problem_type), real(problem_weight, c_double), data_weight), &
                   myrank)
    call tfx_check(tfx_sensit_repartition_into(sensit_matrix%c_handle(), rows_in_hbm(problem_type), problem_type, &
                                               nelements_at_cpu, myrank, nbproc), myrank)
    call tfx_check(tfx_sensit_rows_destroy(rows_in_hbm(problem_type)), myrank)
    rows_in_hbm(problem_type) = c_null_ptr
    nnz_loc = sensit_matrix%get_number_elements()
  else
    call tfx_check(tfx_read_sensitivity_kernel_into(sensit_matrix%c_handle(), pc, sensit_dir(par), data_weight, &
                                                    par%depth_weighting_type, problem_type, myrank, nbproc, &
                                                    nelements_at_cpu, nnz_loc), myrank)
  endif

  call MPI_Allreduce(nnz_loc, nnz_total, 1, MPI_INTEGER8, MPI_SUM, MPI_COMM_WORLD, ierr)
  if (myrank == 0) print *, 'nnz_total (of the read kernel)  = ', nnz_total

  call read_depth_weight(par, column_weight, myrank, nbproc)

  if (myrank == 0) print *, 'Finished reading the sensitivity kernel.'
end subroutine read_sensitivity_kernel

!===============================================================================================
subroutine read_sensitivity_metadata(par, nbproc_sensit, problem_type, myrank)
  class(t_parameters_base), intent(in) :: par
  integer, intent(in) :: problem_type
  integer, intent(in) :: myrank
  integer, intent(out) :: nbproc_sensit

  type(tfx_sensit_params) :: pc
  integer :: ptype_par, ierr
  real(c_double) :: comp_error
  integer(c_int64_t) :: nnz_total

  if (myrank == 0) then
    call fill_params(par, pc, ptype_par)
    pc%problem_type = problem_type
    call tfx_check(tfx_read_sensitivity_metadata(pc, sensit_dir(par), par%depth_weighting_type, nbproc_sensit, &
                                                 comp_error, nnz_total), myrank)
    print *, "COMPRESSION ERROR (read) =", comp_error
  endif
  call MPI_Bcast(nbproc_sensit, 1, MPI_INTEGER, 0, MPI_COMM_WORLD, ierr)
end subroutine read_sensitivity_metadata

end module sensitivity_gravmag
