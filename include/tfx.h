/*
 * tfx.h -- C ABI of libtfx: the B200-native replacement for the Tomofast-x inversion hot path.
 *
 * The reference (TOMOFAST/Tomofast-x, Fortran 2008 + MPI) has no plugin/FFI interface; the drop-in
 * boundary is the set of Fortran MODULE interfaces used by problem_joint_gravmag.F90 and
 * joint_inverse_problem.F90.  Every entry point below replaces one module procedure and cites it
 * (paths relative to the reference checkout).  fortran/ holds same-named ISO_C_BINDING shim modules
 * that forward to these symbols; INTEGRATION.md shows how a maintainer wires them in.
 *
 * Conventions
 *  - plain C types only; all indices crossing the ABI are 1-BASED like the reference's.
 *  - vectors are real(8) (CUSTOM_REAL = 8), matrix values real(4) (MATRIX_PRECISION = 4),
 *    src/global_typedefs.F90:39-45.
 *  - vector arguments may be HOST pointers (copied to/from the device inside the call) or DEVICE
 *    pointers (used in place); the library detects which with cudaPointerGetAttributes.
 *  - every function returns 0 on success; otherwise a negative code and tfx_last_error() holds the
 *    message the reference would have passed to exit_MPI (src/utils/mpi_tools.F90:30-54).
 *  - one process drives one GPU (one MPI rank per GPU); calls are synchronous for the caller.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef TFX_H
#define TFX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfx_matrix tfx_matrix;   /* replaces type(t_sparse_matrix), sparse_matrix.f90:31-98 */

/* ---- runtime --------------------------------------------------------------------------------- */
int         tfx_version(void);
const char *tfx_last_error(void);
/* Selects the GPU of this rank (device < 0: LOCAL_RANK env or 0) and creates the stream.
 * Replaces nothing in the reference (MPI_Init stays in program_tomofastx.F90:56). */
int         tfx_init(int device);
/* Stream ordering: the library works on its own NON-BLOCKING CUDA stream and every entry point returns after that
 * stream has drained, so a host program (the Fortran reference: single-threaded, blocking) needs nothing. A caller that
 * hands over DEVICE pointers written by its own streams (another library, torch) must finish that work first
 * (cudaStreamSynchronize / cudaDeviceSynchronize): the library does not wait for foreign streams. */
int         tfx_finalize(void);
int         tfx_device_synchronize(void);
/* Number of GPU kernels launched by the library so far (bench.py's gpu_launches). */
uint64_t    tfx_launch_count(void);
/* CUDA-event stopwatch on the library stream: start records an event, stop records a second one, waits
 * for it and returns the elapsed device time in milliseconds. */
int         tfx_timer_start(void);
int         tfx_timer_stop(double *ms);
/* Options: "dense_detect" (1: finalize() turns an uncompressed CSR into the dense block);
 * "strict_order" (1: LSQR sums in the reference's sequential order -- slow parity mode);
 * "profile_sweeps" (1: CUDA events around every fused sweep launch);
 * "t16_min_nnz" (matrices with at least this many entries get the T16 layouts; default 4194304);
 * "t16_tile" (0: automatic tile size, else a power of two <= 16384 -- tests);
 * "lsqr_poll" (iterations between two reads of the device-side done flag, default 8);
 * "wavelet_cols" (1, default: column-layout wavelet kernel where the axis fits its tile; 0: generic kernel);
 * "wavelet_slab_mb" (size of the i3 slabs of the L2-blocked axis-1 / axis-2 passes, 0: whole volume per pass);
 * "lsqr_graph" (1, default: the iteration body of the split LSQR path is captured once and replayed as a CUDA graph
 *   when the matrix has fewer than 2e8 entries -- the launch-bound regime);
 * "t16_bank_deal" (1, default: the T16 builder deals the entries of long segments over the shared-memory banks);
 * "t16_async" (bit 0 / bit 1: long segments of the TILES / DIRECT kernel through the cp.async ring; default 3);
 * "t16_direct_max" (gathered ranges up to this many elements use one DIRECT tile; default 16384);
 * "sensit_row_blocks" (1: tfx_sensit_repartition_into / tfx_read_sensitivity_kernel_into build one independent row
 *   block per call -- bounded build memory for kernels near the HBM capacity; such matrices cannot be exported);
 * "grav_shared_nodes" / "mag_shared_nodes" (1, default: forward-kernel lines on structured grids evaluate the prism
 *   corner terms once per grid node -- and the magnetic edge terms once per grid edge -- instead of once per cell);
 * "dense_vec4" (1, default: 512 threads x float4 rows; 0: 1024 threads x float2 rows), "dense_f2f_rows" (row vectors
 *   per thread whose second use converts with F2F; default 2), "dense_stream_only" (diagnostic: the sweep's TMA ring
 *   without the products -- results are meaningless, only the time is);
 * "dense_block_rows" (0, default: 10240; data rows per dense block of an uncompressed kernel -- more rows become several
 *   dense row blocks on the split LSQR path; tests lower it);
 * "wavelet_fuse12" (1, default: Haar axis-1 pass fused with the three lowest axis-2 scales), "wavelet_dist" (1, default:
 *   distributed transform on plane-owner / column-owner layouts; 0: all-gather), "wavelet_p2p" (1, default: its layout
 *   changes through cudaIpc peer memory; 0: ncclSend/ncclRecv), "wavelet_tile_kb" (0: automatic tile size);
 * "sensit_cand_cap" (0, default: 16384; candidate-list capacity of the k-th select of the row pipeline -- tests force
 *   the fallback passes with 1);
 * "t16_tma" (1: long segments through the cp.async.bulk + mbarrier ring; measured slower, default 0);
 * "trace" (1: wall-clock phases of the assembly on stderr). */
int         tfx_set_option(const char *name, int value);

/* Memory helpers for callers that keep their vectors on the device (or in pinned host memory)
 * between calls; thin wrappers of cudaMalloc / cudaMallocHost / cudaMemcpy(Default) / cudaMemGetInfo. */
int tfx_device_alloc(void **p, int64_t bytes);
int tfx_device_free(void *p);
int tfx_host_alloc(void **p, int64_t bytes);
int tfx_host_free(void *p);
int tfx_memcpy(void *dst, const void *src, int64_t bytes);
/* cudaMemset on the library stream (waits for it): zero-fills device-resident vectors such as b_RHS. */
int tfx_device_memset(void *dst, int value, int64_t bytes);
int tfx_device_mem_info(int64_t *free_bytes, int64_t *total_bytes);

/* ---- communicator: stands in for MPI_COMM_WORLD on the solver's reductions ---------------------
 * lsqr_solver2.F90:214 (MPI_Allreduce of u) and :514 (scalar Allreduce) become ncclAllReduce.
 * Rank 0 creates the id, the host application broadcasts the 128 bytes (MPI_Bcast in the Fortran
 * shim, torch.distributed in the Python host), every rank calls tfx_comm_init. */
int tfx_comm_unique_id(char id[128]);
int tfx_comm_init(int nranks, int rank, const char id[128]);
int tfx_comm_finalize(void);
int tfx_comm_allreduce_sum(double *buf, int64_t count);            /* host or device pointer */

/* ---- module sparse_matrix (src/inversion/sparse_matrix.f90) ----------------------------------- */
int tfx_sparse_matrix_initialize(tfx_matrix **m, int32_t nl, int32_t ncolumns, int64_t nnz,
                                 int32_t myrank, int32_t nl_empty);             /* :105-130 */
int tfx_sparse_matrix_destroy(tfx_matrix *m);
int tfx_sparse_matrix_reset(tfx_matrix *m);                                    /* :135-151 */
int tfx_sparse_matrix_finalize(tfx_matrix *m, int32_t myrank);                 /* :157-208 (+ device upload) */
int tfx_sparse_matrix_add(tfx_matrix *m, double value, int32_t column, int32_t myrank);      /* :213-229 */
int tfx_sparse_matrix_add_row(tfx_matrix *m, int32_t nel_add, const float *values,
                              const int32_t *columns, int32_t myrank);         /* :234-249 */
int tfx_sparse_matrix_new_row(tfx_matrix *m, int32_t myrank);                  /* :254-276 */
int tfx_sparse_matrix_add_empty_rows(tfx_matrix *m, int32_t nrows, int32_t myrank);          /* :281-293 */
int tfx_sparse_matrix_mult_vector(tfx_matrix *m, const double *x, double *b);  /* :298-307 */
int tfx_sparse_matrix_add_mult_vector(tfx_matrix *m, const double *x, double *b);            /* :313-329 */
int tfx_sparse_matrix_part_mult_vector(tfx_matrix *m, int32_t nelements, const double *x, int32_t ndata,
                                       double *b, int32_t line_start, int32_t param_shift,
                                       int32_t myrank);                        /* :335-367 */
int tfx_sparse_matrix_trans_mult_vector(tfx_matrix *m, const double *x, double *b);          /* :373-382 */
int tfx_sparse_matrix_add_trans_mult_vector(tfx_matrix *m, const double *x, double *b);      /* :388-405 */
/* The reference declares the four products `pure` (sparse_matrix.f90:298,313,373,388). A Fortran 2008 pure procedure may
 * only call pure procedures, and a pure FUNCTION may not have intent(out/inout) dummies (C1276), so the shim binds these
 * void variants as `pure subroutine`s: same products, the return code is LATCHED inside the library (first failure wins,
 * with its message) and handed to the next non-pure call through tfx_take_latched_error() -- fortran/tfx_c_api.f90
 * tfx_check() asks for it on every call, so a failed product aborts the run at the next library call. */
void tfx_sparse_matrix_mult_vector_v(tfx_matrix *m, const double *x, double *b);
void tfx_sparse_matrix_add_mult_vector_v(tfx_matrix *m, const double *x, double *b);
void tfx_sparse_matrix_trans_mult_vector_v(tfx_matrix *m, const double *x, double *b);
void tfx_sparse_matrix_add_trans_mult_vector_v(tfx_matrix *m, const double *x, double *b);
/* Returns the latched code (0 = none) and clears it; tfx_last_error() then holds the latched message. */
int tfx_take_latched_error(void);
int tfx_sparse_matrix_normalize_columns(tfx_matrix *m, double *column_norm);    /* :414-443 (host builder) */
int32_t tfx_sparse_matrix_get_total_row_number(const tfx_matrix *m);           /* :448-453 */
int32_t tfx_sparse_matrix_get_current_row_number(const tfx_matrix *m);         /* :458-463 */
int32_t tfx_sparse_matrix_get_ncolumns(const tfx_matrix *m);                   /* :468-473 */
int64_t tfx_sparse_matrix_get_number_elements(const tfx_matrix *m);            /* :478-483 */
int64_t tfx_sparse_matrix_get_nnz(const tfx_matrix *m);                        /* :488-493 */
/* Bulk constructor from arrays in the reference's storage (1-based ija/ijl/rowptr); finalized. */
int tfx_sparse_matrix_from_arrays(tfx_matrix **m, int32_t nl, int32_t ncolumns, int32_t nl_nonempty,
                                  int64_t nel, const float *sa, const int32_t *ija, const int64_t *ijl,
                                  const int32_t *rowptr);
/* 0: compressed-segment (CSR + CSR of the transpose); 1: dense column-major block;
 * 2: T16 tiled layouts with 16-bit in-tile indices (big compressed matrices, option "t16_min_nnz"). */
int tfx_sparse_matrix_storage_kind(const tfx_matrix *m);
/* Mean device time (ms, CUDA events on the library stream) of `reps` back-to-back products b = A x
 * (transposed = 0) or b = A^T x (1) on DEVICE-resident vectors: the SpMV GB/s figures of bench.py. */
int tfx_sparse_matrix_time_product(tfx_matrix *m, int transposed, const double *x, double *b, int reps, double *ms);
/* Device memory held by the matrix (all representations), bytes. */
int64_t tfx_sparse_matrix_device_bytes(const tfx_matrix *m);
/* Frees the generic CSR copies of a matrix that has the T16 layouts (export / strict_order unavailable after). */
int tfx_sparse_matrix_drop_csr(tfx_matrix *m);
/* Copies the device-resident matrix back in the reference's CSR storage (for tests / file writers).
 * Pass NULL arrays to query sizes only. */
int tfx_sparse_matrix_export(tfx_matrix *m, int64_t *nel, int32_t *nl_nonempty, float *sa, int32_t *ija,
                             int64_t *ijl, int32_t *rowptr);

/* ---- module wavelet_transform (src/utils/wavelet_transform.F90) -------------------------------- */
int tfx_forward_wavelet(double *s, int32_t n1, int32_t n2, int32_t n3, int32_t wavelet_type);   /* :37-51 */
int tfx_inverse_wavelet(double *s, int32_t n1, int32_t n2, int32_t n3, int32_t wavelet_type);   /* :56-70 */
int tfx_Haar3D(double *s, int32_t n1, int32_t n2, int32_t n3);                                  /* :75-153 */
int tfx_iHaar3D(double *s, int32_t n1, int32_t n2, int32_t n3);                                 /* :158-236 */
int tfx_DaubD43D(double *s, int32_t n1, int32_t n2, int32_t n3);                                /* :243-367 */
int tfx_iDaubD43D(double *s, int32_t n1, int32_t n2, int32_t n3);                               /* :374-498 */
/* Diagnostic: non-zero when the last transform of a distributed vector ran on the plane-owner / column-owner layouts (two
 * all-to-all exchanges: 1 through ncclSend/ncclRecv, 2 through peer memory), 0 when it gathered the slabs into a full
 * volume on every GPU (slab layout did not qualify). */
int tfx_wavelet_last_distributed(void);
/* module wavelet_utils: apply_wavelet_transform (src/inversion/wavelet_utils.F90:37-72):
 * v(nelements, ncomponents, nproblems) holds this rank's cell slab of every volume. With nbproc > 1 (needs
 * tfx_comm_init) the volume is transformed in place across the GPUs -- axis-1/2 passes on the k-planes a rank owns, one
 * NVLink all-to-all, the axis-3 pass on its share of the k-lines, one all-to-all back (or, when a slab is thinner than a
 * plane: slabs all-gathered on every GPU, transformed, own slab kept) -- the reference's gather to rank 0 / serial
 * transform / scatter (:57-67) without the serial section, bit-identical; model_full is not needed. The two all-to-all
 * layout changes go through peer memory when the ranks' buffers can be mapped into each other (cudaIpc over NVLink: one
 * kernel per exchange stores straight into the peers' buffers), else through grouped ncclSend/ncclRecv (option
 * "wavelet_p2p" = 0 forces that path). */
int tfx_apply_wavelet_transform(int32_t nelements, int32_t nx, int32_t ny, int32_t nz, int32_t ncomponents,
                                double *v, int32_t fwd, int32_t compression_type, int32_t nproblems,
                                const int32_t *solve_problem, int32_t myrank, int32_t nbproc);

/* ---- t_model%calculate_data (src/inversion/model.F90:220-307) -------------------------------------
 * data_calc = (S W(model / column_weight)) / problem_weight / data_weight for one problem of the (joint)
 * matrix: rows [line_start, line_start + ndata*ndata_components), columns shifted by param_shift (part_mult_vector).
 * model_val(nelements, ncomponents), column_weight(nelements), data_weight and data_calc(ndata_components, ndata)
 * may be host or device pointers; the MPI_Allreduce of :293 is an NCCL all-reduce when nbproc > 1. */
int tfx_calculate_data(tfx_matrix *matrix_sensit, int32_t nelements, int32_t ncomponents, const double *model_val,
                       int32_t ndata, int32_t ndata_components, double problem_weight, const double *column_weight,
                       const double *data_weight, double *data_calc, int32_t compression_type,
                       int32_t nx, int32_t ny, int32_t nz, int32_t line_start, int32_t param_shift,
                       int32_t myrank, int32_t nbproc);

/* rescale_model (src/inversion/model.F90:312-324; called on delta_model at joint_inverse_problem.F90:569-571) and
 * t_model%update (model.F90:194-200): with tfx_apply_wavelet_transform, tfx_calculate_data, the constraint producers
 * and tfx_lsqr_solve_sensit they keep the whole major iteration of problem_joint_gravmag.F90:473-547 on device
 * buffers. Host or device pointers. */
int tfx_rescale_model(int32_t nelements, int32_t ncomponents, double *model, const double *weight);
int tfx_model_update(int32_t nelements, int32_t ncomponents, double *val, const double *delta_model);

/* ---- module weights_gravmag: calculate_depth_weight (src/forward/gravmag/weights_gravmag.f90:46-199) ----
 * column_weight(nelements) for the rank's cells nsmaller+1 .. nsmaller+nelements of the full grid (grid arrays
 * hold all nelements_total cells): type 1 depth weighting (:71-79, calc_depth_weight_pixel :204-223), type 2
 * distance weighting (:81-138, O(ncells*ndata) pow evaluations), type 3 minimum distance (:140-161); then the
 * cell-volume scaling (:170-175), normalisation by the global maximum (:228-250; an NCCL max all-reduce when
 * nbproc > 1) and inversion (:189-195). Grid, data and output pointers may be host or device memory. Aborts with
 * the reference's messages ("non-positive depth", "Zero depth weight norm", "Zero damping weight", "Not known
 * depth weight type"). */
int tfx_calculate_depth_weight(int32_t depth_weighting_type, double depth_weighting_power,
                               double depth_weighting_beta, double Z0, int32_t nelements_total,
                               const double *X1, const double *X2, const double *Y1, const double *Y2,
                               const double *Z1, const double *Z2, int32_t ndata, const double *data_X,
                               const double *data_Y, const double *data_Z, int32_t nsmaller, int32_t nelements,
                               double *column_weight, int32_t myrank, int32_t nbproc);

/* ---- constraint-matrix producers on the device (csrc/cons.cu) --------------------------------------
 * Each call appends its rows to `matrix` (initialize ... [producers] ... finalize; host-built rows may precede or
 * follow) in the reference's add() order and writes the matching entries of b_RHS(nrows), the constraint part of
 * the right-hand side (b_RHS(lc:), joint_inverse_problem.F90:465), indexed by the matrix row number like the
 * reference does (get_current_row_number). Array arguments may be host or device pointers. With nbproc > 1 the
 * rank's slab position is taken from the communicator (get_nsmaller, parallel_tools.f90:68-86). */
/* t_damping%add (src/inversion/damping.F90:97-201, add_RHS :206-232, get_norm_multiplier :249-261): nx*ny*nz rows,
 * the rank's nelements diagonal entries alpha*problem_weight[*Lp multiplier][*local_weight] at columns
 * param_shift + i; model, model_ref, column_weight and local_weight (may be NULL) are the rank's slabs. Also the
 * ADMM term (joint_inverse_problem.F90:497-527: alpha = rho_ADMM, model_ref = x0_ADMM). cost = sum(b_RHS block^2). */
int tfx_damping_add(tfx_matrix *matrix, int32_t nrows, double *b_RHS, double alpha, double problem_weight,
                    double norm_power, int32_t compression_type, int32_t nx, int32_t ny, int32_t nz,
                    int32_t nelements, const double *column_weight, const double *model, const double *model_ref,
                    int32_t param_shift, int32_t wavelet_domain, const double *local_weight,
                    int32_t myrank, int32_t nbproc, double *cost);
/* t_damping_gradient%add (src/inversion/damping_gradient.F90:93-203; forward differences of gradient.F90:88-93,
 * boundary values of :175-225): nx*ny*nz rows for one direction (1 x, 2 y, 3 z) of one model component.
 * dX(nx), dY(ny), dZ(nz): t_grad_grid cell sizes (grid.F90:359-403); val_full and local_weight: full grid;
 * column_weight: the rank's slab. cost = sum of the squared gradient component. */
int tfx_damping_gradient_add(tfx_matrix *matrix, int32_t nrows, double *b_RHS, double beta, double problem_weight,
                             int32_t nx, int32_t ny, int32_t nz, const double *dX, const double *dY, const double *dZ,
                             int32_t nelements, const double *val_full, const double *column_weight,
                             const double *local_weight, int32_t param_shift, int32_t direction,
                             int32_t myrank, int32_t nbproc, double *cost);
/* t_cross_gradient%calculate with add = .true. and vec_field_type = 0 (src/inversion/cross_gradient.F90:220-391;
 * calculate_tau :455-567, calculate_tau_backward :676-740): 3*nx*ny*nz rows (x, y, z component per cell), columns
 * ind and ind + nparams_loc. model1/model2: full grid; column_weight1/2: the rank's slab; der_type 1 (forward) or
 * 2 (central), anything else aborts like the reference (:281-283). Outputs: cost[3] (:298-300) and, when not NULL,
 * cross_grad(nx*ny*nz) = |tau| per cell (:293-296). */
int tfx_cross_gradient_calculate(tfx_matrix *matrix, int32_t nrows, double *b_RHS, int32_t nx, int32_t ny, int32_t nz,
                                 const double *dX, const double *dY, const double *dZ, int32_t nparams_loc,
                                 const double *model1, const double *model2, const double *column_weight1,
                                 const double *column_weight2, int32_t der_type, double glob_weight,
                                 const int32_t keep_model_constant[2], int32_t myrank, int32_t nbproc,
                                 double cost[3], double *cross_grad);
/* t_admm_method%iterate_admm_arrays (src/inversion/admm_method.F90:70-134): z = P_C(x + u), u += x - z,
 * x0 = z - u; xmin/xmax(nlithos, nelements) Fortran order; z and u are updated in place. */
int tfx_admm_iterate_admm_arrays(int32_t nelements, int32_t nlithos, const double *xmin, const double *xmax,
                                 const double *x, double *z, double *u, double *x0);

/* ---- module lsqr_solver (src/inversion/lsqr_solver2.F90) --------------------------------------- */
int tfx_lsqr_solve(int32_t nlines, int32_t nelements, int32_t niter, double rmin, double gamma,
                   tfx_matrix *matrix, double *u, double *x, int32_t myrank);                    /* :321-473 */
/* u (the right-hand side b_RHS on entry) is the solver's work array and is destroyed, like in the reference. With
 * nbproc > 1 a HOST u is read and written back only on the data rows and on this rank's constraint rows (the rows of
 * matrix_cons it stores, the rows shared with a neighbouring slab, on rank 0 also the rows stored nowhere); a HOST x only
 * on the columns of the problems being solved (zero elsewhere). */
int tfx_lsqr_solve_sensit(int32_t nlines, int32_t ncolumns, int32_t niter, double rmin, double gamma,
                          double target_misfit, tfx_matrix *matrix_sensit, tfx_matrix *matrix_cons,
                          double *u, double *x, const int32_t solve_problem[2], int32_t nelements,
                          int32_t nx, int32_t ny, int32_t nz, int32_t ncomponents,
                          int32_t compression_type, int32_t wavelet_domain, double *memory,
                          int32_t myrank, int32_t nbproc);                                       /* :47-308 */
/* Diagnostics of the last solve: r = phibar/b1 after every executed iteration (the reference only
 * prints the final one, :302-306), number of executed iterations, 1 if the fused single-sweep path
 * ran. r_hist may be NULL. */
int tfx_lsqr_last_history(double *r_hist, int32_t capacity, int32_t *iters, int32_t *fused);
/* Loop bodies executed by the last solve and the iteration count the reference prints (`iter - 1`, :302-306 / :467):
 * they differ by one only when lsqr_solve leaves through its small-rhobar exit, which sits before `iter = iter + 1`
 * (:459-465; lsqr_solve_sensit checks after it, :281-289). */
int tfx_lsqr_last_iterations(int32_t *executed, int32_t *reported);
/* Device time of the last solve measured with CUDA events on the library's stream: the iteration loop
 * (excluding the initialisation before the reference's `do while`, lsqr_solver2.F90:120-157) and, with
 * option "profile_sweeps" = 1, the summed duration / count of the fused sweep kernel launches. */
int tfx_lsqr_last_timing(double *loop_ms, double *sweep_ms, int32_t *nsweeps);

/* ---- module sensitivity_gravmag (src/forward/gravmag/sensitivity_gravmag.F90) ------------------ */
typedef struct tfx_sensit_params {
  int32_t problem_type;        /* 1 gravity, 2 magnetic (select type at :126-137)                    */
  int32_t nx, ny, nz;          /* full grid                                                          */
  int32_t ndata;               /* number of stations                                                 */
  int32_t ndata_components;    /* par%ndata_components                                               */
  int32_t nmodel_components;   /* par%nmodel_components                                              */
  int32_t data_type;           /* gravity: 1 gz, 2 gradiometry (zz only with 1 data component)       */
  int32_t compression_type;    /* 0 none, 1 Haar, 2 Daubechies D4                                    */
  double  compression_rate;    /* :64-77                                                             */
  double  problem_weight;      /* ipar%problem_weight(problem)                                       */
  double  mi, md, theta, intensity;   /* magnetic field (magnetic_field.f90:64-110)                  */
  int32_t cell0, ncells_local; /* 0-based first local cell and count (column slab of this rank);
                                  compression requires the full grid on every rank                   */
  int32_t param_shift;         /* column shift of this problem inside the joint matrix (:685-686)    */
  int32_t ncolumns;            /* total columns of matrix_sensit (2*ncomp*nelements, jip.F90:213)    */
} tfx_sensit_params;

/* calculate_and_write_sensit (:82-410) + read_sensitivity_kernel (:648-883) without the disk round
 * trip: evaluates the kernels on the GPU, applies column weight / wavelet / threshold / real(4)
 * rounding / problem*data weight exactly in the reference's order and leaves the finalized matrix
 * on the device. grid arrays hold the FULL grid (nx*ny*nz cells); column_weight_full(nx*ny*nz);
 * data_weight(ndata_components, ndata). Outputs (may be NULL): sensit_nnz(nx*ny*nz) per-column
 * counts (:267,293), comp_error (:350), nnz_total. */
int tfx_calculate_sensit(tfx_matrix **matrix_sensit, const tfx_sensit_params *par,
                         const double *X1, const double *X2, const double *Y1, const double *Y2,
                         const double *Z1, const double *Z2,
                         const double *data_X, const double *data_Y, const double *data_Z,
                         const double *column_weight_full, const double *data_weight,
                         int32_t *sensit_nnz, double *comp_error, int64_t *nnz_total);

/* ---- multi-GPU assembly: rows sharded by data, re-partitioned to nnz-balanced column slabs ---------
 * The reference's three stages with the per-rank files kept in HBM and the per-row MPI_Scatterv
 * replaced by one all-to-all over NVLink (csrc/sensit_dist.cu):
 *   calculate_and_write_sensit  (sensitivity_gravmag.F90:82-410)  -> tfx_sensit_assemble_rows
 *   get_load_balancing_nelements / calculate_new_partitioning (:470-524, :573-642)
 *                                                                  -> tfx_get_load_balancing_nelements
 *   read_sensitivity_kernel     (:648-883)                         -> tfx_sensit_repartition            */
typedef struct tfx_sensit_rows tfx_sensit_rows;   /* stands in for the file sensit_<type>_<nbproc>_<rank> */

/* Row pipeline for the stations of rank `myrank` (even split of par->ndata, parallel_tools.f90:46-86).
 * par->cell0 / ncells_local / param_shift / ncolumns are ignored (full grid, columns k*N + p).
 * Outputs are reduced over the communicator when nbproc > 1 (requires tfx_comm_init with nbproc ranks):
 * sensit_nnz(nx*ny*nz) (:322), nnz_total (:327), comp_error (:346-353). data_weight and problem_weight
 * are applied in real(4) exactly like read_sensitivity_kernel does (:837-843). */
int tfx_sensit_assemble_rows(tfx_sensit_rows **rows, const tfx_sensit_params *par,
                             const double *X1, const double *X2, const double *Y1, const double *Y2,
                             const double *Z1, const double *Z2,
                             const double *data_X, const double *data_Y, const double *data_Z,
                             const double *column_weight_full, const double *data_weight,
                             int32_t myrank, int32_t nbproc,
                             int32_t *sensit_nnz, double *comp_error, int64_t *nnz_total);
/* Multiplies the values of a row set assembled with unit weights by real(problem_weight * data_weight(d, i), 4)
 * in real(4) (:837-843): what read_sensitivity_kernel does after reading the unweighted file. */
int tfx_sensit_rows_apply_weights(tfx_sensit_rows *rows, double problem_weight, const double *data_weight);
int tfx_sensit_rows_info(const tfx_sensit_rows *rows, int32_t *data0, int32_t *ndata_loc, int64_t *nnz_local);
int tfx_sensit_rows_destroy(tfx_sensit_rows *rows);
/* get_load_balancing_nelements (:470-524): host integer work, no GPU needed. Joint inversions pass the
 * sum of both problems' sensit_nnz (:610-625). */
int tfx_get_load_balancing_nelements(int32_t nelements_total, const int32_t *sensit_nnz, int32_t nbproc,
                                     int64_t *nnz_at_cpu_new, int32_t *nelements_at_cpu_new);
/* Builds this rank's column slab (cells nsmaller+1 .. nsmaller+nelements_at_cpu(myrank+1), all data rows,
 * LOCAL columns (p - nsmaller) + (k-1)*nelements + param_shift(problem_slot), :685-686,:834; ncolumns =
 * 2*nmodel_components*nelements, joint_inverse_problem.F90:213-214) and consumes the row set.
 * Without a communicator (single process) the row set must hold all data rows (assembled with
 * nbproc = 1) and the slab of any (myrank, nbproc) can be built -- used by tests and by hosts that drive
 * the slabs one after the other. */
int tfx_sensit_repartition(tfx_matrix **matrix_sensit, tfx_sensit_rows *rows, int32_t problem_slot,
                           const int32_t *nelements_at_cpu, int32_t myrank, int32_t nbproc);

/* Same, appending the rows to a matrix under construction (tfx_sparse_matrix_initialize ... finalize): the
 * reference calls read_sensitivity_kernel once per problem on jinv%matrix_sensit
 * (problem_joint_gravmag.F90:241-248); finalize() then builds the device representations. */
int tfx_sensit_repartition_into(tfx_matrix *matrix_sensit, tfx_sensit_rows *rows, int32_t problem_slot,
                                const int32_t *nelements_at_cpu, int32_t myrank, int32_t nbproc);

/* ---- the reference's on-disk sensitivity formats (csrc/sensit_io.cu) -------------------------------
 * `dir` is the SENSIT folder (path_output/SENSIT or par%sensit_path). Byte order is big-endian like the
 * reference build (-fconvert=big-endian, Makefile:51). */
int tfx_create_sensit_directory(const char *dir);                               /* file_utils.F90:31-40 */
/* The rank's stream file sensit_<grav|magn>_<nbproc>_<rank> (:143-148,:183,:306-309). The row set must have
 * been assembled with problem_weight = 1 and unit data weights (the file stores the unweighted kernel). */
int tfx_write_sensit_file(const tfx_sensit_rows *rows, const char *dir);
/* sensit_<..>_meta.txt (:359-376) and sensit_<..>_nnz (:381-392; skipped when sensit_nnz is NULL). Rank 0 only. */
int tfx_write_sensit_metadata(const tfx_sensit_params *par, const char *dir, int32_t nbproc,
                              int32_t depth_weighting_type, double comp_error, int64_t nnz_total,
                              const int32_t *sensit_nnz);
/* read_sensitivity_metadata (:974-1037) incl. its consistency checks; outputs may be NULL. */
int tfx_read_sensitivity_metadata(const tfx_sensit_params *par, const char *dir, int32_t depth_weighting_type,
                                  int32_t *nbproc_sensit, double *comp_error, int64_t *nnz_total);
int tfx_read_sensit_nnz(const tfx_sensit_params *par, const char *dir, int32_t *sensit_nnz);   /* :529-568 */
int tfx_write_depth_weight(const tfx_sensit_params *par, const char *dir, const double *column_weight_full); /* :415-465 */
int tfx_read_depth_weight(const tfx_sensit_params *par, const char *dir, double *column_weight_full);        /* :888-969 */
/* read_sensitivity_kernel (:648-883): scans the files of all writer ranks, keeps the column slab of `myrank`
 * (nelements_at_cpu from tfx_get_load_balancing_nelements), applies the index shift (:834) and the real(4)
 * weights (:837-843) and leaves the finalized matrix on the device. */
int tfx_read_sensitivity_kernel(tfx_matrix **matrix_sensit, const tfx_sensit_params *par, const char *dir,
                                const double *data_weight, int32_t depth_weighting_type, int32_t problem_slot,
                                int32_t myrank, int32_t nbproc, const int32_t *nelements_at_cpu,
                                int64_t *nnz_local);

int tfx_read_sensitivity_kernel_into(tfx_matrix *matrix_sensit, const tfx_sensit_params *par, const char *dir,
                                     const double *data_weight, int32_t depth_weighting_type, int32_t problem_slot,
                                     int32_t myrank, int32_t nbproc, const int32_t *nelements_at_cpu,
                                     int64_t *nnz_local);

/* One raw sensitivity line per station (no weighting), for kernel parity tests:
 * lines(ncells, nmodel_components, ndata_components, ndata_batch) Fortran order. */
int tfx_sensit_lines(const tfx_sensit_params *par, const double *X1, const double *X2, const double *Y1,
                     const double *Y2, const double *Z1, const double *Z2, int32_t ndata_batch,
                     const double *data_X, const double *data_Y, const double *data_Z, double *lines);

/* Keeps a device copy of the model grid (grid_type: X1..Z2 of the ncells cells, src/model/grid.F90) for the following
 * tfx_calculate_sensit / tfx_sensit_assemble_rows / tfx_sensit_lines calls that pass THE SAME six host arrays, instead of
 * uploading 48 B per cell on every call (row blocks, the two problems of a joint inversion). The caller must not change
 * the arrays while they are pinned. tfx_grid_unpin releases the copy. */
int tfx_grid_pin(int32_t ncells, const double *X1, const double *X2, const double *Y1, const double *Y2,
                 const double *Z1, const double *Z2);
int tfx_grid_unpin(void);

/* Diagnostics for tests/test_gpu_mathx.py: the forward kernels' own log / atan2 (csrc/mathx.cuh: CUDA's algorithms with
 * the polynomial coefficients read from the constant bank) next to the CUDA library's, element by element (host arrays):
 * out = [tfx_log(x) | log(x) | tfx_atan2(y, x) | atan2(y, x)], 4 * n doubles. */
int tfx_debug_math(int64_t n, const double *y, const double *x, double *out);

#ifdef __cplusplus
}
#endif
#endif /* TFX_H */
