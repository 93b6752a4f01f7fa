// matrix.h -- the t_sparse_matrix replacement (host builder + device representations).
#pragma once

#include <stdint.h>

#include <vector>

#include "kernels.h"

namespace tfx {

// Matrix entries sorted by (row, column) on the device: the currency between the row pipeline, the
// re-partitioner and the matrix builder.
struct RowTriplets {
  DevBuf<int32_t> idx;     // 0-based column
  DevBuf<int32_t> rowid;   // 0-based matrix row
  DevBuf<float> val;
  int64_t nnz = 0;
};

// Mirrors t_sparse_matrix (src/inversion/sparse_matrix.f90:31-98). The host-side builder keeps the
// reference's exact storage (sa real(4), ija int32 1-based, ijl int64 1-based, rowptr int32);
// finalize() validates it like the reference and mirrors it to the device:
//   fwd : CSR of A    (forward products, gathers x by column)
//   trn : CSR of A^T  (transposed products as gathers -- no scatter-add, deterministic)
//   dense (optional): column-major f32 block when every stored row holds the same contiguous
//                     column range (the uncompressed kernel), or when assembled on the device.
struct Matrix {
  int64_t nnz = 0, nel = 0, nel_last = 0;
  int32_t nl = 0, nl_nonempty = 0, nl_nonempty_allocated = 0, nl_current = 0, nl_current_all = 0, ncolumns = 0;
  std::vector<float> sa;
  std::vector<int32_t> ija;
  std::vector<int64_t> ijl;
  std::vector<int32_t> rowptr;
  bool finalized = false;
  bool device_only = false;   // assembled on the device: no host arrays, builder calls are errors
  int tag = 0;

  SegMatrix fwd, trn;
  bool has_seg = false;
  T16Matrix t16f, t16t;       // tiled 16-bit layouts of A (forward) and A^T (transposed), big matrices only
  bool has_t16 = false;
  // Device-resident rows appended by read_sensitivity_kernel / the re-partitioner before finalize()
  // (the reference appends one problem after the other with add_row / new_row, sensitivity_gravmag.F90:846-853).
  RowTriplets pend;
  DenseCM dense;
  bool has_dense = false;
  int32_t dense_row0 = 0;     // 0-based global row of dense row 0

  // Row-blocked storage: a big compressed kernel assembled batch of stations after batch of stations (option
  // "sensit_row_blocks"). Block b is an independent finalized device matrix holding the matrix rows
  // [block_row0[b], block_row0[b] + blocks[b]->nl); S x is the concatenation of the blocks' products, S^T u the sum.
  // Building a block needs ~3x its own footprint for a moment, never 3x the whole matrix.
  std::vector<Matrix *> blocks;
  std::vector<int32_t> block_row0;
  bool has_blocks = false;

  Matrix() {}
  Matrix(const Matrix &) = delete;
  Matrix &operator=(const Matrix &) = delete;
  ~Matrix() { clear_blocks(); }
  void clear_blocks() {
    for (Matrix *b : blocks) delete b;
    blocks.clear(); block_row0.clear(); has_blocks = false;
  }
  int64_t device_nnz() const {
    if (has_blocks) { int64_t n = 0; for (const Matrix *b : blocks) n += b->device_nnz(); return n; }
    return has_dense ? (int64_t)dense.nrows * dense.ncols : fwd.nnz;
  }
};

// Products with a compressed matrix (T16 / CSR representations, row blocks); the dense block has its own sweep.
//   matrix_fwd  : y(nl) (+)= S x, x read at (column - xshift)
//   matrix_trans: y(ncolumns) (+)= S^T u
int matrix_fwd(Matrix &m, const double *d_x, double *d_y, bool accumulate, int32_t xshift, const int *d_done, cudaStream_t st);
int matrix_trans(Matrix &m, const double *d_u, double *d_y, bool accumulate, const int *d_done, cudaStream_t st);
// Appends the rows held as device triplets (row ids relative to the batch) as one more row block.
int matrix_append_block(Matrix &M, RowTriplets &R, int32_t nrows, int32_t ncolumns);
extern int g_opt_sensit_row_blocks;
extern int g_opt_trace;
void trace(const char *label);
extern int g_opt_sensit_cand_cap;   // > 0: candidate-list capacity of the k-th select (tests force the fallback passes with 1)

int matrix_upload(Matrix &m, bool allow_dense);
// Builds the T16 layouts from fwd/trn when the matrix is big enough (option "t16_min_nnz").
int matrix_build_t16(Matrix &m);

// ---- sensit.cu / sensit_dist.cu -----------------------------------------------------------------
}  // namespace tfx
#include "../../include/tfx.h"
namespace tfx {
int assemble_rows_device(const tfx_sensit_params &P, const GridDev &g, const double *d_dx, const double *d_dy,
                         const double *d_dz, const double *d_cw, const double *h_dw, int32_t data0, int32_t ndata_loc,
                         RowTriplets &R, DevBuf<int32_t> &dnnz, std::vector<long long> &seg_end, double *err_sum);
int matrix_from_triplets(Matrix &M, int32_t nl, int32_t ncolumns, RowTriplets &R);
// Appends nrows matrix rows held as device triplets (row ids relative to the appended block) to a matrix that
// is still being built; finalize() turns the accumulated rows into the device representations.
int matrix_append_triplets(Matrix &M, RowTriplets &R, int32_t nrows, int32_t ncolumns);
// Moves the rows built on the host so far into the pending device rows (see sensit.cu).
int matrix_flush_host_rows(Matrix &M);
// The device copy of a grid for one call: the pinned one (tfx_grid_pin, same host arrays) or a fresh upload.
struct GridHold {
  GridDev own;
  GridDev *g = nullptr;
};
int grid_acquire(GridHold &h, int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                 const double *Z1, const double *Z2, int32_t nx, int32_t ny, int32_t nz);
int upload_grid(GridDev &g, int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                const double *Z1, const double *Z2);

// ---- lsqr.cu ----------------------------------------------------------------------------------
struct LsqrParams {
  int32_t nlines = 0, ncolumns = 0, niter = 0;
  double rmin = 0, gamma = 0, target_misfit = 0;
  int32_t solve_problem[2] = {1, 0};
  int32_t nelements = 0, nx = 0, ny = 0, nz = 0, ncomponents = 1, compression_type = 0;
  bool wavelet_domain = true;
  int32_t myrank = 0, nbproc = 1;
  bool single_matrix = false;   // lsqr_solve (tests): no constraint matrix, no wavelet
  // Caller's host vectors (null when the caller passed device memory): the solver copies only the rows of u and the
  // columns of x it works on (data rows + this rank's constraint rows; the active problems' columns).
  double *host_u = nullptr, *host_x = nullptr;
};

struct LsqrResult {
  int32_t iters = 0;      // loop bodies executed
  int32_t reported_iters = 0;   // the reference's printed `iter - 1` (one less than iters on lsqr_solve's small-rhobar exit)
  int32_t status = 0;     // 0 ok, 1: |b| = 0
  double r = 1.0;
  bool fused = false;
  double loop_ms = 0.0;   // device time of the iteration loop (CUDA events on the library stream)
  double sweep_ms = 0.0;  // summed device time of the fused sweep kernel launches (option profile_sweeps)
  int32_t nsweeps = 0;
  std::vector<double> history;
};

// Option "strict_order": reproduce the reference's sequential summation order (slow parity mode).
extern int g_opt_strict_order;
extern int g_opt_profile_sweeps;
// Option "lsqr_graph": 1 (default) replays the split-path iteration body as a CUDA graph for small matrices.
extern int g_opt_lsqr_graph;
// Option "lsqr_poll": iterations between two reads of the device-side done flag (default 8).
extern int g_opt_lsqr_poll;

int lsqr_run(const LsqrParams &p, Matrix *S, Matrix *C, double *d_u, double *d_x, LsqrResult &res);
int lsqr_run_strict(const LsqrParams &p, Matrix *S, Matrix *C, double *d_u, double *d_x, LsqrResult &res);

// ---- comm.cu ----------------------------------------------------------------------------------
int comm_nranks();
int comm_rank();
// Position of this rank's slab in the concatenation of all ranks' slabs (get_nsmaller with arbitrary slab sizes,
// parallel_tools.f90:68-86 + :91-110) and the total; one small all-gather.
int comm_slab_offset(int64_t mine, int64_t *offset, int64_t *total);
int comm_allreduce_max(double *d_buf, size_t count, cudaStream_t st);
// In-place wavelet transform of a distributed volume: d_slab holds cells [nsmaller, nsmaller + nelements) of the
// nx*ny*nz volume (wavelet_utils.F90:57-67 without the rank-0 bottleneck). data.cu
int wavelet_slab_device(double *d_slab, int64_t nelements, int64_t nsmaller, int nx, int ny, int nz, int wavelet_type,
                        bool forward, cudaStream_t st);
// The same with the slab layout of all ranks known (offsets: nranks + 1 prefix entries, comm_slab_offsets): no host
// synchronisation inside, usable in the LSQR loop / under stream capture.
int wavelet_slab_device_off(double *d_slab, const std::vector<int64_t> &offsets, int nx, int ny, int nz, int wavelet_type,
                            bool forward, cudaStream_t st);
int comm_slab_offsets(int64_t mine, std::vector<int64_t> &offsets);
int comm_allgatherv_f64(double *d_full, const int64_t *offsets, cudaStream_t st);
int comm_exchange_f64(const double *d_send, const int64_t *send_off, const int64_t *send_cnt, double *d_recv,
                      const int64_t *recv_off, const int64_t *recv_cnt, cudaStream_t st);
int comm_allreduce_sum(double *d_buf, size_t count, cudaStream_t st);
int comm_allreduce_sum_i32(int32_t *d_buf, size_t count, cudaStream_t st);
int comm_allreduce_sum_u8(uint8_t *d_buf, size_t count, cudaStream_t st);
int comm_allreduce_sum_i64(int64_t *d_buf, size_t count, cudaStream_t st);
int comm_allgather_i64(const int64_t *d_send, int64_t *d_recv, size_t count, cudaStream_t st);
int comm_alltoallv_4b(const void *d_send, const int64_t *send_off, void *d_recv, const int64_t *recv_off,
                      cudaStream_t st);
int comm_unique_id(char id[128]);
int comm_init(int nranks, int rank, const char id[128]);
int comm_finalize();

}  // namespace tfx

// Row-sharded kernel of one problem, resident on the device (stands in for the file
// sensit_<type>_<nbproc>_<rank>; csrc/sensit_dist.cu builds it, csrc/sensit_io.cu writes / reads the file).
struct tfx_sensit_rows {
  tfx_sensit_params par;
  int32_t data0 = 0, ndata_loc = 0;   // stations [data0, data0 + ndata_loc) of par.ndata
  int32_t myrank = 0, nbproc = 1;
  bool unit_weights = true;           // problem_weight * data_weight == 1 for every row (file content is unweighted)
  tfx::RowTriplets R;                 // idx = k*N + p (0-based, no problem shift), rowid = global matrix row
  std::vector<long long> seg_end;     // running entry count after each (idata, d, k) segment
};

// The opaque handle of include/tfx.h.
struct tfx_matrix {
  tfx::Matrix m;
};
