// kernels.h -- internal interfaces between the translation units of libtfx.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace tfx {

// ---- wavelet.cu -------------------------------------------------------------------------------
// In-place 3-D transform of a device-resident Fortran-ordered volume s(n1,n2,n3).
int wavelet3d_device(double *d_s, int n1, int n2, int n3, int wavelet_type, bool forward, cudaStream_t st);
// The same transform applied to nvol volumes stored back to back (one set of three launches).
int wavelet3d_device_batch(double *d_s, int n1, int n2, int n3, long long nvol, int wavelet_type, bool forward,
                           cudaStream_t st);

extern int g_opt_wavelet_slab_mb;
extern int g_opt_wavelet_fuse12;
extern int g_opt_wavelet_cols;
extern int g_opt_wavelet_tile_kb;

int wavelet_axis_device(double *d_s, int L, long long inner, long long outer, int wavelet_type, bool forward,
                        cudaStream_t st);
int wavelet_axes12_device(double *d_s, int n1, int n2, long long nplanes, int wavelet_type, bool forward, cudaStream_t st);
extern int g_opt_wavelet_dist;
extern int g_opt_wavelet_p2p;   // 1: layout changes of the distributed transform through peer memory (cudaIpc), 0: NCCL
void wavelet_peer_reset();

// ---- csr.cu -----------------------------------------------------------------------------------
// A compressed-segment matrix view on the device. For the forward product it is the CSR of A
// (segments = stored rows, idx = column); for the transposed product it is the CSR of A^T
// (segments = non-empty columns, idx = global row). Long segments are cut into work items of at
// most kItemLen entries; each item is summed by one warp/CTA in a fixed order and the per-segment
// partials are added in item order, so every product is run-to-run deterministic.
struct SegMatrix {
  int64_t nnz = 0;
  int32_t nseg = 0;        // stored (non-empty) segments
  int32_t nout = 0;        // length of the output vector
  int32_t nin = 0;         // length of the input vector
  int32_t nitems = 0;
  int32_t max_items_per_seg = 0;
  double avg_len = 0.0;
  DevBuf<int64_t> ptr;     // [nseg+1] 0-based entry offsets
  DevBuf<int32_t> idx;     // [nnz]    0-based input index
  DevBuf<float> val;       // [nnz]
  DevBuf<int32_t> segmap;  // [nseg]   0-based output index of each segment
  DevBuf<int32_t> item_seg;    // [nitems]
  DevBuf<int64_t> item_beg;    // [nitems+1] (item i covers [item_beg[i], item_end[i]))
  DevBuf<int64_t> item_end;    // [nitems]
  DevBuf<int32_t> seg_item0;   // [nseg+1] first item of each segment
  DevBuf<double> partial;      // [nitems]
  // identity_items: no segment is longer than kItemLen, so item i IS segment i (beg = ptr[i], end = ptr[i+1]) and the
  // four item tables are not materialised -- the case of every constraint matrix and of the column-major copies of
  // compressed kernels (tens of millions of short segments: the tables cost 24 B per segment and a host loop to build).
  bool identity_items = false;
  bool empty() const { return nnz == 0 || nseg == 0; }
  void release() {
    ptr.release(); idx.release(); val.release(); segmap.release(); item_seg.release(); item_beg.release();
    item_end.release(); seg_item0.release(); partial.release();
    nnz = 0; nseg = 0; nitems = 0; identity_items = false;
  }
};

static const int kItemLen = 8192;

// Builds the item table from host copies of ptr (0-based, nseg+1 entries).
int seg_build_items(SegMatrix &m, const int64_t *h_ptr);
// The same when ptr only exists on the device and the longest segment (max_len) is known.
int seg_set_identity_items(SegMatrix &m);

// y[segmap[s] - out_lo] (+)= sum_k val[k] * x[idx[k] - xshift], restricted to segments whose output
// index lies in [out_lo, out_hi). accumulate == false zeroes y[0 .. out_hi-out_lo) first.
// `done` (device flag, may be null): kernels exit immediately when *done != 0.
int seg_spmv(SegMatrix &m, const double *d_x, double *d_y, bool accumulate, int32_t out_lo, int32_t out_hi,
             int32_t xshift, const int *d_done, cudaStream_t st);

// y += x (device vectors).
int vec_add_inplace(double *y, const double *x, size_t n, cudaStream_t st);

// ---- t16.cu -----------------------------------------------------------------------------------
// Tiled layout with 16-bit in-tile indices (see t16.cu). One instance serves one product direction.
enum T16Mode { T16_DIRECT = 0, T16_TILES = 1 };
struct T16Matrix {
  bool valid = false;
  int64_t nnz = 0, nnz_padded = 0;
  int32_t nout_total = 0;  // length of the output vector
  int32_t out0 = 0, nseg = 0;      // outputs covered: [out0, out0 + nseg)
  int32_t in0 = 0, nin = 0;        // gathered elements covered: [in0, in0 + nin)
  int32_t tile = 0, ntiles = 0;
  int32_t nsplit = 1;              // TILES: work items per tile
  T16Mode mode = T16_DIRECT;
  DevBuf<float> val;               // [nnz_padded]
  DevBuf<uint16_t> key;            // [nnz_padded]
  DevBuf<int64_t> ptr;             // [ntiles * nseg + 1], tile-major
  DevBuf<double> partial;          // TILES: [ntiles][nseg], written once per product
  DevBuf<int> counter;             // dynamic work distribution
  void release() {
    valid = false;
    val.release(); key.release(); ptr.release(); partial.release();
    nnz = nnz_padded = 0; nseg = ntiles = 0;
  }
  int64_t bytes() const { return nnz_padded * 6 + ((int64_t)ntiles * nseg + 1) * 8 + (int64_t)partial.n * 8; }
};
extern int g_opt_t16_min_nnz;
extern int g_opt_t16_tile;
extern int g_opt_t16_async;
extern int g_opt_t16_direct_max;
extern int g_opt_t16_bank_deal;
extern int g_opt_t16_blk;
extern int g_opt_t16_tma;
extern int g_opt_t16_long_seg;
// Builds the layout from a compressed-segment matrix (segments = outputs, idx = gathered index). Leaves
// T.valid == false (and returns 0) when the source does not qualify (indices not strictly ascending).
int t16_build(const SegMatrix &src, T16Matrix &T, cudaStream_t st);
// y[out0 .. out0+nseg) (+)= A x ; when !accumulate the rest of y[0 .. nout_total) is zeroed.
// Element g of the layout's gathered range is read from d_x[g - xshift].
int t16_spmv(T16Matrix &m, const double *d_x, double *d_y, bool accumulate, int32_t xshift, const int *d_done,
             cudaStream_t st);

// ---- dense.cu ---------------------------------------------------------------------------------
// Uncompressed sensitivity block: column-major f32, column j at base + j*ld (ld % 4 == 0), rows
// [0, nrows). No column indices are stored (columns are 1..N, sensitivity_gravmag.F90:288-295).
struct DenseCM {
  DevBuf<float> val;
  int64_t ld = 0;
  int32_t nrows = 0;       // data rows of this block
  int32_t ncols = 0;       // local columns held by this rank
  int32_t col0 = 0;        // 0-based position of column 0 inside the solver's column space
  int32_t grid = 0;        // CTAs of the sweep kernel (== partial buffers)
  int32_t fastcvt_ok = -1; // 1: no zero / subnormal entries (integer f32->f64 path allowed); -1: not scanned yet
  DevBuf<double> partial_q;    // [grid][ld]
  DevBuf<double> partial_n2;   // [grid]
  bool empty() const { return nrows == 0 || ncols == 0; }
};

enum DenseMode { DENSE_FUSED = 0, DENSE_T_ONLY = 1, DENSE_F_ONLY = 2 };

extern int g_opt_dense_stream_only;      // diagnostic: stream the block through the TMA ring without the products
extern int g_opt_dense_f2f_rows;         // row vectors per thread converted with F2F on their second use
extern int g_opt_dense_vec4;             // 1: 512 threads x float4 rows (default), 0: 1024 threads x float2 rows
static const int kDenseMaxRows = 10240;  // register-resident u / q: 20 rows per thread x 512 threads

// One sweep over the block.
//  FUSED : out[j] = nbeta*v[j] + (S^T u)[j] + g[j];  q += S out;  n2 = |out|^2   (nbeta = *d_nbeta)
//  T_ONLY: out[j] = (S^T u)[j]
//  F_ONLY: q += S in  (in = d_v)
// Vectors v/g/out are indexed in the solver's column space (col0 applied inside). g may be null.
// q (nrows) and n2 (1) are written (not accumulated) by the trailing reduction kernel.
int dense_sweep(DenseCM &S, DenseMode mode, const double *d_u, const double *d_v, const double *d_g,
                double *d_out, const double *d_nbeta, double *d_q, double *d_n2, const int *d_done,
                cudaStream_t st, bool accumulate = false);   // accumulate: DENSE_T_ONLY adds to d_out
extern int g_opt_dense_block_rows;   // rows per dense row block (<= kDenseMaxRows; tests lower it)

// ---- assembly.cu ------------------------------------------------------------------------------
struct GridDev {
  int32_t n = 0;
  DevBuf<double> X1, X2, Y1, Y2, Z1, Z2;
  // Tensor-product view, filled by grid_detect_structured(): the cell boxes are exactly (bit for bit) the products
  // of node coordinates xn(nx+1) x yn(ny+1) x zn(nz+1), cells ordered i fastest (model_IO.F90:184-193). Neighbouring
  // cells then share their corners and a prism corner term needs to be evaluated once per NODE instead of once per
  // cell corner (8x fewer sqrt/atan2/log evaluations, bit-identical sums).
  int32_t nx = 0, ny = 0, nz = 0;
  int structured = -1;          // -1 not examined, 0 no, 1 yes
  DevBuf<double> xn, yn, zn;
};
int grid_detect_structured(GridDev &g, int32_t nx, int32_t ny, int32_t nz, cudaStream_t st);
extern int g_opt_grav_shared_nodes;
int debug_math(long long n, const double *d_y, const double *d_x, double *d_out, cudaStream_t st);
extern int g_opt_mag_shared_nodes;

// Fills a dense column-major block with the depth-weighted gravity kernel, reproducing
// graviprism_z (gravity_field.f90:131-195), apply_column_weight (sensitivity_gravmag.F90:228),
// the real(4) store (:290) and the real(4) scaling by problem_weight*data_weight (:837-843).
int assemble_grav_dense(DenseCM &S, const GridDev &g, int32_t cell0, int32_t ncells, int32_t ndata,
                        const double *d_xd, const double *d_yd, const double *d_zd, const double *d_cw,
                        const double *d_dw, double problem_weight, int *d_err, cudaStream_t st);

// Computes one gravity sensitivity line (all cells) per data point into d_lines[b*ncells + p].
// d_cw != nullptr (only when grav_lines_fused_partials() > 0): the lines come out multiplied by the column weight and
// d_partial[b * npartials + i] holds partial sums of their squares (cost_full, sensitivity_gravmag.F90:228-234).
int grav_lines(const GridDev &g, int32_t ndata_batch, const double *d_xd, const double *d_yd, const double *d_zd,
               int data_type, double *d_lines, int *d_err, cudaStream_t st, const double *d_cw = nullptr,
               double *d_partial = nullptr);
int grav_lines_fused_partials(const GridDev &g, int data_type);

// Full-tensor gravity gradiometry lines (gradiprism_full): d_lines[(b*6 + d)*ncells + p], d = XX, YY, ZZ, XY, YZ, ZX.
int grav_full_lines(const GridDev &g, int32_t ndata_batch, const double *d_xd, const double *d_yd, const double *d_zd,
                    double *d_lines, int *d_err, cudaStream_t st);

// Magnetic tensor lines: d_lines[((b*ndc + d)*nmc + k)*ncells + p] (Fortran sensit_line(p,k,d) per data b).
int mag_lines(const GridDev &g, int32_t ndata_batch, const double *d_xd, const double *d_yd, const double *d_zd,
              int nmc, int ndc, double mi, double md, double theta, double intensity, double *d_lines,
              int *d_err, cudaStream_t st);

}  // namespace tfx
