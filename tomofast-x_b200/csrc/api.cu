// api.cu -- the C ABI (include/tfx.h), the t_sparse_matrix replacement and the runtime context.
#include "../../include/tfx.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

namespace tfx {

// ---------------------------------------------------------------------------------------------
// context / errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int g_opt_dense_detect = 1;
int g_opt_profile_sweeps = 0; // 1: CUDA events around every fused sweep launch (bench roofline)
int g_opt_lsqr_graph = 1;     // 1: CUDA-graph replay of the split-path iteration body (launch-bound small matrices)
int g_opt_strict_order = 0;   // 1: LSQR uses the reference's sequential summation order (parity mode)

void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

Context &ctx() {
  static Context c;
  return c;
}

static int init_device(int device) {
  Context &c = ctx();
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(-1, "libtfx: no CUDA device available (there is no CPU fallback)");
  }
  if (device < 0) {
    const char *lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) % ndev : 0;
  }
  if (device >= ndev) return fail(-2, "libtfx: device index out of range");
  if (c.ready && c.device == device) return 0;
  TFX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TFX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(-3, std::string("libtfx is built for sm_100a (B200); found ") + prop.name);
  c.device = device;
  c.num_sms = prop.multiProcessorCount;
  if (!c.stream) {
    TFX_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      TFX_CUDA(cudaStreamCreateWithFlags(&c.side[i], cudaStreamNonBlocking));
      TFX_CUDA(cudaEventCreateWithFlags(&c.ev_join[i], cudaEventDisableTiming));
    }
    TFX_CUDA(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
  }
  c.ready = true;
  return 0;
}

int ensure_init() {
  if (ctx().ready) return 0;
  return init_device(-1);
}

// Pool of staging buffers (at most kPoolMax kept, largest first out).
namespace {
struct StageBuf { double *p; size_t cap; };
std::vector<StageBuf> g_stage_pool;
const size_t kPoolMax = 6;
}  // namespace

VecIO::~VecIO() {
  if (!stage) return;
  if (g_stage_pool.size() < kPoolMax) g_stage_pool.push_back({stage, stage_cap});
  else cudaFree(stage);
  stage = nullptr;
}

int VecIO::bind(double *ptr, size_t count, bool copy_in) {
  n = count;
  if (is_device_ptr(ptr)) {
    dev = ptr;
    host = nullptr;
    return 0;
  }
  host = ptr;
  // smallest pooled buffer that fits
  int best = -1;
  for (size_t i = 0; i < g_stage_pool.size(); ++i)
    if (g_stage_pool[i].cap >= count && (best < 0 || g_stage_pool[i].cap < g_stage_pool[best].cap)) best = (int)i;
  if (best >= 0) {
    stage = g_stage_pool[best].p; stage_cap = g_stage_pool[best].cap;
    g_stage_pool.erase(g_stage_pool.begin() + best);
  } else {
    if (g_stage_pool.size() >= kPoolMax) {   // make room: drop the smallest
      size_t k = 0;
      for (size_t i = 1; i < g_stage_pool.size(); ++i) if (g_stage_pool[i].cap < g_stage_pool[k].cap) k = i;
      cudaFree(g_stage_pool[k].p);
      g_stage_pool.erase(g_stage_pool.begin() + k);
    }
    const size_t cap = std::max<size_t>(count, 1);
    cudaError_t e = cudaMalloc((void **)&stage, cap * sizeof(double));
    if (e != cudaSuccess) {
      stage = nullptr;
      return fail(-101, std::string("cudaMalloc failed (staging): ") + cudaGetErrorString(e));
    }
    stage_cap = cap;
  }
  dev = stage;
  if (copy_in && count) TFX_CUDA(cudaMemcpyAsync(dev, host, count * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
  return 0;
}
int VecIO::copy_back() {
  if (host && n) {
    TFX_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Matrix: device mirror
// ---------------------------------------------------------------------------------------------
template <typename T>
static int upload(DevBuf<T> &d, const std::vector<T> &h) {
  TFX_TRY(d.alloc(h.size()));
  if (!h.empty()) TFX_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx().stream));
  return 0;
}

int matrix_upload(Matrix &m, bool allow_dense) {
  TFX_TRY(ensure_init());
  const int32_t ns = m.nl_nonempty;
  const int64_t nel = m.nel;
  // ---- forward CSR (0-based)
  std::vector<int64_t> ptr((size_t)ns + 1);
  for (int32_t i = 0; i <= ns; ++i) ptr[i] = m.ijl[i] - 1;
  std::vector<int32_t> idx((size_t)nel), segmap((size_t)ns);
  for (int64_t k = 0; k < nel; ++k) idx[k] = m.ija[k] - 1;
  for (int32_t i = 0; i < ns; ++i) segmap[i] = m.rowptr[i] - 1;
  SegMatrix &f = m.fwd;
  f.nnz = nel; f.nseg = ns; f.nout = m.nl; f.nin = m.ncolumns;
  TFX_TRY(upload(f.ptr, ptr));
  TFX_TRY(upload(f.idx, idx));
  std::vector<float> val(m.sa.begin(), m.sa.begin() + nel);
  TFX_TRY(upload(f.val, val));
  TFX_TRY(upload(f.segmap, segmap));
  TFX_TRY(seg_build_items(f, ptr.data()));

  // ---- CSR of the transpose: counting sort by column, stable in stored-row order, so that the
  // entries of a column appear in the order add_trans_mult_vector visits them (sparse_matrix.f90:397-403).
  std::vector<int64_t> cnt((size_t)m.ncolumns + 1, 0);
  for (int64_t k = 0; k < nel; ++k) cnt[idx[k] + 1]++;
  std::vector<int32_t> tmap;
  std::vector<int64_t> tptr;
  std::vector<int64_t> start((size_t)m.ncolumns, 0);
  tptr.push_back(0);
  {
    int64_t run = 0;
    for (int32_t j = 0; j < m.ncolumns; ++j) {
      start[j] = run;
      if (cnt[j + 1] > 0) {
        tmap.push_back(j);
        run += cnt[j + 1];
        tptr.push_back(run);
      }
    }
  }
  std::vector<int32_t> tidx((size_t)nel);
  std::vector<float> tval((size_t)nel);
  for (int32_t i = 0; i < ns; ++i) {
    const int32_t row = segmap[i];
    for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) {
      const int64_t pos = start[idx[k]]++;
      tidx[pos] = row;
      tval[pos] = val[k];
    }
  }
  SegMatrix &t = m.trn;
  t.nnz = nel; t.nseg = (int32_t)tmap.size(); t.nout = m.ncolumns; t.nin = m.nl;
  TFX_TRY(upload(t.ptr, tptr));
  TFX_TRY(upload(t.idx, tidx));
  TFX_TRY(upload(t.val, tval));
  TFX_TRY(upload(t.segmap, tmap));
  TFX_TRY(seg_build_items(t, tptr.data()));
  m.has_seg = true;

  // ---- dense block detection: every row present, same contiguous column range (the uncompressed
  // kernel, sensitivity_gravmag.F90:288-295, possibly shifted by param_shift).
  m.has_dense = false;
  if (allow_dense && g_opt_dense_detect && ns == m.nl && ns >= 1 && ns <= kDenseMaxRows && nel > 0) {
    const int64_t W = ptr[1] - ptr[0];
    bool ok = W > 0 && W * (int64_t)ns == nel;
    const int32_t c0 = ok ? idx[0] : 0;
    for (int32_t i = 0; ok && i < ns; ++i) {
      if (ptr[i + 1] - ptr[i] != W) { ok = false; break; }
      for (int64_t k = 0; k < W; ++k)
        if (idx[ptr[i] + k] != c0 + (int32_t)k) { ok = false; break; }
    }
    if (ok) {
      DenseCM &d = m.dense;
      d.nrows = ns; d.ncols = (int32_t)W; d.col0 = c0; d.ld = ((int64_t)ns + 3) / 4 * 4; d.grid = 0;
      std::vector<float> cm((size_t)d.ld * (size_t)W, 0.0f);
      for (int32_t i = 0; i < ns; ++i)
        for (int64_t k = 0; k < W; ++k) cm[(size_t)k * d.ld + i] = val[ptr[i] + k];
      TFX_TRY(upload(d.val, cm));
      m.has_dense = true;
      m.dense_row0 = 0;
    }
  }
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  if (!m.has_dense) TFX_TRY(matrix_build_t16(m));
  return 0;
}

int matrix_build_t16(Matrix &m) {
  m.has_t16 = false;
  m.t16f.release(); m.t16t.release();
  if (!m.has_seg || m.fwd.nnz < (int64_t)g_opt_t16_min_nnz) return 0;
  cudaStream_t st = ctx().stream;
  TFX_TRY(t16_build(m.fwd, m.t16f, st));
  if (!m.t16f.valid) return 0;
  TFX_TRY(t16_build(m.trn, m.t16t, st));
  if (!m.t16t.valid) { m.t16f.release(); return 0; }
  m.t16f.nout_total = m.nl;
  m.t16t.nout_total = m.ncolumns;
  m.has_t16 = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Builder (mirrors sparse_matrix.f90 line by line in behaviour, including the error texts)
// ---------------------------------------------------------------------------------------------
int matrix_fwd(Matrix &m, const double *d_x, double *d_y, bool accumulate, int32_t xshift, const int *d_done, cudaStream_t st) {
  if (m.has_blocks) {
    Context &c = ctx();
    if (m.blocks.size() < 2 || st != c.stream) {
      for (size_t b = 0; b < m.blocks.size(); ++b)
        TFX_TRY(matrix_fwd(*m.blocks[b], d_x, d_y + m.block_row0[b], accumulate, xshift, d_done, st));
      return 0;
    }
    // the blocks write disjoint row ranges: alternate them over the two side streams (fork / join on the main one)
    TFX_CUDA(cudaEventRecord(c.ev_fork, st));
    for (int i = 0; i < 2; ++i) TFX_CUDA(cudaStreamWaitEvent(c.side[i], c.ev_fork, 0));
    for (size_t b = 0; b < m.blocks.size(); ++b)
      TFX_TRY(matrix_fwd(*m.blocks[b], d_x, d_y + m.block_row0[b], accumulate, xshift, d_done, c.side[b & 1]));
    for (int i = 0; i < 2; ++i) {
      TFX_CUDA(cudaEventRecord(c.ev_join[i], c.side[i]));
      TFX_CUDA(cudaStreamWaitEvent(st, c.ev_join[i], 0));
    }
    return 0;
  }
  if (m.has_t16) return t16_spmv(m.t16f, d_x, d_y, accumulate, xshift, d_done, st);
  if (m.has_seg) return seg_spmv(m.fwd, d_x, d_y, accumulate, 0, m.nl, xshift, d_done, st);
  if (m.has_dense) {
    // a dense row block (uncompressed kernels with more than kDenseMaxRows data rows are stored as several):
    // the forward-only mode of the sweep, x read at column - xshift
    if (!accumulate)
      return dense_sweep(m.dense, DENSE_F_ONLY, nullptr, d_x - xshift, nullptr, nullptr, nullptr, d_y + m.dense_row0, nullptr,
                         d_done, st);
    DevBuf<double> tmp;   // (add_mult_vector through the API: never inside the LSQR loop)
    TFX_TRY(tmp.alloc((size_t)m.dense.nrows));
    TFX_TRY(dense_sweep(m.dense, DENSE_F_ONLY, nullptr, d_x - xshift, nullptr, nullptr, nullptr, tmp.p, nullptr, d_done, st));
    TFX_TRY(vec_add_inplace(d_y + m.dense_row0, tmp.p, (size_t)m.dense.nrows, st));
    TFX_CUDA(cudaStreamSynchronize(st));
    return 0;
  }
  return fail(-21, "sparse_matrix: no compressed device representation");
}
int matrix_trans(Matrix &m, const double *d_u, double *d_y, bool accumulate, const int *d_done, cudaStream_t st) {
  if (m.has_blocks) {
    if (m.blocks.empty() && !accumulate) TFX_CUDA(cudaMemsetAsync(d_y, 0, (size_t)m.ncolumns * 8, st));
    for (size_t b = 0; b < m.blocks.size(); ++b)
      TFX_TRY(matrix_trans(*m.blocks[b], d_u + m.block_row0[b], d_y, accumulate || b > 0, d_done, st));
    return 0;
  }
  if (m.has_t16) return t16_spmv(m.t16t, d_u, d_y, accumulate, 0, d_done, st);
  if (m.has_seg) return seg_spmv(m.trn, d_u, d_y, accumulate, 0, m.ncolumns, 0, d_done, st);
  if (m.has_dense) {
    // the sweep writes the block's own columns only: the others are zero unless accumulating
    if (!accumulate) TFX_CUDA(cudaMemsetAsync(d_y, 0, (size_t)m.ncolumns * 8, st));
    return dense_sweep(m.dense, DENSE_T_ONLY, d_u + m.dense_row0, nullptr, nullptr, d_y, nullptr, nullptr, nullptr, d_done, st,
                       true);
  }
  return fail(-21, "sparse_matrix: no compressed device representation");
}

static int builder_guard(const Matrix &m) {
  if (m.device_only) return fail(-10, "sparse_matrix: this matrix was assembled on the device and cannot be modified");
  return 0;
}

}  // namespace tfx

using namespace tfx;

extern "C" {

int tfx_version(void) { return 100; }
const char *tfx_last_error(void) { return g_err.c_str(); }
int tfx_init(int device) { return init_device(device); }
int tfx_finalize(void) {
  comm_finalize();
  return 0;
}
int tfx_device_synchronize(void) {
  TFX_TRY(ensure_init());
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}
uint64_t tfx_launch_count(void) { return ctx().launches; }

// CUDA-event stopwatch on the library stream (bench.py times kernels with it).
static cudaEvent_t g_timer0 = nullptr, g_timer1 = nullptr;
int tfx_timer_start(void) {
  TFX_TRY(ensure_init());
  if (!g_timer0) { TFX_CUDA(cudaEventCreate(&g_timer0)); TFX_CUDA(cudaEventCreate(&g_timer1)); }
  TFX_CUDA(cudaEventRecord(g_timer0, ctx().stream));
  return 0;
}
int tfx_timer_stop(double *ms) {
  if (!g_timer0) return fail(-5, "tfx_timer_stop without tfx_timer_start");
  TFX_CUDA(cudaEventRecord(g_timer1, ctx().stream));
  TFX_CUDA(cudaEventSynchronize(g_timer1));
  float t = 0.f;
  TFX_CUDA(cudaEventElapsedTime(&t, g_timer0, g_timer1));
  if (ms) *ms = t;
  return 0;
}
// Bytes of device memory held by the matrix representations (values, indices, pointer tables).
int64_t tfx_sparse_matrix_device_bytes(const tfx_matrix *h) {
  const Matrix &m = h->m;
  int64_t b = 0;
  if (m.has_blocks) {
    for (const Matrix *blk : m.blocks) {
      if (blk->has_seg) b += (blk->fwd.nnz + blk->trn.nnz) * 8 + ((int64_t)blk->fwd.nseg + blk->trn.nseg) * 12;
      if (blk->has_t16) b += blk->t16f.bytes() + blk->t16t.bytes();
      if (blk->has_dense) b += (int64_t)blk->dense.ld * blk->dense.ncols * 4;
    }
    return b;
  }
  if (m.has_seg) b += (m.fwd.nnz + m.trn.nnz) * 8 + ((int64_t)m.fwd.nseg + m.trn.nseg) * 12;
  if (m.has_t16) b += m.t16f.bytes() + m.t16t.bytes();
  if (m.has_dense) b += (int64_t)m.dense.ld * m.dense.ncols * 4;
  return b;
}
// Drops the generic CSR copies of a matrix that has the T16 layouts (frees 16 B/nnz); export() and the
// strict_order mode are no longer available for it.
int tfx_sparse_matrix_drop_csr(tfx_matrix *h) {
  Matrix &m = h->m;
  if (!m.has_t16) return fail(-24, "drop_csr: the matrix has no T16 layouts");
  m.fwd.release(); m.trn.release();
  m.has_seg = false;
  return 0;
}
int tfx_set_option(const char *name, int value) {
  if (name && strcmp(name, "dense_detect") == 0) {
    g_opt_dense_detect = value;
    return 0;
  }
  if (name && strcmp(name, "profile_sweeps") == 0) {
    g_opt_profile_sweeps = value;
    return 0;
  }
  if (name && strcmp(name, "lsqr_graph") == 0) {
    g_opt_lsqr_graph = value;
    return 0;
  }
  if (name && strcmp(name, "t16_tma") == 0) {
    g_opt_t16_tma = value;
    return 0;
  }
  if (name && strcmp(name, "t16_blk") == 0) {
    g_opt_t16_blk = value;
    return 0;
  }
  if (name && strcmp(name, "t16_long_seg") == 0) {
    g_opt_t16_long_seg = value;
    return 0;
  }
  if (name && strcmp(name, "wavelet_dist") == 0) {
    g_opt_wavelet_dist = value;
    return 0;
  }
  if (name && strcmp(name, "wavelet_tile_kb") == 0) {
    g_opt_wavelet_tile_kb = value;
    return 0;
  }
  if (name && strcmp(name, "wavelet_cols") == 0) {
    g_opt_wavelet_cols = value;
    return 0;
  }
  if (name && strcmp(name, "wavelet_p2p") == 0) {
    g_opt_wavelet_p2p = value;
    return 0;
  }
  if (name && strcmp(name, "wavelet_fuse12") == 0) {
    g_opt_wavelet_fuse12 = value;
    return 0;
  }
  if (name && strcmp(name, "wavelet_slab_mb") == 0) {
    g_opt_wavelet_slab_mb = value;
    return 0;
  }
  if (name && strcmp(name, "lsqr_poll") == 0) {
    g_opt_lsqr_poll = value;
    return 0;
  }
  if (name && strcmp(name, "strict_order") == 0) {
    g_opt_strict_order = value;
    return 0;
  }
  if (name && strcmp(name, "dense_vec4") == 0) {
    g_opt_dense_vec4 = value;
    return 0;
  }
  if (name && strcmp(name, "dense_stream_only") == 0) {
    g_opt_dense_stream_only = value;
    return 0;
  }
  if (name && strcmp(name, "dense_f2f_rows") == 0) {
    g_opt_dense_f2f_rows = value;
    return 0;
  }
  if (name && strcmp(name, "mag_shared_nodes") == 0) {
    g_opt_mag_shared_nodes = value;
    return 0;
  }
  if (name && strcmp(name, "grav_shared_nodes") == 0) {
    g_opt_grav_shared_nodes = value;
    return 0;
  }
  if (name && strcmp(name, "sensit_row_blocks") == 0) {
    g_opt_sensit_row_blocks = value;
    return 0;
  }
  if (name && strcmp(name, "dense_block_rows") == 0) {
    if (value < 0 || value > kDenseMaxRows) return fail(-4, "dense_block_rows must be in [0, kDenseMaxRows] (0: kDenseMaxRows)");
    g_opt_dense_block_rows = value;
    return 0;
  }
  if (name && strcmp(name, "trace") == 0) {
    g_opt_trace = value;
    return 0;
  }
  if (name && strcmp(name, "sensit_cand_cap") == 0) {
    g_opt_sensit_cand_cap = value;
    return 0;
  }
  if (name && strcmp(name, "t16_bank_deal") == 0) {
    g_opt_t16_bank_deal = value;
    return 0;
  }
  if (name && strcmp(name, "t16_direct_max") == 0) {
    g_opt_t16_direct_max = value;
    return 0;
  }
  if (name && strcmp(name, "t16_async") == 0) {
    g_opt_t16_async = value;
    return 0;
  }
  if (name && strcmp(name, "t16_min_nnz") == 0) {
    g_opt_t16_min_nnz = value;
    return 0;
  }
  if (name && strcmp(name, "t16_tile") == 0) {
    if (value != 0 && (value < 2 || value > 16384 || (value & (value - 1)) != 0))
      return fail(-4, "t16_tile must be 0 (automatic) or a power of two in [2, 16384]");
    g_opt_t16_tile = value;
    return 0;
  }
  return fail(-4, std::string("unknown option: ") + (name ? name : "(null)"));
}

// ---- communicator -------------------------------------------------------------------------------
int tfx_comm_unique_id(char id[128]) { return comm_unique_id(id); }
int tfx_comm_init(int nranks, int rank, const char id[128]) { return comm_init(nranks, rank, id); }
int tfx_comm_finalize(void) {
  wavelet_peer_reset();
  return comm_finalize();
}
int tfx_comm_allreduce_sum(double *buf, int64_t count) {
  TFX_TRY(ensure_init());
  VecIO v;
  TFX_TRY(v.bind(buf, (size_t)count, true));
  TFX_TRY(comm_allreduce_sum(v.dev, (size_t)count, ctx().stream));
  TFX_TRY(v.copy_back());
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

// ---- sparse_matrix ------------------------------------------------------------------------------
int tfx_sparse_matrix_initialize(tfx_matrix **out, int32_t nl, int32_t ncolumns, int64_t nnz, int32_t myrank,
                                 int32_t nl_empty) {
  (void)myrank;
  if (!out) return fail(-11, "sparse_matrix_initialize: null handle");
  if (nnz < 0 || nl < 0) return fail(-12, "Wrong sizes in sparse_matrix_allocate_arrays!");
  tfx_matrix *h = new tfx_matrix();
  Matrix &m = h->m;
  m.nl = nl; m.ncolumns = ncolumns; m.nnz = nnz;
  m.nl_nonempty_allocated = nl - nl_empty;
  if (m.nl_nonempty_allocated < 0) m.nl_nonempty_allocated = 0;
  m.ijl.assign((size_t)m.nl_nonempty_allocated + 1, 0);
  m.rowptr.assign((size_t)std::max(1, m.nl_nonempty_allocated), 0);
  // sa/ija grow on demand up to nnz (the reference allocates nnz up front; the bound is enforced in add()).
  *out = h;
  return 0;
}

int tfx_sparse_matrix_destroy(tfx_matrix *h) {
  delete h;
  return 0;
}

int tfx_sparse_matrix_reset(tfx_matrix *h) {
  Matrix &m = h->m;
  if (m.device_only) {
    // matrix_cons is reset and rebuilt before every solve (joint_inverse_problem.F90:364-373), also when its rows
    // were produced on the device: drop the device representations and go back to an empty builder
    m.fwd.release(); m.trn.release();
    m.dense.val.release(); m.dense.partial_q.release(); m.dense.partial_n2.release();
    m.device_only = false;
    m.nel = 0;
    m.ijl.assign((size_t)m.nl_nonempty_allocated + 1, 0);
    m.rowptr.assign((size_t)std::max(1, m.nl_nonempty_allocated), 0);
  }
  m.pend.idx.release(); m.pend.val.release(); m.pend.rowid.release(); m.pend.nnz = 0;
  m.clear_blocks();
  m.nl_current = 0; m.nl_current_all = 0; m.nel = 0; m.nel_last = 0; m.nl_nonempty = 0;
  m.sa.clear(); m.ija.clear();
  std::fill(m.ijl.begin(), m.ijl.end(), 0);
  std::fill(m.rowptr.begin(), m.rowptr.end(), 0);
  m.finalized = false; m.has_seg = false; m.has_dense = false; m.has_t16 = false;
  m.t16f.release(); m.t16t.release();
  return 0;
}

int tfx_sparse_matrix_add(tfx_matrix *h, double value, int32_t column, int32_t myrank) {
  (void)myrank;
  Matrix &m = h->m;
  TFX_TRY(builder_guard(m));
  if (value == 0.0) return 0;                      // "Do not add zero values to a sparse matrix."
  if (m.nel >= m.nnz) return fail(-13, "Error in total number of elements in sparse_matrix_add!");
  m.sa.push_back((float)value);                    // real(value, MATRIX_PRECISION)
  m.ija.push_back(column);
  m.nel += 1;
  return 0;
}

int tfx_sparse_matrix_add_row(tfx_matrix *h, int32_t nel_add, const float *values, const int32_t *columns,
                              int32_t myrank) {
  (void)myrank;
  Matrix &m = h->m;
  TFX_TRY(builder_guard(m));
  if (m.nel + nel_add > m.nnz) return fail(-14, "Error in total number of elements in sparse_matrix_add_row!");
  m.sa.insert(m.sa.end(), values, values + nel_add);
  m.ija.insert(m.ija.end(), columns, columns + nel_add);
  m.nel += nel_add;
  return 0;
}

int tfx_sparse_matrix_new_row(tfx_matrix *h, int32_t myrank) {
  (void)myrank;
  Matrix &m = h->m;
  TFX_TRY(builder_guard(m));
  if (m.nl_current >= m.nl_nonempty_allocated)
    return fail(-15, "Error in number of rows in sparse_matrix_new_row!\nnl_current=" + std::to_string(m.nl_current) +
                         "\nnl=" + std::to_string(m.nl));
  m.nl_current_all += 1;
  if (m.nel > m.nel_last) {                        // only non-empty rows are stored
    m.nl_current += 1;
    m.ijl[m.nl_current - 1] = m.nel_last + 1;
    m.rowptr[m.nl_current - 1] = m.nl_current_all;
    m.nel_last = m.nel;
  }
  return 0;
}

int tfx_sparse_matrix_add_empty_rows(tfx_matrix *h, int32_t nrows, int32_t myrank) {
  (void)myrank;
  Matrix &m = h->m;
  TFX_TRY(builder_guard(m));
  m.nl_current_all += nrows;
  return 0;
}

int tfx_sparse_matrix_finalize(tfx_matrix *h, int32_t myrank) {
  (void)myrank;
  Matrix &m = h->m;
  if (m.device_only && m.finalized) return 0;
  if (m.has_blocks) {
    if (m.nl_current_all != m.nl)
      return fail(-16, "Error in total number of rows in sparse_matrix_finalize!\nnl_current=" +
                           std::to_string(m.nl_current_all) + "\nnl=" + std::to_string(m.nl));
    if (m.nel != 0 || m.pend.nnz > 0) return fail(-25, "sparse_matrix_finalize: row blocks cannot be mixed with other rows");
    int64_t nel = 0;
    for (const Matrix *b : m.blocks) nel += b->nel;
    m.nel = nel;
    m.device_only = true;
    m.finalized = true;
    return 0;
  }
  if (m.pend.nnz > 0 || m.pend.idx.p) {
    // rows appended on the device (read_sensitivity_kernel / re-partitioner): same row-count check as the reference
    if (m.nl_current_all != m.nl)
      return fail(-16, "Error in total number of rows in sparse_matrix_finalize!\nnl_current=" +
                           std::to_string(m.nl_current) + "\nnl=" + std::to_string(m.nl));
    TFX_TRY(matrix_flush_host_rows(m));   // host-built rows that follow the device-appended ones
    RowTriplets R;
    std::swap(R.idx.p, m.pend.idx.p); std::swap(R.idx.n, m.pend.idx.n);
    std::swap(R.val.p, m.pend.val.p); std::swap(R.val.n, m.pend.val.n);
    std::swap(R.rowid.p, m.pend.rowid.p); std::swap(R.rowid.n, m.pend.rowid.n);
    R.nnz = m.pend.nnz; m.pend.nnz = 0;
    return matrix_from_triplets(m, m.nl, m.ncolumns, R);
  }
  if (m.nl_current_all != m.nl)
    return fail(-16, "Error in total number of rows in sparse_matrix_finalize!\nnl_current=" +
                         std::to_string(m.nl_current) + "\nnl=" + std::to_string(m.nl));
  if (m.nel_last != m.nel)
    return fail(-17, "Elements were added to the matrix after calling new_row() and before calling finalize()!");
  m.ijl[m.nl_current] = m.nel + 1;
  m.nl_nonempty = m.nl_current;
  // validate(), sparse_matrix.f90:188-208
  for (int32_t i = 0; i < m.nl_nonempty; ++i)
    for (int64_t k = m.ijl[i]; k <= m.ijl[i + 1] - 1; ++k) {
      if (k < 1 || k > m.nnz) return fail(-18, "Sparse matrix element-index validation failed!");
      const int32_t j = m.ija[k - 1];
      if (j < 1 || j > m.ncolumns) return fail(-19, "Sparse matrix column-index validation failed!");
    }
  TFX_TRY(matrix_upload(m, true));
  m.finalized = true;
  return 0;
}

int tfx_sparse_matrix_from_arrays(tfx_matrix **out, int32_t nl, int32_t ncolumns, int32_t nl_nonempty, int64_t nel,
                                  const float *sa, const int32_t *ija, const int64_t *ijl, const int32_t *rowptr) {
  TFX_TRY(tfx_sparse_matrix_initialize(out, nl, ncolumns, nel, 0, nl - nl_nonempty));
  Matrix &m = (*out)->m;
  m.sa.assign(sa, sa + nel);
  m.ija.assign(ija, ija + nel);
  for (int32_t i = 0; i <= nl_nonempty; ++i) m.ijl[i] = ijl[i];
  for (int32_t i = 0; i < nl_nonempty; ++i) m.rowptr[i] = rowptr[i];
  m.nel = m.nel_last = nel;
  m.nl_current = nl_nonempty;
  m.nl_current_all = nl;
  int rc = tfx_sparse_matrix_finalize(*out, 0);
  if (rc != 0) {
    delete *out;
    *out = nullptr;
  }
  return rc;
}

// normalize_columns (sparse_matrix.f90:414-443): column norms from the stored real(4) values squared in
// real(4) (`this%sa(k)**2`), accumulated in real(8); values divided in real(8) and rounded back to real(4).
// Host-side builder operation (the reference only calls it from its unit test); a finalized matrix is
// re-mirrored to the device afterwards.
int tfx_sparse_matrix_normalize_columns(tfx_matrix *h, double *column_norm) {
  Matrix &m = h->m;
  TFX_TRY(builder_guard(m));
  if (!column_norm) return fail(-22, "normalize_columns: null column_norm");
  for (int32_t j = 0; j < m.ncolumns; ++j) column_norm[j] = 0.0;
  const int32_t ns = m.finalized ? m.nl_nonempty : m.nl_current;
  const int64_t nel = m.finalized ? m.nel : m.nel_last;
  (void)ns;
  for (int64_t k = 0; k < nel; ++k) {
    const float sq = m.sa[k] * m.sa[k];
    column_norm[m.ija[k] - 1] += (double)sq;
  }
  for (int32_t j = 0; j < m.ncolumns; ++j) column_norm[j] = sqrt(column_norm[j]);
  for (int64_t k = 0; k < nel; ++k) {
    const double cn = column_norm[m.ija[k] - 1];
    if (cn != 0.0) m.sa[k] = (float)((double)m.sa[k] / cn);
  }
  if (m.finalized) TFX_TRY(matrix_upload(m, true));
  return 0;
}

int32_t tfx_sparse_matrix_get_total_row_number(const tfx_matrix *h) { return h->m.nl; }
int32_t tfx_sparse_matrix_get_current_row_number(const tfx_matrix *h) { return h->m.nl_current_all; }
int32_t tfx_sparse_matrix_get_ncolumns(const tfx_matrix *h) { return h->m.ncolumns; }
int64_t tfx_sparse_matrix_get_number_elements(const tfx_matrix *h) { return h->m.nel; }
int64_t tfx_sparse_matrix_get_nnz(const tfx_matrix *h) { return h->m.nnz; }
int tfx_sparse_matrix_storage_kind(const tfx_matrix *h) {
  const Matrix &m = h->m;
  if (m.has_blocks) {
    bool all_dense = !m.blocks.empty();
    for (const Matrix *b : m.blocks) all_dense = all_dense && b->has_dense;
    if (all_dense) return 1;
    for (const Matrix *b : m.blocks) if (!b->has_t16) return 0;
    return m.blocks.empty() ? 0 : 2;
  }
  return m.has_dense ? 1 : (m.has_t16 ? 2 : 0);
}

// Products. kind: 0 forward, 1 transposed.
static int product(Matrix &m, const double *x, double *b, bool accumulate, bool transposed) {
  TFX_TRY(ensure_init());
  if (!m.finalized) return fail(-20, "sparse_matrix: product called before finalize()");
  cudaStream_t st = ctx().stream;
  const size_t nin = transposed ? m.nl : m.ncolumns, nout = transposed ? m.ncolumns : m.nl;
  VecIO vx, vb;
  TFX_TRY(vx.bind(const_cast<double *>(x), nin, true));
  TFX_TRY(vb.bind(b, nout, accumulate));
  if (m.has_blocks) {
    if (transposed) TFX_TRY(matrix_trans(m, vx.dev, vb.dev, accumulate, nullptr, st));
    else TFX_TRY(matrix_fwd(m, vx.dev, vb.dev, accumulate, 0, nullptr, st));
  } else if (m.has_t16) {
    TFX_TRY(t16_spmv(transposed ? m.t16t : m.t16f, vx.dev, vb.dev, accumulate, 0, nullptr, st));
  } else if (m.has_seg) {
    SegMatrix &s = transposed ? m.trn : m.fwd;
    TFX_TRY(seg_spmv(s, vx.dev, vb.dev, accumulate, 0, (int32_t)nout, 0, nullptr, st));
  } else if (m.has_dense) {
    // device-assembled dense block: run the one-product modes of the sweep kernel
    DevBuf<double> tmp;
    if (transposed) {
      TFX_TRY(tmp.alloc(nout));
      TFX_CUDA(cudaMemsetAsync(tmp.p, 0, nout * 8, st));
      TFX_TRY(dense_sweep(m.dense, DENSE_T_ONLY, vx.dev + m.dense_row0, nullptr, nullptr, tmp.p, nullptr, nullptr, nullptr, nullptr, st));
    } else {
      TFX_TRY(tmp.alloc(nout));
      TFX_CUDA(cudaMemsetAsync(tmp.p, 0, nout * 8, st));
      TFX_TRY(dense_sweep(m.dense, DENSE_F_ONLY, nullptr, vx.dev, nullptr, nullptr, nullptr, tmp.p + m.dense_row0, nullptr, nullptr, st));
    }
    // b = (accumulate ? b : 0) + tmp  -- tiny axpy through thrust-free path: reuse cudaMemcpy when !accumulate
    if (!accumulate) {
      TFX_CUDA(cudaMemcpyAsync(vb.dev, tmp.p, nout * 8, cudaMemcpyDeviceToDevice, st));
    } else {
      TFX_TRY(vec_add_inplace(vb.dev, tmp.p, nout, st));
    }
    TFX_CUDA(cudaStreamSynchronize(st));
  } else {
    return fail(-21, "sparse_matrix: no device representation");
  }
  TFX_TRY(vb.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// Times `reps` back-to-back products (device-resident vectors required) with CUDA events on the library
// stream, without host synchronisation between launches; returns the mean milliseconds per product.
int tfx_sparse_matrix_time_product(tfx_matrix *h, int transposed, const double *x, double *b, int reps, double *ms) {
  Matrix &m = h->m;
  TFX_TRY(ensure_init());
  if (!m.finalized) return fail(-20, "sparse_matrix: product called before finalize()");
  if (!is_device_ptr(x) || !is_device_ptr(b)) return fail(-25, "time_product: vectors must be device pointers");
  if (reps < 1) reps = 1;
  cudaStream_t st = ctx().stream;
  cudaEvent_t e0, e1;
  TFX_CUDA(cudaEventCreate(&e0)); TFX_CUDA(cudaEventCreate(&e1));
  const int32_t nout = transposed ? m.ncolumns : m.nl;
  auto once = [&]() -> int {
    if (m.has_blocks) return transposed ? matrix_trans(m, x, b, false, nullptr, st) : matrix_fwd(m, x, b, false, 0, nullptr, st);
    if (m.has_t16) return t16_spmv(transposed ? m.t16t : m.t16f, x, b, false, 0, nullptr, st);
    if (m.has_seg) return seg_spmv(transposed ? m.trn : m.fwd, x, b, false, 0, nout, 0, nullptr, st);
    return fail(-21, "time_product: needs a compressed representation");
  };
  TFX_TRY(once());
  TFX_CUDA(cudaEventRecord(e0, st));
  for (int r = 0; r < reps; ++r) TFX_TRY(once());
  TFX_CUDA(cudaEventRecord(e1, st));
  TFX_CUDA(cudaEventSynchronize(e1));
  float t = 0.f;
  TFX_CUDA(cudaEventElapsedTime(&t, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (ms) *ms = (double)t / reps;
  return 0;
}

int tfx_sparse_matrix_mult_vector(tfx_matrix *h, const double *x, double *b) { return product(h->m, x, b, false, false); }
int tfx_sparse_matrix_add_mult_vector(tfx_matrix *h, const double *x, double *b) { return product(h->m, x, b, true, false); }
int tfx_sparse_matrix_trans_mult_vector(tfx_matrix *h, const double *x, double *b) { return product(h->m, x, b, false, true); }
int tfx_sparse_matrix_add_trans_mult_vector(tfx_matrix *h, const double *x, double *b) { return product(h->m, x, b, true, true); }

// `pure` Fortran callers (see include/tfx.h): the code of the first failed product is kept with its message until a
// non-pure call collects it.
static int g_latched_rc = 0;
static std::string g_latched_msg;
static void latch(int rc) {
  if (rc != 0 && g_latched_rc == 0) {
    g_latched_rc = rc;
    g_latched_msg = g_err;
  }
}
void tfx_sparse_matrix_mult_vector_v(tfx_matrix *h, const double *x, double *b) { latch(product(h->m, x, b, false, false)); }
void tfx_sparse_matrix_add_mult_vector_v(tfx_matrix *h, const double *x, double *b) { latch(product(h->m, x, b, true, false)); }
void tfx_sparse_matrix_trans_mult_vector_v(tfx_matrix *h, const double *x, double *b) { latch(product(h->m, x, b, false, true)); }
void tfx_sparse_matrix_add_trans_mult_vector_v(tfx_matrix *h, const double *x, double *b) { latch(product(h->m, x, b, true, true)); }
int tfx_take_latched_error(void) {
  const int rc = g_latched_rc;
  if (rc != 0) {
    g_err = g_latched_msg;
    g_latched_rc = 0;
    g_latched_msg.clear();
  }
  return rc;
}

int tfx_sparse_matrix_part_mult_vector(tfx_matrix *h, int32_t nelements, const double *x, int32_t ndata, double *b,
                                       int32_t line_start, int32_t param_shift, int32_t myrank) {
  (void)myrank;
  Matrix &m = h->m;
  TFX_TRY(ensure_init());
  if (!m.finalized) return fail(-20, "sparse_matrix: product called before finalize()");
  const int32_t line_end = line_start + ndata - 1;
  if (line_start < 1 || line_start > m.nl_current_all || line_end < 1 || line_end > m.nl_current_all)
    return fail(-22, "Wrong line index in sparse_matrix_part_mult_vector!");
  cudaStream_t st = ctx().stream;
  VecIO vx, vb;
  TFX_TRY(vx.bind(const_cast<double *>(x), (size_t)nelements, true));
  TFX_TRY(vb.bind(b, (size_t)ndata, false));
  if (m.has_blocks) {
    // Only the row blocks that intersect the requested window are multiplied: in a joint matrix the blocks of the other
    // problem hold columns outside [param_shift, param_shift + nelements) and must not be read through x.
    DevBuf<double> full;
    TFX_CUDA(cudaMemsetAsync(vb.dev, 0, (size_t)ndata * 8, st));
    for (size_t bi = 0; bi < m.blocks.size(); ++bi) {
      Matrix &B = *m.blocks[bi];
      const int32_t r0 = m.block_row0[bi];
      const int32_t lo = std::max(r0, line_start - 1), hi = std::min(r0 + B.nl, line_end);
      if (lo >= hi) continue;
      double *dst = vb.dev + (lo - (line_start - 1));
      if (B.has_t16) {
        if (B.t16f.in0 < param_shift || (int64_t)B.t16f.in0 - param_shift + B.t16f.nin > nelements)
          return fail(-24, "part_mult_vector: a row block inside the requested lines holds columns outside "
                           "[param_shift, param_shift + nelements)");
        TFX_TRY(full.alloc((size_t)B.nl));
        TFX_TRY(t16_spmv(B.t16f, vx.dev, full.p, false, param_shift, nullptr, st));
        TFX_CUDA(cudaMemcpyAsync(dst, full.p + (lo - r0), (size_t)(hi - lo) * 8, cudaMemcpyDeviceToDevice, st));
      } else if (B.has_seg) {
        TFX_TRY(seg_spmv(B.fwd, vx.dev, dst, false, lo - r0, hi - r0, param_shift, nullptr, st));
      } else if (B.has_dense) {
        if (B.dense.col0 < param_shift || (int64_t)B.dense.col0 - param_shift + B.dense.ncols > nelements)
          return fail(-24, "part_mult_vector: a row block inside the requested lines holds columns outside "
                           "[param_shift, param_shift + nelements)");
        TFX_TRY(full.alloc((size_t)B.nl));
        TFX_TRY(dense_sweep(B.dense, DENSE_F_ONLY, nullptr, vx.dev - param_shift, nullptr, nullptr, nullptr, full.p, nullptr,
                            nullptr, st));
        TFX_CUDA(cudaMemcpyAsync(dst, full.p + (lo - r0), (size_t)(hi - lo) * 8, cudaMemcpyDeviceToDevice, st));
      } else {
        return fail(-21, "sparse_matrix: row block without a device representation");
      }
    }
    TFX_CUDA(cudaStreamSynchronize(st));
  } else if (m.has_t16 && m.t16f.in0 >= param_shift && m.t16f.in0 - param_shift + m.t16f.nin <= nelements) {
    // all rows through the F layout (x read at column - param_shift), then the requested window
    DevBuf<double> full;
    TFX_TRY(full.alloc((size_t)m.nl));
    TFX_TRY(t16_spmv(m.t16f, vx.dev, full.p, false, param_shift, nullptr, st));
    TFX_CUDA(cudaMemcpyAsync(vb.dev, full.p + (line_start - 1), (size_t)ndata * 8, cudaMemcpyDeviceToDevice, st));
    TFX_CUDA(cudaStreamSynchronize(st));
  } else if (m.has_seg) {
    TFX_TRY(seg_spmv(m.fwd, vx.dev, vb.dev, false, line_start - 1, line_end, param_shift, nullptr, st));
  } else if (m.has_dense) {
    // dense block covers rows [dense_row0, dense_row0 + nrows) and columns [col0, col0 + ncols):
    // x(ija - param_shift) -> element col0 + c - param_shift of x.
    if (line_start - 1 != m.dense_row0 || ndata != m.dense.nrows || m.dense.col0 < param_shift ||
        m.dense.col0 - param_shift + m.dense.ncols > nelements)
      return fail(-23, "part_mult_vector: requested part does not match the device-resident dense block");
    DenseCM &d = m.dense;
    const int32_t saved = d.col0;
    d.col0 = saved - param_shift;
    int rc = dense_sweep(d, DENSE_F_ONLY, nullptr, vx.dev, nullptr, nullptr, nullptr, vb.dev, nullptr, nullptr, st);
    d.col0 = saved;
    TFX_TRY(rc);
  } else {
    return fail(-21, "sparse_matrix: no device representation");
  }
  TFX_TRY(vb.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int tfx_sparse_matrix_export(tfx_matrix *h, int64_t *nel, int32_t *nl_nonempty, float *sa, int32_t *ija, int64_t *ijl,
                             int32_t *rowptr) {
  Matrix &m = h->m;
  if (!m.device_only) {
    if (nel) *nel = m.nel;
    if (nl_nonempty) *nl_nonempty = m.nl_nonempty;
    if (sa) memcpy(sa, m.sa.data(), (size_t)m.nel * 4);
    if (ija) memcpy(ija, m.ija.data(), (size_t)m.nel * 4);
    if (ijl) memcpy(ijl, m.ijl.data(), ((size_t)m.nl_nonempty + 1) * 8);
    if (rowptr) memcpy(rowptr, m.rowptr.data(), (size_t)m.nl_nonempty * 4);
    return 0;
  }
  TFX_TRY(ensure_init());
  if (m.has_blocks) return fail(-29, "sparse_matrix_export: a row-blocked matrix keeps only its product layouts");
  if (m.has_seg) {
    const SegMatrix &f = m.fwd;
    if (nel) *nel = f.nnz;
    if (nl_nonempty) *nl_nonempty = f.nseg;
    if (sa) TFX_CUDA(cudaMemcpy(sa, f.val.p, (size_t)f.nnz * 4, cudaMemcpyDeviceToHost));
    if (ija) {
      TFX_CUDA(cudaMemcpy(ija, f.idx.p, (size_t)f.nnz * 4, cudaMemcpyDeviceToHost));
      for (int64_t k = 0; k < f.nnz; ++k) ija[k] += 1;
    }
    if (ijl) {
      TFX_CUDA(cudaMemcpy(ijl, f.ptr.p, ((size_t)f.nseg + 1) * 8, cudaMemcpyDeviceToHost));
      for (int32_t i = 0; i <= f.nseg; ++i) ijl[i] += 1;
    }
    if (rowptr) {
      TFX_CUDA(cudaMemcpy(rowptr, f.segmap.p, (size_t)f.nseg * 4, cudaMemcpyDeviceToHost));
      for (int32_t i = 0; i < f.nseg; ++i) rowptr[i] += 1;
    }
    return 0;
  }
  if (m.has_dense) {
    const DenseCM &d = m.dense;
    const int64_t n = (int64_t)d.nrows * d.ncols;
    if (nel) *nel = n;
    if (nl_nonempty) *nl_nonempty = d.nrows;
    if (sa || ija) {
      std::vector<float> cm((size_t)d.ld * d.ncols);
      TFX_CUDA(cudaMemcpy(cm.data(), d.val.p, cm.size() * 4, cudaMemcpyDeviceToHost));
      for (int32_t i = 0; i < d.nrows; ++i)
        for (int32_t c = 0; c < d.ncols; ++c) {
          if (sa) sa[(int64_t)i * d.ncols + c] = cm[(size_t)c * d.ld + i];
          if (ija) ija[(int64_t)i * d.ncols + c] = d.col0 + c + 1;
        }
    }
    if (ijl) for (int32_t i = 0; i <= d.nrows; ++i) ijl[i] = (int64_t)i * d.ncols + 1;
    if (rowptr) for (int32_t i = 0; i < d.nrows; ++i) rowptr[i] = m.dense_row0 + i + 1;
    return 0;
  }
  return fail(-21, "sparse_matrix: no device representation");
}

// ---- wavelet_transform --------------------------------------------------------------------------
static int wavelet_call(double *s, int32_t n1, int32_t n2, int32_t n3, int32_t type, bool fwd) {
  TFX_TRY(ensure_init());
  VecIO v;
  TFX_TRY(v.bind(s, (size_t)n1 * n2 * n3, true));
  TFX_TRY(wavelet3d_device(v.dev, n1, n2, n3, type, fwd, ctx().stream));
  TFX_TRY(v.copy_back());
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}
int tfx_forward_wavelet(double *s, int32_t n1, int32_t n2, int32_t n3, int32_t t) { return wavelet_call(s, n1, n2, n3, t, true); }
int tfx_inverse_wavelet(double *s, int32_t n1, int32_t n2, int32_t n3, int32_t t) { return wavelet_call(s, n1, n2, n3, t, false); }
int tfx_Haar3D(double *s, int32_t n1, int32_t n2, int32_t n3) { return wavelet_call(s, n1, n2, n3, 1, true); }
int tfx_iHaar3D(double *s, int32_t n1, int32_t n2, int32_t n3) { return wavelet_call(s, n1, n2, n3, 1, false); }
int tfx_DaubD43D(double *s, int32_t n1, int32_t n2, int32_t n3) { return wavelet_call(s, n1, n2, n3, 2, true); }
int tfx_iDaubD43D(double *s, int32_t n1, int32_t n2, int32_t n3) { return wavelet_call(s, n1, n2, n3, 2, false); }

// ---- lsqr_solver --------------------------------------------------------------------------------
static LsqrResult g_last;

static int lsqr_call(const LsqrParams &p0, tfx_matrix *S, tfx_matrix *C, double *u, double *x) {
  TFX_TRY(ensure_init());
  // Host vectors are staged in pooled device buffers; the solver itself copies what it needs (the data rows and this
  // rank's constraint rows of u, the active problems' columns of x) in and out -- not the whole vectors.
  VecIO vu, vx;
  TFX_TRY(vu.bind(u, (size_t)p0.nlines, false));
  TFX_TRY(vx.bind(x, (size_t)p0.ncolumns, false));
  LsqrParams p = p0;
  p.host_u = vu.host;
  p.host_x = vx.host;
  TFX_TRY(lsqr_run(p, &S->m, C ? &C->m : nullptr, vu.dev, vx.dev, g_last));
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

int tfx_lsqr_solve(int32_t nlines, int32_t nelements, int32_t niter, double rmin, double gamma, tfx_matrix *matrix,
                   double *u, double *x, int32_t myrank) {
  LsqrParams p;
  p.nlines = nlines; p.ncolumns = nelements; p.niter = niter; p.rmin = rmin; p.gamma = gamma;
  p.single_matrix = true; p.myrank = myrank; p.nelements = nelements;
  return lsqr_call(p, matrix, nullptr, u, x);
}

int tfx_lsqr_solve_sensit(int32_t nlines, int32_t ncolumns, int32_t niter, double rmin, double gamma,
                          double target_misfit, tfx_matrix *matrix_sensit, tfx_matrix *matrix_cons, double *u,
                          double *x, const int32_t solve_problem[2], int32_t nelements, int32_t nx, int32_t ny,
                          int32_t nz, int32_t ncomponents, int32_t compression_type, int32_t wavelet_domain,
                          double *memory, int32_t myrank, int32_t nbproc) {
  LsqrParams p;
  p.nlines = nlines; p.ncolumns = ncolumns; p.niter = niter; p.rmin = rmin; p.gamma = gamma;
  p.target_misfit = target_misfit;
  p.solve_problem[0] = solve_problem[0]; p.solve_problem[1] = solve_problem[1];
  p.nelements = nelements; p.nx = nx; p.ny = ny; p.nz = nz; p.ncomponents = ncomponents;
  p.compression_type = compression_type; p.wavelet_domain = wavelet_domain != 0;
  p.myrank = myrank; p.nbproc = nbproc;
  int rc = lsqr_call(p, matrix_sensit, matrix_cons, u, x);
  if (memory) {
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    *memory = (double)(tot - fr) / (1024.0 * 1024.0 * 1024.0);   // device memory in use [GB] (reference: host PSS)
  }
  return rc;
}

// ---- device / pinned-host memory helpers ---------------------------------------------------------
int tfx_device_alloc(void **p, int64_t bytes) {
  TFX_TRY(ensure_init());
  TFX_CUDA(cudaMalloc(p, (size_t)bytes));
  return 0;
}
int tfx_device_free(void *p) {
  if (p) TFX_CUDA(cudaFree(p));
  return 0;
}
int tfx_host_alloc(void **p, int64_t bytes) {
  TFX_TRY(ensure_init());
  TFX_CUDA(cudaMallocHost(p, (size_t)bytes));
  return 0;
}
int tfx_host_free(void *p) {
  if (p) TFX_CUDA(cudaFreeHost(p));
  return 0;
}
int tfx_memcpy(void *dst, const void *src, int64_t bytes) {
  TFX_TRY(ensure_init());
  TFX_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, ctx().stream));
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}
int tfx_device_memset(void *dst, int value, int64_t bytes) {
  TFX_TRY(ensure_init());
  if (bytes > 0) TFX_CUDA(cudaMemsetAsync(dst, value, (size_t)bytes, ctx().stream));
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}
int tfx_device_mem_info(int64_t *free_bytes, int64_t *total_bytes) {
  TFX_TRY(ensure_init());
  size_t f = 0, t = 0;
  TFX_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return 0;
}

int tfx_lsqr_last_timing(double *loop_ms, double *sweep_ms, int32_t *nsweeps) {
  if (loop_ms) *loop_ms = g_last.loop_ms;
  if (sweep_ms) *sweep_ms = g_last.sweep_ms;
  if (nsweeps) *nsweeps = g_last.nsweeps;
  return 0;
}

int tfx_lsqr_last_iterations(int32_t *executed, int32_t *reported) {
  if (executed) *executed = g_last.iters;
  if (reported) *reported = g_last.reported_iters;
  return 0;
}

int tfx_lsqr_last_history(double *r_hist, int32_t capacity, int32_t *iters, int32_t *fused) {
  if (iters) *iters = g_last.iters;
  if (fused) *fused = g_last.fused ? 1 : 0;
  if (r_hist) {
    const int n = std::min<int>(capacity, (int)g_last.history.size());
    for (int i = 0; i < n; ++i) r_hist[i] = g_last.history[i];
  }
  return 0;
}

}  // extern "C"
