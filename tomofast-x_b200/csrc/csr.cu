// csr.cu -- compressed-segment sparse products (forward and transposed) on the device.
//
// Replaces the loops of src/inversion/sparse_matrix.f90: add_mult_vector (:313-329),
// part_mult_vector (:335-367) and add_trans_mult_vector (:388-405). The transposed product does not
// scatter-add (no atomics): the matrix is also held as the CSR of A^T, built once at finalize(), so
// both products are gathers + fixed-order reductions and are run-to-run deterministic.
//
// Values are real(4), vectors real(8); every product is promoted to f64 and accumulated in f64,
// like `b = b + sa(k) * x(ija(k))` in the reference.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <vector>

namespace tfx {

// ---------------------------------------------------------------------------------------------
// Item table
// ---------------------------------------------------------------------------------------------
int seg_set_identity_items(SegMatrix &m) {
  m.identity_items = true;
  m.item_seg.release(); m.item_beg.release(); m.item_end.release(); m.seg_item0.release(); m.partial.release();
  m.nitems = m.nseg;
  m.max_items_per_seg = m.nseg > 0 ? 1 : 0;
  m.avg_len = m.nseg ? (double)m.nnz / (double)m.nseg : 0.0;
  return 0;
}

int seg_build_items(SegMatrix &m, const int64_t *h_ptr) {
  {
    int64_t longest = 0;
    for (int32_t s = 0; s < m.nseg; ++s) longest = std::max(longest, h_ptr[s + 1] - h_ptr[s]);
    if (longest <= kItemLen) return seg_set_identity_items(m);
  }
  m.identity_items = false;
  std::vector<int32_t> item_seg;
  std::vector<int64_t> item_beg, item_end;
  std::vector<int32_t> seg_item0((size_t)m.nseg + 1);
  m.max_items_per_seg = 0;
  for (int32_t s = 0; s < m.nseg; ++s) {
    seg_item0[s] = (int32_t)item_seg.size();
    int64_t b = h_ptr[s], e = h_ptr[s + 1];
    int n = 0;
    for (int64_t k = b; k < e; k += kItemLen) {
      item_seg.push_back(s);
      item_beg.push_back(k);
      item_end.push_back(k + kItemLen < e ? k + kItemLen : e);
      ++n;
    }
    if (n > m.max_items_per_seg) m.max_items_per_seg = n;
  }
  seg_item0[m.nseg] = (int32_t)item_seg.size();
  m.nitems = (int32_t)item_seg.size();
  m.avg_len = m.nitems ? (double)m.nnz / (double)m.nitems : 0.0;
  TFX_TRY(m.item_seg.alloc(item_seg.size()));
  TFX_TRY(m.item_beg.alloc(item_beg.size()));
  TFX_TRY(m.item_end.alloc(item_end.size()));
  TFX_TRY(m.seg_item0.alloc(seg_item0.size()));
  TFX_TRY(m.partial.alloc(item_seg.size()));
  cudaStream_t st = ctx().stream;
  if (m.nitems) {
    TFX_CUDA(cudaMemcpyAsync(m.item_seg.p, item_seg.data(), item_seg.size() * 4, cudaMemcpyHostToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(m.item_beg.p, item_beg.data(), item_beg.size() * 8, cudaMemcpyHostToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(m.item_end.p, item_end.data(), item_end.size() * 8, cudaMemcpyHostToDevice, st));
  }
  TFX_CUDA(cudaMemcpyAsync(m.seg_item0.p, seg_item0.data(), seg_item0.size() * 4, cudaMemcpyHostToDevice, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------
struct SegArgs {
  const int64_t *ptr;        // identity items (item_seg == nullptr): item i = segment i = [ptr[i], ptr[i+1])
  const int64_t *item_beg, *item_end;
  const int32_t *item_seg, *segmap, *idx;
  const float *val;
  const double *x;
  double *partial;   // per item
  double *y;         // direct output when every segment has exactly one item
  int32_t nitems, out_lo, out_hi, xshift;
  int direct;
  const int *done;
};

__device__ __forceinline__ double item_partial_sum(const SegArgs &a, int64_t b, int64_t e, int lane_id, int nlanes) {
  // 4 independent accumulators per lane -> 4 gathers in flight; fixed combination order.
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const double *x = a.x - a.xshift;
  int64_t k = b + lane_id;
  const int64_t stride = nlanes;
  for (; k + 3 * stride < e; k += 4 * stride) {
    float v0 = __ldg(a.val + k), v1 = __ldg(a.val + k + stride), v2 = __ldg(a.val + k + 2 * stride),
          v3 = __ldg(a.val + k + 3 * stride);
    int c0 = __ldg(a.idx + k), c1 = __ldg(a.idx + k + stride), c2 = __ldg(a.idx + k + 2 * stride),
        c3 = __ldg(a.idx + k + 3 * stride);
    double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
    s0 = fma((double)v0, x0, s0);
    s1 = fma((double)v1, x1, s1);
    s2 = fma((double)v2, x2, s2);
    s3 = fma((double)v3, x3, s3);
  }
  for (; k < e; k += stride) s0 = fma((double)__ldg(a.val + k), __ldg(x + __ldg(a.idx + k)), s0);
  return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ int item_segment(const SegArgs &a, int item) { return a.item_seg ? a.item_seg[item] : item; }
__device__ __forceinline__ int64_t item_begin(const SegArgs &a, int item) { return a.item_seg ? a.item_beg[item] : a.ptr[item]; }
__device__ __forceinline__ int64_t item_finish(const SegArgs &a, int item) { return a.item_seg ? a.item_end[item] : a.ptr[item + 1]; }

__device__ __forceinline__ void item_store(const SegArgs &a, int item, int out, double sum) {
  if (a.direct) a.y[out - a.out_lo] += sum;   // one writer per output element
  else a.partial[item] = sum;
}

// One CTA (256 threads) per item: long segments.
__global__ void __launch_bounds__(256) seg_items_block_kernel(SegArgs a) {
  if (a.done && *a.done) return;
  __shared__ double red[32];
  for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
    const int seg = item_segment(a, item);
    const int out = a.segmap[seg];
    if (out < a.out_lo || out >= a.out_hi) continue;
    double s = item_partial_sum(a, item_begin(a, item), item_finish(a, item), threadIdx.x, blockDim.x);
    s = block_sum(s, red);
    if (threadIdx.x == 0) item_store(a, item, out, s);
  }
}

// One warp per item: medium segments.
__global__ void __launch_bounds__(256) seg_items_warp_kernel(SegArgs a) {
  if (a.done && *a.done) return;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int item = blockIdx.x * wpb + (threadIdx.x >> 5); item < a.nitems; item += gridDim.x * wpb) {
    const int seg = item_segment(a, item);
    const int out = a.segmap[seg];
    if (out < a.out_lo || out >= a.out_hi) continue;
    double s = item_partial_sum(a, item_begin(a, item), item_finish(a, item), lane, 32);
    s = warp_sum(s);
    if (lane == 0) item_store(a, item, out, s);
  }
}

// G lanes per item (G = 4, 8, 16): segments of a few to a few dozen entries -- the columns of a cross-gradient /
// gradient-damping block (each cell is touched by ~20 rows), short sensitivity columns. A full warp per item leaves most
// lanes idle there (config E, ncu: 4.0 ms per C^T product with one warp per column, 81 % of the iteration).
template <int G>
__global__ void __launch_bounds__(256) seg_items_group_kernel(SegArgs a) {
  if (a.done && *a.done) return;
  const int sub = threadIdx.x & (G - 1);
  const int gpw = 32 / G, wpb = blockDim.x >> 5;   // groups per warp; the loop bound is uniform over the warp (shuffles)
  for (int item0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * gpw; item0 < a.nitems; item0 += gridDim.x * wpb * gpw) {
    const int item = item0 + ((threadIdx.x & 31) / G);
    const bool valid = item < a.nitems;
    const int out = valid ? a.segmap[item_segment(a, item)] : -1;
    const bool in = valid && !(out < a.out_lo || out >= a.out_hi);
    double s = in ? item_partial_sum(a, item_begin(a, item), item_finish(a, item), sub, G) : 0.0;
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (in && sub == 0) item_store(a, item, out, s);
  }
}

// One thread per item: very short segments (constraint matrices: 1..12 entries per row).
__global__ void __launch_bounds__(256) seg_items_thread_kernel(SegArgs a) {
  if (a.done && *a.done) return;
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < a.nitems; item += gridDim.x * blockDim.x) {
    const int seg = item_segment(a, item);
    const int out = a.segmap[seg];
    if (out < a.out_lo || out >= a.out_hi) continue;
    const double *x = a.x - a.xshift;
    double s = 0.0;
    const int64_t kend = item_finish(a, item);
    for (int64_t k = item_begin(a, item); k < kend; ++k)
      s = fma((double)__ldg(a.val + k), __ldg(x + __ldg(a.idx + k)), s);
    item_store(a, item, out, s);
  }
}

// Adds the per-item partials of each segment in item order.
__global__ void __launch_bounds__(256) seg_combine_kernel(const int32_t *seg_item0, const int32_t *segmap,
                                                          const double *partial, double *y, int32_t nseg,
                                                          int32_t out_lo, int32_t out_hi, const int *done) {
  if (done && *done) return;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += gridDim.x * blockDim.x) {
    const int out = segmap[s];
    if (out < out_lo || out >= out_hi) continue;
    double acc = 0.0;
    for (int i = seg_item0[s]; i < seg_item0[s + 1]; ++i) acc += partial[i];
    y[out - out_lo] += acc;
  }
}

__global__ void __launch_bounds__(256) zero_unless_done_kernel(double *y, int64_t n, const int *done) {
  if (done && *done) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = 0.0;
}

int seg_spmv(SegMatrix &m, const double *d_x, double *d_y, bool accumulate, int32_t out_lo, int32_t out_hi,
             int32_t xshift, const int *d_done, cudaStream_t st) {
  Context &c = ctx();
  const int64_t nout = (int64_t)out_hi - out_lo;
  if (!accumulate && nout > 0) {
    int blocks = (int)std::min<int64_t>((nout + 255) / 256, (int64_t)c.num_sms * 8);
    zero_unless_done_kernel<<<blocks, 256, 0, st>>>(d_y, nout, d_done);
    c.launches++;
  }
  if (m.empty() || m.nitems == 0) return 0;
  SegArgs a;
  a.ptr = m.ptr.p;
  a.item_beg = m.identity_items ? nullptr : m.item_beg.p; a.item_end = m.identity_items ? nullptr : m.item_end.p;
  a.item_seg = m.identity_items ? nullptr : m.item_seg.p; a.segmap = m.segmap.p;
  a.idx = m.idx.p; a.val = m.val.p; a.x = d_x; a.partial = m.partial.p; a.y = d_y;
  a.nitems = m.nitems; a.out_lo = out_lo; a.out_hi = out_hi; a.xshift = xshift;
  a.direct = (m.max_items_per_seg <= 1) ? 1 : 0;
  a.done = d_done;
  if (m.avg_len >= 1024.0) {
    int blocks = std::min(m.nitems, c.num_sms * 16);
    seg_items_block_kernel<<<blocks, 256, 0, st>>>(a);
  } else if (m.avg_len >= 96.0) {
    int blocks = std::min((m.nitems + 7) / 8, c.num_sms * 16);
    seg_items_warp_kernel<<<blocks, 256, 0, st>>>(a);
  } else if (m.avg_len >= 6.0) {
    // lanes per item ~ a third of the average length: 4 (6-24 entries), 8 (24-48), 16 (48-96)
    const int G = m.avg_len >= 48.0 ? 16 : (m.avg_len >= 24.0 ? 8 : 4);
    const int blocks = (int)std::min<int64_t>(((int64_t)m.nitems * G + 255) / 256, (int64_t)c.num_sms * 32);
    if (G == 16) seg_items_group_kernel<16><<<blocks, 256, 0, st>>>(a);
    else if (G == 8) seg_items_group_kernel<8><<<blocks, 256, 0, st>>>(a);
    else seg_items_group_kernel<4><<<blocks, 256, 0, st>>>(a);
  } else {
    int blocks = std::min((m.nitems + 255) / 256, c.num_sms * 16);
    seg_items_thread_kernel<<<blocks, 256, 0, st>>>(a);
  }
  c.launches++;
  if (!a.direct) {
    int blocks = std::min((m.nseg + 255) / 256, c.num_sms * 8);
    seg_combine_kernel<<<blocks, 256, 0, st>>>(m.seg_item0.p, m.segmap.p, m.partial.p, d_y, m.nseg, out_lo, out_hi,
                                               d_done);
    c.launches++;
  }
  TFX_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) vec_add_kernel(double *__restrict__ y, const double *__restrict__ x, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += x[i];
}
int vec_add_inplace(double *y, const double *x, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx().num_sms * 8);
  vec_add_kernel<<<blocks, 256, 0, st>>>(y, x, n);
  ctx().launches++;
  TFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tfx
