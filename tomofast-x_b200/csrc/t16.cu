// t16.cu -- "T16": tiled sparse products with 16-bit in-tile indices (the compressed sensitivity kernel).
//
// Replaces the loops of src/inversion/sparse_matrix.f90 -- add_mult_vector (:313-329) through the F
// layout, add_trans_mult_vector (:388-405) through the T layout -- for the big wavelet-compressed
// matrix_sensit (sensitivity_gravmag.F90:759-856: uniform ~nel_compressed entries per row, ascending
// columns).
//
// B200 design. A product y = A x is a gather from x. x is cut into tiles of <= 16384 elements: the
// tile is staged in shared memory (<= 128 KB of the 227 KB) and every matrix entry addresses it with a
// 16-bit key, so an entry costs 6 bytes of HBM traffic (f32 value + u16 key) instead of the
// reference's 8 (f32 + int32), all random accesses hit shared memory, and HBM sees two pure streams.
// Entries are grouped by (tile, output element) into contiguous segments of even length (2-entry
// packets: one 8-byte and one 4-byte load per lane), with a dense int64 pointer table per tile.
//   T layout: x = u (data rows, usually ONE tile), outputs = columns      -> S^T u   (DIRECT mode)
//   F layout: x = v (column tiles, thousands),     outputs = data rows    -> S v     (TILES mode)
// DIRECT: every output belongs to exactly one segment of the tile -> y is written once.
// TILES : a CTA owns a contiguous range of tiles (balanced by entries) and accumulates its outputs in
//         a CTA-private partial vector; a second kernel adds the partials in CTA order.
// No atomics anywhere; the summation order is fixed by the layout -> run-to-run deterministic.
// A warp handles 32 consecutive outputs: one coalesced pointer load, then either one segment at a
// time with the whole warp (8 independent f64 accumulators per lane + one shuffle tree) or, when all
// 32 segments are short, one segment per lane (no reduction at all).
#include "common.cuh"
#include "kernels.h"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

#include <algorithm>
#include <vector>

namespace tfx {

int g_opt_t16_min_nnz = 1 << 22;   // matrices with fewer entries stay on the generic CSR kernels
int g_opt_t16_tile = 0;            // 0: automatic; otherwise forced tile size (power of two <= 16384), tests

static const int kT16Threads = 1024;
static const int kT16MaxTile = 16384;
static const int kShortSeg = 24;   // all 32 segments of a block <= this many entries -> one segment per lane

struct T16Args {
  const float *val;
  const uint16_t *key;
  const int64_t *ptr;      // [ntiles * nseg + 1]
  const double *x;         // gathered vector, already shifted: element g of the layout is x[g]
  double *y;               // DIRECT: output vector (offset applied: y[o] is output o of the layout)
  double *partial;         // TILES: [grid][nseg]
  const int32_t *cta_tile; // TILES: [grid + 1]
  int32_t nseg, tile, ntiles, nin;   // nin: number of valid gathered elements (in0-relative)
  int32_t t0;              // DIRECT: the tile to process
  int accumulate;          // DIRECT: y += instead of y =
  const int *done;
};

// Sum over one segment [beg, end) (even bounds) with the whole warp; all lanes return the total.
__device__ __forceinline__ double t16_warp_segment(const float *__restrict__ val, const uint16_t *__restrict__ key,
                                                   const double *xs, int64_t beg, int64_t end, int lane) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0;
  int64_t k = beg + 2 * lane;
  for (; k + 192 < end; k += 256) {
    const float2 v0 = __ldg((const float2 *)(val + k));
    const float2 v1 = __ldg((const float2 *)(val + k + 64));
    const float2 v2 = __ldg((const float2 *)(val + k + 128));
    const float2 v3 = __ldg((const float2 *)(val + k + 192));
    const uint32_t k0 = __ldg((const uint32_t *)(key + k));
    const uint32_t k1 = __ldg((const uint32_t *)(key + k + 64));
    const uint32_t k2 = __ldg((const uint32_t *)(key + k + 128));
    const uint32_t k3 = __ldg((const uint32_t *)(key + k + 192));
    a0 = fma((double)v0.x, xs[k0 & 0xffffu], a0);
    a1 = fma((double)v0.y, xs[k0 >> 16], a1);
    a2 = fma((double)v1.x, xs[k1 & 0xffffu], a2);
    a3 = fma((double)v1.y, xs[k1 >> 16], a3);
    a4 = fma((double)v2.x, xs[k2 & 0xffffu], a4);
    a5 = fma((double)v2.y, xs[k2 >> 16], a5);
    a6 = fma((double)v3.x, xs[k3 & 0xffffu], a6);
    a7 = fma((double)v3.y, xs[k3 >> 16], a7);
  }
  for (; k < end; k += 64) {
    const float2 v0 = __ldg((const float2 *)(val + k));
    const uint32_t k0 = __ldg((const uint32_t *)(key + k));
    a0 = fma((double)v0.x, xs[k0 & 0xffffu], a0);
    a1 = fma((double)v0.y, xs[k0 >> 16], a1);
  }
  return warp_sum(((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)));
}

// 32 consecutive outputs [o0, o0 + 32) of tile t: lane j returns the sum of segment o0 + j.
__device__ __forceinline__ double t16_block32(const T16Args &a, const double *xs, int64_t tbase, int o0, int lane) {
  const int o = o0 + lane;
  const int64_t pb = (o <= a.nseg) ? __ldg(a.ptr + tbase + o) : 0;
  int64_t pe = __shfl_down_sync(0xffffffffu, pb, 1);
  if (lane == 31) pe = (o + 1 <= a.nseg) ? __ldg(a.ptr + tbase + o + 1) : 0;
  const int64_t len = (o < a.nseg) ? (pe - pb) : 0;
  int maxlen = (int)min(len, (int64_t)0x7fffffff);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, s));
  double result = 0.0;
  if (maxlen == 0) return result;
  if (maxlen <= kShortSeg) {
    // one segment per lane, sequential (the 32 segments are adjacent in memory: L1-friendly)
    for (int64_t k = pb; k < pb + len; ++k)
      result = fma((double)__ldg(a.val + k), xs[__ldg(a.key + k)], result);
    return result;
  }
  for (int j = 0; j < 32; ++j) {
    const int64_t b = __shfl_sync(0xffffffffu, pb, j);
    const int64_t l = __shfl_sync(0xffffffffu, len, j);
    if (l == 0) continue;
    const double s = t16_warp_segment(a.val, a.key, xs, b, b + l, lane);
    if (lane == j) result = s;
  }
  return result;
}

__device__ __forceinline__ void t16_load_tile(const T16Args &a, double *xs, int t) {
  const int base = t * a.tile;
  for (int i = threadIdx.x; i < a.tile; i += blockDim.x) {
    const int g = base + i;
    xs[i] = (g < a.nin) ? a.x[g] : 0.0;
  }
}

__global__ void __launch_bounds__(kT16Threads, 1) t16_direct_kernel(T16Args a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(16) double xs[];
  t16_load_tile(a, xs, a.t0);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (kT16Threads / 32);
  const int64_t tbase = (int64_t)a.t0 * a.nseg;
  const int nblk = (a.nseg + 31) / 32;
  for (int blk = blockIdx.x * (kT16Threads / 32) + (threadIdx.x >> 5); blk < nblk; blk += nwarps) {
    const double r = t16_block32(a, xs, tbase, blk * 32, lane);
    const int o = blk * 32 + lane;
    if (o < a.nseg) a.y[o] = a.accumulate ? (a.y[o] + r) : r;
  }
}

__global__ void __launch_bounds__(kT16Threads, 1) t16_tiles_kernel(T16Args a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(16) double xs[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nblk = (a.nseg + 31) / 32;
  double *mine = a.partial + (int64_t)blockIdx.x * a.nseg;
  const int t_lo = a.cta_tile[blockIdx.x], t_hi = a.cta_tile[blockIdx.x + 1];
  if (t_lo >= t_hi) {   // idle CTA: its partial vector must still read as zero
    for (int o = threadIdx.x; o < a.nseg; o += blockDim.x) mine[o] = 0.0;
    return;
  }
  for (int t = t_lo; t < t_hi; ++t) {
    t16_load_tile(a, xs, t);
    __syncthreads();
    const int64_t tbase = (int64_t)t * a.nseg;
    for (int blk = wid; blk < nblk; blk += kT16Threads / 32) {
      const double r = t16_block32(a, xs, tbase, blk * 32, lane);
      const int o = blk * 32 + lane;
      if (o < a.nseg) mine[o] = (t == t_lo) ? r : (mine[o] + r);
    }
    __syncthreads();
  }
}

// y[o] (+)= sum_b partial[b][o], b ascending.
__global__ void __launch_bounds__(256) t16_reduce_kernel(const double *__restrict__ partial, int nblocks, int nseg,
                                                         double *__restrict__ y, int accumulate, const int *done) {
  if (done && *done) return;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < nseg; o += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * nseg + o];
    y[o] = accumulate ? (y[o] + s) : s;
  }
}

__global__ void __launch_bounds__(256) t16_zero_kernel(double *y, int64_t n, const int *done) {
  if (done && *done) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// Product
// ---------------------------------------------------------------------------------------------
int t16_spmv(T16Matrix &m, const double *d_x, double *d_y, bool accumulate, int32_t xshift, const int *d_done,
             cudaStream_t st) {
  Context &c = ctx();
  if (!m.valid) return fail(-40, "t16: layout was not built");
  if (!accumulate) {   // outputs outside the covered range
    const int64_t n_lo = m.out0, n_hi = (int64_t)m.nout_total - (m.out0 + m.nseg);
    if (n_lo > 0) {
      t16_zero_kernel<<<(int)std::min<int64_t>((n_lo + 255) / 256, c.num_sms * 8), 256, 0, st>>>(d_y, n_lo, d_done);
      c.launches++;
    }
    if (n_hi > 0) {
      t16_zero_kernel<<<(int)std::min<int64_t>((n_hi + 255) / 256, c.num_sms * 8), 256, 0, st>>>(d_y + m.out0 + m.nseg, n_hi, d_done);
      c.launches++;
    }
  }
  if (m.nseg == 0) return 0;
  T16Args a;
  a.val = m.val.p; a.key = m.key.p; a.ptr = m.ptr.p;
  a.x = d_x + m.in0 - xshift;
  a.y = d_y + m.out0;
  a.partial = m.partial.p; a.cta_tile = m.cta_tile.p;
  a.nseg = m.nseg; a.tile = m.tile; a.ntiles = m.ntiles; a.nin = m.nin;
  a.t0 = 0; a.accumulate = accumulate ? 1 : 0; a.done = d_done;
  const size_t smem = (size_t)m.tile * sizeof(double);
  if (m.mode == T16_DIRECT) {
    static bool attr = false;
    if (!attr) {
      TFX_CUDA(cudaFuncSetAttribute(t16_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT16MaxTile * 8));
      attr = true;
    }
    const int nblk = (m.nseg + 31) / 32;
    const int grid = std::max(1, std::min(c.num_sms, (nblk + 31) / 32));
    for (int t = 0; t < m.ntiles; ++t) {
      a.t0 = t;
      a.accumulate = (accumulate || t > 0) ? 1 : 0;
      t16_direct_kernel<<<grid, kT16Threads, smem, st>>>(a);
      c.launches++;
    }
  } else {
    static bool attr = false;
    if (!attr) {
      TFX_CUDA(cudaFuncSetAttribute(t16_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT16MaxTile * 8));
      attr = true;
    }
    t16_tiles_kernel<<<m.grid, kT16Threads, smem, st>>>(a);
    c.launches++;
    const int blocks = std::max(1, std::min((m.nseg + 255) / 256, c.num_sms * 8));
    t16_reduce_kernel<<<blocks, 256, 0, st>>>(m.partial.p, m.grid, m.nseg, a.y, accumulate ? 1 : 0, d_done);
    c.launches++;
  }
  TFX_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Builder: from a compressed-segment matrix whose segments hold strictly ascending indices.
// ---------------------------------------------------------------------------------------------
namespace {

// flags[0] = 1 when some segment is not strictly ascending; mm[0] = min idx, mm[1] = max idx.
__global__ void __launch_bounds__(256) t16_scan_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                       int nstored, int *flags, int *mm) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  int lo = 0x7fffffff, hi = -1, bad = 0;
  for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < nstored; s += gridDim.x * wpb) {
    const int64_t b = ptr[s], e = ptr[s + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
      const int v = idx[k];
      if (k > b && idx[k - 1] >= v) bad = 1;
      lo = min(lo, v);
      hi = max(hi, v);
    }
  }
  if (bad) atomicExch(&flags[0], 1);
  if (hi >= 0) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

__global__ void __launch_bounds__(256) t16_segof_kernel(const int32_t *__restrict__ segmap, int nstored, int out0,
                                                        int32_t *__restrict__ segof) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nstored; s += gridDim.x * blockDim.x)
    segof[segmap[s] - out0] = s;
}

__device__ __forceinline__ int64_t t16_lower_bound(const int32_t *__restrict__ idx, int64_t b, int64_t e, int target) {
  while (b < e) {
    const int64_t mid = (b + e) >> 1;
    if (idx[mid] < target) b = mid + 1;
    else e = mid;
  }
  return b;
}

// cnt[t * nseg + o] = even-padded number of entries of output o whose index lies in tile t.
__global__ void __launch_bounds__(256) t16_count_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                        const int32_t *__restrict__ segof, int nseg, int ntiles,
                                                        int tile, int in0, int64_t *__restrict__ cnt) {
  const int64_t total = (int64_t)nseg * ntiles;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / nseg), o = (int)(i % nseg);
    const int s = segof[o];
    int64_t n = 0;
    if (s >= 0) {
      const int64_t b = ptr[s], e = ptr[s + 1];
      const int64_t lo = (ntiles == 1) ? b : t16_lower_bound(idx, b, e, in0 + t * tile);
      const int64_t hi = (t + 1 == ntiles) ? e : t16_lower_bound(idx, lo, e, in0 + (t + 1) * tile);
      n = hi - lo;
    }
    cnt[i] = (n + 1) & ~(int64_t)1;
  }
}

// One warp per (tile, output): copies the run into its padded slot.
__global__ void __launch_bounds__(256) t16_fill_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                       const float *__restrict__ sval, const int32_t *__restrict__ segof,
                                                       int nseg, int ntiles, int tile, int in0,
                                                       const int64_t *__restrict__ tptr, float *__restrict__ val,
                                                       uint16_t *__restrict__ key) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int64_t total = (int64_t)nseg * ntiles;
  for (int64_t i = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); i < total; i += (int64_t)gridDim.x * wpb) {
    const int64_t dst = tptr[i];
    if (tptr[i + 1] == dst) continue;
    const int t = (int)(i / nseg), o = (int)(i % nseg);
    const int s = segof[o];
    const int64_t b = ptr[s], e = ptr[s + 1];
    const int64_t lo = (ntiles == 1) ? b : t16_lower_bound(idx, b, e, in0 + t * tile);
    const int64_t hi = (t + 1 == ntiles) ? e : t16_lower_bound(idx, lo, e, in0 + (t + 1) * tile);
    const int base = in0 + t * tile;
    for (int64_t k = lo + lane; k < hi; k += 32) {
      val[dst + (k - lo)] = sval[k];
      key[dst + (k - lo)] = (uint16_t)(idx[k] - base);
    }
    // padding slot (odd run): value 0 contributes exactly 0; key 0 is always a valid tile element
  }
}

__global__ void t16_tilebase_kernel(const int64_t *__restrict__ tptr, int nseg, int ntiles, int64_t *__restrict__ out) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += gridDim.x * blockDim.x)
    out[t] = tptr[(int64_t)t * nseg];
}

int pow2_floor(int64_t v) {
  int p = 1;
  while ((int64_t)p * 2 <= v) p *= 2;
  return p;
}

}  // namespace

int t16_build(const SegMatrix &src, T16Matrix &T, cudaStream_t st) {
  TFX_TRY(ensure_init());
  Context &c = ctx();
  T.release();
  if (src.nnz == 0 || src.nseg == 0) return 0;
  // ---- index range, ordering, output range
  DevBuf<int> flags, mm;
  TFX_TRY(flags.alloc(1)); TFX_TRY(mm.alloc(2));
  int h_mm[2] = {0x7fffffff, -1}, h_flag = 0;
  TFX_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int), st));
  TFX_CUDA(cudaMemcpyAsync(mm.p, h_mm, sizeof(h_mm), cudaMemcpyHostToDevice, st));
  t16_scan_kernel<<<c.num_sms * 8, 256, 0, st>>>(src.ptr.p, src.idx.p, src.nseg, flags.p, mm.p);
  c.launches++;
  TFX_CUDA(cudaMemcpyAsync(&h_flag, flags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaMemcpyAsync(h_mm, mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
  std::vector<int32_t> h_segmap((size_t)src.nseg);
  TFX_CUDA(cudaMemcpyAsync(h_segmap.data(), src.segmap.p, (size_t)src.nseg * 4, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  if (h_flag) return 0;   // indices not strictly ascending inside a segment: stay on the generic kernels
  const int32_t out_lo = *std::min_element(h_segmap.begin(), h_segmap.end());
  const int32_t out_hi = *std::max_element(h_segmap.begin(), h_segmap.end());
  T.out0 = out_lo;
  T.nseg = out_hi - out_lo + 1;
  T.in0 = h_mm[0];
  T.nin = h_mm[1] - h_mm[0] + 1;
  T.nnz = src.nnz;
  // ---- tile size and mode
  int tile;
  if (g_opt_t16_tile > 0) {
    tile = g_opt_t16_tile;
  } else if (T.nin <= kT16MaxTile) {
    tile = T.nin;                                  // one tile: DIRECT
  } else {
    // many tiles: enough of them to balance the CTAs, segments as long as possible otherwise
    tile = std::max(1024, std::min(kT16MaxTile, pow2_floor(T.nin / (16 * (int64_t)c.num_sms))));
  }
  tile = std::max(2, std::min(tile, kT16MaxTile));
  T.tile = tile;
  T.ntiles = (T.nin + tile - 1) / tile;
  const int64_t table = (int64_t)T.nseg * T.ntiles;
  // TILES needs a CTA-private partial vector per CTA; DIRECT (tile after tile) is used when the output
  // side is the long one.
  T.mode = (T.ntiles == 1 || (int64_t)T.nseg > (int64_t)1 << 18) ? T16_DIRECT : T16_TILES;
  if (table > ((int64_t)1 << 31)) return 0;        // pointer table would exceed 16 GiB: keep the generic kernels

  // ---- output -> stored segment
  DevBuf<int32_t> segof;
  TFX_TRY(segof.alloc((size_t)T.nseg));
  TFX_CUDA(cudaMemsetAsync(segof.p, 0xff, (size_t)T.nseg * 4, st));
  t16_segof_kernel<<<std::min(c.num_sms * 8, (src.nseg + 255) / 256), 256, 0, st>>>(src.segmap.p, src.nseg, T.out0, segof.p);
  c.launches++;
  // ---- counts -> pointers
  TFX_TRY(T.ptr.alloc((size_t)table + 1));
  const int cgrid = (int)std::min<int64_t>((table + 255) / 256, (int64_t)c.num_sms * 32);
  t16_count_kernel<<<cgrid, 256, 0, st>>>(src.ptr.p, src.idx.p, segof.p, T.nseg, T.ntiles, tile, T.in0, T.ptr.p);
  c.launches++;
  TFX_CUDA(cudaMemsetAsync(T.ptr.p + table, 0, 8, st));
  {
    thrust::device_ptr<int64_t> P(T.ptr.p);
    thrust::exclusive_scan(thrust::cuda::par.on(st), P, P + table + 1, P);
    c.launches += 2;
  }
  int64_t padded = 0;
  TFX_CUDA(cudaMemcpyAsync(&padded, T.ptr.p + table, 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  T.nnz_padded = padded;
  TFX_TRY(T.val.alloc((size_t)padded + 2));
  TFX_TRY(T.key.alloc((size_t)padded + 2));
  TFX_CUDA(cudaMemsetAsync(T.val.p, 0, ((size_t)padded + 2) * 4, st));
  TFX_CUDA(cudaMemsetAsync(T.key.p, 0, ((size_t)padded + 2) * 2, st));
  const int fgrid = (int)std::min<int64_t>((table + 7) / 8, (int64_t)c.num_sms * 32);
  t16_fill_kernel<<<fgrid, 256, 0, st>>>(src.ptr.p, src.idx.p, src.val.p, segof.p, T.nseg, T.ntiles, tile, T.in0, T.ptr.p,
                                         T.val.p, T.key.p);
  c.launches++;
  // ---- TILES schedule: contiguous tile ranges per CTA, balanced by entries (+ a per-tile overhead)
  if (T.mode == T16_TILES) {
    DevBuf<int64_t> tb;
    TFX_TRY(tb.alloc((size_t)T.ntiles + 1));
    t16_tilebase_kernel<<<std::max(1, std::min(64, (T.ntiles + 256) / 256)), 256, 0, st>>>(T.ptr.p, T.nseg, T.ntiles, tb.p);
    c.launches++;
    std::vector<int64_t> h_tb((size_t)T.ntiles + 1);
    TFX_CUDA(cudaMemcpyAsync(h_tb.data(), tb.p, h_tb.size() * 8, cudaMemcpyDeviceToHost, st));
    TFX_CUDA(cudaStreamSynchronize(st));
    T.grid = std::min(c.num_sms, T.ntiles);
    const double overhead = 4.0 * T.nseg + 2.0 * tile;     // pointer reads + tile load, in entry units
    std::vector<double> cost((size_t)T.ntiles + 1, 0.0);
    for (int t = 0; t < T.ntiles; ++t) cost[t + 1] = cost[t] + (double)(h_tb[t + 1] - h_tb[t]) + overhead;
    std::vector<int32_t> ct((size_t)T.grid + 1, 0);
    int t = 0;
    for (int b = 1; b < T.grid; ++b) {
      const double target = cost[T.ntiles] * b / T.grid;
      while (t < T.ntiles && cost[t + 1] <= target) ++t;
      // leave at least one tile for every remaining CTA only when there are enough tiles
      ct[b] = std::max(ct[b - 1], std::min(t, T.ntiles));
    }
    ct[T.grid] = T.ntiles;
    TFX_TRY(T.cta_tile.alloc(ct.size()));
    TFX_CUDA(cudaMemcpyAsync(T.cta_tile.p, ct.data(), ct.size() * 4, cudaMemcpyHostToDevice, st));
    TFX_TRY(T.partial.alloc((size_t)T.grid * T.nseg));
    TFX_CUDA(cudaStreamSynchronize(st));
  }
  TFX_CUDA(cudaStreamSynchronize(st));
  TFX_CUDA(cudaGetLastError());
  T.valid = true;
  return 0;
}

}  // namespace tfx
