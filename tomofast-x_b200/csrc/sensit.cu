// sensit.cu -- device-resident sensitivity assembly (module sensitivity_gravmag).
//
// Replaces calculate_and_write_sensit (src/forward/gravmag/sensitivity_gravmag.F90:82-410) and
// read_sensitivity_kernel (:648-883) without the disk round trip. Per data row, in the reference's
// order: kernel line (graviprism / magprism) -> * column weight (:228) -> [cost_full (:234) ->
// forward wavelet (:237) -> threshold = (N - nel_compressed)-th smallest |x| (:240-256) -> keep
// |x| > threshold (strict), columns ascending (:258-272)] -> real(4) (:265/:290) -> * real(problem
// weight * data weight, 4) in real(4) (:837-843) -> matrix row (idata, d) = model components k
// concatenated at column shift param_shift + (k-1)*nelements (:759-856).
#include "../../include/tfx.h"

#include <cub/device/device_radix_sort.cuh>
#include <thrust/copy.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/transform.h>
#include <thrust/transform_reduce.h>
#include <thrust/unique.h>

#include <algorithm>
#include <time.h>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

namespace tfx {

namespace {

__global__ void __launch_bounds__(256) k_apply_cw(double *__restrict__ lines, const double *__restrict__ cw, int64_t n,
                                                   int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    lines[i] = __dmul_rn(lines[i], cw[i % n]);   // apply_column_weight, sensitivity_gravmag.F90:1042-1054
}

// vals = real(line(col), 4) * wgt (f32 multiply); idx = col + shift; rowid = row; nnz_count[col]++ (uncompressed path).
__global__ void __launch_bounds__(256) k_dense_segment(const double *__restrict__ line, int n, float wgt, int32_t shift,
                                                        int32_t row, int32_t *__restrict__ idx_out,
                                                        float *__restrict__ val_out, int32_t *__restrict__ rowid_out,
                                                        int32_t *__restrict__ nnz_count) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    idx_out[p] = p + shift;
    val_out[p] = __fmul_rn((float)line[p], wgt);
    rowid_out[p] = row;
    if (nnz_count) atomicAdd(&nnz_count[p], 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Compressed row pipeline, batched (r2): every kernel below works on ALL the lines of a batch (blockIdx.y = line), the
// state of a line lives in device arrays, nothing returns to the host inside a batch. Per line, after the batched
// wavelet transform:
//   threshold = (N - nel_compressed)-th smallest |x|  (sensitivity_gravmag.F90:240-256): exact k-th order statistic by
//     MSD radix select on the IEEE bit patterns (monotone for x >= 0): two 12-bit passes over the line fix the top 24
//     bits (exponent + 13 mantissa bits); the few elements that share them are collected and the remaining 39 bits
//     are resolved on that list by one CTA. (Round 1: eight 8-bit passes over the whole line.) A bucket with more than
//     kCandCap members (massive ties, e.g. a line of zeros) falls back to further passes over the line.
//   keep |x| > threshold (strict), columns ascending (:258-272): chunk counts + discarded cost in one pass, an exclusive
//     scan per line, one ordered write pass (ballot ranks) -- replaces cub::DeviceSelect + a separate cost pass.
// ---------------------------------------------------------------------------------------------------------------------
struct RowState {
  double err_sum;              // sum over segments of sqrt(cost_disc / cost_full) (:285)
  long long nnz;               // entries written so far == offset of the next segment
  int bad;                     // 1: a segment kept more than nel_compressed entries (:273-275)
};

struct LineSel {
  unsigned long long prefix;   // bits of the k-th smallest |x| found so far
  long long rank;              // 0-based rank still to locate inside the current bucket
  unsigned count;              // members of the current bucket
  unsigned ncand;              // candidates collected
  int mode;                    // 0: passes over the line; 1: finished
};

// What a line of the batch is: line l = ((b * ndc) + d) * nmc + k (station b of the batch, data component d, model
// component k) -> matrix row, column shift and combined weight (sensitivity_gravmag.F90:837-843).
struct BatchGeom {
  int n;                       // cells per line
  int ndc, nmc;
  int idata0;                  // global 0-based index of the batch's first station
  int dw0;                     // index of that station in dw (dw holds the rank's stations only)
  int param_shift;
  double problem_weight;
  const double *dw;            // data weights [station][ndc]
};

static const int kSelBits = 12;
static const int kSelBins = 1 << kSelBits;
static const unsigned kCandCap = 16384;
static const int kChunk = 4096;      // elements per compaction chunk: 256 threads x 16
static const int kSumBlocks = 256;   // partial sums per line (fixed -> fixed summation order)

__device__ __forceinline__ unsigned long long abs_bits(double x) {
  return (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffull;
}

// line *= column weight (apply_column_weight, :1042-1054), partial[line][block] = sum of the weighted x^2 (:234).
__global__ void __launch_bounds__(256) k_cw_sumsq(double *__restrict__ lines, const double *__restrict__ cw, int n,
                                                   double *__restrict__ partial) {
  __shared__ double red[32];
  double *line = lines + (size_t)blockIdx.y * n;
  const int stride = gridDim.x * 256;
  double s = 0.0;
  int i = blockIdx.x * 256 + threadIdx.x;
  for (; (long long)i + 3LL * stride < n; i += 4 * stride) {
    double a0 = line[i], a1 = line[i + stride], a2 = line[i + 2 * stride], a3 = line[i + 3 * stride];
    const double c0 = cw[i], c1 = cw[i + stride], c2 = cw[i + 2 * stride], c3 = cw[i + 3 * stride];
    a0 = __dmul_rn(a0, c0); a1 = __dmul_rn(a1, c1); a2 = __dmul_rn(a2, c2); a3 = __dmul_rn(a3, c3);
    line[i] = a0; line[i + stride] = a1; line[i + 2 * stride] = a2; line[i + 3 * stride] = a3;
    s = fma(a0, a0, s); s = fma(a1, a1, s); s = fma(a2, a2, s); s = fma(a3, a3, s);
  }
  for (; i < n; i += stride) {
    const double a = __dmul_rn(line[i], cw[i]);
    line[i] = a;
    s = fma(a, a, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
}
// out[line] = sum_b partial[line][b], b ascending per thread, fixed tree across threads.
__global__ void __launch_bounds__(256) k_sum_partials(const double *__restrict__ partial, int nb, double *__restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += partial[(size_t)blockIdx.x * nb + i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

__global__ void k_sel_init(LineSel *sel, long long rank, int nlines, unsigned n) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < nlines) {
    sel[l].prefix = 0ull;
    sel[l].rank = rank;
    sel[l].count = n;
    sel[l].ncand = 0u;
    sel[l].mode = 0;
  }
}

// hist[line][digit] += members of the line's current bucket whose digit (bits [shift, shift + bits)) is `digit`.
__global__ void __launch_bounds__(256) k_sel_hist(const double *__restrict__ lines, int n, const LineSel *__restrict__ sel,
                                                   unsigned *__restrict__ hist, int shift, int bits) {
  __shared__ unsigned sh[kSelBins];
  const LineSel ls = sel[blockIdx.y];
  if (ls.mode != 0) return;
  const int bins = 1 << bits;
  for (int i = threadIdx.x; i < bins; i += 256) sh[i] = 0u;
  __syncthreads();
  const double *line = lines + (size_t)blockIdx.y * n;
  const unsigned long long himask = (shift + bits >= 63) ? 0ull : ((~0ull << (shift + bits)) & 0x7fffffffffffffffull);
  const unsigned long long prefix = ls.prefix;
  const unsigned bmask = (unsigned)bins - 1u;
  const bool first = (himask == 0ull);
  const int stride = gridDim.x * 256;
  const long long step = 4LL * stride;
  const long long nround = ((long long)n + step - 1) / step * step;   // whole warps stay together for the ballots
  for (long long i0 = blockIdx.x * 256 + threadIdx.x; i0 < nround; i0 += step) {
    unsigned long long b[4];
    bool in[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = i0 + (long long)j * stride;
      in[j] = i < n;
      b[j] = in[j] ? abs_bits(line[i]) : ~0ull;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool act = in[j] && ((b[j] & himask) == prefix);
      const unsigned bin = (unsigned)(b[j] >> shift) & bmask;
      if (first) {
        // every element takes part and the digits cluster (exponents): the lanes that share lane 0's digit are
        // counted with one ballot (a line of equal values would otherwise serialise 32-fold), the rest add singly
        const unsigned b0 = __shfl_sync(0xffffffffu, bin, 0);
        const unsigned same = __ballot_sync(0xffffffffu, act && bin == b0);
        if ((threadIdx.x & 31) == 0 && same) atomicAdd(&sh[b0], __popc(same));
        if (act && bin != b0) atomicAdd(&sh[bin], 1u);
      } else if (act) {
        atomicAdd(&sh[bin], 1u);
      }
    }
  }
  __syncthreads();
  unsigned *h = hist + (size_t)blockIdx.y * kSelBins;
  for (int i = threadIdx.x; i < bins; i += 256)
    if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// The two passes every line goes through (digits = bits 62..51 and 50..39), specialised: both digits and the prefix
// test live in the HIGH word of the double, and the first pass needs no prefix test at all (ncu on the generic kernel:
// issue slots 72 % busy, 64-bit shifts and compares). Same histogram as k_sel_hist(shift = 51 / 39, bits = 12).
template <int PASS>
__global__ void __launch_bounds__(256) k_sel_hist12(const double *__restrict__ lines, int n, const LineSel *__restrict__ sel,
                                                     unsigned *__restrict__ hist) {
  __shared__ unsigned sh[kSelBins];
  const LineSel ls = sel[blockIdx.y];
  if (ls.mode != 0) return;
  for (int i = threadIdx.x; i < kSelBins; i += 256) sh[i] = 0u;
  __syncthreads();
  const double *line = lines + (size_t)blockIdx.y * n;
  const unsigned p12 = (unsigned)(ls.prefix >> 51);
  const int stride = gridDim.x * 256;
  const long long step = 4LL * stride;
  const long long nround = ((long long)n + step - 1) / step * step;   // whole warps stay together for the ballots
  for (long long i0 = blockIdx.x * 256 + threadIdx.x; i0 < nround; i0 += step) {
    unsigned hi[4];
    bool in[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = i0 + (long long)j * stride;
      in[j] = i < n;
      hi[j] = in[j] ? ((unsigned)__double2hiint(line[i]) & 0x7fffffffu) : 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (PASS == 1) {
        // the digits cluster (exponents): the lanes that share lane 0's digit are counted with one ballot (a line of
        // equal values would otherwise serialise 32-fold), the rest add singly
        const unsigned bin = hi[j] >> 19;
        const unsigned b0 = __shfl_sync(0xffffffffu, bin, 0);
        const unsigned same = __ballot_sync(0xffffffffu, in[j] && bin == b0);
        if ((threadIdx.x & 31) == 0 && same) atomicAdd(&sh[b0], __popc(same));
        if (in[j] && bin != b0) atomicAdd(&sh[bin], 1u);
      } else {
        if (in[j] && (hi[j] >> 19) == p12) atomicAdd(&sh[(hi[j] >> 7) & 0xfffu], 1u);
      }
    }
  }
  __syncthreads();
  unsigned *h = hist + (size_t)blockIdx.y * kSelBins;
  for (int i = threadIdx.x; i < kSelBins; i += 256)
    if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// One CTA per line: the digit that holds the wanted rank; the bucket shrinks to that digit. hist is zeroed for reuse.
__global__ void __launch_bounds__(256) k_sel_pick(LineSel *sel, unsigned *__restrict__ hist, int shift, int bits) {
  __shared__ unsigned sh[kSelBins];
  __shared__ unsigned long long tsum[257];
  LineSel *ls = &sel[blockIdx.x];
  if (ls->mode != 0) return;
  const int bins = 1 << bits;
  unsigned *h = hist + (size_t)blockIdx.x * kSelBins;
  for (int i = threadIdx.x; i < bins; i += 256) { sh[i] = h[i]; h[i] = 0u; }
  __syncthreads();
  const int per = (bins + 255) / 256;
  unsigned long long mine = 0ull;
  for (int j = 0; j < per; ++j) {
    const int d = threadIdx.x * per + j;
    if (d < bins) mine += sh[d];
  }
  tsum[threadIdx.x + 1] = mine;
  if (threadIdx.x == 0) tsum[0] = 0ull;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 1; i <= 256; ++i) tsum[i] += tsum[i - 1];
  __syncthreads();
  const unsigned long long r = (unsigned long long)ls->rank;   // every thread reads it before the owner rewrites it
  __syncthreads();
  if (r >= tsum[threadIdx.x] && r < tsum[threadIdx.x + 1]) {
    // exactly one thread: its bins hold more than rr members, so the walk ends inside them
    unsigned long long rr = r - tsum[threadIdx.x];
    int d = threadIdx.x * per;
    for (;; ++d) {
      if (rr < (unsigned long long)sh[d]) break;
      rr -= sh[d];
    }
    ls->prefix |= (unsigned long long)d << shift;
    ls->rank = (long long)rr;
    ls->count = sh[d];
  }
}

// Members of the bucket fixed by the passes so far (all bits >= shift) -> cand[line][*] when they fit.
__global__ void __launch_bounds__(256) k_sel_collect(const double *__restrict__ lines, int n, LineSel *sel,
                                                      unsigned long long *__restrict__ cand, unsigned candcap, int shift) {
  LineSel *ls = &sel[blockIdx.y];
  if (ls->mode != 0 || ls->count > candcap) return;
  const double *line = lines + (size_t)blockIdx.y * n;
  const unsigned long long himask = (~0ull << shift) & 0x7fffffffffffffffull;
  const unsigned long long prefix = ls->prefix;
  unsigned long long *out = cand + (size_t)blockIdx.y * candcap;
  const int stride = gridDim.x * 256;
  int i = blockIdx.x * 256 + threadIdx.x;
  for (; (long long)i + 3LL * stride < n; i += 4 * stride) {
    unsigned long long b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = abs_bits(line[i + j * stride]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((b[j] & himask) == prefix) {
        const unsigned pos = atomicAdd(&ls->ncand, 1u);
        if (pos < candcap) out[pos] = b[j];
      }
  }
  for (; i < n; i += stride) {
    const unsigned long long b = abs_bits(line[i]);
    if ((b & himask) == prefix) {
      const unsigned pos = atomicAdd(&ls->ncand, 1u);
      if (pos < candcap) out[pos] = b;
    }
  }
}

// One CTA per line: the remaining `shift` low bits of the k-th value, on the collected candidates (8-bit passes).
__global__ void __launch_bounds__(256) k_sel_finish(LineSel *sel, const unsigned long long *__restrict__ cand,
                                                     unsigned candcap, int shift) {
  __shared__ unsigned sh[256];
  __shared__ unsigned long long s_prefix;
  __shared__ long long s_rank;
  LineSel *ls = &sel[blockIdx.x];
  if (ls->mode != 0 || ls->count > candcap) return;
  const unsigned nc = ls->ncand;   // == count
  const unsigned long long *c = cand + (size_t)blockIdx.x * candcap;
  if (threadIdx.x == 0) { s_prefix = ls->prefix; s_rank = ls->rank; }
  __syncthreads();
  int sft = shift;
  while (sft > 0) {
    const int bits = min(8, sft);
    sft -= bits;
    const int bins = 1 << bits;
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long himask = (~0ull << (sft + bits)) & 0x7fffffffffffffffull;
    const unsigned long long prefix = s_prefix;
    for (unsigned i = threadIdx.x; i < nc; i += 256) {
      const unsigned long long b = c[i];
      if ((b & himask) == prefix) atomicAdd(&sh[(unsigned)(b >> sft) & (unsigned)(bins - 1)], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long r = s_rank;
      int d = 0;
      for (; d < bins - 1; ++d) {
        if (r < (long long)sh[d]) break;
        r -= sh[d];
      }
      s_prefix = prefix | ((unsigned long long)d << sft);
      s_rank = r;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { ls->prefix = s_prefix; ls->rank = s_rank; ls->mode = 1; }
}

// thr[line] = |k-th value|, floored at 1e-30 (:252-256); no_select: every entry above the floor is kept.
__global__ void k_sel_thr(const LineSel *sel, double *thr, int nlines, int no_select) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nlines) return;
  double t = no_select ? -1.0 : __longlong_as_double((long long)sel[l].prefix);
  if (t < 1.e-30) t = 1.e-30;
  thr[l] = t;
}

// cnt[line][chunk] = kept entries of the chunk, disc[line][chunk] = sum of x^2 over its discarded entries (:283).
__global__ void __launch_bounds__(256) k_cmp_count(const double *__restrict__ lines, int n, const double *__restrict__ thr,
                                                    int *__restrict__ cnt, double *__restrict__ disc) {
  __shared__ double red[32];
  __shared__ int wc[8];
  const double *line = lines + (size_t)blockIdx.y * n;
  const double t = thr[blockIdx.y];
  const long long c0 = (long long)blockIdx.x * kChunk;
  double x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const long long i = c0 + j * 256 + threadIdx.x;
    x[j] = (i < n) ? line[i] : 0.0;
  }
  int k = 0;
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (fabs(x[j]) > t) ++k;
    else s = fma(x[j], x[j], s);
  }
  k = __reduce_add_sync(0xffffffffu, k);
  s = block_sum(s, red);
  if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += wc[w];
    const size_t o = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    cnt[o] = tot;
    disc[o] = s;
  }
}

// One CTA per line: cnt -> exclusive offsets inside the line (in place), nsel[line], cost_disc[line].
__global__ void __launch_bounds__(256) k_cmp_scan(int *__restrict__ cnt, const double *__restrict__ disc, int nchunks,
                                                   int *__restrict__ nsel, double *__restrict__ cost_disc) {
  __shared__ double red[32];
  __shared__ int wsum[8];
  __shared__ int carry_s;
  int *c = cnt + (size_t)blockIdx.x * nchunks;
  const double *d = disc + (size_t)blockIdx.x * nchunks;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  double s = 0.0;
  for (int base = 0; base < nchunks; base += 256) {
    const int i = base + threadIdx.x;
    const int v = (i < nchunks) ? c[i] : 0;
    if (i < nchunks) s += d[i];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < w) woff += wsum[j];
    const int carry = carry_s;
    if (i < nchunks) c[i] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + woff + incl;
    __syncthreads();
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    nsel[blockIdx.x] = carry_s;
    cost_disc[blockIdx.x] = s;
  }
}

// Sequential over the lines of the batch (segment order): running entry offset, the reference's element-count check
// (:273-275), compression error terms (:283-285), seg_end.
__global__ void k_batch_advance(RowState *st, const int *__restrict__ nsel, const double *__restrict__ cost_disc,
                                const double *__restrict__ cost_full, long long *__restrict__ line_off,
                                long long *__restrict__ seg_end, long long iseg0, int nlines, int nel_compressed,
                                long long cap) {
  long long nnz = st->nnz;
  double err = st->err_sum;
  int bad = st->bad;
  for (int l = 0; l < nlines; ++l) {
    const int k = nsel[l];
    if (k > nel_compressed || nnz + k > cap) {
      bad = 1;
      line_off[l] = -1;
    } else {
      line_off[l] = nnz;
      nnz += k;
    }
    err += sqrt(cost_disc[l] / cost_full[l]);
    seg_end[iseg0 + l] = nnz;
  }
  st->nnz = nnz;
  st->err_sum = err;
  st->bad = bad;
}

// Ordered write of the kept entries: value = real(line(col), 4) * wgt in real(4) (:265 / :837-843), column = p + shift,
// and the per-cell entry count sensit_nnz (:267).
__global__ void __launch_bounds__(256) k_cmp_write(const double *__restrict__ lines, BatchGeom g,
                                                    const double *__restrict__ thr, const int *__restrict__ chunk_off,
                                                    const long long *__restrict__ line_off, int32_t *__restrict__ idx_out,
                                                    float *__restrict__ val_out, int32_t *__restrict__ rowid_out,
                                                    int32_t *__restrict__ nnz_count) {
  __shared__ int wcnt[16 * 8];
  const int l = blockIdx.y;
  const long long loff = line_off[l];
  if (loff < 0) return;
  const int n = g.n;
  const double *line = lines + (size_t)l * n;
  const double t = thr[l];
  const long long c0 = (long long)blockIdx.x * kChunk;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double x[16];
  unsigned keep = 0u;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const long long i = c0 + j * 256 + threadIdx.x;
    x[j] = (i < n) ? line[i] : 0.0;
  }
  int rank_in_warp[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const bool kp = fabs(x[j]) > t;
    const unsigned m = __ballot_sync(0xffffffffu, kp);
    if (kp) keep |= 1u << j;
    rank_in_warp[j] = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[j * 8 + w] = __popc(m);
  }
  __syncthreads();
  if (w == 0) {
    // exclusive scan of the 128 (row j, warp w) counts in element order: lane holds 4 consecutive entries
    int v[4], sum = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { v[q] = wcnt[lane * 4 + q]; sum += v[q]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    int run = incl - sum;
#pragma unroll
    for (int q = 0; q < 4; ++q) { wcnt[lane * 4 + q] = run; run += v[q]; }
  }
  __syncthreads();
  if (!keep) return;
  const int sgm_d = (l / g.nmc) % g.ndc, sgm_k = l % g.nmc, sgm_b = l / (g.nmc * g.ndc);
  const int32_t row = (g.idata0 + sgm_b) * g.ndc + sgm_d;
  const int32_t shift = g.param_shift + sgm_k * n;
  // combined_weight = real(problem_weight * data_weight(d, idata), 4)
  const float wgt = (float)(g.problem_weight * g.dw[(size_t)(g.dw0 + sgm_b) * g.ndc + sgm_d]);
  const long long base = loff + chunk_off[(size_t)l * gridDim.x + blockIdx.x];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (keep & (1u << j)) {
      const int p = (int)(c0 + j * 256 + threadIdx.x);
      const long long pos = base + wcnt[j * 8 + w] + rank_in_warp[j];
      idx_out[pos] = p + shift;
      val_out[pos] = __fmul_rn((float)x[j], wgt);
      rowid_out[pos] = row;
      atomicAdd(&nnz_count[p], 1);
    }
  }
}
__global__ void k_advance_dense(RowState *st, int n, long long *seg_end, int64_t iseg) {
  st->nnz += n;
  seg_end[iseg] = st->nnz;
}

template <typename T>
int up(DevBuf<T> &d, const T *h, size_t n) {
  TFX_TRY(d.alloc(n));
  if (n) TFX_CUDA(cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx().stream));
  return 0;
}

int upload_grid_impl(GridDev &g, int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                const double *Z1, const double *Z2) {
  g.n = n;
  TFX_TRY(up(g.X1, X1, n)); TFX_TRY(up(g.X2, X2, n)); TFX_TRY(up(g.Y1, Y1, n));
  TFX_TRY(up(g.Y2, Y2, n)); TFX_TRY(up(g.Z1, Z1, n)); TFX_TRY(up(g.Z2, Z2, n));
  return 0;
}

int kernel_error(int e) {
  if (e == 1) return fail(-70, "Data coordinate coincides with model grid boundary (YZ). Adjust the model grid!");
  if (e == 2) return fail(-70, "Data coordinate coincides with model grid boundary (XZ). Adjust the model grid!");
  if (e == 3) return fail(-70, "Zero denominator in gradiprism_full! Adjust the model grid.");
  if (e == 4) return fail(-70, "Bad log argument in gradiprism_full! Adjust the model grid.");
  if (e == 11) return fail(-70, "The model grid X-boundary coincides with the data position");
  if (e == 12) return fail(-70, "The model grid Y-boundary coincides with the data position");
  return fail(-70, "forward kernel error");
}

// d_cw / d_partial: see grav_lines() (only passed when lines_fused_partials() > 0).
int lines_fused_partials(const tfx_sensit_params &P, const GridDev &g) {
  if (P.problem_type == 1 && P.nmodel_components == 1 && P.data_type == 1 && P.ndata_components == 1)
    return grav_lines_fused_partials(g, 1);
  return 0;
}

int compute_lines(const tfx_sensit_params &P, const GridDev &g, int nb, const double *dx, const double *dy,
                  const double *dz, double *d_lines, int *d_err, cudaStream_t st, const double *d_cw = nullptr,
                  double *d_partial = nullptr) {
  if (P.problem_type == 1) {
    if (P.nmodel_components != 1) return fail(-71, "gravity: nmodel_components must be 1");
    if (P.data_type == 1 && P.ndata_components == 1)
      return grav_lines(g, nb, dx, dy, dz, 1, d_lines, d_err, st, d_cw, d_partial);
    if (P.data_type == 2 && P.ndata_components == 1) return grav_lines(g, nb, dx, dy, dz, 2, d_lines, d_err, st);
    if (P.data_type == 2 && P.ndata_components == 6) return grav_full_lines(g, nb, dx, dy, dz, d_lines, d_err, st);
    if (P.data_type == 2) return fail(-72, "Wrong number of gravity gradiometry data components!");   // :210-212
    return fail(-72, "gravity: unknown data_type (1: gz, 2: gradiometry)");
  }
  if (P.problem_type == 2)
    return mag_lines(g, nb, dx, dy, dz, P.nmodel_components, P.ndata_components, P.mi, P.md, P.theta, P.intensity,
                     d_lines, d_err, st);
  return fail(-73, "unknown problem_type");
}

}  // namespace

int upload_grid(GridDev &g, int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                const double *Z1, const double *Z2) {
  return upload_grid_impl(g, n, X1, X2, Y1, Y2, Z1, Z2);
}

// ---- pinned grid (tfx_grid_pin): the six cell-box arrays of a 512x512x128 grid are 1.6 GB of pageable host memory,
// 0.3-0.5 s per upload; an application that assembles several row sets on the same grid (row blocks, the two problems of
// a joint inversion) uploads it once.
static GridDev *g_pin = nullptr;
static const double *g_pin_ptr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

int grid_acquire(GridHold &h, int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                 const double *Z1, const double *Z2, int32_t nx, int32_t ny, int32_t nz) {
  const double *ptr[6] = {X1, X2, Y1, Y2, Z1, Z2};
  bool same = g_pin != nullptr && g_pin->n == n;
  for (int i = 0; i < 6 && same; ++i) same = (g_pin_ptr[i] == ptr[i]);
  if (same) {
    h.g = g_pin;
  } else {
    TFX_TRY(upload_grid_impl(h.own, n, X1, X2, Y1, Y2, Z1, Z2));
    h.g = &h.own;
  }
  TFX_TRY(grid_detect_structured(*h.g, nx, ny, nz, ctx().stream));
  return 0;
}

// ---- phase timing on stderr (option "trace"): where the wall-clock time of an assembly goes
int g_opt_trace = 0;
void trace(const char *label) {
  static double last = 0.0;
  if (!g_opt_trace) return;
  cudaStreamSynchronize(ctx().stream);
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  const double now = ts.tv_sec + 1e-9 * ts.tv_nsec;
  fprintf(stderr, "[tfx trace] %-44s +%8.1f ms\n", label, last > 0.0 ? 1e3 * (now - last) : 0.0);
  last = now;
}

// ---------------------------------------------------------------------------------------------
// Row pipeline for the stations [data0, data0 + ndata_loc) (the reference's data loop,
// sensitivity_gravmag.F90:189-318, for one rank's share of the data). Entries come out sorted by
// (matrix row, column): R.idx = 0-based column (param_shift and (k-1)*N applied), R.rowid = 0-based
// GLOBAL matrix row idata*ndc + d. seg_end[i] = running entry count after segment i (segments in
// the order idata, d, k). dnnz (N): per-cell entry counts (sensit_nnz, :267/:293).
// ---------------------------------------------------------------------------------------------
int assemble_rows_device(const tfx_sensit_params &P, const GridDev &g, const double *d_dx, const double *d_dy,
                         const double *d_dz, const double *d_cw, const double *h_dw, int32_t data0, int32_t ndata_loc,
                         RowTriplets &R, DevBuf<int32_t> &dnnz, std::vector<long long> &seg_end, double *err_sum) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const int32_t N = P.nx * P.ny * P.nz;
  const int32_t ndc = P.ndata_components, nmc = P.nmodel_components;
  const int32_t nel_compressed = (P.compression_type > 0) ? (int32_t)(P.compression_rate * (double)N) : N;
  const int64_t nseg_lines = (int64_t)ndata_loc * ndc * nmc;
  const int64_t cap = (int64_t)nel_compressed * nseg_lines;   // upper bound of nnz
  TFX_TRY(R.idx.alloc((size_t)cap)); TFX_TRY(R.val.alloc((size_t)cap)); TFX_TRY(R.rowid.alloc((size_t)cap));
  R.nnz = 0;
  DevBuf<int> derr;
  TFX_TRY(dnnz.alloc(N)); TFX_TRY(derr.alloc(1));
  TFX_CUDA(cudaMemsetAsync(dnnz.p, 0, (size_t)N * 4, st));
  TFX_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), st));
  seg_end.assign((size_t)nseg_lines, 0);
  if (err_sum) *err_sum = 0.0;
  if (ndata_loc <= 0) return 0;

  // batch of stations whose lines are resident at once (<= ~1 GiB, <= 4096 lines: blockIdx.y = line)
  const size_t per_station = (size_t)N * nmc * ndc;
  int B = (int)std::max<size_t>(1, std::min<size_t>((size_t)ndata_loc, ((size_t)1 << 27) / per_station));
  B = std::max(1, std::min(B, 4096 / (ndc * nmc)));
  DevBuf<double> dl;
  TFX_TRY(dl.alloc(per_station * B));
  const int vgrid = c.num_sms * 8;
  const int maxlines = B * ndc * nmc;

  // device-resident state of the row pipeline (no host synchronisation inside a batch)
  const bool compressed = P.compression_type > 0;
  const long long rank = (long long)N - nel_compressed - 1;   // 0-based rank of sorted(N - nel_compressed), :240-251
  const bool no_select = nel_compressed >= N;
  const int nchunks = (N + kChunk - 1) / kChunk;
  // structured-grid gravity: the line kernel applies the column weight and leaves partial sums of the squares
  const int nfused = compressed ? lines_fused_partials(P, g) : 0;
  const int gsum = nfused > 0 ? nfused : std::min(kSumBlocks, (N + 1023) / 1024);
  const unsigned candcap = std::min<unsigned>(g_opt_sensit_cand_cap > 0 ? (unsigned)g_opt_sensit_cand_cap : kCandCap, (unsigned)N);
  DevBuf<RowState> dst;
  DevBuf<long long> dsegend, dlineoff;
  DevBuf<LineSel> dsel;
  DevBuf<unsigned> dhist;
  DevBuf<unsigned long long> dcand;
  DevBuf<double> dpartial, dcostfull, dcostdisc, dthr, ddisc, ddw;
  DevBuf<int> dcnt, dnsel;
  TFX_TRY(dst.alloc(1));
  TFX_TRY(dsegend.alloc((size_t)nseg_lines));
  TFX_CUDA(cudaMemsetAsync(dst.p, 0, sizeof(RowState), st));
  TFX_TRY(up(ddw, h_dw + (size_t)data0 * ndc, (size_t)ndata_loc * ndc));
  if (compressed) {
    TFX_TRY(dlineoff.alloc((size_t)maxlines)); TFX_TRY(dsel.alloc((size_t)maxlines));
    TFX_TRY(dpartial.alloc((size_t)maxlines * gsum)); TFX_TRY(dcostfull.alloc((size_t)maxlines));
    TFX_TRY(dcostdisc.alloc((size_t)maxlines)); TFX_TRY(dthr.alloc((size_t)maxlines)); TFX_TRY(dnsel.alloc((size_t)maxlines));
    TFX_TRY(dcnt.alloc((size_t)maxlines * nchunks)); TFX_TRY(ddisc.alloc((size_t)maxlines * nchunks));
    if (!no_select) {
      TFX_TRY(dhist.alloc((size_t)maxlines * kSelBins)); TFX_TRY(dcand.alloc((size_t)maxlines * candcap));
      TFX_CUDA(cudaMemsetAsync(dhist.p, 0, (size_t)maxlines * kSelBins * sizeof(unsigned), st));
    }
  }

  int64_t iseg = 0;
  for (int32_t b0 = 0; b0 < ndata_loc; b0 += B) {
    const int nb = std::min<int>(B, ndata_loc - b0);
    TFX_TRY(compute_lines(P, g, nb, d_dx + data0 + b0, d_dy + data0 + b0, d_dz + data0 + b0, dl.p, derr.p, st,
                          nfused > 0 ? d_cw : nullptr, nfused > 0 ? dpartial.p : nullptr));
    const int nseg_b = nb * ndc * nmc;   // segments (lines) of this batch, stored back to back in dl
    if (compressed) {
      // column weight (:228) and cost_full (:234) in one pass (inside the line kernel when it can), then ONE batched
      // wavelet transform of all lines (:237)
      if (nfused == 0) k_cw_sumsq<<<dim3(gsum, nseg_b), 256, 0, st>>>(dl.p, d_cw, N, dpartial.p);
      k_sum_partials<<<nseg_b, 256, 0, st>>>(dpartial.p, gsum, dcostfull.p);
      c.launches += 2;
      TFX_TRY(wavelet3d_device_batch(dl.p, P.nx, P.ny, P.nz, nseg_b, P.compression_type, true, st));
      // line-parallel grid of the passes over the lines: ~8 CTAs per SM in total, whole lines per blockIdx.y
      const int gx = std::max(1, std::min((N + 1023) / 1024, (vgrid + nseg_b - 1) / nseg_b));
      const dim3 glines(gx, nseg_b);
      const int gl = (nseg_b + 255) / 256;
      if (!no_select) {
        k_sel_init<<<gl, 256, 0, st>>>(dsel.p, rank, nseg_b, (unsigned)N);
        // top 24 bits with two passes over the line, candidates, the low 39 bits on the candidates
        k_sel_hist12<1><<<glines, 256, 0, st>>>(dl.p, N, dsel.p, dhist.p);
        k_sel_pick<<<nseg_b, 256, 0, st>>>(dsel.p, dhist.p, 51, 12);
        k_sel_hist12<2><<<glines, 256, 0, st>>>(dl.p, N, dsel.p, dhist.p);
        k_sel_pick<<<nseg_b, 256, 0, st>>>(dsel.p, dhist.p, 39, 12);
        k_sel_collect<<<glines, 256, 0, st>>>(dl.p, N, dsel.p, dcand.p, candcap, 39);
        k_sel_finish<<<nseg_b, 256, 0, st>>>(dsel.p, dcand.p, candcap, 39);
        // lines whose bucket did not fit the candidate list (mode still 0) go on with passes over the line; the
        // others leave these kernels at once
        static const int kRest[4][2] = {{27, 12}, {15, 12}, {3, 12}, {0, 3}};
        for (int q = 0; q < 4; ++q) {
          k_sel_hist<<<glines, 256, 0, st>>>(dl.p, N, dsel.p, dhist.p, kRest[q][0], kRest[q][1]);
          k_sel_pick<<<nseg_b, 256, 0, st>>>(dsel.p, dhist.p, kRest[q][0], kRest[q][1]);
        }
        c.launches += 15;
      }
      k_sel_thr<<<gl, 256, 0, st>>>(dsel.p, dthr.p, nseg_b, no_select ? 1 : 0);
      BatchGeom bg;
      bg.n = N; bg.ndc = ndc; bg.nmc = nmc; bg.idata0 = data0 + b0; bg.dw0 = b0; bg.param_shift = P.param_shift;
      bg.problem_weight = P.problem_weight; bg.dw = ddw.p;
      const dim3 gchunks(nchunks, nseg_b);
      k_cmp_count<<<gchunks, 256, 0, st>>>(dl.p, N, dthr.p, dcnt.p, ddisc.p);
      k_cmp_scan<<<nseg_b, 256, 0, st>>>(dcnt.p, ddisc.p, nchunks, dnsel.p, dcostdisc.p);
      k_batch_advance<<<1, 1, 0, st>>>(dst.p, dnsel.p, dcostdisc.p, dcostfull.p, dlineoff.p, dsegend.p, iseg, nseg_b,
                                       nel_compressed, cap);
      k_cmp_write<<<gchunks, 256, 0, st>>>(dl.p, bg, dthr.p, dcnt.p, dlineoff.p, R.idx.p, R.val.p, R.rowid.p, dnnz.p);
      c.launches += 5;
      iseg += nseg_b;
    } else {
      k_apply_cw<<<vgrid, 256, 0, st>>>(dl.p, d_cw, N, (int64_t)per_station * nb);
      c.launches++;
      for (int b = 0; b < nb; ++b) {
        const int32_t idata = data0 + b0 + b;   // 0-based global station
        for (int d = 0; d < ndc; ++d) {
          const int32_t row = idata * ndc + d;
          // combined_weight = real(problem_weight * data_weight(d, idata), 4)
          const float wgt = (float)(P.problem_weight * h_dw[(size_t)idata * ndc + d]);
          for (int k = 0; k < nmc; ++k, ++iseg) {
            const int sgm = (b * ndc + d) * nmc + k;
            const double *line = dl.p + (size_t)sgm * N;
            const int32_t shift = P.param_shift + k * N;   // 0-based column = p + shift
            // uncompressed general path: the offset is known on the host
            k_dense_segment<<<std::min(vgrid, (N + 255) / 256), 256, 0, st>>>(line, N, wgt, shift, row,
                                                                             R.idx.p + iseg * (int64_t)N, R.val.p + iseg * (int64_t)N,
                                                                             R.rowid.p + iseg * (int64_t)N, dnnz.p);
            k_advance_dense<<<1, 1, 0, st>>>(dst.p, N, dsegend.p, iseg);
            c.launches += 2;
          }
        }
      }
    }
    TFX_CUDA(cudaGetLastError());
  }
  {
    // forward-kernel errors (a station on a grid boundary ...) are fatal: one look at the flag after the last batch
    int e = 0;
    TFX_CUDA(cudaMemcpyAsync(&e, derr.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    TFX_CUDA(cudaStreamSynchronize(st));
    if (e) return kernel_error(e);
  }
  RowState hst;
  TFX_CUDA(cudaMemcpyAsync(&hst, dst.p, sizeof(RowState), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaMemcpyAsync(seg_end.data(), dsegend.p, (size_t)nseg_lines * sizeof(long long), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  if (hst.bad) return fail(-79, "Wrong number of elements in calculate_and_write_sensit!");
  R.nnz = hst.nnz;
  if (err_sum) *err_sum = hst.err_sum;
  return 0;
}

__global__ void __launch_bounds__(256) k_add_i32(int32_t *__restrict__ a, int64_t n, int32_t delta) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] += delta;
}

// pend <- pend ++ R (row ids already absolute); R is released.
static int pend_concat(Matrix &M, RowTriplets &R) {
  cudaStream_t st = ctx().stream;
  if (M.pend.nnz == 0) {
    M.pend.idx.release(); M.pend.val.release(); M.pend.rowid.release();
    std::swap(M.pend.idx.p, R.idx.p); std::swap(M.pend.idx.n, R.idx.n);
    std::swap(M.pend.val.p, R.val.p); std::swap(M.pend.val.n, R.val.n);
    std::swap(M.pend.rowid.p, R.rowid.p); std::swap(M.pend.rowid.n, R.rowid.n);
    M.pend.nnz = R.nnz;
    if (!M.pend.idx.p) { TFX_TRY(M.pend.idx.alloc(1)); TFX_TRY(M.pend.val.alloc(1)); TFX_TRY(M.pend.rowid.alloc(1)); }
  } else if (R.nnz > 0) {
    RowTriplets cat;
    const int64_t a = M.pend.nnz, b = R.nnz;
    TFX_TRY(cat.idx.alloc((size_t)(a + b))); TFX_TRY(cat.val.alloc((size_t)(a + b))); TFX_TRY(cat.rowid.alloc((size_t)(a + b)));
    TFX_CUDA(cudaMemcpyAsync(cat.idx.p, M.pend.idx.p, (size_t)a * 4, cudaMemcpyDeviceToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(cat.idx.p + a, R.idx.p, (size_t)b * 4, cudaMemcpyDeviceToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(cat.val.p, M.pend.val.p, (size_t)a * 4, cudaMemcpyDeviceToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(cat.val.p + a, R.val.p, (size_t)b * 4, cudaMemcpyDeviceToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(cat.rowid.p, M.pend.rowid.p, (size_t)a * 4, cudaMemcpyDeviceToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(cat.rowid.p + a, R.rowid.p, (size_t)b * 4, cudaMemcpyDeviceToDevice, st));
    TFX_CUDA(cudaStreamSynchronize(st));
    M.pend.idx.release(); M.pend.val.release(); M.pend.rowid.release();
    std::swap(M.pend.idx.p, cat.idx.p); std::swap(M.pend.idx.n, cat.idx.n);
    std::swap(M.pend.val.p, cat.val.p); std::swap(M.pend.val.n, cat.val.n);
    std::swap(M.pend.rowid.p, cat.rowid.p); std::swap(M.pend.rowid.n, cat.rowid.n);
    M.pend.nnz = a + b;
  }
  TFX_CUDA(cudaStreamSynchronize(st));
  R.idx.release(); R.val.release(); R.rowid.release(); R.nnz = 0;
  return 0;
}

// Rows built on the host with add() / add_row() / new_row() so far become device-resident pending rows, so that
// host-built blocks (e.g. the reference's clustering constraints) and device-produced blocks can follow each other
// in one matrix. Row order is preserved: pending rows always precede the host rows that were added after them.
int matrix_flush_host_rows(Matrix &M) {
  if (M.nel == 0) return 0;
  if (M.nel_last != M.nel)
    return fail(-17, "Elements were added to the matrix after calling new_row() and before appending device rows!");
  std::vector<int32_t> rowid((size_t)M.nel), idx((size_t)M.nel);
  for (int32_t s = 0; s < M.nl_current; ++s) {
    const int64_t beg = M.ijl[(size_t)s] - 1, end = (s + 1 < M.nl_current) ? M.ijl[(size_t)s + 1] - 1 : M.nel;
    for (int64_t k = beg; k < end; ++k) {
      rowid[(size_t)k] = M.rowptr[(size_t)s] - 1;
      idx[(size_t)k] = M.ija[(size_t)k] - 1;
      if (idx[(size_t)k] < 0 || idx[(size_t)k] >= M.ncolumns)
        return fail(-19, "Sparse matrix column-index validation failed!");
    }
  }
  RowTriplets H;
  TFX_TRY(up(H.rowid, rowid.data(), rowid.size()));
  TFX_TRY(up(H.idx, idx.data(), idx.size()));
  TFX_TRY(up(H.val, M.sa.data(), (size_t)M.nel));
  H.nnz = M.nel;
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  TFX_TRY(pend_concat(M, H));
  M.sa.clear(); M.ija.clear();
  M.nel = M.nel_last = 0;
  M.nl_current = 0;
  return 0;
}

int matrix_append_triplets(Matrix &M, RowTriplets &R, int32_t nrows, int32_t ncolumns) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  if (M.finalized) return fail(-26, "sparse_matrix: rows cannot be appended to a finalized matrix");
  if (M.ncolumns != ncolumns)
    return fail(-27, "sparse_matrix: appended rows have " + std::to_string(ncolumns) + " columns, the matrix " +
                         std::to_string(M.ncolumns));
  if (M.nl_current_all + nrows > M.nl) return fail(-28, "Error in total number of rows in sparse_matrix (append)!");
  if (M.pend.nnz + M.nel + R.nnz > M.nnz)
    return fail(-15, "Error in nnz or nl in sparse_matrix_add! nnz=" + std::to_string(M.nnz));   // the reference's capacity check (:222)
  TFX_TRY(matrix_flush_host_rows(M));
  if (R.nnz > 0 && M.nl_current_all != 0) {
    k_add_i32<<<(int)std::min<int64_t>((R.nnz + 255) / 256, (int64_t)c.num_sms * 16), 256, 0, st>>>(R.rowid.p, R.nnz,
                                                                                                   M.nl_current_all);
    c.launches++;
  }
  TFX_TRY(pend_concat(M, R));
  M.nl_current_all += nrows;
  return 0;
}

int g_opt_dense_block_rows = 0;    // 0: kDenseMaxRows
int g_opt_sensit_cand_cap = 0;     // 0: kCandCap
int g_opt_sensit_row_blocks = 0;   // 1: tfx_sensit_repartition_into / read_sensitivity_kernel_into build one row block per call

// One more row block: the batch's rows become an independent finalized device matrix (T16 layouts when it is big
// enough; its generic CSR copies are then dropped -- 12 B/nnz stay resident).
int matrix_append_block(Matrix &M, RowTriplets &R, int32_t nrows, int32_t ncolumns) {
  if (M.finalized) return fail(-26, "sparse_matrix: rows cannot be appended to a finalized matrix");
  if (M.ncolumns != ncolumns)
    return fail(-27, "sparse_matrix: appended rows have " + std::to_string(ncolumns) + " columns, the matrix " +
                         std::to_string(M.ncolumns));
  if (M.nel != 0 || M.pend.nnz != 0) return fail(-25, "sparse_matrix: row blocks cannot be mixed with other rows");
  if (M.nl_current_all + nrows > M.nl) return fail(-28, "Error in total number of rows in sparse_matrix (append)!");
  int64_t have = 0;
  for (const Matrix *b : M.blocks) have += b->nel;
  if (have + R.nnz > M.nnz)
    return fail(-15, "Error in nnz or nl in sparse_matrix_add! nnz=" + std::to_string(M.nnz));   // capacity check (:222)
  Matrix *blk = new Matrix();
  int rc = matrix_from_triplets(*blk, nrows, ncolumns, R);
  if (rc) { delete blk; return rc; }
  if (blk->has_t16) { blk->fwd.release(); blk->trn.release(); blk->has_seg = false; }
  M.blocks.push_back(blk);
  M.block_row0.push_back(M.nl_current_all);
  M.has_blocks = true;
  M.nl_current_all += nrows;
  return 0;
}

// keys (sorted, n entries) -> unique keys + exclusive offsets on the host.
static int runs_of_sorted_keys(const int32_t *d_keys, int64_t n, int32_t max_unique, std::vector<int32_t> &uniq,
                               std::vector<int64_t> &ptr) {
  cudaStream_t st = ctx().stream;
  auto pol = thrust::cuda::par.on(st);
  uniq.clear();
  ptr.assign(1, 0);
  if (n <= 0) return 0;
  DevBuf<int32_t> ucols, ucnt;
  const size_t cap = (size_t)std::min<int64_t>(n, max_unique);
  TFX_TRY(ucols.alloc(cap)); TFX_TRY(ucnt.alloc(cap));
  thrust::device_ptr<const int32_t> K(d_keys);
  thrust::device_ptr<int32_t> UC(ucols.p), UN(ucnt.p);
  size_t nu = 0;
  TFX_THRUST(nu = (size_t)(thrust::reduce_by_key(pol, K, K + n, thrust::make_constant_iterator<int32_t>(1), UC, UN).first - UC));
  uniq.resize(nu);
  std::vector<int32_t> cnt(nu);
  TFX_CUDA(cudaMemcpyAsync(uniq.data(), ucols.p, nu * 4, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaMemcpyAsync(cnt.data(), ucnt.p, nu * 4, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  ptr.resize(nu + 1);
  for (size_t i = 0; i < nu; ++i) ptr[i + 1] = ptr[i] + cnt[i];
  ctx().launches += 2;
  return 0;
}

__global__ void __launch_bounds__(256) k_iota(int32_t *__restrict__ p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}
template <typename T>
__global__ void __launch_bounds__(256) k_gather(const T *__restrict__ src, const int32_t *__restrict__ perm, int64_t n,
                                                T *__restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[(uint32_t)perm[i]];
}

// keys (sorted, n entries) -> S.segmap (unique keys), S.ptr (offsets) and identity items, all on the device: no host
// loop over the segments (a cross-gradient block has 10^8 of them per rank, the column-major copy of a compressed kernel
// 3e7 per row block). *ok = false (nothing built) when some run is longer than kItemLen: the caller takes the host path.
static int seg_from_sorted_keys_device(const int32_t *d_keys, int64_t n, int32_t max_unique, SegMatrix &S, bool *ok) {
  cudaStream_t st = ctx().stream;
  auto pol = thrust::cuda::par.on(st);
  *ok = false;
  if (n <= 0) return 0;
  DevBuf<int32_t> ucnt;
  const size_t cap = (size_t)std::min<int64_t>(n, max_unique);
  TFX_TRY(S.segmap.alloc(cap)); TFX_TRY(ucnt.alloc(cap));
  thrust::device_ptr<const int32_t> K(d_keys);
  thrust::device_ptr<int32_t> UC(S.segmap.p), UN(ucnt.p);
  size_t nu = 0;
  int longest = 0;
  TFX_THRUST(nu = (size_t)(thrust::reduce_by_key(pol, K, K + n, thrust::make_constant_iterator<int32_t>(1), UC, UN).first - UC));
  TFX_THRUST(longest = thrust::reduce(pol, UN, UN + nu, 0, thrust::maximum<int32_t>()));
  ctx().launches += 3;
  if (longest > kItemLen) return 0;
  TFX_TRY(S.ptr.alloc(nu + 1));
  thrust::device_ptr<int64_t> P(S.ptr.p);
  TFX_THRUST(thrust::exclusive_scan(pol, UN, UN + nu, P, (int64_t)0));
  TFX_CUDA(cudaMemcpyAsync(S.ptr.p + nu, &n, 8, cudaMemcpyHostToDevice, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  ctx().launches += 1;
  S.nseg = (int32_t)nu;
  TFX_TRY(seg_set_identity_items(S));
  *ok = true;
  return 0;
}

// Finalized device matrix from entries sorted by (row, column). Takes ownership of R's buffers.
// Matrix rows without entries are not stored (new_row(), sparse_matrix.f90:266-274).
int matrix_from_triplets(Matrix &M, int32_t nl, int32_t ncolumns, RowTriplets &R) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  auto pol = thrust::cuda::par.on(st);
  const int64_t nnz = R.nnz;
  M.nl = nl; M.nl_current_all = nl; M.ncolumns = ncolumns;
  M.device_only = true;

  trace("from_triplets: begin");
  // ---- forward representation: runs of equal row ids
  SegMatrix &F = M.fwd;
  F.nnz = nnz; F.nout = nl; F.nin = ncolumns;
  bool on_device = false;
  TFX_TRY(seg_from_sorted_keys_device(R.rowid.p, nnz, nl, F, &on_device));
  std::vector<int64_t> ptr;
  std::vector<int32_t> segmap;
  if (!on_device) TFX_TRY(runs_of_sorted_keys(R.rowid.p, nnz, nl, segmap, ptr));
  std::swap(F.idx.p, R.idx.p); std::swap(F.idx.n, R.idx.n);
  std::swap(F.val.p, R.val.p); std::swap(F.val.n, R.val.n);
  if (!on_device) {
    F.nseg = (int32_t)segmap.size();
    TFX_TRY(up(F.ptr, ptr.data(), ptr.size()));
    TFX_TRY(up(F.segmap, segmap.data(), segmap.size()));
    TFX_TRY(seg_build_items(F, ptr.data()));
  }

  trace("from_triplets: forward segments");
  // ---- transpose on the device: stable sort by column keeps the row order inside each column.
  // (r2) The sort moves (column, position) pairs only -- a CUB radix sort over the significant column bits with double
  // buffers -- and the row ids / values are gathered through the permutation afterwards, one array at a time: 28 B per
  // entry at the peak instead of ~45 B with thrust::stable_sort_by_key on zipped values. Row blocks near the HBM
  // capacity can be 1.6x larger for the same memory.
  SegMatrix &T = M.trn;
  T.nnz = nnz; T.nout = ncolumns; T.nin = nl;
  std::vector<int64_t> tptr(1, 0);
  std::vector<int32_t> tmap;
  on_device = false;
  if (nnz >= (int64_t)0xFFFFFFF0LL) {
    // more entries than a 32-bit position can address: the (memory-hungrier) zipped sort
    DevBuf<int32_t> keys;
    TFX_TRY(keys.alloc((size_t)nnz));
    TFX_TRY(T.val.alloc((size_t)nnz));
    std::swap(T.idx.p, R.rowid.p); std::swap(T.idx.n, R.rowid.n);
    TFX_CUDA(cudaMemcpyAsync(keys.p, F.idx.p, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, st));
    TFX_CUDA(cudaMemcpyAsync(T.val.p, F.val.p, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, st));
    thrust::device_ptr<int32_t> K(keys.p), Rw(T.idx.p);
    thrust::device_ptr<float> V(T.val.p);
    TFX_THRUST(thrust::stable_sort_by_key(pol, K, K + nnz, thrust::make_zip_iterator(thrust::make_tuple(Rw, V))));
    c.launches += 2;
    TFX_TRY(seg_from_sorted_keys_device(keys.p, nnz, ncolumns, T, &on_device));
    if (!on_device) TFX_TRY(runs_of_sorted_keys(keys.p, nnz, ncolumns, tmap, tptr));
  } else if (nnz > 0) {
    DevBuf<int32_t> keys, keys_alt, perm, perm_alt;   // perm holds 32-bit UNSIGNED positions
    TFX_TRY(keys.alloc((size_t)nnz)); TFX_TRY(keys_alt.alloc((size_t)nnz));
    TFX_TRY(perm.alloc((size_t)nnz)); TFX_TRY(perm_alt.alloc((size_t)nnz));
    TFX_CUDA(cudaMemcpyAsync(keys.p, F.idx.p, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, st));
    k_iota<<<(int)std::min<int64_t>((nnz + 255) / 256, (int64_t)c.num_sms * 16), 256, 0, st>>>(perm.p, nnz);
    cub::DoubleBuffer<int32_t> dk(keys.p, keys_alt.p), dv(perm.p, perm_alt.p);
    int end_bit = 1;
    while (end_bit < 31 && ((int64_t)1 << end_bit) < (int64_t)ncolumns) ++end_bit;
    size_t tb = 0;
    TFX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, nnz, 0, end_bit, st));
    DevBuf<unsigned char> tmp;
    TFX_TRY(tmp.alloc(tb + 16));
    TFX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, nnz, 0, end_bit, st));
    c.launches += 5;
    TFX_CUDA(cudaStreamSynchronize(st));
    tmp.release();
    if (dk.Current() == keys.p) keys_alt.release(); else keys.release();
    if (dv.Current() == perm.p) perm_alt.release(); else perm.release();
    const int32_t *skeys = dk.Current(), *sperm = dv.Current();
    TFX_TRY(seg_from_sorted_keys_device(skeys, nnz, ncolumns, T, &on_device));
    if (!on_device) TFX_TRY(runs_of_sorted_keys(skeys, nnz, ncolumns, tmap, tptr));
    keys.release(); keys_alt.release();
    const int gg = (int)std::min<int64_t>((nnz + 255) / 256, (int64_t)c.num_sms * 16);
    TFX_TRY(T.idx.alloc((size_t)nnz));
    k_gather<int32_t><<<gg, 256, 0, st>>>(R.rowid.p, sperm, nnz, T.idx.p);     // row ids become the gathered index of A^T
    TFX_CUDA(cudaStreamSynchronize(st));
    R.rowid.release();
    TFX_TRY(T.val.alloc((size_t)nnz));
    k_gather<float><<<gg, 256, 0, st>>>(F.val.p, sperm, nnz, T.val.p);
    c.launches += 2;
    TFX_CUDA(cudaStreamSynchronize(st));
  } else {
    TFX_TRY(T.idx.alloc(1)); TFX_TRY(T.val.alloc(1));
    R.rowid.release();
  }
  if (!on_device) {
    T.nseg = (int32_t)tmap.size();
    TFX_TRY(up(T.ptr, tptr.data(), tptr.size()));
    TFX_TRY(up(T.segmap, tmap.data(), tmap.size()));
    TFX_TRY(seg_build_items(T, tptr.data()));
  }

  M.has_seg = true;
  trace("from_triplets: transposed copy (sort + gathers)");
  TFX_TRY(matrix_build_t16(M));
  trace("from_triplets: T16 layouts");
  if (M.nnz < nnz) M.nnz = nnz;   // a matrix created by initialize() keeps its capacity (reset() + rebuild)
  M.nel = nnz;
  M.nl_nonempty = F.nseg;
  M.finalized = true;
  return 0;
}

}  // namespace tfx

using namespace tfx;

extern "C" int tfx_sensit_lines(const tfx_sensit_params *par, const double *X1, const double *X2, const double *Y1,
                                const double *Y2, const double *Z1, const double *Z2, int32_t nb, const double *data_X,
                                const double *data_Y, const double *data_Z, double *lines) {
  TFX_TRY(ensure_init());
  cudaStream_t st = ctx().stream;
  const tfx_sensit_params &P = *par;
  const int32_t N = P.nx * P.ny * P.nz;
  GridHold gh;
  TFX_TRY(grid_acquire(gh, N, X1, X2, Y1, Y2, Z1, Z2, P.nx, P.ny, P.nz));
  GridDev &g = *gh.g;
  DevBuf<double> dx, dy, dz, dl;
  DevBuf<int> derr;
  TFX_TRY(up(dx, data_X, nb)); TFX_TRY(up(dy, data_Y, nb)); TFX_TRY(up(dz, data_Z, nb));
  const size_t per = (size_t)N * P.nmodel_components * P.ndata_components;
  TFX_TRY(dl.alloc(per * nb));
  TFX_TRY(derr.alloc(1));
  TFX_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), st));
  TFX_TRY(compute_lines(P, g, nb, dx.p, dy.p, dz.p, dl.p, derr.p, st));
  int e = 0;
  TFX_CUDA(cudaMemcpyAsync(&e, derr.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaMemcpyAsync(lines, dl.p, per * nb * sizeof(double), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  if (e) return kernel_error(e);
  return 0;
}

extern "C" int tfx_grid_pin(int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                            const double *Z1, const double *Z2) {
  TFX_TRY(ensure_init());
  if (n <= 0 || !X1 || !X2 || !Y1 || !Y2 || !Z1 || !Z2) return fail(-88, "grid_pin: wrong arguments");
  if (g_pin) { delete g_pin; g_pin = nullptr; }
  GridDev *g = new GridDev();
  int rc = upload_grid_impl(*g, n, X1, X2, Y1, Y2, Z1, Z2);
  if (!rc && cudaStreamSynchronize(ctx().stream) != cudaSuccess) rc = fail(-100, "grid_pin: upload failed");
  if (rc) { delete g; return rc; }
  g_pin = g;
  const double *ptr[6] = {X1, X2, Y1, Y2, Z1, Z2};
  for (int i = 0; i < 6; ++i) g_pin_ptr[i] = ptr[i];
  return 0;
}
extern "C" int tfx_grid_unpin(void) {
  if (g_pin) { delete g_pin; g_pin = nullptr; }
  for (int i = 0; i < 6; ++i) g_pin_ptr[i] = nullptr;
  return 0;
}

extern "C" int tfx_debug_math(int64_t n, const double *y, const double *x, double *out) {
  TFX_TRY(ensure_init());
  cudaStream_t st = ctx().stream;
  if (n <= 0) return 0;
  DevBuf<double> dy, dx, dout;
  TFX_TRY(up(dy, y, (size_t)n)); TFX_TRY(up(dx, x, (size_t)n));
  TFX_TRY(dout.alloc((size_t)n * 4));
  TFX_TRY(debug_math(n, dy.p, dx.p, dout.p, st));
  TFX_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int tfx_calculate_sensit(tfx_matrix **out, const tfx_sensit_params *par, const double *X1, const double *X2,
                                    const double *Y1, const double *Y2, const double *Z1, const double *Z2,
                                    const double *data_X, const double *data_Y, const double *data_Z,
                                    const double *column_weight_full, const double *data_weight, int32_t *sensit_nnz,
                                    double *comp_error, int64_t *nnz_total) {
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const tfx_sensit_params &P = *par;
  if (!out) return fail(-74, "calculate_sensit: null handle");
  if (P.compression_rate < 0 || P.compression_rate > 1)
    return fail(-75, "Wrong compression rate! It must be between 0 and 1.");
  const int64_t N64 = (int64_t)P.nx * P.ny * P.nz;
  if (N64 <= 0 || N64 > 2000000000LL) return fail(-76, "calculate_sensit: wrong grid size");
  const int32_t N = (int32_t)N64;
  const int32_t ndc = P.ndata_components, nmc = P.nmodel_components;
  // get_nel_compressed, sensitivity_gravmag.F90:64-77
  const int32_t nel_compressed = (P.compression_type > 0) ? (int32_t)(P.compression_rate * (double)N) : N;
  const int32_t nl = P.ndata * ndc;

  GridHold gh;
  TFX_TRY(grid_acquire(gh, N, X1, X2, Y1, Y2, Z1, Z2, P.nx, P.ny, P.nz));
  GridDev &g = *gh.g;
  DevBuf<double> dx, dy, dz, dcw, ddw;
  DevBuf<int> derr;
  TFX_TRY(up(dx, data_X, P.ndata)); TFX_TRY(up(dy, data_Y, P.ndata)); TFX_TRY(up(dz, data_Z, P.ndata));
  TFX_TRY(up(dcw, column_weight_full, N));
  TFX_TRY(up(ddw, data_weight, (size_t)P.ndata * ndc));
  TFX_TRY(derr.alloc(1));
  TFX_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), st));

  tfx_matrix *h = new tfx_matrix();
  Matrix &M = h->m;
  M.device_only = true;
  M.nl = nl; M.nl_current_all = nl; M.ncolumns = P.ncolumns;

  // ------------------------------------------------------------------ uncompressed gravity: dense block(s)
  if (P.compression_type == 0 && P.problem_type == 1 && P.data_type == 1 && ndc == 1 && nmc == 1) {
    const int32_t ncl = P.ncells_local > 0 ? P.ncells_local : N;
    if (P.cell0 < 0 || P.cell0 + ncl > N) { delete h; return fail(-77, "calculate_sensit: wrong local cell range"); }
    // One block of <= kDenseMaxRows stations feeds the fused single-sweep LSQR path; more stations become several dense
    // row blocks (4 B per entry each, split-path LSQR: one transposed and one forward sweep per block and iteration).
    const int32_t brows = (g_opt_dense_block_rows > 0) ? g_opt_dense_block_rows : kDenseMaxRows;
    const int32_t nblk = (P.ndata + brows - 1) / brows;
    for (int32_t bi = 0; bi < nblk; ++bi) {
      const int32_t r0 = bi * brows, nr = std::min(brows, P.ndata - r0);
      Matrix *B = (nblk == 1) ? &M : new Matrix();
      int rc = assemble_grav_dense(B->dense, g, P.cell0, ncl, nr, dx.p + r0, dy.p + r0, dz.p + r0, dcw.p, ddw.p + r0,
                                   P.problem_weight, derr.p, st);
      int e = 0;
      if (!rc && cudaMemcpyAsync(&e, derr.p, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = -100;
      if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(-100, "calculate_sensit: dense assembly failed");
      if (!rc && e) rc = kernel_error(e);
      if (rc) {
        if (nblk > 1) delete B;
        delete h;
        return rc;
      }
      B->dense.col0 = P.param_shift;   // local columns 1..nelements of this rank, shifted for the problem
      B->has_dense = true; B->dense_row0 = 0;
      B->nnz = B->nel = (int64_t)nr * ncl;
      B->finalized = true;
      if (nblk > 1) {
        B->device_only = true;
        B->nl = B->nl_current_all = B->nl_nonempty = nr;
        B->ncolumns = P.ncolumns;
        M.blocks.push_back(B);
        M.block_row0.push_back(r0);
        M.has_blocks = true;
      }
    }
    M.nnz = M.nel = (int64_t)P.ndata * ncl;
    M.nl_nonempty = nl;
    M.finalized = true;
    if (sensit_nnz) for (int32_t p = 0; p < N; ++p) sensit_nnz[p] = (p >= P.cell0 && p < P.cell0 + ncl) ? P.ndata : 0;
    if (comp_error) *comp_error = 0.0;
    if (nnz_total) *nnz_total = M.nel;
    *out = h;
    return 0;
  }

  // ------------------------------------------------------------------ general row pipeline
  if (P.cell0 != 0 || (P.ncells_local != 0 && P.ncells_local != N)) {
    delete h;
    return fail(-78, "calculate_sensit: the compressed / magnetic pipeline needs the full grid on the rank");
  }
  const int64_t nseg_lines = (int64_t)P.ndata * ndc * nmc;
  RowTriplets R;
  DevBuf<int32_t> dnnz;
  std::vector<long long> seg_end;
  double err_sum = 0.0;
  std::vector<double> h_dw((size_t)P.ndata * ndc);
  memcpy(h_dw.data(), data_weight, h_dw.size() * sizeof(double));
  int rc = assemble_rows_device(P, g, dx.p, dy.p, dz.p, dcw.p, h_dw.data(), 0, P.ndata, R, dnnz, seg_end, &err_sum);
  if (rc) { delete h; return rc; }
  const int64_t nnz = R.nnz;
  rc = matrix_from_triplets(M, nl, P.ncolumns, R);
  if (rc) { delete h; return rc; }
  if (sensit_nnz) TFX_CUDA(cudaMemcpyAsync(sensit_nnz, dnnz.p, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  if (comp_error) *comp_error = (P.compression_type > 0) ? err_sum / (double)nseg_lines : 0.0;   // :346-353
  if (nnz_total) *nnz_total = nnz;
  *out = h;
  return 0;
}
