// wavelet.cu -- 3-D separable Haar / Daubechies-D4 lifting transforms on the device.
//
// Replaces src/utils/wavelet_transform.F90 (Haar3D :75-153, iHaar3D :158-236, DaubD43D :243-367,
// iDaubD43D :374-498). The reference sweeps whole (n-1)-D slabs once per lifting step and per
// scale (3-5 passes over the volume per scale, ~25 scales for 512x512x128). Here one kernel per
// axis keeps a tile of complete lines in shared memory and runs ALL scales of that axis on it:
// 3 passes over the volume in total (16 B/element each), every global access coalesced.
//
// HBM traffic: three axis passes = 3 x 16 B per element for volumes beyond L2. (An L2-blocked variant -- axis 1 and axis 2
// slab of i3 after slab of i3, option "wavelet_slab_mb" -- is bit-identical but measured slower, see g_opt_wavelet_slab_mb.)
// Tiles are filled with cp.async (LDGSTS, 8 B per element straight into the shared-memory layout): a CTA has its whole
// tile in flight at once instead of one register-staged load per thread.
//
// The volume s(n1,n2,n3) (Fortran order) is viewed per axis as A[outer][L][inner]:
//   axis 1: inner = 1,     L = n1, outer = n2*n3
//   axis 2: inner = n1,    L = n2, outer = n3
//   axis 3: inner = n1*n2, L = n3, outer = 1
// A CTA owns the lines (o0..o0+TO) x (i0..i0+TI) and stores them as smem[l][line] with an odd
// line pitch, so both the coalesced global<->smem copy and the (pair, line) lifting passes are
// bank-conflict free.
//
// Arithmetic is written with explicit round-to-nearest intrinsics (no FMA contraction), in the
// reference's per-element operation order, so results are bit-identical to the CPU restatement.
#include "common.cuh"
#include "kernels.h"

#include <math.h>

#include <algorithm>

namespace tfx {

struct WaveConst {
  double sq2, c0, c1, c2, c3, c4;
  double isq2;       // RN(1/sqrt(2)), seed of the exact fast division
  double half_sq2;   // 0.4999 * sqrt(2): acceptance bound of the fast division in units of ulp(q)
};

__device__ __forceinline__ int ilog2_floor(int n) { return 31 - __clz(n); }

// Number of complete (low, high) pairs at a scale, wavelet_transform.F90:97-101.
__device__ __forceinline__ int npairs(int L, int step) {
  int ngmin = step / 2;  // 0-based index of the first high element
  return (L - 1 - ngmin) / step + 1;
}

template <int TYPE, bool FWD>
__global__ void __launch_bounds__(512) wavelet_axis_kernel(double *__restrict__ s, int L, long long inner,
                                                            long long outer, int TI, int TO, int pitch,
                                                            WaveConst k) {
  extern __shared__ double tile[];
  const long long i0 = (long long)blockIdx.x * TI;
  const long long o0 = (long long)blockIdx.y * TO;
  const int ti_n = (int)min((long long)TI, inner - i0);
  const int to_n = (int)min((long long)TO, outer - o0);
  const int NL = TO * TI;  // line slots (some may be unused at the edges)
  const long long total = (long long)TO * L * TI;

  // Element e of the tile = (ti, l, to), ti fastest: thread order == global memory order inside each outer chunk.
  // The mixed-radix digits are advanced incrementally (blockDim.x per step): two integer divisions per thread and
  // loop instead of three 64-bit ones per element (they were ~90 % of the kernel's instructions).
  const int nt = (int)blockDim.x;
  const int d_ti = nt % TI, d_r = nt / TI, d_l = d_r % L, d_to = d_r / L;
  const int total_i = (int)total;
  {
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    int ti = (int)threadIdx.x % TI, r0 = (int)threadIdx.x / TI;
    int l = r0 % L, to = r0 / L;
    for (int e = threadIdx.x; e < total_i; e += nt) {
      if (ti < ti_n && to < to_n) {
        const double *src = s + ((o0 + to) * L + l) * inner + (i0 + ti);
        const uint32_t dst = tile_s + (uint32_t)(((size_t)l * pitch + to * TI + ti) * sizeof(double));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
      }
      ti += d_ti;
      int cr = 0;
      if (ti >= TI) { ti -= TI; cr = 1; }
      l += d_l + cr;
      int cl = 0;
      if (l >= L) { l -= L; cl = 1; }
      to += d_to + cl;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  // (line, pair) decomposition of the lifting work items: w = p * NL + line, advanced the same way
  const int w_line0 = (int)threadIdx.x % NL, w_p0 = (int)threadIdx.x / NL, w_dline = nt % NL, w_dp = nt / NL;

  const int nscale = (L >= 2) ? ilog2_floor(L) : 0;
  if (FWD) {
    for (int istep = 1; istep <= nscale; ++istep) {
      const int step = 1 << istep, half = step >> 1, ng = npairs(L, step);
      const int items = ng * NL;
      if (TYPE == 1) {
        // Haar: predict / update / normalise are pair-local (wavelet_transform.F90:103-149).
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {
          if (line >= NL) { line -= NL; ++p; }
          double *lo = &tile[(size_t)(p * step) * pitch + line];
          double *hi = &tile[(size_t)(p * step + half) * pitch + line];
          double h = __dsub_rn(*hi, *lo);
          double l = __dadd_rn(*lo, __dmul_rn(h, 0.5));
          *lo = __dmul_rn(l, k.sq2);
          *hi = __ddiv_rn(h, k.sq2);
        }
        __syncthreads();
      } else {
        // D4, wavelet_transform.F90:280-363.
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {  // update 1
          if (line >= NL) { line -= NL; ++p; }
          double *lo = &tile[(size_t)(p * step) * pitch + line];
          double hi = tile[(size_t)(p * step + half) * pitch + line];
          *lo = __dadd_rn(*lo, __dmul_rn(hi, k.c0));
        }
        __syncthreads();
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {  // predict (periodic among the pairs)
          if (line >= NL) { line -= NL; ++p; }
          int pm = (p == 0) ? ng - 1 : p - 1;
          double lo = tile[(size_t)(p * step) * pitch + line];
          double lom = tile[(size_t)(pm * step) * pitch + line];
          double *hi = &tile[(size_t)(p * step + half) * pitch + line];
          *hi = __dsub_rn(__dsub_rn(*hi, __dmul_rn(lo, k.c1)), __dmul_rn(lom, k.c2));
        }
        __syncthreads();
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {  // update 2 + normalise low
          if (line >= NL) { line -= NL; ++p; }
          int pp = (p == ng - 1) ? 0 : p + 1;
          double hin = tile[(size_t)(pp * step + half) * pitch + line];
          double *lo = &tile[(size_t)(p * step) * pitch + line];
          *lo = __dmul_rn(__dsub_rn(*lo, hin), k.c3);
        }
        __syncthreads();
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {  // normalise high
          if (line >= NL) { line -= NL; ++p; }
          double *hi = &tile[(size_t)(p * step + half) * pitch + line];
          *hi = __dmul_rn(*hi, k.c4);
        }
        __syncthreads();
      }
    }
  } else {
    for (int istep = nscale; istep >= 1; --istep) {
      const int step = 1 << istep, half = step >> 1, ng = npairs(L, step);
      const int items = ng * NL;
      if (TYPE == 1) {
        // iHaar, wavelet_transform.F90:186-232.
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {
          if (line >= NL) { line -= NL; ++p; }
          double *lo = &tile[(size_t)(p * step) * pitch + line];
          double *hi = &tile[(size_t)(p * step + half) * pitch + line];
          double l = __ddiv_rn(*lo, k.sq2);
          double h = __dmul_rn(*hi, k.sq2);
          l = __dsub_rn(l, __dmul_rn(h, 0.5));
          *lo = l;
          *hi = __dadd_rn(h, l);
        }
        __syncthreads();
      } else {
        // iD4, wavelet_transform.F90:411-494.
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {  // normalise
          if (line >= NL) { line -= NL; ++p; }
          double *lo = &tile[(size_t)(p * step) * pitch + line];
          double *hi = &tile[(size_t)(p * step + half) * pitch + line];
          *lo = __dmul_rn(*lo, k.c4);
          *hi = __dmul_rn(*hi, k.c3);
        }
        __syncthreads();
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {  // undo update 2
          if (line >= NL) { line -= NL; ++p; }
          int pp = (p == ng - 1) ? 0 : p + 1;
          double hin = tile[(size_t)(pp * step + half) * pitch + line];
          double *lo = &tile[(size_t)(p * step) * pitch + line];
          *lo = __dadd_rn(*lo, hin);
        }
        __syncthreads();
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {  // undo predict
          if (line >= NL) { line -= NL; ++p; }
          int pm = (p == 0) ? ng - 1 : p - 1;
          double lo = tile[(size_t)(p * step) * pitch + line];
          double lom = tile[(size_t)(pm * step) * pitch + line];
          double *hi = &tile[(size_t)(p * step + half) * pitch + line];
          *hi = __dadd_rn(__dadd_rn(*hi, __dmul_rn(lo, k.c1)), __dmul_rn(lom, k.c2));
        }
        __syncthreads();
        for (int w = threadIdx.x, line = w_line0, p = w_p0; w < items; w += nt, line += w_dline, p += w_dp) {  // undo update 1
          if (line >= NL) { line -= NL; ++p; }
          double hi = tile[(size_t)(p * step + half) * pitch + line];
          double *lo = &tile[(size_t)(p * step) * pitch + line];
          *lo = __dsub_rn(*lo, __dmul_rn(hi, k.c0));
        }
        __syncthreads();
      }
    }
  }

  // ---- store
  {
    int ti = (int)threadIdx.x % TI, r0 = (int)threadIdx.x / TI;
    int l = r0 % L, to = r0 / L;
    for (int e = threadIdx.x; e < total_i; e += nt) {
      if (ti < ti_n && to < to_n)
        s[((o0 + to) * L + l) * inner + (i0 + ti)] = tile[(size_t)l * pitch + to * TI + ti];
      ti += d_ti;
      int cr = 0;
      if (ti >= TI) { ti -= TI; cr = 1; }
      l += d_l + cr;
      int cl = 0;
      if (l >= L) { l -= L; cl = 1; }
      to += d_to + cl;
    }
  }
}

static WaveConst make_consts();

// ---------------------------------------------------------------------------------------------------------------------
// Column-layout kernel (round 2). ncu on the kernel above (profiles/r2_wavelet_axis_v1_ncu_summary.csv): 175 thread-
// instructions per element and pass, ALU pipe 60 %, issue slots 85 % busy, DRAM 28 % -- bound by the integer index
// arithmetic of the (line, pair) work items, not by memory and not by FP64. Here a thread owns ONE line (tile column c)
// for the whole kernel: its global base / stride are computed once, shared-memory addresses are row * pitch + c, and
//   * Haar runs three scales per shared-memory round trip in registers: a thread loads 8 rows of its column (stride B =
//     8^round), lifts the pairs (0,1)(2,3)(4,5)(6,7), (0,2)(4,6), (0,4) and stores them back -- a chunk of 8 rows is
//     closed under those three scales, so there is one block barrier per three scales;
//   * D4 keeps its four phases per scale (neighbour pairs, periodic wrap) with the same cheap addressing.
// Tile = L rows x NC columns (lines); rows of the tile are contiguous in global memory for the axis-2 / axis-3 passes
// (lanes read consecutive addresses); for axis 1 (lines contiguous along l) the copy runs with the lanes along l into an
// odd-pitch tile (conflict-free transposition). Same arithmetic, same order per element: bit-identical.
// ---------------------------------------------------------------------------------------------------------------------
// x / sqrt(2), correctly rounded, without the ~20-instruction IEEE division in the common case.
// q = x*ci corrected once with the exact residual (Markstein) is within one ulp of t = x/c; r = x - q*c is then exactly
// representable and equals c*(t - q). If |r| <= 0.4999*c*ulp(q) and q is a normal number that is not a power of two
// (below a power of two the spacing halves), q is the floating-point number nearest to t, i.e. q == RN(x/c). Anything
// else -- one division in several thousand, zeros excepted -- takes __ddiv_rn. The result is ALWAYS the correctly
// rounded quotient, so the transform stays bit-identical to the reference's `/ sqrt(2._CUSTOM_REAL)`.
// Out of line on purpose: inlined, the compiler if-converts the rare branch and runs the whole IEEE division next to the
// fast path for every element (ncu: one MUFU.RCP64H per division).
__device__ __noinline__ double div_slow(double x, double c) { return __ddiv_rn(x, c); }

__device__ __forceinline__ double div_sq2(double x, const WaveConst &k) {
  const double q0 = __dmul_rn(x, k.isq2);
  if (x == 0.0) return q0;                       // +-0 / c = +-0
  const double q = __fma_rn(__fma_rn(-q0, k.sq2, x), k.isq2, q0);
  const double r = __fma_rn(-q, k.sq2, x);
  const int hi = __double2hiint(q);
  const int e = (hi >> 20) & 0x7ff;
  const bool pow2 = ((hi & 0xfffff) | __double2loint(q)) == 0;
  const double bound = __dmul_rn(__hiloint2double((e - 52) << 20, 0), k.half_sq2);   // 0.4999 * c * ulp(q)
  if (e > 64 && e < 1980 && !pow2 && fabs(r) <= bound) return q;
  return div_slow(x, k.sq2);
}

template <bool FWD>
__device__ __forceinline__ void haar_pair(double &lo, double &hi, const WaveConst &k) {
  if (FWD) {   // wavelet_transform.F90:103-149
    const double h = __dsub_rn(hi, lo);
    const double l = __dadd_rn(lo, __dmul_rn(h, 0.5));
    lo = __dmul_rn(l, k.sq2);
    hi = div_sq2(h, k);
  } else {     // :186-232
    double l = div_sq2(lo, k);
    const double h = __dmul_rn(hi, k.sq2);
    l = __dsub_rn(l, __dmul_rn(h, 0.5));
    lo = l;
    hi = __dadd_rn(h, l);
  }
}

// Where the lines of a pass live. Line q (0 <= q < nlines), element l (0 <= l < L):
//   TRANSPOSE (lines contiguous along l):  s[(q / per) * ostride + (q % per) * rstride + l]
//   otherwise (lines strided):             s[(q / inner) * ostride + (q % inner) + l * rstride]
// A whole-volume axis pass has per = 1, ostride = L (axis 1) or ostride = L * inner, rstride = inner (axes 2 / 3); the
// other values address every 8th row of axis 2 (the Haar passes that go with the fused kernel, see run3d()).
struct ColsGeom {
  long long inner, nlines, per, ostride, rstride;
  int L, NC, pitch;
  int nscale;     // scales of this pass (the reference derives them from the FULL axis length)
  // fused Haar axis-1 + axis-2 kernel: the tile's columns are NC consecutive axis-2 positions of ONE plane
  int n2, tpp;    // axis-2 length, tiles per plane (0: not fused)
  int nscale2;    // axis-2 scales done on the tile's groups of 8 columns (<= 3)
  int skip0;      // inverse: columns j % 8 == 0 already went through the axis-1 pass (with the high axis-2 scales)
};

template <int TYPE, bool FWD, bool TRANSPOSE, bool FUSE>
__global__ void __launch_bounds__(256, 4) wavelet_cols_kernel(double *__restrict__ s, ColsGeom G, WaveConst k) {
  extern __shared__ double tile[];
  const int L = G.L, NC = G.NC, pitch = G.pitch;
  const long long inner = G.inner;
  const int nt = (int)blockDim.x;
  const int c = (int)threadIdx.x % NC, rid = (int)threadIdx.x / NC, nrid = nt / NC;
  long long q0;
  int ncol;
  if (FUSE) {
    const long long plane = blockIdx.x / G.tpp;
    const int jt = (int)(blockIdx.x - plane * G.tpp);
    q0 = plane * G.n2 + (long long)jt * NC;
    ncol = min(NC, G.n2 - jt * NC);
  } else {
    q0 = (long long)blockIdx.x * NC;
    ncol = (int)min((long long)NC, G.nlines - q0);
  }
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);

  // ---- load
  long long colbase = 0;
  if (TRANSPOSE) {   // inner == 1: line q = s[q*L .. q*L + L); lanes along l, a warp takes whole lines
    // (running pointers only: the (job / chunk) index arithmetic of the first version was 2/3 of the kernel's
    // instructions, profiles/r2_wavelet_cols_v1_ncu_summary.csv)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nt >> 5;
    const uint32_t dstep = (uint32_t)(32 * pitch * (int)sizeof(double));
    for (int cc = w; cc < ncol; cc += nw) {
      const long long q = q0 + cc, qo = q / G.per;
      const double *src = s + qo * G.ostride + (q - qo * G.per) * G.rstride + lane;
      uint32_t dst = tile_s + (uint32_t)((lane * pitch + cc) * (int)sizeof(double));
#pragma unroll 4
      for (int l = lane; l < L; l += 32) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
        src += 32;
        dst += dstep;
      }
    }
  } else {
    const long long q = q0 + c;
    const long long o = q / inner, i = q - o * inner;
    colbase = o * G.ostride + i;
    if (c < ncol) {
      const double *src = s + colbase + (long long)rid * G.rstride;
      uint32_t dst = tile_s + (uint32_t)((rid * pitch + c) * (int)sizeof(double));
#pragma unroll 4
      for (int l = rid; l < L; l += nrid) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
        src += (long long)nrid * G.rstride;
        dst += (uint32_t)(nrid * pitch * (int)sizeof(double));
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int nscale = G.nscale;
  double *col = tile + c;
  if (TYPE == 1) {
    // ---- Haar: rounds of three scales in registers. Round t works on the rows that are multiples of B = 8^t.
    const int nround = (nscale + 2) / 3;
    for (int tt = 0; tt < nround; ++tt) {
      const int t = FWD ? tt : nround - 1 - tt;
      const int sh = 3 * t;                       // B = 1 << sh
      const int s1 = sh + 1;                      // scales s1, s1 + 1, s1 + 2 (those <= nscale)
      const int nchunk = (L + (8 << sh) - 1) >> (sh + 3);
      const size_t rs = (size_t)pitch << sh;       // distance of two rows of the chunk in the tile
      const bool sc2 = s1 + 1 <= nscale, sc3 = s1 + 2 <= nscale;
      for (int g = (FUSE && G.skip0 && (c & 7) == 0) ? nchunk : rid; g < nchunk; g += nrid) {
        const int r0 = g << (sh + 3);
        double *base = col + (size_t)r0 * pitch;
        double a[8];
        if (r0 + (7 << sh) < L) {
          // all eight rows inside the line (every chunk but the last of a line whose length is not a multiple of 8B)
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = base[j * rs];
          if (FWD) {
            haar_pair<true>(a[0], a[1], k); haar_pair<true>(a[2], a[3], k);
            haar_pair<true>(a[4], a[5], k); haar_pair<true>(a[6], a[7], k);
            if (sc2) { haar_pair<true>(a[0], a[2], k); haar_pair<true>(a[4], a[6], k); }
            if (sc3) haar_pair<true>(a[0], a[4], k);
          } else {
            if (sc3) haar_pair<false>(a[0], a[4], k);
            if (sc2) { haar_pair<false>(a[0], a[2], k); haar_pair<false>(a[4], a[6], k); }
            haar_pair<false>(a[0], a[1], k); haar_pair<false>(a[2], a[3], k);
            haar_pair<false>(a[4], a[5], k); haar_pair<false>(a[6], a[7], k);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) base[j * rs] = a[j];
          continue;
        }
        // ragged end of the line: a pair (lo, hi) of a scale exists when the row of hi is inside the line (npairs(), :97-101)
        bool in[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          in[j] = r0 + (j << sh) < L;
          a[j] = in[j] ? base[j * rs] : 0.0;
        }
        if (FWD) {
          if (in[1]) haar_pair<true>(a[0], a[1], k);
          if (in[3]) haar_pair<true>(a[2], a[3], k);
          if (in[5]) haar_pair<true>(a[4], a[5], k);
          if (in[7]) haar_pair<true>(a[6], a[7], k);
          if (sc2) {
            if (in[2]) haar_pair<true>(a[0], a[2], k);
            if (in[6]) haar_pair<true>(a[4], a[6], k);
          }
          if (sc3 && in[4]) haar_pair<true>(a[0], a[4], k);
        } else {
          if (sc3 && in[4]) haar_pair<false>(a[0], a[4], k);
          if (sc2) {
            if (in[2]) haar_pair<false>(a[0], a[2], k);
            if (in[6]) haar_pair<false>(a[4], a[6], k);
          }
          if (in[1]) haar_pair<false>(a[0], a[1], k);
          if (in[3]) haar_pair<false>(a[2], a[3], k);
          if (in[5]) haar_pair<false>(a[4], a[5], k);
          if (in[7]) haar_pair<false>(a[6], a[7], k);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (in[j]) base[j * rs] = a[j];
      }
      __syncthreads();
    }
    if (FUSE && G.nscale2 > 0) {
      // ---- the three lowest axis-2 scales on the tile's groups of 8 columns (8 consecutive axis-2 positions, the first
      // a multiple of 8): rows of the tile are independent, a thread takes (row i, group g) with the lanes along i
      // (odd pitch: conflict-free). A pair exists when its high element is inside the axis (npairs(), :97-101).
      const int ngrp = (ncol + 7) >> 3;
      const bool sc2 = G.nscale2 >= 2, sc3 = G.nscale2 >= 3;
      for (int w = threadIdx.x; w < L * ngrp; w += nt) {
        const int g = w / L, i = w - g * L;
        double *base = tile + (size_t)i * pitch + g * 8;
        const int nin = ncol - g * 8;   // columns of this group inside the axis (>= 1)
        double a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = (j < nin) ? base[j] : 0.0;
        if (FWD) {
          if (nin > 1) haar_pair<true>(a[0], a[1], k);
          if (nin > 3) haar_pair<true>(a[2], a[3], k);
          if (nin > 5) haar_pair<true>(a[4], a[5], k);
          if (nin > 7) haar_pair<true>(a[6], a[7], k);
          if (sc2) {
            if (nin > 2) haar_pair<true>(a[0], a[2], k);
            if (nin > 6) haar_pair<true>(a[4], a[6], k);
          }
          if (sc3 && nin > 4) haar_pair<true>(a[0], a[4], k);
        } else {
          if (sc3 && nin > 4) haar_pair<false>(a[0], a[4], k);
          if (sc2) {
            if (nin > 2) haar_pair<false>(a[0], a[2], k);
            if (nin > 6) haar_pair<false>(a[4], a[6], k);
          }
          if (nin > 1) haar_pair<false>(a[0], a[1], k);
          if (nin > 3) haar_pair<false>(a[2], a[3], k);
          if (nin > 5) haar_pair<false>(a[4], a[5], k);
          if (nin > 7) haar_pair<false>(a[6], a[7], k);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < nin) base[j] = a[j];
      }
      __syncthreads();
    }
  } else {
    // ---- D4: four phases per scale (wavelet_transform.F90:280-363 forward, :411-494 inverse); pair p of scale istep =
    // rows p*step (low) and p*step + half (high); the neighbour pairs wrap periodically among the ng complete pairs.
    for (int ii = 1; ii <= nscale; ++ii) {
      const int istep = FWD ? ii : nscale + 1 - ii;
      const int step = 1 << istep, half = step >> 1, ng = npairs(L, step);
      const size_t sp = (size_t)step * pitch, hp = (size_t)half * pitch;
      if (FWD) {
        for (int p = rid; p < ng; p += nrid) {                      // update 1
          double *lo = col + p * sp;
          *lo = __dadd_rn(*lo, __dmul_rn(lo[hp], k.c0));
        }
        __syncthreads();
        for (int p = rid; p < ng; p += nrid) {                      // predict
          const int pm = (p == 0) ? ng - 1 : p - 1;
          double *lo = col + p * sp;
          lo[hp] = __dsub_rn(__dsub_rn(lo[hp], __dmul_rn(*lo, k.c1)), __dmul_rn(col[pm * sp], k.c2));
        }
        __syncthreads();
        for (int p = rid; p < ng; p += nrid) {                      // update 2 + normalise low
          const int pp = (p == ng - 1) ? 0 : p + 1;
          double *lo = col + p * sp;
          *lo = __dmul_rn(__dsub_rn(*lo, col[pp * sp + hp]), k.c3);
        }
        __syncthreads();
        for (int p = rid; p < ng; p += nrid) {                      // normalise high
          double *hi = col + p * sp + hp;
          *hi = __dmul_rn(*hi, k.c4);
        }
        __syncthreads();
      } else {
        for (int p = rid; p < ng; p += nrid) {                      // normalise
          double *lo = col + p * sp;
          *lo = __dmul_rn(*lo, k.c4);
          lo[hp] = __dmul_rn(lo[hp], k.c3);
        }
        __syncthreads();
        for (int p = rid; p < ng; p += nrid) {                      // undo update 2
          const int pp = (p == ng - 1) ? 0 : p + 1;
          double *lo = col + p * sp;
          *lo = __dadd_rn(*lo, col[pp * sp + hp]);
        }
        __syncthreads();
        for (int p = rid; p < ng; p += nrid) {                      // undo predict
          const int pm = (p == 0) ? ng - 1 : p - 1;
          double *lo = col + p * sp;
          lo[hp] = __dadd_rn(__dadd_rn(lo[hp], __dmul_rn(*lo, k.c1)), __dmul_rn(col[pm * sp], k.c2));
        }
        __syncthreads();
        for (int p = rid; p < ng; p += nrid) {                      // undo update 1
          double *lo = col + p * sp;
          *lo = __dsub_rn(*lo, __dmul_rn(lo[hp], k.c0));
        }
        __syncthreads();
      }
    }
  }

  // ---- store
  if (TRANSPOSE) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nt >> 5;
    for (int cc = w; cc < ncol; cc += nw) {
      const long long q = q0 + cc, qo = q / G.per;
      double *dst = s + qo * G.ostride + (q - qo * G.per) * G.rstride + lane;
      const double *src = tile + (size_t)lane * pitch + cc;
#pragma unroll 4
      for (int l = lane; l < L; l += 32) {
        *dst = *src;
        dst += 32;
        src += (size_t)32 * pitch;
      }
    }
  } else if (c < ncol) {
    double *dst = s + colbase + (long long)rid * G.rstride;
    const double *src = col + (size_t)rid * pitch;
#pragma unroll 4
    for (int l = rid; l < L; l += nrid) {
      *dst = *src;
      dst += (long long)nrid * G.rstride;
      src += (size_t)nrid * pitch;
    }
  }
}

int g_opt_wavelet_cols = 1;   // 0: always the generic (line, pair) kernel
int g_opt_wavelet_tile_kb = 0;   // 0: automatic (32 KB tiles, 64 KB for axes longer than 256)

// Columns (lines) per CTA of the column-layout kernel for lines of length L; 0: the axis does not fit any tile.
static int cols_tile(int L, bool transpose, long long inner, bool fuse = false) {
  if (!g_opt_wavelet_cols || L < 2) return 0;
  // the largest power of two in [8, 256] whose tile stays within the budget (small tiles: many CTAs per SM in different
  // phases -- load / lift / store -- keep the memory pipeline busy)
  // measured (512x512x128 Haar): 16 KB tiles 0.51 ms, 32 KB 0.46 ms, 64 KB 0.39 ms; 256x256x64: 0.086 / 0.075 / 0.078 ms
  // fused Haar kernel (512x512x128): 64 KB tiles 0.295 ms per transform, 32 KB 0.281 ms, 128 KB 0.50 ms
  const size_t budget = (g_opt_wavelet_tile_kb > 0 ? (size_t)std::max(8, g_opt_wavelet_tile_kb) : ((L > 256 && !fuse) ? 64 : 32)) * 1024;
  int NC = 256;
  while (NC > 8 && (size_t)L * NC * sizeof(double) > budget) NC >>= 1;
  if ((size_t)L * (NC + 1) * sizeof(double) > 140 * 1024) return 0;   // axis too long for any tile: generic kernel
  if (!transpose && inner < NC && inner < 16) return 0;   // tiny inner extents: lanes would not read contiguous memory
  return NC;
}

// Launches the column-layout kernel on the lines described by G (L, inner, nlines, per, ostride, rstride, nscale and, for
// the fused kernel, n2 / nscale2 / skip0 filled in by the caller; NC, pitch, tpp here).
template <int TYPE, bool FWD>
static int launch_cols_geom(double *d_s, ColsGeom G, bool transpose, bool fuse, long long nplanes, cudaStream_t st) {
  const int NC = cols_tile(G.L, transpose, G.inner, fuse);
  if (NC == 0) return fail(-23, "wavelet: internal error (column kernel launched on an axis that does not fit)");
  G.NC = NC;
  G.pitch = transpose ? NC + 1 : NC;
  G.tpp = fuse ? (G.n2 + NC - 1) / NC : 0;
  const size_t smem = (size_t)G.L * G.pitch * sizeof(double);
  const long long grid = fuse ? nplanes * G.tpp : (G.nlines + NC - 1) / NC;
  if (grid > 0x7fffffffLL) return fail(-23, "wavelet: volume too large for one launch");
  if (grid <= 0) return 0;
#define TFX_WCOLS(T, F)                                                                                     \
  do {                                                                                                      \
    auto kern = wavelet_cols_kernel<TYPE, FWD, T, F>;                                                       \
    TFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 144 * 1024));          \
    kern<<<(unsigned)grid, 256, smem, st>>>(d_s, G, make_consts());                                         \
  } while (0)
  if (fuse) {
    if (TYPE != 1 || !transpose) return fail(-23, "wavelet: internal error (fused kernel is Haar / axis 1 only)");
    TFX_WCOLS(true, (TYPE == 1));
  } else if (transpose) {
    TFX_WCOLS(true, false);
  } else {
    TFX_WCOLS(false, false);
  }
#undef TFX_WCOLS
  ctx().launches++;
  TFX_CUDA(cudaGetLastError());
  return 0;
}

static int host_ilog2(long long n) {
  int k = 0;
  while ((1LL << (k + 1)) <= n) ++k;
  return k;
}

// One whole-volume axis pass with the column-layout kernel when the axis length fits its tile (*launched = 1).
template <int TYPE, bool FWD>
static int launch_cols(double *d_s, int L, long long inner, long long outer, cudaStream_t st, int *launched) {
  *launched = 0;
  const bool transpose = (inner == 1);
  if (cols_tile(L, transpose, inner) == 0) return 0;
  ColsGeom G = {};
  G.L = L; G.inner = inner; G.nlines = inner * outer;
  G.per = 1; G.ostride = transpose ? (long long)L : (long long)L * inner; G.rstride = transpose ? 0 : inner;
  G.nscale = host_ilog2(L);
  if ((G.nlines + 7) / 8 > 0x7fffffffLL) return 0;
  TFX_TRY((launch_cols_geom<TYPE, FWD>(d_s, G, transpose, false, 0, st)));
  *launched = 1;
  return 0;
}

static WaveConst make_consts() {
  // wavelet_transform.F90:251-255 and the sqrt(2) of the Haar normalisation (:138-139).
  WaveConst k;
  k.sq2 = sqrt(2.0);
  k.c0 = sqrt(3.0);
  k.c1 = sqrt(3.0) / 4.0;
  k.c2 = (sqrt(3.0) - 2.0) / 4.0;
  k.c3 = (sqrt(3.0) - 1.0) / sqrt(2.0);
  k.c4 = (sqrt(3.0) + 1.0) / sqrt(2.0);
  k.isq2 = 1.0 / k.sq2;
  k.half_sq2 = 0.4999 * k.sq2;
  return k;
}

template <int TYPE, bool FWD>
static int launch_axis(double *d_s, int L, long long inner, long long outer, cudaStream_t st) {
  if (L < 2) return 0;  // nscale == 0: nothing to do
  {
    int launched = 0;
    TFX_TRY((launch_cols<TYPE, FWD>(d_s, L, inner, outer, st, &launched)));
    if (launched) return 0;
  }
  const size_t kHardMax = 200 * 1024;
  // Tile budget. The passes are latency-bound (load, log2(L) block barriers, store), so for short axes many small
  // CTAs in flight win (32 KB tiles, 256 threads: measured 0.104 ms per 256x256x64 transform against 0.121 ms with
  // 64 KB / 512 threads); long axes need the bigger tile to keep >= 16 lines per CTA (512x512x128: 0.72 vs 0.87 ms).
  const size_t kBudget = (L <= 256) ? 32 * 1024 : 64 * 1024;
  const int kWThreads = (L <= 256) ? 256 : 512;
  size_t budget = std::min<size_t>(kHardMax, std::max<size_t>(kBudget, (size_t)L * 17 * sizeof(double)));
  long long max_lines = (long long)(budget / sizeof(double)) / L - 1;
  if (max_lines < 1) {
    budget = kHardMax;
    max_lines = (long long)(budget / sizeof(double)) / L - 1;
  }
  if (max_lines < 1) return fail(-20, "wavelet: axis length " + std::to_string(L) + " does not fit shared memory");
  int TI, TO;
  if (inner <= 32 && inner <= max_lines) {
    // whole inner extent in the tile: batch consecutive outer chunks (tile stays contiguous in memory)
    TI = (int)inner;
    TO = (int)std::min<long long>(outer, max_lines / TI);
  } else {
    int cap = (int)std::min<long long>(std::min<long long>(inner, max_lines), 32);
    TI = 1;
    while (TI * 2 <= cap) TI *= 2;
    TO = 1;
  }
  long long gx = (inner + TI - 1) / TI;
  // keep at least ~2 waves of CTAs when the volume allows it
  while (TO > 1 && gx * ((outer + TO - 1) / TO) < 2LL * 148) TO = (TO + 1) / 2;
  long long gy = (outer + TO - 1) / TO;
  if (gy > 65535) {
    long long per = 65535LL * TO;
    for (long long ob = 0; ob < outer; ob += per) {
      long long on = std::min(per, outer - ob);
      TFX_TRY((launch_axis<TYPE, FWD>(d_s + ob * L * inner, L, inner, on, st)));
    }
    return 0;
  }
  int NL = TI * TO;
  int pitch = NL | 1;  // odd pitch: conflict-free for 64-bit accesses in both thread mappings
  size_t smem = (size_t)L * pitch * sizeof(double);
  auto kern = wavelet_axis_kernel<TYPE, FWD>;
  TFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHardMax + 8 * 1024));
  dim3 grid((unsigned)gx, (unsigned)gy);
  kern<<<grid, smem > 100 * 1024 ? 512 : kWThreads, smem, st>>>(d_s, L, inner, outer, TI, TO, pitch, make_consts());
  ctx().launches++;
  TFX_CUDA(cudaGetLastError());
  return 0;
}

// nvol volumes stored back to back are one volume with nvol times the outer extent for every axis pass (lines never
// mix): one set of three launches for a whole batch, full waves of CTAs instead of 1.2 per small volume.
// i3-slab size of the L2-blocked axis-1 / axis-2 passes; 0 (default): whole volume per pass. Measured on B200 with the
// column kernels (gpurun_out r2_wavelet_time_e, 512x512x128 Haar): no slabs 0.405 ms, 32 MB slabs 0.537 ms, 64 MB 0.443 ms,
// 100 MB 0.441 ms -- the small launches lose more (partial waves, launch gaps) than the L2 hits win, so it stays off.
int g_opt_wavelet_slab_mb = 0;

int g_opt_wavelet_fuse12 = 1;   // 1: Haar axis-1 pass fused with the three lowest axis-2 scales (see run3d)

// Haar, axes 1 and 2 of nplanes planes (n1 x n2 each) in 1 + 1/8 (forward) or 1 + 2/8 (inverse) passes over the data
// instead of 2. The fused kernel holds NC complete axis-1 lines = NC consecutive axis-2 positions of one plane: after
// the axis-1 scales it runs the three lowest axis-2 scales on the groups of 8 columns (a group of 8 starting at a
// multiple of 8 is closed under scales 1-3). The remaining axis-2 scales only touch the rows j = 0 (mod 8) -- the
// transform of the line (j / 8) of length ceil(n2 / 8) with nscale - 3 scales -- in a pass over 1/8 of the plane.
// Every element goes through the reference's operations in the reference's order:
//   forward (wavelet_transform.F90:75-153): axis 1, axis 2 scales 1..3 | axis 2 scales 4..nscale
//   inverse (:158-236, same axis order, scales downwards): axis 1 on the rows j = 0 (mod 8) | axis 2 scales nscale..4 on
//   them | axis 1 on the other rows, then axis 2 scales 3..1 on all of them (fused kernel, skip0)
template <int TYPE, bool FWD>
static int haar_axes12_fused(double *d_s, int n1, int n2, long long nplanes, cudaStream_t st, int *launched) {
  *launched = 0;
  if (TYPE != 1 || !g_opt_wavelet_fuse12 || n2 < 2 || n1 < 2) return 0;
  if (cols_tile(n1, true, 1) == 0) return 0;
  const int ns2 = host_ilog2(n2), m2 = (n2 + 7) / 8;
  const bool high = ns2 > 3;
  if (high && cols_tile(m2, false, n1) == 0) return 0;
  ColsGeom F = {};   // fused kernel
  F.L = n1; F.inner = 1; F.nlines = (long long)n2 * nplanes; F.per = 1; F.ostride = n1; F.rstride = 0;
  F.nscale = host_ilog2(n1); F.n2 = n2; F.nscale2 = std::min(3, ns2); F.skip0 = (!FWD && high) ? 1 : 0;
  ColsGeom H = {};   // axis-2 scales 4.. on the rows j = 0 (mod 8): lines of length m2, element stride 8 * n1
  H.L = m2; H.inner = n1; H.nlines = (long long)n1 * nplanes; H.per = 1; H.ostride = (long long)n1 * n2;
  H.rstride = 8LL * n1; H.nscale = ns2 - 3;
  ColsGeom S = {};   // axis 1 on the rows j = 0 (mod 8): m2 lines per plane, 8 * n1 apart
  S.L = n1; S.inner = 1; S.nlines = (long long)m2 * nplanes; S.per = m2; S.ostride = (long long)n1 * n2;
  S.rstride = 8LL * n1; S.nscale = F.nscale;
  if (FWD) {
    TFX_TRY((launch_cols_geom<TYPE, FWD>(d_s, F, true, true, nplanes, st)));
    if (high) TFX_TRY((launch_cols_geom<TYPE, FWD>(d_s, H, false, false, 0, st)));
  } else {
    if (high) {
      TFX_TRY((launch_cols_geom<TYPE, FWD>(d_s, S, true, false, 0, st)));
      TFX_TRY((launch_cols_geom<TYPE, FWD>(d_s, H, false, false, 0, st)));
    }
    TFX_TRY((launch_cols_geom<TYPE, FWD>(d_s, F, true, true, nplanes, st)));
  }
  *launched = 1;
  return 0;
}

template <int TYPE, bool FWD>
static int run3d(double *d_s, int n1, int n2, int n3, long long nvol, cudaStream_t st) {
  const long long plane = (long long)n1 * n2, nplanes = (long long)n3 * nvol;
  const long long bytes = plane * nplanes * (long long)sizeof(double);
  long long per = nplanes;   // planes per slab
  if (g_opt_wavelet_slab_mb > 0 && bytes > 2LL * g_opt_wavelet_slab_mb * (1 << 20))
    per = std::max<long long>(1, (long long)g_opt_wavelet_slab_mb * (1 << 20) / (plane * (long long)sizeof(double)));
  for (long long k0 = 0; k0 < nplanes; k0 += per) {
    const long long nk = std::min(per, nplanes - k0);
    double *slab = d_s + k0 * plane;
    int fused = 0;
    TFX_TRY((haar_axes12_fused<TYPE, FWD>(slab, n1, n2, nk, st, &fused)));
    if (fused) continue;
    TFX_TRY((launch_axis<TYPE, FWD>(slab, n1, 1, (long long)n2 * nk, st)));
    TFX_TRY((launch_axis<TYPE, FWD>(slab, n2, n1, nk, st)));
  }
  TFX_TRY((launch_axis<TYPE, FWD>(d_s, n3, plane, nvol, st)));
  return 0;
}

// forward_wavelet / inverse_wavelet dispatch, wavelet_transform.F90:37-70.
int wavelet3d_device_batch(double *d_s, int n1, int n2, int n3, long long nvol, int wavelet_type, bool forward,
                           cudaStream_t st) {
  if (n1 < 1 || n2 < 1 || n3 < 1 || nvol < 1) return fail(-21, "wavelet: wrong grid size");
  if (wavelet_type == 1)
    return forward ? run3d<1, true>(d_s, n1, n2, n3, nvol, st) : run3d<1, false>(d_s, n1, n2, n3, nvol, st);
  if (wavelet_type == 2)
    return forward ? run3d<2, true>(d_s, n1, n2, n3, nvol, st) : run3d<2, false>(d_s, n1, n2, n3, nvol, st);
  return fail(-22, "Unknown wavelet type!");
}
int wavelet3d_device(double *d_s, int n1, int n2, int n3, int wavelet_type, bool forward, cudaStream_t st) {
  return wavelet3d_device_batch(d_s, n1, n2, n3, 1, wavelet_type, forward, st);
}

// One axis pass (all scales of that axis) over lines A[outer][L][inner]: the building block of the distributed
// transform (data.cu), which runs axes 1 / 2 on the planes a rank owns and axis 3 on its share of the k-lines.
int wavelet_axis_device(double *d_s, int L, long long inner, long long outer, int wavelet_type, bool forward,
                        cudaStream_t st) {
  if (L < 1 || inner < 1 || outer < 1) return 0;
  if (wavelet_type == 1)
    return forward ? launch_axis<1, true>(d_s, L, inner, outer, st) : launch_axis<1, false>(d_s, L, inner, outer, st);
  if (wavelet_type == 2)
    return forward ? launch_axis<2, true>(d_s, L, inner, outer, st) : launch_axis<2, false>(d_s, L, inner, outer, st);
  return fail(-22, "Unknown wavelet type!");
}

// Axes 1 and 2 of nplanes independent n1 x n2 planes stored back to back (layout A of the distributed transform): the
// fused Haar scheme when it applies, else the two axis passes.
int wavelet_axes12_device(double *d_s, int n1, int n2, long long nplanes, int wavelet_type, bool forward, cudaStream_t st) {
  if (n1 < 1 || n2 < 1 || nplanes < 1) return 0;
  if (wavelet_type != 1 && wavelet_type != 2) return fail(-22, "Unknown wavelet type!");
  int fused = 0;
  if (wavelet_type == 1) {
    if (forward) TFX_TRY((haar_axes12_fused<1, true>(d_s, n1, n2, nplanes, st, &fused)));
    else TFX_TRY((haar_axes12_fused<1, false>(d_s, n1, n2, nplanes, st, &fused)));
  }
  if (fused) return 0;
  TFX_TRY(wavelet_axis_device(d_s, n1, 1, (long long)n2 * nplanes, wavelet_type, forward, st));
  return wavelet_axis_device(d_s, n2, n1, nplanes, wavelet_type, forward, st);
}

}  // namespace tfx
