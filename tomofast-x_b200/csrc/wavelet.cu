// wavelet.cu -- 3-D separable Haar / Daubechies-D4 lifting transforms on the device.
//
// Replaces src/utils/wavelet_transform.F90 (Haar3D :75-153, iHaar3D :158-236, DaubD43D :243-367,
// iDaubD43D :374-498). The reference sweeps whole (n-1)-D slabs once per lifting step and per
// scale (3-5 passes over the volume per scale, ~25 scales for 512x512x128). Here one kernel per
// axis keeps a tile of complete lines in shared memory and runs ALL scales of that axis on it:
// 3 passes over the volume in total (16 B/element each), every global access coalesced.
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
    int ti = (int)threadIdx.x % TI, r0 = (int)threadIdx.x / TI;
    int l = r0 % L, to = r0 / L;
    for (int e = threadIdx.x; e < total_i; e += nt) {
      if (ti < ti_n && to < to_n)
        tile[(size_t)l * pitch + to * TI + ti] = s[((o0 + to) * L + l) * inner + (i0 + ti)];
      ti += d_ti;
      int cr = 0;
      if (ti >= TI) { ti -= TI; cr = 1; }
      l += d_l + cr;
      int cl = 0;
      if (l >= L) { l -= L; cl = 1; }
      to += d_to + cl;
    }
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

static WaveConst make_consts() {
  // wavelet_transform.F90:251-255 and the sqrt(2) of the Haar normalisation (:138-139).
  WaveConst k;
  k.sq2 = sqrt(2.0);
  k.c0 = sqrt(3.0);
  k.c1 = sqrt(3.0) / 4.0;
  k.c2 = (sqrt(3.0) - 2.0) / 4.0;
  k.c3 = (sqrt(3.0) - 1.0) / sqrt(2.0);
  k.c4 = (sqrt(3.0) + 1.0) / sqrt(2.0);
  return k;
}

template <int TYPE, bool FWD>
static int launch_axis(double *d_s, int L, long long inner, long long outer, cudaStream_t st) {
  if (L < 2) return 0;  // nscale == 0: nothing to do
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
template <int TYPE, bool FWD>
static int run3d(double *d_s, int n1, int n2, int n3, long long nvol, cudaStream_t st) {
  TFX_TRY((launch_axis<TYPE, FWD>(d_s, n1, 1, (long long)n2 * n3 * nvol, st)));
  TFX_TRY((launch_axis<TYPE, FWD>(d_s, n2, n1, (long long)n3 * nvol, st)));
  TFX_TRY((launch_axis<TYPE, FWD>(d_s, n3, (long long)n1 * n2, nvol, st)));
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

}  // namespace tfx
