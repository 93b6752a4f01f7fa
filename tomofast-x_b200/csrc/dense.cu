// dense.cu -- single-pass sweep over an uncompressed (dense, column-major f32) sensitivity block.
//
// Reference path: the two products per LSQR iteration, `v = v + S^T u` (lsqr_solver2.F90:228,236 ->
// sparse_matrix.f90:388-405) and `u = u + S v` (lsqr_solver2.F90:209 -> sparse_matrix.f90:313-329),
// each streaming the whole matrix (8 B/nnz as CSR). Without compression every row holds columns
// 1..N (sensitivity_gravmag.F90:288-295), so the block is stored as bare f32 values (4 B/nnz).
//
// B200 design: the number of data rows is small (<= ~10^4) while the number of columns is huge, so
// u and the accumulators of S*vhat live in REGISTERS (20 rows per thread, 512 threads per CTA, one
// CTA per SM), and S is stored column-major so that one column is one contiguous 16 B-aligned
// burst. A column is staged ONCE in shared memory by TMA (cp.async.bulk + mbarrier ring) and used
// twice while it is there:
//     t_j    = sum_i S_ij u_i                       (transposed product, block-wide tree reduction)
//     vhat_j = -beta v_j + t_j + g_j                (LSQR's v update, un-normalised)
//     q_i   += S_ij vhat_j                          (forward product of the NEXT iteration)
// By linearity S*(vhat/alpha) = (S*vhat)/alpha, so the normalisation of v by alpha = |vhat| (only
// known after the sweep) is applied afterwards to the short vector q. One LSQR iteration therefore
// reads S exactly once: 4 B/nnz of HBM traffic instead of the reference's 16 B/nnz.
//
// Instruction budget (ncu, round 1, profiles/r1_dense_sweep_v4_*): the TMA ring alone streams at 7.3 TB/s, the
// kernel is bound by instruction issue and the MIO path (LDS, shuffles, F2F), so every design choice removes
// instructions per matrix entry (11.4 now): one LDS.128 per 4 rows; two columns per block barrier with one
// transposed shuffle reduction; the f32->f64 conversion of the second use split between F2F (XU pipe) and an
// integer rebuild of the double from the f32 bits (ALU/FMA pipes); ring slots laid out so that every thread reads
// its rows without bounds predicates.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace tfx {

static const int kThreads = 1024;
static const int kMaxSlots = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_WAIT;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE_WAIT:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier, L2 evict-first (the
// matrix is streamed once per sweep and must not displace the vectors).
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t make_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async_8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// f32 -> f64 without the XU pipe: rebias the exponent and spread the mantissa with integer ops.
// Exact for every normal float; zeros / subnormals (|x| < 1.2e-38) are NOT handled -- the host only
// selects this path for blocks that contain neither (DenseCM::fastcvt_ok).
__device__ __forceinline__ double f32bits_to_f64(uint32_t b) {
  const uint32_t hi = ((uint32_t)((int32_t)b >> 3) & 0x8FFFFFFFu) + 0x38000000u;
  const uint32_t lo = b << 29;
  return __hiloint2double((int)hi, (int)lo);
}

struct DenseArgs {
  const float *S;
  long long ld;
  int nrows, ncols, col0;
  const double *u, *v, *g;
  double *out;
  const double *nbeta;
  double *partial_q, *partial_n2;
  int ns;
  unsigned col_bytes;    // bytes of one column in global memory and in a ring slot (ld * 4)
  unsigned ring_bytes;   // ns * col_bytes + zeroed guard so that row index K*1024-1 is always readable
  const int *done;
  int accumulate;        // DENSE_T_ONLY: out[j] += t_j (row blocks after the first one) instead of out[j] = t_j
};

// Sum of `a` over the warp for column 0 and of `b` for column 1 with ONE transposed butterfly:
// on return even lanes hold sum(a), odd lanes hold sum(b).
__device__ __forceinline__ double warp_sum2(double a, double b, int lane) {
  const bool odd = lane & 1;
  double keep = odd ? b : a;
  const double send = odd ? a : b;
  keep += __shfl_xor_sync(0xffffffffu, send, 1);
#pragma unroll
  for (int o = 2; o <= 16; o <<= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
  return keep;
}

// Row ownership: thread t owns rows VW*(t + NT*m) + h, h < VW, m < KV (VW consecutive floats = one LDS.64 / LDS.128).
//   <NT = 1024, VW = 2>: 32 warps, 10 rows per thread at 10^4 rows (64 registers per thread)
//   <NT =  512, VW = 4>: 16 warps, 20 rows per thread (128 registers): half the shared-memory load instructions and
//                        half the shuffle reductions per matrix entry -- the kernel is issue/MIO-bound, not HBM-bound,
//                        once the SM clock drops under the power cap.
template <int VW>
struct RowVec;
template <>
struct RowVec<2> {
  typedef float2 type;
  static __device__ __forceinline__ void get(const float2 &f, float (&o)[2]) { o[0] = f.x; o[1] = f.y; }
};
template <>
struct RowVec<4> {
  typedef float4 type;
  static __device__ __forceinline__ void get(const float4 &f, float (&o)[4]) { o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = f.w; }
};

template <int KV, int MODE, int FM, int NT, int VW>
__global__ void __launch_bounds__(NT, 1) dense_sweep_kernel(DenseArgs a) {
  if (a.done && *a.done) return;
  typedef typename RowVec<VW>::type vec_t;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *ring = smem;
  uint64_t *full = (uint64_t *)(smem + a.ring_bytes);
  double *red = (double *)(full + kMaxSlots);   // [2 buffers][2 columns][32 warps]
  double *vq = red + 128;                       // [kMaxSlots]
  double *gq = vq + kMaxSlots;                  // [kMaxSlots]

  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int c_lo = (int)((long long)a.ncols * blockIdx.x / gridDim.x);
  const int c_hi = (int)((long long)a.ncols * (blockIdx.x + 1) / gridDim.x);
  const int ncl = c_hi - c_lo;
  const int ns = a.ns;
  const bool has_g = (MODE == DENSE_FUSED) && (a.g != nullptr);

  // Every word of the ring (+ guard) must always hold a finite float: rows beyond nrows are read
  // unpredicated (their u is 0 and their accumulators are never stored).
  for (unsigned i = t * 16u; i < a.ring_bytes; i += NT * 16u) *(uint4 *)(ring + i) = make_uint4(0, 0, 0, 0);
  if (t < 128) red[t] = 0.0;                    // warps that do not exist contribute 0 to the second stage

  double ur[VW * KV], acc[VW * KV];
#pragma unroll
  for (int m = 0; m < KV; ++m) {
#pragma unroll
    for (int h = 0; h < VW; ++h) {
      const int row = VW * (t + NT * m) + h;
      ur[VW * m + h] = (MODE != DENSE_F_ONLY && row < a.nrows) ? a.u[row] : 0.0;
      acc[VW * m + h] = 0.0;
    }
  }
  const double nbeta = (MODE == DENSE_FUSED) ? *a.nbeta : 0.0;
  double n2 = 0.0;
  uint64_t policy = 0;
  __syncthreads();

  if (t == 0) {
    for (int s = 0; s < ns; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    policy = make_evict_first_policy();
    const int npro = min(ns, ncl);
    for (int s = 0; s < npro; ++s) {
      mbar_expect_tx(&full[s], a.col_bytes);
      tma_load_1d(ring + (size_t)s * a.col_bytes, a.S + (long long)(c_lo + s) * a.ld, a.col_bytes, &full[s], policy);
      if (MODE != DENSE_T_ONLY) {
        cp_async_8(&vq[s], a.v + a.col0 + c_lo + s);
        if (has_g) cp_async_8(&gq[s], a.g + a.col0 + c_lo + s);
      }
    }
    cp_async_commit();
  }
  __syncthreads();

  // Two columns per step: one block barrier and one pair of transposed reductions per step.
  // (s0, ph0) = ring slot and mbarrier parity of column j0, advanced incrementally (no divisions).
  int s0 = 0;
  uint32_t ph0 = 0;
  int sp0 = 0, sp1 = 0;   // slots of the previous step (refilled after this step's barrier)
  for (int j0 = 0; j0 < ncl; j0 += 2) {
    const bool two = (j0 + 1 < ncl);
    int s1 = s0 + 1;
    uint32_t ph1 = ph0;
    if (s1 == ns) { s1 = 0; ph1 ^= 1u; }
    const vec_t *col0p = (const vec_t *)(ring + (size_t)s0 * a.col_bytes) + t;
    const vec_t *col1p = (const vec_t *)(ring + (size_t)s1 * a.col_bytes) + t;
    mbar_wait(&full[s0], ph0);
    if (two) mbar_wait(&full[s1], ph1);
    const int buf = (j0 >> 1) & 1;

    if (MODE != DENSE_F_ONLY && MODE != 3) {
      // ---- transposed product: partial dots of this thread's rows (F2F conversions, XU pipe); two chains per column
      double p0 = 0.0, p1 = 0.0, p0b = 0.0, p1b = 0.0;
#pragma unroll
      for (int m = 0; m < KV; ++m) {
        float f[VW];
        RowVec<VW>::get(col0p[NT * m], f);
#pragma unroll
        for (int h = 0; h < VW; h += 2) {
          p0 = fma((double)f[h], ur[VW * m + h], p0);
          p0b = fma((double)f[h + 1], ur[VW * m + h + 1], p0b);
        }
      }
      if (two) {
#pragma unroll
        for (int m = 0; m < KV; ++m) {
          float f[VW];
          RowVec<VW>::get(col1p[NT * m], f);
#pragma unroll
          for (int h = 0; h < VW; h += 2) {
            p1 = fma((double)f[h], ur[VW * m + h], p1);
            p1b = fma((double)f[h + 1], ur[VW * m + h + 1], p1b);
          }
        }
      }
      const double r = warp_sum2(p0 + p0b, p1 + p1b, lane);
      if (lane < 2) red[(buf * 2 + lane) * 32 + wid] = r;
    }
    if (t == 0 && MODE != DENSE_T_ONLY) cp_async_wait_all();   // v_j / g_j were requested >= 1 step ago
    __syncthreads();

    // Every thread has finished with the slots of the previous step: refill them.
    if (t == 0 && j0 >= 2) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int jn = j0 - 2 + c + ns;
        const int sn = c ? sp1 : sp0;
        if (jn < ncl) {
          mbar_expect_tx(&full[sn], a.col_bytes);
          tma_load_1d(ring + (size_t)sn * a.col_bytes, a.S + (long long)(c_lo + jn) * a.ld, a.col_bytes, &full[sn], policy);
          if (MODE != DENSE_T_ONLY) {
            cp_async_8(&vq[sn], a.v + a.col0 + c_lo + jn);
            if (has_g) cp_async_8(&gq[sn], a.g + a.col0 + c_lo + jn);
          }
        }
      }
      cp_async_commit();
    }

    double x0, x1 = 0.0;
    if (MODE == DENSE_F_ONLY || MODE == 3) {
      x0 = vq[s0];
      if (two) x1 = vq[s1];
    } else {
      const double r = warp_sum2(red[(buf * 2 + 0) * 32 + lane], red[(buf * 2 + 1) * 32 + lane], lane);
      const double t0 = __shfl_sync(0xffffffffu, r, 0), t1 = __shfl_sync(0xffffffffu, r, 1);
      if (MODE == DENSE_FUSED) {
        x0 = fma(nbeta, vq[s0], t0);             // v = -beta v ; v = v + S^T u   (lsqr_solver2.F90:225,236)
        if (has_g) x0 += gq[s0];                 //                + C^T u_cons  (:238)
        n2 = fma(x0, x0, n2);
        if (two) {
          x1 = fma(nbeta, vq[s1], t1);
          if (has_g) x1 += gq[s1];
          n2 = fma(x1, x1, n2);
        }
      } else {
        x0 = t0;
        x1 = t1;
      }
      if (t == 0) {
        double *o = a.out + a.col0 + c_lo + j0;
        if (MODE == DENSE_T_ONLY && a.accumulate) {
          o[0] += x0;
          if (two) o[1] += x1;
        } else {
          o[0] = x0;
          if (two) o[1] = x1;
        }
      }
    }

    if (MODE != DENSE_T_ONLY && MODE != 3) {
      // ---- forward product with the columns that are still in shared memory
#pragma unroll
      for (int m = 0; m < KV; ++m) {
        float f[VW];
        RowVec<VW>::get(col0p[NT * m], f);
#pragma unroll
        for (int h = 0; h < VW; ++h) {
          const double d = (m >= FM) ? f32bits_to_f64(__float_as_uint(f[h])) : (double)f[h];
          acc[VW * m + h] = fma(d, x0, acc[VW * m + h]);
        }
      }
      if (two) {
#pragma unroll
        for (int m = 0; m < KV; ++m) {
          float f[VW];
          RowVec<VW>::get(col1p[NT * m], f);
#pragma unroll
          for (int h = 0; h < VW; ++h) {
            const double d = (m >= FM) ? f32bits_to_f64(__float_as_uint(f[h])) : (double)f[h];
            acc[VW * m + h] = fma(d, x1, acc[VW * m + h]);
          }
        }
      }
    }
    // advance to the next pair of columns
    sp0 = s0; sp1 = s1;
    s0 = s1 + 1; ph0 = ph1;
    if (s0 == ns) { s0 = 0; ph0 ^= 1u; }
  }

  if (MODE != DENSE_T_ONLY) {
#pragma unroll
    for (int m = 0; m < KV; ++m) {
#pragma unroll
      for (int h = 0; h < VW; ++h) {
        const int row = VW * (t + NT * m) + h;
        if (row < a.nrows) a.partial_q[(long long)blockIdx.x * a.ld + row] = acc[VW * m + h];
      }
    }
  }
  if (MODE == DENSE_FUSED && t == 0) a.partial_n2[blockIdx.x] = n2;
}

// q[row] = sum_b partial_q[b][row] (b ascending), n2 = sum_b partial_n2[b].
__global__ void __launch_bounds__(256) dense_reduce_kernel(const double *partial_q, long long ld, int nblocks, int nrows,
                                                           double *q, const double *partial_n2, double *n2,
                                                           const int *done) {
  if (done && *done) return;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (q != nullptr && row < nrows) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial_q[(long long)b * ld + row];
    q[row] = s;
  }
  if (n2 != nullptr && row == 0) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial_n2[b];
    *n2 = s;
  }
}

// Flags a block that contains zeros or subnormal floats (the integer f32->f64 path cannot convert them).
__global__ void __launch_bounds__(256) dense_scan_kernel(const float *__restrict__ S, long long ld, int nrows,
                                                         long long ncols, int *flag) {
  // one column per CTA step, float4 loads (ld % 4 == 0, columns are 16-byte aligned); rows >= nrows are padding
  int bad = 0;
  const int nv = nrows >> 2;
  for (long long j = blockIdx.x; j < ncols; j += gridDim.x) {
    const float4 *col = (const float4 *)(S + j * ld);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const float4 f = col[i];
      const uint32_t m = 0x7F800000u;
      if (!(__float_as_uint(f.x) & m) || !(__float_as_uint(f.y) & m) || !(__float_as_uint(f.z) & m) ||
          !(__float_as_uint(f.w) & m))
        bad = 1;
    }
    for (int i = 4 * nv + threadIdx.x; i < nrows; i += blockDim.x)
      if ((__float_as_uint(S[j * ld + i]) & 0x7F800000u) == 0u) bad = 1;
  }
  if (bad) atomicExch(flag, 1);
}

int dense_scan_fastcvt(DenseCM &S, cudaStream_t st) {
  DevBuf<int> flag;
  TFX_TRY(flag.alloc(1));
  TFX_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
  dense_scan_kernel<<<ctx().num_sms * 8, 256, 0, st>>>(S.val.p, S.ld, S.nrows, S.ncols, flag.p);
  ctx().launches++;
  int h = 0;
  TFX_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  S.fastcvt_ok = (h == 0) ? 1 : 0;
  return 0;
}

int g_opt_dense_stream_only = 0;

// FM = number of row vectors per thread (of KV) whose SECOND use converts with F2F (XU pipe) instead of the integer
// path (ALU/FMA pipes, 4 instructions per entry): FM >= KV is the all-F2F kernel that is also correct for blocks
// with zeros / subnormals.
template <int K, int FM, int NT, int VW>
static int launch_k(DenseMode mode, const DenseArgs &a, int grid, size_t smem, cudaStream_t st) {
  if (g_opt_dense_stream_only && mode == DENSE_FUSED) {
    // diagnostic: the TMA ring, the barriers and the refills without the two products (results are meaningless):
    // the streaming ceiling of this access pattern, to tell an HBM-side bound from an SM-side one
    auto k = dense_sweep_kernel<K, 3, 99, NT, VW>;
    TFX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    k<<<grid, NT, smem, st>>>(a);
    return 0;
  }
  switch (mode) {
    case DENSE_FUSED: {
      auto k = dense_sweep_kernel<K, DENSE_FUSED, FM, NT, VW>;
      TFX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      k<<<grid, NT, smem, st>>>(a);
      break;
    }
    case DENSE_T_ONLY: {
      auto k = dense_sweep_kernel<K, DENSE_T_ONLY, 99, NT, VW>;
      TFX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      k<<<grid, NT, smem, st>>>(a);
      break;
    }
    default: {
      auto k = dense_sweep_kernel<K, DENSE_F_ONLY, FM, NT, VW>;
      TFX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      k<<<grid, NT, smem, st>>>(a);
      break;
    }
  }
  return 0;
}

// Option "dense_f2f_rows": row vectors per thread converted with F2F on their second use (0, 2 or 99 = all).
int g_opt_dense_f2f_rows = 2;

template <int K, int NT, int VW>
static int launch_kf(DenseMode mode, const DenseArgs &a, int grid, size_t smem, bool fast, cudaStream_t st) {
  if (!fast || g_opt_dense_f2f_rows >= K) return launch_k<K, 99, NT, VW>(mode, a, grid, smem, st);
  if (g_opt_dense_f2f_rows >= 2 && K > 2) return launch_k<K, 2, NT, VW>(mode, a, grid, smem, st);
  return launch_k<K, 0, NT, VW>(mode, a, grid, smem, st);
}

// Option "dense_vec4": 1 = 512 threads x float4 rows (default), 0 = 1024 threads x float2 rows.
int g_opt_dense_vec4 = 1;

int dense_sweep(DenseCM &S, DenseMode mode, const double *d_u, const double *d_v, const double *d_g, double *d_out,
                const double *d_nbeta, double *d_q, double *d_n2, const int *d_done, cudaStream_t st, bool accumulate) {
  Context &c = ctx();
  if (S.empty()) return 0;
  if (S.nrows > kDenseMaxRows)
    return fail(-30, "dense sweep: more than " + std::to_string(kDenseMaxRows) + " data rows per block is not supported yet");
  if (S.fastcvt_ok < 0) TFX_TRY(dense_scan_fastcvt(S, st));
  const bool vec4 = g_opt_dense_vec4 != 0;
  // dense_vec4 = 2: 640 threads x float4 rows (20 warps per SM instead of 16, 16 rows per thread at 10^4 rows)
  const bool wide = g_opt_dense_vec4 == 2 && S.nrows > 4 * 512 * 3;
  const int NT = wide ? 640 : (vec4 ? 512 : kThreads), VW = vec4 ? 4 : 2;
  const int KV = (S.nrows + VW * NT - 1) / (VW * NT);                       // row vectors per thread
  const int K = VW * KV;
  const unsigned col_bytes = (unsigned)(S.ld * sizeof(float));
  const size_t span = (size_t)K * NT * sizeof(float);                       // bytes a thread block may read per slot
  const size_t guard = (span > col_bytes) ? ((span - col_bytes + 15) / 16 * 16) : 0;
  const size_t tail = kMaxSlots * sizeof(uint64_t) + (128 + 2 * kMaxSlots) * sizeof(double);
  const size_t budget = 227 * 1024 - 256;
  if (budget < tail + guard + 3 * (size_t)col_bytes) return fail(-31, "dense sweep: column does not fit the shared-memory ring");
  int ns = (int)std::min<size_t>(kMaxSlots, (budget - tail - guard) / col_bytes);
  const size_t ring_bytes = (size_t)ns * col_bytes + guard;
  const size_t smem = ring_bytes + tail;
  if (S.grid <= 0) {
    S.grid = std::min(c.num_sms, S.ncols);
    TFX_TRY(S.partial_q.alloc((size_t)S.grid * S.ld));
    TFX_TRY(S.partial_n2.alloc((size_t)S.grid));
  }
  DenseArgs a;
  a.S = S.val.p; a.ld = S.ld; a.nrows = S.nrows; a.ncols = S.ncols; a.col0 = S.col0;
  a.u = d_u; a.v = d_v; a.g = d_g; a.out = d_out; a.nbeta = d_nbeta;
  a.partial_q = S.partial_q.p; a.partial_n2 = S.partial_n2.p;
  a.ns = ns; a.col_bytes = col_bytes; a.ring_bytes = (unsigned)ring_bytes; a.done = d_done;
  a.accumulate = (mode == DENSE_T_ONLY && accumulate) ? 1 : 0;
  const bool fast = S.fastcvt_ok == 1;
  int rc;
  if (wide) {
    switch (KV) {
      case 3: rc = launch_kf<3, 640, 4>(mode, a, S.grid, smem, fast, st); break;
      default: rc = launch_kf<4, 640, 4>(mode, a, S.grid, smem, fast, st); break;
    }
  } else if (vec4) {
    switch (KV) {
      case 1: rc = launch_kf<1, 512, 4>(mode, a, S.grid, smem, fast, st); break;
      case 2: rc = launch_kf<2, 512, 4>(mode, a, S.grid, smem, fast, st); break;
      case 3: rc = launch_kf<3, 512, 4>(mode, a, S.grid, smem, fast, st); break;
      case 4: rc = launch_kf<4, 512, 4>(mode, a, S.grid, smem, fast, st); break;
      default: rc = launch_kf<5, 512, 4>(mode, a, S.grid, smem, fast, st); break;
    }
  } else {
    switch (KV) {
      case 1: rc = launch_kf<1, 1024, 2>(mode, a, S.grid, smem, fast, st); break;
      case 2: rc = launch_kf<2, 1024, 2>(mode, a, S.grid, smem, fast, st); break;
      case 3: rc = launch_kf<3, 1024, 2>(mode, a, S.grid, smem, fast, st); break;
      case 4: rc = launch_kf<4, 1024, 2>(mode, a, S.grid, smem, fast, st); break;
      default: rc = launch_kf<5, 1024, 2>(mode, a, S.grid, smem, fast, st); break;
    }
  }
  TFX_TRY(rc);
  c.launches++;
  if (mode != DENSE_T_ONLY) {
    int blocks = (S.nrows + 255) / 256;
    dense_reduce_kernel<<<blocks, 256, 0, st>>>(S.partial_q.p, S.ld, S.grid, S.nrows, d_q, S.partial_n2.p,
                                                mode == DENSE_FUSED ? d_n2 : nullptr, d_done);
    c.launches++;
  }
  TFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tfx
