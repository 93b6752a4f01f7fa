// dense.cu -- single-pass sweep over an uncompressed (dense, column-major f32) sensitivity block.
//
// Reference path: the two products per LSQR iteration, `v = v + S^T u` (lsqr_solver2.F90:228,236 ->
// sparse_matrix.f90:388-405) and `u = u + S v` (lsqr_solver2.F90:209 -> sparse_matrix.f90:313-329),
// each streaming the whole matrix (8 B/nnz as CSR). Without compression every row holds columns
// 1..N (sensitivity_gravmag.F90:288-295), so the block is stored as bare f32 values (4 B/nnz).
//
// B200 design: the number of data rows is small (<= ~10^4) while the number of columns is huge, so
// u and the accumulators of S*vhat live in REGISTERS (10 rows per thread, 1024 threads per CTA, one
// CTA per SM), and S is stored column-major so that one column is one contiguous 16 B-aligned
// burst. A column is staged ONCE in shared memory by TMA (cp.async.bulk + mbarrier ring) and used
// twice while it is there:
//     t_j    = sum_i S_ij u_i                       (transposed product, block-wide tree reduction)
//     vhat_j = -beta v_j + t_j + g_j                (LSQR's v update, un-normalised)
//     q_i   += S_ij vhat_j                          (forward product of the NEXT iteration)
// By linearity S*(vhat/alpha) = (S*vhat)/alpha, so the normalisation of v by alpha = |vhat| (only
// known after the sweep) is applied afterwards to the short vector q. One LSQR iteration therefore
// reads S exactly once: 4 B/nnz of HBM traffic instead of the reference's 16 B/nnz.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace tfx {

static const int kThreads = 1024;

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
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

struct DenseArgs {
  const float *S;
  long long ld;
  int nrows, ncols, col0;
  const double *u, *v, *g;
  double *out;
  const double *nbeta;
  double *partial_q, *partial_n2;
  int ns;
  unsigned col_bytes;
  const int *done;
};

template <int K, int MODE>
__global__ void __launch_bounds__(kThreads, 1) dense_sweep_kernel(DenseArgs a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *ring = smem;
  uint64_t *full = (uint64_t *)(smem + (size_t)a.ns * a.col_bytes);
  double *red = (double *)(full + 8);   // [2][32]
  double *vq = red + 64;                // [8]
  double *gq = vq + 8;                  // [8]

  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int c_lo = (int)((long long)a.ncols * blockIdx.x / gridDim.x);
  const int c_hi = (int)((long long)a.ncols * (blockIdx.x + 1) / gridDim.x);
  const int ncl = c_hi - c_lo;
  const int ns = a.ns;
  const bool has_g = (MODE == DENSE_FUSED) && (a.g != nullptr);

  double ur[K], acc[K];
#pragma unroll
  for (int m = 0; m < K; ++m) {
    const int row = t + kThreads * m;
    ur[m] = (MODE != DENSE_F_ONLY && row < a.nrows) ? a.u[row] : 0.0;
    acc[m] = 0.0;
  }
  const double nbeta = (MODE == DENSE_FUSED) ? *a.nbeta : 0.0;
  double n2 = 0.0;
  uint64_t policy = 0;

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
      cp_async_commit();
    }
  }
  __syncthreads();

  for (int j = 0; j < ncl; ++j) {
    const int s = j % ns;
    const uint32_t parity = (uint32_t)((j / ns) & 1);
    const float *col = (const float *)(ring + (size_t)s * a.col_bytes);
    mbar_wait(&full[s], parity);

    double tj = 0.0;
    if (MODE != DENSE_F_ONLY) {
      // ---- transposed product: partial dot of this thread's rows, then fixed-order tree reduction
      double p = 0.0;
#pragma unroll
      for (int m = 0; m < K; ++m) {
        const int row = t + kThreads * m;
        const float f = (row < a.nrows) ? col[row] : 0.0f;
        p = fma((double)f, ur[m], p);
      }
      p = warp_sum(p);
      if (lane == 0) red[(j & 1) * 32 + wid] = p;
    }
    if (t == 0 && MODE != DENSE_T_ONLY) cp_async_wait_1();   // v_j / g_j prefetched >= 1 iteration ago
    __syncthreads();

    // Every thread has finished with the slot of column j-1: refill it.
    if (t == 0 && j >= 1) {
      const int jn = j - 1 + ns;
      if (jn < ncl) {
        const int sn = (j - 1) % ns;
        mbar_expect_tx(&full[sn], a.col_bytes);
        tma_load_1d(ring + (size_t)sn * a.col_bytes, a.S + (long long)(c_lo + jn) * a.ld, a.col_bytes, &full[sn], policy);
        if (MODE != DENSE_T_ONLY) {
          cp_async_8(&vq[sn], a.v + a.col0 + c_lo + jn);
          if (has_g) cp_async_8(&gq[sn], a.g + a.col0 + c_lo + jn);
        }
      }
      cp_async_commit();   // (possibly empty) group keeps the wait_group accounting uniform
    }

    double xj;
    if (MODE == DENSE_F_ONLY) {
      xj = vq[s];
    } else {
      tj = warp_sum(red[(j & 1) * 32 + lane]);
      if (MODE == DENSE_FUSED) {
        xj = fma(nbeta, vq[s], tj);            // v = -beta v ; v = v + S^T u   (lsqr_solver2.F90:225,236)
        if (has_g) xj += gq[s];                //                + C^T u_cons  (:238)
        n2 = fma(xj, xj, n2);
      } else {
        xj = tj;
      }
      if (t == 0) a.out[a.col0 + c_lo + j] = xj;
    }

    if (MODE != DENSE_T_ONLY) {
      // ---- forward product with the column that is still in shared memory
#pragma unroll
      for (int m = 0; m < K; ++m) {
        const int row = t + kThreads * m;
        const float f = (row < a.nrows) ? col[row] : 0.0f;
        acc[m] = fma((double)f, xj, acc[m]);
      }
    }
  }

  if (MODE != DENSE_T_ONLY) {
#pragma unroll
    for (int m = 0; m < K; ++m) {
      const int row = t + kThreads * m;
      if (row < a.nrows) a.partial_q[(long long)blockIdx.x * a.ld + row] = acc[m];
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

template <int K>
static int launch_k(DenseMode mode, const DenseArgs &a, int grid, size_t smem, cudaStream_t st) {
  switch (mode) {
    case DENSE_FUSED: {
      auto k = dense_sweep_kernel<K, DENSE_FUSED>;
      TFX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      k<<<grid, kThreads, smem, st>>>(a);
      break;
    }
    case DENSE_T_ONLY: {
      auto k = dense_sweep_kernel<K, DENSE_T_ONLY>;
      TFX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      k<<<grid, kThreads, smem, st>>>(a);
      break;
    }
    default: {
      auto k = dense_sweep_kernel<K, DENSE_F_ONLY>;
      TFX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      k<<<grid, kThreads, smem, st>>>(a);
      break;
    }
  }
  return 0;
}

int dense_sweep(DenseCM &S, DenseMode mode, const double *d_u, const double *d_v, const double *d_g, double *d_out,
                const double *d_nbeta, double *d_q, double *d_n2, const int *d_done, cudaStream_t st) {
  Context &c = ctx();
  if (S.empty()) return 0;
  if (S.nrows > kDenseMaxRows)
    return fail(-30, "dense sweep: more than " + std::to_string(kDenseMaxRows) + " data rows per block is not supported yet");
  const unsigned col_bytes = (unsigned)(S.ld * sizeof(float));
  const size_t tail = 8 * sizeof(uint64_t) + (64 + 16) * sizeof(double);
  const size_t budget = 227 * 1024 - 1024;
  int ns = (int)std::min<size_t>(8, (budget - tail) / col_bytes);
  if (ns < 3) return fail(-31, "dense sweep: column does not fit the shared-memory ring");
  const size_t smem = (size_t)ns * col_bytes + tail;
  if (S.grid <= 0) {
    S.grid = std::min(c.num_sms, S.ncols);
    TFX_TRY(S.partial_q.alloc((size_t)S.grid * S.ld));
    TFX_TRY(S.partial_n2.alloc((size_t)S.grid));
  }
  DenseArgs a;
  a.S = S.val.p; a.ld = S.ld; a.nrows = S.nrows; a.ncols = S.ncols; a.col0 = S.col0;
  a.u = d_u; a.v = d_v; a.g = d_g; a.out = d_out; a.nbeta = d_nbeta;
  a.partial_q = S.partial_q.p; a.partial_n2 = S.partial_n2.p;
  a.ns = ns; a.col_bytes = col_bytes; a.done = d_done;
  const int K = (S.nrows + kThreads - 1) / kThreads;
  int rc;
  switch (K) {
    case 1: rc = launch_k<1>(mode, a, S.grid, smem, st); break;
    case 2: rc = launch_k<2>(mode, a, S.grid, smem, st); break;
    case 3: rc = launch_k<3>(mode, a, S.grid, smem, st); break;
    case 4: rc = launch_k<4>(mode, a, S.grid, smem, st); break;
    case 5: rc = launch_k<5>(mode, a, S.grid, smem, st); break;
    case 6: rc = launch_k<6>(mode, a, S.grid, smem, st); break;
    case 7: rc = launch_k<7>(mode, a, S.grid, smem, st); break;
    case 8: rc = launch_k<8>(mode, a, S.grid, smem, st); break;
    case 9: rc = launch_k<9>(mode, a, S.grid, smem, st); break;
    default: rc = launch_k<10>(mode, a, S.grid, smem, st); break;
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
