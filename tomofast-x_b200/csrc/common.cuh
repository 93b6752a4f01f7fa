// common.cuh -- shared declarations for libtfx (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <exception>
#include <string>

namespace tfx {

// ---------------------------------------------------------------------------------------------
// Error handling. The reference convention is fatal-abort with a message (exit_MPI,
// src/utils/mpi_tools.F90:30-54); the C ABI returns a non-zero code and keeps the message for
// tfx_last_error(); the Fortran shim turns it into exit_MPI.
// ---------------------------------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define TFX_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      return ::tfx::fail(-100, std::string("CUDA error: ") + cudaGetErrorString(_e) +       \
                                   " at " __FILE__ ":" + std::to_string(__LINE__));         \
    }                                                                                       \
  } while (0)

// Thrust / CUB signal failures (typically cudaErrorMemoryAllocation of their scratch space) with C++ exceptions; no
// exception may cross the C ABI (a Fortran / C host would abort without the library's message).
#define TFX_THRUST(stmt)                                                                           \
  do {                                                                                             \
    try {                                                                                          \
      stmt;                                                                                        \
    } catch (const std::exception &_ex) {                                                          \
      cudaGetLastError();                                                                          \
      return ::tfx::fail(-101, std::string("device scratch allocation failed (thrust): ") + _ex.what()); \
    }                                                                                              \
  } while (0)

#define TFX_TRY(expr)              \
  do {                             \
    int _rc = (expr);              \
    if (_rc != 0) return _rc;      \
  } while (0)

// Process-wide context: one process drives one GPU (one rank per GPU, like one MPI rank).
struct Context {
  int device = -1;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  // Two side streams + events: independent row blocks of a forward product are issued alternately on them so that the
  // tail of one block's kernel (last CTAs still walking their tiles) overlaps the start of the next block's.
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  bool ready = false;
  // Kernel launch counter (bench.py's "gpu_launches" claim).
  unsigned long long launches = 0;
};
Context &ctx();
int ensure_init();

inline bool is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// RAII device buffer.
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  int alloc(size_t count) {
    if (count <= n && p) return 0;
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(-101, std::string("cudaMalloc failed (") + std::to_string(count * sizeof(T)) +
                            " bytes): " + cudaGetErrorString(e));
    }
    n = count;
    return 0;
  }
  int zero() {
    if (p) TFX_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx().stream));
    return 0;
  }
};

// Vector that may live on the host (caller-owned, copied in/out) or already on the device.
// Host vectors are staged through a small pool of device buffers that survive the call, so a host
// application that calls the solver once per major iteration does not pay cudaMalloc/cudaFree each time.
struct VecIO {
  double *dev = nullptr;       // device pointer to use
  double *host = nullptr;      // original host pointer (nullptr when caller passed device memory)
  double *stage = nullptr;     // pooled staging buffer when host != nullptr
  size_t stage_cap = 0;
  size_t n = 0;
  VecIO() {}
  VecIO(const VecIO &) = delete;
  VecIO &operator=(const VecIO &) = delete;
  ~VecIO();
  int bind(double *ptr, size_t count, bool copy_in);
  int copy_back();
};

// ---------------------------------------------------------------------------------------------
// LSQR scalar state kept on the device (no host round trip inside the loop).
// Mirrors the locals of lsqr_solve_sensit (src/inversion/lsqr_solver2.F90:66-69).
// ---------------------------------------------------------------------------------------------
struct LsqrScalars {
  double alpha, beta, rho, rhobar, phi, phibar, theta;
  double b1, c, r, s, t1, t2;
  double inv_alpha;     // 1/alpha of the last normalisation of v (1 when alpha == 0: v is left as is)
  double inv_beta;      // 1/beta of the last normalisation of u (1 when beta == 0)
  double neg_alpha;     // -alpha
  double neg_beta;      // -beta
  double misfit;
  int iter;             // next iteration number (1-based, like the reference's `iter`)
  int done;             // loop finished (niter / rmin / rho == 0 / small rhobar / misfit / |b| == 0)
  int status;           // 0 ok; 1 |b| = 0; <0 fatal (zero initial norms)
  int executed;         // loop bodies executed
  int was_active;       // the iteration whose scalars were just computed was live (done was 0 before)
  int do_update;        // x/w update of that iteration must be applied (rho != 0)
  // Deferred normalisation (split path): the solver keeps uhat = beta*u and vhat = alpha*v and folds the factors into
  // the next update instead of rescaling the vectors (S(vhat/alpha) = (S vhat)/alpha, S^T(uhat/beta) = (S^T uhat)/beta).
  double su;            // u = su * uhat  (1/beta, 1 when beta == 0)
  double sv;            // v = sv * vhat  (1/alpha, 1 when alpha == 0)
  double cu, cq;        // next u update: uhat = cu*uhat + cq*q
  double cv;            // next v update: vhat = cv*vhat + su*v2
};

// ---------------------------------------------------------------------------------------------
// Device helpers.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum (fixed tree). `red` must hold >= 32 doubles. All threads get the sum.
__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = (lane < nw) ? red[lane] : 0.0;
  t = warp_sum(t);
  return t;
}
#endif

}  // namespace tfx
