// lsqr.cu -- device-resident LSQR (Paige & Saunders) over the stacked operator [S; C].
//
// Replaces lsqr_solve_sensit (src/inversion/lsqr_solver2.F90:47-308), lsqr_solve (:321-473),
// normalize (:501-530) and apply_soft_thresholding (:478-494).
//
// * All scalars (alpha, beta, rhobar, phibar, ...) live in one small device struct; the host only
//   enqueues kernels and polls the `done` flag every few iterations, so the loop stops at exactly
//   the iteration the reference stops at (iter > niter, r <= rmin, rho == 0, |rhobar| < 1e-30,
//   misfit target) without a host round trip per iteration: once `done` is set every later kernel
//   returns immediately.
// * Column-sharded multi-GPU (the reference's own decomposition, lsqr_solver2.F90:16): every rank
//   owns a slab of columns; the products S_loc v_loc are summed over ranks (MPI_Allreduce of u at
//   :214 -> ncclAllReduce here) and |v|^2 is a scalar all-reduce (:514).
// * FUSED path (uncompressed S, no wavelet inside the loop): one sweep over S per iteration does
//   S^T u, the v update and S vhat of the next iteration (dense.cu). The normalisation by alpha is
//   applied afterwards to the short vector: u = -alpha u + (S vhat)/alpha  [linearity].
// * SPLIT path (compressed S, or wavelet / misfit inside the loop): the reference's order, two
//   products per iteration.
#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

#include <math.h>

#include <algorithm>

namespace tfx {

static const int kVecThreads = 256;

// ---------------------------------------------------------------------------------------------
// Vector kernels (fixed grids -> fixed summation order -> deterministic)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVecThreads) k_sumsq_partial(const double *__restrict__ x, int64_t n,
                                                                double *__restrict__ partial, const int *done) {
  if (*done) return;
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fma(x[i], x[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// u = -alpha*u + qscale*q, and the partial sums of |u|^2.  (lsqr_solver2.F90:194-214)
__global__ void __launch_bounds__(kVecThreads) k_u_update(double *__restrict__ u, const double *__restrict__ q,
                                                           int64_t n, const LsqrScalars *sc, int scale_q_by_inv_alpha,
                                                           double *__restrict__ partial) {
  if (sc->done) return;
  __shared__ double red[32];
  const double na = sc->neg_alpha;
  const double qs = scale_q_by_inv_alpha ? sc->inv_alpha : 1.0;
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = fma(na, u[i], qs * q[i]);
    u[i] = v;
    s = fma(v, v, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out = sum(partial[0..n)) in a fixed order (single block).
__global__ void __launch_bounds__(kVecThreads) k_final_sum(const double *__restrict__ partial, int n, double *out,
                                                            const int *done) {
  if (*done) return;
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s;
}

// beta = sqrt(sumsq); u-normalisation factors. init: also b1 and the |b| = 0 early return
// (lsqr_solver2.F90:123-134, :218-222).
__global__ void k_scal_beta(const double *sumsq, LsqrScalars *sc, int init) {
  if (sc->done) return;
  const double beta = sqrt(*sumsq);
  sc->beta = beta;
  sc->neg_beta = -beta;
  if (beta != 0.0) {
    sc->inv_beta = 1.0 / beta;
  } else {
    sc->inv_beta = 1.0;   // normalize() returns ierr = -1 and leaves the vector untouched
    if (init) {
      sc->status = 1;     // "|b| = 0, the model is exact": x = 0 is returned
      sc->done = 1;
    }
  }
  if (init) sc->b1 = beta;
}

__global__ void __launch_bounds__(kVecThreads) k_scale(double *__restrict__ x, int64_t n, const double *factor,
                                                        const int *done) {
  if (*done) return;
  const double f = *factor;
  if (f == 1.0) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = f * x[i];
}

// v = -beta*v + v2   (lsqr_solver2.F90:225,236)
__global__ void __launch_bounds__(kVecThreads) k_v_update(double *__restrict__ v, const double *__restrict__ v2,
                                                           int64_t n, const LsqrScalars *sc) {
  if (sc->done) return;
  const double nb = sc->neg_beta;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = fma(nb, v[i], v2[i]);
}

// Scalar recurrences after |v|^2 is known (lsqr_solver2.F90:150-157 for init, :241-289 in the loop).
__global__ void k_scal_alpha(const double *sumsq, LsqrScalars *sc, int init, int niter, double rmin,
                             double *__restrict__ hist) {
  if (sc->done) {
    sc->was_active = 0;
    return;
  }
  sc->was_active = 1;
  const double alpha = sqrt(*sumsq);
  sc->alpha = alpha;
  sc->neg_alpha = -alpha;
  sc->inv_alpha = (alpha != 0.0) ? 1.0 / alpha : 1.0;
  if (init) {
    sc->do_update = 0;
    if (alpha == 0.0) {   // "Could not normalize initial v, zero denominator!"
      sc->status = -3;
      sc->done = 1;
      return;
    }
    sc->rhobar = alpha;
    sc->phibar = sc->beta;
    sc->iter = 1;
    sc->r = 1.0;
    if (!(sc->iter <= niter && sc->r > rmin)) sc->done = 1;
    return;
  }
  const double beta = sc->beta;
  const double rho = sqrt(sc->rhobar * sc->rhobar + beta * beta);
  sc->rho = rho;
  if (rho == 0.0) {   // "rho = 0. Exiting." -- leaves the loop before the x/w update
    sc->do_update = 0;
    sc->done = 1;
    return;
  }
  const double rho_inv = 1.0 / rho;
  const double c = sc->rhobar * rho_inv;
  const double s = beta * rho_inv;
  const double theta = s * alpha;
  sc->rhobar = -c * alpha;
  const double phi = c * sc->phibar;
  sc->phibar = s * sc->phibar;
  sc->c = c; sc->s = s; sc->theta = theta; sc->phi = phi;
  sc->t1 = phi * rho_inv;
  sc->t2 = -theta * rho_inv;
  sc->do_update = 1;
  sc->r = sc->phibar / sc->b1;
  hist[sc->iter - 1] = sc->r;
  sc->iter += 1;
  sc->executed += 1;
  if (fabs(sc->rhobar) < (double)1.e-30f) sc->done = 1;                   // :286-289 (single-precision literal)
  if (!(sc->iter <= niter && sc->r > rmin)) sc->done = 1;                  // loop condition :163
}

// v = v/alpha ; x = t1*w + x ; w = t2*w + v ; optional soft threshold (lsqr_solver2.F90:241,269-275).
__global__ void __launch_bounds__(kVecThreads) k_xw_update(double *__restrict__ v, double *__restrict__ x,
                                                            double *__restrict__ w, int64_t n, const LsqrScalars *sc,
                                                            int init, double gamma) {
  if (!sc->was_active) return;
  const double ia = sc->inv_alpha;
  const int upd = sc->do_update;
  const double t1 = sc->t1, t2 = sc->t2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double vi = (ia == 1.0) ? v[i] : ia * v[i];
    v[i] = vi;
    if (init) {
      w[i] = vi;
    } else if (upd) {
      const double wi = w[i];
      double xi = fma(t1, wi, x[i]);
      w[i] = fma(t2, wi, vi);
      if (gamma != 0.0) {
        if (fabs(xi) <= gamma) xi = 0.0;
        else if (xi <= -gamma) xi = xi + gamma;
        else if (xi >= gamma) xi = xi - gamma;
      }
      x[i] = xi;
    }
  }
}

// Fused path, columns NOT covered by the dense block (e.g. the unused second problem of the joint
// column space): vhat = -beta v + g there, plus the partial sums of |vhat|^2 over those columns.
__global__ void __launch_bounds__(kVecThreads) k_outside_update(double *__restrict__ v, const double *__restrict__ g,
                                                                 int64_t n, int64_t blk0, int64_t blk1,
                                                                 const LsqrScalars *sc, double *__restrict__ partial) {
  if (sc->done) return;
  __shared__ double red[32];
  const double nb = sc->neg_beta;
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= blk0 && i < blk1) continue;
    const double val = fma(nb, v[i], g ? g[i] : 0.0);
    v[i] = val;
    s = fma(val, val, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// *out += sum(partial[0..n))
__global__ void __launch_bounds__(kVecThreads) k_final_sum_add(const double *__restrict__ partial, int n, double *out,
                                                                const int *done) {
  if (*done) return;
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out += s;
}

// misfit = sqrt(sum((Sx - b0)^2)/n) (lsqr_solver2.F90:183-188)
__global__ void __launch_bounds__(kVecThreads) k_diffsq_partial(const double *__restrict__ a,
                                                                 const double *__restrict__ b, int64_t n,
                                                                 double *__restrict__ partial, const int *done) {
  if (*done) return;
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = a[i] - b[i];
    s = fma(d, d, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void k_scal_misfit(const double *sumsq, LsqrScalars *sc, double target, int n) {
  if (sc->done) return;
  sc->misfit = sqrt(*sumsq / (double)n);
  if (sc->misfit <= target) sc->done = 1;   // "Reached the target misfit, exiting the loop."
}

// ---------------------------------------------------------------------------------------------
// Driver
// ---------------------------------------------------------------------------------------------
namespace {

struct Timing {
  cudaEvent_t loop0 = nullptr, loop1 = nullptr;
  std::vector<cudaEvent_t> sw;   // pairs around the dominant sweep kernel (option "profile_sweeps")
  size_t used = 0;
  int ensure() {
    if (!loop0) { TFX_CUDA(cudaEventCreate(&loop0)); TFX_CUDA(cudaEventCreate(&loop1)); }
    return 0;
  }
  int next(cudaEvent_t *e) {
    if (used == sw.size()) {
      cudaEvent_t n;
      TFX_CUDA(cudaEventCreate(&n));
      sw.push_back(n);
    }
    *e = sw[used++];
    return 0;
  }
};
Timing &timing() {
  static Timing t;
  return t;
}

struct Work {
  DevBuf<double> v, w, v2, g, q, b0, sx, partial, red;
  DevBuf<LsqrScalars> sc;
  DevBuf<double> hist;
};

Work &work() {
  static Work w;
  return w;
}

inline int vec_grid(int64_t n) {
  int64_t b = (n + kVecThreads - 1) / kVecThreads;
  int64_t cap = (int64_t)ctx().num_sms * 8;
  return (int)std::max<int64_t>(1, std::min(b, cap));
}

// apply_wavelet_transform (src/inversion/wavelet_utils.F90:37-72): every active problem and component of
// v(nelements, ncomponents, 2) is one nx*ny*nz volume; with several ranks the slabs are assembled into the full
// volume on every GPU (wavelet_slab_device, data.cu) instead of the reference's gather to rank 0 / scatter.
int apply_wavelet(const LsqrParams &p, double *d_v, bool fwd, int64_t nsmaller, cudaStream_t st) {
  for (int i = 0; i < 2; ++i) {
    if (!p.solve_problem[i]) continue;
    for (int k = 0; k < p.ncomponents; ++k) {
      double *vol = d_v + ((size_t)i * p.ncomponents + k) * (size_t)p.nelements;
      TFX_TRY(wavelet_slab_device(vol, p.nelements, nsmaller, p.nx, p.ny, p.nz, p.compression_type, fwd, st));
    }
  }
  return 0;
}

}  // namespace

int lsqr_run(const LsqrParams &p, Matrix *S, Matrix *C, double *d_u, double *d_x, LsqrResult &res) {
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  Work &W = work();
  const int64_t ncol = p.ncolumns, nlines = p.nlines;
  const int32_t nls = S->nl;
  const int32_t ncons = C ? C->nl : 0;
  const bool have_C = C && C->has_seg && !C->fwd.empty();
  const bool wav = (!p.single_matrix && p.compression_type > 0 && !p.wavelet_domain);
  const bool misfit = (!p.single_matrix && p.target_misfit > 0.0);
  const int nranks = comm_nranks();

  // Sanity checks of the reference (lsqr_solver2.F90:85-89, :342-345).
  if (nls + ncons != nlines || S->ncolumns != ncol || (C && C->ncolumns != ncol))
    return fail(-50, p.single_matrix ? "Wrong matrix size in lsqr_solve! Exiting."
                                     : "Wrong matrix sizes in lsqr_solve_sensit! Exiting.");
  if (!S->finalized || (C && !C->finalized)) return fail(-51, "lsqr: matrix is not finalized");
  int64_t nsmaller = 0;
  if (wav) {
    int64_t total = 0;
    TFX_TRY(comm_slab_offset(p.nelements, &nsmaller, &total));
    if (total != (int64_t)p.nx * p.ny * p.nz)
      return fail(-53, "lsqr: the ranks' nelements must add up to nx*ny*nz when the wavelet transform runs inside the loop");
  }

  if (g_opt_strict_order) {
    if (S->has_blocks) return fail(-57, "lsqr: strict_order needs the CSR copies, a row-blocked matrix has none");
    return lsqr_run_strict(p, S, C, d_u, d_x, res);
  }

  const bool dense_ok = S->has_dense && S->dense.nrows == nls && S->dense_row0 == 0 && S->dense.nrows <= kDenseMaxRows;
  if (S->has_dense && !dense_ok && !S->has_seg) return fail(-54, "lsqr: dense sensitivity block does not cover all data rows");
  if (!dense_ok && !S->has_t16 && !S->has_seg && !S->has_blocks)
    return fail(-55, "lsqr: the sensitivity matrix has no device representation");
  const bool fused = dense_ok && !wav && !misfit;
  res.fused = fused;
  res.history.clear();
  res.iters = 0;
  res.status = 0;

  Timing &T = timing();
  TFX_TRY(T.ensure());
  T.used = 0;
  const int GV = vec_grid(std::max<int64_t>(ncol, nlines));
  TFX_TRY(W.v.alloc(ncol)); TFX_TRY(W.w.alloc(ncol)); TFX_TRY(W.v2.alloc(ncol)); TFX_TRY(W.g.alloc(ncol));
  TFX_TRY(W.q.alloc(nlines + 1));
  TFX_TRY(W.partial.alloc((size_t)c.num_sms * 8 + 8)); TFX_TRY(W.red.alloc(8));
  TFX_TRY(W.sc.alloc(1)); TFX_TRY(W.hist.alloc(std::max(1, p.niter)));
  if (misfit) { TFX_TRY(W.b0.alloc(nls)); TFX_TRY(W.sx.alloc(nls)); }
  double *v = W.v.p, *w = W.w.p, *v2 = W.v2.p, *g = W.g.p, *q = W.q.p, *partial = W.partial.p, *red = W.red.p;
  LsqrScalars *sc = W.sc.p;
  const int *done = &sc->done;
  TFX_CUDA(cudaMemsetAsync(v, 0, ncol * 8, st));
  TFX_CUDA(cudaMemsetAsync(w, 0, ncol * 8, st));
  TFX_CUDA(cudaMemsetAsync(v2, 0, ncol * 8, st));
  TFX_CUDA(cudaMemsetAsync(g, 0, ncol * 8, st));
  TFX_CUDA(cudaMemsetAsync(q, 0, (nlines + 1) * 8, st));
  TFX_CUDA(cudaMemsetAsync(d_x, 0, ncol * 8, st));                         // x = 0 (:120)
  TFX_CUDA(cudaMemsetAsync(sc, 0, sizeof(LsqrScalars), st));
  TFX_CUDA(cudaMemsetAsync(W.hist.p, 0, std::max(1, p.niter) * 8, st));
  if (misfit) TFX_CUDA(cudaMemcpyAsync(W.b0.p, d_u, (size_t)nls * 8, cudaMemcpyDeviceToDevice, st));   // :110

#define LAUNCHED() c.launches++
  // ---- beta = |u| ; u = u / beta ; b1 = beta (:123-134)
  k_sumsq_partial<<<GV, kVecThreads, 0, st>>>(d_u, nlines, partial, done); LAUNCHED();
  k_final_sum<<<1, kVecThreads, 0, st>>>(partial, GV, red, done); LAUNCHED();
  k_scal_beta<<<1, 1, 0, st>>>(red, sc, 1); LAUNCHED();
  k_scale<<<GV, kVecThreads, 0, st>>>(d_u, nlines, &sc->inv_beta, done); LAUNCHED();

  // Products with S, by representation.
  auto S_trans = [&](const double *u_d, double *out) -> int {   // out(ncol) = S^T u_d
    if (dense_ok) {
      TFX_CUDA(cudaMemsetAsync(out, 0, ncol * 8, st));
      return dense_sweep(S->dense, DENSE_T_ONLY, u_d, nullptr, nullptr, out, nullptr, nullptr, nullptr, done, st);
    }
    return matrix_trans(*S, u_d, out, false, done, st);
  };
  auto S_fwd = [&](const double *xin, double *out) -> int {     // out(nls) = S xin
    if (dense_ok)
      return dense_sweep(S->dense, DENSE_F_ONLY, nullptr, xin, nullptr, nullptr, nullptr, out, nullptr, done, st);
    return matrix_fwd(*S, xin, out, false, 0, done, st);
  };

  if (fused) {
    // =========================== FUSED PATH ===========================
    const bool outside = (S->dense.col0 != 0 || S->dense.ncols != ncol);
    auto sweep = [&]() -> int {
      // g = C^T u_c ; vhat = -beta v + S^T u_d + g ; q_d = S vhat ; n2 = |vhat|^2 ; q_c = C vhat
      if (have_C) TFX_TRY(seg_spmv(C->trn, d_u + nls, g, false, 0, (int32_t)ncol, 0, done, st));
      cudaEvent_t ea = nullptr, eb = nullptr;
      if (g_opt_profile_sweeps && T.used < 4096) {
        TFX_TRY(T.next(&ea)); TFX_TRY(T.next(&eb));
        TFX_CUDA(cudaEventRecord(ea, st));
      }
      TFX_TRY(dense_sweep(S->dense, DENSE_FUSED, d_u, v, have_C ? g : nullptr, v, &sc->neg_beta, q, q + nlines, done, st));
      if (eb) TFX_CUDA(cudaEventRecord(eb, st));
      if (outside && have_C) {
        k_outside_update<<<GV, kVecThreads, 0, st>>>(v, g, ncol, S->dense.col0, (int64_t)S->dense.col0 + S->dense.ncols,
                                                     sc, partial); LAUNCHED();
        k_final_sum_add<<<1, kVecThreads, 0, st>>>(partial, GV, q + nlines, done); LAUNCHED();
      }
      if (ncons > 0) {
        if (have_C) TFX_TRY(seg_spmv(C->fwd, v, q + nls, false, 0, ncons, 0, done, st));
      }
      if (nranks > 1) TFX_TRY(comm_allreduce_sum(q, (size_t)nlines + 1, st));
      return 0;
    };
    // Columns outside the dense block (e.g. the unused second problem) stay zero in v: the sweep only
    // rewrites its own column range, and C^T u_c contributions there are added below when present.
    // (without a constraint matrix those columns are identically zero and need no work)
    // init: neg_beta multiplies v = 0, so vhat = S^T u (+ C^T u_c)
    TFX_TRY(sweep());
    k_scal_alpha<<<1, 1, 0, st>>>(q + nlines, sc, 1, p.niter, p.rmin, W.hist.p); LAUNCHED();
    k_xw_update<<<GV, kVecThreads, 0, st>>>(v, d_x, w, ncol, sc, 1, p.gamma); LAUNCHED();
    const int chk = (S->device_nnz() > (int64_t)2e8) ? 1 : 16;
    int host_done = 0;
    TFX_CUDA(cudaEventRecord(T.loop0, st));
    for (int it = 1; it <= p.niter && !host_done; ++it) {
      // u = -alpha u + (S vhat, C vhat)/alpha ; beta = |u| ; u /= beta
      k_u_update<<<GV, kVecThreads, 0, st>>>(d_u, q, nlines, sc, 1, partial); LAUNCHED();
      k_final_sum<<<1, kVecThreads, 0, st>>>(partial, GV, red, done); LAUNCHED();
      k_scal_beta<<<1, 1, 0, st>>>(red, sc, 0); LAUNCHED();
      k_scale<<<GV, kVecThreads, 0, st>>>(d_u, nlines, &sc->inv_beta, done); LAUNCHED();
      TFX_TRY(sweep());
      k_scal_alpha<<<1, 1, 0, st>>>(q + nlines, sc, 0, p.niter, p.rmin, W.hist.p); LAUNCHED();
      k_xw_update<<<GV, kVecThreads, 0, st>>>(v, d_x, w, ncol, sc, 0, p.gamma); LAUNCHED();
      if (it % chk == 0 || it == p.niter) {
        TFX_CUDA(cudaMemcpyAsync(&host_done, &sc->done, sizeof(int), cudaMemcpyDeviceToHost, st));
        TFX_CUDA(cudaStreamSynchronize(st));
      }
    }
  } else {
    // =========================== SPLIT PATH (reference order) ===========================
    // v = S^T u_d [inverse wavelet] + C^T u_c ; alpha = |v| ; v /= alpha ; w = v   (:137-157)
    TFX_TRY(S_trans(d_u, v2));
    if (wav) TFX_TRY(apply_wavelet(p, v2, false, nsmaller, st));
    k_v_update<<<GV, kVecThreads, 0, st>>>(v, v2, ncol, sc); LAUNCHED();
    if (have_C) TFX_TRY(seg_spmv(C->trn, d_u + nls, v, true, 0, (int32_t)ncol, 0, done, st));
    k_sumsq_partial<<<GV, kVecThreads, 0, st>>>(v, ncol, partial, done); LAUNCHED();
    k_final_sum<<<1, kVecThreads, 0, st>>>(partial, GV, red, done); LAUNCHED();
    if (nranks > 1) TFX_TRY(comm_allreduce_sum(red, 1, st));
    k_scal_alpha<<<1, 1, 0, st>>>(red, sc, 1, p.niter, p.rmin, W.hist.p); LAUNCHED();
    k_xw_update<<<GV, kVecThreads, 0, st>>>(v, d_x, w, ncol, sc, 1, p.gamma); LAUNCHED();
    const int chk = (S->device_nnz() > (int64_t)2e8) ? 1 : 16;
    int host_done = 0;
    TFX_CUDA(cudaEventRecord(T.loop0, st));
    // One iteration of the reference's loop body (:160-290) as a stream of launches.
    auto iter_body = [&]() -> int {
      if (misfit) {   // :168-189
        TFX_CUDA(cudaMemcpyAsync(v2, d_x, ncol * 8, cudaMemcpyDeviceToDevice, st));
        if (wav) TFX_TRY(apply_wavelet(p, v2, true, nsmaller, st));
        TFX_TRY(S_fwd(v2, W.sx.p));
        if (nranks > 1) TFX_TRY(comm_allreduce_sum(W.sx.p, nls, st));
        k_diffsq_partial<<<vec_grid(nls), kVecThreads, 0, st>>>(W.sx.p, W.b0.p, nls, partial, done); LAUNCHED();
        k_final_sum<<<1, kVecThreads, 0, st>>>(partial, vec_grid(nls), red + 1, done); LAUNCHED();
        k_scal_misfit<<<1, 1, 0, st>>>(red + 1, sc, p.target_misfit, nls); LAUNCHED();
      }
      // q = [S W(v); C v]  (:200-211), summed over ranks (:214)
      const double *vin = v;
      if (wav) {
        TFX_CUDA(cudaMemcpyAsync(v2, v, ncol * 8, cudaMemcpyDeviceToDevice, st));
        TFX_TRY(apply_wavelet(p, v2, true, nsmaller, st));
        vin = v2;
      }
      TFX_TRY(S_fwd(vin, q));
      if (ncons > 0) {
        if (have_C) TFX_TRY(seg_spmv(C->fwd, v, q + nls, false, 0, ncons, 0, done, st));
      }
      if (nranks > 1) TFX_TRY(comm_allreduce_sum(q, (size_t)nlines, st));
      // u = -alpha u + q ; beta = |u| ; u /= beta (:194-222)
      k_u_update<<<GV, kVecThreads, 0, st>>>(d_u, q, nlines, sc, 0, partial); LAUNCHED();
      k_final_sum<<<1, kVecThreads, 0, st>>>(partial, GV, red, done); LAUNCHED();
      k_scal_beta<<<1, 1, 0, st>>>(red, sc, 0); LAUNCHED();
      k_scale<<<GV, kVecThreads, 0, st>>>(d_u, nlines, &sc->inv_beta, done); LAUNCHED();
      // v = -beta v + W^-1(S^T u_d) + C^T u_c ; alpha = |v| (:225-245)
      TFX_TRY(S_trans(d_u, v2));
      if (wav) TFX_TRY(apply_wavelet(p, v2, false, nsmaller, st));
      k_v_update<<<GV, kVecThreads, 0, st>>>(v, v2, ncol, sc); LAUNCHED();
      if (have_C) TFX_TRY(seg_spmv(C->trn, d_u + nls, v, true, 0, (int32_t)ncol, 0, done, st));
      k_sumsq_partial<<<GV, kVecThreads, 0, st>>>(v, ncol, partial, done); LAUNCHED();
      k_final_sum<<<1, kVecThreads, 0, st>>>(partial, GV, red, done); LAUNCHED();
      if (nranks > 1) TFX_TRY(comm_allreduce_sum(red, 1, st));
      k_scal_alpha<<<1, 1, 0, st>>>(red, sc, 0, p.niter, p.rmin, W.hist.p); LAUNCHED();
      k_xw_update<<<GV, kVecThreads, 0, st>>>(v, d_x, w, ncol, sc, 0, p.gamma); LAUNCHED();
      return 0;
    };
    // Launch-bound regime (small matrices: ~16 launches of a few microseconds each per iteration): the body is
    // captured once into a CUDA graph -- after a first, directly launched iteration has done every lazy allocation --
    // and replayed; all kernel arguments are iteration-independent (the scalars live in *sc on the device).
    cudaGraphExec_t gexec = nullptr;
    unsigned long long launches_per_iter = 0;
    const bool want_graph = g_opt_lsqr_graph && nranks == 1 && !S->has_blocks && p.niter >= 4 &&
                            S->device_nnz() <= (int64_t)2e8;
    for (int it = 1; it <= p.niter && !host_done; ++it) {
      if (gexec) {
        TFX_CUDA(cudaGraphLaunch(gexec, st));
        c.launches += launches_per_iter;
      } else {
        TFX_TRY(iter_body());
        if (want_graph && it == 1) {
          cudaGraph_t graph = nullptr;
          const unsigned long long l0 = c.launches;
          if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            const int rc = iter_body();
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc == 0 && ce == cudaSuccess && graph && cudaGraphInstantiate(&gexec, graph, 0) == cudaSuccess)
              launches_per_iter = c.launches - l0;
            else
              gexec = nullptr;
            if (graph) cudaGraphDestroy(graph);
          }
          c.launches = l0;                 // nothing ran during the capture
          (void)cudaGetLastError();        // a failed capture falls back to direct launches
        }
      }
      if (it % chk == 0 || it == p.niter) {
        TFX_CUDA(cudaMemcpyAsync(&host_done, &sc->done, sizeof(int), cudaMemcpyDeviceToHost, st));
        TFX_CUDA(cudaStreamSynchronize(st));
      }
    }
    if (gexec) cudaGraphExecDestroy(gexec);
  }
#undef LAUNCHED
  TFX_CUDA(cudaEventRecord(T.loop1, st));
  TFX_CUDA(cudaGetLastError());
  LsqrScalars h;
  TFX_CUDA(cudaMemcpyAsync(&h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  res.iters = h.executed;
  res.status = h.status;
  res.r = h.r;
  if (h.status == -3) return fail(-56, "Could not normalize initial v, zero denominator!");
  {
    float ms = 0.f;
    TFX_CUDA(cudaEventElapsedTime(&ms, T.loop0, T.loop1));
    res.loop_ms = ms;
    res.sweep_ms = 0.0;
    res.nsweeps = 0;
    for (size_t i = 0; i + 1 < T.used; i += 2) {
      float t = 0.f;
      TFX_CUDA(cudaEventElapsedTime(&t, T.sw[i], T.sw[i + 1]));
      res.sweep_ms += t;
      res.nsweeps += 1;
    }
  }
  res.history.resize((size_t)std::max(0, h.executed));
  if (h.executed > 0)
    TFX_CUDA(cudaMemcpy(res.history.data(), W.hist.p, (size_t)h.executed * 8, cudaMemcpyDeviceToHost));
  return 0;
}

}  // namespace tfx
