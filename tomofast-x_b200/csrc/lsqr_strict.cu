// lsqr_strict.cu -- LSQR in the reference's exact operation order (option "strict_order").
//
// Purpose: parity evidence. LSQR's residual history is chaotic in its mid-phase: scaling the right-hand
// side of config A by (1 + 2e-16) moves r_k of the REFERENCE ALGORITHM ITSELF by up to 1e-3 relative at
// iterations 13-40 (tests/test_oracle_mansf.py measures this), so no implementation with a different
// summation order can match those iterates to 1e-6. This mode removes the difference instead of
// tolerating it: every sum runs sequentially in the order of the Fortran loops, products and additions
// are rounded separately (the reference is built without FMA contraction, Makefile:51), norms follow
// libgfortran's norm2 / sum(x**2), and the scalar recurrences run on the host in plain IEEE double.
// One thread per matrix row (or column): slow by design, used only by the parity tests.
//
// Follows lsqr_solve_sensit (src/inversion/lsqr_solver2.F90:47-308) and lsqr_solve (:321-473) line by line.
#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

#include <math.h>

#include <algorithm>

namespace tfx {

namespace {

// b(i_all) = b(i_all) + sa(k) * x(ija(k)), k ascending (sparse_matrix.f90:322-327 / :397-403 via A^T).
__global__ void __launch_bounds__(128) ks_seq_spmv(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                   const float *__restrict__ val, const int32_t *__restrict__ segmap,
                                                   int nseg, const double *__restrict__ x, double *__restrict__ y) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int out = segmap[s];
  double acc = y[out];
  for (int64_t k = ptr[s]; k < ptr[s + 1]; ++k) acc = __dadd_rn(acc, __dmul_rn((double)val[k], x[idx[k]]));
  y[out] = acc;
}
__global__ void __launch_bounds__(256) ks_scale(double *x, int64_t n, double f) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = __dmul_rn(f, x[i]);
}
__global__ void __launch_bounds__(256) ks_add(double *v, const double *v2, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = __dadd_rn(v[i], v2[i]);
}
// x = t1*w + x ; w = t2*w + v (lsqr_solver2.F90:269-270), soft threshold (:478-494)
__global__ void __launch_bounds__(256) ks_xw(double *x, double *w, const double *v, int64_t n, double t1, double t2,
                                             double gamma) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double wi = w[i];
    double xi = __dadd_rn(__dmul_rn(t1, wi), x[i]);
    w[i] = __dadd_rn(__dmul_rn(t2, wi), v[i]);
    if (gamma != 0.0) {
      if (fabs(xi) <= gamma) xi = 0.0;
      else if (xi <= -gamma) xi = __dadd_rn(xi, gamma);
      else if (xi >= gamma) xi = __dsub_rn(xi, gamma);
    }
    x[i] = xi;
  }
}
// libgfortran norm2 (scaled sum of squares), one thread, element order.
__global__ void ks_norm2(const double *x, int64_t n, double *out) {
  double scale = 1.0, ssq = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double xi = x[i];
    if (xi != 0.0) {
      const double a = fabs(xi);
      if (scale < a) {
        const double v = scale / a;
        ssq = __dadd_rn(1.0, __dmul_rn(ssq, __dmul_rn(v, v)));
        scale = a;
      } else {
        const double v = a / scale;
        ssq = __dadd_rn(ssq, __dmul_rn(v, v));
      }
    }
  }
  *out = __dmul_rn(scale, sqrt(ssq));
}
// s = sum(x**2), one thread, element order (lsqr_solver2.F90:513).
__global__ void ks_sumsq(const double *x, int64_t n, double *out) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s = __dadd_rn(s, __dmul_rn(x[i], x[i]));
  *out = s;
}
__global__ void ks_diffsq(const double *a, const double *b, int64_t n, double *out) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double d = __dsub_rn(a[i], b[i]);
    s = __dadd_rn(s, __dmul_rn(d, d));
  }
  *out = s;
}

inline int vgrid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 1184)); }

}  // namespace

int lsqr_run_strict(const LsqrParams &p, Matrix *S, Matrix *C, double *d_u, double *d_x, LsqrResult &res) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const int64_t ncol = p.ncolumns, nlines = p.nlines;
  const int32_t nls = S->nl;
  const bool have_C = C && C->has_seg && !C->fwd.empty();
  const bool wav = (!p.single_matrix && p.compression_type > 0 && !p.wavelet_domain);
  const bool misfit_on = (!p.single_matrix && p.target_misfit > 0.0);
  if (!S->has_seg) return fail(-57, "lsqr(strict_order): the sensitivity matrix has no compressed-row representation");
  if (comm_nranks() > 1) return fail(-58, "lsqr(strict_order): single rank only");

  DevBuf<double> bv, bw, bv2, bb0, bsx, bs;
  TFX_TRY(bv.alloc(ncol)); TFX_TRY(bw.alloc(ncol)); TFX_TRY(bv2.alloc(ncol)); TFX_TRY(bs.alloc(4));
  if (misfit_on) { TFX_TRY(bb0.alloc(nls)); TFX_TRY(bsx.alloc(nls)); }
  double *v = bv.p, *w = bw.p, *v2 = bv2.p;
  TFX_CUDA(cudaMemsetAsync(d_x, 0, ncol * 8, st));
  if (misfit_on) TFX_CUDA(cudaMemcpyAsync(bb0.p, d_u, (size_t)nls * 8, cudaMemcpyDeviceToDevice, st));

  auto scalar = [&](double *dptr, double &h) -> int {
    TFX_CUDA(cudaMemcpyAsync(&h, dptr, 8, cudaMemcpyDeviceToHost, st));
    TFX_CUDA(cudaStreamSynchronize(st));
    return 0;
  };
  auto spmv = [&](SegMatrix &m, const double *x, double *y) {
    if (m.nseg > 0) {
      ks_seq_spmv<<<(m.nseg + 127) / 128, 128, 0, st>>>(m.ptr.p, m.idx.p, m.val.p, m.segmap.p, m.nseg, x, y);
      c.launches++;
    }
  };
  auto wavelet = [&](double *vec, bool fwd) -> int {
    for (int i = 0; i < 2; ++i) {
      if (!p.solve_problem[i]) continue;
      for (int k = 0; k < p.ncomponents; ++k)
        TFX_TRY(wavelet3d_device(vec + ((size_t)i * p.ncomponents + k) * (size_t)p.nelements, p.nx, p.ny, p.nz,
                                 p.compression_type, fwd, st));
    }
    return 0;
  };
  // v = S^T u_d [inverse wavelet] ... (+ C^T u_c)
  auto trans_products = [&](bool first) -> int {
    if (p.single_matrix && !first) {   // lsqr_solve: add_trans_mult_vector accumulates straight into v (:414)
      spmv(S->trn, d_u, v);
      return 0;
    }
    TFX_CUDA(cudaMemsetAsync(v2, 0, ncol * 8, st));
    spmv(S->trn, d_u, v2);
    if (wav) TFX_TRY(wavelet(v2, false));
    if (first) {
      TFX_CUDA(cudaMemcpyAsync(v, v2, ncol * 8, cudaMemcpyDeviceToDevice, st));   // v = v2 (:145)
    } else {
      ks_add<<<vgrid(ncol), 256, 0, st>>>(v, v2, ncol);                           // v = v + v2 (:236)
      c.launches++;
    }
    if (have_C) spmv(C->trn, d_u + nls, v);                                       // (:147, :238)
    return 0;
  };

  res.history.clear(); res.iters = 0; res.status = 0; res.fused = false; res.r = 1.0;
  double alpha, beta, rho, rhobar, phi, phibar, theta, b1, cc, r, s, t1, t2, rho_inv, nrm, mis = 0.0;
  ks_norm2<<<1, 1, 0, st>>>(d_u, nlines, bs.p); c.launches++;
  TFX_TRY(scalar(bs.p, nrm));
  if (nrm == 0.0) { res.status = 1; return 0; }                                   // |b| = 0 (:123-126)
  beta = nrm;
  ks_scale<<<vgrid(nlines), 256, 0, st>>>(d_u, nlines, 1.0 / beta); c.launches++;
  b1 = beta;
  TFX_TRY(trans_products(true));
  ks_sumsq<<<1, 1, 0, st>>>(v, ncol, bs.p); c.launches++;
  TFX_TRY(scalar(bs.p, nrm));
  alpha = sqrt(nrm);
  if (alpha == 0.0) return fail(-56, "Could not normalize initial v, zero denominator!");
  ks_scale<<<vgrid(ncol), 256, 0, st>>>(v, ncol, 1.0 / alpha); c.launches++;
  rhobar = alpha; phibar = beta;
  TFX_CUDA(cudaMemcpyAsync(w, v, ncol * 8, cudaMemcpyDeviceToDevice, st));
  int iter = 1;
  r = 1.0;
  while (iter <= p.niter && r > p.rmin) {
    if (misfit_on) {                                                              // :168-189
      TFX_CUDA(cudaMemcpyAsync(v2, d_x, ncol * 8, cudaMemcpyDeviceToDevice, st));
      if (wav) TFX_TRY(wavelet(v2, true));
      TFX_CUDA(cudaMemsetAsync(bsx.p, 0, (size_t)nls * 8, st));
      spmv(S->fwd, v2, bsx.p);
      ks_diffsq<<<1, 1, 0, st>>>(bsx.p, bb0.p, nls, bs.p); c.launches++;
      TFX_TRY(scalar(bs.p, nrm));
      mis = sqrt(nrm / (double)nls);
      if (mis <= p.target_misfit) break;
    }
    ks_scale<<<vgrid(nlines), 256, 0, st>>>(d_u, nlines, -alpha); c.launches++;   // u = -alpha*u (:195)
    const double *vin = v;
    if (wav) {
      TFX_CUDA(cudaMemcpyAsync(v2, v, ncol * 8, cudaMemcpyDeviceToDevice, st));
      TFX_TRY(wavelet(v2, true));
      vin = v2;
    }
    spmv(S->fwd, vin, d_u);                                                       // :209
    if (have_C) spmv(C->fwd, v, d_u + nls);                                       // :211
    ks_norm2<<<1, 1, 0, st>>>(d_u, nlines, bs.p); c.launches++;                   // :218
    TFX_TRY(scalar(bs.p, beta));
    if (beta != 0.0) { ks_scale<<<vgrid(nlines), 256, 0, st>>>(d_u, nlines, 1.0 / beta); c.launches++; }
    ks_scale<<<vgrid(ncol), 256, 0, st>>>(v, ncol, -beta); c.launches++;          // :225
    TFX_TRY(trans_products(false));
    ks_sumsq<<<1, 1, 0, st>>>(v, ncol, bs.p); c.launches++;                       // :241
    TFX_TRY(scalar(bs.p, nrm));
    alpha = sqrt(nrm);
    if (alpha != 0.0) { ks_scale<<<vgrid(ncol), 256, 0, st>>>(v, ncol, 1.0 / alpha); c.launches++; }
    rho = sqrt(rhobar * rhobar + beta * beta);                                    // :248
    if (rho == 0.0) break;
    rho_inv = 1.0 / rho;
    cc = rhobar * rho_inv; s = beta * rho_inv; theta = s * alpha; rhobar = -cc * alpha;
    phi = cc * phibar; phibar = s * phibar; t1 = phi * rho_inv; t2 = -theta * rho_inv;
    ks_xw<<<vgrid(ncol), 256, 0, st>>>(d_x, w, v, ncol, t1, t2, p.gamma); c.launches++;
    r = phibar / b1;
    res.history.push_back(r);
    if (p.single_matrix) {   // lsqr_solve: the small-rhobar exit sits before iter = iter + 1 (:459-465)
      if (fabs(rhobar) < (double)1.e-30f) break;
      iter += 1;
    } else {                 // lsqr_solve_sensit: after it (:281-289)
      iter += 1;
      if (fabs(rhobar) < (double)1.e-30f) break;
    }
  }
  (void)mis;
  TFX_CUDA(cudaStreamSynchronize(st));
  TFX_CUDA(cudaGetLastError());
  res.iters = (int32_t)res.history.size();
  res.reported_iters = iter - 1;
  res.r = r;
  return 0;
}

}  // namespace tfx
