// cons.cu -- producers of the constraint matrix `matrix_cons` on the device:
//   damping_add            src/inversion/damping.F90:97-261            (model damping and the ADMM term)
//   damping_gradient_add   src/inversion/damping_gradient.F90:93-203   (+ gradient.F90:71-225, grid.F90:409-426)
//   cross_gradient_calculate  src/inversion/cross_gradient.F90:220-391,455-567,676-740
//   iterate_admm_arrays    src/inversion/admm_method.F90:70-134
// The reference rebuilds matrix_cons on the host before every solve (joint_inverse_problem.F90:364-544). Here each
// producer is a pair of kernels (count, fill) that writes the rows as (row, column, value) entries straight into HBM
// in the reference's add() order -- zero values dropped (sparse_matrix.f90:219), values rounded to real(4) (:226) --
// and appends them to the matrix under construction; finalize() builds the product layouts. All arithmetic uses
// explicit round-to-nearest operations in the reference's order (no FMA contraction), so entries and right-hand
// sides are bit-identical to the host code except where pow() is involved (Lp-norm multiplier).
#include "../../include/tfx.h"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/scan.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

namespace tfx {
namespace {

constexpr int kCT = 256;

inline int cons_grid(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + kCT - 1) / kCT, (int64_t)ctx().num_sms * 8));
}

// ---- generic two-pass emitter -------------------------------------------------------------------
template <class Gen>
__global__ void __launch_bounds__(kCT) k_cons_count(Gen g, int64_t nrows, int32_t *__restrict__ cnt) {
  for (int64_t r = blockIdx.x * (int64_t)kCT + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * kCT) {
    int32_t c = 0;
    g.row(r, true, [&](int32_t, double v) { if (v != 0.0) ++c; });
    cnt[r] = c;
  }
}
template <class Gen>
__global__ void __launch_bounds__(kCT) k_cons_fill(Gen g, int64_t nrows, const int64_t *__restrict__ off,
                                                   int32_t *__restrict__ idx, int32_t *__restrict__ rowid,
                                                   float *__restrict__ val) {
  for (int64_t r = blockIdx.x * (int64_t)kCT + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * kCT) {
    int64_t o = off[r];
    g.row(r, false, [&](int32_t col, double v) {
      if (v != 0.0) { idx[o] = col - 1; rowid[o] = (int32_t)r; val[o] = (float)v; ++o; }
    });
  }
}

struct ToI64 {
  __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};

// Runs the generator over `nrows` rows and appends them to the matrix.
template <class Gen>
int emit_rows(Matrix &M, const Gen &g, int64_t nrows) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  DevBuf<int32_t> cnt;
  DevBuf<int64_t> off;
  TFX_TRY(cnt.alloc((size_t)nrows + 1));
  TFX_TRY(off.alloc((size_t)nrows + 1));
  TFX_CUDA(cudaMemsetAsync(cnt.p + nrows, 0, 4, st));
  k_cons_count<<<cons_grid(nrows), kCT, 0, st>>>(g, nrows, cnt.p);
  auto first = thrust::make_transform_iterator(thrust::device_pointer_cast(cnt.p), ToI64());
  TFX_THRUST(thrust::exclusive_scan(thrust::cuda::par.on(st), first, first + nrows + 1, thrust::device_pointer_cast(off.p)));
  int64_t nnz = 0;
  TFX_CUDA(cudaMemcpyAsync(&nnz, off.p + nrows, 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  c.launches += 2;
  RowTriplets R;
  TFX_TRY(R.idx.alloc((size_t)std::max<int64_t>(nnz, 1)));
  TFX_TRY(R.rowid.alloc((size_t)std::max<int64_t>(nnz, 1)));
  TFX_TRY(R.val.alloc((size_t)std::max<int64_t>(nnz, 1)));
  R.nnz = nnz;
  if (nnz > 0) {
    k_cons_fill<<<cons_grid(nrows), kCT, 0, st>>>(g, nrows, off.p, R.idx.p, R.rowid.p, R.val.p);
    c.launches++;
  }
  TFX_CUDA(cudaGetLastError());
  return matrix_append_triplets(M, R, (int32_t)nrows, M.ncolumns);
}

// Deterministic sum of t[offset + i*stride], i < n: fixed grid, fixed tree.
__global__ void __launch_bounds__(kCT) k_sum_partial(const double *__restrict__ t, int64_t n, int64_t stride, int64_t offset,
                                                     double *__restrict__ part) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kCT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kCT) s += t[offset + i * stride];
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
int strided_sum(const double *d_t, int64_t n, int64_t stride, int64_t offset, double *result) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const int g = (int)std::max<int64_t>(1, std::min<int64_t>((n + kCT - 1) / kCT, (int64_t)c.num_sms * 4));
  DevBuf<double> part;
  TFX_TRY(part.alloc((size_t)g));
  k_sum_partial<<<g, kCT, 0, st>>>(d_t, n, stride, offset, part.p);
  c.launches++;
  std::vector<double> h((size_t)g);
  TFX_CUDA(cudaMemcpyAsync(h.data(), part.p, (size_t)g * 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  double s = 0.0;
  for (int i = 0; i < g; ++i) s += h[(size_t)i];
  *result = s;
  return 0;
}

// ---- damping ------------------------------------------------------------------------------------
// model_diff = (m - m_ref) / column_weight, 0 where the weight is 0 (damping.F90:117-126)
__global__ void __launch_bounds__(kCT) k_model_diff(const double *__restrict__ m, const double *__restrict__ ref,
                                                    const double *__restrict__ cw, int64_t n, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)kCT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kCT) {
    const double w = cw[i];
    out[i] = (w != 0.0) ? __ddiv_rn(__dsub_rn(m[i], ref[i]), w) : 0.0;
  }
}
__device__ __forceinline__ double norm_multiplier(double model_diff, double norm_power) {   // damping.F90:249-261
  return (model_diff != 0.0) ? pow(fabs(model_diff), norm_power / 2.0 - 1.0) : 1.0;
}
struct DampingGen {
  const double *diff, *lw;      // local slab (nelements); lw may be null
  double alpha, pw, norm_power;
  int64_t nsmaller, nelements;
  int32_t param_shift;
  template <class Emit>
  __device__ void row(int64_t r, bool, Emit emit) const {
    const int64_t i = r - nsmaller;
    if (i < 0 || i >= nelements) return;                               // add_empty_rows (:151,:171)
    double value = __dmul_rn(alpha, pw);                               // :155
    if (norm_power != 2.0) value = __dmul_rn(value, norm_multiplier(diff[i], norm_power));
    if (lw) value = __dmul_rn(value, lw[i]);
    emit(param_shift + (int32_t)i + 1, value);                         // :167
  }
};
// b_RHS(i) of the rank's slab (damping_add_RHS, :213-227); the rest of the block is zero before the gather
__global__ void __launch_bounds__(kCT) k_damping_rhs(const double *__restrict__ diff, const double *__restrict__ lw,
                                                     double alpha, double pw, double norm_power, int64_t nsmaller,
                                                     int64_t nelements, int64_t ntotal, double *__restrict__ b) {
  for (int64_t r = blockIdx.x * (int64_t)kCT + threadIdx.x; r < ntotal; r += (int64_t)gridDim.x * kCT) {
    const int64_t i = r - nsmaller;
    double v = 0.0;
    if (i >= 0 && i < nelements) {
      v = __dmul_rn(-__dmul_rn(alpha, pw), diff[i]);
      if (norm_power != 2.0) v = __dmul_rn(v, norm_multiplier(diff[i], norm_power));
      if (lw) v = __dmul_rn(v, lw[i]);
    }
    b[r] = v;
  }
}
__global__ void __launch_bounds__(kCT) k_square(const double *__restrict__ a, int64_t n, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)kCT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kCT) out[i] = __dmul_rn(a[i], a[i]);
}

// ---- gradients ----------------------------------------------------------------------------------
struct GradGrid {
  const double *dX, *dY, *dZ;
  int32_t nx, ny, nz;
  __device__ __forceinline__ int32_t ind(int32_t i, int32_t j, int32_t k) const {   // grid.F90:409-426 (1-based, -1 outside)
    if (i < 1 || i > nx || j < 1 || j > ny || k < 1 || k > nz) return -1;
    return i + (j - 1) * nx + (k - 1) * nx * ny;
  }
  __device__ __forceinline__ double par(const double *val, int32_t i, int32_t j, int32_t k) const {   // gradient.F90:175-225
    if (i == nx + 1 || j == ny + 1 || k == nz + 1) return 0.0;
    if (i == 0 || j == 0 || k == 0) return 0.0;
    return val[(int64_t)(i - 1) + (int64_t)(j - 1) * nx + (int64_t)(k - 1) * nx * ny];
  }
  // get_grad (gradient.F90:71-170): type 0 backward, 1 forward, 2 central
  __device__ __forceinline__ void grad(const double *val, int32_t i, int32_t j, int32_t k, int type, double g[3]) const {
    const double hx = dX[i - 1], hy = dY[j - 1], hz = dZ[k - 1];
    if (type == 0) {
      const double c = par(val, i, j, k);
      g[0] = __ddiv_rn(__dsub_rn(c, par(val, i - 1, j, k)), hx);
      g[1] = __ddiv_rn(__dsub_rn(c, par(val, i, j - 1, k)), hy);
      g[2] = __ddiv_rn(__dsub_rn(c, par(val, i, j, k - 1)), hz);
    } else if (type == 1) {
      const double c = par(val, i, j, k);
      g[0] = __ddiv_rn(__dsub_rn(par(val, i + 1, j, k), c), hx);
      g[1] = __ddiv_rn(__dsub_rn(par(val, i, j + 1, k), c), hy);
      g[2] = __ddiv_rn(__dsub_rn(par(val, i, j, k + 1), c), hz);
    } else {
      g[0] = __ddiv_rn(__ddiv_rn(__dsub_rn(par(val, i + 1, j, k), par(val, i - 1, j, k)), 2.0), hx);
      g[1] = __ddiv_rn(__ddiv_rn(__dsub_rn(par(val, i, j + 1, k), par(val, i, j - 1, k)), 2.0), hy);
      g[2] = __ddiv_rn(__ddiv_rn(__dsub_rn(par(val, i, j, k + 1), par(val, i, j, k - 1)), 2.0), hz);
    }
  }
  __device__ __forceinline__ void ijk(int64_t p0, int32_t &i, int32_t &j, int32_t &k) const {
    i = (int32_t)(p0 % nx) + 1;
    j = (int32_t)((p0 / nx) % ny) + 1;
    k = (int32_t)(p0 / ((int64_t)nx * ny)) + 1;
  }
};

struct DampGradGen {
  GradGrid G;
  const double *val_full, *cw, *lw;     // full model, local column weight, full local weight
  double beta, pw;
  int64_t nsmaller, nelements;
  int32_t param_shift, direction;
  double *b;                            // b_RHS of the appended block (device), may be null in the fill pass
  double *cost_term;                    // gradient_val^2 per row
  template <class Emit>
  __device__ void row(int64_t r, bool first_pass, Emit emit) const {
    int32_t i, j, k;
    G.ijk(r, i, j, k);
    const bool edge = (direction == 1) ? (i == G.nx) : (direction == 2) ? (j == G.ny) : (k == G.nz);
    if (edge) {                                                        // new_row + cycle (:126,:141,:156)
      if (first_pass) cost_term[r] = 0.0;
      return;
    }
    double g[3];
    G.grad(val_full, i, j, k, 1, g);
    const double delta = (direction == 1) ? G.dX[i - 1] : (direction == 2) ? G.dY[j - 1] : G.dZ[k - 1];
    int32_t ind[2];
    ind[0] = (direction == 1) ? G.ind(i + 1, j, k) : (direction == 2) ? G.ind(i, j + 1, k) : G.ind(i, j, k + 1);
    ind[1] = G.ind(i, j, k);
    const double gradient_val = g[direction - 1];
    double v[2];
    v[0] = __ddiv_rn(1.0, delta);
    v[1] = -v[0];
    for (int l = 0; l < 2; ++l) {                                      // :177-184
      if (ind[l] > nsmaller && ind[l] <= nsmaller + nelements) {
        const int32_t loc = ind[l] - (int32_t)nsmaller;
        const double value = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(v[l], pw), beta), cw[loc - 1]), lw[r]);
        emit(param_shift + loc, value);
      }
    }
    if (first_pass) {
      b[r] = __dmul_rn(-__dmul_rn(__dmul_rn(pw, beta), gradient_val), lw[r]);      // :189
      cost_term[r] = __dmul_rn(gradient_val, gradient_val);                        // :192
    }
  }
};

struct Tau {
  double val[3];
  double dm1[4][3], dm2[4][3];
  int32_t ind[4][3];
};
__device__ __forceinline__ void tau_zero(Tau &t) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    t.val[c] = 0.0;
#pragma unroll
    for (int l = 0; l < 4; ++l) { t.dm1[l][c] = 0.0; t.dm2[l][c] = 0.0; t.ind[l][c] = 0; }
  }
}
__device__ __forceinline__ void cross_product(const double a[3], const double b[3], double out[3]) {   // vector.f90:107-116
  out[0] = __dsub_rn(__dmul_rn(a[1], b[2]), __dmul_rn(a[2], b[1]));
  out[1] = __dsub_rn(__dmul_rn(a[2], b[0]), __dmul_rn(a[0], b[2]));
  out[2] = __dsub_rn(__dmul_rn(a[0], b[1]), __dmul_rn(a[1], b[0]));
}
#define DV(a, b) __ddiv_rn((a), (b))
#define SB(a, b) __dsub_rn((a), (b))
// cross_gradient_calculate_tau (cross_gradient.F90:455-567): der_type 1 forward, otherwise central
__device__ void calc_tau(const GradGrid &G, const double *m1, const double *m2, int32_t i, int32_t j, int32_t k, int der_type,
                         Tau &t) {
  double g1[3], g2[3];
  tau_zero(t);
  G.grad(m1, i, j, k, der_type == 1 ? 1 : 2, g1);
  G.grad(m2, i, j, k, der_type == 1 ? 1 : 2, g2);
  cross_product(g1, g2, t.val);
  double sx = G.dX[i - 1], sy = G.dY[j - 1], sz = G.dZ[k - 1];
  if (der_type != 1) { sx = 2.0 * sx; sy = 2.0 * sy; sz = 2.0 * sz; }
  // x
  t.dm1[0][0] = DV(g2[2], sy);  t.dm2[0][0] = DV(-g1[2], sy);
  t.dm1[1][0] = DV(-g2[1], sz); t.dm2[1][0] = DV(g1[1], sz);
  t.ind[0][0] = G.ind(i, j + 1, k); t.ind[1][0] = G.ind(i, j, k + 1);
  if (der_type == 1) {
    t.dm1[2][0] = -SB(DV(g2[2], sy), DV(g2[1], sz)); t.dm2[2][0] = -SB(DV(g1[1], sz), DV(g1[2], sy));
    t.ind[2][0] = G.ind(i, j, k);
  } else {
    t.dm1[2][0] = -t.dm1[0][0]; t.dm2[2][0] = -t.dm2[0][0];
    t.dm1[3][0] = -t.dm1[1][0]; t.dm2[3][0] = -t.dm2[1][0];
    t.ind[2][0] = G.ind(i, j - 1, k); t.ind[3][0] = G.ind(i, j, k - 1);
  }
  // y
  t.dm1[0][1] = DV(-g2[2], sx); t.dm2[0][1] = DV(g1[2], sx);
  t.dm1[1][1] = DV(g2[0], sz);  t.dm2[1][1] = DV(-g1[0], sz);
  t.ind[0][1] = G.ind(i + 1, j, k); t.ind[1][1] = G.ind(i, j, k + 1);
  if (der_type == 1) {
    t.dm1[2][1] = -SB(DV(g2[0], sz), DV(g2[2], sx)); t.dm2[2][1] = -SB(DV(g1[2], sx), DV(g1[0], sz));
    t.ind[2][1] = G.ind(i, j, k);
  } else {
    t.dm1[2][1] = -t.dm1[0][1]; t.dm2[2][1] = -t.dm2[0][1];
    t.dm1[3][1] = -t.dm1[1][1]; t.dm2[3][1] = -t.dm2[1][1];
    t.ind[2][1] = G.ind(i - 1, j, k); t.ind[3][1] = G.ind(i, j, k - 1);
  }
  // z
  t.dm1[0][2] = DV(g2[1], sx);  t.dm2[0][2] = DV(-g1[1], sx);
  t.dm1[1][2] = DV(-g2[0], sy); t.dm2[1][2] = DV(g1[0], sy);
  t.ind[0][2] = G.ind(i + 1, j, k); t.ind[1][2] = G.ind(i, j + 1, k);
  if (der_type == 1) {
    t.dm1[2][2] = -SB(DV(g2[1], sx), DV(g2[0], sy)); t.dm2[2][2] = -SB(DV(g1[0], sy), DV(g1[1], sx));
    t.ind[2][2] = G.ind(i, j, k);
  } else {
    t.dm1[2][2] = -t.dm1[0][2]; t.dm2[2][2] = -t.dm2[0][2];
    t.dm1[3][2] = -t.dm1[1][2]; t.dm2[3][2] = -t.dm2[1][2];
    t.ind[2][2] = G.ind(i - 1, j, k); t.ind[3][2] = G.ind(i, j - 1, k);
  }
}
// cross_gradient_calculate_tau_backward (:676-740)
__device__ void calc_tau_backward(const GradGrid &G, const double *m1, const double *m2, int32_t i, int32_t j, int32_t k,
                                  Tau &t) {
  double g1[3], g2[3];
  tau_zero(t);
  G.grad(m1, i, j, k, 0, g1);
  G.grad(m2, i, j, k, 0, g2);
  cross_product(g1, g2, t.val);
  const double sx = G.dX[i - 1], sy = G.dY[j - 1], sz = G.dZ[k - 1];
  t.dm1[0][0] = DV(-g2[2], sy); t.dm1[1][0] = DV(g2[1], sz);  t.dm1[2][0] = SB(DV(g2[2], sy), DV(g2[1], sz));
  t.dm2[0][0] = DV(g1[2], sy);  t.dm2[1][0] = DV(-g1[1], sz); t.dm2[2][0] = SB(DV(g1[1], sz), DV(g1[2], sy));
  t.ind[0][0] = G.ind(i, j - 1, k); t.ind[1][0] = G.ind(i, j, k - 1); t.ind[2][0] = G.ind(i, j, k);
  t.dm1[0][1] = DV(g2[2], sx);  t.dm1[1][1] = DV(-g2[0], sz); t.dm1[2][1] = SB(DV(g2[0], sz), DV(g2[2], sx));
  t.dm2[0][1] = DV(-g1[2], sx); t.dm2[1][1] = DV(g1[0], sz);  t.dm2[2][1] = SB(DV(g1[2], sx), DV(g1[0], sz));
  t.ind[0][1] = G.ind(i - 1, j, k); t.ind[1][1] = G.ind(i, j, k - 1); t.ind[2][1] = G.ind(i, j, k);
  t.dm1[0][2] = DV(-g2[1], sx); t.dm1[1][2] = DV(g2[0], sy);  t.dm1[2][2] = SB(DV(g2[1], sx), DV(g2[0], sy));
  t.dm2[0][2] = DV(g1[1], sx);  t.dm2[1][2] = DV(-g1[0], sy); t.dm2[2][2] = SB(DV(g1[0], sy), DV(g1[1], sx));
  t.ind[0][2] = G.ind(i - 1, j, k); t.ind[1][2] = G.ind(i, j - 1, k); t.ind[2][2] = G.ind(i, j, k);
}
#undef DV
#undef SB

struct CrossGradGen {
  GradGrid G;
  const double *m1, *m2, *cw1, *cw2;    // full models, local column weights
  double glob_weight;
  int64_t nsmaller, nparams_loc;
  int32_t der_type, keep1, keep2;
  double *b;                            // b_RHS of the appended block (3 rows per cell)
  double *cost_term;                    // tau%val(c)^2 per row
  double *cross_grad;                   // |tau| per cell, may be null
  template <class Emit>
  __device__ void row(int64_t r, bool first_pass, Emit emit) const {
    const int64_t p0 = r / 3;
    const int c = (int)(r - p0 * 3);
    int32_t i, j, k;
    G.ijk(p0, i, j, k);
    Tau t;
    const bool left = (i == 1 || j == 1 || k == 1), right = (i == G.nx || j == G.ny || k == G.nz);
    if (left && right) tau_zero(t);                                    // :262-266
    else if (der_type == 2 && left) calc_tau(G, m1, m2, i, j, k, 1, t);
    else if (right) calc_tau_backward(G, m1, m2, i, j, k, t);
    else calc_tau(G, m1, m2, i, j, k, der_type, t);
    const int nderiv = (der_type == 1) ? 3 : 4;                        // :203-215
    for (int l = 0; l < nderiv; ++l) {                                 // :305-325
      int32_t ind = t.ind[l][c];
      if (ind > nsmaller && ind <= nsmaller + nparams_loc) {
        ind -= (int32_t)nsmaller;
        const double d1 = keep1 ? 0.0 : t.dm1[l][c], d2 = keep2 ? 0.0 : t.dm2[l][c];      // :286-287
        emit(ind, __dmul_rn(__dmul_rn(d1, cw1[ind - 1]), glob_weight));
        emit(ind + (int32_t)nparams_loc, __dmul_rn(__dmul_rn(d2, cw2[ind - 1]), glob_weight));
      }
    }
    if (first_pass) {
      b[r] = __dmul_rn(-t.val[c], glob_weight);                        // :323
      cost_term[r] = __dmul_rn(t.val[c], t.val[c]);                    // :298-300
      if (cross_grad && c == 0)                                        // vector_get_norm (vector.f90:133-139)
        cross_grad[p0] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(t.val[0], t.val[0]), __dmul_rn(t.val[1], t.val[1])),
                                        __dmul_rn(t.val[2], t.val[2])));
    }
  }
};

// admm_method_iterate_admm_arrays (admm_method.F90:70-134); xmin/xmax(nlithos, n) Fortran order
__global__ void __launch_bounds__(kCT) k_admm(int64_t n, int32_t nlithos, const double *__restrict__ xmin,
                                              const double *__restrict__ xmax, const double *__restrict__ x,
                                              double *__restrict__ z, double *__restrict__ u, double *__restrict__ x0) {
  for (int64_t i = blockIdx.x * (int64_t)kCT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kCT) {
    const double arg = __dadd_rn(x[i], u[i]);
    bool inside = false;
    double zi = 0.0;
    for (int32_t j = 0; j < nlithos; ++j)
      if (xmin[j + (int64_t)nlithos * i] <= arg && arg <= xmax[j + (int64_t)nlithos * i]) { inside = true; zi = arg; break; }
    if (!inside) {
      double mindist = 1.e30, closest = 0.0;
      for (int32_t j = 0; j < nlithos; ++j) {
        double v = fabs(__dsub_rn(xmin[j + (int64_t)nlithos * i], arg));
        if (v < mindist) { mindist = v; closest = xmin[j + (int64_t)nlithos * i]; }
        v = fabs(__dsub_rn(xmax[j + (int64_t)nlithos * i], arg));
        if (v < mindist) { mindist = v; closest = xmax[j + (int64_t)nlithos * i]; }
      }
      zi = closest;
    }
    const double ui = __dsub_rn(__dadd_rn(u[i], x[i]), zi);            // u = u + x - z (:129)
    z[i] = zi;
    u[i] = ui;
    x0[i] = __dsub_rn(zi, ui);                                         // x0 = z - u (:131)
  }
}

int check_block(const Matrix &M, int32_t nrows_b, int64_t block_rows, const char *who) {
  if (M.finalized) return fail(-26, std::string(who) + ": the matrix is already finalized");
  if ((int64_t)M.nl_current_all + block_rows > (int64_t)M.nl)
    return fail(-28, std::string(who) + ": the rows do not fit into the matrix");
  if ((int64_t)M.nl_current_all + block_rows > (int64_t)nrows_b)
    return fail(-28, std::string(who) + ": the rows do not fit into b_RHS");
  return 0;
}

}  // namespace
}  // namespace tfx

using namespace tfx;

extern "C" int tfx_damping_add(tfx_matrix *matrix, int32_t nrows, double *b_RHS, double alpha, double problem_weight,
                               double norm_power, int32_t compression_type, int32_t nx, int32_t ny, int32_t nz,
                               int32_t nelements, const double *column_weight, const double *model,
                               const double *model_ref, int32_t param_shift, int32_t wavelet_domain,
                               const double *local_weight, int32_t myrank, int32_t nbproc, double *cost) {
  (void)myrank;
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  if (!matrix) return fail(-82, "damping_add: null matrix");
  Matrix &M = matrix->m;
  const int64_t ntotal = (int64_t)nx * ny * nz;
  if (nbproc > 1 && comm_nranks() != nbproc) return fail(-24, "damping_add: nbproc does not match the communicator");
  int64_t nsmaller = 0, total = nelements;
  if (nbproc > 1) TFX_TRY(comm_slab_offset(nelements, &nsmaller, &total));
  if (total != ntotal) return fail(-98, "Sanity check failed in damping_add!");                 // :176-177
  TFX_TRY(check_block(M, nrows, ntotal, "damping_add"));
  const int64_t row_beg0 = M.nl_current_all;                                                    // 0-based first row (:145)
  VecIO vcw, vm, vref, vlw, vb;
  TFX_TRY(vcw.bind(const_cast<double *>(column_weight), (size_t)nelements, true));
  TFX_TRY(vm.bind(const_cast<double *>(model), (size_t)nelements, true));
  TFX_TRY(vref.bind(const_cast<double *>(model_ref), (size_t)nelements, true));
  if (local_weight) TFX_TRY(vlw.bind(const_cast<double *>(local_weight), (size_t)nelements, true));
  TFX_TRY(vb.bind(b_RHS, (size_t)nrows, true));
  DevBuf<double> diff, sq;
  TFX_TRY(diff.alloc((size_t)std::max(nelements, 1)));
  k_model_diff<<<cons_grid(nelements), kCT, 0, st>>>(vm.dev, vref.dev, vcw.dev, nelements, diff.p);
  c.launches++;
  if (compression_type > 0 && wavelet_domain)                                                   // :128-142
    TFX_TRY(wavelet_slab_device(diff.p, nelements, nsmaller, nx, ny, nz, compression_type, true, st));
  DampingGen g{diff.p, local_weight ? vlw.dev : nullptr, alpha, problem_weight, norm_power, nsmaller, nelements, param_shift};
  TFX_TRY(emit_rows(M, g, ntotal));
  double *blk = vb.dev + row_beg0;
  k_damping_rhs<<<cons_grid(ntotal), kCT, 0, st>>>(diff.p, local_weight ? vlw.dev : nullptr, alpha, problem_weight, norm_power,
                                                   nsmaller, nelements, ntotal, blk);
  c.launches++;
  if (nbproc > 1) TFX_TRY(comm_allreduce_sum(blk, (size_t)ntotal, st));                         // get_full_array_in_place (:230)
  TFX_TRY(sq.alloc((size_t)ntotal));
  k_square<<<cons_grid(ntotal), kCT, 0, st>>>(blk, ntotal, sq.p);
  c.launches++;
  double cst = 0.0;
  TFX_TRY(strided_sum(sq.p, ntotal, 1, 0, &cst));                                               // :190
  if (cost) *cost = cst;
  TFX_TRY(vb.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int tfx_damping_gradient_add(tfx_matrix *matrix, int32_t nrows, double *b_RHS, double beta,
                                        double problem_weight, int32_t nx, int32_t ny, int32_t nz, const double *dX,
                                        const double *dY, const double *dZ, int32_t nelements, const double *val_full,
                                        const double *column_weight, const double *local_weight, int32_t param_shift,
                                        int32_t direction, int32_t myrank, int32_t nbproc, double *cost) {
  (void)myrank;
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  if (!matrix) return fail(-82, "damping_gradient_add: null matrix");
  if (direction < 1 || direction > 3) return fail(-99, "Wrong direction in damping_gradient_add!");
  Matrix &M = matrix->m;
  const int64_t ntotal = (int64_t)nx * ny * nz;
  if (nbproc > 1 && comm_nranks() != nbproc) return fail(-24, "damping_gradient_add: nbproc does not match the communicator");
  int64_t nsmaller = 0, total = nelements;
  if (nbproc > 1) TFX_TRY(comm_slab_offset(nelements, &nsmaller, &total));
  if (total != ntotal) return fail(-98, "damping_gradient_add: the ranks' nelements must add up to nx*ny*nz");
  TFX_TRY(check_block(M, nrows, ntotal, "damping_gradient_add"));
  const int64_t row_beg0 = M.nl_current_all;
  VecIO vx, vy, vz, vval, vcw, vlw, vb;
  TFX_TRY(vx.bind(const_cast<double *>(dX), (size_t)nx, true));
  TFX_TRY(vy.bind(const_cast<double *>(dY), (size_t)ny, true));
  TFX_TRY(vz.bind(const_cast<double *>(dZ), (size_t)nz, true));
  TFX_TRY(vval.bind(const_cast<double *>(val_full), (size_t)ntotal, true));
  TFX_TRY(vcw.bind(const_cast<double *>(column_weight), (size_t)nelements, true));
  TFX_TRY(vlw.bind(const_cast<double *>(local_weight), (size_t)ntotal, true));
  TFX_TRY(vb.bind(b_RHS, (size_t)nrows, true));
  DevBuf<double> term;
  TFX_TRY(term.alloc((size_t)ntotal));
  DampGradGen g{{vx.dev, vy.dev, vz.dev, nx, ny, nz}, vval.dev, vcw.dev, vlw.dev, beta, problem_weight, nsmaller, nelements,
                param_shift, direction, vb.dev + row_beg0, term.p};
  TFX_TRY(emit_rows(M, g, ntotal));
  double cst = 0.0;
  TFX_TRY(strided_sum(term.p, ntotal, 1, 0, &cst));
  if (cost) *cost = cst;
  TFX_TRY(vb.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int tfx_cross_gradient_calculate(tfx_matrix *matrix, int32_t nrows, double *b_RHS, int32_t nx, int32_t ny,
                                            int32_t nz, const double *dX, const double *dY, const double *dZ,
                                            int32_t nparams_loc, const double *model1, const double *model2,
                                            const double *column_weight1, const double *column_weight2,
                                            int32_t der_type, double glob_weight, const int32_t keep_model_constant[2],
                                            int32_t myrank, int32_t nbproc, double cost[3], double *cross_grad) {
  (void)myrank;
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  if (!matrix) return fail(-82, "cross_gradient_calculate: null matrix");
  if (der_type != 1 && der_type != 2) return fail(-99, "Unsupported derivative type!");          // :281-283
  Matrix &M = matrix->m;
  const int64_t ntotal = (int64_t)nx * ny * nz;
  if (nbproc > 1 && comm_nranks() != nbproc) return fail(-24, "cross_gradient_calculate: nbproc does not match the communicator");
  int64_t nsmaller = 0, total = nparams_loc;
  if (nbproc > 1) TFX_TRY(comm_slab_offset(nparams_loc, &nsmaller, &total));
  if (total != ntotal) return fail(-98, "cross_gradient_calculate: the ranks' nparams_loc must add up to nx*ny*nz");
  TFX_TRY(check_block(M, nrows, 3 * ntotal, "cross_gradient_calculate"));
  const int64_t row_beg0 = M.nl_current_all;
  VecIO vx, vy, vz, v1, v2, w1, w2, vb, vcg;
  TFX_TRY(vx.bind(const_cast<double *>(dX), (size_t)nx, true));
  TFX_TRY(vy.bind(const_cast<double *>(dY), (size_t)ny, true));
  TFX_TRY(vz.bind(const_cast<double *>(dZ), (size_t)nz, true));
  TFX_TRY(v1.bind(const_cast<double *>(model1), (size_t)ntotal, true));
  TFX_TRY(v2.bind(const_cast<double *>(model2), (size_t)ntotal, true));
  TFX_TRY(w1.bind(const_cast<double *>(column_weight1), (size_t)nparams_loc, true));
  TFX_TRY(w2.bind(const_cast<double *>(column_weight2), (size_t)nparams_loc, true));
  TFX_TRY(vb.bind(b_RHS, (size_t)nrows, true));
  if (cross_grad) TFX_TRY(vcg.bind(cross_grad, (size_t)ntotal, false));
  DevBuf<double> term;
  TFX_TRY(term.alloc((size_t)(3 * ntotal)));
  const int32_t k1 = keep_model_constant ? keep_model_constant[0] : 0, k2 = keep_model_constant ? keep_model_constant[1] : 0;
  CrossGradGen g{{vx.dev, vy.dev, vz.dev, nx, ny, nz}, v1.dev, v2.dev, w1.dev, w2.dev, glob_weight, nsmaller, nparams_loc,
                 der_type, k1 != 0, k2 != 0, vb.dev + row_beg0, term.p, cross_grad ? vcg.dev : nullptr};
  TFX_TRY(emit_rows(M, g, 3 * ntotal));
  if (cost)
    for (int cpt = 0; cpt < 3; ++cpt) TFX_TRY(strided_sum(term.p, ntotal, 3, cpt, &cost[cpt]));
  TFX_TRY(vb.copy_back());
  if (cross_grad) TFX_TRY(vcg.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int tfx_admm_iterate_admm_arrays(int32_t nelements, int32_t nlithos, const double *xmin, const double *xmax,
                                            const double *x, double *z, double *u, double *x0) {
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  if (nlithos < 1) return fail(-99, "admm: nlithos must be positive");
  VecIO vmin, vmax, vx, vz, vu, v0;
  TFX_TRY(vmin.bind(const_cast<double *>(xmin), (size_t)nelements * nlithos, true));
  TFX_TRY(vmax.bind(const_cast<double *>(xmax), (size_t)nelements * nlithos, true));
  TFX_TRY(vx.bind(const_cast<double *>(x), (size_t)nelements, true));
  TFX_TRY(vz.bind(z, (size_t)nelements, true));
  TFX_TRY(vu.bind(u, (size_t)nelements, true));
  TFX_TRY(v0.bind(x0, (size_t)nelements, false));
  k_admm<<<cons_grid(nelements), kCT, 0, st>>>(nelements, nlithos, vmin.dev, vmax.dev, vx.dev, vz.dev, vu.dev, v0.dev);
  c.launches++;
  TFX_CUDA(cudaGetLastError());
  TFX_TRY(vz.copy_back());
  TFX_TRY(vu.copy_back());
  TFX_TRY(v0.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}
