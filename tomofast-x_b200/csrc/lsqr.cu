// lsqr.cu -- device-resident LSQR (Paige & Saunders) over the stacked operator [S; C].
//
// Replaces lsqr_solve_sensit (src/inversion/lsqr_solver2.F90:47-308), lsqr_solve (:321-473),
// normalize (:501-530) and apply_soft_thresholding (:478-494).
//
// * All scalars (alpha, beta, rhobar, phibar, ...) live in one small device struct; the host only enqueues kernels and
//   polls the `done` flag every few iterations, so the loop stops at exactly the iteration the reference stops at
//   (iter > niter, r <= rmin, rho == 0, |rhobar| < 1e-30, misfit target) without a host round trip per iteration: once
//   `done` is set every later kernel returns immediately.
// * Column-sharded multi-GPU (the reference's own decomposition, lsqr_solver2.F90:16): every rank owns a slab of
//   columns. The reference all-reduces the whole u (nlines = data rows + constraint rows, :214). Here only the rows that
//   really receive contributions from several ranks travel: the data rows, the constraint rows whose stored entries
//   straddle a slab boundary ("shared" rows, found once per solve) and one scalar. A constraint row whose entries all
//   live on one rank (every damping / ADMM row, every cross-gradient row away from a slab boundary) is OWNED by that
//   rank: its u element is kept, updated and normed there only, and its |u|^2 partial rides in the scalar slot of the
//   same all-reduce -- config B: 33.6 MB -> 80 KB per iteration, and the constraint-row work is split N ways.
// * FUSED path (uncompressed S, no wavelet inside the loop): one sweep over S per iteration does S^T u, the v update and
//   S vhat of the next iteration (dense.cu). The normalisation by alpha is applied afterwards to the short vector:
//   u = -alpha u + (S vhat)/alpha  [linearity].
// * SPLIT path (compressed S, or wavelet / misfit inside the loop): two products per iteration. Neither u nor v is ever
//   rescaled: the solver carries uhat = beta*u and vhat = alpha*v and folds 1/beta, 1/alpha into the coefficients of
//   the next update (LsqrScalars::su/sv/cu/cq/cv), the constraint block's forward product is fused into the u update
//   (one pass over the rows of C), every norm is finished by the last CTA of the kernel that produced the vector
//   (fixed partial order -> deterministic) together with the scalar recurrences: 5 launches per iteration besides the
//   two products, no host synchronisation (the done flag is polled every 8 iterations).
#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

#include <math.h>
#include <string.h>

#include <thrust/copy.h>
#include <thrust/count.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>

namespace tfx {

static const int kVecThreads = 256;
int g_opt_lsqr_poll = 8;   // iterations between two reads of the device-side done flag

// ---------------------------------------------------------------------------------------------
// Helpers shared by the kernels
// ---------------------------------------------------------------------------------------------
// Finishes a grid-wide sum: every CTA deposits its partial, the last one to arrive adds them in index order (the
// result does not depend on which CTA is last) and returns true with the total in `total` for all its threads.
__device__ __forceinline__ bool grid_sum_last(double s, double *red, double *partial, unsigned *ticket, double &total) {
  __shared__ int s_last;
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s;
    __threadfence();
    const unsigned t = atomicInc(ticket, gridDim.x - 1);   // wraps to 0 after the last CTA: ready for the next launch
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double t = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(partial + i);
  total = block_sum(t, red);
  return true;
}

// beta = sqrt(sumsq) and what follows from it (lsqr_solver2.F90:123-134 for init, :218-225 in the loop).
__device__ __forceinline__ void scal_beta_update(LsqrScalars *sc, double sumsq, int init) {
  const double beta = sqrt(sumsq);
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
  sc->su = sc->inv_beta;
  sc->cv = init ? 0.0 : -beta * sc->sv;
  if (init) sc->b1 = beta;
}

// Scalar recurrences after |vhat|^2 is known (lsqr_solver2.F90:150-157 for init, :241-289 in the loop).
// unnorm: the split path keeps uhat = beta*u, so the next u update multiplies the stored vector by -alpha*su.
__device__ __forceinline__ void scal_alpha_update(LsqrScalars *sc, double sumsq, int init, int niter, double rmin,
                                                  double *__restrict__ hist, int single_matrix, int unnorm) {
  sc->was_active = 1;
  const double alpha = sqrt(sumsq);
  sc->alpha = alpha;
  sc->neg_alpha = -alpha;
  sc->inv_alpha = (alpha != 0.0) ? 1.0 / alpha : 1.0;
  sc->sv = sc->inv_alpha;
  sc->cu = unnorm ? -alpha * sc->su : -alpha;
  sc->cq = sc->inv_alpha;
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
  hist[sc->executed] = sc->r;
  sc->executed += 1;
  const bool small = fabs(sc->rhobar) < (double)1.e-30f;   // single-precision literal in the reference
  if (single_matrix) {
    // lsqr_solve leaves the loop BEFORE iter = iter + 1 (:459-465): it reports one iteration less on this exit
    if (small) sc->done = 1;
    else sc->iter += 1;
  } else {
    sc->iter += 1;                                          // lsqr_solve_sensit: iter = iter + 1, then the check (:281-289)
    if (small) sc->done = 1;
  }
  if (!(sc->iter <= niter && sc->r > rmin)) sc->done = 1;   // loop condition :163 / :384
}

// ---------------------------------------------------------------------------------------------
// Kernels (fixed grids -> fixed summation order -> deterministic)
// ---------------------------------------------------------------------------------------------
struct ConsArgs {
  const int64_t *ptr;          // CSR of the local constraint block: stored rows
  const int32_t *idx;
  const float *val;
  const int32_t *segmap;       // stored row -> constraint row
  const int32_t *seg_slot;     // stored row -> slot among the shared rows, -1: owned by this rank (null: all owned)
  int32_t nseg;
  const int32_t *empty_rows;   // constraint rows without entries on any rank (kept by rank 0)
  int32_t nempty;
  double *uc;                  // u + nls
  const double *v;             // the vector the rows multiply
  double *qsh;                 // q + nls: partial products of the shared rows (all-reduced with the data rows)
  LsqrScalars *sc;
  int phase;                   // 0: sum of squares of the owned rows of b; 1: uhat = cu*uhat + cq*(C v)
  int do_owned, do_shared;
  double *partial;
  unsigned *ticket;
  double *out;                 // receives the owned rows' sum of squares
};

// u_c = -alpha u_c + C v with the forward product of the constraint block fused in (lsqr_solver2.F90:194-211): one
// pass over the stored rows of C. LANES = 1: thread per row (damping / gradient rows, 1-12 entries); 32: warp per row.
template <int LANES>
__global__ void __launch_bounds__(kVecThreads) k_cons_rows(ConsArgs a) {
  if (a.sc->done) return;
  __shared__ double red[32];
  const double cu = a.phase ? a.sc->cu : 1.0, cq = a.phase ? a.sc->cq : 0.0;
  const int lane = threadIdx.x & (LANES - 1);
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x / LANES;
  double s = 0.0;
  for (int64_t seg = gid; seg < a.nseg; seg += gstride) {
    const int slot = a.seg_slot ? a.seg_slot[seg] : -1;
    if (slot < 0 ? !a.do_owned : !a.do_shared) continue;
    double dot = 0.0;
    if (a.phase) {
      for (int64_t k = a.ptr[seg] + lane; k < a.ptr[seg + 1]; k += LANES)
        dot = fma((double)__ldg(a.val + k), a.v[__ldg(a.idx + k)], dot);
      if (LANES > 1) dot = warp_sum(dot);
    }
    if (lane == 0) {
      if (slot >= 0) {
        a.qsh[slot] = dot;
      } else {
        const int row = a.segmap[seg];
        const double un = fma(cu, a.uc[row], cq * dot);
        if (a.phase) a.uc[row] = un;
        s = fma(un, un, s);
      }
    }
  }
  if (!a.do_owned) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.nempty; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = a.empty_rows[i];
    const double un = cu * a.uc[row];
    if (a.phase) a.uc[row] = un;
    s = fma(un, un, s);
  }
  double total;
  if (grid_sum_last(s, red, a.partial, a.ticket, total) && threadIdx.x == 0) *a.out = total;
}

struct DataArgs {
  double *u;
  const double *q;             // [0, nls): S vhat summed over ranks; [nls, nls + nshared): shared constraint rows
  int32_t nls, nshared;
  const int32_t *shared_rows;
  LsqrScalars *sc;
  int phase, init;
  const double *owned_sum;     // sum of squares of the rank-owned constraint rows, summed over ranks
  double *partial;
  unsigned *ticket;
};

// The replicated rows (data rows + shared constraint rows): uhat = cu*uhat + cq*q, then beta = |uhat| with the owned
// rows' share (lsqr_solver2.F90:194-222).
__global__ void __launch_bounds__(kVecThreads) k_data_rows(DataArgs a) {
  if (a.sc->done) return;
  __shared__ double red[32];
  const double cu = a.phase ? a.sc->cu : 1.0, cq = a.phase ? a.sc->cq : 0.0;
  const int64_t n = (int64_t)a.nls + a.nshared;
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = (i < a.nls) ? i : (int64_t)a.nls + a.shared_rows[i - a.nls];
    double un = a.u[row];
    if (a.phase) {
      un = fma(cu, un, cq * a.q[i]);
      a.u[row] = un;
    }
    s = fma(un, un, s);
  }
  double total;
  if (grid_sum_last(s, red, a.partial, a.ticket, total) && threadIdx.x == 0)
    scal_beta_update(a.sc, total + *a.owned_sum, a.init);
}

struct VArgs {
  double *v;
  const double *v2;
  int64_t c0, c1;              // active column window
  LsqrScalars *sc;
  int init, niter, single_matrix, finish;   // finish: single rank -> alpha and the recurrences right here
  double rmin;
  double *hist, *partial, *out;
  unsigned *ticket;
};

// vhat = cv*vhat + su*v2 (v = -beta v + S^T u + C^T u_c, lsqr_solver2.F90:225-238) and |vhat|^2 (:241).
__global__ void __launch_bounds__(kVecThreads) k_v_rows(VArgs a) {
  if (a.sc->done) {
    if (blockIdx.x == 0 && threadIdx.x == 0) a.sc->was_active = 0;   // k_xw of a dead iteration must not run
    return;
  }
  __shared__ double red[32];
  const double cv = a.sc->cv, su = a.sc->su;
  double s = 0.0;
  for (int64_t i = a.c0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.c1; i += (int64_t)gridDim.x * blockDim.x) {
    const double t = su * a.v2[i];
    const double vn = a.init ? t : fma(cv, a.v[i], t);
    a.v[i] = vn;
    s = fma(vn, vn, s);
  }
  double total;
  if (grid_sum_last(s, red, a.partial, a.ticket, total) && threadIdx.x == 0) {
    if (a.finish) scal_alpha_update(a.sc, total, a.init, a.niter, a.rmin, a.hist, a.single_matrix, 1);
    else *a.out = total;
  }
}

__global__ void k_scal_alpha(const double *sumsq, LsqrScalars *sc, int init, int niter, double rmin,
                             double *__restrict__ hist, int single_matrix, int unnorm) {
  if (sc->done) {
    sc->was_active = 0;
    return;
  }
  scal_alpha_update(sc, *sumsq, init, niter, rmin, hist, single_matrix, unnorm);
}

// x = t1*w + x ; w = t2*w + v ; optional soft threshold (lsqr_solver2.F90:269-275), v = sv*vhat on the fly
// (write_v: the fused path stores the normalised v, the dense sweep reads it).
__global__ void __launch_bounds__(kVecThreads) k_xw_update(double *__restrict__ v, double *__restrict__ x,
                                                            double *__restrict__ w, int64_t c0, int64_t c1,
                                                            const LsqrScalars *sc, int init, double gamma, int write_v) {
  if (!sc->was_active) return;
  const double ia = sc->sv;
  const int upd = sc->do_update;
  const double t1 = sc->t1, t2 = sc->t2;
  for (int64_t i = c0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < c1; i += (int64_t)gridDim.x * blockDim.x) {
    const double vi = (ia == 1.0) ? v[i] : ia * v[i];
    if (write_v) v[i] = vi;
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

// u *= factor on the rows this rank maintains: data rows, shared rows, the stored rows it owns, rank 0's empty rows.
__global__ void __launch_bounds__(kVecThreads) k_scale_rows(double *__restrict__ u, int32_t nls, const int32_t *shared_rows,
                                                             int32_t nshared, const int32_t *segmap, const int32_t *seg_slot,
                                                             int32_t nseg, const int32_t *empty_rows, int32_t nempty,
                                                             const double *factor, const int *done) {
  if (done && *done) return;
  const double f = *factor;
  if (f == 1.0) return;
  const int64_t n = (int64_t)nls + nshared + nseg + nempty;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row;
    if (i < nls) row = i;
    else if (i < (int64_t)nls + nshared) row = (int64_t)nls + shared_rows[i - nls];
    else if (i < (int64_t)nls + nshared + nseg) {
      const int64_t s = i - nls - nshared;
      if (seg_slot && seg_slot[s] >= 0) continue;    // shared: already scaled above
      row = (int64_t)nls + segmap[s];
    } else row = (int64_t)nls + empty_rows[i - nls - nshared - nseg];
    u[row] = f * u[row];
  }
}

// Fused path, columns of the active window NOT covered by the dense block: vhat = -beta v + g there, plus |vhat|^2.
__global__ void __launch_bounds__(kVecThreads) k_outside_update(double *__restrict__ v, const double *__restrict__ g,
                                                                 int64_t c0, int64_t c1, int64_t blk0, int64_t blk1,
                                                                 const LsqrScalars *sc, double *__restrict__ partial,
                                                                 unsigned *ticket, double *n2) {
  if (sc->done) return;
  __shared__ double red[32];
  const double nb = sc->neg_beta;
  double s = 0.0;
  for (int64_t i = c0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < c1; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= blk0 && i < blk1) continue;
    const double val = fma(nb, v[i], g ? g[i] : 0.0);
    v[i] = val;
    s = fma(val, val, s);
  }
  double total;
  if (grid_sum_last(s, red, partial, ticket, total) && threadIdx.x == 0) *n2 += total;
}

// misfit = sqrt(sum((Sx - b0)^2)/n) (lsqr_solver2.F90:183-188)
__global__ void __launch_bounds__(kVecThreads) k_misfit(const double *__restrict__ a, const double *__restrict__ b, int64_t n,
                                                         LsqrScalars *sc, double target, double *partial, unsigned *ticket) {
  if (sc->done) return;
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = a[i] - b[i];
    s = fma(d, d, s);
  }
  double total;
  if (grid_sum_last(s, red, partial, ticket, total) && threadIdx.x == 0) {
    sc->misfit = sqrt(total / (double)n);
    if (sc->misfit <= target) sc->done = 1;   // "Reached the target misfit, exiting the loop."
  }
}

// ---- ownership plan of the constraint rows -----------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mark_rows(const int32_t *__restrict__ segmap, int32_t nseg, uint8_t *cnt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nseg; i += gridDim.x * blockDim.x) cnt[segmap[i]] = 1;
}
__global__ void __launch_bounds__(256) k_seg_slots(const int32_t *__restrict__ segmap, int32_t nseg,
                                                    const uint8_t *__restrict__ cnt, const int32_t *__restrict__ shared_rows,
                                                    int32_t nshared, int32_t *__restrict__ seg_slot) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nseg; i += gridDim.x * blockDim.x) {
    const int row = segmap[i];
    int slot = -1;
    if (cnt[row] >= 2) {   // lower bound in the ascending list of shared rows
      int lo = 0, hi = nshared;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (shared_rows[mid] < row) lo = mid + 1;
        else hi = mid;
      }
      slot = lo;
    }
    seg_slot[i] = slot;
  }
}
struct CntIs {
  const uint8_t *cnt;
  int lo, hi;   // selects rows with lo <= cnt <= hi
  __device__ bool operator()(int i) const { return cnt[i] >= lo && cnt[i] <= hi; }
};

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

struct ConsPlan {
  int32_t ncons = 0, nseg = 0, nshared = 0, nempty = 0;
  int32_t win_lo = 0, win_hi = 0;   // constraint rows this rank reads / writes lie in [win_lo, win_hi)
  bool warp_rows = false;
  DevBuf<int32_t> shared_rows, seg_slot, empty_rows;
  DevBuf<uint8_t> cnt;
};

struct Work {
  DevBuf<double> v, w, v2, g, q, b0, sx, partial, red;
  DevBuf<unsigned> ticket;
  DevBuf<LsqrScalars> sc;
  DevBuf<double> hist;
  ConsPlan plan;
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

int read_i32(const int32_t *d, int32_t *h) {
  TFX_CUDA(cudaMemcpy(h, d, 4, cudaMemcpyDeviceToHost));
  return 0;
}

// Who keeps which constraint row. Single rank: every row is owned; rows without stored entries form the empty list.
// Several ranks: the per-row count of ranks holding entries is summed over the ranks (one uint8 all-reduce per solve);
// count 1 -> owned by that rank, >= 2 -> shared (replicated like a data row), 0 -> kept by rank 0.
int build_plan(ConsPlan &P, Matrix *C, int nranks, int rank, cudaStream_t st) {
  Context &c = ctx();
  P.ncons = C ? C->nl : 0;
  P.nseg = (C && C->has_seg) ? C->fwd.nseg : 0;
  P.nshared = P.nempty = 0;
  P.win_lo = P.win_hi = 0;
  P.warp_rows = P.nseg > 0 && C->fwd.avg_len > 24.0;
  if (P.ncons == 0) return 0;
  if (C && !C->has_seg && C->nel > 0) return fail(-52, "lsqr: the constraint matrix has no compressed-row representation");
  if (nranks == 1 && P.nseg == P.ncons) {   // every row stored (the usual damping / ADMM block)
    P.win_lo = 0; P.win_hi = P.ncons;
    return 0;
  }
  auto pol = thrust::cuda::par.on(st);
  TFX_TRY(P.cnt.alloc((size_t)P.ncons));
  TFX_CUDA(cudaMemsetAsync(P.cnt.p, 0, (size_t)P.ncons, st));
  if (P.nseg > 0) {
    k_mark_rows<<<std::min(c.num_sms * 8, (P.nseg + 255) / 256), 256, 0, st>>>(C->fwd.segmap.p, P.nseg, P.cnt.p);
    c.launches++;
  }
  if (nranks > 1) TFX_TRY(comm_allreduce_sum_u8(P.cnt.p, (size_t)P.ncons, st));
  thrust::counting_iterator<int> first(0);
  if (nranks > 1) {
    TFX_THRUST(P.nshared = (int32_t)thrust::count_if(pol, first, first + P.ncons, CntIs{P.cnt.p, 2, 255}));
    TFX_TRY(P.shared_rows.alloc((size_t)P.nshared));
    thrust::device_ptr<int32_t> out(P.shared_rows.p);
    if (P.nshared > 0) TFX_THRUST(thrust::copy_if(pol, first, first + P.ncons, out, CntIs{P.cnt.p, 2, 255}));
    c.launches += 2;
    if (P.nseg > 0) {
      TFX_TRY(P.seg_slot.alloc((size_t)P.nseg));
      k_seg_slots<<<std::min(c.num_sms * 8, (P.nseg + 255) / 256), 256, 0, st>>>(C->fwd.segmap.p, P.nseg, P.cnt.p,
                                                                                  P.shared_rows.p, P.nshared, P.seg_slot.p);
      c.launches++;
    }
  }
  if (rank == 0) {
    TFX_THRUST(P.nempty = (int32_t)thrust::count_if(pol, first, first + P.ncons, CntIs{P.cnt.p, 0, 0}));
    TFX_TRY(P.empty_rows.alloc((size_t)P.nempty));
    thrust::device_ptr<int32_t> out(P.empty_rows.p);
    if (P.nempty > 0) TFX_THRUST(thrust::copy_if(pol, first, first + P.ncons, out, CntIs{P.cnt.p, 0, 0}));
    c.launches += 2;
  }
  TFX_CUDA(cudaStreamSynchronize(st));
  P.cnt.release();   // ncons bytes: only needed to classify
  // window of the rows this rank touches (host buffers are copied in / out only there)
  int32_t lo = P.ncons, hi = 0, a, b;
  if (P.nseg > 0) {
    TFX_TRY(read_i32(C->fwd.segmap.p, &a)); TFX_TRY(read_i32(C->fwd.segmap.p + P.nseg - 1, &b));
    lo = std::min(lo, a); hi = std::max(hi, b + 1);
  }
  if (P.nshared > 0) {
    TFX_TRY(read_i32(P.shared_rows.p, &a)); TFX_TRY(read_i32(P.shared_rows.p + P.nshared - 1, &b));
    lo = std::min(lo, a); hi = std::max(hi, b + 1);
  }
  if (P.nempty > 0) {
    TFX_TRY(read_i32(P.empty_rows.p, &a)); TFX_TRY(read_i32(P.empty_rows.p + P.nempty - 1, &b));
    lo = std::min(lo, a); hi = std::max(hi, b + 1);
  }
  if (lo > hi) lo = hi = 0;
  P.win_lo = lo; P.win_hi = hi;
  return 0;
}

// Column extent [lo, hi) of the stored entries of a matrix (0-based), from its device representations.
int col_extent(Matrix &m, int64_t *lo, int64_t *hi) {
  if (m.has_blocks) {
    for (Matrix *b : m.blocks) TFX_TRY(col_extent(*b, lo, hi));
    return 0;
  }
  if (m.has_dense && !m.dense.empty()) {
    *lo = std::min<int64_t>(*lo, m.dense.col0);
    *hi = std::max<int64_t>(*hi, (int64_t)m.dense.col0 + m.dense.ncols);
  }
  if (m.has_seg && m.trn.nseg > 0) {
    int32_t a, b;
    TFX_TRY(read_i32(m.trn.segmap.p, &a)); TFX_TRY(read_i32(m.trn.segmap.p + m.trn.nseg - 1, &b));
    *lo = std::min<int64_t>(*lo, a); *hi = std::max<int64_t>(*hi, (int64_t)b + 1);
  } else if (m.has_t16 && m.t16t.valid && m.t16t.nseg > 0) {
    *lo = std::min<int64_t>(*lo, m.t16t.out0); *hi = std::max<int64_t>(*hi, (int64_t)m.t16t.out0 + m.t16t.nseg);
  }
  return 0;
}

// apply_wavelet_transform (src/inversion/wavelet_utils.F90:37-72): every active problem and component of
// v(nelements, ncomponents, 2) is one nx*ny*nz volume; with several ranks the slabs are assembled into the full
// volume on every GPU (wavelet_slab_device, data.cu) instead of the reference's gather to rank 0 / scatter.
int apply_wavelet(const LsqrParams &p, double *d_v, bool fwd, const std::vector<int64_t> &offsets, cudaStream_t st) {
  for (int i = 0; i < 2; ++i) {
    if (!p.solve_problem[i]) continue;
    for (int k = 0; k < p.ncomponents; ++k) {
      double *vol = d_v + ((size_t)i * p.ncomponents + k) * (size_t)p.nelements;
      TFX_TRY(wavelet_slab_device_off(vol, offsets, p.nx, p.ny, p.nz, p.compression_type, fwd, st));
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
  const int rank = comm_rank();

  // Sanity checks of the reference (lsqr_solver2.F90:85-89, :342-345).
  if (nls + ncons != nlines || S->ncolumns != ncol || (C && C->ncolumns != ncol))
    return fail(-50, p.single_matrix ? "Wrong matrix size in lsqr_solve! Exiting."
                                     : "Wrong matrix sizes in lsqr_solve_sensit! Exiting.");
  if (!S->finalized || (C && !C->finalized)) return fail(-51, "lsqr: matrix is not finalized");
  std::vector<int64_t> nsmaller;   // slab offsets of all ranks (wavelet inside the loop)
  if (wav) {
    TFX_TRY(comm_slab_offsets(p.nelements, nsmaller));
    if (nsmaller.back() != (int64_t)p.nx * p.ny * p.nz)
      return fail(-53, "lsqr: the ranks' nelements must add up to nx*ny*nz when the wavelet transform runs inside the loop");
  }

  // Host vectors: only the rows / columns this rank works on are copied (see below); parity mode copies everything.
  auto copy_in_all = [&]() -> int {
    if (p.host_u) TFX_CUDA(cudaMemcpyAsync(d_u, p.host_u, (size_t)nlines * 8, cudaMemcpyHostToDevice, st));
    return 0;
  };
  auto copy_out_all = [&]() -> int {
    if (p.host_u) TFX_CUDA(cudaMemcpyAsync(p.host_u, d_u, (size_t)nlines * 8, cudaMemcpyDeviceToHost, st));
    if (p.host_x) TFX_CUDA(cudaMemcpyAsync(p.host_x, d_x, (size_t)ncol * 8, cudaMemcpyDeviceToHost, st));
    TFX_CUDA(cudaStreamSynchronize(st));
    return 0;
  };

  if (g_opt_strict_order) {
    if (S->has_blocks) return fail(-57, "lsqr: strict_order needs the CSR copies, a row-blocked matrix has none");
    TFX_TRY(copy_in_all());
    TFX_TRY(lsqr_run_strict(p, S, C, d_u, d_x, res));
    return copy_out_all();
  }

  const bool dense_ok = S->has_dense && S->dense.nrows == nls && S->dense_row0 == 0 && S->dense.nrows <= kDenseMaxRows;
  if (S->has_dense && !dense_ok && !S->has_seg) return fail(-54, "lsqr: dense sensitivity block does not cover all data rows");
  if (!dense_ok && !S->has_t16 && !S->has_seg && !S->has_blocks)
    return fail(-55, "lsqr: the sensitivity matrix has no device representation");
  const bool fused = dense_ok && !wav && !misfit;
  res.fused = fused;
  res.history.clear();
  res.iters = 0;
  res.reported_iters = 0;
  res.status = 0;

  // ---- active column window: the columns of the problems being solved (joint_inverse_problem.F90:213-214 lays the two
  // problems side by side; the half of a single-problem run that belongs to the other problem is identically zero in
  // S, C, v, w and x). Verified against the matrices' column extents; any entry outside -> the whole range is swept.
  int64_t a0 = 0, a1 = ncol;
  if (!p.single_matrix && (int64_t)2 * p.ncomponents * p.nelements == ncol && (p.solve_problem[0] != 0) != (p.solve_problem[1] != 0)) {
    const int64_t half = (int64_t)p.ncomponents * p.nelements;
    const int64_t w0 = p.solve_problem[0] ? 0 : half, w1 = w0 + half;
    int64_t lo = ncol, hi = 0;
    TFX_TRY(col_extent(*S, &lo, &hi));
    if (C) TFX_TRY(col_extent(*C, &lo, &hi));
    if (lo >= hi || (lo >= w0 && hi <= w1)) { a0 = w0; a1 = w1; }
  }
  const int64_t nact = a1 - a0;

  ConsPlan &P = W.plan;
  TFX_TRY(build_plan(P, C, nranks, rank, st));
  const int32_t nshared = P.nshared;
  const int64_t nq = (int64_t)nls + nshared;       // all-reduced vector part of q, two scalar slots behind it

  Timing &T = timing();
  TFX_TRY(T.ensure());
  T.used = 0;
  const int GV = vec_grid(nact);
  const int GD = vec_grid(nq);
  const int GC = vec_grid(std::max<int64_t>((int64_t)P.nseg * (P.warp_rows ? 32 : 1), P.nempty));
  TFX_TRY(W.v.alloc(ncol)); TFX_TRY(W.w.alloc(ncol)); TFX_TRY(W.v2.alloc(ncol));
  if (fused) TFX_TRY(W.g.alloc(ncol));
  TFX_TRY(W.q.alloc(nq + 2));
  TFX_TRY(W.partial.alloc((size_t)c.num_sms * 8 + 8)); TFX_TRY(W.red.alloc(8));
  TFX_TRY(W.ticket.alloc(8));
  TFX_TRY(W.sc.alloc(1)); TFX_TRY(W.hist.alloc(std::max(1, p.niter)));
  if (misfit) { TFX_TRY(W.b0.alloc(nls)); TFX_TRY(W.sx.alloc(nls)); }
  double *v = W.v.p, *w = W.w.p, *v2 = W.v2.p, *g = W.g.p, *q = W.q.p, *partial = W.partial.p, *red = W.red.p;
  unsigned *ticket = W.ticket.p;
  LsqrScalars *sc = W.sc.p;
  const int *done = &sc->done;
  // q[nq] travels with the vector part of the all-reduce: |vhat|^2 in the fused path (known after the sweep), the owned
  // rows' |u|^2 in the split path; the fused path reduces the owned sum on its own (q[nq + 1]) once alpha is known.
  double *slot_n2 = q + nq, *slot_own = fused ? q + nq + 1 : q + nq;

  // ---- right-hand side in: data rows + this rank's window of the constraint rows
  if (p.host_u) {
    TFX_CUDA(cudaMemcpyAsync(d_u, p.host_u, (size_t)nls * 8, cudaMemcpyHostToDevice, st));
    if (P.win_hi > P.win_lo)
      TFX_CUDA(cudaMemcpyAsync(d_u + nls + P.win_lo, p.host_u + nls + P.win_lo, (size_t)(P.win_hi - P.win_lo) * 8,
                               cudaMemcpyHostToDevice, st));
  }
  TFX_CUDA(cudaMemsetAsync(v, 0, ncol * 8, st));
  TFX_CUDA(cudaMemsetAsync(w, 0, ncol * 8, st));
  TFX_CUDA(cudaMemsetAsync(v2, 0, ncol * 8, st));
  if (fused) TFX_CUDA(cudaMemsetAsync(g, 0, ncol * 8, st));
  TFX_CUDA(cudaMemsetAsync(q, 0, (nq + 2) * 8, st));
  TFX_CUDA(cudaMemsetAsync(d_x, 0, ncol * 8, st));                         // x = 0 (:120)
  TFX_CUDA(cudaMemsetAsync(sc, 0, sizeof(LsqrScalars), st));
  TFX_CUDA(cudaMemsetAsync(ticket, 0, 8 * sizeof(unsigned), st));
  TFX_CUDA(cudaMemsetAsync(W.hist.p, 0, std::max(1, p.niter) * 8, st));
  if (misfit) TFX_CUDA(cudaMemcpyAsync(W.b0.p, d_u, (size_t)nls * 8, cudaMemcpyDeviceToDevice, st));   // :110

#define LAUNCHED() c.launches++
  auto cons_rows = [&](int phase, const double *vin, bool owned, bool shared) -> int {
    if (ncons == 0) return 0;
    if (!owned && (!shared || nshared == 0)) return 0;
    if (P.nseg == 0 && P.nempty == 0) {   // this rank keeps no constraint row: its share of the owned sum is 0
      if (owned && nranks > 1) TFX_CUDA(cudaMemsetAsync(slot_own, 0, 8, st));
      return 0;
    }
    ConsArgs a;
    a.ptr = P.nseg ? C->fwd.ptr.p : nullptr; a.idx = P.nseg ? C->fwd.idx.p : nullptr; a.val = P.nseg ? C->fwd.val.p : nullptr;
    a.segmap = P.nseg ? C->fwd.segmap.p : nullptr;
    a.seg_slot = (nranks > 1 && P.nseg) ? P.seg_slot.p : nullptr;
    a.nseg = P.nseg; a.empty_rows = P.empty_rows.p; a.nempty = P.nempty;
    a.uc = d_u + nls; a.v = vin; a.qsh = q + nls; a.sc = sc; a.phase = phase;
    a.do_owned = owned ? 1 : 0; a.do_shared = shared ? 1 : 0;
    a.partial = partial; a.ticket = ticket; a.out = slot_own;
    if (P.warp_rows) k_cons_rows<32><<<GC, kVecThreads, 0, st>>>(a);
    else k_cons_rows<1><<<GC, kVecThreads, 0, st>>>(a);
    LAUNCHED();
    return 0;
  };
  auto data_rows = [&](int phase, int init) {
    DataArgs a;
    a.u = d_u; a.q = q; a.nls = nls; a.nshared = nshared; a.shared_rows = P.shared_rows.p; a.sc = sc;
    a.phase = phase; a.init = init; a.owned_sum = slot_own; a.partial = partial; a.ticket = ticket + 1;
    k_data_rows<<<GD, kVecThreads, 0, st>>>(a);
    LAUNCHED();
  };
  auto scale_rows = [&](const double *factor, const int *dn) {
    const int64_t n = (int64_t)nls + nshared + P.nseg + P.nempty;
    k_scale_rows<<<vec_grid(n), kVecThreads, 0, st>>>(d_u, nls, P.shared_rows.p, nshared,
                                                      P.nseg ? C->fwd.segmap.p : nullptr,
                                                      (nranks > 1 && P.nseg) ? P.seg_slot.p : nullptr, P.nseg,
                                                      P.empty_rows.p, P.nempty, factor, dn);
    LAUNCHED();
  };

  // ---- beta = |u| ; b1 = beta (:123-134). u itself is normalised only in the fused path.
  TFX_TRY(cons_rows(0, nullptr, true, false));
  if (nranks > 1 && ncons > 0) TFX_TRY(comm_allreduce_sum(slot_own, 1, st));
  data_rows(0, 1);

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
  int host_done = 0;
  const int chk = std::max(1, g_opt_lsqr_poll);
  auto poll = [&](int it) -> int {
    if (it % chk == 0 || it == p.niter) {
      TFX_CUDA(cudaMemcpyAsync(&host_done, &sc->done, sizeof(int), cudaMemcpyDeviceToHost, st));
      TFX_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
  };

  if (fused) {
    // =========================== FUSED PATH ===========================
    scale_rows(&sc->inv_beta, done);                                         // u = u / beta
    const int64_t blk0 = S->dense.col0, blk1 = (int64_t)S->dense.col0 + S->dense.ncols;
    const bool outside = have_C && (blk0 > a0 || blk1 < a1);
    auto sweep = [&]() -> int {
      // g = C^T u_c ; vhat = -beta v + S^T u_d + g ; q_d = S vhat ; n2 = |vhat|^2 ; shared rows of C vhat
      if (have_C) TFX_TRY(seg_spmv(C->trn, d_u + nls, g, false, 0, (int32_t)ncol, 0, done, st));
      cudaEvent_t ea = nullptr, eb = nullptr;
      if (g_opt_profile_sweeps && T.used < 4096) {
        TFX_TRY(T.next(&ea)); TFX_TRY(T.next(&eb));
        TFX_CUDA(cudaEventRecord(ea, st));
      }
      TFX_TRY(dense_sweep(S->dense, DENSE_FUSED, d_u, v, have_C ? g : nullptr, v, &sc->neg_beta, q, slot_n2, done, st));
      if (eb) TFX_CUDA(cudaEventRecord(eb, st));
      if (outside) {
        k_outside_update<<<GV, kVecThreads, 0, st>>>(v, g, a0, a1, blk0, blk1, sc, partial, ticket + 2, slot_n2); LAUNCHED();
      }
      TFX_TRY(cons_rows(1, v, false, true));                                  // shared rows: partial dots with vhat
      if (nranks > 1) {
        TFX_TRY(comm_allreduce_sum(q, (size_t)nq + 1, st));   // one collective: data rows, shared rows, |vhat|^2
      }
      return 0;
    };
    // init: neg_beta multiplies v = 0, so vhat = S^T u (+ C^T u_c)
    TFX_TRY(sweep());
    k_scal_alpha<<<1, 1, 0, st>>>(slot_n2, sc, 1, p.niter, p.rmin, W.hist.p, p.single_matrix ? 1 : 0, 0); LAUNCHED();
    TFX_CUDA(cudaEventRecord(T.loop0, st));
    for (int it = 1; it <= p.niter && !host_done; ++it) {
      // u_c = -alpha u_c + (C vhat)/alpha on the owned rows (vhat still un-normalised), their |u|^2 summed over ranks;
      // then v = vhat/alpha, x, w; then the replicated rows and beta; u /= beta
      TFX_TRY(cons_rows(1, v, true, false));
      if (nranks > 1 && ncons > 0) TFX_TRY(comm_allreduce_sum(slot_own, 1, st));
      k_xw_update<<<GV, kVecThreads, 0, st>>>(v, d_x, w, a0, a1, sc, it == 1 ? 1 : 0, p.gamma, 1); LAUNCHED();
      data_rows(1, 0);
      scale_rows(&sc->inv_beta, done);
      TFX_TRY(sweep());
      k_scal_alpha<<<1, 1, 0, st>>>(slot_n2, sc, 0, p.niter, p.rmin, W.hist.p, p.single_matrix ? 1 : 0, 0); LAUNCHED();
      TFX_TRY(poll(it));
    }
    // the x/w update of the last executed iteration
    k_xw_update<<<GV, kVecThreads, 0, st>>>(v, d_x, w, a0, a1, sc, 0, p.gamma, 1); LAUNCHED();
  } else {
    // =========================== SPLIT PATH (reference order, deferred normalisation) ===========================
    // vhat = (S^T uhat_d [inverse wavelet] + C^T uhat_c) * su ; alpha = |vhat| ; w = v   (:137-157)
    auto v_rows = [&](int init) {
      VArgs a;
      a.v = v; a.v2 = v2; a.c0 = a0; a.c1 = a1; a.sc = sc; a.init = init; a.niter = p.niter;
      a.single_matrix = p.single_matrix ? 1 : 0; a.finish = (nranks == 1) ? 1 : 0; a.rmin = p.rmin;
      a.hist = W.hist.p; a.partial = partial; a.out = red; a.ticket = ticket + 2;
      k_v_rows<<<GV, kVecThreads, 0, st>>>(a);
      LAUNCHED();
    };
    auto v_step = [&](int init) -> int {
      TFX_TRY(S_trans(d_u, v2));
      if (wav) TFX_TRY(apply_wavelet(p, v2, false, nsmaller, st));
      if (have_C) TFX_TRY(seg_spmv(C->trn, d_u + nls, v2, true, 0, (int32_t)ncol, 0, done, st));
      v_rows(init);
      if (nranks > 1) {
        TFX_TRY(comm_allreduce_sum(red, 1, st));
        k_scal_alpha<<<1, 1, 0, st>>>(red, sc, init, p.niter, p.rmin, W.hist.p, p.single_matrix ? 1 : 0, 1); LAUNCHED();
      }
      k_xw_update<<<GV, kVecThreads, 0, st>>>(v, d_x, w, a0, a1, sc, init, p.gamma, 0); LAUNCHED();
      return 0;
    };
    TFX_TRY(v_step(1));
    TFX_CUDA(cudaEventRecord(T.loop0, st));
    // One iteration of the reference's loop body (:160-290) as a stream of launches.
    auto iter_body = [&]() -> int {
      if (misfit) {   // :168-189
        TFX_CUDA(cudaMemcpyAsync(v2, d_x, ncol * 8, cudaMemcpyDeviceToDevice, st));
        if (wav) TFX_TRY(apply_wavelet(p, v2, true, nsmaller, st));
        TFX_TRY(S_fwd(v2, W.sx.p));
        if (nranks > 1) TFX_TRY(comm_allreduce_sum(W.sx.p, nls, st));
        k_misfit<<<vec_grid(nls), kVecThreads, 0, st>>>(W.sx.p, W.b0.p, nls, sc, p.target_misfit, partial, ticket + 3); LAUNCHED();
      }
      // q = [S W(vhat); C vhat]  (:200-211), summed over ranks (:214); uhat = cu*uhat + cq*q ; beta = |uhat| (:194-222)
      const double *vin = v;
      if (wav) {
        TFX_CUDA(cudaMemcpyAsync(v2, v, ncol * 8, cudaMemcpyDeviceToDevice, st));
        TFX_TRY(apply_wavelet(p, v2, true, nsmaller, st));
        vin = v2;
      }
      TFX_TRY(S_fwd(vin, q));
      TFX_TRY(cons_rows(1, v, true, true));
      if (nranks > 1) TFX_TRY(comm_allreduce_sum(q, (size_t)nq + 1, st));
      data_rows(1, 0);
      // vhat = cv*vhat + su*(W^-1(S^T uhat_d) + C^T uhat_c) ; alpha = |vhat| (:225-245) ; x, w (:269-275)
      TFX_TRY(v_step(0));
      return 0;
    };
    // Launch-bound regime (small matrices: a dozen launches of a few microseconds each per iteration): the body is
    // captured once into a CUDA graph -- after a first, directly launched iteration has done every lazy allocation --
    // and replayed; all kernel arguments are iteration-independent (the scalars live in *sc on the device). NCCL
    // collectives are captured like kernels, so the multi-rank body replays too.
    cudaGraphExec_t gexec = nullptr;
    unsigned long long launches_per_iter = 0;
    const bool want_graph = g_opt_lsqr_graph && !S->has_blocks && p.niter >= 4 && S->device_nnz() <= (int64_t)2e8;
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
      TFX_TRY(poll(it));
    }
    if (gexec) cudaGraphExecDestroy(gexec);
  }
  TFX_CUDA(cudaEventRecord(T.loop1, st));
  // the reference leaves the normalised u behind (it is the solver's work array, destroyed): same here
  if (!fused) scale_rows(&sc->su, nullptr);
#undef LAUNCHED
  TFX_CUDA(cudaGetLastError());
  LsqrScalars h;
  TFX_CUDA(cudaMemcpyAsync(&h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
  // ---- results out: x on the active window (zero elsewhere), u on the rows this rank maintains
  if (p.host_x) {
    if (a0 > 0) memset(p.host_x, 0, (size_t)a0 * 8);
    if (a1 < ncol) memset(p.host_x + a1, 0, (size_t)(ncol - a1) * 8);
    TFX_CUDA(cudaMemcpyAsync(p.host_x + a0, d_x + a0, (size_t)nact * 8, cudaMemcpyDeviceToHost, st));
  }
  if (p.host_u) {
    TFX_CUDA(cudaMemcpyAsync(p.host_u, d_u, (size_t)nls * 8, cudaMemcpyDeviceToHost, st));
    if (P.win_hi > P.win_lo)
      TFX_CUDA(cudaMemcpyAsync(p.host_u + nls + P.win_lo, d_u + nls + P.win_lo, (size_t)(P.win_hi - P.win_lo) * 8,
                               cudaMemcpyDeviceToHost, st));
  }
  TFX_CUDA(cudaStreamSynchronize(st));
  res.iters = h.executed;
  res.reported_iters = h.iter - 1;
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
