// weights.cu -- calculate_depth_weight (src/forward/gravmag/weights_gravmag.f90:46-199) on the device.
//
// Types 2 (distance weighting, Li & Oldenburg 2000 Eq. 19) and 3 (minimum distance) are O(ncells * ndata) with
// 8 sqrt + 8 pow per (cell, station): one thread per cell, stations staged through shared memory, the sum over
// stations taken in the reference's order (j ascending) so the result does not depend on the launch shape.
#include "../../include/tfx.h"

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

namespace tfx {
namespace {

constexpr int kWThreads = 128;
constexpr int kWStations = 512;     // stations per shared-memory chunk (3 x 4 KB)

__device__ __forceinline__ double cell_volume(double x1, double x2, double y1, double y2, double z1, double z2) {
  return fabs(__dmul_rn(__dmul_rn(x2 - x1, y2 - y1), z2 - z1));      // grid.F90:284-293
}

// weight(i) for cells [cell0, cell0 + n) before normalisation, already multiplied by sqrt(volume) (:71-175).
template <int TYPE>
__global__ void __launch_bounds__(kWThreads) k_depth_weight(const double *__restrict__ X1, const double *__restrict__ X2,
                                                            const double *__restrict__ Y1, const double *__restrict__ Y2,
                                                            const double *__restrict__ Z1, const double *__restrict__ Z2,
                                                            int64_t cell0, int64_t n, int32_t ndata,
                                                            const double *__restrict__ xd, const double *__restrict__ yd,
                                                            const double *__restrict__ zd, double power, double beta,
                                                            double Z0, double *__restrict__ w, int *__restrict__ err) {
  __shared__ double sx[TYPE == 1 ? 1 : kWStations], sy[TYPE == 1 ? 1 : kWStations], sz[TYPE == 1 ? 1 : kWStations];
  const int64_t i = blockIdx.x * (int64_t)kWThreads + threadIdx.x;
  const bool live = i < n;
  const int64_t p = cell0 + (live ? i : 0);
  const double x1 = X1[p], x2 = X2[p], y1 = Y1[p], y2 = Y2[p], z1 = Z1[p], z2 = Z2[p];
  const double vol = cell_volume(x1, x2, y1, y2, z1, z2);
  double weight = 0.0;
  if (TYPE == 1) {
    const double depth = 0.5 * (z1 + z2);                            // :204-223
    if (depth + Z0 > 0.0) weight = pow(depth + Z0, -power / 2.0);
    else if (live) atomicCAS(err, 0, 1);
  } else if (TYPE == 2) {
    const double R0 = 0.1, dfactor = 0.25;                           // :86-90
    const double dhx = dfactor * fabs(x2 - x1), dhy = dfactor * fabs(y2 - y1), dhz = dfactor * fabs(z2 - z1);
    const double ax = x1 + dhx, bx = x2 - dhx, ay = y1 + dhy, by = y2 - dhy, az = z1 + dhz, bz = z2 - dhz;
    double wr = 0.0;
    for (int32_t j0 = 0; j0 < ndata; j0 += kWStations) {
      const int32_t m = min(kWStations, ndata - j0);
      __syncthreads();
      for (int32_t t = threadIdx.x; t < m; t += kWThreads) { sx[t] = xd[j0 + t]; sy[t] = yd[j0 + t]; sz[t] = zd[j0 + t]; }
      __syncthreads();
      for (int32_t t = 0; t < m; ++t) {
        double dX[2], dY[2], dZ[2];
        dX[0] = __dmul_rn(ax - sx[t], ax - sx[t]); dX[1] = __dmul_rn(bx - sx[t], bx - sx[t]);
        dY[0] = __dmul_rn(ay - sy[t], ay - sy[t]); dY[1] = __dmul_rn(by - sy[t], by - sy[t]);
        dZ[0] = __dmul_rn(az - sz[t], az - sz[t]); dZ[1] = __dmul_rn(bz - sz[t], bz - sz[t]);
        double integral = 0.0;
#pragma unroll
        for (int ii = 0; ii < 2; ++ii)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const double R = sqrt(__dadd_rn(__dadd_rn(dX[ii], dY[jj]), dZ[kk]));
              integral = __dadd_rn(integral, 1.0 / pow(R + R0, power));
            }
        integral = __dmul_rn(integral, vol) / 8.0;
        wr = __dadd_rn(wr, __dmul_rn(integral, integral));
      }
    }
    weight = __dmul_rn(1.0 / sqrt(vol), pow(wr, beta / 4.0));        // :137
  } else {
    const double R0 = 0.01;                                          // :140-161
    const double cx = 0.5 * (x1 + x2), cy = 0.5 * (y1 + y2), cz = 0.5 * (z1 + z2);
    double mindist = 1.e30;
    for (int32_t j0 = 0; j0 < ndata; j0 += kWStations) {
      const int32_t m = min(kWStations, ndata - j0);
      __syncthreads();
      for (int32_t t = threadIdx.x; t < m; t += kWThreads) { sx[t] = xd[j0 + t]; sy[t] = yd[j0 + t]; sz[t] = zd[j0 + t]; }
      __syncthreads();
      for (int32_t t = 0; t < m; ++t) {
        const double a = cx - sx[t], b = cy - sy[t], c = cz - sz[t];
        const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
        if (dist < mindist) mindist = dist;
      }
    }
    weight = sqrt(1.0 / pow(mindist + R0, power));
  }
  if (live) w[i] = __dmul_rn(weight, sqrt(vol));                     // :170-175
}

// max over the slab (normalize_depth_weight, :228-250): per-CTA maxima, then one CTA.
__global__ void __launch_bounds__(256) k_max_partial(const double *__restrict__ w, int64_t n, double *__restrict__ part) {
  __shared__ double red[256];
  double m = -1.e300;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) m = fmax(m, w[i]);
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256) k_max_final(const double *__restrict__ part, int np, double *__restrict__ out) {
  __shared__ double red[256];
  double m = -1.e300;
  for (int i = threadIdx.x; i < np; i += 256) m = fmax(m, part[i]);
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}
// w = 1 / (w / norm), flags: 2 zero norm, 3 zero weight (:177-195, :240-247)
__global__ void __launch_bounds__(256) k_normalize_invert(double *__restrict__ w, int64_t n, const double *__restrict__ norm,
                                                          int *__restrict__ err) {
  const double nrm = norm[0];
  if (nrm == 0.0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicCAS(err, 0, 2);
    return;
  }
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double v = __ddiv_rn(w[i], nrm);
    if (v != 0.0) w[i] = __ddiv_rn(1.0, v);
    else atomicCAS(err, 0, 3);
  }
}

}  // namespace
}  // namespace tfx

using namespace tfx;

extern "C" int tfx_calculate_depth_weight(int32_t depth_weighting_type, double depth_weighting_power,
                                          double depth_weighting_beta, double Z0, int32_t nelements_total,
                                          const double *X1, const double *X2, const double *Y1, const double *Y2,
                                          const double *Z1, const double *Z2, int32_t ndata, const double *data_X,
                                          const double *data_Y, const double *data_Z, int32_t nsmaller,
                                          int32_t nelements, double *column_weight, int32_t myrank, int32_t nbproc) {
  (void)myrank;
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  if (depth_weighting_type < 1 || depth_weighting_type > 3) return fail(-97, "Not known depth weight type!");
  if (nsmaller < 0 || nelements < 0 || (int64_t)nsmaller + nelements > nelements_total)
    return fail(-97, "calculate_depth_weight: wrong cell slab");
  if (nbproc > 1 && comm_nranks() != nbproc)
    return fail(-24, "calculate_depth_weight: nbproc does not match the communicator (tfx_comm_init)");
  const size_t N = (size_t)nelements_total;
  VecIO g[6], d[3], w;
  const double *gp[6] = {X1, X2, Y1, Y2, Z1, Z2};
  for (int k = 0; k < 6; ++k) TFX_TRY(g[k].bind(const_cast<double *>(gp[k]), N, true));
  const double *dp[3] = {data_X, data_Y, data_Z};
  for (int k = 0; k < 3; ++k) TFX_TRY(d[k].bind(const_cast<double *>(dp[k]), (size_t)std::max(ndata, 1), ndata > 0));
  TFX_TRY(w.bind(column_weight, (size_t)nelements, false));
  DevBuf<int> err;
  TFX_TRY(err.alloc(1));
  TFX_TRY(err.zero());
  const int np = std::max(1, std::min<int>((int)((nelements + 255) / 256), c.num_sms * 8));
  DevBuf<double> part;
  TFX_TRY(part.alloc((size_t)np + 1));
  if (nelements > 0) {
    const unsigned grid = (unsigned)(((int64_t)nelements + kWThreads - 1) / kWThreads);
#define TFX_W_LAUNCH(T)                                                                                               \
  k_depth_weight<T><<<grid, kWThreads, 0, st>>>(g[0].dev, g[1].dev, g[2].dev, g[3].dev, g[4].dev, g[5].dev, nsmaller, \
                                                nelements, ndata, d[0].dev, d[1].dev, d[2].dev, depth_weighting_power, \
                                                depth_weighting_beta, Z0, w.dev, err.p)
    if (depth_weighting_type == 1) TFX_W_LAUNCH(1);
    else if (depth_weighting_type == 2) TFX_W_LAUNCH(2);
    else TFX_W_LAUNCH(3);
#undef TFX_W_LAUNCH
    c.launches++;
  }
  k_max_partial<<<np, 256, 0, st>>>(w.dev, nelements, part.p);
  k_max_final<<<1, 256, 0, st>>>(part.p, np, part.p + np);
  c.launches += 2;
  if (nbproc > 1) TFX_TRY(comm_allreduce_max(part.p + np, 1, st));   // mpi_allreduce(MPI_MAX), :237-240
  k_normalize_invert<<<np, 256, 0, st>>>(w.dev, nelements, part.p + np, err.p);
  c.launches++;
  int h_err = 0;
  TFX_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  TFX_CUDA(cudaGetLastError());
  if (h_err == 1) return fail(-97, "Error: non-positive depth in calc_depth_weight_pixel!");
  if (h_err == 2) return fail(-97, "Zero depth weight norm! Exiting.");
  if (h_err == 3) return fail(-97, "Zero damping weight! Exiting.");
  TFX_TRY(w.copy_back());
  return 0;
}
