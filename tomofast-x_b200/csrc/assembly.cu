// assembly.cu -- forward-kernel evaluation on the device (right rectangular prisms).
//
// Replaces graviprism_z / gradiprism_zz / gradiprism_full (src/forward/gravmag/grav/gravity_field.f90:131-195, :314-364,
// :207-309)
// and magprism / sharmbox (src/forward/gravmag/mag/magnetic_field.f90:118-457) together with the
// weighting / real(4) rounding steps of calculate_and_write_sensit (sensitivity_gravmag.F90:228,290)
// and read_sensitivity_kernel (:837-843).
//
// This stage is FP64-transcendental bound (8 corners x {sqrt, atan2, 2 log} per cell and station),
// not memory bound and not a GEMM: tensor cores do not apply. One thread evaluates one
// (station, cell) pair; stations vary fastest across a warp so that the cell box is a broadcast
// load and the column-major store of the dense block is coalesced.
#include "common.cuh"
#include "kernels.h"
#include "mathx.cuh"

#include <math.h>

#include <algorithm>

namespace tfx {

// PI, src/global_typedefs.F90:52.
#define TFX_PI 3.1415926535897932385

// graviprism_z for one cell / one station, gravity_field.f90:151-192. Returns gz (without G).
__device__ __forceinline__ double grav_gz(double x1, double x2, double y1, double y2, double z1, double z2, double xd,
                                          double yd, double zd, int *err) {
  const double twopi = 2.0 * TFX_PI;
  const double XX[2] = {xd - x1, xd - x2};
  const double YY[2] = {yd - y1, yd - y2};
  const double ZZ[2] = {zd - z1, zd - z2};
  double gz = 0.0;
#pragma unroll
  for (int K = 0; K < 2; ++K)
#pragma unroll
    for (int L = 0; L < 2; ++L)
#pragma unroll
      for (int M = 0; M < 2; ++M) {
        const double dmu = ((K + L + M) & 1) ? 1.0 : -1.0;   // signo(K)*signo(L)*signo(M), signo = (-1, +1)
        const double Rs =
            sqrt(__dadd_rn(__dadd_rn(__dmul_rn(XX[K], XX[K]), __dmul_rn(YY[L], YY[L])), __dmul_rn(ZZ[M], ZZ[M])));
        double arg3 = tfx_atan2(__dmul_rn(XX[K], YY[L]), __dmul_rn(ZZ[M], Rs));
        if (arg3 < 0) arg3 = arg3 + twopi;
        double arg4 = Rs + XX[K];
        double arg5 = Rs + YY[L];
        if (arg4 <= 0.) *err = 1;   // "Data coordinate coincides with model grid boundary (YZ)"
        if (arg5 <= 0.) *err = 2;   // "... (XZ)"
        arg4 = tfx_log(arg4);
        arg5 = tfx_log(arg5);
        const double term = __dsub_rn(__dsub_rn(__dmul_rn(ZZ[M], arg3), __dmul_rn(XX[K], arg5)), __dmul_rn(YY[L], arg4));
        gz = __dadd_rn(gz, __dmul_rn(dmu, term));
      }
  return gz;
}

// gradiprism_zz, gravity_field.f90:331-361.
__device__ __forceinline__ double grav_gzz(double x1, double x2, double y1, double y2, double z1, double z2, double xd,
                                           double yd, double zd) {
  const double twopi = 2.0 * TFX_PI;
  const double XX[2] = {xd - x1, xd - x2};
  const double YY[2] = {yd - y1, yd - y2};
  const double ZZ[2] = {-(zd - z1), -(zd - z2)};
  double gzz = 0.0;
#pragma unroll
  for (int K = 0; K < 2; ++K)
#pragma unroll
    for (int L = 0; L < 2; ++L)
#pragma unroll
      for (int M = 0; M < 2; ++M) {
        const double dmu = ((K + L + M) & 1) ? 1.0 : -1.0;
        const double Rs =
            sqrt(__dadd_rn(__dadd_rn(__dmul_rn(XX[K], XX[K]), __dmul_rn(YY[L], YY[L])), __dmul_rn(ZZ[M], ZZ[M])));
        double vzz = -tfx_atan2(__dmul_rn(XX[K], YY[L]), __dmul_rn(Rs, ZZ[M]));
        if (vzz < 0) vzz = vzz + twopi;
        gzz = __dadd_rn(gzz, __dmul_rn(dmu, vzz));
      }
  return gzz;
}

// gradiprism_full, gravity_field.f90:207-309: the six tensor components of one prism, out[] in the order the caller stores
// them (sensitivity_gravmag.F90:207-209): XX, YY, ZZ, XY, YZ, ZX. err: 3 zero denominator (:271-273), 4 bad log argument
// (:278-280). Same operation order as the reference, no FMA contraction.
__device__ __forceinline__ void grav_full(double x1, double x2, double y1, double y2, double z1, double z2, double xd,
                                          double yd, double zd, double (&out)[6], int *err) {
  const double twopi = 2.0 * TFX_PI;
  const double XX[2] = {xd - x1, xd - x2};
  const double YY[2] = {yd - y1, yd - y2};
  const double ZZ[2] = {-(zd - z1), -(zd - z2)};
  double gxx = 0.0, gxy = 0.0, gyy = 0.0, gzx = 0.0, gyz = 0.0, gzz = 0.0;
#pragma unroll
  for (int K = 0; K < 2; ++K)
#pragma unroll
    for (int L = 0; L < 2; ++L)
#pragma unroll
      for (int M = 0; M < 2; ++M) {
        const double dmu = ((K + L + M) & 1) ? 1.0 : -1.0;
        const double xx = __dmul_rn(XX[K], XX[K]), zz = __dmul_rn(ZZ[M], ZZ[M]);
        const double Rs = sqrt(__dadd_rn(__dadd_rn(xx, __dmul_rn(YY[L], YY[L])), zz));
        const double xy = __dmul_rn(XX[K], YY[L]), rz = __dmul_rn(Rs, ZZ[M]);
        double vxx = tfx_atan2(xy, __dadd_rn(__dadd_rn(xx, rz), zz));
        double vyy = tfx_atan2(xy, __dsub_rn(__dadd_rn(__dmul_rn(Rs, Rs), rz), xx));
        double vzz = -tfx_atan2(xy, rz);
        if (vxx < 0) vxx = vxx + twopi;
        if (vyy < 0) vyy = vyy + twopi;
        if (vzz < 0) vzz = vzz + twopi;
        const double arg1 = Rs + ZZ[M];
        const double arg21 = Rs - YY[L], arg22 = Rs + YY[L];
        const double arg31 = Rs - XX[K], arg32 = Rs + XX[K];
        if (arg22 == 0. || arg32 == 0.) { *err = 3; continue; }
        const double arg2 = __ddiv_rn(arg21, arg22);
        const double arg3 = __ddiv_rn(arg31, arg32);
        if (arg1 <= 0. || arg2 <= 0. || arg3 <= 0.) { *err = 4; continue; }
        const double vxy = tfx_log(arg1);
        const double vzx = __dmul_rn(0.5, tfx_log(arg2));
        const double vyz = __dmul_rn(0.5, tfx_log(arg3));
        gxx = __dadd_rn(gxx, __dmul_rn(dmu, vxx));
        gyy = __dadd_rn(gyy, __dmul_rn(dmu, vyy));
        gzz = __dadd_rn(gzz, __dmul_rn(dmu, vzz));
        gxy = __dadd_rn(gxy, __dmul_rn(dmu, vxy));
        gyz = __dadd_rn(gyz, __dmul_rn(dmu, vyz));
        gzx = __dadd_rn(gzx, __dmul_rn(dmu, vzx));
      }
  out[0] = gxx; out[1] = gyy; out[2] = gzz; out[3] = gxy; out[4] = gyz; out[5] = gzx;
}

// One corner term of graviprism_z (gravity_field.f90:163-188): Z*atan2(X*Y, Z*R) - X*log(R+Y) - Y*log(R+X).
__device__ __forceinline__ double grav_corner_term(double X, double Y, double Z, int *err) {
  const double twopi = 2.0 * TFX_PI;
  const double Rs = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(X, X), __dmul_rn(Y, Y)), __dmul_rn(Z, Z)));
  double arg3 = tfx_atan2(__dmul_rn(X, Y), __dmul_rn(Z, Rs));
  if (arg3 < 0) arg3 = arg3 + twopi;
  double arg4 = Rs + X;
  double arg5 = Rs + Y;
  if (arg4 <= 0.) *err = 1;   // "Data coordinate coincides with model grid boundary (YZ)"
  if (arg5 <= 0.) *err = 2;   // "... (XZ)"
  arg4 = tfx_log(arg4);
  arg5 = tfx_log(arg5);
  return __dsub_rn(__dsub_rn(__dmul_rn(Z, arg3), __dmul_rn(X, arg5)), __dmul_rn(Y, arg4));
}

// G_grav = 6.674e-11 is a single-precision literal in the reference (gravity_field.f90:26).
__device__ __forceinline__ double g_grav() { return (double)6.674e-11f; }

// ---------------------------------------------------------------------------------------------
// Dense (no compression) block: S[col*ld + row], one thread per (row, col).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grav_dense_kernel(float *__restrict__ S, long long ld, int nrows, int ncols,
                                                         int cell0, const double *__restrict__ X1,
                                                         const double *__restrict__ X2, const double *__restrict__ Y1,
                                                         const double *__restrict__ Y2, const double *__restrict__ Z1,
                                                         const double *__restrict__ Z2, const double *__restrict__ xd,
                                                         const double *__restrict__ yd, const double *__restrict__ zd,
                                                         const double *__restrict__ cw, const double *__restrict__ dw,
                                                         double problem_weight, int cols_per_block, int *err) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const double px = xd[row], py = yd[row], pz = zd[row];
  // combined_weight = real(problem_weight * data_weight, 4), sensitivity_gravmag.F90:837.
  const float wgt = (float)(problem_weight * dw[row]);
  const int cbeg = blockIdx.y * cols_per_block;
  const int cend = min(ncols, cbeg + cols_per_block);
  int e = 0;
  // Consecutive cells of a structured grid share a face: the four corner terms of the x2 face of cell c are, bit for
  // bit, the four terms of the x1 face of cell c+1 (same station, same corner coordinates), so they are carried over
  // instead of being recomputed -- half the sqrt/atan2/log work; the sum keeps the reference's order (K outer).
  double carry[4] = {0.0, 0.0, 0.0, 0.0};
  double px2 = 0.0, py1 = 0.0, py2 = 0.0, pz1 = 0.0, pz2 = 0.0;
  bool have = false;
  for (int c = cbeg; c < cend; ++c) {
    const int p = cell0 + c;
    const double x1 = X1[p], x2 = X2[p], y1 = Y1[p], y2 = Y2[p], z1 = Z1[p], z2 = Z2[p];
    const double XX[2] = {px - x1, px - x2}, YY[2] = {py - y1, py - y2}, ZZ[2] = {pz - z1, pz - z2};
    const bool reuse = have && x1 == px2 && y1 == py1 && y2 == py2 && z1 == pz1 && z2 == pz2;
    double gz = 0.0;
#pragma unroll
    for (int L = 0; L < 2; ++L)
#pragma unroll
      for (int M = 0; M < 2; ++M) {
        const double t = reuse ? carry[2 * L + M] : grav_corner_term(XX[0], YY[L], ZZ[M], &e);
        gz = __dadd_rn(gz, ((L + M) & 1) ? t : -t);              // dmu = signo(1)*signo(L)*signo(M), K = 1 -> -1
      }
#pragma unroll
    for (int L = 0; L < 2; ++L)
#pragma unroll
      for (int M = 0; M < 2; ++M) {
        const double t = grav_corner_term(XX[1], YY[L], ZZ[M], &e);
        carry[2 * L + M] = t;
        gz = __dadd_rn(gz, ((1 + L + M) & 1) ? t : -t);
      }
    px2 = x2; py1 = y1; py2 = y2; pz1 = z1; pz2 = z2; have = true;
    const double line = __dmul_rn(__dmul_rn(g_grav(), gz), cw[p]);   // LineZ = G*gz (:192); * column weight (:1051)
    const float v = __fmul_rn((float)line, wgt);                      // real(.,4) (:290) ; * combined_weight (:842)
    S[(long long)c * ld + row] = v;
  }
  if (e) atomicExch(err, e);
}

int assemble_grav_dense(DenseCM &S, const GridDev &g, int32_t cell0, int32_t ncells, int32_t ndata,
                        const double *d_xd, const double *d_yd, const double *d_zd, const double *d_cw,
                        const double *d_dw, double problem_weight, int *d_err, cudaStream_t st) {
  S.nrows = ndata;
  S.ncols = ncells;
  S.ld = ((int64_t)ndata + 3) / 4 * 4;
  TFX_TRY(S.val.alloc((size_t)S.ld * (size_t)ncells));
  if (S.ld != ndata) TFX_CUDA(cudaMemsetAsync(S.val.p, 0, (size_t)S.ld * ncells * sizeof(float), st));
  const int cols_per_block = 64;
  dim3 grid((ndata + 255) / 256, (ncells + cols_per_block - 1) / cols_per_block);
  if (grid.y > 65535) {
    // chunk the columns over several launches
    const int per = 65535 * cols_per_block;
    for (int c0 = 0; c0 < ncells; c0 += per) {
      const int cn = std::min(per, ncells - c0);
      dim3 gsub((ndata + 255) / 256, (cn + cols_per_block - 1) / cols_per_block);
      grav_dense_kernel<<<gsub, 256, 0, st>>>(S.val.p + (long long)c0 * S.ld, S.ld, ndata, cn, cell0 + c0, g.X1.p,
                                              g.X2.p, g.Y1.p, g.Y2.p, g.Z1.p, g.Z2.p, d_xd, d_yd, d_zd, d_cw, d_dw,
                                              problem_weight, cols_per_block, d_err);
      ctx().launches++;
    }
  } else {
    grav_dense_kernel<<<grid, 256, 0, st>>>(S.val.p, S.ld, ndata, ncells, cell0, g.X1.p, g.X2.p, g.Y1.p, g.Y2.p,
                                            g.Z1.p, g.Z2.p, d_xd, d_yd, d_zd, d_cw, d_dw, problem_weight,
                                            cols_per_block, d_err);
    ctx().launches++;
  }
  TFX_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Full lines (one per station) for the compression pipeline: lines[b*n + p].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grav_lines_kernel(double *__restrict__ lines, int n, int nb,
                                                         const double *__restrict__ X1, const double *__restrict__ X2,
                                                         const double *__restrict__ Y1, const double *__restrict__ Y2,
                                                         const double *__restrict__ Z1, const double *__restrict__ Z2,
                                                         const double *__restrict__ xd, const double *__restrict__ yd,
                                                         const double *__restrict__ zd, int data_type, int *err) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double x1 = X1[p], x2 = X2[p], y1 = Y1[p], y2 = Y2[p], z1 = Z1[p], z2 = Z2[p];
  int e = 0;
  for (int b = blockIdx.y; b < nb; b += gridDim.y) {
    double v;
    if (data_type == 1) v = grav_gz(x1, x2, y1, y2, z1, z2, xd[b], yd[b], zd[b], &e);
    else v = grav_gzz(x1, x2, y1, y2, z1, z2, xd[b], yd[b], zd[b]);
    lines[(long long)b * n + p] = __dmul_rn(g_grav(), v);
  }
  if (e) atomicExch(err, e);
}

// ---- structured grids: one corner term per node -------------------------------------------------
// flag = 1 when some cell box is not the tensor product of the first row / column / pile of boxes, or when
// neighbouring boxes do not share their faces bit for bit.
__global__ void __launch_bounds__(256) grid_structured_kernel(const double *__restrict__ X1, const double *__restrict__ X2,
                                                              const double *__restrict__ Y1, const double *__restrict__ Y2,
                                                              const double *__restrict__ Z1, const double *__restrict__ Z2,
                                                              int nx, int ny, int nz, int *flag) {
  const long long n = (long long)nx * ny * nz;
  int bad = 0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(p % nx), j = (int)((p / nx) % ny), k = (int)(p / ((long long)nx * ny));
    const long long pi = i, pj = (long long)j * nx, pk = (long long)k * nx * ny;
    if (X1[p] != X1[pi] || X2[p] != X2[pi] || Y1[p] != Y1[pj] || Y2[p] != Y2[pj] || Z1[p] != Z1[pk] || Z2[p] != Z2[pk]) bad = 1;
    if (i + 1 < nx && X2[pi] != X1[pi + 1]) bad = 1;
    if (j + 1 < ny && Y2[pj] != Y1[pj + nx]) bad = 1;
    if (k + 1 < nz && Z2[pk] != Z1[pk + (long long)nx * ny]) bad = 1;
  }
  if (bad) atomicExch(flag, 1);
}
__global__ void __launch_bounds__(256) grid_nodes_kernel(const double *__restrict__ A1, const double *__restrict__ A2, int n,
                                                         long long stride, double *__restrict__ nodes) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x)
    nodes[i] = (i < n) ? A1[(long long)i * stride] : A2[(long long)(n - 1) * stride];
}

int g_opt_grav_shared_nodes = 1;

int grid_detect_structured(GridDev &g, int32_t nx, int32_t ny, int32_t nz, cudaStream_t st) {
  if (g.structured >= 0 && g.nx == nx && g.ny == ny && g.nz == nz) return 0;
  g.nx = nx; g.ny = ny; g.nz = nz;
  g.structured = 0;
  if ((long long)nx * ny * nz != g.n || nx < 1 || ny < 1 || nz < 1) return 0;
  DevBuf<int> flag;
  TFX_TRY(flag.alloc(1));
  TFX_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
  grid_structured_kernel<<<ctx().num_sms * 8, 256, 0, st>>>(g.X1.p, g.X2.p, g.Y1.p, g.Y2.p, g.Z1.p, g.Z2.p, nx, ny, nz, flag.p);
  int h = 0;
  TFX_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  ctx().launches++;
  if (h) return 0;
  TFX_TRY(g.xn.alloc((size_t)nx + 1)); TFX_TRY(g.yn.alloc((size_t)ny + 1)); TFX_TRY(g.zn.alloc((size_t)nz + 1));
  grid_nodes_kernel<<<(nx + 256) / 256, 256, 0, st>>>(g.X1.p, g.X2.p, nx, 1, g.xn.p);
  grid_nodes_kernel<<<(ny + 256) / 256, 256, 0, st>>>(g.Y1.p, g.Y2.p, ny, nx, g.yn.p);
  grid_nodes_kernel<<<(nz + 256) / 256, 256, 0, st>>>(g.Z1.p, g.Z2.p, nz, (long long)nx * ny, g.zn.p);
  ctx().launches += 3;
  TFX_CUDA(cudaGetLastError());
  g.structured = 1;
  return 0;
}

// graviprism_z (gravity_field.f90:151-192) on a structured grid. A CTA owns a tile of TX x TY x TZ cells: it evaluates
// the corner term  Z*atan2(X*Y, Z*R) - X*log(R+Y) - Y*log(R+X)  once for each of the (TX+1)(TY+1)(TZ+1) nodes of the
// tile into shared memory (1.3 evaluations per cell instead of 8), then every cell adds its 8 corner terms with the
// reference's signs in the reference's loop order (K outer, L, M inner; dmu = +-1 is an exact sign flip), so the line is
// bit-identical to the per-cell evaluation of grav_gz().
namespace {
constexpr int kNTX = 32, kNTY = 8, kNTZ = 8;
}
__global__ void __launch_bounds__(256, 4) grav_lines_nodes_kernel(double *__restrict__ lines, int nx, int ny, int nz, int nb,
                                                               const double *__restrict__ xn, const double *__restrict__ yn,
                                                               const double *__restrict__ zn, const double *__restrict__ xd,
                                                               const double *__restrict__ yd, const double *__restrict__ zd,
                                                               const double *__restrict__ cw, double *__restrict__ partial,
                                                               int *err) {
  // cw != nullptr: the line is stored already multiplied by the column weight (apply_column_weight,
  // sensitivity_gravmag.F90:1042-1054); partial != nullptr: partial[station][tile] = this tile's share of the sum of the
  // weighted squares (cost_full, :234) -- one pass over the line less in the row pipeline.
  __shared__ double T[kNTZ + 1][kNTY + 1][kNTX + 1];
  __shared__ double red[32];
  const double twopi = 2.0 * TFX_PI;
  const int tiles_x = (nx + kNTX - 1) / kNTX, tiles_y = (ny + kNTY - 1) / kNTY;
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, tz = blockIdx.x / (tiles_x * tiles_y);
  const int i0 = tx * kNTX, j0 = ty * kNTY, k0 = tz * kNTZ;
  const long long n = (long long)nx * ny * nz;
  int e = 0;
  for (int b = blockIdx.y; b < nb; b += gridDim.y) {
    const double px = xd[b], py = yd[b], pz = zd[b];
    for (int idx = threadIdx.x; idx < (kNTX + 1) * (kNTY + 1) * (kNTZ + 1); idx += 256) {
      const int a = idx % (kNTX + 1), bb = (idx / (kNTX + 1)) % (kNTY + 1), c = idx / ((kNTX + 1) * (kNTY + 1));
      const int gi = i0 + a, gj = j0 + bb, gk = k0 + c;
      double term = 0.0;
      if (gi <= nx && gj <= ny && gk <= nz) {
        const double X = px - xn[gi], Y = py - yn[gj], Z = pz - zn[gk];
        const double Rs = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(X, X), __dmul_rn(Y, Y)), __dmul_rn(Z, Z)));
        double arg3 = tfx_atan2(__dmul_rn(X, Y), __dmul_rn(Z, Rs));
        if (arg3 < 0) arg3 = arg3 + twopi;
        double arg4 = Rs + X;
        double arg5 = Rs + Y;
        if (arg4 <= 0.) e = 1;   // "Data coordinate coincides with model grid boundary (YZ)"
        if (arg5 <= 0.) e = 2;   // "... (XZ)"
        arg4 = tfx_log(arg4);
        arg5 = tfx_log(arg5);
        term = __dsub_rn(__dsub_rn(__dmul_rn(Z, arg3), __dmul_rn(X, arg5)), __dmul_rn(Y, arg4));
      }
      T[c][bb][a] = term;
    }
    __syncthreads();
    double ssq = 0.0;
    constexpr int kCellsPerThread = kNTX * kNTY * kNTZ / 256;
    static_assert(kCellsPerThread * 256 == kNTX * kNTY * kNTZ, "tile cells must be a multiple of the block size");
    // the column weights of this thread's cells first (independent loads in flight together: the weight is streamed from
    // HBM once per station and there is little arithmetic in this phase to hide its latency behind)
    double cwv[kCellsPerThread];
    long long pv[kCellsPerThread];
#pragma unroll
    for (int it = 0; it < kCellsPerThread; ++it) {
      const int idx = threadIdx.x + it * 256;
      const int a = idx % kNTX, bb = (idx / kNTX) % kNTY, c = idx / (kNTX * kNTY);
      const int gi = i0 + a, gj = j0 + bb, gk = k0 + c;
      const bool in = gi < nx && gj < ny && gk < nz;
      pv[it] = in ? gi + (long long)gj * nx + (long long)gk * nx * ny : -1;
      cwv[it] = (in && cw) ? cw[pv[it]] : 1.0;
    }
#pragma unroll
    for (int it = 0; it < kCellsPerThread; ++it) {
      const int idx = threadIdx.x + it * 256;
      const int a = idx % kNTX, bb = (idx / kNTX) % kNTY, c = idx / (kNTX * kNTY);
      if (pv[it] >= 0) {
        double gz = 0.0;
#pragma unroll
        for (int K = 0; K < 2; ++K)
#pragma unroll
          for (int L = 0; L < 2; ++L)
#pragma unroll
            for (int M = 0; M < 2; ++M) {
              const double t = T[c + M][bb + L][a + K];
              gz = __dadd_rn(gz, ((K + L + M) & 1) ? t : -t);
            }
        double v = __dmul_rn(g_grav(), gz);
        if (cw) {
          v = __dmul_rn(v, cwv[it]);
          ssq = fma(v, v, ssq);
        }
        lines[(long long)b * n + pv[it]] = v;
      }
    }
    if (partial) {
      ssq = block_sum(ssq, red);   // (contains the barriers that separate this station's tile from the next one's)
      if (threadIdx.x == 0) partial[(long long)b * gridDim.x + blockIdx.x] = ssq;
    }
    __syncthreads();
  }
  if (e) atomicExch(err, e);
}

// Full-tensor gradiometry lines: lines[(b*6 + d)*n + p] (Fortran sensit_line(p, 1, d) of station b).
__global__ void __launch_bounds__(256) grav_full_lines_kernel(double *__restrict__ lines, int n, int nb,
                                                              const double *__restrict__ X1, const double *__restrict__ X2,
                                                              const double *__restrict__ Y1, const double *__restrict__ Y2,
                                                              const double *__restrict__ Z1, const double *__restrict__ Z2,
                                                              const double *__restrict__ xd, const double *__restrict__ yd,
                                                              const double *__restrict__ zd, int *err) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double x1 = X1[p], x2 = X2[p], y1 = Y1[p], y2 = Y2[p], z1 = Z1[p], z2 = Z2[p];
  int e = 0;
  for (int b = blockIdx.y; b < nb; b += gridDim.y) {
    double v[6];
    grav_full(x1, x2, y1, y2, z1, z2, xd[b], yd[b], zd[b], v, &e);
#pragma unroll
    for (int d = 0; d < 6; ++d) lines[((long long)b * 6 + d) * n + p] = __dmul_rn(g_grav(), v[d]);
  }
  if (e) atomicExch(err, e);
}

int grav_full_lines(const GridDev &g, int32_t nb, const double *d_xd, const double *d_yd, const double *d_zd,
                    double *d_lines, int *d_err, cudaStream_t st) {
  dim3 grid((g.n + 255) / 256, std::min(nb, 1024));
  grav_full_lines_kernel<<<grid, 256, 0, st>>>(d_lines, g.n, nb, g.X1.p, g.X2.p, g.Y1.p, g.Y2.p, g.Z1.p, g.Z2.p, d_xd,
                                               d_yd, d_zd, d_err);
  ctx().launches++;
  TFX_CUDA(cudaGetLastError());
  return 0;
}

// Number of per-line partial sums the fused column-weight path of grav_lines() writes; 0: that path does not apply (the
// caller weights the lines and sums their squares itself).
int grav_lines_fused_partials(const GridDev &g, int data_type) {
  if (data_type == 1 && g.structured == 1 && g_opt_grav_shared_nodes)
    return ((g.nx + kNTX - 1) / kNTX) * ((g.ny + kNTY - 1) / kNTY) * ((g.nz + kNTZ - 1) / kNTZ);
  return 0;
}

int grav_lines(const GridDev &g, int32_t nb, const double *d_xd, const double *d_yd, const double *d_zd, int data_type,
               double *d_lines, int *d_err, cudaStream_t st, const double *d_cw, double *d_partial) {
  if (data_type == 1 && g.structured == 1 && g_opt_grav_shared_nodes) {
    const int tiles = ((g.nx + kNTX - 1) / kNTX) * ((g.ny + kNTY - 1) / kNTY) * ((g.nz + kNTZ - 1) / kNTZ);
    dim3 grid(tiles, std::min(nb, 1024));
    grav_lines_nodes_kernel<<<grid, 256, 0, st>>>(d_lines, g.nx, g.ny, g.nz, nb, g.xn.p, g.yn.p, g.zn.p, d_xd, d_yd, d_zd,
                                                  d_cw, d_cw ? d_partial : nullptr, d_err);
    ctx().launches++;
    TFX_CUDA(cudaGetLastError());
    return 0;
  }
  if (d_cw) return fail(-24, "grav_lines: the fused column weight needs the structured-grid kernel");
  dim3 grid((g.n + 255) / 256, std::min(nb, 1024));
  grav_lines_kernel<<<grid, 256, 0, st>>>(d_lines, g.n, nb, g.X1.p, g.X2.p, g.Y1.p, g.Y2.p, g.Z1.p, g.Z2.p, d_xd,
                                          d_yd, d_zd, data_type, d_err);
  ctx().launches++;
  TFX_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Magnetic tensor of one prism (Sharma 1966), magnetic_field.f90:321-457, eps = 0.
// ---------------------------------------------------------------------------------------------
__device__ void sharmbox_dev(double x0, double y0, double z0, double x1, double y1, double z1, double x2, double y2,
                             double z2, double tsx[3], double tsy[3], double tsz[3], int *err) {
  const double rx1 = x1 - x0, rx2 = x2 - x0, ry1 = y1 - y0, ry2 = y2 - y0, rz1 = z1 - z0, rz2 = z2 - z0;
  if (rx1 == 0. || rx2 == 0.) *err = 11;
  if (ry1 == 0. || ry2 == 0.) *err = 12;
  const double rx1sq = __dmul_rn(rx1, rx1), rx2sq = __dmul_rn(rx2, rx2), ry1sq = __dmul_rn(ry1, ry1),
               ry2sq = __dmul_rn(ry2, ry2), rz1sq = __dmul_rn(rz1, rz1), rz2sq = __dmul_rn(rz2, rz2);
  // R = ry^2 + rx^2 first, then rz^2 + R (the reference's association, :361-373)
  double R1 = __dadd_rn(ry2sq, rx2sq), R2 = __dadd_rn(ry2sq, rx1sq), R3 = __dadd_rn(ry1sq, rx2sq),
         R4 = __dadd_rn(ry1sq, rx1sq);
  double a1 = sqrt(__dadd_rn(rz2sq, R2)), a2 = sqrt(__dadd_rn(rz2sq, R1)), a3 = sqrt(__dadd_rn(rz1sq, R1)),
         a4 = sqrt(__dadd_rn(rz1sq, R2)), a5 = sqrt(__dadd_rn(rz2sq, R3)), a6 = sqrt(__dadd_rn(rz2sq, R4)),
         a7 = sqrt(__dadd_rn(rz1sq, R4)), a8 = sqrt(__dadd_rn(rz1sq, R3));
  // ts_xx (:376-383)
  double t = tfx_atan2(__dmul_rn(ry1, rz2), __dmul_rn(rx2, a5));
  t = __dsub_rn(t, tfx_atan2(__dmul_rn(ry2, rz2), __dmul_rn(rx2, a2)));
  t = __dadd_rn(t, tfx_atan2(__dmul_rn(ry2, rz1), __dmul_rn(rx2, a3)));
  t = __dsub_rn(t, tfx_atan2(__dmul_rn(ry1, rz1), __dmul_rn(rx2, a8)));
  t = __dadd_rn(t, tfx_atan2(__dmul_rn(ry2, rz2), __dmul_rn(rx1, a1)));
  t = __dsub_rn(t, tfx_atan2(__dmul_rn(ry1, rz2), __dmul_rn(rx1, a6)));
  t = __dadd_rn(t, tfx_atan2(__dmul_rn(ry1, rz1), __dmul_rn(rx1, a7)));
  t = __dsub_rn(t, tfx_atan2(__dmul_rn(ry2, rz1), __dmul_rn(rx1, a4)));
  tsx[0] = t;
  // ts_yx (:386-389)
  t = tfx_log(__ddiv_rn(rz2 + a2, rz1 + a3));
  t = __dsub_rn(t, tfx_log(__ddiv_rn(rz2 + a1, rz1 + a4)));
  t = __dadd_rn(t, tfx_log(__ddiv_rn(rz2 + a6, rz1 + a7)));
  t = __dsub_rn(t, tfx_log(__ddiv_rn(rz2 + a5, rz1 + a8)));
  tsy[0] = t;
  // ts_yy (:392-399)
  t = tfx_atan2(__dmul_rn(rx1, rz2), __dmul_rn(ry2, a1));
  t = __dsub_rn(t, tfx_atan2(__dmul_rn(rx2, rz2), __dmul_rn(ry2, a2)));
  t = __dadd_rn(t, tfx_atan2(__dmul_rn(rx2, rz1), __dmul_rn(ry2, a3)));
  t = __dsub_rn(t, tfx_atan2(__dmul_rn(rx1, rz1), __dmul_rn(ry2, a4)));
  t = __dadd_rn(t, tfx_atan2(__dmul_rn(rx2, rz2), __dmul_rn(ry1, a5)));
  t = __dsub_rn(t, tfx_atan2(__dmul_rn(rx1, rz2), __dmul_rn(ry1, a6)));
  t = __dadd_rn(t, tfx_atan2(__dmul_rn(rx1, rz1), __dmul_rn(ry1, a7)));
  t = __dsub_rn(t, tfx_atan2(__dmul_rn(rx2, rz1), __dmul_rn(ry1, a8)));
  tsy[1] = t;
  // ts_yz (:404-422)
  R1 = __dadd_rn(ry2sq, rz1sq); R2 = __dadd_rn(ry2sq, rz2sq); R3 = __dadd_rn(ry1sq, rz1sq); R4 = __dadd_rn(ry1sq, rz2sq);
  a1 = sqrt(__dadd_rn(rx1sq, R1)); a2 = sqrt(__dadd_rn(rx2sq, R1)); a3 = sqrt(__dadd_rn(rx1sq, R2));
  a4 = sqrt(__dadd_rn(rx2sq, R2)); a5 = sqrt(__dadd_rn(rx1sq, R3)); a6 = sqrt(__dadd_rn(rx2sq, R3));
  a7 = sqrt(__dadd_rn(rx1sq, R4)); a8 = sqrt(__dadd_rn(rx2sq, R4));
  t = tfx_log(__ddiv_rn(rx1 + a1, rx2 + a2));
  t = __dsub_rn(t, tfx_log(__ddiv_rn(rx1 + a3, rx2 + a4)));
  t = __dadd_rn(t, tfx_log(__ddiv_rn(rx1 + a7, rx2 + a8)));
  t = __dsub_rn(t, tfx_log(__ddiv_rn(rx1 + a5, rx2 + a6)));
  tsy[2] = t;
  // ts_xz (:424-442)
  R1 = __dadd_rn(rx2sq, rz1sq); R2 = __dadd_rn(rx2sq, rz2sq); R3 = __dadd_rn(rx1sq, rz1sq); R4 = __dadd_rn(rx1sq, rz2sq);
  a1 = sqrt(__dadd_rn(ry1sq, R1)); a2 = sqrt(__dadd_rn(ry2sq, R1)); a3 = sqrt(__dadd_rn(ry1sq, R2));
  a4 = sqrt(__dadd_rn(ry2sq, R2)); a5 = sqrt(__dadd_rn(ry1sq, R3)); a6 = sqrt(__dadd_rn(ry2sq, R3));
  a7 = sqrt(__dadd_rn(ry1sq, R4)); a8 = sqrt(__dadd_rn(ry2sq, R4));
  t = tfx_log(__ddiv_rn(ry1 + a1, ry2 + a2));
  t = __dsub_rn(t, tfx_log(__ddiv_rn(ry1 + a3, ry2 + a4)));
  t = __dadd_rn(t, tfx_log(__ddiv_rn(ry1 + a7, ry2 + a8)));
  t = __dsub_rn(t, tfx_log(__ddiv_rn(ry1 + a5, ry2 + a6)));
  tsx[2] = t;
  tsz[2] = -1 * (tsx[0] + tsy[1]);   // Gauss (:446)
  tsz[1] = tsy[2];
  tsx[1] = tsy[0];
  tsz[0] = tsx[2];
}

struct MagPar {
  double magv[3];
  double mult;   // intensity (susceptibility model) or mu0*T2nT (magnetisation model), :286-292
  int nmc, ndc;
};

// magprism, magnetic_field.f90:135-295.
__global__ void __launch_bounds__(128) mag_lines_kernel(double *__restrict__ lines, int n, int nb,
                                                        const double *__restrict__ X1, const double *__restrict__ X2,
                                                        const double *__restrict__ Y1, const double *__restrict__ Y2,
                                                        const double *__restrict__ Z1, const double *__restrict__ Z2,
                                                        const double *__restrict__ xd, const double *__restrict__ yd,
                                                        const double *__restrict__ zd, MagPar mp, int *err) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double gx1 = X1[p], gx2 = X2[p], gy1 = Y1[p], gy2 = Y2[p], gz1 = Z1[p], gz2 = Z2[p];
  int e = 0;
  for (int b = blockIdx.y; b < nb; b += gridDim.y) {
    const double Xd = xd[b], Yd = yd[b], Zd = zd[b];
    double tx[3], ty[3], tz[3];
    if ((gx1 < Xd) && (gx2 > Xd) && (gy1 < Yd) && (gy2 > Yd) && (gz1 < Zd) && (gz2 > Zd)) {
      // Station inside the cell: six sub-prisms around a small void (:139-224).
      double width = (double)0.1f;   // single-precision literal in the reference (:144)
      const double min_clr = fmin(fmin(fmin(fabs(Xd - gx1), fabs(Xd - gx2)), fmin(fabs(Yd - gy1), fabs(Yd - gy2))),
                                  fmin(fabs(Zd - gz1), fabs(Zd - gz2)));
      if (width > min_clr) width = 0.5 * min_clr;
      const double bx1[6] = {gx1, gx1, gx1, Xd + width, Xd - width, Xd - width};
      const double bx2[6] = {gx2, gx2, Xd - width, gx2, Xd + width, Xd + width};
      const double by1[6] = {gy1, gy1, gy1, gy1, gy1, Yd + width};
      const double by2[6] = {gy2, gy2, gy2, gy2, Yd - width, gy2};
      const double bz1[6] = {gz1, Zd + width, Zd - width, Zd - width, Zd - width, Zd - width};
      const double bz2[6] = {Zd - width, gz2, Zd + width, Zd + width, Zd + width, Zd + width};
      for (int q = 0; q < 3; ++q) tx[q] = ty[q] = tz[q] = 0.0;
      for (int j = 0; j < 6; ++j) {
        double ax[3], ay[3], az[3];
        sharmbox_dev(Xd, Yd, Zd, bx1[j], by1[j], bz1[j], bx2[j], by2[j], bz2[j], ax, ay, az, &e);
        for (int q = 0; q < 3; ++q) {
          tx[q] = __dadd_rn(tx[q], ax[q]);
          ty[q] = __dadd_rn(ty[q], ay[q]);
          tz[q] = __dadd_rn(tz[q], az[q]);
        }
      }
    } else {
      sharmbox_dev(Xd, Yd, Zd, gx1, gy1, gz1, gx2, gy2, gz2, tx, ty, tz, &e);
    }
    const double fourpi = 4.0 * TFX_PI;
    double *out = lines + (long long)b * mp.ndc * mp.nmc * n;   // (p, k, d) Fortran order per station
#define OUT(k, d) out[((long long)(d)*mp.nmc + (k)) * n + p]
#define FIN(x) __ddiv_rn(__dmul_rn(mp.mult, (x)), fourpi)
#define DOT3(a) __dadd_rn(__dadd_rn(__dmul_rn(a[0], mp.magv[0]), __dmul_rn(a[1], mp.magv[1])), __dmul_rn(a[2], mp.magv[2]))
    if (mp.nmc == 1) {
      const double mx = DOT3(tx), my = DOT3(ty), mz = DOT3(tz);
      if (mp.ndc == 1) {
        OUT(0, 0) = FIN(__dadd_rn(__dadd_rn(__dmul_rn(mx, mp.magv[0]), __dmul_rn(my, mp.magv[1])), __dmul_rn(mz, mp.magv[2])));
      } else {
        OUT(0, 0) = FIN(mx); OUT(0, 1) = FIN(my); OUT(0, 2) = FIN(mz);
      }
    } else {
      for (int k = 0; k < 3; ++k) {
        if (mp.ndc == 1) {
          OUT(k, 0) = FIN(__dadd_rn(__dadd_rn(__dmul_rn(tx[k], mp.magv[0]), __dmul_rn(ty[k], mp.magv[1])), __dmul_rn(tz[k], mp.magv[2])));
        } else {
          OUT(k, 0) = FIN(tx[k]); OUT(k, 1) = FIN(ty[k]); OUT(k, 2) = FIN(tz[k]);
        }
      }
    }
#undef OUT
#undef FIN
#undef DOT3
  }
  if (e) atomicExch(err, e);
}

// The cell that contains the station: six sub-prisms around it (magnetic_field.f90:139-224). Out of line: one cell per
// station at most takes this path, and inlined it sets the register count of the whole kernel.
__device__ __noinline__ void mag_incell_dev(double Xd, double Yd, double Zd, double gx1, double gy1, double gz1, double gx2,
                                            double gy2, double gz2, double *tx, double *ty, double *tz, int *e) {
  double width = (double)0.1f;
  const double min_clr = fmin(fmin(fmin(fabs(Xd - gx1), fabs(Xd - gx2)), fmin(fabs(Yd - gy1), fabs(Yd - gy2))),
                              fmin(fabs(Zd - gz1), fabs(Zd - gz2)));
  if (width > min_clr) width = 0.5 * min_clr;
  const double bx1[6] = {gx1, gx1, gx1, Xd + width, Xd - width, Xd - width};
  const double bx2[6] = {gx2, gx2, Xd - width, gx2, Xd + width, Xd + width};
  const double by1[6] = {gy1, gy1, gy1, gy1, gy1, Yd + width};
  const double by2[6] = {gy2, gy2, gy2, gy2, Yd - width, gy2};
  const double bz1[6] = {gz1, Zd + width, Zd - width, Zd - width, Zd - width, Zd - width};
  const double bz2[6] = {Zd - width, gz2, Zd + width, Zd + width, Zd + width, Zd + width};
  for (int q = 0; q < 3; ++q) tx[q] = ty[q] = tz[q] = 0.0;
  for (int j = 0; j < 6; ++j) {
    double ax[3], ay[3], az[3];
    sharmbox_dev(Xd, Yd, Zd, bx1[j], by1[j], bz1[j], bx2[j], by2[j], bz2[j], ax, ay, az, e);
    for (int q = 0; q < 3; ++q) {
      tx[q] = __dadd_rn(tx[q], ax[q]);
      ty[q] = __dadd_rn(ty[q], ay[q]);
      tz[q] = __dadd_rn(tz[q], az[q]);
    }
  }
}

// ---- structured grids: shared corner / edge terms -------------------------------------------------
// sharmbox (magnetic_field.f90:321-457) is a signed sum of terms that belong either to ONE corner of the prism
//   A = atan2(ry*rz, rx*a)  (ts_xx),   B = atan2(rx*rz, ry*a)  (ts_yy),        a = sqrt(rz^2 + (ry^2 + rx^2))
// or to ONE edge of it (the log of a ratio of the two end corners)
//   Ez = log((rz2 + a_hi) / (rz1 + a_lo))  (ts_yx, edges along z, the association of a above),
//   Ex = log((rx1 + a') / (rx2 + a'))      (ts_yz, edges along x, a' = sqrt(rx^2 + (ry^2 + rz^2))),
//   Ey = log((ry1 + a'') / (ry2 + a''))    (ts_xz, edges along y, a'' = sqrt(ry^2 + (rx^2 + rz^2))).
// On a structured grid a corner is shared by 8 cells and an edge by 4: a CTA evaluates the terms of a
// 32 x 8 x 8-cell tile once into shared memory (2 atan2 per node, 1 log per edge instead of 16 atan2 + 12 log per
// cell) and every cell adds them with the reference's signs in the reference's order -- bit-identical to
// sharmbox_dev(). The cell that contains the station takes the six-sub-prism branch (:139-224) as before.
namespace {
constexpr int kMTX = 32, kMTY = 8, kMTZ = 8;
constexpr int kMNodes = (kMTX + 1) * (kMTY + 1) * (kMTZ + 1);
constexpr int kMEz = (kMTX + 1) * (kMTY + 1) * kMTZ, kMEx = kMTX * (kMTY + 1) * (kMTZ + 1), kMEy = (kMTX + 1) * kMTY * (kMTZ + 1);
constexpr int kMagThreads = 512;   // 2 CTAs of 512 threads per SM (shared memory: 2 x 101 KB), 64 registers
constexpr size_t kMagSmem = (size_t)(2 * kMNodes + kMEz + kMEx + kMEy) * sizeof(double);
}

__global__ void __launch_bounds__(kMagThreads, 2) mag_lines_nodes_kernel(double *__restrict__ lines, int nx, int ny, int nz, int nb,
                                                              const double *__restrict__ xn, const double *__restrict__ yn,
                                                              const double *__restrict__ zn, const double *__restrict__ xd,
                                                              const double *__restrict__ yd, const double *__restrict__ zd,
                                                              MagPar mp, int *err) {
  extern __shared__ double msm[];
  double *sA = msm, *sB = sA + kMNodes, *sEz = sB + kMNodes, *sEx = sEz + kMEz, *sEy = sEx + kMEx;
#define NA(a, b, c) sA[((c) * (kMTY + 1) + (b)) * (kMTX + 1) + (a)]
#define NB(a, b, c) sB[((c) * (kMTY + 1) + (b)) * (kMTX + 1) + (a)]
#define EZ(a, b, c) sEz[((c) * (kMTY + 1) + (b)) * (kMTX + 1) + (a)]     /* edge (a, b) from node level c to c+1 */
#define EX(a, b, c) sEx[((c) * (kMTY + 1) + (b)) * kMTX + (a)]           /* edge (b, c) from node a to a+1     */
#define EY(a, b, c) sEy[((c) * kMTY + (b)) * (kMTX + 1) + (a)]           /* edge (a, c) from node b to b+1     */
  const int tiles_x = (nx + kMTX - 1) / kMTX, tiles_y = (ny + kMTY - 1) / kMTY;
  const int tx_ = blockIdx.x % tiles_x, ty_ = (blockIdx.x / tiles_x) % tiles_y, tz_ = blockIdx.x / (tiles_x * tiles_y);
  const int i0 = tx_ * kMTX, j0 = ty_ * kMTY, k0 = tz_ * kMTZ;
  const long long n = (long long)nx * ny * nz;
  int e = 0;
  for (int b_ = blockIdx.y; b_ < nb; b_ += gridDim.y) {
    const double Xd = xd[b_], Yd = yd[b_], Zd = zd[b_];
    // ---- corner terms
    for (int idx = threadIdx.x; idx < kMNodes; idx += kMagThreads) {
      const int a = idx % (kMTX + 1), b = (idx / (kMTX + 1)) % (kMTY + 1), c = idx / ((kMTX + 1) * (kMTY + 1));
      const int gi = min(i0 + a, nx), gj = min(j0 + b, ny), gk = min(k0 + c, nz);
      const double rx = xn[gi] - Xd, ry = yn[gj] - Yd, rz = zn[gk] - Zd;
      if (i0 + a <= nx && j0 + b <= ny && k0 + c <= nz) {
        if (rx == 0.) e = 11;
        if (ry == 0.) e = 12;
      }
      const double as1 = sqrt(__dadd_rn(__dmul_rn(rz, rz), __dadd_rn(__dmul_rn(ry, ry), __dmul_rn(rx, rx))));
      sA[idx] = tfx_atan2(__dmul_rn(ry, rz), __dmul_rn(rx, as1));
      sB[idx] = tfx_atan2(__dmul_rn(rx, rz), __dmul_rn(ry, as1));
    }
    // ---- edge terms
    for (int idx = threadIdx.x; idx < kMEz; idx += kMagThreads) {
      const int a = idx % (kMTX + 1), b = (idx / (kMTX + 1)) % (kMTY + 1), c = idx / ((kMTX + 1) * (kMTY + 1));
      const int gi = min(i0 + a, nx), gj = min(j0 + b, ny), g1 = min(k0 + c, nz), g2 = min(k0 + c + 1, nz);
      const double rx = xn[gi] - Xd, ry = yn[gj] - Yd, rz1 = zn[g1] - Zd, rz2 = zn[g2] - Zd;
      const double R = __dadd_rn(__dmul_rn(ry, ry), __dmul_rn(rx, rx));
      const double a_lo = sqrt(__dadd_rn(__dmul_rn(rz1, rz1), R)), a_hi = sqrt(__dadd_rn(__dmul_rn(rz2, rz2), R));
      sEz[idx] = tfx_log(__ddiv_rn(rz2 + a_hi, rz1 + a_lo));
    }
    for (int idx = threadIdx.x; idx < kMEx; idx += kMagThreads) {
      const int a = idx % kMTX, b = (idx / kMTX) % (kMTY + 1), c = idx / (kMTX * (kMTY + 1));
      const int g1 = min(i0 + a, nx), g2 = min(i0 + a + 1, nx), gj = min(j0 + b, ny), gk = min(k0 + c, nz);
      const double rx1 = xn[g1] - Xd, rx2 = xn[g2] - Xd, ry = yn[gj] - Yd, rz = zn[gk] - Zd;
      const double R = __dadd_rn(__dmul_rn(ry, ry), __dmul_rn(rz, rz));
      const double a1 = sqrt(__dadd_rn(__dmul_rn(rx1, rx1), R)), a2 = sqrt(__dadd_rn(__dmul_rn(rx2, rx2), R));
      sEx[idx] = tfx_log(__ddiv_rn(rx1 + a1, rx2 + a2));
    }
    for (int idx = threadIdx.x; idx < kMEy; idx += kMagThreads) {
      const int a = idx % (kMTX + 1), b = (idx / (kMTX + 1)) % kMTY, c = idx / ((kMTX + 1) * kMTY);
      const int gi = min(i0 + a, nx), g1 = min(j0 + b, ny), g2 = min(j0 + b + 1, ny), gk = min(k0 + c, nz);
      const double rx = xn[gi] - Xd, ry1 = yn[g1] - Yd, ry2 = yn[g2] - Yd, rz = zn[gk] - Zd;
      const double R = __dadd_rn(__dmul_rn(rx, rx), __dmul_rn(rz, rz));
      const double a1 = sqrt(__dadd_rn(__dmul_rn(ry1, ry1), R)), a2 = sqrt(__dadd_rn(__dmul_rn(ry2, ry2), R));
      sEy[idx] = tfx_log(__ddiv_rn(ry1 + a1, ry2 + a2));
    }
    __syncthreads();
    // ---- cells
    for (int idx = threadIdx.x; idx < kMTX * kMTY * kMTZ; idx += kMagThreads) {
      const int a = idx % kMTX, b = (idx / kMTX) % kMTY, c = idx / (kMTX * kMTY);
      const int gi = i0 + a, gj = j0 + b, gk = k0 + c;
      if (gi >= nx || gj >= ny || gk >= nz) continue;
      const double gx1 = xn[gi], gx2 = xn[gi + 1], gy1 = yn[gj], gy2 = yn[gj + 1], gz1 = zn[gk], gz2 = zn[gk + 1];
      double tx[3], ty[3], tz[3];
      if ((gx1 < Xd) && (gx2 > Xd) && (gy1 < Yd) && (gy2 > Yd) && (gz1 < Zd) && (gz2 > Zd)) {
        mag_incell_dev(Xd, Yd, Zd, gx1, gy1, gz1, gx2, gy2, gz2, tx, ty, tz, &e);
      } else {
        // corner (x, y, z) in {1, 2}^3  ->  node (a + x - 1, b + y - 1, c + z - 1); orders and signs of :376-442
        double t = NA(a + 1, b, c + 1);
        t = __dsub_rn(t, NA(a + 1, b + 1, c + 1));
        t = __dadd_rn(t, NA(a + 1, b + 1, c));
        t = __dsub_rn(t, NA(a + 1, b, c));
        t = __dadd_rn(t, NA(a, b + 1, c + 1));
        t = __dsub_rn(t, NA(a, b, c + 1));
        t = __dadd_rn(t, NA(a, b, c));
        t = __dsub_rn(t, NA(a, b + 1, c));
        tx[0] = t;
        t = EZ(a + 1, b + 1, c);
        t = __dsub_rn(t, EZ(a, b + 1, c));
        t = __dadd_rn(t, EZ(a, b, c));
        t = __dsub_rn(t, EZ(a + 1, b, c));
        ty[0] = t;
        t = NB(a, b + 1, c + 1);
        t = __dsub_rn(t, NB(a + 1, b + 1, c + 1));
        t = __dadd_rn(t, NB(a + 1, b + 1, c));
        t = __dsub_rn(t, NB(a, b + 1, c));
        t = __dadd_rn(t, NB(a + 1, b, c + 1));
        t = __dsub_rn(t, NB(a, b, c + 1));
        t = __dadd_rn(t, NB(a, b, c));
        t = __dsub_rn(t, NB(a + 1, b, c));
        ty[1] = t;
        t = EX(a, b + 1, c);
        t = __dsub_rn(t, EX(a, b + 1, c + 1));
        t = __dadd_rn(t, EX(a, b, c + 1));
        t = __dsub_rn(t, EX(a, b, c));
        ty[2] = t;
        t = EY(a + 1, b, c);
        t = __dsub_rn(t, EY(a + 1, b, c + 1));
        t = __dadd_rn(t, EY(a, b, c + 1));
        t = __dsub_rn(t, EY(a, b, c));
        tx[2] = t;
        tz[2] = -1 * (tx[0] + ty[1]);   // Gauss (:446)
        tz[1] = ty[2];
        tx[1] = ty[0];
        tz[0] = tx[2];
      }
      const double fourpi = 4.0 * TFX_PI;
      const long long p = gi + (long long)gj * nx + (long long)gk * nx * ny;
      double *out = lines + (long long)b_ * mp.ndc * mp.nmc * n;
#define OUT(k, d) out[((long long)(d)*mp.nmc + (k)) * n + p]
#define FIN(x) __ddiv_rn(__dmul_rn(mp.mult, (x)), fourpi)
#define DOT3(v) __dadd_rn(__dadd_rn(__dmul_rn(v[0], mp.magv[0]), __dmul_rn(v[1], mp.magv[1])), __dmul_rn(v[2], mp.magv[2]))
      if (mp.nmc == 1) {
        const double mx = DOT3(tx), my = DOT3(ty), mz = DOT3(tz);
        if (mp.ndc == 1) {
          OUT(0, 0) = FIN(__dadd_rn(__dadd_rn(__dmul_rn(mx, mp.magv[0]), __dmul_rn(my, mp.magv[1])), __dmul_rn(mz, mp.magv[2])));
        } else {
          OUT(0, 0) = FIN(mx); OUT(0, 1) = FIN(my); OUT(0, 2) = FIN(mz);
        }
      } else {
        for (int k = 0; k < 3; ++k) {
          if (mp.ndc == 1) {
            OUT(k, 0) = FIN(__dadd_rn(__dadd_rn(__dmul_rn(tx[k], mp.magv[0]), __dmul_rn(ty[k], mp.magv[1])), __dmul_rn(tz[k], mp.magv[2])));
          } else {
            OUT(k, 0) = FIN(tx[k]); OUT(k, 1) = FIN(ty[k]); OUT(k, 2) = FIN(tz[k]);
          }
        }
      }
#undef OUT
#undef FIN
#undef DOT3
    }
    __syncthreads();
  }
#undef NA
#undef NB
#undef EZ
#undef EX
#undef EY
  if (e) atomicExch(err, e);
}

int g_opt_mag_shared_nodes = 1;

int mag_lines(const GridDev &g, int32_t nb, const double *d_xd, const double *d_yd, const double *d_zd, int nmc, int ndc,
              double mi, double md, double theta, double intensity, double *d_lines, int *d_err, cudaStream_t st) {
  if (!((nmc == 1 || nmc == 3) && (ndc == 1 || ndc == 3)))
    return fail(-40, "Wrong number of model/data components in magnetic_field_magprism!");
  MagPar mp;
  // dircos, magnetic_field.f90:91-110 (host libm; three scalars).
  const double d2rad = TFX_PI / 180.0;
  const double decl2 = fmod(450.0 - md, 360.0);
  const double xincl = mi * d2rad, xdecl = decl2 * d2rad, xazim = theta * d2rad;
  mp.magv[0] = cos(xincl) * cos(xdecl - xazim);
  mp.magv[1] = cos(xincl) * sin(xdecl - xazim);
  mp.magv[2] = sin(xincl);
  const double mu0 = 4.0 * TFX_PI * 1.e-7, T2nT = 1.e+9;
  mp.mult = (nmc == 1) ? intensity : (mu0 * T2nT);
  mp.nmc = nmc;
  mp.ndc = ndc;
  if (g.structured == 1 && g_opt_mag_shared_nodes) {
    static bool attr = false;
    if (!attr) {
      TFX_CUDA(cudaFuncSetAttribute(mag_lines_nodes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMagSmem));
      attr = true;
    }
    const int tiles = ((g.nx + kMTX - 1) / kMTX) * ((g.ny + kMTY - 1) / kMTY) * ((g.nz + kMTZ - 1) / kMTZ);
    dim3 grid(tiles, std::min(nb, 1024));
    mag_lines_nodes_kernel<<<grid, kMagThreads, kMagSmem, st>>>(d_lines, g.nx, g.ny, g.nz, nb, g.xn.p, g.yn.p, g.zn.p, d_xd, d_yd, d_zd,
                                                        mp, d_err);
    ctx().launches++;
    TFX_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid((g.n + 127) / 128, std::min(nb, 1024));
  mag_lines_kernel<<<grid, 128, 0, st>>>(d_lines, g.n, nb, g.X1.p, g.X2.p, g.Y1.p, g.Y2.p, g.Z1.p, g.Z2.p, d_xd, d_yd,
                                         d_zd, mp, d_err);
  ctx().launches++;
  TFX_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) debug_math_kernel(long long n, const double *__restrict__ y,
                                                         const double *__restrict__ x, double *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out[i] = tfx_log(x[i]);
    out[n + i] = log(x[i]);
    out[2 * n + i] = tfx_atan2(y[i], x[i]);
    out[3 * n + i] = atan2(y[i], x[i]);
  }
}
int debug_math(long long n, const double *d_y, const double *d_x, double *d_out, cudaStream_t st) {
  debug_math_kernel<<<ctx().num_sms * 4, 256, 0, st>>>(n, d_y, d_x, d_out);
  ctx().launches++;
  TFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tfx
