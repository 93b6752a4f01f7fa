// data.cu -- the callers either side of the products: apply_wavelet_transform on a distributed model vector
// (src/inversion/wavelet_utils.F90:37-72) and t_model%calculate_data (src/inversion/model.F90:220-307), kept on
// the device so the model never leaves HBM between solves.
#include "../../include/tfx.h"

#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

namespace tfx {

namespace {
DevBuf<double> &full_scratch() {
  static DevBuf<double> b;
  return b;
}

// model_scaled(i, k) = val(i, k) / column_weight(i), 0 where the weight is 0 (model.F90:243-251)
__global__ void __launch_bounds__(256) k_scale_model(const double *__restrict__ val, const double *__restrict__ cw,
                                                      int64_t nelements, int64_t total, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const double w = cw[i % nelements];
    out[i] = (w != 0.0) ? __ddiv_rn(val[i], w) : 0.0;
  }
}
// data_calc = data_calc / problem_weight / data_weight (model.F90:296-304)
__global__ void __launch_bounds__(256) k_unweight_data(double *__restrict__ d, const double *__restrict__ dw, int64_t n,
                                                        double problem_weight) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = __ddiv_rn(__ddiv_rn(d[i], problem_weight), dw[i]);
}
// rescale_model (model.F90:312-324): model(i, k) *= weight(i); model_update (:194-200): val += delta
__global__ void __launch_bounds__(256) k_rescale_model(double *__restrict__ m, const double *__restrict__ w, int64_t nelements,
                                                       int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    m[i] = __dmul_rn(m[i], w[i % nelements]);
}
__global__ void __launch_bounds__(256) k_model_update(double *__restrict__ v, const double *__restrict__ d, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = __dadd_rn(v[i], d[i]);
}
}  // namespace

// Every rank drops its slab at its place in a full volume, the slabs are all-gathered over NVLink (each GPU receives the
// other ranks' cells once: half the bytes of an all-reduce into a zeroed volume, no additions), every GPU transforms the
// identical volume and keeps its own cells: get_full_array + scatter_full_array (parallel_tools.f90:147,250) without the
// rank-0 serial section, bit-identical to the serial transform.
int wavelet_slab_device_off(double *d_slab, const std::vector<int64_t> &offsets, int nx, int ny, int nz, int wavelet_type,
                            bool forward, cudaStream_t st) {
  const int64_t N = (int64_t)nx * ny * nz;
  const int nranks = comm_nranks(), rank = comm_rank();
  if (nranks <= 1) {
    if (offsets.size() < 2 || offsets[1] - offsets[0] != N)
      return fail(-24, "apply_wavelet_transform: nelements must equal nx*ny*nz on a single rank");
    return wavelet3d_device(d_slab, nx, ny, nz, wavelet_type, forward, st);
  }
  if ((int)offsets.size() != nranks + 1 || offsets[(size_t)nranks] != N)
    return fail(-24, "apply_wavelet_transform: the ranks' nelements must add up to nx*ny*nz");
  const int64_t nsmaller = offsets[(size_t)rank], nelements = offsets[(size_t)rank + 1] - nsmaller;
  DevBuf<double> &F = full_scratch();
  TFX_TRY(F.alloc((size_t)N));
  TFX_CUDA(cudaMemcpyAsync(F.p + nsmaller, d_slab, (size_t)nelements * 8, cudaMemcpyDeviceToDevice, st));
  TFX_TRY(comm_allgatherv_f64(F.p, offsets.data(), st));
  TFX_TRY(wavelet3d_device(F.p, nx, ny, nz, wavelet_type, forward, st));
  TFX_CUDA(cudaMemcpyAsync(d_slab, F.p + nsmaller, (size_t)nelements * 8, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int wavelet_slab_device(double *d_slab, int64_t nelements, int64_t nsmaller, int nx, int ny, int nz, int wavelet_type,
                        bool forward, cudaStream_t st) {
  (void)nsmaller;
  std::vector<int64_t> offsets;
  TFX_TRY(comm_slab_offsets(nelements, offsets));
  return wavelet_slab_device_off(d_slab, offsets, nx, ny, nz, wavelet_type, forward, st);
}

}  // namespace tfx

using namespace tfx;

extern "C" int tfx_apply_wavelet_transform(int32_t nelements, int32_t nx, int32_t ny, int32_t nz, int32_t ncomponents,
                                           double *v, int32_t fwd, int32_t compression_type, int32_t nproblems,
                                           const int32_t *solve_problem, int32_t myrank, int32_t nbproc) {
  (void)myrank;
  TFX_TRY(ensure_init());
  if (nbproc > 1 && comm_nranks() != nbproc)
    return fail(-24, "apply_wavelet_transform: nbproc does not match the communicator (tfx_comm_init)");
  int64_t nsmaller = 0, total = nelements;
  if (nbproc > 1) TFX_TRY(comm_slab_offset(nelements, &nsmaller, &total));
  if (total != (int64_t)nx * ny * nz)
    return fail(-24, "apply_wavelet_transform: the ranks' nelements must add up to nx*ny*nz");
  VecIO io;
  TFX_TRY(io.bind(v, (size_t)nelements * ncomponents * nproblems, true));
  for (int i = 0; i < nproblems; ++i) {
    if (!solve_problem[i]) continue;
    for (int k = 0; k < ncomponents; ++k)
      TFX_TRY(wavelet_slab_device(io.dev + ((size_t)i * ncomponents + k) * nelements, nelements, nsmaller, nx, ny, nz,
                                  compression_type, fwd != 0, ctx().stream));
  }
  TFX_TRY(io.copy_back());
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

// t_model%calculate_data (model.F90:220-307): d = (S W(m / cw)) / problem_weight / data_weight for one problem of
// the (joint) matrix. model_val(nelements, ncomponents), column_weight(nelements), data_weight / data_calc
// (ndata_components, ndata); all four may be host or device pointers.
extern "C" int tfx_calculate_data(tfx_matrix *matrix_sensit, int32_t nelements, int32_t ncomponents, const double *model_val,
                                  int32_t ndata, int32_t ndata_components, double problem_weight,
                                  const double *column_weight, const double *data_weight, double *data_calc,
                                  int32_t compression_type, int32_t nx, int32_t ny, int32_t nz, int32_t line_start,
                                  int32_t param_shift, int32_t myrank, int32_t nbproc) {
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  if (!matrix_sensit) return fail(-82, "calculate_data: null matrix");
  if (problem_weight == 0.0) return fail(-96, "Zero problem weight in model_calculate_data!");
  if (nbproc > 1 && comm_nranks() != nbproc)
    return fail(-24, "calculate_data: nbproc does not match the communicator (tfx_comm_init)");
  const int64_t nm = (int64_t)nelements * ncomponents, nd = (int64_t)ndata * ndata_components;
  VecIO vm, vcw, vdw, vd;
  TFX_TRY(vm.bind(const_cast<double *>(model_val), (size_t)nm, true));
  TFX_TRY(vcw.bind(const_cast<double *>(column_weight), (size_t)nelements, true));
  TFX_TRY(vdw.bind(const_cast<double *>(data_weight), (size_t)nd, true));
  TFX_TRY(vd.bind(data_calc, (size_t)nd, false));
  DevBuf<double> scaled;
  TFX_TRY(scaled.alloc((size_t)nm));
  const int grid = (int)std::min<int64_t>((nm + 255) / 256, (int64_t)c.num_sms * 8);
  k_scale_model<<<grid, 256, 0, st>>>(vm.dev, vcw.dev, nelements, nm, scaled.p);
  c.launches++;
  if (compression_type > 0) {
    int64_t nsmaller = 0, total = nelements;
    if (nbproc > 1) TFX_TRY(comm_slab_offset(nelements, &nsmaller, &total));
    if (total != (int64_t)nx * ny * nz) return fail(-24, "calculate_data: the ranks' nelements must add up to nx*ny*nz");
    for (int k = 0; k < ncomponents; ++k)
      TFX_TRY(wavelet_slab_device(scaled.p + (size_t)k * nelements, nelements, nsmaller, nx, ny, nz, compression_type, true, st));
  }
  TFX_TRY(tfx_sparse_matrix_part_mult_vector(matrix_sensit, (int32_t)nm, scaled.p, (int32_t)nd, vd.dev, line_start,
                                             param_shift, myrank));
  if (nbproc > 1) TFX_TRY(comm_allreduce_sum(vd.dev, (size_t)nd, st));               // MPI_Allreduce, model.F90:293
  k_unweight_data<<<(int)std::min<int64_t>((nd + 255) / 256, (int64_t)c.num_sms * 8), 256, 0, st>>>(vd.dev, vdw.dev, nd,
                                                                                                    problem_weight);
  c.launches++;
  TFX_TRY(vd.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// rescale_model (src/inversion/model.F90:312-324), applied to delta_model after the solve
// (joint_inverse_problem.F90:569-571): model(nelements, ncomponents) *= weight(nelements).
extern "C" int tfx_rescale_model(int32_t nelements, int32_t ncomponents, double *model, const double *weight) {
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const int64_t total = (int64_t)nelements * ncomponents;
  VecIO vm, vw;
  TFX_TRY(vm.bind(model, (size_t)total, true));
  TFX_TRY(vw.bind(const_cast<double *>(weight), (size_t)nelements, true));
  k_rescale_model<<<(int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)c.num_sms * 8)), 256, 0, st>>>(
      vm.dev, vw.dev, nelements, total);
  c.launches++;
  TFX_TRY(vm.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// t_model%update (src/inversion/model.F90:194-200): val(nelements, ncomponents) += delta_model.
extern "C" int tfx_model_update(int32_t nelements, int32_t ncomponents, double *val, const double *delta_model) {
  TFX_TRY(ensure_init());
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const int64_t total = (int64_t)nelements * ncomponents;
  VecIO vv, vd;
  TFX_TRY(vv.bind(val, (size_t)total, true));
  TFX_TRY(vd.bind(const_cast<double *>(delta_model), (size_t)total, true));
  k_model_update<<<(int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)c.num_sms * 8)), 256, 0, st>>>(
      vv.dev, vd.dev, total);
  c.launches++;
  TFX_TRY(vv.copy_back());
  TFX_CUDA(cudaStreamSynchronize(st));
  return 0;
}
