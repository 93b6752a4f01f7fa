// sensit_dist.cu -- multi-GPU sensitivity assembly: rows sharded by data, re-partitioned to column slabs.
//
// The reference computes the kernel "parallelized by data" (every rank evaluates its share of the
// stations and writes one stream file, sensitivity_gravmag.F90:179-318), derives an nnz-balanced
// column partitioning from the per-cell entry counts (get_load_balancing_nelements, :470-524;
// calculate_new_partitioning, :573-642) and re-reads the files "parallelized by model": rank 0 reads
// every row and MPI_Scatterv's its pieces to the column owners (read_sensitivity_kernel, :648-883).
//
// Here the three stages keep the same meaning but the files become HBM-resident row shards and the
// per-row Scatterv becomes ONE all-to-all over NVLink (grouped ncclSend/ncclRecv):
//   1. tfx_sensit_assemble_rows : row pipeline (sensit.cu) for this rank's stations -> (row, col, value)
//      entries on the device; sensit_nnz / nnz_total / compression error reduced over ranks (:322-353).
//   2. tfx_get_load_balancing_nelements : the reference's partitioner (host, integer work).
//   3. tfx_sensit_repartition   : per (row segment, destination) bounds by binary search (columns are
//      ascending inside a segment, :258-272), pack per destination with the destination's local column
//      index (:834), exchange, and build the column-slab matrix (CSR, CSR of the transpose, T16).
// Pieces arrive in source-rank order == global row order, so the received entries are already sorted by
// (row, column) and the row order inside every column equals the reference's (add_row order, :846).
#include "../../include/tfx.h"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "matrix.h"

namespace tfx {

// parallel_tools.f90:46-63 / :68-86
static int32_t nelements_at_cpu_even(int32_t total, int32_t rank, int32_t nbproc) {
  int32_t n = total / nbproc;
  if (rank + 1 <= total - n * nbproc) n += 1;
  return n;
}
static int32_t nsmaller_even(int32_t total, int32_t rank, int32_t nbproc) {
  int32_t s = 0;
  for (int32_t r = 0; r < rank; ++r) s += nelements_at_cpu_even(total, r, nbproc);
  return s;
}

// bound[s*(P+1) + r] = first entry of segment s whose cell index is >= cum[r] (r = 0..P).
__global__ void __launch_bounds__(256) k_piece_bounds(const int32_t *__restrict__ idx, const int64_t *__restrict__ seg_beg,
                                                       int64_t nseg, int32_t nmc, int32_t N, int32_t P,
                                                       const int32_t *__restrict__ cum, int64_t *__restrict__ bound) {
  const int64_t total = nseg * (P + 1);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = t / (P + 1);
    const int32_t r = (int32_t)(t % (P + 1));
    const int32_t k = (int32_t)(s % nmc);
    const int64_t target = (int64_t)k * N + cum[r];
    int64_t lo = seg_beg[s], hi = seg_beg[s + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)idx[mid] < target) lo = mid + 1;
      else hi = mid;
    }
    bound[t] = lo;
  }
}
// cnt[r*nseg + s] = entries of segment s owned by destination r.
__global__ void __launch_bounds__(256) k_piece_counts(const int64_t *__restrict__ bound, int64_t nseg, int32_t P,
                                                       int64_t *__restrict__ cnt) {
  const int64_t total = nseg * P;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / nseg, s = t % nseg;
    cnt[t] = bound[s * (P + 1) + r + 1] - bound[s * (P + 1) + r];
  }
}
// One CTA per segment (grid-stride): copies each destination's piece to its place in the send buffers and
// rewrites the column as the destination's LOCAL column (sensitivity_gravmag.F90:834):
//   local = (p - nsmaller_r) + k*nelements_r + param_shift_r,  param_shift_r = (slot-1)*nelements_r*nmc (:685-686)
__global__ void __launch_bounds__(256) k_pack_pieces(const int32_t *__restrict__ idx, const float *__restrict__ val,
                                                      const int32_t *__restrict__ rowid, const int64_t *__restrict__ bound,
                                                      const int64_t *__restrict__ off, int64_t nseg, int32_t nmc,
                                                      int32_t N, int32_t P, const int32_t *__restrict__ cum, int32_t slot,
                                                      int32_t only_dest, int64_t out_base, int32_t *__restrict__ s_idx,
                                                      float *__restrict__ s_val, int32_t *__restrict__ s_row) {
  for (int64_t s = blockIdx.x; s < nseg; s += gridDim.x) {
    const int32_t k = (int32_t)(s % nmc);
    for (int32_t r = 0; r < P; ++r) {
      if (only_dest >= 0 && r != only_dest) continue;
      const int64_t b = bound[s * (P + 1) + r], e = bound[s * (P + 1) + r + 1];
      const int64_t o = off[(int64_t)r * nseg + s] - out_base;
      const int32_t nel_r = cum[r + 1] - cum[r];
      const int32_t delta = -k * N - cum[r] + k * nel_r + slot * nel_r * nmc;
      for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        s_idx[o + (i - b)] = idx[i] + delta;
        s_val[o + (i - b)] = val[i];
        s_row[o + (i - b)] = rowid[i];
      }
    }
  }
}

// val[i] *= wgt[rowid[i]] in real(4): sensit_compressed(j) * combined_weight (sensitivity_gravmag.F90:837-843)
__global__ void __launch_bounds__(256) k_apply_row_weight(float *__restrict__ val, const int32_t *__restrict__ rowid,
                                                           int64_t n, const float *__restrict__ wgt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    val[i] = __fmul_rn(val[i], wgt[rowid[i]]);
}

template <typename T>
static int up(DevBuf<T> &d, const T *h, size_t n) {
  TFX_TRY(d.alloc(n));
  if (n) TFX_CUDA(cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx().stream));
  return 0;
}

}  // namespace tfx

using namespace tfx;

extern "C" int tfx_get_load_balancing_nelements(int32_t nelements_total, const int32_t *sensit_nnz, int32_t nbproc,
                                                int64_t *nnz_at_cpu_new, int32_t *nelements_at_cpu_new) {
  if (nbproc < 1 || nelements_total < nbproc) return fail(-80, "Wrong cpu in get_load_balancing_nelements!");
  int64_t nnz_total = 0;
  for (int32_t p = 0; p < nelements_total; ++p) nnz_total += sensit_nnz[p];
  // target cumulative nnz after rank c: (c+1) * (nnz_total / nbproc), the remainder goes to the last rank (:491-493)
  const int64_t share = nnz_total / nbproc;
  int32_t cpu = 0, nel_new = 0;
  int64_t nnz_new = 0, running = 0;
  for (int32_t c = 0; c < nbproc; ++c) { nnz_at_cpu_new[c] = 0; nelements_at_cpu_new[c] = 0; }
  for (int32_t p = 0; p < nelements_total; ++p) {
    nnz_new += sensit_nnz[p];
    running += sensit_nnz[p];
    nel_new += 1;
    const bool last = (p == nelements_total - 1);
    bool cut = last;
    if (!cut && cpu < nbproc - 1) {
      const int64_t target = share * (int64_t)(cpu + 1);
      cut = running >= target;
    }
    if (cut) {
      if (cpu >= nbproc) return fail(-80, "Wrong cpu in get_load_balancing_nelements!");
      nnz_at_cpu_new[cpu] = nnz_new;
      nelements_at_cpu_new[cpu] = nel_new;
      nnz_new = 0; nel_new = 0;
      ++cpu;
    }
  }
  if (cpu != nbproc) return fail(-80, "Wrong cpu in get_load_balancing_nelements!");
  int64_t chk = 0;
  for (int32_t c = 0; c < nbproc; ++c) chk += nnz_at_cpu_new[c];
  if (chk != nnz_total) return fail(-81, "Wrong nnz_at_cpu_new in get_load_balancing_nelements!");
  return 0;
}

extern "C" int tfx_sensit_rows_destroy(tfx_sensit_rows *rows) {
  delete rows;
  return 0;
}

// Applies combined_weight = real(problem_weight * data_weight(d, idata), 4) to a row set assembled with unit
// weights (the state in which it can be written to the reference's files): the in-HBM shortcut of
// read_sensitivity_kernel's weighting (:837-843), bit-identical to weighting at assembly time.
extern "C" int tfx_sensit_rows_apply_weights(tfx_sensit_rows *rows, double problem_weight, const double *data_weight) {
  TFX_TRY(ensure_init());
  if (!rows) return fail(-82, "sensit_rows: null handle");
  if (!rows->unit_weights) return fail(-89, "sensit_rows_apply_weights: the rows already carry weights");
  const tfx_sensit_params &P = rows->par;
  const size_t nl = (size_t)P.ndata * P.ndata_components;
  std::vector<float> w(nl);
  bool unit = true;
  for (size_t i = 0; i < nl; ++i) {
    w[i] = (float)(problem_weight * data_weight[i]);
    unit = unit && (w[i] == 1.0f);
  }
  rows->par.problem_weight = problem_weight;
  if (unit || rows->R.nnz == 0) return 0;
  DevBuf<float> dw;
  TFX_TRY(up(dw, w.data(), nl));
  Context &c = ctx();
  k_apply_row_weight<<<(int)std::min<int64_t>((rows->R.nnz + 255) / 256, (int64_t)c.num_sms * 16), 256, 0, c.stream>>>(
      rows->R.val.p, rows->R.rowid.p, rows->R.nnz, dw.p);
  c.launches++;
  TFX_CUDA(cudaStreamSynchronize(c.stream));
  rows->unit_weights = false;
  return 0;
}

extern "C" int tfx_sensit_rows_info(const tfx_sensit_rows *rows, int32_t *data0, int32_t *ndata_loc, int64_t *nnz_local) {
  if (!rows) return fail(-82, "sensit_rows: null handle");
  if (data0) *data0 = rows->data0;
  if (ndata_loc) *ndata_loc = rows->ndata_loc;
  if (nnz_local) *nnz_local = rows->R.nnz;
  return 0;
}

extern "C" int tfx_sensit_assemble_rows(tfx_sensit_rows **out, const tfx_sensit_params *par, const double *X1,
                                        const double *X2, const double *Y1, const double *Y2, const double *Z1,
                                        const double *Z2, const double *data_X, const double *data_Y,
                                        const double *data_Z, const double *column_weight_full,
                                        const double *data_weight, int32_t myrank, int32_t nbproc,
                                        int32_t *sensit_nnz, double *comp_error, int64_t *nnz_total) {
  TFX_TRY(ensure_init());
  cudaStream_t st = ctx().stream;
  if (!out || !par) return fail(-74, "sensit_assemble_rows: null handle");
  if (par->compression_rate < 0 || par->compression_rate > 1)
    return fail(-75, "Wrong compression rate! It must be between 0 and 1.");
  if (nbproc < 1 || myrank < 0 || myrank >= nbproc) return fail(-83, "sensit_assemble_rows: wrong rank");
  if (nbproc > 1 && comm_nranks() != nbproc)
    return fail(-84, "sensit_assemble_rows: nbproc does not match the communicator (tfx_comm_init)");
  const int64_t N64 = (int64_t)par->nx * par->ny * par->nz;
  if (N64 <= 0 || N64 * par->nmodel_components > 2000000000LL) return fail(-76, "calculate_sensit: wrong grid size");
  const int32_t N = (int32_t)N64;
  const int32_t ndc = par->ndata_components, nmc = par->nmodel_components;

  tfx_sensit_rows *h = new tfx_sensit_rows();
  h->par = *par;
  h->par.param_shift = 0;                 // columns stay k*N + p until the destination is known
  h->par.cell0 = 0; h->par.ncells_local = N;
  h->myrank = myrank; h->nbproc = nbproc;
  h->unit_weights = true;
  for (int64_t i = 0; i < (int64_t)par->ndata * ndc; ++i)
    if ((float)(par->problem_weight * data_weight[i]) != 1.0f) { h->unit_weights = false; break; }
  h->ndata_loc = nelements_at_cpu_even(par->ndata, myrank, nbproc);   // sensitivity_gravmag.F90:179-180
  h->data0 = nsmaller_even(par->ndata, myrank, nbproc);

  GridHold gh;
  DevBuf<double> dx, dy, dz, dcw;
  DevBuf<int32_t> dnnz;
  double err_sum = 0.0;
  trace("assemble_rows: begin");
  int rc = grid_acquire(gh, N, X1, X2, Y1, Y2, Z1, Z2, par->nx, par->ny, par->nz);
  trace("assemble_rows: grid on the device");
  if (!rc) rc = up(dx, data_X, par->ndata);
  if (!rc) rc = up(dy, data_Y, par->ndata);
  if (!rc) rc = up(dz, data_Z, par->ndata);
  if (!rc) rc = up(dcw, column_weight_full, N);
  trace("assemble_rows: stations + column weight");
  if (!rc)
    rc = assemble_rows_device(h->par, *gh.g, dx.p, dy.p, dz.p, dcw.p, data_weight, h->data0, h->ndata_loc, h->R, dnnz,
                              h->seg_end, &err_sum);
  trace("assemble_rows: row pipeline");
  if (rc) { delete h; return rc; }

  // reductions over ranks: sensit_nnz (:322), nnz_total (:327), compression error (:346-353)
  DevBuf<int64_t> dtot;
  DevBuf<double> derr;
  if (dtot.alloc(1) || derr.alloc(1)) { delete h; return -101; }
  int64_t tot = h->R.nnz;
  TFX_CUDA(cudaMemcpyAsync(dtot.p, &tot, 8, cudaMemcpyHostToDevice, st));
  TFX_CUDA(cudaMemcpyAsync(derr.p, &err_sum, 8, cudaMemcpyHostToDevice, st));
  if (nbproc > 1) {
    rc = comm_allreduce_sum_i32(dnnz.p, (size_t)N, st);
    if (!rc) rc = comm_allreduce_sum_i64(dtot.p, 1, st);
    if (!rc) rc = comm_allreduce_sum(derr.p, 1, st);
    if (rc) { delete h; return rc; }
  }
  TFX_CUDA(cudaMemcpyAsync(&tot, dtot.p, 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaMemcpyAsync(&err_sum, derr.p, 8, cudaMemcpyDeviceToHost, st));
  if (sensit_nnz) TFX_CUDA(cudaMemcpyAsync(sensit_nnz, dnnz.p, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  if (nnz_total) *nnz_total = tot;
  if (comp_error)
    *comp_error = (par->compression_type > 0) ? err_sum / ((double)par->ndata * ndc * nmc) : 0.0;
  *out = h;
  trace("assemble_rows: reductions + sensit_nnz");
  return 0;
}

// Core of the re-partitioning: leaves this rank's slab as triplets sorted by (row, local column).
static int repartition_core(tfx_sensit_rows *rows, int32_t problem_slot, const int32_t *nelements_at_cpu, int32_t myrank,
                            int32_t nbproc, RowTriplets &Rx, int32_t *nl_out, int32_t *ncolumns_out) {
  TFX_TRY(ensure_init());
  trace("repartition: begin");
  Context &c = ctx();
  cudaStream_t st = c.stream;
  if (!rows) return fail(-82, "sensit_repartition: null handle");
  if (problem_slot != 1 && problem_slot != 2) return fail(-85, "sensit_repartition: problem_slot must be 1 or 2");
  if (nbproc < 1 || myrank < 0 || myrank >= nbproc) return fail(-83, "sensit_repartition: wrong rank");
  const tfx_sensit_params &P = rows->par;
  const int32_t N = P.nx * P.ny * P.nz, nmc = P.nmodel_components, ndc = P.ndata_components;
  const int nranks = comm_nranks();
  // Single-process mode: the handle holds ALL rows and the slab of `myrank` is built without an exchange.
  const bool single = (nranks == 1);
  if (single && rows->ndata_loc != P.ndata)
    return fail(-86, "sensit_repartition: without a communicator the row set must hold all data rows");
  if (!single && (nranks != nbproc || rows->nbproc != nbproc || rows->myrank != myrank))
    return fail(-84, "sensit_repartition: nbproc / myrank do not match the communicator and the row set");
  std::vector<int32_t> cum((size_t)nbproc + 1, 0);
  for (int32_t r = 0; r < nbproc; ++r) {
    if (nelements_at_cpu[r] < 0) return fail(-87, "sensit_repartition: negative nelements_at_cpu");
    cum[r + 1] = cum[r] + nelements_at_cpu[r];
  }
  if (cum[nbproc] != N) return fail(-87, "sensit_repartition: nelements_at_cpu does not sum to nx*ny*nz");

  RowTriplets &R = rows->R;
  const int64_t nseg = (int64_t)rows->ndata_loc * ndc * nmc;
  const int64_t nnz_loc = R.nnz;
  const int32_t P1 = nbproc + 1;

  // ---- piece bounds, counts, send offsets
  std::vector<int64_t> seg_beg((size_t)nseg + 1, 0);
  for (int64_t s = 0; s < nseg; ++s) seg_beg[s + 1] = rows->seg_end[(size_t)s];
  DevBuf<int64_t> d_segbeg, d_bound, d_off;
  DevBuf<int32_t> d_cum;
  TFX_TRY(up(d_segbeg, seg_beg.data(), seg_beg.size()));
  TFX_TRY(up(d_cum, cum.data(), cum.size()));
  TFX_TRY(d_bound.alloc((size_t)std::max<int64_t>(nseg * P1, 1)));
  TFX_TRY(d_off.alloc((size_t)(nseg * nbproc + 1)));
  std::vector<int64_t> send_off((size_t)nbproc + 1, 0);
  if (nseg > 0) {
    const int g1 = (int)std::min<int64_t>((nseg * P1 + 255) / 256, (int64_t)c.num_sms * 16);
    k_piece_bounds<<<g1, 256, 0, st>>>(R.idx.p, d_segbeg.p, nseg, nmc, N, nbproc, d_cum.p, d_bound.p);
    k_piece_counts<<<g1, 256, 0, st>>>(d_bound.p, nseg, nbproc, d_off.p);
    TFX_CUDA(cudaMemsetAsync(d_off.p + nseg * nbproc, 0, 8, st));
    thrust::device_ptr<int64_t> O(d_off.p);
    TFX_THRUST(thrust::exclusive_scan(thrust::cuda::par.on(st), O, O + nseg * nbproc + 1, O));
    c.launches += 4;
    for (int32_t r = 0; r <= nbproc; ++r)
      TFX_CUDA(cudaMemcpyAsync(&send_off[r], d_off.p + (int64_t)r * nseg, 8, cudaMemcpyDeviceToHost, st));
    TFX_CUDA(cudaStreamSynchronize(st));
    if (send_off[nbproc] != nnz_loc) return fail(-88, "sensit_repartition: piece counts do not add up to nnz");
  }

  // ---- counts of every (source, destination) pair
  std::vector<int64_t> counts((size_t)nbproc * nbproc, 0);   // counts[src*nbproc + dst]
  if (single) {
    // only the row "source 0 -> myrank" matters
    counts[(size_t)myrank] = send_off[myrank + 1] - send_off[myrank];
  } else {
    std::vector<int64_t> mine((size_t)nbproc);
    for (int32_t r = 0; r < nbproc; ++r) mine[r] = send_off[r + 1] - send_off[r];
    DevBuf<int64_t> d_mine, d_all;
    TFX_TRY(up(d_mine, mine.data(), mine.size()));
    TFX_TRY(d_all.alloc(counts.size()));
    TFX_TRY(comm_allgather_i64(d_mine.p, d_all.p, (size_t)nbproc, st));
    TFX_CUDA(cudaMemcpyAsync(counts.data(), d_all.p, counts.size() * 8, cudaMemcpyDeviceToHost, st));
    TFX_CUDA(cudaStreamSynchronize(st));
  }

  // ---- pack (destination-major, segment order inside a destination)
  RowTriplets S;   // send buffers
  const int32_t only_dest = single ? myrank : -1;
  int64_t send_total = single ? (send_off[myrank + 1] - send_off[myrank]) : nnz_loc;
  TFX_TRY(S.idx.alloc((size_t)std::max<int64_t>(send_total, 1)));
  TFX_TRY(S.val.alloc((size_t)std::max<int64_t>(send_total, 1)));
  TFX_TRY(S.rowid.alloc((size_t)std::max<int64_t>(send_total, 1)));
  if (nseg > 0 && send_total > 0) {
    const int64_t base = single ? send_off[myrank] : 0;   // single: the destination's piece starts at 0
    const int g2 = (int)std::min<int64_t>(nseg, (int64_t)c.num_sms * 16);
    k_pack_pieces<<<g2, 256, 0, st>>>(R.idx.p, R.val.p, R.rowid.p, d_bound.p, d_off.p, nseg, nmc, N, nbproc, d_cum.p,
                                      problem_slot - 1, only_dest, base, S.idx.p, S.val.p, S.rowid.p);
    c.launches++;
  }
  TFX_CUDA(cudaStreamSynchronize(st));
  R.idx.release(); R.val.release(); R.rowid.release(); R.nnz = 0;
  rows->seg_end.clear();
  rows->ndata_loc = 0;

  // ---- exchange
  if (single) {
    std::swap(Rx.idx.p, S.idx.p); std::swap(Rx.idx.n, S.idx.n);
    std::swap(Rx.val.p, S.val.p); std::swap(Rx.val.n, S.val.n);
    std::swap(Rx.rowid.p, S.rowid.p); std::swap(Rx.rowid.n, S.rowid.n);
    Rx.nnz = send_total;
  } else {
    std::vector<int64_t> recv_off((size_t)nbproc + 1, 0);
    for (int32_t q = 0; q < nbproc; ++q) recv_off[q + 1] = recv_off[q] + counts[(size_t)q * nbproc + myrank];
    const int64_t nrecv = recv_off[nbproc];
    TFX_TRY(Rx.idx.alloc((size_t)std::max<int64_t>(nrecv, 1)));
    TFX_TRY(Rx.val.alloc((size_t)std::max<int64_t>(nrecv, 1)));
    TFX_TRY(Rx.rowid.alloc((size_t)std::max<int64_t>(nrecv, 1)));
    TFX_TRY(comm_alltoallv_4b(S.idx.p, send_off.data(), Rx.idx.p, recv_off.data(), st));
    TFX_TRY(comm_alltoallv_4b(S.val.p, send_off.data(), Rx.val.p, recv_off.data(), st));
    TFX_TRY(comm_alltoallv_4b(S.rowid.p, send_off.data(), Rx.rowid.p, recv_off.data(), st));
    TFX_CUDA(cudaStreamSynchronize(st));
    S.idx.release(); S.val.release(); S.rowid.release();
    Rx.nnz = nrecv;
  }

  // the column slab of this rank: all data rows, local columns (joint_inverse_problem.F90:213-214)
  *nl_out = P.ndata * ndc;
  *ncolumns_out = 2 * nmc * nelements_at_cpu[myrank];
  return 0;
}

extern "C" int tfx_sensit_repartition(tfx_matrix **out, tfx_sensit_rows *rows, int32_t problem_slot,
                                      const int32_t *nelements_at_cpu, int32_t myrank, int32_t nbproc) {
  if (!out) return fail(-82, "sensit_repartition: null handle");
  RowTriplets Rx;
  int32_t nl = 0, ncolumns = 0;
  TFX_TRY(repartition_core(rows, problem_slot, nelements_at_cpu, myrank, nbproc, Rx, &nl, &ncolumns));
  tfx_matrix *h = new tfx_matrix();
  int rc = matrix_from_triplets(h->m, nl, ncolumns, Rx);
  if (rc) { delete h; return rc; }
  TFX_CUDA(cudaStreamSynchronize(ctx().stream));
  *out = h;
  return 0;
}

// Same, but the rows are APPENDED to a matrix under construction (initialize ... finalize), like the reference's
// read_sensitivity_kernel which is called once per problem on jinv%matrix_sensit (problem_joint_gravmag.F90:241-248).
extern "C" int tfx_sensit_repartition_into(tfx_matrix *matrix_sensit, tfx_sensit_rows *rows, int32_t problem_slot,
                                           const int32_t *nelements_at_cpu, int32_t myrank, int32_t nbproc) {
  if (!matrix_sensit) return fail(-82, "sensit_repartition_into: null handle");
  RowTriplets Rx;
  int32_t nl = 0, ncolumns = 0;
  TFX_TRY(repartition_core(rows, problem_slot, nelements_at_cpu, myrank, nbproc, Rx, &nl, &ncolumns));
  if (g_opt_sensit_row_blocks) return matrix_append_block(matrix_sensit->m, Rx, nl, ncolumns);
  return matrix_append_triplets(matrix_sensit->m, Rx, nl, ncolumns);
}
