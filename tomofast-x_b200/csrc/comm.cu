// comm.cu -- the solver's reductions over NVLink/NVSwitch (NCCL), one rank per GPU.
//
// Stands in for the MPI_COMM_WORLD collectives on the hot path: MPI_Allreduce(u, nlines)
// (src/inversion/lsqr_solver2.F90:214), the scalar Allreduce of normalize() (:514) and
// MPI_Allreduce(data_calc) (src/inversion/model.F90:293).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, the copy torch already loaded when the host
// is Python), so single-GPU use and the CPU-side symbol checks need no NCCL at link time.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "matrix.h"

namespace tfx {

namespace {
struct Nccl {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};
Nccl &N() {
  static Nccl n;
  return n;
}

int load_nccl() {
  Nccl &n = N();
  if (n.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.h) break;
  }
  if (!n.h) return fail(-60, std::string("cannot load NCCL: ") + dlerror());
  n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(n.h, "ncclGetUniqueId");
  n.CommInitRank = (decltype(n.CommInitRank))dlsym(n.h, "ncclCommInitRank");
  n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.h, "ncclCommDestroy");
  n.AllReduce = (decltype(n.AllReduce))dlsym(n.h, "ncclAllReduce");
  n.GetErrorString = (decltype(n.GetErrorString))dlsym(n.h, "ncclGetErrorString");
  n.AllGather = (decltype(n.AllGather))dlsym(n.h, "ncclAllGather");
  n.Send = (decltype(n.Send))dlsym(n.h, "ncclSend");
  n.Recv = (decltype(n.Recv))dlsym(n.h, "ncclRecv");
  n.Broadcast = (decltype(n.Broadcast))dlsym(n.h, "ncclBroadcast");
  n.GroupStart = (decltype(n.GroupStart))dlsym(n.h, "ncclGroupStart");
  n.GroupEnd = (decltype(n.GroupEnd))dlsym(n.h, "ncclGroupEnd");
  if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllReduce || !n.AllGather || !n.Send || !n.Recv ||
      !n.GroupStart || !n.GroupEnd || !n.Broadcast)
    return fail(-61, "NCCL symbols missing");
  return 0;
}
}  // namespace

int comm_nranks() { return N().comm ? N().nranks : 1; }

int comm_allreduce_sum(double *d_buf, size_t count, cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) return 0;
  ncclResult_t r = n.AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, n.comm, st);
  if (r != ncclSuccess) return fail(-62, std::string("ncclAllReduce failed: ") + (n.GetErrorString ? n.GetErrorString(r) : "?"));
  return 0;
}

int comm_rank() { return N().comm ? N().rank : 0; }

static int nccl_fail(const char *what, ncclResult_t r) {
  Nccl &n = N();
  return fail(-62, std::string(what) + " failed: " + (n.GetErrorString ? n.GetErrorString(r) : "?"));
}

// MPI_Allreduce(sensit_nnz, MPI_INTEGER) (sensitivity_gravmag.F90:322) and the INTEGER8 sum of nnz (:327).
int comm_allreduce_sum_i32(int32_t *d_buf, size_t count, cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) return 0;
  ncclResult_t r = n.AllReduce(d_buf, d_buf, count, ncclInt32, ncclSum, n.comm, st);
  return r == ncclSuccess ? 0 : nccl_fail("ncclAllReduce", r);
}
// Per-row count of ranks holding entries of a constraint row (lsqr.cu build_plan): small integers, one byte each.
int comm_allreduce_sum_u8(uint8_t *d_buf, size_t count, cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) return 0;
  ncclResult_t r = n.AllReduce(d_buf, d_buf, count, ncclUint8, ncclSum, n.comm, st);
  return r == ncclSuccess ? 0 : nccl_fail("ncclAllReduce", r);
}
int comm_allreduce_sum_i64(int64_t *d_buf, size_t count, cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) return 0;
  ncclResult_t r = n.AllReduce(d_buf, d_buf, count, ncclInt64, ncclSum, n.comm, st);
  return r == ncclSuccess ? 0 : nccl_fail("ncclAllReduce", r);
}
int comm_allreduce_max(double *d_buf, size_t count, cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) return 0;
  ncclResult_t r = n.AllReduce(d_buf, d_buf, count, ncclDouble, ncclMax, n.comm, st);
  return r == ncclSuccess ? 0 : nccl_fail("ncclAllReduce", r);
}
// recv[r*count .. (r+1)*count) = rank r's send[0 .. count)
int comm_allgather_i64(const int64_t *d_send, int64_t *d_recv, size_t count, cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) {
    if (d_send != d_recv) TFX_CUDA(cudaMemcpyAsync(d_recv, d_send, count * 8, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  ncclResult_t r = n.AllGather(d_send, d_recv, count, ncclInt64, n.comm, st);
  return r == ncclSuccess ? 0 : nccl_fail("ncclAllGather", r);
}
// All-to-all of 4-byte elements with per-peer counts (element offsets, nranks + 1 entries each): the
// device-to-device replacement of the per-row MPI_Scatterv of read_sensitivity_kernel
// (sensitivity_gravmag.F90:818-829). One grouped ncclSend/ncclRecv pair per peer over NVLink.
int comm_alltoallv_4b(const void *d_send, const int64_t *send_off, void *d_recv, const int64_t *recv_off,
                      cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) {
    const int64_t cnt = send_off[1] - send_off[0];
    if (cnt > 0)
      TFX_CUDA(cudaMemcpyAsync((char *)d_recv + recv_off[0] * 4, (const char *)d_send + send_off[0] * 4, (size_t)cnt * 4,
                               cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  ncclResult_t r = n.GroupStart();
  if (r != ncclSuccess) return nccl_fail("ncclGroupStart", r);
  for (int q = 0; q < n.nranks; ++q) {
    const int64_t ns = send_off[q + 1] - send_off[q], nr = recv_off[q + 1] - recv_off[q];
    if (ns > 0) {
      r = n.Send((const char *)d_send + send_off[q] * 4, (size_t)ns, ncclInt32, q, n.comm, st);
      if (r != ncclSuccess) { n.GroupEnd(); return nccl_fail("ncclSend", r); }
    }
    if (nr > 0) {
      r = n.Recv((char *)d_recv + recv_off[q] * 4, (size_t)nr, ncclInt32, q, n.comm, st);
      if (r != ncclSuccess) { n.GroupEnd(); return nccl_fail("ncclRecv", r); }
    }
  }
  r = n.GroupEnd();
  return r == ncclSuccess ? 0 : nccl_fail("ncclGroupEnd", r);
}

// In-place all-gather of slabs of different lengths: rank r's slab already sits at d_full + offsets[r]
// (offsets has nranks + 1 entries). Replaces get_full_array (MPI_Gatherv to
// rank 0, parallel_tools.f90:147) + the scatter back: every GPU ends up with the whole vector, moving half the bytes of
// the all-reduce-into-zeros it replaces.
int comm_allgatherv_f64(double *d_full, const int64_t *offsets, cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) return 0;
  // One grouped ncclBroadcast per slab. Measured at 8 GPUs on a 1.07 GB volume with nnz-balanced (unequal) slabs
  // (profiles/r2_bench_8gpu_a.jsonl vs gpurun_out/r2_bench_8gpu_e.jsonl): 3.8 ms this way, 5.8 ms as grouped
  // ncclSend/ncclRecv pairs -- the rank with the largest slab would have to push it to 7 peers itself, the broadcast lets
  // the NVSwitch replicate it.
  ncclResult_t r = n.GroupStart();
  if (r != ncclSuccess) return nccl_fail("ncclGroupStart", r);
  for (int q = 0; q < n.nranks; ++q) {
    const int64_t cnt = offsets[q + 1] - offsets[q];
    if (cnt <= 0) continue;
    r = n.Broadcast(d_full + offsets[q], d_full + offsets[q], (size_t)cnt, ncclDouble, q, n.comm, st);
    if (r != ncclSuccess) { n.GroupEnd(); return nccl_fail("ncclBroadcast", r); }
  }
  r = n.GroupEnd();
  return r == ncclSuccess ? 0 : nccl_fail("ncclGroupEnd", r);
}

// Grouped point-to-point exchange of doubles: for every peer q, send_cnt[q] doubles from d_send + send_off[q] and
// recv_cnt[q] doubles into d_recv + recv_off[q] (the own block is NOT copied: callers place it themselves).
int comm_exchange_f64(const double *d_send, const int64_t *send_off, const int64_t *send_cnt, double *d_recv,
                      const int64_t *recv_off, const int64_t *recv_cnt, cudaStream_t st) {
  Nccl &n = N();
  if (!n.comm || n.nranks <= 1) return 0;
  ncclResult_t r = n.GroupStart();
  if (r != ncclSuccess) return nccl_fail("ncclGroupStart", r);
  for (int q = 0; q < n.nranks; ++q) {
    if (q == n.rank) continue;
    if (send_cnt[q] > 0) {
      r = n.Send(d_send + send_off[q], (size_t)send_cnt[q], ncclDouble, q, n.comm, st);
      if (r != ncclSuccess) { n.GroupEnd(); return nccl_fail("ncclSend", r); }
    }
    if (recv_cnt[q] > 0) {
      r = n.Recv(d_recv + recv_off[q], (size_t)recv_cnt[q], ncclDouble, q, n.comm, st);
      if (r != ncclSuccess) { n.GroupEnd(); return nccl_fail("ncclRecv", r); }
    }
  }
  r = n.GroupEnd();
  return r == ncclSuccess ? 0 : nccl_fail("ncclGroupEnd", r);
}

// Slab lengths of all ranks as prefix offsets (nranks + 1 entries): one small all-gather + host read.
int comm_slab_offsets(int64_t mine, std::vector<int64_t> &offsets) {
  Nccl &n = N();
  offsets.assign(2, 0);
  offsets[1] = mine;
  if (!n.comm || n.nranks <= 1) return 0;
  cudaStream_t st = ctx().stream;
  DevBuf<int64_t> d_mine, d_all;
  TFX_TRY(d_mine.alloc(1)); TFX_TRY(d_all.alloc((size_t)n.nranks));
  TFX_CUDA(cudaMemcpyAsync(d_mine.p, &mine, 8, cudaMemcpyHostToDevice, st));
  TFX_TRY(comm_allgather_i64(d_mine.p, d_all.p, 1, st));
  std::vector<int64_t> all((size_t)n.nranks);
  TFX_CUDA(cudaMemcpyAsync(all.data(), d_all.p, all.size() * 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  offsets.assign((size_t)n.nranks + 1, 0);
  for (int r = 0; r < n.nranks; ++r) offsets[(size_t)r + 1] = offsets[(size_t)r] + all[(size_t)r];
  return 0;
}

int comm_slab_offset(int64_t mine, int64_t *offset, int64_t *total) {
  Nccl &n = N();
  *offset = 0; *total = mine;
  if (!n.comm || n.nranks <= 1) return 0;
  cudaStream_t st = ctx().stream;
  DevBuf<int64_t> d_mine, d_all;
  TFX_TRY(d_mine.alloc(1)); TFX_TRY(d_all.alloc((size_t)n.nranks));
  TFX_CUDA(cudaMemcpyAsync(d_mine.p, &mine, 8, cudaMemcpyHostToDevice, st));
  TFX_TRY(comm_allgather_i64(d_mine.p, d_all.p, 1, st));
  std::vector<int64_t> all((size_t)n.nranks);
  TFX_CUDA(cudaMemcpyAsync(all.data(), d_all.p, all.size() * 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  int64_t off = 0, tot = 0;
  for (int r = 0; r < n.nranks; ++r) {
    if (r < n.rank) off += all[(size_t)r];
    tot += all[(size_t)r];
  }
  *offset = off; *total = tot;
  return 0;
}

int comm_unique_id(char id[128]) {
  TFX_TRY(load_nccl());
  ncclUniqueId uid;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclResult_t r = N().GetUniqueId(&uid);
  if (r != ncclSuccess) return fail(-63, "ncclGetUniqueId failed");
  memcpy(id, &uid, 128);
  return 0;
}

int comm_init(int nranks, int rank, const char id[128]) {
  TFX_TRY(ensure_init());
  Nccl &n = N();
  if (nranks <= 1) {
    n.nranks = 1;
    n.rank = 0;
    return 0;
  }
  TFX_TRY(load_nccl());
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclResult_t r = n.CommInitRank(&n.comm, nranks, uid, rank);
  if (r != ncclSuccess) return fail(-64, std::string("ncclCommInitRank failed: ") + (n.GetErrorString ? n.GetErrorString(r) : "?"));
  n.nranks = nranks;
  n.rank = rank;
  // NCCL sets up the ring/tree and the point-to-point channels lazily on first use (seconds with 8 peers): pay for
  // it here, once, instead of inside the first LSQR iteration / the first re-partitioning.
  {
    cudaStream_t st = ctx().stream;
    DevBuf<double> w;
    DevBuf<int32_t> a, b;
    TFX_TRY(w.alloc(8)); TFX_TRY(a.alloc((size_t)nranks)); TFX_TRY(b.alloc((size_t)nranks));
    TFX_TRY(w.zero()); TFX_TRY(a.zero());
    TFX_TRY(comm_allreduce_sum(w.p, 8, st));
    std::vector<int64_t> off((size_t)nranks + 1);
    for (int q = 0; q <= nranks; ++q) off[(size_t)q] = q;
    TFX_TRY(comm_alltoallv_4b(a.p, off.data(), b.p, off.data(), st));
    TFX_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

int comm_finalize() {
  Nccl &n = N();
  if (n.comm) {
    n.CommDestroy(n.comm);
    n.comm = nullptr;
  }
  n.nranks = 1;
  n.rank = 0;
  return 0;
}

}  // namespace tfx
