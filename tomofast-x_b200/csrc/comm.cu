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
  if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllReduce) return fail(-61, "NCCL symbols missing");
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
