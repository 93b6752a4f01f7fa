// data.cu -- the callers either side of the products: apply_wavelet_transform on a distributed model vector
// (src/inversion/wavelet_utils.F90:37-72) and t_model%calculate_data (src/inversion/model.F90:220-307), kept on
// the device so the model never leaves HBM between solves.
#include "../../include/tfx.h"

#include <string.h>

#include <algorithm>
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

int g_opt_wavelet_dist = 1;   // 1: plane-owner / column-owner transform when the slabs allow it, 0: always gather
int g_wavelet_last_dist = 0;  // diagnostic: 1 / 2 when the last slab transform ran distributed with NCCL / peer-memory exchanges

namespace {
struct DistBufs {
  DevBuf<double> A, B, stage;
};
DistBufs &dist_bufs() {
  static DistBufs b;
  return b;
}
inline int64_t clampi(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Peer-memory exchange (r2). The two layout changes of the distributed transform move ~V/N doubles per GPU each; through
// grouped ncclSend/ncclRecv they went pack (strided copy into a staging buffer) -> NCCL -> unpack. Here every rank's A, B
// and staging buffers are mapped into all the other processes (cudaIpc over NVLink / NVSwitch peer access) and ONE kernel
// per layout change reads the local layout and stores straight into the peers' buffers at their final places -- pack,
// transfer and unpack in the same pass, all NVLinks busy at once. Ordering between the ranks: a one-element all-reduce on
// the stream after each exchange (every rank's stores are complete when its kernel is, and nobody leaves the collective
// before everybody entered it). Falls back to the NCCL path when IPC mapping is not available.
// ---------------------------------------------------------------------------------------------------------------------
int g_opt_wavelet_p2p = 1;

namespace {
const int kMaxPeers = 16;
struct PeerPtrs {
  double *p[kMaxPeers];
};
struct PeerState {
  bool failed = false;                 // IPC not available: NCCL path from now on
  int nr = 0;
  std::vector<size_t> capA, capB, capS; // capacities (doubles) of EVERY rank's buffers: the same numbers on all ranks
  DevBuf<double> A, B, S, flag;
  PeerPtrs pA, pB, pS;                  // pX.p[r]: rank r's buffer in this process (own buffer for r == me)
  bool mapped = false;
};
PeerState &peer_state() {
  static PeerState s;
  return s;
}

void peer_unmap(PeerState &P, int me) {
  if (!P.mapped) return;
  for (int r = 0; r < P.nr; ++r) {
    if (r == me) continue;
    if (P.pA.p[r]) cudaIpcCloseMemHandle(P.pA.p[r]);
    if (P.pB.p[r]) cudaIpcCloseMemHandle(P.pB.p[r]);
    if (P.pS.p[r]) cudaIpcCloseMemHandle(P.pS.p[r]);
    P.pA.p[r] = P.pB.p[r] = P.pS.p[r] = nullptr;
  }
  P.mapped = false;
}

// Makes sure every rank's buffers hold needX[r] doubles and are mapped everywhere. All ranks call this with the same
// arguments (they all know the whole plan), so they agree on when buffers are re-allocated and handles re-exchanged.
int peer_ensure(PeerState &P, const std::vector<size_t> &needA, const std::vector<size_t> &needB,
                const std::vector<size_t> &needS, int me, int nr, cudaStream_t st) {
  if (P.failed || nr > kMaxPeers) return 1;
  if (P.nr != nr) {
    peer_unmap(P, me);
    P.nr = nr;
    P.capA.assign((size_t)nr, 0); P.capB.assign((size_t)nr, 0); P.capS.assign((size_t)nr, 0);
  }
  bool grow = !P.mapped;
  for (int r = 0; r < nr; ++r)
    if (needA[r] > P.capA[r] || needB[r] > P.capB[r] || needS[r] > P.capS[r]) grow = true;
  if (!grow) return 0;
  TFX_CUDA(cudaStreamSynchronize(st));
  peer_unmap(P, me);
  for (int r = 0; r < nr; ++r) {
    P.capA[r] = std::max(P.capA[r], needA[r]); P.capB[r] = std::max(P.capB[r], needB[r]);
    P.capS[r] = std::max(P.capS[r], needS[r]);
  }
  // every rank has passed its last use of the old buffers before any of them is freed
  TFX_TRY(P.flag.alloc(8));
  TFX_CUDA(cudaMemsetAsync(P.flag.p, 0, 64, st));
  TFX_TRY(comm_allreduce_sum(P.flag.p, 1, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  P.A.release(); P.B.release(); P.S.release();
  TFX_TRY(P.A.alloc(P.capA[me])); TFX_TRY(P.B.alloc(P.capB[me])); TFX_TRY(P.S.alloc(P.capS[me]));
  // handles: 3 x 64 bytes per rank, all-gathered as int64
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  int64_t mine[24];
  cudaIpcMemHandle_t h[3];
  int bad = 0;
  if (cudaIpcGetMemHandle(&h[0], P.A.p) != cudaSuccess || cudaIpcGetMemHandle(&h[1], P.B.p) != cudaSuccess ||
      cudaIpcGetMemHandle(&h[2], P.S.p) != cudaSuccess) {
    cudaGetLastError();
    bad = 1;
    memset(h, 0, sizeof(h));
  }
  memcpy(mine, h, sizeof(h));
  DevBuf<int64_t> dmine, dall;
  TFX_TRY(dmine.alloc(24)); TFX_TRY(dall.alloc((size_t)24 * nr));
  TFX_CUDA(cudaMemcpyAsync(dmine.p, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  TFX_TRY(comm_allgather_i64(dmine.p, dall.p, 24, st));
  std::vector<int64_t> all((size_t)24 * nr);
  TFX_CUDA(cudaMemcpyAsync(all.data(), dall.p, all.size() * 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < nr; ++r) P.pA.p[r] = P.pB.p[r] = P.pS.p[r] = nullptr;
  P.pA.p[me] = P.A.p; P.pB.p[me] = P.B.p; P.pS.p[me] = P.S.p;
  for (int r = 0; r < nr && !bad; ++r) {
    if (r == me) continue;
    cudaIpcMemHandle_t hr[3];
    memcpy(hr, all.data() + (size_t)24 * r, sizeof(hr));
    void *a = nullptr, *b = nullptr, *c = nullptr;
    if (cudaIpcOpenMemHandle(&a, hr[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&b, hr[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&c, hr[2], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      bad = 1;
    }
    P.pA.p[r] = (double *)a; P.pB.p[r] = (double *)b; P.pS.p[r] = (double *)c;
  }
  P.mapped = true;
  // everybody maps or nobody uses the path
  double hb = bad ? 1.0 : 0.0;
  TFX_CUDA(cudaMemcpyAsync(P.flag.p, &hb, 8, cudaMemcpyHostToDevice, st));
  TFX_TRY(comm_allreduce_sum(P.flag.p, 1, st));
  TFX_CUDA(cudaMemcpyAsync(&hb, P.flag.p, 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  if (hb != 0.0) {
    peer_unmap(P, me);
    P.failed = true;
    return 1;
  }
  return 0;
}

// Unpack of the B -> slabs exchange in one launch: the piece of source rank q (blockIdx.y) is the contiguous run
// [first row from column c0, then whole rows of width W_q from column pa_q, the last row cut short]; element e of it is
// cell (row, column) of the volume and lands at d_slab[row * plane + column - off_me].
struct UnpackArgs {
  const double *src[kMaxPeers];
  long long cnt[kMaxPeers], w0[kMaxPeers], c0[kMaxPeers], paq[kMaxPeers], Wq[kMaxPeers];
  long long k0, plane, off_me;
  int nr;
};
__global__ void __launch_bounds__(256) k_unpack_slab(double *__restrict__ slab, UnpackArgs a) {
  const int q = blockIdx.y;
  const long long n = a.cnt[q];
  if (n <= 0) return;
  const double *src = a.src[q];
  const long long w0 = a.w0[q], c0 = a.c0[q], paq = a.paq[q], Wq = a.Wq[q];
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < n; e += 256LL * gridDim.x) {
    long long row, col;
    if (e < w0) {
      row = a.k0; col = c0 + e;
    } else {
      const long long r = (e - w0) / Wq;
      row = a.k0 + 1 + r; col = paq + (e - w0) - r * Wq;
    }
    slab[row * a.plane + col - a.off_me] = src[e];
  }
}
}  // namespace
// Drops the mappings of the other processes' buffers (before the communicator goes away).
void wavelet_peer_reset() {
  PeerState &P = peer_state();
  peer_unmap(P, comm_rank());
  P.A.release(); P.B.release(); P.S.release();
  P.nr = 0;
  P.failed = false;
}
namespace {
int peer_barrier(PeerState &P, cudaStream_t st) { return comm_allreduce_sum(P.flag.p + 1, 1, st); }

// Layout A -> layout B of every rank: rank q receives the columns [pa[q], pa[q+1]) of my nk planes as the rows
// ka_me .. ka_me + nk of its B (row length W_q). blockIdx.z = q, blockIdx.y = plane, x over the columns.
struct ScatterArgs {
  PeerPtrs dst;
  long long pa[kMaxPeers + 1];
  long long plane, ka_me;
  int nk, nr;
};
__global__ void __launch_bounds__(256) k_scatter_columns(const double *__restrict__ A, ScatterArgs a) {
  const int q = blockIdx.z;
  const long long p0 = a.pa[q], W = a.pa[q + 1] - p0;
  double *dst = a.dst.p[q];
  for (long long k = blockIdx.y; k < a.nk; k += gridDim.y) {
    const double *src = A + k * a.plane + p0;
    double *d = dst + (a.ka_me + k) * W;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < W; i += 256LL * gridDim.x) d[i] = src[i];
  }
}
// Layout B -> the slabs: the cells of slab q inside my columns are one contiguous range [b0[q], b1[q]) of my B; they go
// to rank q's staging buffer at offset roff[q] (where rank q's unpack expects what comes from me). q == me is skipped
// (unpacked straight from B).
struct RangeArgs {
  PeerPtrs dst;
  long long b0[kMaxPeers], cnt[kMaxPeers], roff[kMaxPeers];
  int nr, me;
};
__global__ void __launch_bounds__(256) k_scatter_ranges(const double *__restrict__ B, RangeArgs a) {
  const int q = blockIdx.y;
  if (q == a.me) return;
  const double *src = B + a.b0[q];
  double *d = a.dst.p[q] + a.roff[q];
  const long long n = a.cnt[q];
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) d[i] = src[i];
}
}  // namespace

// Distributed 3-D transform of a volume held as contiguous cell slabs (i fastest, then j, then k; slab r = cells
// [off[r], off[r+1])), WITHOUT assembling the volume anywhere:
//   layout A: rank r holds the complete k-planes whose first cell lies in its slab -> the axis-1 and axis-2 passes are
//             local (they never leave a plane); only the piece of a plane that straddles a slab boundary moves
//             (<= one plane, to the neighbouring rank);
//   layout B: rank q holds, for ALL k, the cells p in [pa[q], pa[q+1]) of every plane -> the axis-3 pass is local;
//   A -> B and B -> slabs are two all-to-all exchanges over NVLink: ~2 V/N doubles per GPU instead of the (N-1) V/N an
//   all-gather receives, and every GPU transforms V/N cells instead of V.
// Same kernels on the same lines in the same order: bit-identical to the serial transform. Returns 1 in *done when the
// slab layout qualifies (every slab starts at most one plane before a plane it owns), 0 to fall back to the gather.
static int wavelet_slab_dist(double *d_slab, const std::vector<int64_t> &off, int nx, int ny, int nz, int wavelet_type,
                             bool forward, cudaStream_t st, int *done) {
  *done = 0;
  const int nr = comm_nranks(), me = comm_rank();
  const int64_t plane = (int64_t)nx * ny, n3 = nz;
  if (!g_opt_wavelet_dist || nr <= 1 || plane < nr || n3 < 1) return 0;
  // ---- ownership of planes (A) and of plane columns (B)
  std::vector<int64_t> ka((size_t)nr + 1), pa((size_t)nr + 1);
  for (int r = 0; r <= nr; ++r) {
    ka[r] = (r == nr) ? n3 : (off[r] + plane - 1) / plane;      // first plane whose first cell is >= off[r]
    pa[r] = plane * r / nr;
  }
  for (int r = 0; r < nr; ++r) {
    if (ka[r] >= ka[r + 1]) return 0;                              // a rank without a plane of its own
    if (r > 0 && (ka[r] - 1) * plane < off[r - 1]) return 0;       // its leading piece would skip a rank
  }
  const int64_t nk = ka[me + 1] - ka[me];                          // planes this rank owns
  const int64_t lead = ka[me] * plane - off[me];                   // cells of my slab that belong to rank me-1's plane
  const int64_t own = off[me + 1] - ka[me] * plane;                // cells of my planes I hold myself
  const int64_t trail = ka[me + 1] * plane - off[me + 1];          // ... and the rest comes from rank me+1
  const int64_t W = pa[me + 1] - pa[me];
  auto range_in = [&](int q, int r, int64_t *b0, int64_t *b1) {     // range of rank q's B that belongs to slab r
    const int64_t Wq = pa[q + 1] - pa[q];
    const int64_t k0 = off[r] / plane, s0 = off[r] - k0 * plane;
    const int64_t k1 = (off[r + 1] - 1) / plane, e1 = off[r + 1] - k1 * plane;
    *b0 = k0 * Wq + clampi(s0, pa[q], pa[q + 1]) - pa[q];
    *b1 = k1 * Wq + clampi(e1, pa[q], pa[q + 1]) - pa[q];
    if (*b1 < *b0) *b1 = *b0;
  };
  // ---- buffers: mapped into every process (peer-memory exchange) or private (NCCL exchange)
  PeerState &P = peer_state();
  bool p2p = false;
  if (g_opt_wavelet_p2p) {
    std::vector<size_t> needA((size_t)nr), needB((size_t)nr), needS((size_t)nr);
    for (int r = 0; r < nr; ++r) {
      const int64_t nkr = ka[r + 1] - ka[r];
      needA[r] = (size_t)(nkr * plane);
      needB[r] = (size_t)(n3 * (pa[r + 1] - pa[r]));
      needS[r] = (size_t)std::max<int64_t>(nkr * plane, off[r + 1] - off[r]);
    }
    const int rc = peer_ensure(P, needA, needB, needS, me, nr, st);
    if (rc < 0) return rc;
    p2p = (rc == 0);
  }
  DistBufs &D = dist_bufs();
  if (!p2p) {
    TFX_TRY(D.A.alloc((size_t)(nk * plane)));
    TFX_TRY(D.B.alloc((size_t)(n3 * W)));
    TFX_TRY(D.stage.alloc((size_t)std::max<int64_t>(nk * plane, off[me + 1] - off[me])));
  }
  double *const bufA = p2p ? P.A.p : D.A.p, *const bufB = p2p ? P.B.p : D.B.p, *const bufS = p2p ? P.S.p : D.stage.p;
  std::vector<int64_t> soff((size_t)nr, 0), scnt((size_t)nr, 0), roff((size_t)nr, 0), rcnt((size_t)nr, 0);

  // ---- slabs -> layout A
  TFX_CUDA(cudaMemcpyAsync(bufA, d_slab + lead, (size_t)own * 8, cudaMemcpyDeviceToDevice, st));
  if (p2p) {
    if (me > 0 && lead > 0) {
      const int64_t own_prev = off[me] - ka[me - 1] * plane;          // cells of rank me-1's planes it holds itself
      TFX_CUDA(cudaMemcpyAsync(P.pA.p[me - 1] + own_prev, d_slab, (size_t)lead * 8, cudaMemcpyDefault, st));
    }
    TFX_TRY(peer_barrier(P, st));
  } else {
    if (me > 0) { soff[me - 1] = 0; scnt[me - 1] = lead; }
    if (me + 1 < nr) { roff[me + 1] = own; rcnt[me + 1] = trail; }
    TFX_TRY(comm_exchange_f64(d_slab, soff.data(), scnt.data(), bufA, roff.data(), rcnt.data(), st));
  }

  // ---- axes 1 and 2 on my planes
  TFX_TRY(wavelet_axes12_device(bufA, nx, ny, nk, wavelet_type, forward, st));

  // ---- A -> B: to rank q the columns [pa[q], pa[q+1]) of my planes; what arrives from rank r are the rows
  // [ka[r], ka[r+1]) of B, already in place
  if (p2p) {
    ScatterArgs sa;
    sa.dst = P.pB;
    int64_t wmax = 0;
    for (int q = 0; q <= nr; ++q) sa.pa[q] = pa[q];
    for (int q = 0; q < nr; ++q) wmax = std::max(wmax, pa[q + 1] - pa[q]);
    sa.plane = plane; sa.ka_me = ka[me]; sa.nk = (int)nk; sa.nr = nr;
    const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((wmax + 1023) / 1024, 64)),
                    (unsigned)std::min<int64_t>(nk, 1024), (unsigned)nr);
    k_scatter_columns<<<grid, 256, 0, st>>>(bufA, sa);
    ctx().launches++;
    TFX_CUDA(cudaGetLastError());
    TFX_TRY(peer_barrier(P, st));
  } else {
    int64_t pos = 0;
    for (int q = 0; q < nr; ++q) {
      const int64_t Wq = pa[q + 1] - pa[q];
      if (q == me) {
        TFX_CUDA(cudaMemcpy2DAsync(bufB + ka[me] * W, (size_t)W * 8, bufA + pa[me], (size_t)plane * 8, (size_t)W * 8,
                                   (size_t)nk, cudaMemcpyDeviceToDevice, st));
        scnt[q] = rcnt[q] = 0;
        continue;
      }
      TFX_CUDA(cudaMemcpy2DAsync(bufS + pos, (size_t)Wq * 8, bufA + pa[q], (size_t)plane * 8, (size_t)Wq * 8,
                                 (size_t)nk, cudaMemcpyDeviceToDevice, st));
      soff[q] = pos; scnt[q] = nk * Wq;
      pos += nk * Wq;
      roff[q] = ka[q] * W; rcnt[q] = (ka[q + 1] - ka[q]) * W;
    }
    TFX_TRY(comm_exchange_f64(bufS, soff.data(), scnt.data(), bufB, roff.data(), rcnt.data(), st));
  }

  // ---- axis 3 on my columns
  TFX_TRY(wavelet_axis_device(bufB, nz, W, 1, wavelet_type, forward, st));

  // ---- B -> slabs: the cells of slab r inside my columns are ONE contiguous range of B (suffix of the first row, whole
  // rows, prefix of the last row); the receiver unpacks rows of width W_q at stride `plane`.
  {
    int64_t pos = 0;
    for (int q = 0; q < nr; ++q) {
      int64_t b0, b1;
      range_in(me, q, &b0, &b1);                                       // what I send to rank q
      soff[q] = b0; scnt[q] = (q == me) ? 0 : b1 - b0;
      range_in(q, me, &b0, &b1);                                       // what rank q sends to me
      roff[q] = pos; rcnt[q] = (q == me) ? 0 : b1 - b0;
      if (q != me) pos += b1 - b0;
    }
    if (p2p) {
      RangeArgs ra;
      ra.dst = P.pS; ra.nr = nr; ra.me = me;
      int64_t cmax = 0;
      for (int q = 0; q < nr; ++q) {
        ra.b0[q] = soff[q]; ra.cnt[q] = scnt[q];
        // where rank q expects my piece: behind the pieces of the ranks before me (rank q itself sends nothing)
        int64_t o = 0;
        for (int sidx = 0; sidx < me; ++sidx) {
          if (sidx == q) continue;
          int64_t b0, b1;
          range_in(sidx, q, &b0, &b1);
          o += b1 - b0;
        }
        ra.roff[q] = o;
        cmax = std::max(cmax, scnt[q]);
      }
      const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((cmax + 1023) / 1024, 512)), (unsigned)nr);
      k_scatter_ranges<<<grid, 256, 0, st>>>(bufB, ra);
      ctx().launches++;
      TFX_CUDA(cudaGetLastError());
      TFX_TRY(peer_barrier(P, st));
    } else {
      TFX_TRY(comm_exchange_f64(bufB, soff.data(), scnt.data(), bufS, roff.data(), rcnt.data(), st));
    }
    // unpack (the own block straight from B): one launch for all the sources
    const int64_t k0 = off[me] / plane, s0 = off[me] - k0 * plane;
    const int64_t k1 = (off[me + 1] - 1) / plane, e1 = off[me + 1] - k1 * plane;
    UnpackArgs ua;
    ua.k0 = k0; ua.plane = plane; ua.off_me = off[me]; ua.nr = nr;
    int64_t cmax = 0;
    for (int q = 0; q < nr; ++q) {
      int64_t b0, b1;
      range_in(q, me, &b0, &b1);
      ua.src[q] = (q == me) ? bufB + b0 : bufS + roff[q];
      ua.cnt[q] = b1 - b0;
      ua.paq[q] = pa[q]; ua.Wq[q] = pa[q + 1] - pa[q];
      ua.c0[q] = clampi(s0, pa[q], pa[q + 1]);
      // elements of the first row: up to the end of the column range, or up to e1 when the slab ends in that row
      ua.w0[q] = ((k1 == k0) ? clampi(e1, pa[q], pa[q + 1]) : pa[q + 1]) - ua.c0[q];
      cmax = std::max<int64_t>(cmax, ua.cnt[q]);
    }
    if (cmax > 0) {
      const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((cmax + 1023) / 1024, 512)), (unsigned)nr);
      k_unpack_slab<<<grid, 256, 0, st>>>(d_slab, ua);
      ctx().launches++;
      TFX_CUDA(cudaGetLastError());
    }
    // (peer-memory path: B and the staging buffer are overwritten by the peers' next exchanges only after the barrier
    // that follows the NEXT transform's first exchange, which every rank enters after this unpack)
  }
  *done = 1;
  g_wavelet_last_dist = p2p ? 2 : 1;
  return 0;
}

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
  {
    int done = 0;
    TFX_TRY(wavelet_slab_dist(d_slab, offsets, nx, ny, nz, wavelet_type, forward, st, &done));
    if (done) return 0;
  }
  g_wavelet_last_dist = 0;
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

extern "C" int tfx_wavelet_last_distributed(void) { return tfx::g_wavelet_last_dist; }

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
