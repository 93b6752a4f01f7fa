// t16.cu -- "T16": tiled sparse products with 16-bit in-tile indices (the compressed sensitivity kernel).
//
// Replaces the loops of src/inversion/sparse_matrix.f90 -- add_mult_vector (:313-329) through the F
// layout, add_trans_mult_vector (:388-405) through the T layout -- for the big wavelet-compressed
// matrix_sensit (sensitivity_gravmag.F90:759-856: uniform ~nel_compressed entries per row, ascending
// columns).
//
// B200 design. A product y = A x is a gather from x. x is cut into tiles of <= 16384 elements: the
// tile is staged in shared memory (<= 128 KB of the 227 KB) and every matrix entry addresses it with a
// 16-bit key, so an entry costs 6 bytes of HBM traffic (f32 value + u16 key) instead of the
// reference's 8 (f32 + int32), all random accesses hit shared memory, and HBM sees two pure streams.
// Entries are grouped by (tile, output element) into contiguous segments padded to multiples of 4 entries, with a
// dense int64 pointer table per tile.
//   T layout: x = u (data rows: one tile up to 16384 rows, a few tiles beyond), outputs = columns   -> S^T u
//   F layout: x = v (column tiles, thousands),                                  outputs = data rows -> S v
// DIRECT: one tile; every output belongs to exactly one segment -> y is written once.
// TILES : a CTA parks on a tile (gathered slice in shared memory), its warps draw blocks of 32 outputs from the tile's
//         counter and write partial[tile][output]; a second kernel adds the partials in tile order.
// No atomics on the data path; the summation order is fixed by the layout -> run-to-run deterministic.
// A warp handles 32 consecutive outputs: one coalesced pointer load, then
//   * segments longer than 256 entries are streamed back to back by the whole warp through a per-warp cp.async ring
//     (t16_long_async: 6-8 sixteen-byte packets per lane in flight, across segment ends); the builder deals their
//     entries over the 16 shared-memory bank classes, so the gathers of a half-warp are conflict-free;
//   * runs of short segments are streamed as one contiguous range and every lane adds up its own segment (no
//     reduction at all).
// ncu (round 1, profiles/r1_t16_*_v3_ncu_summary.csv): 2.24 ms per product on the 2.1e9-entry bench matrix, 5.7 TB/s of
// DRAM reads (0.88 of the measured copy peak).
#include "common.cuh"
#include "kernels.h"

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

#include <algorithm>
#include <vector>

namespace tfx {

int g_opt_t16_min_nnz = 1 << 22;   // matrices with fewer entries stay on the generic CSR kernels
int g_opt_t16_tile = 0;            // 0: automatic; otherwise forced tile size (power of two <= 16384), tests
// Long segments through the cp.async ring when it fits next to the tile: bit 0 = TILES, bit 1 = DIRECT. Measured on
// the bench matrix (2.1e9 nnz): forward 3.10 -> 2.72 ms with the ring, -> 2.28 ms with the ring AND the bank-dealt
// segment order of the builder; transposed 2.45 (registers) -> 2.56 (ring alone) -> 2.28 ms (ring + bank dealing).
int g_opt_t16_async = 3;
int g_opt_t16_bank_deal = 1;       // 1: long segments are dealt over the shared-memory banks by the builder
// 1: long segments through the TMA (cp.async.bulk + mbarrier) ring, 0 (default): per-lane cp.async ring. Measured on B200
// (gpurun_out/r2_t16_ab.txt, 256x256x64 / 10 000 stations, 2.1e9 nnz): forward 2.31 ms with the cp.async ring, 3.45 ms
// with the TMA ring; transposed 2.36 / 2.40 ms. A round is 1.5 KB -- two bulk copies of 1 KB + 0.5 KB issued by one lane
// and a 32-lane mbarrier spin per round cost more than the 128 LDGSTS they replace; TMA pays from several KB per copy
// (the dense sweep's 40 KB columns), which this per-warp, per-segment streaming cannot offer without giving up the
// fixed summation order. Kept as an option for the record and for the tests (bit-identical results).
int g_opt_t16_tma = 0;
int g_opt_t16_blk = 0;            // 0: automatic (32 / 16 / 8 outputs per warp task by the number of outputs)
int g_opt_t16_long_seg = 256;     // segments longer than this take the whole-warp path (<= 256)
int g_opt_t16_direct_max = 16384;   // gathered ranges up to this many elements use one DIRECT tile; longer ones TILES

static const int kT16Threads = 768;      // DIRECT: one CTA per SM
static const int kT16TilesThreads = 384; // TILES: two CTAs per SM (barrier / tile-load waits of one overlap the other)
static const int kT16TilesMaxTile = 8192;
static const int kT16MaxTile = 16384;
static const int kLongSeg = 256;    // longer segments are summed by the whole warp, one at a time
static const int kFlatMax = 1024;   // entries per flat run (512 packet products = 4 KB of shared memory per warp)
static const int kDirectChunk = 1;  // DIRECT: blocks of 32 outputs a warp draws at a time

// Shared-memory placement of gathered element i of a tile. Wavelet coefficients of level l sit at indices
// = 2^(l-1) mod 2^l (interleaved lifting layout), so the gathers of one segment hit power-of-two strides:
// folding the higher index nibbles into the low one spreads them over the 16 eight-byte bank pairs.
// The keys stored in the matrix are already swizzled (builder), only the tile load pays for it.
__host__ __device__ __forceinline__ uint32_t t16_swz(uint32_t i) { return i ^ (((i >> 4) ^ (i >> 8) ^ (i >> 12)) & 15u); }

// Does the kernel of this layout stream its long segments through the cp.async ring? (decided identically by the
// builder, which lays the segments out for the ring's 4-entry packets, and by the launcher)
static bool t16_uses_ring(T16Mode mode, int tile);

struct T16Args {
  const float *val;
  const uint16_t *key;
  const int64_t *ptr;      // [ntiles * nseg + 1]
  const double *x;         // gathered vector, already shifted: element g of the layout is x[g]
  double *y;               // DIRECT: output vector (offset applied: y[o] is output o of the layout)
  double *partial;         // TILES: [ntiles][nseg], every element written exactly once per product
  int *counter;            // dynamic work distribution (results do not depend on it: write-once outputs)
  int32_t nseg, tile, ntiles, nin;   // nin: number of valid gathered elements (in0-relative)
  int32_t nsplit;          // TILES: work items per tile
  int32_t t0;              // DIRECT: the tile to process
  int accumulate;          // DIRECT: y += instead of y =
  int async_ring;          // long segments go through the cp.async ring (the warp strip holds kAsyncRingBytes)
  int wstrip;              // doubles per warp strip (flat-run products / cp.async ring)
  int blk;                 // outputs a warp takes at a time (32; 16 or 8 when a tile has few outputs, so that every warp
                           // of the CTA parked on it finds work -- row blocks of a few hundred stations)
  int long_seg;            // segments longer than this are summed by the whole warp (<= kLongSeg)
  const int *done;
};

// Lane-partial sum over the rest of a long segment: entries [k0 + 2*lane + 64*i, end), 8 packets in flight.
__device__ __forceinline__ double t16_long_partial(const float *__restrict__ val, const uint16_t *__restrict__ key,
                                                   const double *xs, int k, int end) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0;
  for (; k + 448 < end; k += 512) {
    float2 v[8];
    uint32_t kk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = __ldg((const float2 *)(val + k + 64 * i));
      kk[i] = __ldg((const uint32_t *)(key + k + 64 * i));
    }
    a0 = fma((double)v[0].x, xs[kk[0] & 0xffffu], a0); a0 = fma((double)v[0].y, xs[kk[0] >> 16], a0);
    a1 = fma((double)v[1].x, xs[kk[1] & 0xffffu], a1); a1 = fma((double)v[1].y, xs[kk[1] >> 16], a1);
    a2 = fma((double)v[2].x, xs[kk[2] & 0xffffu], a2); a2 = fma((double)v[2].y, xs[kk[2] >> 16], a2);
    a3 = fma((double)v[3].x, xs[kk[3] & 0xffffu], a3); a3 = fma((double)v[3].y, xs[kk[3] >> 16], a3);
    a4 = fma((double)v[4].x, xs[kk[4] & 0xffffu], a4); a4 = fma((double)v[4].y, xs[kk[4] >> 16], a4);
    a5 = fma((double)v[5].x, xs[kk[5] & 0xffffu], a5); a5 = fma((double)v[5].y, xs[kk[5] >> 16], a5);
    a6 = fma((double)v[6].x, xs[kk[6] & 0xffffu], a6); a6 = fma((double)v[6].y, xs[kk[6] >> 16], a6);
    a7 = fma((double)v[7].x, xs[kk[7] & 0xffffu], a7); a7 = fma((double)v[7].y, xs[kk[7] >> 16], a7);
  }
  if (k < end) {   // tail: up to 8 predicated packets, all in flight together
    float2 v[8];
    uint32_t kk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool in = k + 64 * i < end;
      v[i] = in ? __ldg((const float2 *)(val + k + 64 * i)) : make_float2(0.f, 0.f);
      kk[i] = in ? __ldg((const uint32_t *)(key + k + 64 * i)) : 0u;
    }
    a0 = fma((double)v[0].x, xs[kk[0] & 0xffffu], a0); a0 = fma((double)v[0].y, xs[kk[0] >> 16], a0);
    a1 = fma((double)v[1].x, xs[kk[1] & 0xffffu], a1); a1 = fma((double)v[1].y, xs[kk[1] >> 16], a1);
    a2 = fma((double)v[2].x, xs[kk[2] & 0xffffu], a2); a2 = fma((double)v[2].y, xs[kk[2] >> 16], a2);
    a3 = fma((double)v[3].x, xs[kk[3] & 0xffffu], a3); a3 = fma((double)v[3].y, xs[kk[3] >> 16], a3);
    a4 = fma((double)v[4].x, xs[kk[4] & 0xffffu], a4); a4 = fma((double)v[4].y, xs[kk[4] >> 16], a4);
    a5 = fma((double)v[5].x, xs[kk[5] & 0xffffu], a5); a5 = fma((double)v[5].y, xs[kk[5] >> 16], a5);
    a6 = fma((double)v[6].x, xs[kk[6] & 0xffffu], a6); a6 = fma((double)v[6].y, xs[kk[6] >> 16], a6);
    a7 = fma((double)v[7].x, xs[kk[7] & 0xffffu], a7); a7 = fma((double)v[7].y, xs[kk[7] >> 16], a7);
  }
  return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ---- asynchronous long path -------------------------------------------------------------------------------------
// The long segments of a block of 32 outputs are streamed back to back through a per-warp shared-memory ring filled
// with cp.async (LDGSTS): the bytes in flight sit in shared memory instead of registers, so a lane keeps
// (kAsyncW - 1) .. kAsyncW rounds = 6 .. 8 sixteen-byte value packets (+ their keys) in flight -- twice what the
// register-staged path can afford -- and the stream does not drain at segment ends: the producer cursor runs kAsyncW
// rounds ahead of the consumer cursor, across segments. Every lane reads back only what it copied itself
// (cp.async.wait_group is per thread), so no warp- or block-level synchronisation is involved.
// One round = 256 consecutive entries of a segment = kAsyncR packets of 4 entries per lane; segments are padded to
// multiples of 4 entries by the builder (16-byte aligned value packets, 8-byte aligned key packets).
static const int kAsyncW = 4;                       // rounds in the ring
static const int kAsyncR = 2;                       // packets (of 4 entries) per lane and round
static const int kAsyncRingBytes = kAsyncW * kAsyncR * 32 * 24;   // 6 KB per warp

__device__ __forceinline__ uint32_t t16_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t16_cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void t16_cp_async8(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void t16_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void t16_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ double t16_long_async(const float *__restrict__ val, const uint16_t *__restrict__ key,
                                                 const double *xs, unsigned char *ring, unsigned longmask, int rel,
                                                 int len, int lane, double result) {
  const uint32_t rv = t16_smem_u32(ring) + lane * 16;                                   // value packets: 512 B per slot
  const uint32_t rk = t16_smem_u32(ring) + kAsyncW * kAsyncR * 512 + lane * 8;          // key packets:   256 B per slot
  const float4 *lv = (const float4 *)ring + lane;
  const uint2 *lk = (const uint2 *)(ring + kAsyncW * kAsyncR * 512) + lane;
  // producer cursor (copies) and consumer cursor (sums): segment, first entry of the next round, end of the segment
  unsigned pm = longmask, cm = longmask;
  int pj = __ffs(pm) - 1;
  int pk = __shfl_sync(0xffffffffu, rel, pj);
  int p_end = pk + __shfl_sync(0xffffffffu, len, pj);
  int cj = pj, ck = pk, c_end = p_end;
  auto issue = [&](int slot) {
    if (pm != 0u) {
#pragma unroll
      for (int i = 0; i < kAsyncR; ++i) {
        const int e = pk + 128 * i + 4 * lane;
        if (e < p_end) {
          t16_cp_async16(rv + (slot * kAsyncR + i) * 512, val + e);
          t16_cp_async8(rk + (slot * kAsyncR + i) * 256, key + e);
        }
      }
      pk += 128 * kAsyncR;
      if (pk >= p_end) {
        pm &= pm - 1;
        if (pm != 0u) {
          pj = __ffs(pm) - 1;
          pk = __shfl_sync(0xffffffffu, rel, pj);
          p_end = pk + __shfl_sync(0xffffffffu, len, pj);
        }
      }
    }
    t16_cp_commit();
  };
#pragma unroll
  for (int s = 0; s < kAsyncW; ++s) issue(s);
  int slot = 0;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  while (cm != 0u) {
    t16_cp_wait<kAsyncW - 1>();
#pragma unroll
    for (int i = 0; i < kAsyncR; ++i) {
      const int e = ck + 128 * i + 4 * lane;
      if (e < c_end) {
        const float4 v = lv[(slot * kAsyncR + i) * 32];
        const uint2 k = lk[(slot * kAsyncR + i) * 32];
        a0 = fma((double)v.x, xs[k.x & 0xffffu], a0);
        a1 = fma((double)v.y, xs[k.x >> 16], a1);
        a2 = fma((double)v.z, xs[k.y & 0xffffu], a2);
        a3 = fma((double)v.w, xs[k.y >> 16], a3);
      }
    }
    issue(slot);                                   // refill the slot that was just consumed
    slot = (slot + 1 == kAsyncW) ? 0 : slot + 1;
    ck += 128 * kAsyncR;
    if (ck >= c_end) {                             // end of the consumed segment: reduce, hand the sum to its lane
      const double t = warp_sum((a0 + a1) + (a2 + a3));
      if (lane == cj) result = t;
      a0 = a1 = a2 = a3 = 0.0;
      cm &= cm - 1;
      if (cm != 0u) {
        cj = __ffs(cm) - 1;
        ck = __shfl_sync(0xffffffffu, rel, cj);
        c_end = ck + __shfl_sync(0xffffffffu, len, cj);
      }
    }
  }
  t16_cp_wait<0>();                                // only empty groups are left; the strip is reused by the flat path
  return result;
}

// ---- TMA long path -----------------------------------------------------------------------------------------------
// Same ring, filled by the TMA engine: one elected lane issues ONE cp.async.bulk (UBLKCP) for the 256 values of a round
// and one for their keys, completion is signalled on an mbarrier per ring slot (complete_tx), all lanes wait on the
// slot's phase and read their packets. 2 bulk copies per round replace 128 LDGSTS of the warp; the producer cursor
// still runs kAsyncW rounds ahead across segment ends. Value packets are 16-byte aligned (segments are padded to 4
// entries); key ranges are 8-byte aligned, so the key copy starts at the 16-byte boundary below (<= 8 bytes early) and
// the readers add the offset. Ring strip per warp: [kAsyncW x 1024 B values][kAsyncW x 528 B keys][kAsyncW mbarriers].
static const int kBulkKeySlot = 512 + 16;
static const int kBulkRingBytes = kAsyncW * 1024 + kAsyncW * kBulkKeySlot + kAsyncW * 8 + 32;   // 6272 B per warp

__device__ __forceinline__ void t16_mbar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t16_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t16_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void t16_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// Once per warp and kernel: the slot barriers of the warp's strip.
__device__ __forceinline__ void t16_bulk_ring_init(unsigned char *ring, int lane) {
  if (lane == 0) {
    const uint32_t bars = t16_smem_u32(ring) + kAsyncW * 1024 + kAsyncW * kBulkKeySlot;
#pragma unroll
    for (int s = 0; s < kAsyncW; ++s) t16_mbar_init(bars + 8 * s);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

// `rounds` counts the rounds this warp has pushed through its ring since the kernel started (slot = rounds % kAsyncW,
// phase parity = (rounds / kAsyncW) & 1); every call drains what it issued, so the count is the same for producer and
// consumer between calls.
__device__ __forceinline__ double t16_long_bulk(const float *__restrict__ val, const uint16_t *__restrict__ key,
                                                const double *xs, unsigned char *ring, unsigned longmask, int rel,
                                                int len, int lane, double result, unsigned &rounds) {
  const uint32_t rbase = t16_smem_u32(ring);
  const uint32_t kbase = rbase + kAsyncW * 1024;
  const uint32_t bars = kbase + kAsyncW * kBulkKeySlot;
  // the strip may hold flat-run products written through the generic proxy: order them before the async-proxy writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  unsigned pm = longmask, cm = longmask;
  int pj = __ffs(pm) - 1;
  int pk = __shfl_sync(0xffffffffu, rel, pj);
  int p_end = pk + __shfl_sync(0xffffffffu, len, pj);
  int cj = pj, ck = pk, c_end = p_end;
  unsigned pr = rounds, cr = rounds;   // producer / consumer round counters
  auto issue = [&]() {
    if (pm == 0u) return;
    const int n = min(256, p_end - pk);                 // entries of this round (a multiple of 4)
    const unsigned slot = pr & (kAsyncW - 1);
    if (lane == 0) {
      const uint16_t *ksrc = key + pk;
      const uint32_t kal = (uint32_t)((uintptr_t)ksrc & 15u);           // 0 or 8
      const uint32_t vbytes = (uint32_t)n * 4u, kbytes = (kal + (uint32_t)n * 2u + 15u) & ~15u;
      const uint32_t bar = bars + 8 * slot;
      t16_mbar_expect_tx(bar, vbytes + kbytes);
      t16_bulk_load(rbase + slot * 1024, val + pk, vbytes, bar);
      t16_bulk_load(kbase + slot * kBulkKeySlot, (const unsigned char *)ksrc - kal, kbytes, bar);
    }
    ++pr;
    pk += 256;
    if (pk >= p_end) {
      pm &= pm - 1;
      if (pm != 0u) {
        pj = __ffs(pm) - 1;
        pk = __shfl_sync(0xffffffffu, rel, pj);
        p_end = pk + __shfl_sync(0xffffffffu, len, pj);
      }
    }
  };
#pragma unroll
  for (int s = 0; s < kAsyncW; ++s) issue();
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  while (cm != 0u) {
    const unsigned slot = cr & (kAsyncW - 1);
    t16_mbar_wait(bars + 8 * slot, (cr / kAsyncW) & 1u);
    const unsigned char *vs = ring + slot * 1024 + lane * 16;
    const unsigned char *ks = ring + kAsyncW * 1024 + slot * kBulkKeySlot + (((uintptr_t)(key + ck)) & 15u) + lane * 8;
#pragma unroll
    for (int i = 0; i < kAsyncR; ++i) {
      const int e = ck + 128 * i + 4 * lane;
      if (e < c_end) {
        const float4 v = *(const float4 *)(vs + 512 * i);
        const uint2 k = *(const uint2 *)(ks + 256 * i);
        a0 = fma((double)v.x, xs[k.x & 0xffffu], a0);
        a1 = fma((double)v.y, xs[k.x >> 16], a1);
        a2 = fma((double)v.z, xs[k.y & 0xffffu], a2);
        a3 = fma((double)v.w, xs[k.y >> 16], a3);
      }
    }
    ++cr;
    __syncwarp();                                  // every lane has read the slot: it may be refilled
    issue();
    ck += 256;
    if (ck >= c_end) {                             // end of the consumed segment: reduce, hand the sum to its lane
      const double t = warp_sum((a0 + a1) + (a2 + a3));
      if (lane == cj) result = t;
      a0 = a1 = a2 = a3 = 0.0;
      cm &= cm - 1;
      if (cm != 0u) {
        cj = __ffs(cm) - 1;
        ck = __shfl_sync(0xffffffffu, rel, cj);
        c_end = ck + __shfl_sync(0xffffffffu, len, cj);
      }
    }
  }
  rounds = cr;
  return result;
}

// 32 consecutive outputs [o0, o0 + 32) of tile t: lane j returns the sum of segment o0 + j.
//  * segments longer than kLongSeg entries: one at a time with the whole warp (t16_long_partial);
//  * all others: FLAT -- maximal runs of consecutive segments spanning <= kFlatMax entries are streamed as
//    one contiguous range (every lane loads packets, all loads independent and coalesced), the packet
//    products are parked in the warp's shared-memory strip and every lane then adds up its own segment.
__device__ __forceinline__ double t16_block32(const T16Args &a, const double *xs, double *wbuf, int64_t tbase, int o0,
                                              int lane, unsigned &rounds) {
  const int o = o0 + lane;
  // lanes past the last output of the block (o0 + blk) or of the layout read the end pointer: empty segments, offsets
  // stay monotone
  const int olim = min(o0 + a.blk, a.nseg);
  const int64_t pb = __ldg(a.ptr + tbase + min(o, olim));
  int64_t pe = __shfl_down_sync(0xffffffffu, pb, 1);
  if (lane == 31) pe = __ldg(a.ptr + tbase + min(o + 1, olim));
  const int64_t pb0 = __shfl_sync(0xffffffffu, pb, 0);
  const int len = (int)(pe - pb);                              // a segment never exceeds the tile size
  const int rel = (int)(pb - pb0);                             // 32 segments span < 2^31 entries
  double result = 0.0;
  if (__ballot_sync(0xffffffffu, len > 0) == 0u) return result;
  const float *__restrict__ val = a.val + pb0;                 // pb0 is even: 8-byte aligned packets
  const uint16_t *__restrict__ key = a.key + pb0;
  const bool is_long = len > a.long_seg;
  unsigned longmask = __ballot_sync(0xffffffffu, is_long);
  const unsigned flatmask = ~longmask;
  if (a.async_ring == 2) {
    if (longmask) result = t16_long_bulk(val, key, xs, (unsigned char *)wbuf, longmask, rel, len, lane, result, rounds);
  } else if (a.async_ring) {
    if (longmask) result = t16_long_async(val, key, xs, (unsigned char *)wbuf, longmask, rel, len, lane, result);
  } else {
    while (longmask) {
      const int j = __ffs(longmask) - 1;
      longmask &= longmask - 1;
      const int r = __shfl_sync(0xffffffffu, rel, j);
      const int l = __shfl_sync(0xffffffffu, len, j);
      const double s = warp_sum(t16_long_partial(val, key, xs, r + 2 * lane, r + l));
      if (lane == j) result = s;
    }
  }
  int j0 = 0;
  while (j0 < 32) {
    if (!((flatmask >> j0) & 1u)) { ++j0; continue; }
    const int start = __shfl_sync(0xffffffffu, rel, j0);
    const bool fits = !is_long && lane >= j0 && (rel + len - start) <= kFlatMax;
    const unsigned fm = __ballot_sync(0xffffffffu, fits) >> j0;      // bit 0 is set: a flat segment always fits
    const int cnt = (fm == 0xffffffffu) ? 32 : (__ffs(~fm) - 1);       // consecutive segments in this run
    const int j1 = j0 + cnt;
    const int endrel = __shfl_sync(0xffffffffu, rel + len, j1 - 1);
    const int npk = (endrel - start) >> 1;
    for (int p = lane; p < npk; p += 256) {
      float2 v[8];
      uint32_t kk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool in = p + 32 * i < npk;
        v[i] = in ? __ldg((const float2 *)(val + start + 2 * (p + 32 * i))) : make_float2(0.f, 0.f);
        kk[i] = in ? __ldg((const uint32_t *)(key + start + 2 * (p + 32 * i))) : 0u;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (p + 32 * i < npk)
          wbuf[p + 32 * i] = fma((double)v[i].y, xs[kk[i] >> 16], (double)v[i].x * xs[kk[i] & 0xffffu]);
    }
    __syncwarp();
    if (lane >= j0 && lane < j1) {
      double s = 0.0;
      const int q0 = (rel - start) >> 1, q1 = (rel + len - start) >> 1;
      for (int q = q0; q < q1; ++q) s += wbuf[q];
      result = s;
    }
    __syncwarp();
    j0 = j1;
  }
  return result;
}

__device__ __forceinline__ void t16_load_tile(const T16Args &a, double *xs, int t) {
  const int base = t * a.tile;
  const int padded = (a.tile + 15) & ~15;
  for (int i = threadIdx.x; i < padded; i += blockDim.x) {
    const int g = base + i;
    xs[t16_swz((uint32_t)i)] = (i < a.tile && g < a.nin) ? a.x[g] : 0.0;
  }
}

// DIRECT: one tile; every warp draws kDirectChunk blocks of 32 outputs at a time from a global counter
// (the nonzeros are concentrated in the shallow cells, i.e. in a few column ranges: static splits starve).
__global__ void __launch_bounds__(kT16Threads, 1) t16_direct_kernel(T16Args a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(16) double xs[];
  t16_load_tile(a, xs, a.t0);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *wbuf = xs + ((a.tile + 15) & ~15) + wid * a.wstrip;
  unsigned rounds = 0;
  if (a.async_ring == 2) t16_bulk_ring_init((unsigned char *)wbuf, lane);
  const int64_t tbase = (int64_t)a.t0 * a.nseg;
  const int nblk = (a.nseg + a.blk - 1) / a.blk;
  __syncthreads();
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(a.counter, kDirectChunk);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= nblk) break;
    const int b_hi = min(nblk, base + kDirectChunk);
    for (int blk = base; blk < b_hi; ++blk) {
      const double r = t16_block32(a, xs, wbuf, tbase, blk * a.blk, lane, rounds);
      const int o = blk * a.blk + lane;
      if (lane < a.blk && o < a.nseg) a.y[o] = a.accumulate ? (a.y[o] + r) : r;
    }
  }
}

// TILES: a CTA parks on a tile (gathered slice in shared memory) and its warps draw blocks of 32 outputs from
// that tile's counter; CTAs start on evenly spread tiles and walk forward, skipping exhausted tiles, so heavy
// (dense, shallow-depth) tiles are finished by several CTAs together. partial[tile][output] is written exactly
// once per product, hence the result does not depend on who computed what.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) t16_tiles_kernel(T16Args a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(16) double xs[];
  __shared__ int s_next;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *wbuf = xs + ((a.tile + 15) & ~15) + wid * a.wstrip;
  unsigned rounds = 0;
  if (a.async_ring == 2) t16_bulk_ring_init((unsigned char *)wbuf, lane);
  const int nblk = (a.nseg + a.blk - 1) / a.blk;
  const int start = (int)((int64_t)blockIdx.x * a.ntiles / gridDim.x);
  int i = 0;   // tiles visited so far (relative to start)
  for (;;) {
    __syncthreads();   // everybody is done with xs and s_next
    if (wid == 0) {    // find the next tile that still has undrawn blocks, 32 candidates at a time
      int found = a.ntiles;
      for (int base = i; base < a.ntiles && found == a.ntiles; base += 32) {
        const int c = base + lane;
        bool open = false;
        if (c < a.ntiles) {
          int t = start + c;
          if (t >= a.ntiles) t -= a.ntiles;
          open = *((volatile int *)(a.counter + t)) < nblk;
        }
        const unsigned m = __ballot_sync(0xffffffffu, open);
        if (m) found = base + __ffs(m) - 1;
      }
      if (lane == 0) s_next = found;
    }
    __syncthreads();
    i = s_next;
    if (i >= a.ntiles) break;
    int t = start + i;
    if (t >= a.ntiles) t -= a.ntiles;
    t16_load_tile(a, xs, t);
    __syncthreads();
    const int64_t tbase = (int64_t)t * a.nseg;
    for (;;) {
      int blk = 0;
      if (lane == 0) blk = atomicAdd(a.counter + t, 1);
      blk = __shfl_sync(0xffffffffu, blk, 0);
      if (blk >= nblk) break;
      const double r = t16_block32(a, xs, wbuf, tbase, blk * a.blk, lane, rounds);
      const int o = blk * a.blk + lane;
      if (lane < a.blk && o < a.nseg) a.partial[tbase + o] = r;
    }
    ++i;
  }
}

// y[o] (+)= sum_t partial[t][o], t ascending.
__global__ void __launch_bounds__(256) t16_reduce_kernel(const double *__restrict__ partial, int ntiles, int nseg,
                                                         double *__restrict__ y, int accumulate, const int *done) {
  if (done && *done) return;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < nseg; o += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int t = 0; t < ntiles; ++t) s += partial[(int64_t)t * nseg + o];
    y[o] = accumulate ? (y[o] + s) : s;
  }
}

__global__ void __launch_bounds__(256) t16_zero_kernel(double *y, int64_t n, const int *done) {
  if (done && *done) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = 0.0;
}

static bool t16_uses_ring(T16Mode mode, int tile) {
  const size_t tile_bytes = (size_t)((tile + 15) & ~15) * sizeof(double);
  const size_t smem_max = 227 * 1024 - 64;
  const int bit = (mode == T16_TILES) ? 1 : 2;
  return (g_opt_t16_async & bit) && tile_bytes + (size_t)(kT16Threads / 32) * kBulkRingBytes <= smem_max;
}

// ---------------------------------------------------------------------------------------------
// Product
// ---------------------------------------------------------------------------------------------
int t16_spmv(T16Matrix &m, const double *d_x, double *d_y, bool accumulate, int32_t xshift, const int *d_done,
             cudaStream_t st) {
  Context &c = ctx();
  if (!m.valid) return fail(-40, "t16: layout was not built");
  if (!accumulate) {   // outputs outside the covered range
    const int64_t n_lo = m.out0, n_hi = (int64_t)m.nout_total - (m.out0 + m.nseg);
    if (n_lo > 0) {
      t16_zero_kernel<<<(int)std::min<int64_t>((n_lo + 255) / 256, c.num_sms * 8), 256, 0, st>>>(d_y, n_lo, d_done);
      c.launches++;
    }
    if (n_hi > 0) {
      t16_zero_kernel<<<(int)std::min<int64_t>((n_hi + 255) / 256, c.num_sms * 8), 256, 0, st>>>(d_y + m.out0 + m.nseg, n_hi, d_done);
      c.launches++;
    }
  }
  if (m.nseg == 0) return 0;
  T16Args a;
  a.val = m.val.p; a.key = m.key.p; a.ptr = m.ptr.p;
  a.x = d_x + m.in0 - xshift;
  a.y = d_y + m.out0;
  a.partial = m.partial.p; a.counter = m.counter.p;
  a.nseg = m.nseg; a.tile = m.tile; a.ntiles = m.ntiles; a.nin = m.nin; a.nsplit = m.nsplit;
  a.t0 = 0; a.accumulate = accumulate ? 1 : 0; a.done = d_done;
  // outputs per warp task: fewer when the layout has few outputs (every tile visit must feed all 24 warps of a CTA)
  a.blk = (m.nseg >= 32 * 48) ? 32 : (m.nseg >= 16 * 48 ? 16 : 8);
  if (g_opt_t16_blk > 0) a.blk = g_opt_t16_blk;
  a.long_seg = std::max(8, std::min(g_opt_t16_long_seg, kLongSeg));
  const int nblk = (m.nseg + a.blk - 1) / a.blk;
  const size_t tile_bytes = (size_t)((m.tile + 15) & ~15) * sizeof(double);
  const size_t smem_max = 227 * 1024 - 64;
  // option t16_tma (default 1): the ring is filled by cp.async.bulk (TMA) instead of per-lane cp.async
  const size_t strip_flat = (size_t)(kFlatMax / 2) * sizeof(double), strip_async = g_opt_t16_tma ? (size_t)kBulkRingBytes : (size_t)kAsyncRingBytes;
  if (m.mode == T16_DIRECT) {
    const int warps = kT16Threads / 32;
    const bool use_async = t16_uses_ring(T16_DIRECT, m.tile);
    const size_t strip = use_async ? strip_async : strip_flat;
    const size_t smem = tile_bytes + warps * strip;
    a.async_ring = use_async ? (g_opt_t16_tma ? 2 : 1) : 0;
    a.wstrip = (int)(strip / sizeof(double));
    static bool attr = false;
    if (!attr) {
      TFX_CUDA(cudaFuncSetAttribute(t16_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
      attr = true;
    }
    const int nitems = (nblk + kDirectChunk - 1) / kDirectChunk;
    const int grid = std::max(1, std::min(c.num_sms, (nitems + kT16Threads / 32 - 1) / (kT16Threads / 32)));
    for (int t = 0; t < m.ntiles; ++t) {
      a.t0 = t;
      a.accumulate = (accumulate || t > 0) ? 1 : 0;
      TFX_CUDA(cudaMemsetAsync(m.counter.p, 0, sizeof(int), st));
      t16_direct_kernel<<<grid, kT16Threads, smem, st>>>(a);
      c.launches++;
    }
  } else {
    // register-staged long path: two CTAs of 384 threads per SM (barrier / tile-load waits of one overlap the other);
    // cp.async ring: one CTA of 768 threads (same 24 warps per SM, the ring needs the second CTA's shared memory)
    const bool use_async = t16_uses_ring(T16_TILES, m.tile);
    TFX_CUDA(cudaMemsetAsync(m.counter.p, 0, sizeof(int) * (size_t)m.ntiles, st));
    if (use_async) {
      const size_t smem = tile_bytes + (kT16Threads / 32) * strip_async;
      a.async_ring = g_opt_t16_tma ? 2 : 1;
      a.wstrip = (int)(strip_async / sizeof(double));
      static bool attr = false;
      if (!attr) {
        TFX_CUDA(cudaFuncSetAttribute(t16_tiles_kernel<kT16Threads, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        attr = true;
      }
      const int grid = c.num_sms;   // several CTAs may park on the same tile (they share its block counter)
      t16_tiles_kernel<kT16Threads, 1><<<grid, kT16Threads, smem, st>>>(a);
    } else {
      const size_t smem = tile_bytes + (kT16TilesThreads / 32) * strip_flat;
      a.async_ring = 0;
      a.wstrip = (int)(strip_flat / sizeof(double));
      static bool attr = false;
      if (!attr) {
        TFX_CUDA(cudaFuncSetAttribute(t16_tiles_kernel<kT16TilesThreads, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (kT16TilesMaxTile + (kT16TilesThreads / 32) * (kFlatMax / 2)) * 8));
        attr = true;
      }
      const int grid = 2 * c.num_sms;
      t16_tiles_kernel<kT16TilesThreads, 2><<<grid, kT16TilesThreads, smem, st>>>(a);
    }
    c.launches++;
    const int blocks = std::max(1, std::min((m.nseg + 255) / 256, c.num_sms * 8));
    t16_reduce_kernel<<<blocks, 256, 0, st>>>(m.partial.p, m.ntiles, m.nseg, a.y, accumulate ? 1 : 0, d_done);
    c.launches++;
  }
  TFX_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Builder: from a compressed-segment matrix whose segments hold strictly ascending indices.
// ---------------------------------------------------------------------------------------------
namespace {

// flags[0] = 1 when some segment is not strictly ascending; mm[0] = min idx, mm[1] = max idx.
__global__ void __launch_bounds__(256) t16_scan_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                       int nstored, int *flags, int *mm) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  int lo = 0x7fffffff, hi = -1, bad = 0;
  for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < nstored; s += gridDim.x * wpb) {
    const int64_t b = ptr[s], e = ptr[s + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
      const int v = idx[k];
      if (k > b && idx[k - 1] >= v) bad = 1;
      lo = min(lo, v);
      hi = max(hi, v);
    }
  }
  if (bad) atomicExch(&flags[0], 1);
  if (hi >= 0) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

__global__ void __launch_bounds__(256) t16_segof_kernel(const int32_t *__restrict__ segmap, int nstored, int out0,
                                                        int32_t *__restrict__ segof) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nstored; s += gridDim.x * blockDim.x)
    segof[segmap[s] - out0] = s;
}

__device__ __forceinline__ int64_t t16_lower_bound(const int32_t *__restrict__ idx, int64_t b, int64_t e, int target) {
  while (b < e) {
    const int64_t mid = (b + e) >> 1;
    if (idx[mid] < target) b = mid + 1;
    else e = mid;
  }
  return b;
}

// cnt[t * nseg + o] = number of entries of output o whose index lies in tile t, padded to a multiple of 4 (16-byte
// value packets / 8-byte key packets of the cp.async path; the 2-entry packets of the other paths stay aligned too).
__global__ void __launch_bounds__(256) t16_count_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                        const int32_t *__restrict__ segof, int nseg, int ntiles,
                                                        int tile, int in0, int64_t *__restrict__ cnt) {
  const int64_t total = (int64_t)nseg * ntiles;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / nseg), o = (int)(i % nseg);
    const int s = segof[o];
    int64_t n = 0;
    if (s >= 0) {
      const int64_t b = ptr[s], e = ptr[s + 1];
      const int64_t lo = (ntiles == 1) ? b : t16_lower_bound(idx, b, e, in0 + t * tile);
      const int64_t hi = (t + 1 == ntiles) ? e : t16_lower_bound(idx, lo, e, in0 + (t + 1) * tile);
      n = hi - lo;
    }
    cnt[i] = (n + 3) & ~(int64_t)3;
  }
}

// One warp per (tile, output): copies the run into its padded slot.
//
// Long runs (the ones the whole-warp paths stream) are re-ordered so that the shared-memory gathers of a half-warp
// hit 16 different 8-byte banks: the entries are dealt round-robin over the 16 bank classes of their (swizzled) keys,
// and the resulting sequence is laid out so that the 16 entries one gather instruction reads for a half-warp are 16
// consecutive entries of that sequence. `stride` = entries per lane packet of the kernel that will read the layout
// (2: register-staged path, 4: cp.async ring). The order of summation inside a segment changes with it -- it stays
// fixed by the layout, i.e. deterministic. ncu before: 6 shared-memory wavefronts per 32-lane gather (ideal 2), the
// gathers alone filled 60 % of the shared-memory pipe.
__global__ void __launch_bounds__(256) t16_fill_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                       const float *__restrict__ sval, const int32_t *__restrict__ segof,
                                                       int nseg, int ntiles, int tile, int in0,
                                                       const int64_t *__restrict__ tptr, float *__restrict__ val,
                                                       uint16_t *__restrict__ key, int stride, int reorder) {
  __shared__ int s_cnt[8][16], s_run[8][16];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int64_t total = (int64_t)nseg * ntiles;
  const int blk = 16 * stride;   // entries one half-warp reads per packet row
  for (int64_t i = blockIdx.x * (int64_t)wpb + w; i < total; i += (int64_t)gridDim.x * wpb) {
    const int64_t dst = tptr[i];
    const int64_t npad = tptr[i + 1] - dst;
    if (npad == 0) continue;
    const int t = (int)(i / nseg), o = (int)(i % nseg);
    const int s = segof[o];
    const int64_t b = ptr[s], e = ptr[s + 1];
    const int64_t lo = (ntiles == 1) ? b : t16_lower_bound(idx, b, e, in0 + t * tile);
    const int64_t hi = (t + 1 == ntiles) ? e : t16_lower_bound(idx, lo, e, in0 + (t + 1) * tile);
    const int base = in0 + t * tile;
    if (!reorder || npad <= kLongSeg) {
      for (int64_t k = lo + lane; k < hi; k += 32) {
        val[dst + (k - lo)] = sval[k];
        key[dst + (k - lo)] = (uint16_t)t16_swz((uint32_t)(idx[k] - base));
      }
      continue;   // padding slots: value 0 contributes exactly 0; key 0 is always a valid tile element
    }
    const int n = (int)(hi - lo);
    const int nfull = n / blk * blk;   // only whole packet rows are permuted, the tail keeps the dealing order
    if (lane < 16) { s_cnt[w][lane] = 0; s_run[w][lane] = 0; }
    __syncwarp();
    for (int k = lane; k < n; k += 32) atomicAdd(&s_cnt[w][t16_swz((uint32_t)(idx[lo + k] - base)) & 15u], 1);
    __syncwarp();
    for (int k0 = 0; k0 < n; k0 += 32) {
      const int k = k0 + lane;
      const bool act = k < n;
      const unsigned amask = __ballot_sync(0xffffffffu, act);
      if (act) {
        const uint32_t kswz = t16_swz((uint32_t)(idx[lo + k] - base));
        const int c = (int)(kswz & 15u);
        const unsigned peers = __match_any_sync(amask, c);
        const int g = s_run[w][c] + __popc(peers & ((1u << lane) - 1u));   // index of this entry inside its class
        __syncwarp(amask);
        if (lane == __ffs(peers) - 1) s_run[w][c] += __popc(peers);
        // place in the dealing sequence: g full rounds over the classes that still have entries, then the classes below
        int q = 0;
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          const int m = s_cnt[w][cc];
          q += min(m, g) + ((cc < c && m > g) ? 1 : 0);
        }
        const int pos = (q < nfull) ? (q / blk * blk + stride * (q & 15) + ((q >> 4) % stride)) : q;
        val[dst + pos] = sval[lo + k];
        key[dst + pos] = (uint16_t)kswz;
      }
      __syncwarp();
    }
  }
}

}  // namespace

int t16_build(const SegMatrix &src, T16Matrix &T, cudaStream_t st) {
  TFX_TRY(ensure_init());
  Context &c = ctx();
  T.release();
  if (src.nnz == 0 || src.nseg == 0) return 0;
  // ---- index range, ordering, output range
  DevBuf<int> flags, mm;
  TFX_TRY(flags.alloc(1)); TFX_TRY(mm.alloc(2));
  int h_mm[2] = {0x7fffffff, -1}, h_flag = 0;
  TFX_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int), st));
  TFX_CUDA(cudaMemcpyAsync(mm.p, h_mm, sizeof(h_mm), cudaMemcpyHostToDevice, st));
  t16_scan_kernel<<<c.num_sms * 8, 256, 0, st>>>(src.ptr.p, src.idx.p, src.nseg, flags.p, mm.p);
  c.launches++;
  TFX_CUDA(cudaMemcpyAsync(&h_flag, flags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaMemcpyAsync(h_mm, mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
  std::vector<int32_t> h_segmap((size_t)src.nseg);
  TFX_CUDA(cudaMemcpyAsync(h_segmap.data(), src.segmap.p, (size_t)src.nseg * 4, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  if (h_flag) return 0;   // indices not strictly ascending inside a segment: stay on the generic kernels
  const int32_t out_lo = *std::min_element(h_segmap.begin(), h_segmap.end());
  const int32_t out_hi = *std::max_element(h_segmap.begin(), h_segmap.end());
  T.out0 = out_lo;
  T.nseg = out_hi - out_lo + 1;
  T.in0 = h_mm[0];
  T.nin = h_mm[1] - h_mm[0] + 1;
  T.nnz = src.nnz;
  // ---- mode and tile size. DIRECT: one tile (or tile after tile when the output side is the long one);
  // TILES: one partial per (tile, output), the largest tile (longest segments) that leaves a few per CTA.
  // A long output side (S^T u: outputs = columns) with MANY gathered tiles would need a huge partial table: DIRECT,
  // tile after tile. With a handful of tiles (data rows beyond 16384: 2-8 tiles) one TILES launch over all
  // (tile, block) items balances far better than one DIRECT launch per tile (measured on 8 GPUs, 80 000 rows x 0.5 M
  // columns per rank: 5.3 ms as 5 DIRECT launches).
  const bool few_tiles = (int64_t)(T.nin + kT16TilesMaxTile - 1) / kT16TilesMaxTile <= 16;
  const bool long_out = (int64_t)T.nseg > (int64_t)1 << 18 && !(T.nin > g_opt_t16_direct_max && few_tiles);
  int tile;
  if (g_opt_t16_tile > 0) tile = g_opt_t16_tile;
  else if (T.nin <= g_opt_t16_direct_max || long_out) tile = std::min(T.nin, kT16MaxTile);
  else tile = kT16TilesMaxTile;   // the grid no longer depends on the tile count: the longest segments win
  tile = std::max(2, std::min(tile, kT16MaxTile));
  T.mode = ((T.nin + tile - 1) / tile == 1 || long_out) ? T16_DIRECT : T16_TILES;
  if (T.mode == T16_TILES) tile = std::min(tile, kT16TilesMaxTile);
  T.tile = tile;
  T.ntiles = (T.nin + tile - 1) / tile;
  const int64_t table = (int64_t)T.nseg * T.ntiles;
  if (table > ((int64_t)1 << 31)) return 0;        // pointer table would exceed 16 GiB: keep the generic kernels
  // The layout pays when segments are long. A matrix with ~1 entry per (tile, output) -- e.g. the diagonal damping
  // block alpha*I, 4.2e6 entries but a 1e9-entry pointer table -- stays on the generic CSR kernels.
  // (not applied when the layouts are forced with option t16_min_nnz = 0, as the layout tests do)
  if (g_opt_t16_min_nnz > 0 && table > std::max<int64_t>(T.nnz / 2, (int64_t)1 << 16)) return 0;

  // ---- output -> stored segment
  DevBuf<int32_t> segof;
  TFX_TRY(segof.alloc((size_t)T.nseg));
  TFX_CUDA(cudaMemsetAsync(segof.p, 0xff, (size_t)T.nseg * 4, st));
  t16_segof_kernel<<<std::min(c.num_sms * 8, (src.nseg + 255) / 256), 256, 0, st>>>(src.segmap.p, src.nseg, T.out0, segof.p);
  c.launches++;
  // ---- counts -> pointers
  TFX_TRY(T.ptr.alloc((size_t)table + 1));
  const int cgrid = (int)std::min<int64_t>((table + 255) / 256, (int64_t)c.num_sms * 32);
  t16_count_kernel<<<cgrid, 256, 0, st>>>(src.ptr.p, src.idx.p, segof.p, T.nseg, T.ntiles, tile, T.in0, T.ptr.p);
  c.launches++;
  TFX_CUDA(cudaMemsetAsync(T.ptr.p + table, 0, 8, st));
  {
    thrust::device_ptr<int64_t> P(T.ptr.p);
    TFX_THRUST(thrust::exclusive_scan(thrust::cuda::par.on(st), P, P + table + 1, P));
    c.launches += 2;
  }
  int64_t padded = 0;
  TFX_CUDA(cudaMemcpyAsync(&padded, T.ptr.p + table, 8, cudaMemcpyDeviceToHost, st));
  TFX_CUDA(cudaStreamSynchronize(st));
  T.nnz_padded = padded;
  TFX_TRY(T.val.alloc((size_t)padded + 8));
  TFX_TRY(T.key.alloc((size_t)padded + 8));
  TFX_CUDA(cudaMemsetAsync(T.val.p, 0, ((size_t)padded + 8) * 4, st));
  TFX_CUDA(cudaMemsetAsync(T.key.p, 0, ((size_t)padded + 8) * 2, st));
  const int fgrid = (int)std::min<int64_t>((table + 7) / 8, (int64_t)c.num_sms * 32);
  // packet width of the kernel that will stream the long segments of this layout (see t16_spmv)
  const bool ring = t16_uses_ring(T.mode, tile);
  t16_fill_kernel<<<fgrid, 256, 0, st>>>(src.ptr.p, src.idx.p, src.val.p, segof.p, T.nseg, T.ntiles, tile, T.in0, T.ptr.p,
                                         T.val.p, T.key.p, ring ? 4 : 2, g_opt_t16_bank_deal);
  c.launches++;
  // ---- work distribution: one counter per tile (TILES) / one counter (DIRECT)
  TFX_TRY(T.counter.alloc((size_t)T.ntiles + 1));
  T.nsplit = 1;
  if (T.mode == T16_TILES) TFX_TRY(T.partial.alloc((size_t)table));
  TFX_CUDA(cudaStreamSynchronize(st));
  TFX_CUDA(cudaGetLastError());
  T.valid = true;
  return 0;
}

}  // namespace tfx
