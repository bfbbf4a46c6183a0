// tc_gram.cuh -- the tcgen05 / TMEM / TMA Gram mainloop shared by the batch-hard mining kernel,
// the gallery top-k kernel and the all-pairs cosine kernel (sm_100a only).
//
//   G[i][j] = sum_k A[i][k] * B[j][k]      A: (M, Kp) bf16, B: (N, Kp) bf16, both K-major, fp32 accumulate
//
// One CTA = 18 warps: warps 0-15 epilogue (warp w reads TMEM lane quarter w%4, 64-column slice w/4; four
// warps per scheduler hide the TMEM / L1 latencies of one another), warp 16 TMA producer (one elected lane),
// warp 17 MMA issuer (one elected lane) + TMEM owner.
// Tile = 128 rows x 256 columns x K (UMMA M=128, N=256, K=16, cta_group::1), accumulator
// double-buffered in TMEM (2 x 256 columns = all 512), so the epilogue of tile t overlaps the MMAs of
// tile t+1.  The A row-block (128 x Kp) stays resident in shared memory while the CTA walks its
// column tiles; B streams through a ring of 32 KB K-block stages (256 rows x 64 bf16, SWIZZLE_128B).
// The Gram tile never leaves the SM: the epilogue functor consumes it 32 columns at a time straight
// from TMEM registers.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace efm {
namespace tc {

// Two tile shapes, selected by the template parameter MH (number of 128-row UMMA halves per CTA tile):
//   MH = 1: 128 rows x 256 columns  (A block 16 KB / K-block, B stage 32 KB)   -- any K <= 512
//   MH = 2: 256 rows x 128 columns  (A block 32 KB / K-block, B stage 16 KB)   -- K <= 256; halves the bytes
//           every SM must pull from L2 per MAC (the 128 x 256 shape is bound by the per-SM L2 port)
// Either way the CTA tile is 32768 accumulators = 2 x 256 TMEM columns with double buffering.
constexpr int BK = 64;
template <int MH> struct Shape {
  static constexpr int BM = 128 * MH;
  static constexpr int BN = 256 / MH;
  static constexpr int A_KB_BYTES = BM * BK * 2;
  static constexpr int B_ST_BYTES = BN * BK * 2;
};
constexpr int BM = 128, BN = 256;          // MH = 1 shape (host-side padding granularity)
constexpr int EPI_WARPS = 16;             // 4 per TMEM lane quarter: each owns a 64-column slice of the tile
constexpr int EPI_SLICES = EPI_WARPS / 4;
constexpr int COLS_PER_WARP = 64;          // BN / (EPI_SLICES / MH) for both shapes
constexpr int PRODUCER_WARP = EPI_WARPS;
constexpr int MMA_WARP = EPI_WARPS + 1;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int MAX_STAGES = 8;
constexpr int MAX_NK = 8;                 // K <= 512
constexpr int TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 232448;        // 227 KB opt-in maximum per CTA

struct GramParams {
  unsigned long long* trace;   // profiling aid (env EFM_GRAM_TRACE=<device ptr>): 16 u64 per CTA -- globaltimer stamps 0-7 of the
                               // tile pipeline, 8-12 of the exchange warp, 13/14 = ns / count of the producer's peer waits
  int debug;             // profiling aid (env EFM_GRAM_DEBUG): 1 = skip the epilogue work, 2 = TMEM loads only
  int M, N;              // valid rows of A / rows of B (= Gram columns)
  int nk;                // K blocks of 64 (Kp = 64 * nk)
  int stages;            // B ring depth
  int row_blocks, col_tiles;
  int sched_mode;        // 0: contiguous tile ranges over the whole grid; 1: `splits` CTAs per row block
  int splits;
  // Column blocks ("phases", multi-GPU mining): the N columns are `phases` equal blocks, one per rank, visited
  // block-major starting at block `phase0` (the local rank's rows, already resident).  Before the first tile of any
  // other block the kernel waits until that block's arrival counter (incremented by this kernel's own exchange warps as
  // their copies over NVLink complete, see CommParams) has reached `ready_target`.  phases = 1: one block, no waits.
  int phases, phase0;
  const unsigned long long* ready;       // [phases] arrival counters in LOCAL memory; nullptr = nothing to wait for
  unsigned long long ready_target;
};

// ---- peer exchange fused into the Gram kernel (one process per GPU, peers mapped through CUDA IPC) -------------
// Every rank owns an exchange region with the same layout; rank r's rows live in block r of every array.  The exchange is
// a PULL: a rank's prep kernel leaves its block in its OWN region, the Gram kernel announces it to the peers with one flag
// store per peer (the data was written by a previous kernel, so no fence is needed) and the exchange warps of every other
// rank copy the block into their own region with loads over NVLink -- block by block in the order of the receiver's
// block-major tile schedule, each warp taking its share of every block.  After a block: a DEVICE-scope fence and one
// add on the local arrival counter the TMA producer / epilogue warps wait on.  No store ever crosses NVLink except the
// flags, and no system-scope fence is executed on an SM that runs the tensor pipeline (measured: every MEMBAR.SYS of a
// push-style exchange costs 10-60 us while the links are busy and stalls the SM's TMA traffic for that long).
constexpr int MAX_PEERS = 8;
constexpr int MAX_XCHG_SEGS = 6;
constexpr int COMM_WARPS = 2;            // exchange warps per CTA (COMM kernels): 2 x 32 lanes x 8 x 16 B = 8 KB of loads in flight per SM
constexpr int PULL_DEPTH = 8;            // 16-byte loads in flight per lane.  Measured at 8 GPUs: 16 is SLOWER (Gram kernel 334 vs 252 us) --
                                         // the remote loads (~8 us round trip under load) fill the SM's pending-request slots and the
                                         // epilogue's own loads queue behind them
struct XchgSeg { unsigned long long off; unsigned long long bytes; };   // offset of block 0 of an array inside a region; bytes per rank block
struct CommParams {
  int world, rank;
  int nseg_a, nseg;                      // segments [0, nseg_a): stage A (what mining needs), [nseg_a, nseg): stage B (only read after mining)
  uint8_t* peer[MAX_PEERS];              // base of rank r's exchange region as mapped in this process (peer[rank] = own)
  XchgSeg seg[MAX_XCHG_SEGS];
  unsigned long long ready_off;          // u64[world] in every region: ready[r] = last step whose block rank r has announced
  unsigned long long cnt_a_off, cnt_b_off;   // u64[world] LOCAL arrival counters per source rank, stage A / stage B
  unsigned long long step;
  // Second job of the exchange warps, after their copies: once stage A of every rank has arrived (the labels of the whole
  // global batch are local), register this warp's share of the rows in the per-label member table that the finalize
  // kernel reads -- the table build is hidden behind the Gram tiles instead of being a kernel of its own.
  const int32_t* tab_lab;                // (tab_rows) labels, nullptr = no table job
  long tab_rows;
  unsigned long long* hkeys; int32_t* hcnt; int32_t* hmem; int32_t* hhead; int32_t* hnext; int32_t* hslot; int tbits;
};

// register `row` with its label: open-addressing table of 64-bit tags; the first 8 members of a label are stored inline
// (no pointer chasing in the finalize kernel), further members go to an overflow chain.  Tables start as all-ones bytes.
__device__ __forceinline__ void label_table_insert(int32_t lab, long row, unsigned long long* __restrict__ hkeys,
                                                   int32_t* __restrict__ hcnt, int32_t* __restrict__ hmem,
                                                   int32_t* __restrict__ hhead, int32_t* __restrict__ hnext,
                                                   int32_t* __restrict__ hslot, int tbits) {
  const unsigned long long tag = (unsigned long long)(uint32_t)lab;     // upper half 0: never equals EMPTY (~0)
  const uint32_t mask = (1u << tbits) - 1u;
  uint32_t sl = ((uint32_t)lab * 2654435761u) >> (32 - tbits);
  for (;;) {
    const unsigned long long old = atomicCAS(&hkeys[sl], ~0ull, tag);
    if (old == ~0ull || old == tag) break;
    sl = (sl + 1) & mask;
  }
  hslot[row] = (int32_t)sl;
  const int pos = atomicAdd(&hcnt[sl], 1) + 1;                       // counter starts at -1
  if (pos < 8) hmem[sl * 8 + pos] = (int32_t)row;
  else hnext[row] = atomicExch(&hhead[sl], (int32_t)row);
}

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires), so waiting warps do not burn issue slots of the warps that share their scheduler.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(200000u)
      : "memory");
  return done;
}
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("efmtl: mbarrier timeout block %d thread %d bar 0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > 20000u) mbar_timeout(bar, parity);               // >= ~2 s of hardware sleeps
}
// ---- 2-CTA (cta_group::2) helpers: the pair shares every B stage; CTA rank 0 issues the MMAs ---------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 remote;\n\t"
      "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t}"
      ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
constexpr int TRACE_SLOTS = 16;
#define EFM_TRACE(slot) do { if (g.trace && (threadIdx.x & 31) == 0) g.trace[(size_t)blockIdx.x * TRACE_SLOTS + (slot)] = gtimer(); } while (0)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}
// ---- cross-GPU helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_sys_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// acquire-release fence at system scope (MEMBAR.ALL.SYS): cheaper than __threadfence_system(), which is a
// sequentially-consistent fence (MEMBAR.SC.SYS)
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
static __device__ __noinline__ void peer_timeout(const void* p, unsigned long long have, unsigned long long want) {
  printf("efmtl: peer wait timeout block %d thread %d counter %p = %llu < %llu\n", blockIdx.x, threadIdx.x, p, have, want);
  __trap();
}
// One thread polls (acquire, system scope) with back-off.  Bounded: a lost peer traps instead of hanging the GPU.
__device__ __forceinline__ void wait_counter_thread(const unsigned long long* p, unsigned long long target) {
  unsigned long long v = ld_acquire_sys_u64(p);
  if (v < target) {
    const unsigned long long t0 = gtimer();
    while ((v = ld_relaxed_sys_u64(p)) < target) {                           // relaxed polls ...
      __nanosleep(200);
      if (gtimer() - t0 > 8000000000ull) peer_timeout(p, v, target);         // 8 s
    }
    (void)ld_acquire_sys_u64(p);                                             // ... one acquire once the value is there
  }
}
// Whole warp calls; lane 0 polls, the rest of the warp is ordered behind it by the warp barrier.
__device__ __forceinline__ void wait_counter(const unsigned long long* p, unsigned long long target) {
  if ((threadIdx.x & 31) == 0) wait_counter_thread(p, target);
  __syncwarp();
}
// one warp copies its share of a block of a segment: 16-byte units, PULL_DEPTH loads (over NVLink) in flight per lane
__device__ __forceinline__ void pull_share(uint8_t* dst, const uint8_t* src, unsigned long long bytes, int part, int nparts, int lane) {
  const long U = (long)(bytes >> 4);
  const long lo = U * part / nparts, hi = U * (part + 1) / nparts;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  long u = lo + lane;
  for (; u + (PULL_DEPTH - 1) * 32 < hi; u += PULL_DEPTH * 32) {
    uint4 v[PULL_DEPTH];
#pragma unroll
    for (int k = 0; k < PULL_DEPTH; ++k) v[k] = __ldcg(s4 + u + k * 32);
#pragma unroll
    for (int k = 0; k < PULL_DEPTH; ++k) d4[u + k * 32] = v[k];
  }
  for (; u < hi; u += 32) d4[u] = __ldcg(s4 + u);
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// The exchange of one warp (`part` of `nparts` over the whole grid): for every peer block in schedule order, wait for the
// owner's announcement, copy this warp's share of every segment of the stage, publish it locally.
__device__ __forceinline__ void comm_pull(const CommParams& cp, int part, int nparts, unsigned long long* trace) {
  const int lane = threadIdx.x & 31;
  uint8_t* own = cp.peer[cp.rank];
  const unsigned long long* ready = reinterpret_cast<const unsigned long long*>(own + cp.ready_off);
  for (int stage = 0; stage < 2; ++stage) {
    if (stage && trace && lane == 0) trace[9] = gtimer();            // stage A complete (this warp's shares)
    const int s0 = stage ? cp.nseg_a : 0, s1 = stage ? cp.nseg : cp.nseg_a;
    if (s0 == s1) continue;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(own + (stage ? cp.cnt_b_off : cp.cnt_a_off));
    for (int ph = 1; ph < cp.world; ++ph) {
      int blk = cp.rank + ph; if (blk >= cp.world) blk -= cp.world;  // the block this rank's tile schedule reaches at phase ph
      const unsigned long long tw = trace ? gtimer() : 0ull;
      wait_counter(ready + blk, cp.step);                            // the owner's prep of this step has completed
      if (trace && lane == 0) trace[15] += gtimer() - tw;
      const uint8_t* src = cp.peer[blk];
      for (int s = s0; s < s1; ++s) {
        const unsigned long long o = cp.seg[s].off + (unsigned long long)blk * cp.seg[s].bytes;
        pull_share(own + o, src + o, cp.seg[s].bytes, part, nparts, lane);
      }
      __threadfence();                                               // device scope: the copies are in L2
      __syncwarp();
      if (lane == 0) atomicAdd(cnt + blk, 1ull);
    }
  }
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled TMA load global -> shared, completion on an mbarrier (c0 = innermost coordinate)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart
constexpr uint64_t SMEM_DESC_BASE = (uint64_t(1) << 16)              // leading byte offset (unused for SW128 K-major)
                                    | (uint64_t(1024 >> 4) << 32)    // stride byte offset between 8-row groups
                                    | (uint64_t(1) << 46)            // descriptor version 1 (sm_100)
                                    | (uint64_t(2) << 61);           // SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) { return SMEM_DESC_BASE | uint64_t((addr & 0x3FFFFu) >> 4); }

// instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M=128, N given
__host__ __device__ constexpr uint32_t idesc_bf16_m128(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24); }
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread t gets lane base+t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait + make the destination registers data-dependent on the wait, so no consumer can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// ---- tile schedule -------------------------------------------------------------------------------------
// Tiles are visited phase by phase (one phase = one column block; a single phase unless the columns come from several
// ranks).  Inside a phase a CTA (pair) owns a contiguous range of the phase's tiles in row-major order -- the same range in
// every phase -- so its A row block stays resident while it walks column tiles.  Odd phases walk the range BACKWARDS
// (serpentine): a phase starts in the row block the previous one ended in, so a phase boundary costs no A reload, no
// accumulator drain and no row-state flush.  All three roles (TMA, MMA, epilogue) run the same iterator.
template <int CG>
struct TileIter {
  int CTp, phases, phase0;   // column tiles per phase
  int sched_mode, splits, row_blocks;
  int ph, left, dir;         // current phase; tiles left in it after the current one; walking direction
  int rb, ctl, ct0;          // row block; column tile inside the phase's block; first column tile of that block
  bool new_rb, new_phase;
  __device__ __forceinline__ explicit TileIter(const GramParams& g)
      : CTp(g.col_tiles / g.phases), phases(g.phases), phase0(g.phase0), sched_mode(g.sched_mode), splits(g.splits),
        row_blocks(g.row_blocks), ph(-1), left(0), dir(1), rb(-1), ctl(0), ct0(0), new_rb(false), new_phase(false) {}
  __device__ __forceinline__ void range(long& t0, long& t1) const {
    const long CT = CTp;
    if (sched_mode == 0) {
      const long T = (long)row_blocks * CT, G = gridDim.x / CG, c = blockIdx.x / CG;   // CG = 2: one range per CTA pair
      t0 = c * T / G;
      t1 = (c + 1) * T / G;
    } else {                                                         // `splits` CTAs (pairs) per row block; single phase
      const long c = blockIdx.x / CG;
      const long r = c / splits, s = c % splits;
      t0 = r * CT + s * CT / splits;
      t1 = r * CT + (s + 1) * CT / splits;
    }
  }
  __device__ __forceinline__ bool next() {
    if (left > 0) {
      --left;
      new_phase = false;
      if (dir > 0) { if (++ctl == CTp) { ctl = 0; ++rb; new_rb = true; } else new_rb = false; }
      else { if (--ctl < 0) { ctl = CTp - 1; --rb; new_rb = true; } else new_rb = false; }
      return true;
    }
    while (++ph < phases) {
      long t0, t1;
      range(t0, t1);                                                 // the only divisions: once per phase
      if (t1 > t0) {
        left = (int)(t1 - t0) - 1;
        dir = (ph & 1) ? -1 : 1;
        const long start = dir > 0 ? t0 : t1 - 1;
        const int nrb = (int)(start / CTp);
        ctl = (int)(start % CTp);
        int blk = phase0 + ph; if (blk >= phases) blk -= phases;
        ct0 = blk * CTp;
        new_rb = nrb != rb;                                          // rb = -1 before the first tile
        rb = nrb;
        new_phase = true;
        return true;
      }
    }
    return false;
  }
  __device__ __forceinline__ int ct() const { return ct0 + ctl; }    // global column tile
  __device__ __forceinline__ int block() const { return ct0 / CTp; }
};

template <int MH> inline size_t smem_bytes(int nk, int stages, int extra) {
  return 1024 /*alignment slack*/ + (size_t)nk * Shape<MH>::A_KB_BYTES + (size_t)stages * Shape<MH>::B_ST_BYTES + 256 /*barriers*/ + extra;
}
template <int MH> inline int pick_stages(int nk, int extra) {
  int s = (int)((SMEM_LIMIT - 1024 - 256 - extra - (long)nk * Shape<MH>::A_KB_BYTES) / Shape<MH>::B_ST_BYTES);
  return s > MAX_STAGES ? MAX_STAGES : s;
}

// ---- the kernel ----------------------------------------------------------------------------------------
// Epi supplies:  struct Params;  struct State;
//   static __device__ void row_begin(const Params&, const GramParams&, int row, State&);
//   static constexpr int SMEM_PER_WARP;   (bytes of scratch shared memory per epilogue warp, may be 0)
//   static constexpr int CW;              (16 or 32: columns per TMEM load / chunk)
//   static __device__ void chunk(const Params&, const GramParams&, State&, int row, const uint32_t (&acc)[CW], int col0,
//                                float* scratch /* the warp's scratch */, int bc /* (tile parity) * 4 + chunk ordinal */);
//   static __device__ void prefetch(const Params&, const GramParams&, float* scratch, int buf, int colw);  (may be a no-op)
//   static __device__ void prefetch_wait();
//   static __device__ void row_end(const Params&, const GramParams&, int row, int part, State&);
// row = anchor/query row index of this thread, col0 = first Gram column of the 32-column chunk held in
// acc (fp32 bit patterns); part = this warp's column slice (0..3) for partial-result slots.
// CG = 2 (requires MH = 1): two CTAs of a cluster form a 256 x 256 tile.  Each CTA keeps its own 128 A rows resident and
// loads HALF of every B stage (128 of the 256 Gram columns, 16 KB); CTA rank 0 issues tcgen05.mma.cta_group::2 (M = 256)
// which reads both halves, so every SM pulls half the B bytes from L2 and reads half the B bytes from shared memory per
// MAC.  Barriers: full / a_full / acc_empty live in the leader (TMA of both CTAs credits the leader, the peer's epilogue
// arrives remotely); empty / a_empty / acc_full are signalled in both CTAs by multicast tcgen05.commit.
// COMM: COMM_WARPS extra warps copy the peers' rows into this rank's exchange region while the tiles run (see CommParams).
template <class Epi, int MH, int CG, bool COMM>
__global__ void __launch_bounds__(THREADS + (COMM ? 32 * COMM_WARPS : 0), 1)
gram_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GramParams g,
            const typename Epi::Params ep, const CommParams cp) {
  static_assert(CG == 1 || MH == 1, "the 2-CTA path uses the 128 x 256 per-CTA shape");
  constexpr int BM = Shape<MH>::BM, BN = Shape<MH>::BN;              // per-CTA accumulator tile
  constexpr int A_KB_BYTES = Shape<MH>::A_KB_BYTES, B_ST_BYTES = Shape<MH>::B_ST_BYTES / CG;   // CG = 2: half a B stage per CTA
  constexpr int B_ROWS = BN / CG;                                    // Gram columns this CTA loads per stage
  constexpr int ACC_COLS = 256;                                      // TMEM columns per accumulator buffer (MH * BN)
  constexpr uint32_t IDESC = CG == 2 ? idesc_bf16(256, BN) : idesc_bf16_m128(BN);
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;           // 0 = leader
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B atoms need 1024 B alignment
  const uint32_t sA = base;
  const uint32_t sB = sA + (uint32_t)g.nk * A_KB_BYTES;
  const uint32_t sBar = sB + (uint32_t)g.stages * B_ST_BYTES;
  // barrier slots (8 B each)
  const uint32_t bar_full = sBar;                       // [MAX_STAGES]
  const uint32_t bar_empty = sBar + 8 * MAX_STAGES;     // [MAX_STAGES]
  const uint32_t bar_a_full = sBar + 16 * MAX_STAGES;
  const uint32_t bar_a_empty = bar_a_full + 8;
  const uint32_t bar_acc_full = bar_a_empty + 8;        // [2]
  const uint32_t bar_acc_empty = bar_acc_full + 16;     // [2]
  const uint32_t tmem_slot = bar_acc_empty + 16;        // u32
  float* scratch = reinterpret_cast<float*>(smem_raw + (sBar + 256 - smem_u32(smem_raw))) +
                   ((threadIdx.x >> 5) & (EPI_WARPS - 1)) * (Epi::SMEM_PER_WARP / 4);   // epilogue warps are warps 0..15
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) EFM_TRACE(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_empty, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(bar_acc_full + 8 * a, 1); mbar_init(bar_acc_empty + 8 * a, EPI_WARPS * CG); }
    fence_barrier_init();
  }
  if (warp == PRODUCER_WARP && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == MMA_WARP) {
    if constexpr (CG == 2) { tmem_alloc_2cta(tmem_slot, TMEM_COLS); tmem_relinquish_2cta(); }
    else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // barriers + TMEM of both CTAs are live past this point
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();                                                     // the next kernel's CTAs may take this SM as soon as we exit
  pdl_wait();                                                        // everything above overlapped the previous kernel's tail
  if (threadIdx.x == 0) EFM_TRACE(1);

  if (warp == PRODUCER_WARP) {
    // whole warp runs the (warp-uniform) loop so counters live in uniform registers; one elected lane issues
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0, a_empty_phase = 0;
    TileIter<CG> it(g);
    bool first = true;
    while (it.next()) {
      const int rb = it.rb, ct = it.ct();
      if (it.new_phase && it.ph > 0 && g.ready) {                    // a peer's column block: wait until it has landed
        const unsigned long long tw = g.trace ? gtimer() : 0ull;
        wait_counter(g.ready + it.block(), g.ready_target);
        fence_proxy_async_all();                                     // the exchange warps' generic-proxy writes -> our TMA (async proxy) reads
        if (g.trace && lane == 0) {
          g.trace[(size_t)blockIdx.x * TRACE_SLOTS + 13] += gtimer() - tw;
          g.trace[(size_t)blockIdx.x * TRACE_SLOTS + 14] += 1;
        }
      }
      if (it.new_rb) {                                               // new row block: (re)load the resident A block
        if (!first) { mbar_wait(bar_a_empty, a_empty_phase); a_empty_phase ^= 1; }
        if (leader) {
          if constexpr (CG == 2) {
            if (crank == 0) mbar_expect_tx(bar_a_full, (uint32_t)g.nk * A_KB_BYTES * 2);      // both CTAs' A blocks
            for (int kb = 0; kb < g.nk; ++kb)
              tma_load_2d_2cta(sA + kb * A_KB_BYTES, &tmA, bar_a_full, kb * BK, (rb * 2 + (int)crank) * BM);
          } else {
            mbar_expect_tx(bar_a_full, (uint32_t)g.nk * A_KB_BYTES);
            for (int kb = 0; kb < g.nk; ++kb) tma_load_2d(sA + kb * A_KB_BYTES, &tmA, bar_a_full, kb * BK, rb * BM);
          }
        }
      }
      for (int kb = 0; kb < g.nk; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (leader) {
          if constexpr (CG == 2) {                                  // this CTA's half of the B stage -> leader's barrier
            if (crank == 0) mbar_expect_tx(bar_full + 8 * stage, B_ST_BYTES * 2);
            tma_load_2d_2cta(sB + stage * B_ST_BYTES, &tmB, bar_full + 8 * stage, kb * BK, ct * BN + (int)crank * B_ROWS);
          } else if (g.debug & 4) { mbar_arrive(bar_full + 8 * stage); }    // profiling: no B traffic
          else {
            mbar_expect_tx(bar_full + 8 * stage, B_ST_BYTES);
            tma_load_2d(sB + stage * B_ST_BYTES, &tmB, bar_full + 8 * stage, kb * BK, ct * BN);
          }
        }
        if (++stage == g.stages) { stage = 0; phase ^= 1; }
      }
      if (first) EFM_TRACE(2);
      first = false;
    }
  } else if (warp == MMA_WARP && crank == 0) {
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0, a_full_phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint64_t descA0 = smem_desc(sA), descB0 = smem_desc(sB);
    TileIter<CG> nx(g);
    bool has = nx.next();
    while (has) {
      const bool new_rb = nx.new_rb;
      has = nx.next();                                               // nx now describes the tile after this one
      mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);             // epilogue drained this accumulator
      if (new_rb) { mbar_wait(bar_a_full, a_full_phase); a_full_phase ^= 1; }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
      for (int kb = 0; kb < g.nk; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (leader) {
          const uint64_t da = descA0 + (uint64_t)((kb * A_KB_BYTES) >> 4);
          const uint64_t db = descB0 + (uint64_t)((stage * B_ST_BYTES) >> 4);
          if (!(g.debug & 8)) {                                      // profiling: no MMA issue
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
              for (int h = 0; h < MH; ++h) {                         // row half h: A rows [128h, +128) -> TMEM columns [h*BN, +BN)
                if constexpr (CG == 2)
                  umma_bf16_2cta(d_tmem, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), IDESC, (kb | k) ? 1u : 0u);
                else
                  umma_bf16(d_tmem + h * BN, da + (uint64_t)((h * (128 * BK * 2) + k * 32) >> 4), db + (uint64_t)((k * 32) >> 4),
                            IDESC, (kb | k) ? 1u : 0u);
              }
            }
          }
          if constexpr (CG == 2) umma_commit_2cta(bar_empty + 8 * stage);
          else umma_commit(bar_empty + 8 * stage);                   // stage reusable once these MMAs retire
        }
        if (++stage == g.stages) { stage = 0; phase ^= 1; }
      }
      const bool last_of_rb = !has || nx.new_rb;
      if (leader) {
        if constexpr (CG == 2) {
          umma_commit_2cta(bar_acc_full + 8 * acc);
          if (last_of_rb) umma_commit_2cta(bar_a_empty);
        } else {
          umma_commit(bar_acc_full + 8 * acc);                       // accumulator ready for the epilogue
          if (last_of_rb) umma_commit(bar_a_empty);                  // A row-block may be overwritten
        }
      }
      if (!has) EFM_TRACE(5);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < EPI_WARPS) {
    // ---------------- epilogue warps ----------------
    constexpr int COL_SLICES = EPI_SLICES / MH;                      // column slices of 64 per row half
    const int quarter = warp & 3, slice = warp >> 2;
    const int rowhalf = slice / COL_SLICES, cs = slice % COL_SLICES;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    int acc = 0; uint32_t acc_phase = 0;
    typename Epi::State st;
    int row = 0;
    const int part = cs;
    constexpr int CW = Epi::CW;                                      // columns per TMEM load / epilogue chunk
    TileIter<CG> nx(g);
    bool has = nx.next();
    if (has) {
      if (nx.ph > 0 && g.ready) wait_counter(g.ready + nx.block(), g.ready_target);   // no tile in the local block
      Epi::prefetch(ep, g, scratch, 0, nx.ct() * BN + cs * COLS_PER_WARP);
    }
    int i = 0;
    while (has) {
      if (nx.new_rb) {
        if (i != 0) Epi::row_end(ep, g, row, part, st);
        row = (nx.rb * CG + (int)crank) * BM + rowhalf * 128 + quarter * 32 + lane;
        Epi::row_begin(ep, g, row, st);
      }
      const int colw = nx.ct() * BN + cs * COLS_PER_WARP;            // first column of this warp's slice
      has = nx.next();                                               // nx now describes the tile after this one
      Epi::prefetch_wait();                                          // this tile's per-column data has landed
      if (has) {                                                     // next tile's data: in flight during this tile
        if (nx.new_phase && g.ready) wait_counter(g.ready + nx.block(), g.ready_target);   // ... once its block has arrived
        Epi::prefetch(ep, g, scratch, (i + 1) & 1, nx.ct() * BN + cs * COLS_PER_WARP);
      }
      mbar_wait(bar_acc_full + 8 * acc, acc_phase);
      tc_fence_after();
      if (i == 0 && warp == 0) EFM_TRACE(3);
      const uint32_t t_acc = tmem_base + lane_base + (uint32_t)(acc * ACC_COLS + rowhalf * BN + cs * COLS_PER_WARP);
      if ((g.debug & 3) != 1) {
#pragma unroll
        for (int c = 0; c < COLS_PER_WARP / CW; ++c) {
          const int col0 = colw + c * CW;
          if (col0 < g.N) {                                          // warp-uniform
            uint32_t v[CW];
            tmem_ld(t_acc + c * CW, v);
            tmem_ld_wait(v);
            if ((g.debug & 3) != 2) Epi::chunk(ep, g, st, row, v, col0, scratch, (i & 1) * 4 + c);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(bar_acc_empty + 8 * acc, 0);   // the leader's MMA thread waits for both CTAs
        else mbar_arrive(bar_acc_empty + 8 * acc);
      }
      if (warp == 0) { if (i == 0) EFM_TRACE(4); if (!has) EFM_TRACE(6); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      ++i;
    }
    if (i > 0) Epi::row_end(ep, g, row, part, st);
  } else if (COMM && warp > MMA_WARP) {
    const int cw = warp - (MMA_WARP + 1);
    const int part = (int)blockIdx.x * COMM_WARPS + cw, nparts = (int)gridDim.x * COMM_WARPS;
    const bool tr = g.trace && cw == 0;
    if (tr) EFM_TRACE(8);
    if (cp.world > 1) {
      if (blockIdx.x == 0 && cw == 0) {                               // announce this rank's block: one flag store per peer
        fence_acq_rel_sys();                                           // (one warp of one CTA, nothing in flight yet: cheap)
        if (lane < cp.world && lane != cp.rank)
          st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(cp.peer[lane] + cp.ready_off) + cp.rank, cp.step);
      }
      comm_pull(cp, part, nparts, tr ? g.trace + (size_t)blockIdx.x * TRACE_SLOTS : nullptr);
    }
    if (tr) EFM_TRACE(10);
    if (cp.tab_lab) {
      if (g.ready)
        for (int r = 0; r < cp.world; ++r)
          if (r != cp.rank) wait_counter(g.ready + r, g.ready_target);       // labels of every rank are here
      if (tr) EFM_TRACE(11);
      const long lo = cp.tab_rows * part / nparts, hi = cp.tab_rows * (part + 1) / nparts;
      for (long row = lo + lane; row < hi; row += 32)
        label_table_insert(__ldcg(cp.tab_lab + row), row, cp.hkeys, cp.hcnt, cp.hmem, cp.hhead, cp.hnext, cp.hslot, cp.tbits);
      if (tr) EFM_TRACE(12);
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody may still touch the peer's smem / TMEM
  tc_fence_after();
  if (warp == MMA_WARP) {
    __syncwarp();
    if constexpr (CG == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) EFM_TRACE(7);
}

// ---- host side: tensor maps + launch ----------------------------------------------------------------------
// 2-D bf16 row-major (rows, Kp) tensor map with a (64 x box_rows) box, SWIZZLE_128B, zero OOB fill.
int make_tmap_bf16(CUtensorMap* out, const void* base, long rows, long Kp, int box_rows);

template <class Epi, int MH, int CG, bool COMM = false>
int launch_gram_mh(const void* A, const void* B, GramParams g, const typename Epi::Params& ep, int grid, cudaStream_t st,
                   const CommParams* cp = nullptr, bool pdl = false) {
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, A, g.M, (long)g.nk * BK, 128);       // A box: 128 rows per TMA (MH = 2 issues it per 256-row block)
  if (MH == 2) rc = make_tmap_bf16(&tmA, A, g.M, (long)g.nk * BK, Shape<MH>::BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, B, g.N, (long)g.nk * BK, Shape<MH>::BN / CG);
  if (rc) return rc;
  if (g.phases < 1) g.phases = 1;
  const size_t smem = 1024 + (size_t)g.nk * Shape<MH>::A_KB_BYTES + (size_t)g.stages * (Shape<MH>::B_ST_BYTES / CG) + 256 +
                      Epi::SMEM_PER_WARP * EPI_WARPS;
  static bool attr_done[64] = {};                                    // per instantiation and per device (the attribute is per device)
  int dev = 0;
  EFM_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    EFM_CUDA(cudaFuncSetAttribute(gram_kernel<Epi, MH, CG, COMM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  CommParams none{};
  const CommParams& c = cp ? *cp : none;
  constexpr unsigned NT = THREADS + (COMM ? 32 * COMM_WARPS : 0);
  LaunchCfg lc(dim3((unsigned)grid), dim3(NT), smem, st, pdl, CG);
  EFM_CUDA(cudaLaunchKernelEx(&lc.cfg, gram_kernel<Epi, MH, CG, COMM>, tmA, tmB, g, ep, c));
  EFM_LAUNCH_CHECK("gram_kernel");
  return EFM_OK;
}

// tile shape: 128 x 256 by default -- measured faster than 256 x 128 on B200 for K = 256 (34.9 vs 38.9 us at
// B = 8192): the deeper B ring (4 x 32 KB in flight) matters more than the halved L2 traffic.  EFM_GRAM_MH=2
// selects the other shape for experiments.
inline int pick_mh(int M, int nk) {
  (void)M;
  if (const char* e = getenv("EFM_GRAM_MH")) { const int v = atoi(e); if (v == 1 || (v == 2 && nk <= 4)) return v; }
  return 1;
}

// 2-CTA pairs (cta_group::2): 128 x 256 per-CTA shape, at least one full pair of row blocks
inline int pick_cg(int M, int mh, int sched_mode) {
  (void)sched_mode;
  if (mh != 1 || M <= 128) return 1;
  if (const char* e = getenv("EFM_GRAM_CG")) { const int v = atoi(e); if (v == 1 || v == 2) return v; }
  return 2;
}

// fills row_blocks / col_tiles / stages / schedule and returns the grid size
inline int plan_gram(GramParams& g, int mh, int sched_mode, int smem_extra = 0, int cg = 1) {
  const int bm = 128 * mh * cg, bn = 256 / mh;
  { const char* e = getenv("EFM_GRAM_DEBUG"); g.debug = e ? atoi(e) : 0; }
  { const char* e = getenv("EFM_GRAM_TRACE"); g.trace = e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr; }
  g.stages = 0;
  if (const char* e = getenv("EFM_GRAM_STAGES")) { const int v = atoi(e); if (v >= 1) g.stages = -v; }
  g.row_blocks = (int)cdiv(g.M, bm);
  g.col_tiles = (int)cdiv(g.N, bn);
  { const int forced = g.stages < 0 ? -g.stages : 0;
    g.stages = mh == 2 ? pick_stages<2>(g.nk, smem_extra) : pick_stages<1>(g.nk, smem_extra);
    if (cg == 2) {                                                   // half-size stages: twice as many fit
      g.stages = (int)((SMEM_LIMIT - 1024 - 256 - smem_extra - (long)g.nk * Shape<1>::A_KB_BYTES) / (Shape<1>::B_ST_BYTES / 2));
      if (g.stages > MAX_STAGES) g.stages = MAX_STAGES;
    }
    if (forced && forced < g.stages) g.stages = forced; }
  g.sched_mode = sched_mode;
  if (g.phases < 1) g.phases = 1;                                    // caller sets phases > 1 (col_tiles % phases == 0, mode 0)
  const long T = (long)g.row_blocks * (g.col_tiles / g.phases);      // tiles of one phase: every CTA (pair) takes a share of each
  const int sms = num_sms();
  if (sched_mode == 0) {
    g.splits = 1;
    if (cg == 2) { const long pairs = sms / 2; return 2 * (int)(T < pairs ? T : pairs); }
    return (int)(T < sms ? T : sms);
  }
  // mode 1: every row block is cut into `splits` column ranges so that the grid covers the SMs
  int splits = (int)cdiv((long)(sms / cg), g.row_blocks);            // one wave of CTAs (pairs): fixed cost per CTA is ~8 us
  if (splits > g.col_tiles) splits = g.col_tiles;
  if (splits < 1) splits = 1;
  g.splits = splits;
  return g.row_blocks * splits * cg;
}

template <class Epi>
int launch_gram(const void* A, const void* B, int mh, const GramParams& g, const typename Epi::Params& ep, int grid,
                cudaStream_t st, int cg = 1, bool pdl = false) {
  EFM_REQUIRE(g.stages >= 1, EFM_E_UNSUPPORTED, "gram: K = %d leaves no shared memory for the B pipeline", g.nk * BK);
  if (cg == 2) return launch_gram_mh<Epi, 1, 2>(A, B, g, ep, grid, st, nullptr, pdl);
  if (mh == 2) return launch_gram_mh<Epi, 2, 1>(A, B, g, ep, grid, st, nullptr, pdl);
  return launch_gram_mh<Epi, 1, 1>(A, B, g, ep, grid, st, nullptr, pdl);
}

}  // namespace tc
}  // namespace efm
