// bh.cu -- fused batch-hard triplet loss, forward + backward (SURVEY 8 row a2', extends a1/a2/a3).
//
// Step = 4 launches on one stream, the B x B distance matrix never touches HBM:
//   1. bh_prep      per row: (optional) L2 normalise, round to the mining precision, squared norm of the
//                   mining row, copy labels, reset the mined keys, zero dE
//   2. mining       G = Xa . Xall^T tile by tile; per anchor row running
//                   argmax_{same label, j != i}(|x_j|^2 - 2 g_ij) and argmin_{other label}(...),
//                   merged across CTAs with packed 64-bit atomics (value, index) -> lowest index on ties
//                     * bf16 rows: tcgen05 tiles (tc_gram.cuh) with the mining epilogue below
//                     * fp32 rows: CUDA-core tiled kernel (parity path, rel 1e-5)
//   3. bh_finalize  per anchor: decode indices, recompute d_ap, d_an directly in fp32 (no Gram
//                   cancellation), hinge, scatter-add 2w(n-p), 2w(p-a), 2w(a-n) into dE
//   4. bh_norm_bwd  (normalize only) dE = (dX - x (x . dX)) / |e| in place
#include <string.h>

#include "tc_gram.cuh"

namespace efm {
namespace {

constexpr int WARPS = 8;

struct BhLayout {
  bool use_tc;
  bool xm_f32;          // fp32 mining rows materialised in the workspace
  int Dp;               // K padded to a multiple of 64 (tensor-core path)
  int tbits;            // log2 of the label hash-table size
  size_t kpos, kneg, sq, nrm, lab, xm, hkeys, hcnt, hhead, rheadp, rheadn, hmem, hnext, hslot, rnextp, rnextn, wgt, clr_begin, clr_bytes, total;
};

BhLayout bh_layout(long B_local, long B_all, long D, int dtype, int flags) {
  BhLayout L{};
  L.Dp = (int)(cdiv(D, tc::BK) * tc::BK);
  L.use_tc = (dtype != EFM_F32 || (flags & EFM_BH_MINE_BF16)) && L.Dp <= tc::MAX_NK * tc::BK;
  L.xm_f32 = !L.use_tc && (dtype != EFM_F32 || (flags & EFM_BH_NORMALIZE));
  L.tbits = 4;
  while ((1L << L.tbits) < 2 * B_all) ++L.tbits;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.kpos = take((size_t)B_local * 8);
  L.kneg = take((size_t)B_local * 8);
  const size_t Bpad = (size_t)cdiv(B_all, tc::BN) * tc::BN;  // sq/labels padded to whole column tiles (sq = +inf there)
  L.sq = take(Bpad * 4);
  L.nrm = take((size_t)B_all * 4);
  L.lab = take(Bpad * 4);
  L.xm = take(L.use_tc ? (size_t)B_all * L.Dp * 2 : (L.xm_f32 ? (size_t)B_all * D * 4 : 0));
  // per-label chains (tensor-core path): open-addressing table of (label -> head row), next pointers, slot per row
  // one memset(0xFF) resets [hkeys | hcnt | hhead | rheadp | rheadn]: empty label table, empty chains (-1 heads)
  L.hkeys = take(L.use_tc ? (size_t)8 << L.tbits : 0);
  L.hcnt = take(L.use_tc ? (size_t)4 << L.tbits : 0);        // members - 1 per label slot (starts at -1)
  L.hhead = take(L.use_tc ? (size_t)4 << L.tbits : 0);       // overflow chain (members beyond the 8 inline ones)
  L.rheadp = take((size_t)B_all * 4);                        // rows that chose r as hardest positive (chain heads)
  L.rheadn = take((size_t)B_all * 4);                        // rows that chose r as hardest negative
  L.clr_begin = L.use_tc ? L.hkeys : L.rheadp;
  L.clr_bytes = off - L.clr_begin;
  L.hmem = take(L.use_tc ? (size_t)32 << L.tbits : 0);       // first 8 member rows of every label slot, inline
  L.hnext = take(L.use_tc ? (size_t)B_all * 4 : 0);
  L.hslot = take(L.use_tc ? (size_t)B_all * 4 : 0);
  L.rnextp = take((size_t)B_local * 4);
  L.rnextn = take((size_t)B_local * 4);
  L.wgt = take((size_t)B_local * 4);                         // 2 * head_grad_i for active anchors, else 0
  L.total = off;
  return L;
}

// 8 consecutive elements -> fp32 (128-bit loads for 16-bit types, 2 x 128-bit for fp32)
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&o)[8]) {
  if constexpr (sizeof(T) == 2) {
    union { uint4 raw; T v[8]; } u;
    u.raw = __ldg(reinterpret_cast<const uint4*>(p));
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = El<T>::to_f(u.v[k]);
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
}

// ------------------------------------------------------------------------------------------ prep
// One warp per row of E_all.  VEC8: D % 8 == 0 and 16-byte aligned rows -> each lane owns 8-element slabs.
//
// Tensor-core path (use_tc): the MINING ROWS are the raw rows rounded to bf16 (for bf16 input: the rows themselves, and
// xm_bf16 == nullptr means "the Gram kernel reads E directly, write nothing"), zero-padded in K to a multiple of 64.  The
// normalisation is not applied to the rows but to the Gram tile: with normalize, sq[row] receives the COLUMN SCALE
// -1/|e_row| (mining metric: cosine, v_ij = g_ij * sq_j is monotone in the squared distance of the normalised rows for a
// fixed anchor); without, sq[row] = |bf16(e_row)|^2 (v_ij = sq_j - 2 g_ij).  Columns beyond B_all inside the last Gram tile
// get sq = +inf: v = +inf (plain) or 0 * inf = NaN (cosine), neither can ever be mined -- no edge path in the kernel.
// CUDA-core fp32 path (!use_tc): normalised fp32 rows in xm_f32 (when write_f32), sq = their squared norms.
template <typename T, bool VEC8>
__global__ void __launch_bounds__(WARPS * 32)
bh_prep_kernel(const T* __restrict__ E, const int32_t* __restrict__ labels, long B_all, int D, int Dp, int normalize,
               int use_tc, int write_f32, long row_offset, long B_local, float* __restrict__ sq,
               float* __restrict__ nrm, int32_t* __restrict__ lab_ws, __nv_bfloat16* __restrict__ xm_bf16,
               float* __restrict__ xm_f32, unsigned long long* __restrict__ kpos,
               unsigned long long* __restrict__ kneg, float* __restrict__ dE,
               unsigned long long* __restrict__ hkeys, int32_t* __restrict__ hcnt, int32_t* __restrict__ hmem,
               int32_t* __restrict__ hhead, int32_t* __restrict__ hnext, int32_t* __restrict__ hslot, int tbits,
               uint4* __restrict__ clr, long clr_n16) {
  pdl_trigger();                                    // the Gram kernel's prologue may overlap this kernel's tail
  // reset of the label table and the gradient chain heads (all-ones bytes), folded into this kernel when nothing in it
  // touches them (the table is then built inside the Gram kernel): saves a memset node and its launch gap
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < clr_n16; i += (long)gridDim.x * blockDim.x)
    clr[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  const long row = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= B_all) return;
  const int lane = threadIdx.x & 31;
  const T* er = E + row * D;
  float n = 1.f, s2 = 0.f;
  if constexpr (VEC8) {
    if (D <= 256) {                                  // whole row lives in registers: one pass
      float v[8];
      const bool act = lane * 8 < D;
      if (act) load8<T>(er + lane * 8, v);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
      }
      if (normalize) {
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) ss = fmaf(v[k], v[k], ss);
        n = sqrtf(warp_sum(ss));
      }
      if (use_tc) {
        if (xm_bf16 || !normalize) {
          union { uint4 raw; __nv_bfloat16 b[8]; } u;
#pragma unroll
          for (int k = 0; k < 8; ++k) { u.b[k] = __float2bfloat16_rn(v[k]); const float vb = __bfloat162float(u.b[k]); s2 = fmaf(vb, vb, s2); }
          if (xm_bf16 && lane * 8 < Dp) *reinterpret_cast<uint4*>(xm_bf16 + row * Dp + lane * 8) = u.raw;
        }
      } else {
        if (normalize) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = v[k] / n;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s2 = fmaf(v[k], v[k], s2);
        if (write_f32 && act) {
          float4* o = reinterpret_cast<float4*>(xm_f32 + row * D + lane * 8);
          o[0] = make_float4(v[0], v[1], v[2], v[3]); o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    } else {
      if (normalize) {
        float ss = 0.f;
        for (int d = lane * 8; d < D; d += 256) { float v[8]; load8<T>(er + d, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) ss = fmaf(v[k], v[k], ss); }
        n = sqrtf(warp_sum(ss));
      }
      if (!use_tc || xm_bf16 || !normalize) {
        for (int d = lane * 8; d < Dp; d += 256) {
          float v[8];
          if (d < D) load8<T>(er + d, v);
          else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.f;
          }
          if (use_tc) {
            union { uint4 raw; __nv_bfloat16 b[8]; } u;
#pragma unroll
            for (int k = 0; k < 8; ++k) { u.b[k] = __float2bfloat16_rn(v[k]); const float vb = __bfloat162float(u.b[k]); s2 = fmaf(vb, vb, s2); }
            if (xm_bf16) *reinterpret_cast<uint4*>(xm_bf16 + row * Dp + d) = u.raw;
          } else if (d < D) {
            if (normalize) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = v[k] / n;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) s2 = fmaf(v[k], v[k], s2);
            if (write_f32) {
              float4* o = reinterpret_cast<float4*>(xm_f32 + row * D + d);
              o[0] = make_float4(v[0], v[1], v[2], v[3]); o[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
        }
      }
    }
    if (dE) {
      float4* z = reinterpret_cast<float4*>(dE + row * D);
      for (int d4 = lane; d4 < D / 4; d4 += 32) z[d4] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    if (normalize) {
      float ss = 0.f;
      for (int d = lane; d < D; d += 32) { const float v = El<T>::to_f(er[d]); ss = fmaf(v, v, ss); }
      n = sqrtf(warp_sum(ss));
    }
    if (use_tc) {
      if (xm_bf16 || !normalize) {
        for (int d = lane; d < Dp; d += 32) {
          const float v = d < D ? El<T>::to_f(er[d]) : 0.f;
          const __nv_bfloat16 b = __float2bfloat16_rn(v);
          if (xm_bf16) xm_bf16[row * Dp + d] = b;
          const float vb = __bfloat162float(b);
          s2 = fmaf(vb, vb, s2);
        }
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        float v = El<T>::to_f(er[d]);
        if (normalize) v = v / n;
        if (write_f32) xm_f32[row * D + d] = v;
        s2 = fmaf(v, v, s2);
      }
    }
    if (dE) for (int d = lane; d < D; d += 32) dE[row * D + d] = 0.f;
  }
  s2 = warp_sum(s2);
  if (lane == 0) {
    sq[row] = (use_tc && normalize) ? -1.f / n : s2;
    nrm[row] = n;
    const int32_t lab = labels ? labels[row] : 0;
    if (lab_ws) {
      lab_ws[row] = lab;
      const long Bpad = (B_all + tc::BN - 1) / tc::BN * tc::BN;
      for (long j = B_all + row; j < Bpad; j += B_all) { sq[j] = INFINITY; lab_ws[j] = 0; }
    }
    const long li = row - row_offset;
    if (kpos && li >= 0 && li < B_local) { kpos[li] = KEY_POS_EMPTY; kneg[li] = KEY_NEG_EMPTY; }
    if (hkeys) tc::label_table_insert(lab, row, hkeys, hcnt, hmem, hhead, hnext, hslot, tbits);
  }
}

// ------------------------------------------------------------------------------------------ mining epilogue (tcgen05 path)
// Hardest NEGATIVE only, branch-free per element: v_ij = same_label ? +inf : |x_j|^2 - 2 g_ij (same order as
// d_ij for a fixed anchor; the anchor itself is a same-label column).  Per 32-column chunk: 8 group-of-4 minima,
// their minimum, and -- only when that beats the running best -- the index of the winning GROUP.  The column
// inside the group is resolved by bh_finalize with exact fp32 distances (4 candidates per anchor), so the hot
// loop carries no per-element index bookkeeping.  Hardest POSITIVES are mined in bh_finalize from the per-label
// chains (a handful of candidates per anchor).
struct MineEpi {
  static constexpr int SMEM_PER_WARP = 2 * 512;      // double-buffered {sq[64], labels[64]} of the warp's column slice
  static constexpr int CW = 16;
  struct Params {
    const float* sq;          // |x_j|^2 of the mining rows, (B_all)
    const int32_t* labels;    // (B_all)
    unsigned long long* kneg; // (B_local): (ord(v) << 32) | first column of the winning group of 4
    int row_offset;
    int norm;                 // 1: cosine metric, sq holds the column scale -1/|e_j| and v_ij = g_ij * sq_j; 0: v_ij = sq_j - 2 g_ij
  };
  struct State {
    int lab_i;
    float nbest;
    int ngrp;
    bool valid;
  };
  static __device__ __forceinline__ void row_begin(const Params& p, const tc::GramParams& g, int row, State& s) {
    s.valid = row < g.M;
    s.lab_i = s.valid ? p.labels[p.row_offset + row] : 0;
    s.nbest = INFINITY;
    s.ngrp = -1;
  }
  // per-column data of the warp's 64-column slice -> shared memory, one tile ahead (16 B per lane, cp.async)
  static __device__ __forceinline__ void prefetch(const Params& p, const tc::GramParams&, float* scratch, int buf, int colw) {
    const int lane = threadIdx.x & 31;
    float* dst = scratch + buf * 128 + lane * 4;                     // [0,64): sq, [64,128): labels (bit patterns)
    const void* src = lane < 16 ? (const void*)(p.sq + colw + lane * 4) : (const void*)(p.labels + colw + (lane - 16) * 4);
    tc::cp_async16(dst, src);
    tc::cp_async_commit();
  }
  static __device__ __forceinline__ void prefetch_wait() { tc::cp_async_wait_all(); __syncwarp(); }
  static __device__ __forceinline__ void chunk(const Params& p, const tc::GramParams& g, State& s, int row,
                                               const uint32_t (&acc)[16], int col0, float* scratch, int bc) {
    // no edge handling: sq/labels are padded to whole tiles with |x|^2 = +inf and TMA zero-fills the Gram there
    const float* tile = scratch + (bc >> 2) * 128;                   // this tile's {sq[64], labels[64]}
    const float4* sq4 = reinterpret_cast<const float4*>(tile + (bc & 3) * 16);
    const int4* lb4 = reinterpret_cast<const int4*>(tile + 64 + (bc & 3) * 16);
    float pm[4];                                     // group-of-4 minima
    if (p.norm) {                                    // warp-uniform
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 s4 = sq4[q];
        const int4 l4 = lb4[q];
        // padding columns: g = 0, scale = +inf -> NaN, which fminf drops
        const float v0 = (l4.x == s.lab_i) ? INFINITY : __uint_as_float(acc[4 * q + 0]) * s4.x;
        const float v1 = (l4.y == s.lab_i) ? INFINITY : __uint_as_float(acc[4 * q + 1]) * s4.y;
        const float v2 = (l4.z == s.lab_i) ? INFINITY : __uint_as_float(acc[4 * q + 2]) * s4.z;
        const float v3 = (l4.w == s.lab_i) ? INFINITY : __uint_as_float(acc[4 * q + 3]) * s4.w;
        pm[q] = fminf(fminf(fminf(v0, v1), v2), v3);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 s4 = sq4[q];
        const int4 l4 = lb4[q];
        const float v0 = fmaf(-2.f, __uint_as_float(acc[4 * q + 0]), (l4.x == s.lab_i) ? INFINITY : s4.x);
        const float v1 = fmaf(-2.f, __uint_as_float(acc[4 * q + 1]), (l4.y == s.lab_i) ? INFINITY : s4.y);
        const float v2 = fmaf(-2.f, __uint_as_float(acc[4 * q + 2]), (l4.z == s.lab_i) ? INFINITY : s4.z);
        const float v3 = fmaf(-2.f, __uint_as_float(acc[4 * q + 3]), (l4.w == s.lab_i) ? INFINITY : s4.w);
        pm[q] = fminf(fminf(fminf(v0, v1), v2), v3);                 // FMNMX3 + FMNMX
      }
    }
    const float m = fminf(fminf(fminf(pm[0], pm[1]), pm[2]), pm[3]);
    // ties go to the lower column whatever the visiting order (the multi-GPU schedule is block-major and serpentine)
    if (m < s.nbest || (m == s.nbest && col0 < s.ngrp)) {
      int q = 3;
      q = (pm[2] == m) ? 2 : q;
      q = (pm[1] == m) ? 1 : q;
      q = (pm[0] == m) ? 0 : q;
      s.nbest = m;
      s.ngrp = col0 + 4 * q;
    }
  }
  static __device__ __forceinline__ void row_end(const Params& p, const tc::GramParams&, int row, int, State& s) {
    if (s.valid && s.ngrp >= 0) atomicMin(&p.kneg[row], key_neg(s.nbest, s.ngrp));
  }
};

// ------------------------------------------------------------------------------------------ mining, fp32 CUDA-core path
// 64 x 64 tile per step, 16 x 16 threads, 4 x 4 micro-tile; grid = (column splits, row blocks).
constexpr int FM = 64, FN = 64, FK = 16;
__global__ void __launch_bounds__(256)
bh_mine_f32_kernel(const float* __restrict__ X, const float* __restrict__ sq, const int32_t* __restrict__ labels,
                   int B_all, int row_offset, int B_local, int D, unsigned long long* __restrict__ kpos,
                   unsigned long long* __restrict__ kneg) {
  __shared__ float As[FK][FM + 4];
  __shared__ float Bs[FK][FN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int rb = blockIdx.y;
  const int CT = (B_all + FN - 1) / FN;
  const int ct0 = (int)((long)blockIdx.x * CT / gridDim.x), ct1 = (int)((long)(blockIdx.x + 1) * CT / gridDim.x);
  const int arow0 = rb * FM;                                    // local anchor index of tile row 0
  int lab_i[4], gi[4];
  float pbest[4], nbest[4];
  int pidx[4], nidx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int li = arow0 + ty * 4 + i;
    gi[i] = row_offset + li;
    lab_i[i] = li < B_local ? labels[gi[i]] : 0;
    pbest[i] = -INFINITY; nbest[i] = INFINITY; pidx[i] = -1; nidx[i] = -1;
  }
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;  // loader: row lr, k offset lk..lk+3
  for (int ct = ct0; ct < ct1; ++ct) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += FK) {
      {
        const int ar = arow0 + lr;                               // local anchor
        const int br = ct * FN + lr;                             // column (global row)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = k0 + lk + q;
          As[lk + q][lr] = (ar < B_local && k < D) ? X[(long)(row_offset + ar) * D + k] : 0.f;
          Bs[lk + q][lr] = (br < B_all && k < D) ? X[(long)br * D + k] : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = ct * FN + tx * 4 + j;
      if (col < B_all) {
        const float sj = sq[col];
        const int lj = labels[col];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = fmaf(-2.f, acc[i][j], sj);
          if (lj == lab_i[i]) {
            if (col != gi[i] && v > pbest[i]) { pbest[i] = v; pidx[i] = col; }
          } else if (v < nbest[i]) { nbest[i] = v; nidx[i] = col; }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int li = arow0 + ty * 4 + i;
    if (li < B_local) {
      if (pidx[i] >= 0) atomicMax(&kpos[li], key_pos(pbest[i], pidx[i]));
      if (nidx[i] >= 0) atomicMin(&kneg[li], key_neg(nbest[i], nidx[i]));
    }
  }
}

// ------------------------------------------------------------------------------------------ finalize
// Multi-GPU step (see efm_bh_dist_fwd_bwd): what bh_finalize waits for, where it publishes an anchor's mining result and
// whom it tells.  Results stay in LOCAL memory (this rank's exchange region, where the peers' chain kernels read them over
// NVLink): the kernel itself causes no remote traffic at all.
struct DistFin {
  int world, rank;                              // world = 0: single-GPU call, nothing of this applies
  const unsigned long long* wait_b;             // [world] local arrival counters of exchange stage B (the original rows read below)
  unsigned long long target_b;
  int32_t* rip; int32_t* rin; float* rw;        // (B_all) result arrays of this step in the own region, indexed by global row
  unsigned long long* done;                     // LOCAL counter in the own region: +1 per finished block; the peers' chain kernels
                                                // poll it over NVLink (a system-scope fence + remote atomic per block costs tens of
                                                // microseconds per block while the links are busy)
};
// "how far is mining row b from mining row a", warp-cooperative, fp32 on the bf16 mining rows: squared distance by direct
// differences (plain), or -(a . b) / |e_b| (cosine metric: `scale_b` = -1/|e_b|; the anchor's own norm is a common factor)
__device__ __forceinline__ float xm_far(const __nv_bfloat16* a, const __nv_bfloat16* b, int Dp, int lane, int normalize, float scale_b) {
  float s = 0.f;
  for (int d = lane * 8; d < Dp; d += 256) {
    float x[8], y[8];
    load8<__nv_bfloat16>(a + d, x);
    load8<__nv_bfloat16>(b + d, y);
    if (normalize) {
#pragma unroll
      for (int k = 0; k < 8; ++k) s = fmaf(x[k], y[k], s);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float t = x[k] - y[k]; s = fmaf(t, t, s); }
    }
  }
  s = warp_sum(s);
  return normalize ? s * scale_b : s;
}

// One warp per anchor.  mode 0: kpos/kneg hold exact indices (fp32 mining kernel).  mode 1 (tensor-core path):
// hardest positive from the anchor's label chain, hardest negative = best non-same-label column inside the
// group of 4 the Gram kernel selected; both by exact fp32 distances on the bf16 mining rows.  mode 2: indices given.
template <typename T, bool VEC8>
__global__ void __launch_bounds__(WARPS * 32, 6)
bh_finalize_kernel(const T* __restrict__ E, const float* __restrict__ nrm, int normalize, long B_all, int D,
                   long row_offset, long B_local, float margin, const float* __restrict__ head_grad, int detach,
                   int mode, const unsigned long long* __restrict__ kpos, const unsigned long long* __restrict__ kneg,
                   const __nv_bfloat16* __restrict__ xm, int Dp, const float* __restrict__ cs, const int32_t* __restrict__ lab,
                   const int32_t* __restrict__ hcnt, const int32_t* __restrict__ hmem,
                   const int32_t* __restrict__ hhead, const int32_t* __restrict__ hnext,
                   const int32_t* __restrict__ hslot, float* __restrict__ loss, int32_t* __restrict__ idx_p,
                   int32_t* __restrict__ idx_n, float* __restrict__ wgt, int32_t* __restrict__ rheadp,
                   int32_t* __restrict__ rheadn, int32_t* __restrict__ rnextp, int32_t* __restrict__ rnextn,
                   const DistFin df) {
  pdl_trigger();
  pdl_wait();                                                        // the mining kernel has completed
  if (df.world > 1 && df.wait_b) {                                   // the peers' original rows (stage B) must have landed
    if ((int)threadIdx.x < df.world && (int)threadIdx.x != df.rank) tc::wait_counter_thread(df.wait_b + threadIdx.x, df.target_b);
    __syncthreads();
  }
  const long li = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (li < B_local) {
  const int lane = threadIdx.x & 31;
  const long gi = row_offset + li;
  int ip = -1, in = -1;
  if (mode == 0) {
    const unsigned long long kp = kpos[li], kn = kneg[li];
    ip = (kp == KEY_POS_EMPTY) ? -1 : (int)(0xffffffffu - (uint32_t)(kp & 0xffffffffull));
    in = (kn == KEY_NEG_EMPTY) ? -1 : (int)(uint32_t)(kn & 0xffffffffull);
  } else if (mode == 1) {
    // Candidates: up to 7 inline same-label members (positives) + the 4 columns of the negative group.  They are
    // compacted and spread over the warp (2/4/8/16 lanes each), so every lane sums an equal share of one candidate
    // row and all distances cost one memory round trip.
    const __nv_bfloat16* xi = xm + gi * Dp;
    const int my = lab[gi];
    const int sl = hslot[gi];
    const unsigned long long kn = kneg[li];
    const int g0 = (kn == KEY_NEG_EMPTY) ? -1 : (int)(uint32_t)(kn & 0xffffffffull);
    // second round trip: member count, inline members (read unconditionally, masked by the count afterwards) and the
    // labels of the group's columns all depend on the first one only
    int mine = -1;                                                   // lane l < 8: member l; lanes 8-11: group column
    int clab = my;
    const int c = g0 + lane - 8;
    if (lane < 8) mine = hmem[sl * 8 + lane];
    else if (lane < 12 && g0 >= 0 && c < B_all) clab = lab[c];
    const int cnt = hcnt[sl] + 1;
    if (lane < 8) { if (lane >= cnt || mine == gi) mine = -1; }
    else if (clab != my) mine = c;
    const unsigned vmask = __ballot_sync(0xffffffffu, mine >= 0);
    const int ncand = __popc(vmask);
    int lanes_per = ncand <= 2 ? 16 : (ncand <= 4 ? 8 : (ncand <= 8 ? 4 : 2));
    while (lanes_per > 2 && Dp / lanes_per < 8) lanes_per >>= 1;     // at least one 8-element slab per lane
    const int slot = lane / lanes_per, part = lane % lanes_per;
    // source lane of the slot-th valid candidate
    int src = -1;
    { unsigned m = vmask; for (int k = 0; k < slot && m; ++k) m &= m - 1; if (m && slot < ncand) src = __ffs(m) - 1; }
    const int cand = __shfl_sync(0xffffffffu, mine, src < 0 ? 0 : src);
    const bool is_pos = src >= 0 && src < 8;
    // "how far" score of the candidate: squared distance of the mining rows (plain) or -(x_i . x_c) / |e_c| (cosine metric:
    // cs = -1/|e|, the anchor's norm is a common positive factor) -- larger = farther in both cases
    float d = 0.f, csc = 1.f;
    if (src >= 0) {
      const __nv_bfloat16* xc = xm + (long)cand * Dp;
      if (normalize) csc = cs[cand];
      // the lanes of one candidate read adjacent 16-byte pieces: every 32-byte sector is requested once
      if (normalize) {
        for (int e = part * 8; e < Dp; e += lanes_per * 8) {
          float x[8], y[8];
          load8<__nv_bfloat16>(xi + e, x);
          load8<__nv_bfloat16>(xc + e, y);
#pragma unroll
          for (int k = 0; k < 8; ++k) d = fmaf(x[k], y[k], d);
        }
      } else {
        for (int e = part * 8; e < Dp; e += lanes_per * 8) {
          float x[8], y[8];
          load8<__nv_bfloat16>(xi + e, x);
          load8<__nv_bfloat16>(xc + e, y);
#pragma unroll
          for (int k = 0; k < 8; ++k) { const float t = x[k] - y[k]; d = fmaf(t, t, d); }
        }
      }
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { const float od = __shfl_xor_sync(0xffffffffu, d, o); if (o < lanes_per) d += od; }
    if (normalize) d *= csc;
    // hardest positive: max d (lowest index on ties); hardest negative: min d (lowest index on ties)
    float pd = (src >= 0 && is_pos) ? d : -INFINITY; int pi = (src >= 0 && is_pos) ? cand : -1;
    float nd = (src >= 0 && !is_pos) ? d : INFINITY; int ni = (src >= 0 && !is_pos) ? cand : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float opd = __shfl_xor_sync(0xffffffffu, pd, o); const int opi = __shfl_xor_sync(0xffffffffu, pi, o);
      if (opi >= 0 && (pi < 0 || opd > pd || (opd == pd && opi < pi))) { pd = opd; pi = opi; }
      const float ond = __shfl_xor_sync(0xffffffffu, nd, o); const int oni = __shfl_xor_sync(0xffffffffu, ni, o);
      if (oni >= 0 && (ni < 0 || ond < nd || (ond == nd && oni < ni))) { nd = ond; ni = oni; }
    }
    ip = pi; in = ni;
    if (cnt > 8) {                                                   // rare: labels with more than 8 rows
      for (int j = hhead[sl]; j >= 0; j = hnext[j]) {
        if (j == gi) continue;
        const float dj = xm_far(xi, xm + (long)j * Dp, Dp, lane, normalize, normalize ? cs[j] : 1.f);
        if (ip < 0 || dj > pd || (dj == pd && j < ip)) { pd = dj; ip = j; }
      }
    }
  } else {
    ip = idx_p[li]; in = idx_n[li];
  }
  if (mode != 2 && lane == 0) { idx_p[li] = ip; idx_n[li] = in; }
  float l = 0.f;
  if (!(ip < 0 || in < 0 || ip >= B_all || in >= B_all)) {          // else: no valid triplet, loss 0, no gradient
  const T* ea = E + gi * D; const T* epp = E + (long)ip * D; const T* en = E + (long)in * D;
  // x = e * (1/|e|): one reciprocal per row instead of a division per element (<= 1 ulp from e/|e|)
  const float ra = normalize ? 1.f / nrm[gi] : 1.f, rp = normalize ? 1.f / nrm[ip] : 1.f, rn = normalize ? 1.f / nrm[in] : 1.f;
  float s = 0.f;
  if constexpr (VEC8) {
    for (int d = lane * 8; d < D; d += 256) {
      float a[8], p[8], n[8];
      load8<T>(ea + d, a); load8<T>(epp + d, p); load8<T>(en + d, n);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float av = a[k] * ra;
        const float dp = fmaf(p[k], rp, -av), dn = fmaf(n[k], rn, -av);
        s += dp * dp - dn * dn;
      }
    }
  } else {
    for (int d = lane; d < D; d += 32) {
      const float a = El<T>::to_f(ea[d]) * ra;
      const float dp = fmaf(El<T>::to_f(epp[d]), rp, -a), dn = fmaf(El<T>::to_f(en[d]), rn, -a);
      s += dp * dp - dn * dn;
    }
  }
  s = warp_sum(s);
  l = fmaxf(s + margin, 0.f);
  }
  if (lane == 0) {
    if (loss) loss[li] = l;
    // gradient bookkeeping: weight of this anchor and membership in the chains of its positive / negative row
    const float w2 = (l > 0.f) ? 2.f * (head_grad ? head_grad[li] : 1.f) : 0.f;       // relu'(0) = 0
    if (wgt) wgt[li] = w2;
    if (wgt && rheadp && w2 != 0.f) {
      const bool p_ok = !detach || (ip >= row_offset && ip < row_offset + B_local);
      const bool n_ok = !detach || (in >= row_offset && in < row_offset + B_local);
      if (p_ok) rnextp[li] = atomicExch(&rheadp[ip], (int32_t)li);
      if (n_ok) rnextn[li] = atomicExch(&rheadn[in], (int32_t)li);
    }
    // multi-GPU exact gradient: (idx_p, idx_n, weight) of every anchor is what the other ranks need from this one
    // (12 bytes per anchor instead of a dense reduce-scatter of dE); they fetch it from here once this block has signalled
    if (df.world > 0) { df.rip[gi] = ip; df.rin[gi] = in; df.rw[gi] = w2; }
  }
  }
  if (df.world > 1) {                                              // results of this block are in L2: count the block as done
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(df.done, 1ull);
  }
}

// ------------------------------------------------------------------------------------------ gradient (gather form)
// One warp per row r of E_all: dX_r = [r is an active anchor] w_r (x_n - x_p) + sum_{j: p(j)=r} w_j (x_r - x_j)
// + sum_{j: n(j)=r} w_j (x_j - x_r), then (normalize) dE_r = (dX_r - x_r (x_r . dX_r)) / |e_r|.  Every output row is
// written exactly once with plain stores: no zero-fill, no floating-point atomics, no separate normalise-backward pass.
template <typename T, int VEC>
__device__ __forceinline__ void load_slab(const T* p, int d, int D, float rinv, float (&o)[VEC]) {
  if constexpr (VEC == 8) {
    if (d < D) {
      load8<T>(p + d, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = o[k] * rinv;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
    }
  } else {
    o[0] = d < D ? El<T>::to_f(p[d]) * rinv : 0.f;
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(WARPS * 32, 6)
bh_grad_kernel(const T* __restrict__ E, const float* __restrict__ nrm, int normalize, long B_all, int D,
               long row_offset, long B_local, const int32_t* __restrict__ idx_p, const int32_t* __restrict__ idx_n,
               const float* __restrict__ wgt, const int32_t* __restrict__ rheadp, const int32_t* __restrict__ rheadn,
               const int32_t* __restrict__ rnextp, const int32_t* __restrict__ rnextn, float* __restrict__ dE,
               long row_begin) {
  pdl_wait();
  // rows [row_begin, row_begin + B_all) are processed (B_all = number of output rows); anchors are rows
  // [row_offset, row_offset + B_local) and index idx_p / idx_n / wgt / rnext*; rhead* and dE are indexed by the
  // processed row (r - row_begin).  Fused single-call form: row_begin = 0, B_all = all rows.  Row-sliced form
  // (efm_bh_grad_rows): row_begin = first local row, anchors = every row of the gathered batch.
  const long rl = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (rl >= B_all) return;
  const long r = row_begin + rl;
  dE += -row_begin * D;                                           // dE[r * D] addresses output row r - row_begin
  rheadp -= row_begin; rheadn -= row_begin;
  const int lane = threadIdx.x & 31;
  const T* er = E + r * D;
  const float nr = normalize ? 1.f / nrm[r] : 1.f;              // reciprocal norm of this row
  const long li = r - row_offset;
  const bool anchor = li >= 0 && li < B_local;
  const float wa = anchor ? wgt[li] : 0.f;
  const int ap = wa != 0.f ? idx_p[li] : -1, an = wa != 0.f ? idx_n[li] : -1;
  const int hp = rheadp[r], hn = rheadn[r];
  constexpr int SLAB = 32 * VEC;
  const bool single = D <= SLAB;
  float dot = 0.f;
  float acc[VEC], x[VEC];
  for (int d0 = 0; d0 < D; d0 += SLAB) {
    const int d = d0 + lane * VEC;
    load_slab<T, VEC>(er, d, D, nr, x);
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    if (wa != 0.f) {
      float p[VEC], n[VEC];
      load_slab<T, VEC>(E + (long)ap * D, d, D, normalize ? 1.f / nrm[ap] : 1.f, p);
      load_slab<T, VEC>(E + (long)an * D, d, D, normalize ? 1.f / nrm[an] : 1.f, n);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = wa * (n[k] - p[k]);
    }
    for (int j = hp; j >= 0; j = rnextp[j]) {                        // r is the hardest positive of anchor j
      const long a = row_offset + j;
      float av[VEC];
      load_slab<T, VEC>(E + a * D, d, D, normalize ? 1.f / nrm[a] : 1.f, av);
      const float w = wgt[j];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w, x[k] - av[k], acc[k]);
    }
    for (int j = hn; j >= 0; j = rnextn[j]) {                        // r is the hardest negative of anchor j
      const long a = row_offset + j;
      float av[VEC];
      load_slab<T, VEC>(E + a * D, d, D, normalize ? 1.f / nrm[a] : 1.f, av);
      const float w = wgt[j];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w, av[k] - x[k], acc[k]);
    }
    if (normalize) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) dot = fmaf(x[k], acc[k], dot);
    }
    if (!(single && normalize) && d < D) {                           // raw dX (final if not normalising)
      if constexpr (VEC == 8) {
        float4* o = reinterpret_cast<float4*>(dE + r * D + d);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]); o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      } else {
        dE[r * D + d] = acc[0];
      }
    }
  }
  if (!normalize) return;
  dot = warp_sum(dot);
  if (single) {
    const int d = lane * VEC;
    if (d < D) {
      if constexpr (VEC == 8) {
        float4* o = reinterpret_cast<float4*>(dE + r * D + d);
        o[0] = make_float4((acc[0] - x[0] * dot) * nr, (acc[1] - x[1] * dot) * nr, (acc[2] - x[2] * dot) * nr, (acc[3] - x[3] * dot) * nr);
        o[1] = make_float4((acc[4] - x[4] * dot) * nr, (acc[5] - x[5] * dot) * nr, (acc[6] - x[6] * dot) * nr, (acc[7] - x[7] * dot) * nr);
      } else {
        dE[r * D + d] = (acc[0] - x[0] * dot) * nr;
      }
    }
    return;
  }
  for (int d0 = 0; d0 < D; d0 += SLAB) {                             // second pass over the raw dX row
    const int d = d0 + lane * VEC;
    if (d >= D) continue;
    load_slab<T, VEC>(er, d, D, nr, x);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float* o = dE + r * D + d + k;
      *o = (*o - x[k] * dot) * nr;
    }
  }
}

}  // namespace
}  // namespace efm

using namespace efm;

// Profiling aid: stage-boundary events of the batch-hard step (see efm_bh_profile_events in efmtl.h).
// Boundary i is recorded after: 0 nothing (step start), 1 memset + prep, 2 Gram + mining (+ exchange), 3 (unused),
// 4 finalize, 5 chains (multi-GPU exact gradient only), 6 gradient.
static thread_local cudaEvent_t g_prof_ev[EFM_BH_PROFILE_EVENTS] = {};
static thread_local int g_prof_n = 0;
extern "C" int efm_bh_profile_events(void* const* events, int n) {
  EFM_REQUIRE(n >= 0 && n <= EFM_BH_PROFILE_EVENTS && (n == 0 || events), EFM_E_ARG, "efm_bh_profile_events: n=%d", n);
  for (int i = 0; i < n; ++i) g_prof_ev[i] = static_cast<cudaEvent_t>(events[i]);
  g_prof_n = n;
  return EFM_OK;
}
#define EFM_BH_STAGE(i, st)                                                            \
  do {                                                                                 \
    if ((i) < g_prof_n && g_prof_ev[i]) EFM_CUDA(cudaEventRecord(g_prof_ev[i], (st))); \
  } while (0)

// The bf16 mining rows of the tensor-core path: the caller's rows themselves when they already are bf16 with K a multiple
// of 64 (nothing is copied), else the rounded / zero-padded copy bh_prep writes into the workspace.
static bool bh_mine_in_place(const BhLayout& L, const void* E_all, long D, int dtype) {
  return L.use_tc && dtype == EFM_BF16 && D == L.Dp && aligned(E_all, 16);
}

// true when the step's Gram kernel is the variant with the extra warps (CTA pairs): they build the per-label member table
// while the tiles run, so bh_prep does not have to
static bool bh_table_in_gram(const BhLayout& L, long B_local) {
  if (!L.use_tc) return false;
  const int mh = tc::pick_mh((int)B_local, L.Dp / tc::BK);
  return tc::pick_cg((int)B_local, mh, 0) == 2;
}

static int bh_launch_mining(const BhLayout& L, const void* E_all, uint8_t* ws, long B_all, long row_offset, long B_local,
                            long D, int dtype, int normalize, cudaStream_t st, bool pdl = false, int build_table = 0) {
  // build_table: 0 = plain kernel, 1 = variant with the extra warps, which build the label table, 2 = same variant, table
  // left alone (efm_bh_mine: re-registering the rows of a populated table would corrupt its overflow chains)
  auto* kpos = reinterpret_cast<unsigned long long*>(ws + L.kpos);
  auto* kneg = reinterpret_cast<unsigned long long*>(ws + L.kneg);
  float* sq = reinterpret_cast<float*>(ws + L.sq);
  int32_t* lab = reinterpret_cast<int32_t*>(ws + L.lab);
  const auto* xm_bf16 = bh_mine_in_place(L, E_all, D, dtype) ? static_cast<const __nv_bfloat16*>(E_all)
                                                              : reinterpret_cast<const __nv_bfloat16*>(ws + L.xm);
  float* xm_f32 = reinterpret_cast<float*>(ws + L.xm);
  if (L.use_tc) {
    tc::GramParams g{};
    g.M = (int)B_local; g.N = (int)B_all; g.nk = L.Dp / tc::BK;
    const int mh = tc::pick_mh(g.M, g.nk);
    const int cg = tc::pick_cg(g.M, mh, 0);
    if (const char* e = getenv("EFM_GRAM_PHASES")) {                 // test hook: the multi-GPU block-major column schedule
      const int ph = atoi(e), ct = (int)cdiv(B_all, 256 / mh);       // (no waits: every block is local)
      if (ph > 1 && ct % ph == 0) { g.phases = ph; g.phase0 = (int)((row_offset * ph) / B_all) % ph; }
    }
    const int grid = tc::plan_gram(g, mh, 0, MineEpi::SMEM_PER_WARP * tc::EPI_WARPS, cg);
    MineEpi::Params ep{sq, lab, kneg, (int)row_offset, normalize};
    if (build_table && cg == 2) {                                    // single-GPU use of the exchange warps: the label table only
      EFM_REQUIRE(g.stages >= 1, EFM_E_UNSUPPORTED, "gram: K = %d leaves no shared memory for the B pipeline", g.nk * tc::BK);
      tc::CommParams cp{};
      cp.world = 1;
      cp.tab_lab = build_table == 1 ? lab : nullptr; cp.tab_rows = B_all;
      cp.hkeys = reinterpret_cast<unsigned long long*>(ws + L.hkeys); cp.hcnt = reinterpret_cast<int32_t*>(ws + L.hcnt);
      cp.hmem = reinterpret_cast<int32_t*>(ws + L.hmem); cp.hhead = reinterpret_cast<int32_t*>(ws + L.hhead);
      cp.hnext = reinterpret_cast<int32_t*>(ws + L.hnext); cp.hslot = reinterpret_cast<int32_t*>(ws + L.hslot); cp.tbits = L.tbits;
      return tc::launch_gram_mh<MineEpi, 1, 2, true>(xm_bf16 + row_offset * L.Dp, xm_bf16, g, ep, grid, st, &cp, pdl);
    }
    return tc::launch_gram<MineEpi>(xm_bf16 + row_offset * L.Dp, xm_bf16, mh, g, ep, grid, st, cg, pdl);
  }
  const float* X = L.xm_f32 ? xm_f32 : static_cast<const float*>(E_all);
  const int row_blocks = (int)cdiv(B_local, FM);
  int splits = (int)cdiv(4L * num_sms(), row_blocks);
  const int CT = (int)cdiv(B_all, FN);
  if (splits > CT) splits = CT;
  if (splits < 1) splits = 1;
  bh_mine_f32_kernel<<<dim3(splits, row_blocks), 256, 0, st>>>(X, sq, lab, (int)B_all, (int)row_offset, (int)B_local,
                                                                (int)D, kpos, kneg);
  EFM_LAUNCH_CHECK("bh_mine_f32");
  return EFM_OK;
}

static bool bh_vec8(const void* E, long D, const float* dE) {
  return (D % 8 == 0) && aligned(E, 16) && (!dE || aligned(dE, 16));
}

#define EFM_BH_DISPATCH(dtype, vec, T, V, ...)                                              \
  EFM_DISPATCH_DTYPE(dtype, T, {                                                            \
    if (vec) { constexpr bool V = true; __VA_ARGS__; } else { constexpr bool V = false; __VA_ARGS__; } \
  })

extern "C" size_t efm_bh_workspace_bytes(long B_local, long B_all, long D, int dtype, int flags) {
  if (B_local < 0 || B_all <= 0 || D <= 0) return 0;
  return bh_layout(B_local, B_all, D, dtype, flags).total;
}

extern "C" int efm_bh_triplet_fwd_bwd(const void* E_all, const int32_t* labels_all, long B_all, long row_offset,
                                      long B_local, long D, int dtype, float margin, const float* head_grad, int flags,
                                      float* loss, int32_t* idx_p, int32_t* idx_n, float* dE_all, void* workspace,
                                      size_t ws_bytes, void* stream) {
  EFM_REQUIRE(B_all > 0 && B_local >= 0 && D > 0 && row_offset >= 0 && row_offset + B_local <= B_all, EFM_E_ARG,
              "efm_bh_triplet_fwd_bwd: bad shape B_all=%ld row_offset=%ld B_local=%ld D=%ld", B_all, row_offset, B_local, D);
  EFM_REQUIRE(B_all < (1L << 30) && D < (1 << 20), EFM_E_ARG, "efm_bh_triplet_fwd_bwd: shape too large");
  EFM_REQUIRE(E_all && labels_all && loss && idx_p && idx_n && workspace, EFM_E_ARG, "efm_bh_triplet_fwd_bwd: null pointer");
  EFM_REQUIRE(dtype == EFM_F32 || dtype == EFM_BF16 || dtype == EFM_F16, EFM_E_ARG, "efm_bh_triplet_fwd_bwd: dtype %d", dtype);
  EFM_REQUIRE(aligned(workspace, 256), EFM_E_ALIGN, "efm_bh_triplet_fwd_bwd: workspace must be 256-byte aligned");
  const BhLayout L = bh_layout(B_local, B_all, D, dtype, flags);
  EFM_REQUIRE(ws_bytes >= L.total, EFM_E_WORKSPACE, "efm_bh_triplet_fwd_bwd: workspace %zu < required %zu", ws_bytes, L.total);
  if (B_local == 0) {
    if (dE_all) EFM_CUDA(cudaMemsetAsync(dE_all, 0, (size_t)B_all * D * 4, static_cast<cudaStream_t>(stream)));
    return EFM_OK;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto* kpos = reinterpret_cast<unsigned long long*>(ws + L.kpos);
  auto* kneg = reinterpret_cast<unsigned long long*>(ws + L.kneg);
  float* sq = reinterpret_cast<float*>(ws + L.sq);
  float* nrm = reinterpret_cast<float*>(ws + L.nrm);
  int32_t* lab = reinterpret_cast<int32_t*>(ws + L.lab);
  auto* xm_bf16 = reinterpret_cast<__nv_bfloat16*>(ws + L.xm);
  float* xm_f32 = reinterpret_cast<float*>(ws + L.xm);
  auto* hkeys = L.use_tc ? reinterpret_cast<unsigned long long*>(ws + L.hkeys) : nullptr;
  int32_t* hcnt = reinterpret_cast<int32_t*>(ws + L.hcnt);
  int32_t* hmem = reinterpret_cast<int32_t*>(ws + L.hmem);
  int32_t* hhead = reinterpret_cast<int32_t*>(ws + L.hhead);
  int32_t* hnext = reinterpret_cast<int32_t*>(ws + L.hnext);
  int32_t* hslot = reinterpret_cast<int32_t*>(ws + L.hslot);
  int32_t* rheadp = reinterpret_cast<int32_t*>(ws + L.rheadp);
  int32_t* rheadn = reinterpret_cast<int32_t*>(ws + L.rheadn);
  int32_t* rnextp = reinterpret_cast<int32_t*>(ws + L.rnextp);
  int32_t* rnextn = reinterpret_cast<int32_t*>(ws + L.rnextn);
  float* wgt = dE_all ? reinterpret_cast<float*>(ws + L.wgt) : nullptr;
  const int normalize = (flags & EFM_BH_NORMALIZE) ? 1 : 0;
  const int detach = (flags & EFM_BH_DETACH_GATHERED) ? 1 : 0;
  const unsigned grid_all = (unsigned)cdiv(B_all, WARPS), grid_loc = (unsigned)cdiv(B_local, WARPS);
  const bool vec = bh_vec8(E_all, D, dE_all);

  EFM_BH_STAGE(0, st);
  const bool in_place = bh_mine_in_place(L, E_all, D, dtype);
  const __nv_bfloat16* xm_rows = in_place ? static_cast<const __nv_bfloat16*>(E_all) : xm_bf16;
  const bool table_in_gram = bh_table_in_gram(L, B_local);
  // empty label table + empty chains: by the prep kernel itself when it does not insert into the table, else a memset first
  if (!table_in_gram) EFM_CUDA(cudaMemsetAsync(ws + L.clr_begin, 0xFF, L.clr_bytes, st));
  uint4* clr = table_in_gram ? reinterpret_cast<uint4*>(ws + L.clr_begin) : nullptr;
  const long clr_n16 = table_in_gram ? (long)(L.clr_bytes / 16) : 0;
  EFM_BH_DISPATCH(dtype, vec, T, V, {
    bh_prep_kernel<T, V><<<grid_all, WARPS * 32, 0, st>>>((const T*)E_all, labels_all, B_all, (int)D, L.Dp, normalize,
                                                          L.use_tc ? 1 : 0, L.xm_f32 ? 1 : 0, row_offset, B_local, sq, nrm,
                                                          lab, in_place ? nullptr : xm_bf16, xm_f32, kpos, kneg, nullptr,
                                                          table_in_gram ? nullptr : hkeys, hcnt, hmem, hhead, hnext, hslot, L.tbits,
                                                          clr, clr_n16);
  });
  EFM_LAUNCH_CHECK("bh_prep");
  EFM_BH_STAGE(1, st);
  {
    int rc = bh_launch_mining(L, E_all, ws, B_all, row_offset, B_local, D, dtype, normalize, st, true, table_in_gram ? 1 : 0);
    if (rc) return rc;
  }
  EFM_BH_STAGE(2, st);
  EFM_BH_DISPATCH(dtype, vec, T, V, {
    EFM_LAUNCH((bh_finalize_kernel<T, V>), grid_loc, WARPS * 32, 0, st, true, (const T*)E_all, nrm, normalize, B_all, (int)D, row_offset,
               B_local, margin, head_grad, detach, L.use_tc ? 1 : 0, kpos, kneg, xm_rows, L.Dp, sq, lab, hcnt, hmem, hhead, hnext, hslot,
               loss, idx_p, idx_n, wgt, rheadp, rheadn, rnextp, rnextn, DistFin{});
  });
  EFM_LAUNCH_CHECK("bh_finalize");
  EFM_BH_STAGE(4, st);
  if (dE_all) {
    EFM_DISPATCH_DTYPE(dtype, T, {
      if (vec) EFM_LAUNCH((bh_grad_kernel<T, 8>), grid_all, WARPS * 32, 0, st, true, (const T*)E_all, nrm, normalize, B_all, (int)D,
                          row_offset, B_local, idx_p, idx_n, wgt, rheadp, rheadn, rnextp, rnextn, dE_all, 0L);
      else EFM_LAUNCH((bh_grad_kernel<T, 1>), grid_all, WARPS * 32, 0, st, true, (const T*)E_all, nrm, normalize, B_all, (int)D,
                      row_offset, B_local, idx_p, idx_n, wgt, rheadp, rheadn, rnextp, rnextn, dE_all, 0L);
    });
    EFM_LAUNCH_CHECK("bh_grad");
    EFM_BH_STAGE(6, st);
  }
  return EFM_OK;
}

// Backward for given mined indices (head_grad per anchor): used when the loss vector was returned
// un-reduced and autograd later supplies d(total)/d(loss_i).
struct BwdLayout { size_t nrm, sq, rheadp, rheadn, rnextp, rnextn, wgt, total; };
static BwdLayout bwd_layout(long B_all) {
  BwdLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t n = (size_t)(B_all > 0 ? B_all : 0) * 4;
  L.nrm = take(n); L.sq = take(n); L.rheadp = take(n); L.rheadn = take(n); L.rnextp = take(n); L.rnextn = take(n); L.wgt = take(n);
  L.total = off;
  return L;
}
extern "C" size_t efm_bh_backward_workspace_bytes(long B_all) { return bwd_layout(B_all).total; }

extern "C" int efm_bh_backward(const void* E_all, long B_all, long row_offset, long B_local, long D, int dtype,
                               float margin, const float* head_grad, int flags, const int32_t* idx_p,
                               const int32_t* idx_n, float* dE_all, void* workspace, size_t ws_bytes, void* stream) {
  EFM_REQUIRE(B_all > 0 && B_local >= 0 && D > 0 && row_offset >= 0 && row_offset + B_local <= B_all, EFM_E_ARG,
              "efm_bh_backward: bad shape B_all=%ld row_offset=%ld B_local=%ld D=%ld", B_all, row_offset, B_local, D);
  EFM_REQUIRE(E_all && idx_p && idx_n && dE_all && workspace, EFM_E_ARG, "efm_bh_backward: null pointer");
  EFM_REQUIRE(aligned(workspace, 256), EFM_E_ALIGN, "efm_bh_backward: workspace must be 256-byte aligned");
  const BwdLayout L = bwd_layout(B_all);
  EFM_REQUIRE(ws_bytes >= L.total, EFM_E_WORKSPACE, "efm_bh_backward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* nrm = reinterpret_cast<float*>(ws + L.nrm);
  float* sq = reinterpret_cast<float*>(ws + L.sq);
  int32_t* rheadp = reinterpret_cast<int32_t*>(ws + L.rheadp);
  int32_t* rheadn = reinterpret_cast<int32_t*>(ws + L.rheadn);
  int32_t* rnextp = reinterpret_cast<int32_t*>(ws + L.rnextp);
  int32_t* rnextn = reinterpret_cast<int32_t*>(ws + L.rnextn);
  float* wgt = reinterpret_cast<float*>(ws + L.wgt);
  const int normalize = (flags & EFM_BH_NORMALIZE) ? 1 : 0;
  const int detach = (flags & EFM_BH_DETACH_GATHERED) ? 1 : 0;
  const unsigned grid_all = (unsigned)cdiv(B_all, WARPS), grid_loc = (unsigned)cdiv(B_local, WARPS);
  const bool vec = bh_vec8(E_all, D, dE_all);
  EFM_CUDA(cudaMemsetAsync(ws + L.rheadp, 0xFF, L.rnextp - L.rheadp, st));
  // row norms: prep with every optional output disabled
  EFM_BH_DISPATCH(dtype, vec, T, V, {
    bh_prep_kernel<T, V><<<grid_all, WARPS * 32, 0, st>>>((const T*)E_all, nullptr, B_all, (int)D, (int)D, normalize, 0, 0, 0, 0,
                                                          sq, nrm, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                          nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0L);
  });
  EFM_LAUNCH_CHECK("bh_prep(stats)");
  if (B_local > 0) {
    EFM_BH_DISPATCH(dtype, vec, T, V, {
      bh_finalize_kernel<T, V><<<grid_loc, WARPS * 32, 0, st>>>((const T*)E_all, nrm, normalize, B_all, (int)D, row_offset,
                                                                B_local, margin, head_grad, detach, 2, nullptr, nullptr, nullptr,
                                                                0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                const_cast<int32_t*>(idx_p), const_cast<int32_t*>(idx_n), wgt,
                                                                rheadp, rheadn, rnextp, rnextn, DistFin{});
    });
    EFM_LAUNCH_CHECK("bh_finalize(bwd)");
  }
  EFM_DISPATCH_DTYPE(dtype, T, {
    if (vec) bh_grad_kernel<T, 8><<<grid_all, WARPS * 32, 0, st>>>((const T*)E_all, nrm, normalize, B_all, (int)D, row_offset,
                                                                    B_local, idx_p, idx_n, wgt, rheadp, rheadn, rnextp, rnextn, dE_all, 0);
    else bh_grad_kernel<T, 1><<<grid_all, WARPS * 32, 0, st>>>((const T*)E_all, nrm, normalize, B_all, (int)D, row_offset,
                                                                B_local, idx_p, idx_n, wgt, rheadp, rheadn, rnextp, rnextn, dE_all, 0);
  });
  EFM_LAUNCH_CHECK("bh_grad");
  return EFM_OK;
}

// ---- multi-GPU exact gradient without a dense reduce-scatter ----------------------------------------------------------
// Every rank all-gathers (idx_p, idx_n, wgt) of all anchors (12 bytes per anchor) and computes the gradient of ITS rows
// only: the anchor-role term of its own anchors plus the positive / negative-role terms of every anchor (of any rank)
// that picked one of its rows.  All embeddings needed are in the gathered batch already.
// Multi-GPU form (pull.world > 1): the results of the peers' anchors live in the PEERS' exchange regions.  Every block
// waits until the local arrival counters of all other ranks have reached `target` (their bh_finalize blocks have all
// signalled), then every thread fetches its anchor's (idx_p, idx_n, weight) from the owner over NVLink and keeps a local
// copy for the gradient kernel.
struct ChainPull {
  int world, rank;
  long rows_per_rank;
  const unsigned long long* done[tc::MAX_PEERS];   // rank r's finalize-completion counter (in r's region, read over NVLink)
  unsigned long long target;                    // its value once r's finalize of this step has finished
  unsigned long long* gate;                     // local: set to `step` by block 0 once every peer is done
  unsigned long long step;
  const int32_t* ip[tc::MAX_PEERS]; const int32_t* in[tc::MAX_PEERS]; const float* w[tc::MAX_PEERS];   // rank r's arrays as mapped here
};
__global__ void __launch_bounds__(256)
bh_chain_kernel(int32_t* idx_p, int32_t* idx_n, float* wgt,
                long n_anchors, long row_begin, long nrows, int32_t* __restrict__ rheadp, int32_t* __restrict__ rheadn,
                int32_t* __restrict__ rnextp, int32_t* __restrict__ rnextn, const ChainPull pull) {
  pdl_trigger();
  pdl_wait();
  if (pull.world > 1) {
    // Block 0 polls the peers' completion counters over NVLink (world - 1 threads, not world - 1 per block) and opens a local
    // gate; every other block spins on the gate in local L2.  Block 0 is dispatched first, so the gate always opens.
    if (blockIdx.x == 0) {
#pragma unroll
      for (int r = 0; r < tc::MAX_PEERS; ++r)
        if ((int)threadIdx.x == r && r < pull.world && r != pull.rank) tc::wait_counter_thread(pull.done[r], pull.target);
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); atomicMax(pull.gate, pull.step); }
    } else {
      if (threadIdx.x == 0) {
        const unsigned long long t0 = tc::gtimer();
        while (*reinterpret_cast<volatile unsigned long long*>(pull.gate) < pull.step) {
          __nanosleep(100);
          if (tc::gtimer() - t0 > 8000000000ull) tc::peer_timeout(pull.gate, 0ull, pull.step);
        }
        __threadfence();
      }
      __syncthreads();
    }
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_anchors) {
      const int owner = (int)(j / pull.rows_per_rank);
      if (owner != pull.rank) {
        int p = 0, n = 0; float w = 0.f;
#pragma unroll
        for (int r = 0; r < tc::MAX_PEERS; ++r)                         // static indices: the pointer arrays stay in constant memory
          if (r == owner) { p = __ldcg(pull.ip[r] + j); n = __ldcg(pull.in[r] + j); w = __ldcg(pull.w[r] + j); }
        idx_p[j] = p; idx_n[j] = n; wgt[j] = w;
        if (w != 0.f) {
          const long pl = p - row_begin, nl = n - row_begin;
          if (pl >= 0 && pl < nrows) rnextp[j] = atomicExch(&rheadp[pl], (int32_t)j);
          if (nl >= 0 && nl < nrows) rnextn[j] = atomicExch(&rheadn[nl], (int32_t)j);
        }
        return;
      }
    }
  }
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_anchors || __ldcg(wgt + j) == 0.f) return;
  const long p = __ldcg(idx_p + j) - row_begin, n = __ldcg(idx_n + j) - row_begin;
  if (p >= 0 && p < nrows) rnextp[j] = atomicExch(&rheadp[p], (int32_t)j);
  if (n >= 0 && n < nrows) rnextn[j] = atomicExch(&rheadn[n], (int32_t)j);
}

struct GradRowsLayout { size_t nrm, sq, rheadp, rheadn, rnextp, rnextn, total; };
static GradRowsLayout grad_rows_layout(long B_all, long B_local) {
  GradRowsLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.nrm = take((size_t)B_all * 4); L.sq = take((size_t)B_all * 4);
  L.rheadp = take((size_t)B_local * 4); L.rheadn = take((size_t)B_local * 4);
  L.rnextp = take((size_t)B_all * 4); L.rnextn = take((size_t)B_all * 4);
  L.total = off;
  return L;
}
extern "C" size_t efm_bh_grad_rows_workspace_bytes(long B_all, long B_local) {
  return grad_rows_layout(B_all > 0 ? B_all : 0, B_local > 0 ? B_local : 0).total;
}

extern "C" int efm_bh_grad_rows(const void* E_all, long B_all, long D, int dtype, int flags, const int32_t* idx_p_all,
                                const int32_t* idx_n_all, const float* wgt_all, long row_offset, long B_local,
                                float* dE_local, void* workspace, size_t ws_bytes, void* stream) {
  EFM_REQUIRE(B_all > 0 && B_local >= 0 && D > 0 && row_offset >= 0 && row_offset + B_local <= B_all, EFM_E_ARG,
              "efm_bh_grad_rows: bad shape B_all=%ld row_offset=%ld B_local=%ld D=%ld", B_all, row_offset, B_local, D);
  if (B_local == 0) return EFM_OK;
  EFM_REQUIRE(E_all && idx_p_all && idx_n_all && wgt_all && dE_local && workspace, EFM_E_ARG, "efm_bh_grad_rows: null pointer");
  EFM_REQUIRE(aligned(workspace, 256), EFM_E_ALIGN, "efm_bh_grad_rows: workspace must be 256-byte aligned");
  const GradRowsLayout L = grad_rows_layout(B_all, B_local);
  EFM_REQUIRE(ws_bytes >= L.total, EFM_E_WORKSPACE, "efm_bh_grad_rows: workspace %zu < required %zu", ws_bytes, L.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* nrm = reinterpret_cast<float*>(ws + L.nrm);
  float* sq = reinterpret_cast<float*>(ws + L.sq);
  int32_t* rheadp = reinterpret_cast<int32_t*>(ws + L.rheadp);
  int32_t* rheadn = reinterpret_cast<int32_t*>(ws + L.rheadn);
  int32_t* rnextp = reinterpret_cast<int32_t*>(ws + L.rnextp);
  int32_t* rnextn = reinterpret_cast<int32_t*>(ws + L.rnextn);
  const int normalize = (flags & EFM_BH_NORMALIZE) ? 1 : 0;
  const bool vec = bh_vec8(E_all, D, dE_local);
  EFM_CUDA(cudaMemsetAsync(ws + L.rheadp, 0xFF, L.rnextp - L.rheadp, st));
  if (normalize) {                                                // row norms of the whole gathered batch
    EFM_BH_DISPATCH(dtype, vec, T, V, {
      bh_prep_kernel<T, V><<<(unsigned)cdiv(B_all, WARPS), WARPS * 32, 0, st>>>((const T*)E_all, nullptr, B_all, (int)D, (int)D, 1,
                                                                                 0, 0, 0, 0, sq, nrm, nullptr, nullptr, nullptr, nullptr,
                                                                                 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                                 nullptr, nullptr, 0, nullptr, 0L);
    });
    EFM_LAUNCH_CHECK("bh_prep(stats)");
  }
  bh_chain_kernel<<<(unsigned)cdiv(B_all, 256), 256, 0, st>>>(const_cast<int32_t*>(idx_p_all), const_cast<int32_t*>(idx_n_all),
                                                                const_cast<float*>(wgt_all), B_all, row_offset, B_local, rheadp,
                                                                rheadn, rnextp, rnextn, ChainPull{});
  EFM_LAUNCH_CHECK("bh_chain");
  const unsigned grid = (unsigned)cdiv(B_local, WARPS);
  EFM_DISPATCH_DTYPE(dtype, T, {
    if (vec) bh_grad_kernel<T, 8><<<grid, WARPS * 32, 0, st>>>((const T*)E_all, nrm, normalize, B_local, (int)D, 0, B_all, idx_p_all,
                                                                idx_n_all, wgt_all, rheadp, rheadn, rnextp, rnextn, dE_local, row_offset);
    else bh_grad_kernel<T, 1><<<grid, WARPS * 32, 0, st>>>((const T*)E_all, nrm, normalize, B_local, (int)D, 0, B_all, idx_p_all,
                                                            idx_n_all, wgt_all, rheadp, rheadn, rnextp, rnextn, dE_local, row_offset);
  });
  EFM_LAUNCH_CHECK("bh_grad(rows)");
  return EFM_OK;
}

// Mining step alone on a workspace prepared by a previous efm_bh_triplet_fwd_bwd call with the same shapes
// (keys are monotone, so re-mining is idempotent).  For profiling / roofline measurement of the dominant kernel.
extern "C" int efm_bh_mine(const void* E_all, long B_all, long row_offset, long B_local, long D, int dtype, int flags,
                           void* workspace, size_t ws_bytes, void* stream) {
  EFM_REQUIRE(B_all > 0 && B_local > 0 && D > 0 && row_offset >= 0 && row_offset + B_local <= B_all, EFM_E_ARG,
              "efm_bh_mine: bad shape");
  EFM_REQUIRE(E_all && workspace && aligned(workspace, 256), EFM_E_ARG, "efm_bh_mine: bad pointer");
  const BhLayout L = bh_layout(B_local, B_all, D, dtype, flags);
  EFM_REQUIRE(ws_bytes >= L.total, EFM_E_WORKSPACE, "efm_bh_mine: workspace too small");
  // same kernel variant as the step; its extra warps leave the (already populated) label table alone
  return bh_launch_mining(L, E_all, static_cast<uint8_t*>(workspace), B_all, row_offset, B_local, D, dtype,
                          (flags & EFM_BH_NORMALIZE) ? 1 : 0, static_cast<cudaStream_t>(stream), false,
                          bh_table_in_gram(L, B_local) ? 2 : 0);
}

// ============================================================================================================
// Multi-GPU step with the exchange fused into the Gram kernel (BASELINE config C3).
//
// One process per GPU.  Every rank owns an EXCHANGE REGION (efm_xchg_alloc: plain cudaMalloc memory, because CUDA IPC
// cannot export sub-allocations of a caching allocator) that all peers map through CUDA IPC.  Layout (the same on
// every rank):  [cnt_a u64[8] | cnt_b u64[8] | done, gate, .. | ready u64[8] | buffer 0 | buffer 1], each buffer =
//   xm (B_all, Dp) bf16 mining rows | sq (B_all) | lab (B_all) | nrm (B_all) | E (B_all, D) original rows (absent when
//   the rows are bf16 with D % 64 == 0: they are their own mining rows) | res_ip (B_all) | res_in (B_all) | res_w (B_all)
// with rank r's rows in block r of every array.  Step s uses buffer s & 1; counters and flags only ever grow (step s waits
// for s * increments-per-step), so nothing is reset between steps.  A peer can be at most one step ahead (its next Gram
// kernel needs this rank's next block), hence two buffers suffice.
//
//   memset        label table + chain heads (local workspace)
//   bh_prep       LOCAL rows only -> block `rank` of xm / sq / lab / nrm in the own region
//   gram_kernel   <MineEpi, COMM>: announces the block to the peers (one flag store each); two extra warps per CTA COPY the
//                 peers' blocks into the own region with loads over NVLink, in the order of the column-block-major tile
//                 schedule (local block first), and bump a local arrival counter per block; the producer / epilogue warps
//                 wait for a block's counter right before its first tile -> the transfer is hidden behind the tiles of the
//                 blocks that are already there; afterwards the same warps build the per-label member table
//   bh_finalize   local anchors; leaves (idx_p, idx_n, weight) in the own region, counts finished blocks in a local counter
//   bh_chain      exact mode: one block waits for the peers' completion counters (read over NVLink), then every thread FETCHES
//                 one anchor's result from its owner (12 B per anchor instead of a dense reduce-scatter of dE) and chains it
//                 if it mined a local row
//   bh_grad       local rows only -> dE_local (B_local, D)
// No NCCL call, no host synchronisation, no system-scope fence inside the step.
namespace efm {
namespace {

struct XchgLayout { int Dp; bool e_is_xm; size_t cnt_a, cnt_b, cnt_c, ready, buf0, buf_bytes, xm, sq, lab, nrm, E, rip, rin, rw, total; };
XchgLayout xchg_layout(long B, int world, long D, int dtype) {
  XchgLayout L{};
  L.Dp = (int)(cdiv(D, tc::BK) * tc::BK);
  const size_t B_all = (size_t)B * world;
  L.cnt_a = 0; L.cnt_b = 64; L.cnt_c = 128; L.ready = 192; L.buf0 = 256;   // cnt_c: [0] finalize-completion counter, [1] chain gate
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.xm = take(B_all * L.Dp * 2);
  L.sq = take(B_all * 4); L.lab = take(B_all * 4); L.nrm = take(B_all * 4);
  // bf16 rows with K a multiple of 64 ARE their own mining rows: one array serves the Gram kernel, finalize and gradient,
  // and there is no stage B.  Other dtypes / widths: the original rows travel as stage B next to the bf16 mining rows.
  L.e_is_xm = dtype == EFM_BF16 && D == L.Dp;
  L.E = take(L.e_is_xm ? 0 : B_all * D * dtype_size(dtype));
  L.rip = take(B_all * 4); L.rin = take(B_all * 4); L.rw = take(B_all * 4);
  L.buf_bytes = off;
  L.total = L.buf0 + 2 * L.buf_bytes;
  return L;
}

struct DistLayout { int tbits; size_t kpos, kneg, hkeys, hcnt, hhead, rheadp, rheadn, clr_begin, clr_bytes, hmem, hnext, hslot, rnextp, rnextn, wgt, total; };
DistLayout dist_layout(long B, long B_all) {
  DistLayout L{};
  L.tbits = 4;
  while ((1L << L.tbits) < 2 * B_all) ++L.tbits;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.kpos = take((size_t)B * 8); L.kneg = take((size_t)B * 8);
  L.hkeys = take((size_t)8 << L.tbits); L.hcnt = take((size_t)4 << L.tbits); L.hhead = take((size_t)4 << L.tbits);
  L.rheadp = take((size_t)B_all * 4); L.rheadn = take((size_t)B_all * 4);
  L.clr_begin = L.hkeys; L.clr_bytes = off - L.clr_begin;
  L.hmem = take((size_t)32 << L.tbits);
  L.hnext = take((size_t)B_all * 4); L.hslot = take((size_t)B_all * 4);
  L.rnextp = take((size_t)B_all * 4); L.rnextn = take((size_t)B_all * 4);
  L.wgt = take((size_t)B * 4);
  L.total = off;
  return L;
}

int dist_check_shape(const char* who, long B, int world, int rank, long D, int dtype) {
  EFM_REQUIRE(world >= 1 && world <= tc::MAX_PEERS && rank >= 0 && rank < world, EFM_E_ARG, "%s: world=%d rank=%d (at most %d ranks)",
              who, world, rank, tc::MAX_PEERS);
  EFM_REQUIRE(B > 0 && D > 0, EFM_E_ARG, "%s: bad shape B_local=%ld D=%ld", who, B, D);
  EFM_REQUIRE(dtype == EFM_F32 || dtype == EFM_BF16 || dtype == EFM_F16, EFM_E_ARG, "%s: dtype %d", who, dtype);
  EFM_REQUIRE(B % 256 == 0, EFM_E_UNSUPPORTED, "%s: B_local=%ld must be a multiple of 256 (whole 256-column tile blocks per rank)", who, B);
  EFM_REQUIRE(D % 8 == 0 && cdiv(D, tc::BK) <= tc::MAX_NK, EFM_E_UNSUPPORTED, "%s: D=%ld must be a multiple of 8 and <= %d", who, D,
              tc::MAX_NK * tc::BK);
  EFM_REQUIRE((long)B * world < (1L << 30), EFM_E_ARG, "%s: global batch too large", who);
  return EFM_OK;
}

}  // namespace
}  // namespace efm

extern "C" size_t efm_bh_xchg_bytes(long B_local, int world, long D, int dtype) {
  if (B_local <= 0 || world < 1 || world > tc::MAX_PEERS || D <= 0) return 0;
  return xchg_layout(B_local, world, D, dtype).total;
}
extern "C" size_t efm_bh_dist_workspace_bytes(long B_local, int world) {
  if (B_local <= 0 || world < 1) return 0;
  return dist_layout(B_local, B_local * world).total;
}

extern "C" int efm_xchg_alloc(size_t bytes, void** region) {
  EFM_REQUIRE(region && bytes > 0, EFM_E_ARG, "efm_xchg_alloc: bad argument");
  void* p = nullptr;
  EFM_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);                              // arrival counters start at 0
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "efm_xchg_alloc: memset"); }
  *region = p;
  return EFM_OK;
}
extern "C" int efm_xchg_free(void* region) {
  if (region) EFM_CUDA(cudaFree(region));
  return EFM_OK;
}
extern "C" int efm_ipc_export(const void* region, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  EFM_REQUIRE(region && handle64, EFM_E_ARG, "efm_ipc_export: null pointer");
  cudaIpcMemHandle_t h;
  EFM_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(region)));
  memcpy(handle64, &h, sizeof(h));
  return EFM_OK;
}
extern "C" int efm_ipc_import(const void* handle64, void** mapped) {
  EFM_REQUIRE(handle64 && mapped, EFM_E_ARG, "efm_ipc_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  EFM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *mapped = p;
  return EFM_OK;
}
extern "C" int efm_ipc_release(void* mapped) {
  if (mapped) EFM_CUDA(cudaIpcCloseMemHandle(mapped));
  return EFM_OK;
}

extern "C" int efm_bh_dist_fwd_bwd(const void* E_local, const int32_t* labels_local, long B_local, long D, int dtype,
                                   float margin, const float* head_grad, int flags, int world, int rank,
                                   void* const* regions, unsigned long long step, float* loss, int32_t* idx_p,
                                   int32_t* idx_n, float* dE_local, void* workspace, size_t ws_bytes, void* stream) {
  { const int rc = dist_check_shape("efm_bh_dist_fwd_bwd", B_local, world, rank, D, dtype); if (rc) return rc; }
  EFM_REQUIRE(E_local && labels_local && regions && loss && idx_p && idx_n && workspace, EFM_E_ARG, "efm_bh_dist_fwd_bwd: null pointer");
  EFM_REQUIRE(step >= 1, EFM_E_ARG, "efm_bh_dist_fwd_bwd: step counts from 1 (the same value on every rank)");
  EFM_REQUIRE(aligned(workspace, 256) && aligned(E_local, 16) && (!dE_local || aligned(dE_local, 16)), EFM_E_ALIGN,
              "efm_bh_dist_fwd_bwd: workspace must be 256-byte, E_local / dE_local 16-byte aligned");
  for (int r = 0; r < world; ++r) EFM_REQUIRE(regions[r] && aligned(regions[r], 256), EFM_E_ARG, "efm_bh_dist_fwd_bwd: region %d missing", r);
  const long B = B_local, B_all = B * world, row0 = B * rank;
  const XchgLayout X = xchg_layout(B, world, D, dtype);
  const DistLayout L = dist_layout(B, B_all);
  EFM_REQUIRE(ws_bytes >= L.total, EFM_E_WORKSPACE, "efm_bh_dist_fwd_bwd: workspace %zu < required %zu", ws_bytes, L.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint8_t* own = static_cast<uint8_t*>(regions[rank]);
  const size_t boff = X.buf0 + (size_t)(step & 1ull) * X.buf_bytes;      // this step's buffer inside every region
  uint8_t* buf = own + boff;
  auto* xm = reinterpret_cast<__nv_bfloat16*>(buf + X.xm);
  float* sq = reinterpret_cast<float*>(buf + X.sq);
  int32_t* lab = reinterpret_cast<int32_t*>(buf + X.lab);
  float* nrm = reinterpret_cast<float*>(buf + X.nrm);
  const void* E_all = X.e_is_xm ? static_cast<const void*>(xm) : static_cast<const void*>(buf + X.E);
  int32_t* rip = reinterpret_cast<int32_t*>(buf + X.rip);
  int32_t* rin = reinterpret_cast<int32_t*>(buf + X.rin);
  float* rw = reinterpret_cast<float*>(buf + X.rw);
  auto* kpos = reinterpret_cast<unsigned long long*>(ws + L.kpos);
  auto* kneg = reinterpret_cast<unsigned long long*>(ws + L.kneg);
  auto* hkeys = reinterpret_cast<unsigned long long*>(ws + L.hkeys);
  int32_t* hcnt = reinterpret_cast<int32_t*>(ws + L.hcnt);
  int32_t* hmem = reinterpret_cast<int32_t*>(ws + L.hmem);
  int32_t* hhead = reinterpret_cast<int32_t*>(ws + L.hhead);
  int32_t* hnext = reinterpret_cast<int32_t*>(ws + L.hnext);
  int32_t* hslot = reinterpret_cast<int32_t*>(ws + L.hslot);
  int32_t* rheadp = reinterpret_cast<int32_t*>(ws + L.rheadp);
  int32_t* rheadn = reinterpret_cast<int32_t*>(ws + L.rheadn);
  int32_t* rnextp = reinterpret_cast<int32_t*>(ws + L.rnextp);
  int32_t* rnextn = reinterpret_cast<int32_t*>(ws + L.rnextn);
  float* wgt = reinterpret_cast<float*>(ws + L.wgt);
  const int normalize = (flags & EFM_BH_NORMALIZE) ? 1 : 0;
  const bool exact = dE_local && !(flags & EFM_BH_DETACH_GATHERED);
  const size_t es = dtype_size(dtype);
  const unsigned grid_loc = (unsigned)(B / WARPS);

  EFM_BH_STAGE(0, st);
  if (!X.e_is_xm)                                                        // original rows next to the bf16 mining rows (stage B source)
    EFM_CUDA(cudaMemcpyAsync(buf + X.E + (size_t)row0 * D * es, E_local, (size_t)B * D * es, cudaMemcpyDeviceToDevice, st));
  // 1. local rows -> block `rank` of this step's buffer
  EFM_DISPATCH_DTYPE(dtype, T, {
    bh_prep_kernel<T, true><<<grid_loc, WARPS * 32, 0, st>>>((const T*)E_local, labels_local, B, (int)D, X.Dp, normalize, 1, 0, 0, B,
                                                             sq + row0, nrm + row0, lab + row0, xm + row0 * X.Dp, nullptr, kpos, kneg,
                                                             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                                             reinterpret_cast<uint4*>(ws + L.clr_begin), (long)(L.clr_bytes / 16));
  });
  EFM_LAUNCH_CHECK("bh_prep(local)");
  EFM_BH_STAGE(1, st);
  // 2. Gram + mining + exchange
  unsigned long long target_ab = 0;
  {
    tc::GramParams g{};
    g.M = (int)B; g.N = (int)B_all; g.nk = X.Dp / tc::BK;
    g.phases = world; g.phase0 = rank;
    const int grid = tc::plan_gram(g, 1, 0, MineEpi::SMEM_PER_WARP * tc::EPI_WARPS, 2);
    EFM_REQUIRE(g.stages >= 1, EFM_E_UNSUPPORTED, "efm_bh_dist_fwd_bwd: D = %ld leaves no shared memory for the B pipeline", D);
    target_ab = step * (unsigned long long)grid * tc::COMM_WARPS;        // every exchange warp adds 1 per (block, stage) per step
    g.ready = reinterpret_cast<const unsigned long long*>(own + X.cnt_a);
    g.ready_target = target_ab;
    tc::CommParams cp{};
    cp.world = world; cp.rank = rank; cp.nseg_a = 4; cp.nseg = X.e_is_xm ? 4 : 5; cp.step = step;
    for (int r = 0; r < world; ++r) cp.peer[r] = static_cast<uint8_t*>(regions[r]);
    cp.seg[0] = {boff + X.xm, (unsigned long long)B * X.Dp * 2};
    cp.seg[1] = {boff + X.sq, (unsigned long long)B * 4};
    cp.seg[2] = {boff + X.lab, (unsigned long long)B * 4};
    cp.seg[3] = {boff + X.nrm, (unsigned long long)B * 4};
    cp.seg[4] = {boff + X.E, (unsigned long long)B * D * es};
    cp.ready_off = X.ready; cp.cnt_a_off = X.cnt_a; cp.cnt_b_off = X.cnt_b;
    cp.tab_lab = lab; cp.tab_rows = B_all;                               // label table of the global batch, built by the same warps
    cp.hkeys = hkeys; cp.hcnt = hcnt; cp.hmem = hmem; cp.hhead = hhead; cp.hnext = hnext; cp.hslot = hslot; cp.tbits = L.tbits;
    MineEpi::Params ep{sq, lab, kneg, (int)row0, normalize};
    const int rc = tc::launch_gram_mh<MineEpi, 1, 2, true>(xm + row0 * X.Dp, xm, g, ep, grid, st, &cp, true);
    if (rc) return rc;
  }
  EFM_BH_STAGE(2, st);
  // 3. finalize the local anchors (waits for stage B: the original rows of every rank, own copy included)
  DistFin df{};
  df.world = world; df.rank = rank;
  df.wait_b = X.e_is_xm ? nullptr : reinterpret_cast<const unsigned long long*>(own + X.cnt_b);
  df.target_b = target_ab;
  df.rip = rip; df.rin = rin; df.rw = rw;                                // published in every mode, so the completion counters of
  df.done = reinterpret_cast<unsigned long long*>(own + X.cnt_c);        // all ranks advance in lockstep whatever mix of flags is used
  const bool chains_local = dE_local && !exact;
  EFM_DISPATCH_DTYPE(dtype, T, {
    EFM_LAUNCH((bh_finalize_kernel<T, true>), grid_loc, WARPS * 32, 0, st, true, (const T*)E_all, nrm, normalize, B_all, (int)D, row0, B,
               margin, head_grad, chains_local ? 1 : 0, 1, kpos, kneg, xm, X.Dp, sq, lab, hcnt, hmem, hhead, hnext, hslot, loss, idx_p,
               idx_n, chains_local ? wgt : nullptr, chains_local ? rheadp : nullptr, rheadn, rnextp, rnextn, df);
  });
  EFM_LAUNCH_CHECK("bh_finalize(dist)");
  EFM_BH_STAGE(4, st);
  if (!dE_local) return EFM_OK;
  // 5./6. gradient of the local rows
  if (exact) {
    ChainPull pull{};
    pull.world = world; pull.rank = rank; pull.rows_per_rank = B;
    pull.target = step * (unsigned long long)grid_loc;                   // one count per finalize block of the owner per step
    pull.gate = reinterpret_cast<unsigned long long*>(own + X.cnt_c) + 1;
    pull.step = step;
    for (int r = 0; r < world; ++r) {
      pull.done[r] = reinterpret_cast<const unsigned long long*>(static_cast<const uint8_t*>(regions[r]) + X.cnt_c);
      const uint8_t* b = static_cast<const uint8_t*>(regions[r]) + boff;
      pull.ip[r] = reinterpret_cast<const int32_t*>(b + X.rip); pull.in[r] = reinterpret_cast<const int32_t*>(b + X.rin);
      pull.w[r] = reinterpret_cast<const float*>(b + X.rw);
    }
    EFM_LAUNCH(bh_chain_kernel, (unsigned)cdiv(B_all, 256), 256, 0, st, true, rip, rin, rw, B_all, row0, B, rheadp, rheadn, rnextp, rnextn,
               pull);
    EFM_LAUNCH_CHECK("bh_chain(dist)");
    EFM_BH_STAGE(5, st);
    EFM_DISPATCH_DTYPE(dtype, T, {
      EFM_LAUNCH((bh_grad_kernel<T, 8>), grid_loc, WARPS * 32, 0, st, true, (const T*)E_all, nrm, normalize, B, (int)D, 0L, B_all, rip, rin,
                 rw, rheadp, rheadn, rnextp, rnextn, dE_local, row0);
    });
  } else {
    EFM_DISPATCH_DTYPE(dtype, T, {
      EFM_LAUNCH((bh_grad_kernel<T, 8>), grid_loc, WARPS * 32, 0, st, true, (const T*)E_all, nrm, normalize, B, (int)D, row0, B, idx_p,
                 idx_n, wgt, rheadp + row0, rheadn + row0, rnextp, rnextn, dE_local, row0);
    });
  }
  EFM_LAUNCH_CHECK("bh_grad(dist)");
  EFM_BH_STAGE(6, st);
  return EFM_OK;
}
