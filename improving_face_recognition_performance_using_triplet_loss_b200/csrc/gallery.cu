// gallery.cu -- cosine gallery scoring and the all-pairs cosine matrix (C4).
//
//  * exact fp32 path: bit-identical restatement of simd_dot (Feature.hpp:273-293) -- 4 accumulator lanes
//    mapped onto 4 adjacent threads per gallery row, non-fused mul/add in the reference's order, then
//    sim = dot / (gnorm * qnorm) with IEEE div/sqrt; HBM-bound (N*D*4 bytes per pass over the gallery).
//  * bf16 path: tcgen05 Gram (tc_gram.cuh) with a top-K epilogue: per query row a register-resident sorted
//    list; a 32-column chunk is only examined element-wise when its maximum can enter the list.
//  * selection / merge kernels order by (score desc, index asc); strict replacement keeps the earlier row on
//    ties, NaN scores are skipped, score < sim_th skipped (Feature.hpp:314-319,358-363).
//  * all-pairs cosine: same mainloop, epilogue scales by 1/(|a||b|) and writes coalesced rows through a
//    per-warp shared-memory transpose.
#include "tc_gram.cuh"

namespace efm {
namespace {

__device__ __forceinline__ long cdiv_dev(long a, long b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------ sorted top-K in registers
template <int KMAX> struct TopK {
  float s[KMAX];
  int id[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { s[k] = -INFINITY; id[k] = -1; }
  }
  // candidate order: (score desc, index asc); empty slots (id < 0) are worst
  static __device__ __forceinline__ bool better(float s1, int i1, float s2, int i2) {
    return i2 < 0 || s1 > s2 || (s1 == s2 && i1 < i2);
  }
  __device__ __forceinline__ bool full() const { return id[KMAX - 1] >= 0; }
  __device__ __forceinline__ void insert(float sc, int idx) {
    if (!better(sc, idx, s[KMAX - 1], id[KMAX - 1])) return;
    s[KMAX - 1] = sc; id[KMAX - 1] = idx;
#pragma unroll
    for (int k = KMAX - 1; k > 0; --k) {
      if (better(s[k], id[k], s[k - 1], id[k - 1])) {
        const float ts = s[k]; s[k] = s[k - 1]; s[k - 1] = ts;
        const int ti = id[k]; id[k] = id[k - 1]; id[k - 1] = ti;
      }
    }
  }
};

__device__ __forceinline__ bool accept_score(float sim, float sim_th, int flags) {
  if (sim != sim) return false;                              // NaN: "no reg data", skipped
  if (sim < sim_th) return false;
  if ((flags & EFM_GAL_POSITIVE_ONLY) && !(sim > 0.f)) return false;   // maxsim starts at 0
  return true;
}

// ------------------------------------------------------------------------------------------ exact simd_dot order
// thread (row, l) owns accumulator lane l of the reference's __m128; QT queries per pass.
template <int QT, bool RAW>
__global__ void __launch_bounds__(256)
gallery_scores_exact_kernel(const float* __restrict__ G, const float* __restrict__ gnorm, long N, int D,
                            const float* __restrict__ Q, int nq, int q0, float* __restrict__ out) {
  extern __shared__ float qs[];                               // QT * D queries + QT norms
  float* qn = qs + QT * D;
  for (int i = threadIdx.x; i < QT * D; i += blockDim.x) {
    const int q = i / D, d = i % D;
    qs[i] = (q0 + q < nq) ? Q[(long)(q0 + q) * D + d] : 0.f;
  }
  __syncthreads();
  const int nquads = (D - 1) / 4;                             // loop runs while i + 4 < len
  // query norms with the same routine: sqrt(simd_dot(q, q))
  if (threadIdx.x < 4 * QT) {
    const int q = threadIdx.x >> 2, l = threadIdx.x & 3;
    const float* x = qs + q * D;
    float a = 0.f;
    for (int i = 0; i < nquads; ++i) a = __fadd_rn(a, __fmul_rn(x[4 * i + l], x[4 * i + l]));
    const unsigned m = 0xffffffffu >> (32 - 4 * QT);
    const float a1 = __shfl_down_sync(m, a, 1, 4), a2 = __shfl_down_sync(m, a, 2, 4), a3 = __shfl_down_sync(m, a, 3, 4);
    if (l == 0) {
      float s = __fadd_rn(__fadd_rn(__fadd_rn(a, a1), a2), a3);
      for (int i = 4 * nquads; i < D; ++i) s = __fadd_rn(s, __fmul_rn(x[i], x[i]));
      qn[q] = __fsqrt_rn(s);
    }
  }
  __syncthreads();
  const int l = threadIdx.x & 3;
  const long rows_per_block = blockDim.x >> 2;
  for (long base = (long)blockIdx.x * rows_per_block; base < N; base += (long)gridDim.x * rows_per_block) {
    const long row = base + (threadIdx.x >> 2);
    const bool ok = row < N;                                   // loop bound is block-uniform (shuffles below)
    const float* g = G + (ok ? row : 0) * D;
    float acc[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) acc[q] = 0.f;
    for (int i = 0; i < nquads; ++i) {
      const float x = __ldg(g + 4 * i + l);
#pragma unroll
      for (int q = 0; q < QT; ++q) acc[q] = __fadd_rn(acc[q], __fmul_rn(x, qs[q * D + 4 * i + l]));
    }
    const float gn = RAW ? 1.f : gnorm[ok ? row : 0];
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      const float a1 = __shfl_down_sync(0xffffffffu, acc[q], 1, 4);
      const float a2 = __shfl_down_sync(0xffffffffu, acc[q], 2, 4);
      const float a3 = __shfl_down_sync(0xffffffffu, acc[q], 3, 4);
      if (l == 0 && ok && q0 + q < nq) {
        float s = __fadd_rn(__fadd_rn(__fadd_rn(acc[q], a1), a2), a3);
        for (int i = 4 * nquads; i < D; ++i) s = __fadd_rn(s, __fmul_rn(__ldg(g + i), qs[q * D + i]));
        out[(long)(q0 + q) * N + row] = RAW ? s : __fdiv_rn(s, __fmul_rn(gn, qn[q]));
      }
    }
  }
}

// sqrt(simd_dot(g, g)) per row -- the `sqrt_simdot` the reference stores at registration (Feature.hpp:476,535)
__global__ void __launch_bounds__(256)
simd_norm_rows_kernel(const float* __restrict__ G, long N, int D, float* __restrict__ out) {
  const int l = threadIdx.x & 3;
  const int nquads = (D - 1) / 4;
  const long rows_per_block = blockDim.x >> 2;
  for (long base = (long)blockIdx.x * rows_per_block; base < N; base += (long)gridDim.x * rows_per_block) {
    const long row = base + (threadIdx.x >> 2);
    const bool ok = row < N;
    const float* g = G + (ok ? row : 0) * D;
    float a = 0.f;
    for (int i = 0; i < nquads; ++i) { const float x = __ldg(g + 4 * i + l); a = __fadd_rn(a, __fmul_rn(x, x)); }
    const float a1 = __shfl_down_sync(0xffffffffu, a, 1, 4), a2 = __shfl_down_sync(0xffffffffu, a, 2, 4),
                a3 = __shfl_down_sync(0xffffffffu, a, 3, 4);
    if (l == 0 && ok) {
      float s = __fadd_rn(__fadd_rn(__fadd_rn(a, a1), a2), a3);
      for (int i = 4 * nquads; i < D; ++i) { const float x = __ldg(g + i); s = __fadd_rn(s, __fmul_rn(x, x)); }
      out[row] = __fsqrt_rn(s);
    }
  }
}

// ------------------------------------------------------------------------------------------ selection over a score matrix
// grid (chunks, nq): block scans rows [c*len, (c+1)*len) of query q, emits a sorted partial list of K.
template <int KMAX>
__global__ void __launch_bounds__(256)
topk_select_kernel(const float* __restrict__ scores, long N, int K, float sim_th, int flags, long idx_base,
                   float* __restrict__ part_s, int32_t* __restrict__ part_i, long nq) {
  __shared__ float ls[256 * KMAX];
  __shared__ int li[256 * KMAX];
  const int chunks = gridDim.x;
  const long q = blockIdx.y;
  const long len = cdiv_dev(N, chunks);
  const long r0 = blockIdx.x * len, r1 = (r0 + len < N) ? r0 + len : N;
  TopK<KMAX> tk;
  tk.init();
  const float* sc = scores + q * N;
  for (long r = r0 + threadIdx.x; r < r1; r += 256) {
    const float sim = sc[r];
    if (accept_score(sim, sim_th, flags)) tk.insert(sim, (int)r);
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { ls[threadIdx.x * KMAX + k] = tk.s[k]; li[threadIdx.x * KMAX + k] = tk.id[k]; }
  __syncthreads();
  // warp 0 pops the K best of the 256 sorted lists: lane owns 8 lists, keeps a head per list
  if (threadIdx.x < 32) {
    int head[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) head[j] = 0;
    for (int k = 0; k < K; ++k) {
      float bs = -INFINITY; int bi = -1, bj = -1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = threadIdx.x * 8 + j;
        if (head[j] < KMAX) {
          const float s = ls[t * KMAX + head[j]]; const int i = li[t * KMAX + head[j]];
          if (i >= 0 && TopK<KMAX>::better(s, i, bs, bi)) { bs = s; bi = i; bj = j; }
        }
      }
      // warp argbest
      float ws = bs; int wi = bi;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, ws, o); const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
        if (oi >= 0 && TopK<KMAX>::better(os, oi, ws, wi)) { ws = os; wi = oi; }
      }
      if (wi >= 0 && wi == bi) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j == bj) head[j]++;
      }
      if (threadIdx.x == 0) {
        const long o = ((long)blockIdx.x * nq + q) * K + k;
        part_s[o] = wi >= 0 ? ws : 0.f;
        part_i[o] = wi >= 0 ? (int32_t)(wi + idx_base) : -1;
      }
    }
  }
}

// K-way merge of `parts` partial lists per query: one block per query.  Every thread folds a strided share of the
// parts*K candidates into a sorted register list, then warp 0 pops the K best of the 256 lists from shared memory.
template <int KMAX>
__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ part_s, const int32_t* __restrict__ part_i, int parts, long nq, int K,
                  float* __restrict__ out_s, int32_t* __restrict__ out_i) {
  __shared__ float ls[256 * KMAX];
  __shared__ int li[256 * KMAX];
  const long q = blockIdx.x;
  TopK<KMAX> tk;
  tk.init();
  const int total = parts * K;
  for (int c = threadIdx.x; c < total; c += 256) {
    const int p = c / K, kk = c % K;
    const long o = ((long)p * nq + q) * K + kk;
    const int i = part_i[o];
    if (i >= 0) tk.insert(part_s[o], i);
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { ls[threadIdx.x * KMAX + k] = tk.s[k]; li[threadIdx.x * KMAX + k] = tk.id[k]; }
  __syncthreads();
  if (threadIdx.x < 32) {
    int head[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) head[j] = 0;
    for (int k = 0; k < K; ++k) {
      float bs = -INFINITY; int bi = -1, bj = -1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = threadIdx.x * 8 + j;
        if (head[j] < KMAX) {
          const float s = ls[t * KMAX + head[j]]; const int i = li[t * KMAX + head[j]];
          if (i >= 0 && TopK<KMAX>::better(s, i, bs, bi)) { bs = s; bi = i; bj = j; }
        }
      }
      float ws = bs; int wi = bi;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, ws, o); const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
        if (oi >= 0 && TopK<KMAX>::better(os, oi, ws, wi)) { ws = os; wi = oi; }
      }
      if (wi >= 0 && wi == bi) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j == bj) head[j]++;
      }
      if (threadIdx.x == 0) { out_s[q * K + k] = wi >= 0 ? ws : 0.f; out_i[q * K + k] = wi; }
    }
  }
}

static int launch_topk_merge(const float* ps, const int32_t* pi, int parts, long nq, int K, float* out_s, int32_t* out_i,
                             cudaStream_t st) {
  if (K == 1) topk_merge_kernel<1><<<(unsigned)nq, 256, 0, st>>>(ps, pi, parts, nq, K, out_s, out_i);
  else if (K <= 5) topk_merge_kernel<5><<<(unsigned)nq, 256, 0, st>>>(ps, pi, parts, nq, K, out_s, out_i);
  else topk_merge_kernel<EFM_GAL_MAX_K><<<(unsigned)nq, 256, 0, st>>>(ps, pi, parts, nq, K, out_s, out_i);
  EFM_LAUNCH_CHECK("topk_merge");
  return EFM_OK;
}

// ------------------------------------------------------------------------------------------ operand prep for the tcgen05 paths
// rows -> bf16, K padded with zeros to Kp; optional fp32 norm of the (bf16-rounded) row and its reciprocal
template <typename T>
__global__ void __launch_bounds__(256)
pad_rows_kernel(const T* __restrict__ X, long rows, int D, int Kp, __nv_bfloat16* __restrict__ out,
                float* __restrict__ nrm_out, float* __restrict__ rnrm_out) {
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.f;
  for (int d = lane; d < Kp; d += 32) {
    const float v = d < D ? El<T>::to_f(X[row * D + d]) : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    if (out) out[row * Kp + d] = b;
    const float vb = __bfloat162float(b);
    ss = fmaf(vb, vb, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) {
    const float n = sqrtf(ss);
    if (nrm_out) nrm_out[row] = n;
    if (rnrm_out) rnrm_out[row] = 1.f / n;
  }
}

__global__ void recip_kernel(const float* __restrict__ x, long n, long npad, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) out[i] = i < n ? 1.f / x[i] : 0.f;
}

// ------------------------------------------------------------------------------------------ tcgen05 top-K epilogue
template <int KMAX> struct TopkEpi {
  static constexpr int SMEM_PER_WARP = 2 * 256;      // double-buffered 1/gnorm of the warp's 64-column slice
  static constexpr int CW = 16;
  struct Params {
    const float* gnorm;   // (N) stored gallery norms
    const float* rgn;     // (N) reciprocals
    const float* qn;      // (nq) query norms
    float sim_th;
    int flags, K;
    long idx_base, nq;
    float* part_s;        // (parts, nq, K)
    int32_t* part_i;
    int col_slices;       // 64-column slices per row half (4 for the 128x256 shape, 2 for 256x128)
    int cg;               // CTAs per tile (2 = cta_group::2 pairs)
  };
  struct State {
    TopK<KMAX> tk;
    float qn, rqn;
    bool valid;
  };
  static __device__ __forceinline__ void row_begin(const Params& p, const tc::GramParams& g, int row, State& s) {
    s.valid = row < g.M;
    s.qn = s.valid ? p.qn[row] : 1.f;
    s.rqn = 1.f / s.qn;
    s.tk.init();
  }
  // 1/gnorm of the warp's 64-column slice -> shared memory one tile ahead (lanes 0-15, 16 B each; rgn is padded)
  static __device__ __forceinline__ void prefetch(const Params& p, const tc::GramParams&, float* scratch, int buf, int colw) {
    const int lane = threadIdx.x & 31;
    if (lane < 16) tc::cp_async16(scratch + buf * 64 + lane * 4, p.rgn + colw + lane * 4);
    tc::cp_async_commit();
  }
  static __device__ __forceinline__ void prefetch_wait() { tc::cp_async_wait_all(); __syncwarp(); }
  static __device__ __forceinline__ void chunk(const Params& p, const tc::GramParams& g, State& s, int row,
                                               const uint32_t (&acc)[16], int col0, float* scratch, int bc) {
    if (!s.valid) return;                            // padding rows of the last query block (their Gram is all zeros)
    const int ncols = g.N - col0;
    float est[16];                                   // g * (1/gnorm): cheap estimate of sim * |q|
    if (ncols >= 16) {
      const float4* r4 = reinterpret_cast<const float4*>(scratch + (bc >> 2) * 64 + (bc & 3) * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 r = r4[q];
        est[4 * q + 0] = __uint_as_float(acc[4 * q + 0]) * r.x;
        est[4 * q + 1] = __uint_as_float(acc[4 * q + 1]) * r.y;
        est[4 * q + 2] = __uint_as_float(acc[4 * q + 2]) * r.z;
        est[4 * q + 3] = __uint_as_float(acc[4 * q + 3]) * r.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) est[k] = (k < ncols) ? __uint_as_float(acc[k]) * p.rgn[col0 + k] : -INFINITY;
    }
    float m = est[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) m = fmaxf(m, est[k]);
    // conservative filter: can the chunk maximum enter the list?  (reciprocal-multiply estimate, 1e-5 slack)
    float cut = p.sim_th;
    if (s.tk.full()) cut = fmaxf(cut, s.tk.s[KMAX - 1]);
    const float cutm = (cut - fabsf(cut) * 1e-5f - 1e-30f) * s.qn;   // compare g / gnorm against cut * |q|
    // NaN anywhere in the chunk?  fmaxf drops NaNs, a sum keeps them (FADDs on the under-used FMA pipe; inf - inf
    // can only produce a false positive, which merely takes the exact path below)
    const float nansum = ((est[0] + est[1]) + (est[2] + est[3])) + ((est[4] + est[5]) + (est[6] + est[7])) +
                         (((est[8] + est[9]) + (est[10] + est[11])) + ((est[12] + est[13]) + (est[14] + est[15])));
    if (!(m >= cutm) && !(nansum != nansum)) return;
    // rare: collect the surviving elements in a bit mask, then evaluate exactly (IEEE division) only those.
    // One copy of the insertion code, accumulators re-read from a local-memory copy by dynamic index (the fully
    // unrolled form made the kernel instruction-cache bound: 55 % stall_no_inst in ncu).
    unsigned mask = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) mask |= ((est[k] >= cutm || est[k] != est[k]) ? 1u : 0u) << k;
    float tmp[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) tmp[k] = __uint_as_float(acc[k]);
#pragma unroll 1
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      if (k >= ncols) break;
      const float sim = __fdiv_rn(tmp[k], __fmul_rn(p.gnorm[col0 + k], s.qn));
      if (accept_score(sim, p.sim_th, p.flags)) s.tk.insert(sim, col0 + k);
    }
  }
  static __device__ __forceinline__ void row_end(const Params& p, const tc::GramParams& g, int row, int cs, State& s) {
    if (!s.valid) return;
    const int part = (int)((blockIdx.x / p.cg) % g.splits) * p.col_slices + cs;                          // (split, column slice)
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < p.K) {
        const long o = ((long)part * p.nq + row) * p.K + k;
        p.part_s[o] = s.tk.id[k] >= 0 ? s.tk.s[k] : 0.f;
        p.part_i[o] = s.tk.id[k] >= 0 ? (int32_t)(s.tk.id[k] + p.idx_base) : -1;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------ all-pairs cosine epilogue
struct StoreEpi {
  static constexpr int CW = 16;
  static constexpr int SMEM_PER_WARP = 32 * 17 * 4;            // 32 rows x 16 columns (+1 pad) transpose tile
  struct Params {
    const float* rna;   // (M) 1/|a_i|
    const float* rnb;   // (N) 1/|b_j|
    float* out;         // (M, ld) with ld >= N
    long ld;            // leading dimension of out in elements (a multiple of 32 keeps every 64-byte store sector-aligned)
  };
  struct State { float rna; int row0; };
  static __device__ __forceinline__ void row_begin(const Params& p, const tc::GramParams& g, int row, State& s) {
    s.rna = row < g.M ? p.rna[row] : 0.f;
    s.row0 = row - (int)(threadIdx.x & 31);                   // first row of this warp's 32-row slab
  }
  static __device__ __forceinline__ void chunk(const Params& p, const tc::GramParams& g, State& s, int row,
                                               const uint32_t (&acc)[16], int col0, float* scratch, int) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) scratch[lane * 17 + k] = __uint_as_float(acc[k]) * s.rna;
    __syncwarp();
    // lanes 0-15 write row r, lanes 16-31 row r+16: two 64-byte segments per store instruction
    const int c = lane & 15, rsel = lane >> 4;
    const int col = col0 + c;
    if (col < g.N) {
      const float rb = p.rnb[col];
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        const int lr = r + 16 * rsel;
        const int orow = s.row0 + lr;
        if (orow < g.M) __stcs(&p.out[(long)orow * p.ld + col], scratch[lr * 17 + c] * rb);
      }
    }
  }
  static __device__ __forceinline__ void prefetch(const Params&, const tc::GramParams&, float*, int, int) {}
  static __device__ __forceinline__ void prefetch_wait() {}
  static __device__ __forceinline__ void row_end(const Params&, const tc::GramParams&, int, int, State&) {}
};

struct TcOperands {
  int Kp;
  bool copy_g;
  size_t gpad, qpad, qn, rgn, parts_s, parts_i, total;
};

TcOperands topk_tc_layout(long N, long D, long nq, int K, int parts_max) {
  TcOperands L{};
  L.Kp = (int)(cdiv(D, tc::BK) * tc::BK);
  L.copy_g = (L.Kp != D);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.gpad = take(L.copy_g ? (size_t)N * L.Kp * 2 : 0);
  L.qpad = take((size_t)nq * L.Kp * 2);
  L.qn = take((size_t)nq * 4);
  L.rgn = take((size_t)cdiv(N, tc::BN) * tc::BN * 4);   // padded to whole column tiles (prefetched by 64-column slices)
  L.parts_s = take((size_t)parts_max * nq * K * 4);
  L.parts_i = take((size_t)parts_max * nq * K * 4);
  L.total = off;
  return L;
}

inline int exact_chunks(long N) {
  long c = cdiv(N, 4096);
  if (c > 1024) c = 1024;
  return (int)(c < 1 ? 1 : c);
}

}  // namespace
}  // namespace efm

using namespace efm;

static int scores_exact(bool raw, const float* G, const float* gnorm, long N, long D, const float* Q, long nq, float* out,
                        void* stream) {
  EFM_REQUIRE(N >= 0 && D > 0 && nq >= 0 && D <= 2048, EFM_E_ARG, "efm_gallery_scores_f32: bad shape N=%ld D=%ld nq=%ld", N, D, nq);
  if (N == 0 || nq == 0) return EFM_OK;
  EFM_REQUIRE(G && (raw || gnorm) && Q && out, EFM_E_ARG, "efm_gallery_scores_f32: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr int QT = 4;
  long blocks = cdiv(N, 64);
  const long cap = 16L * num_sms();
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)(QT * D + QT) * sizeof(float);
  for (long q0 = 0; q0 < nq; q0 += QT) {
    if (raw) gallery_scores_exact_kernel<QT, true><<<(unsigned)blocks, 256, smem, st>>>(G, gnorm, N, (int)D, Q, (int)nq, (int)q0, out);
    else gallery_scores_exact_kernel<QT, false><<<(unsigned)blocks, 256, smem, st>>>(G, gnorm, N, (int)D, Q, (int)nq, (int)q0, out);
    EFM_LAUNCH_CHECK("gallery_scores_exact");
  }
  return EFM_OK;
}

extern "C" int efm_gallery_scores_f32(const float* G, const float* gnorm, long N, long D, const float* Q, long nq,
                                      float* out, void* stream) {
  return scores_exact(false, G, gnorm, N, D, Q, nq, out, stream);
}

extern "C" int efm_simd_dot_rows(const float* G, long N, long D, const float* Q, long nq, float* out, void* stream) {
  return scores_exact(true, G, nullptr, N, D, Q, nq, out, stream);
}

extern "C" int efm_simd_norm_rows(const float* G, long N, long D, float* out, void* stream) {
  EFM_REQUIRE(N >= 0 && D > 0, EFM_E_ARG, "efm_simd_norm_rows: bad shape N=%ld D=%ld", N, D);
  if (N == 0) return EFM_OK;
  EFM_REQUIRE(G && out, EFM_E_ARG, "efm_simd_norm_rows: null pointer");
  long blocks = cdiv(N, 64);
  const long cap = 16L * num_sms();
  if (blocks > cap) blocks = cap;
  simd_norm_rows_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(G, N, (int)D, out);
  EFM_LAUNCH_CHECK("simd_norm_rows");
  return EFM_OK;
}

extern "C" size_t efm_gallery_topk_workspace_bytes(long N, long D, long nq, int K, int dtype) {
  if (N <= 0 || D <= 0 || nq <= 0 || K <= 0) return 0;
  if (dtype == EFM_F32) {
    size_t off = align_up((size_t)nq * N * 4, 256);
    off += align_up((size_t)exact_chunks(N) * nq * K * 4, 256) * 2;
    return off;
  }
  const int parts_max = tc::EPI_SLICES * (int)cdiv(N, tc::BN);
  return topk_tc_layout(N, D, nq, K, parts_max).total;
}

template <int KMAX>
static int topk_tc_launch(const void* Gp, const void* Qp, int mh, int cg, tc::GramParams g, int grid,
                          const typename TopkEpi<KMAX>::Params& ep, cudaStream_t st) {
  return tc::launch_gram<TopkEpi<KMAX>>(Qp, Gp, mh, g, ep, grid, st, cg);
}

extern "C" int efm_gallery_topk(const void* G, const float* gnorm, long N, long D, const void* Q, long nq, int K,
                                float sim_th, int flags, int dtype, long idx_base, float* out_score, int32_t* out_idx,
                                void* workspace, size_t ws_bytes, void* stream) {
  EFM_REQUIRE(N > 0 && D > 0 && nq >= 0 && K >= 1 && K <= EFM_GAL_MAX_K, EFM_E_ARG,
              "efm_gallery_topk: bad shape N=%ld D=%ld nq=%ld K=%d (K <= %d)", N, D, nq, K, EFM_GAL_MAX_K);
  EFM_REQUIRE(N + idx_base < (1L << 31), EFM_E_ARG, "efm_gallery_topk: indices exceed int32");
  if (nq == 0) return EFM_OK;
  EFM_REQUIRE(G && gnorm && Q && out_score && out_idx && workspace, EFM_E_ARG, "efm_gallery_topk: null pointer");
  EFM_REQUIRE(aligned(workspace, 256), EFM_E_ALIGN, "efm_gallery_topk: workspace must be 256-byte aligned");
  const size_t need = efm_gallery_topk_workspace_bytes(N, D, nq, K, dtype);
  EFM_REQUIRE(ws_bytes >= need, EFM_E_WORKSPACE, "efm_gallery_topk: workspace %zu < required %zu", ws_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  if (dtype == EFM_F32) {
    float* scores = reinterpret_cast<float*>(ws);
    const int chunks = exact_chunks(N);
    size_t off = align_up((size_t)nq * N * 4, 256);
    float* ps = reinterpret_cast<float*>(ws + off);
    off += align_up((size_t)chunks * nq * K * 4, 256);
    int32_t* pi = reinterpret_cast<int32_t*>(ws + off);
    int rc = efm_gallery_scores_f32(static_cast<const float*>(G), gnorm, N, D, static_cast<const float*>(Q), nq, scores, stream);
    if (rc) return rc;
    // blockIdx.y limit: loop over query slabs
    for (long qs = 0; qs < nq; qs += 32768) {
      const long qn_ = (nq - qs < 32768) ? nq - qs : 32768;
      dim3 grid(chunks, (unsigned)qn_);
      const float* sc = scores + qs * N;
      // partial layout is (chunks, nq, K) over the FULL nq: offset the base so q indexes the slab
      if (K == 1) topk_select_kernel<1><<<grid, 256, 0, st>>>(sc, N, K, sim_th, flags, idx_base, ps + qs * K, pi + qs * K, nq);
      else if (K <= 4) topk_select_kernel<4><<<grid, 256, 0, st>>>(sc, N, K, sim_th, flags, idx_base, ps + qs * K, pi + qs * K, nq);
      else topk_select_kernel<EFM_GAL_MAX_K><<<grid, 256, 0, st>>>(sc, N, K, sim_th, flags, idx_base, ps + qs * K, pi + qs * K, nq);
      EFM_LAUNCH_CHECK("topk_select");
    }
    return launch_topk_merge(ps, pi, chunks, nq, K, out_score, out_idx, st);
  }

  EFM_REQUIRE(dtype == EFM_BF16 || dtype == EFM_F16, EFM_E_ARG, "efm_gallery_topk: dtype %d", dtype);
  const int Kp = (int)(cdiv(D, tc::BK) * tc::BK);
  EFM_REQUIRE(Kp <= tc::MAX_NK * tc::BK, EFM_E_UNSUPPORTED, "efm_gallery_topk: D=%ld > %d on the tensor-core path", D, tc::MAX_NK * tc::BK);
  tc::GramParams g{};
  g.M = (int)nq; g.N = (int)N; g.nk = Kp / tc::BK;
  const int mh = tc::pick_mh(g.M, g.nk);
  const int cg = tc::pick_cg(g.M, mh, 1);
  const int grid = tc::plan_gram(g, mh, 1, TopkEpi<1>::SMEM_PER_WARP * tc::EPI_WARPS, cg);
  const int col_slices = tc::EPI_SLICES / mh;
  const int parts = g.splits * col_slices;
  const TcOperands L = topk_tc_layout(N, D, nq, K, tc::EPI_SLICES * (int)cdiv(N, tc::BN));
  auto* gpad = reinterpret_cast<__nv_bfloat16*>(ws + L.gpad);
  auto* qpad = reinterpret_cast<__nv_bfloat16*>(ws + L.qpad);
  float* qn = reinterpret_cast<float*>(ws + L.qn);
  float* rgn = reinterpret_cast<float*>(ws + L.rgn);
  float* ps = reinterpret_cast<float*>(ws + L.parts_s);
  int32_t* pi = reinterpret_cast<int32_t*>(ws + L.parts_i);
  const bool copy_g = L.copy_g || dtype == EFM_F16 || !aligned(G, 16);
  EFM_REQUIRE(!copy_g || L.copy_g, EFM_E_ALIGN, "efm_gallery_topk: gallery must be bf16, 16-byte aligned when D %% 64 == 0");
  if (dtype == EFM_BF16) {
    if (copy_g) pad_rows_kernel<__nv_bfloat16><<<(unsigned)cdiv(N, 8), 256, 0, st>>>((const __nv_bfloat16*)G, N, (int)D, Kp, gpad, nullptr, nullptr);
    pad_rows_kernel<__nv_bfloat16><<<(unsigned)cdiv(nq, 8), 256, 0, st>>>((const __nv_bfloat16*)Q, nq, (int)D, Kp, qpad, qn, nullptr);
  } else {
    pad_rows_kernel<__half><<<(unsigned)cdiv(N, 8), 256, 0, st>>>((const __half*)G, N, (int)D, Kp, gpad, nullptr, nullptr);
    pad_rows_kernel<__half><<<(unsigned)cdiv(nq, 8), 256, 0, st>>>((const __half*)Q, nq, (int)D, Kp, qpad, qn, nullptr);
  }
  EFM_LAUNCH_CHECK("pad_rows");
  recip_kernel<<<(unsigned)cdiv(cdiv(N, tc::BN) * tc::BN, 256), 256, 0, st>>>(gnorm, N, cdiv(N, tc::BN) * tc::BN, rgn);
  EFM_LAUNCH_CHECK("recip");
  const void* Gop = copy_g ? (const void*)gpad : G;
  int rc;
  if (K == 1) {
    TopkEpi<1>::Params ep{gnorm, rgn, qn, sim_th, flags, K, idx_base, nq, ps, pi, col_slices, cg};
    rc = topk_tc_launch<1>(Gop, qpad, mh, cg, g, grid, ep, st);
  } else if (K <= 5) {
    TopkEpi<5>::Params ep{gnorm, rgn, qn, sim_th, flags, K, idx_base, nq, ps, pi, col_slices, cg};
    rc = topk_tc_launch<5>(Gop, qpad, mh, cg, g, grid, ep, st);
  } else {
    TopkEpi<EFM_GAL_MAX_K>::Params ep{gnorm, rgn, qn, sim_th, flags, K, idx_base, nq, ps, pi, col_slices, cg};
    rc = topk_tc_launch<EFM_GAL_MAX_K>(Gop, qpad, mh, cg, g, grid, ep, st);
  }
  if (rc) return rc;
  return launch_topk_merge(ps, pi, parts, nq, K, out_score, out_idx, st);
}

extern "C" int efm_topk_merge(const float* part_score, const int32_t* part_idx, int parts, long nq, int K,
                              float* out_score, int32_t* out_idx, void* stream) {
  EFM_REQUIRE(parts >= 1 && nq >= 0 && K >= 1, EFM_E_ARG, "efm_topk_merge: bad arguments parts=%d nq=%ld K=%d", parts, nq, K);
  if (nq == 0) return EFM_OK;
  EFM_REQUIRE(part_score && part_idx && out_score && out_idx, EFM_E_ARG, "efm_topk_merge: null pointer");
  EFM_REQUIRE(K <= EFM_GAL_MAX_K && nq < (1L << 31), EFM_E_ARG, "efm_topk_merge: K <= %d", EFM_GAL_MAX_K);
  return launch_topk_merge(part_score, part_idx, parts, nq, K, out_score, out_idx, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------ all-pairs cosine
extern "C" size_t efm_cosine_allpairs_workspace_bytes(long M, long N, long D, int dtype) {
  (void)dtype;
  if (M <= 0 || N <= 0 || D <= 0) return 0;
  const size_t Kp = cdiv(D, tc::BK) * tc::BK;
  return align_up((size_t)M * Kp * 2, 256) + align_up((size_t)N * Kp * 2, 256) + align_up((size_t)M * 4, 256) +
         align_up((size_t)N * 4, 256);
}

extern "C" int efm_cosine_allpairs(const void* A, const void* B, long M, long N, long D, int dtype, float* out, long ld_out,
                                   void* workspace, size_t ws_bytes, void* stream) {
  EFM_REQUIRE(M >= 0 && N >= 0 && D > 0 && ld_out >= N, EFM_E_ARG, "efm_cosine_allpairs: bad shape M=%ld N=%ld D=%ld ld=%ld", M, N, D, ld_out);
  if (M == 0 || N == 0) return EFM_OK;
  EFM_REQUIRE(A && B && out && workspace, EFM_E_ARG, "efm_cosine_allpairs: null pointer");
  EFM_REQUIRE(aligned(workspace, 256), EFM_E_ALIGN, "efm_cosine_allpairs: workspace must be 256-byte aligned");
  const int Kp = (int)(cdiv(D, tc::BK) * tc::BK);
  EFM_REQUIRE(Kp <= tc::MAX_NK * tc::BK, EFM_E_UNSUPPORTED, "efm_cosine_allpairs: D=%ld > %d", D, tc::MAX_NK * tc::BK);
  EFM_REQUIRE(ws_bytes >= efm_cosine_allpairs_workspace_bytes(M, N, D, dtype), EFM_E_WORKSPACE, "efm_cosine_allpairs: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  size_t off = 0;
  auto* apad = reinterpret_cast<__nv_bfloat16*>(ws + off); off += align_up((size_t)M * Kp * 2, 256);
  auto* bpad = reinterpret_cast<__nv_bfloat16*>(ws + off); off += align_up((size_t)N * Kp * 2, 256);
  float* rna = reinterpret_cast<float*>(ws + off); off += align_up((size_t)M * 4, 256);
  float* rnb = reinterpret_cast<float*>(ws + off);
  EFM_DISPATCH_DTYPE(dtype, T, {
    pad_rows_kernel<T><<<(unsigned)cdiv(M, 8), 256, 0, st>>>((const T*)A, M, (int)D, Kp, apad, nullptr, rna);
    pad_rows_kernel<T><<<(unsigned)cdiv(N, 8), 256, 0, st>>>((const T*)B, N, (int)D, Kp, bpad, nullptr, rnb);
  });
  EFM_LAUNCH_CHECK("pad_rows");
  tc::GramParams g{};
  g.M = (int)M; g.N = (int)N; g.nk = Kp / tc::BK;
  const int mh = tc::pick_mh(g.M, g.nk);
  const int grid = tc::plan_gram(g, mh, 0, StoreEpi::SMEM_PER_WARP * tc::EPI_WARPS);
  StoreEpi::Params ep{rna, rnb, out, ld_out};
  return tc::launch_gram<StoreEpi>(apad, bpad, mh, g, ep, grid, st);
}
