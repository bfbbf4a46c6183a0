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
#include "tc_gram.cuh"

namespace efm {
namespace {

constexpr int WARPS = 8;

struct BhLayout {
  bool use_tc;
  bool xm_f32;          // fp32 mining rows materialised in the workspace
  int Dp;               // K padded to a multiple of 64 (tensor-core path)
  size_t kpos, kneg, sq, nrm, lab, xm, total;
};

BhLayout bh_layout(long B_local, long B_all, long D, int dtype, int flags) {
  BhLayout L{};
  L.Dp = (int)(cdiv(D, tc::BK) * tc::BK);
  L.use_tc = (dtype != EFM_F32 || (flags & EFM_BH_MINE_BF16)) && L.Dp <= tc::MAX_NK * tc::BK;
  L.xm_f32 = !L.use_tc && (dtype != EFM_F32 || (flags & EFM_BH_NORMALIZE));
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.kpos = take((size_t)B_local * 8);
  L.kneg = take((size_t)B_local * 8);
  L.sq = take((size_t)B_all * 4);
  L.nrm = take((size_t)B_all * 4);
  L.lab = take((size_t)B_all * 4 + 1024);          // +slack: the epilogue reads labels/sq in 128-bit vectors
  L.xm = take(L.use_tc ? (size_t)B_all * L.Dp * 2 : (L.xm_f32 ? (size_t)B_all * D * 4 : 0));
  L.total = off;
  return L;
}

// ------------------------------------------------------------------------------------------ prep
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
bh_prep_kernel(const T* __restrict__ E, const int32_t* __restrict__ labels, long B_all, int D, int Dp, int normalize,
               int use_tc, int write_f32, long row_offset, long B_local, float* __restrict__ sq,
               float* __restrict__ nrm, int32_t* __restrict__ lab_ws, __nv_bfloat16* __restrict__ xm_bf16,
               float* __restrict__ xm_f32, unsigned long long* __restrict__ kpos,
               unsigned long long* __restrict__ kneg, float* __restrict__ dE) {
  const long row = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= B_all) return;
  const int lane = threadIdx.x & 31;
  const T* er = E + row * D;
  float n = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = El<T>::to_f(er[d]); ss = fmaf(v, v, ss); }
    n = sqrtf(warp_sum(ss));
  }
  float s2 = 0.f;
  if (use_tc) {
    __nv_bfloat16* xr = xm_bf16 + row * Dp;
    for (int d = lane; d < Dp; d += 32) {
      float v = 0.f;
      if (d < D) { v = El<T>::to_f(er[d]); if (normalize) v = v / n; }
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      xr[d] = b;
      const float vb = __bfloat162float(b);
      s2 = fmaf(vb, vb, s2);
    }
  } else {
    for (int d = lane; d < D; d += 32) {
      float v = El<T>::to_f(er[d]);
      if (normalize) v = v / n;
      if (write_f32) xm_f32[row * D + d] = v;
      s2 = fmaf(v, v, s2);
    }
  }
  s2 = warp_sum(s2);
  if (dE) for (int d = lane; d < D; d += 32) dE[row * D + d] = 0.f;
  if (lane == 0) {
    sq[row] = s2;
    nrm[row] = n;
    if (lab_ws) lab_ws[row] = labels[row];
    const long li = row - row_offset;
    if (kpos && li >= 0 && li < B_local) { kpos[li] = KEY_POS_EMPTY; kneg[li] = KEY_NEG_EMPTY; }
  }
}

// ------------------------------------------------------------------------------------------ mining epilogue (tcgen05 path)
struct MineEpi {
  static constexpr int SMEM_PER_WARP = 0;
  struct Params {
    const float* sq;          // |x_j|^2 of the mining rows, (B_all + pad)
    const int32_t* labels;    // (B_all + pad)
    unsigned long long* kpos; // (B_local)
    unsigned long long* kneg;
    int row_offset;
  };
  struct State {
    int lab_i, gi;
    float pbest, nbest;
    int pidx, nidx;
    bool valid;
  };
  static __device__ __forceinline__ void row_begin(const Params& p, const tc::GramParams& g, int row, State& s) {
    s.valid = row < g.M;
    s.gi = p.row_offset + row;
    s.lab_i = s.valid ? p.labels[s.gi] : 0;
    s.pbest = -INFINITY; s.nbest = INFINITY;
    s.pidx = -1; s.nidx = -1;
  }
  // exact per-element tracking (used when the chunk holds a same-label column, or at the ragged edge)
  static __device__ __forceinline__ void slow(State& s, const float (&v)[32], const int (&lab)[32], int col0, int ncols) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k < ncols) {
        const int col = col0 + k;
        if (lab[k] == s.lab_i) {
          if (col != s.gi && v[k] > s.pbest) { s.pbest = v[k]; s.pidx = col; }
        } else if (v[k] < s.nbest) { s.nbest = v[k]; s.nidx = col; }
      }
    }
  }
  static __device__ __forceinline__ void chunk(const Params& p, const tc::GramParams& g, State& s, int row,
                                               const uint32_t (&acc)[32], int col0, float*) {
    const int ncols = g.N - col0;                    // > 0 guaranteed by the caller
    float v[32];
    int lab[32];
    if (ncols >= 32) {
      const float4* sq4 = reinterpret_cast<const float4*>(p.sq + col0);
      const int4* lb4 = reinterpret_cast<const int4*>(p.labels + col0);
      bool any_same = false;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 s4 = __ldg(sq4 + q);
        const int4 l4 = __ldg(lb4 + q);
        v[4 * q + 0] = fmaf(-2.f, __uint_as_float(acc[4 * q + 0]), s4.x);
        v[4 * q + 1] = fmaf(-2.f, __uint_as_float(acc[4 * q + 1]), s4.y);
        v[4 * q + 2] = fmaf(-2.f, __uint_as_float(acc[4 * q + 2]), s4.z);
        v[4 * q + 3] = fmaf(-2.f, __uint_as_float(acc[4 * q + 3]), s4.w);
        lab[4 * q + 0] = l4.x; lab[4 * q + 1] = l4.y; lab[4 * q + 2] = l4.z; lab[4 * q + 3] = l4.w;
        any_same |= (l4.x == s.lab_i) | (l4.y == s.lab_i) | (l4.z == s.lab_i) | (l4.w == s.lab_i);
      }
      if (any_same) {
        slow(s, v, lab, col0, 32);
      } else {
        float m = v[0];
#pragma unroll
        for (int k = 1; k < 32; ++k) m = fminf(m, v[k]);
        if (m < s.nbest) {
          int kk = 31;
#pragma unroll
          for (int k = 30; k >= 0; --k) kk = (v[k] == m) ? k : kk;
          s.nbest = m; s.nidx = col0 + kk;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const bool in = k < ncols;
        v[k] = in ? fmaf(-2.f, __uint_as_float(acc[k]), p.sq[col0 + k]) : 0.f;
        lab[k] = in ? p.labels[col0 + k] : 0;
      }
      slow(s, v, lab, col0, ncols);
    }
  }
  static __device__ __forceinline__ void row_end(const Params& p, const tc::GramParams&, int row, int, State& s) {
    if (!s.valid) return;
    if (s.pidx >= 0) atomicMax(&p.kpos[row], key_pos(s.pbest, s.pidx));
    if (s.nidx >= 0) atomicMin(&p.kneg[row], key_neg(s.nbest, s.nidx));
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
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
bh_finalize_kernel(const T* __restrict__ E, const float* __restrict__ nrm, int normalize, long B_all, int D,
                   long row_offset, long B_local, float margin, const float* __restrict__ head_grad, int detach,
                   const unsigned long long* __restrict__ kpos, const unsigned long long* __restrict__ kneg,
                   float* __restrict__ loss, int32_t* __restrict__ idx_p, int32_t* __restrict__ idx_n,
                   float* __restrict__ dE) {
  const long li = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (li >= B_local) return;
  const int lane = threadIdx.x & 31;
  int ip, in;
  if (kpos) {                                               // decode the mined keys
    const unsigned long long kp = kpos[li], kn = kneg[li];
    ip = (kp == KEY_POS_EMPTY) ? -1 : (int)(0xffffffffu - (uint32_t)(kp & 0xffffffffull));
    in = (kn == KEY_NEG_EMPTY) ? -1 : (int)(uint32_t)(kn & 0xffffffffull);
    if (lane == 0) { idx_p[li] = ip; idx_n[li] = in; }
  } else {                                                  // backward-only call: indices given
    ip = idx_p[li]; in = idx_n[li];
  }
  if (ip < 0 || in < 0 || ip >= B_all || in >= B_all) { if (lane == 0 && loss) loss[li] = 0.f; return; }
  const long gi = row_offset + li;
  const T* ea = E + gi * D; const T* epp = E + (long)ip * D; const T* en = E + (long)in * D;
  const float na = normalize ? nrm[gi] : 1.f, np_ = normalize ? nrm[ip] : 1.f, nn = normalize ? nrm[in] : 1.f;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    float a = El<T>::to_f(ea[d]), p = El<T>::to_f(epp[d]), n = El<T>::to_f(en[d]);
    if (normalize) { a = a / na; p = p / np_; n = n / nn; }
    const float dp = p - a, dn = n - a;
    s += dp * dp - dn * dn;
  }
  s = warp_sum(s);
  const float l = fmaxf(s + margin, 0.f);
  if (lane == 0 && loss) loss[li] = l;
  if (!dE || !(l > 0.f)) return;
  const float w2 = 2.f * (head_grad ? head_grad[li] : 1.f);
  if (w2 == 0.f) return;
  const bool p_ok = !detach || (ip >= row_offset && ip < row_offset + B_local);
  const bool n_ok = !detach || (in >= row_offset && in < row_offset + B_local);
  for (int d = lane; d < D; d += 32) {
    float a = El<T>::to_f(ea[d]), p = El<T>::to_f(epp[d]), n = El<T>::to_f(en[d]);
    if (normalize) { a = a / na; p = p / np_; n = n / nn; }
    atomicAdd(&dE[gi * D + d], w2 * (n - p));
    if (p_ok) atomicAdd(&dE[(long)ip * D + d], w2 * (p - a));
    if (n_ok) atomicAdd(&dE[(long)in * D + d], w2 * (a - n));
  }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
bh_norm_bwd_kernel(const T* __restrict__ E, const float* __restrict__ nrm, long B_all, int D, float* __restrict__ dE) {
  const long row = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= B_all) return;
  const int lane = threadIdx.x & 31;
  const float n = nrm[row];
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(El<T>::to_f(E[row * D + d]) / n, dE[row * D + d], dot);
  dot = warp_sum(dot);
  for (int d = lane; d < D; d += 32) {
    const float x = El<T>::to_f(E[row * D + d]) / n;
    dE[row * D + d] = (dE[row * D + d] - x * dot) / n;
  }
}

}  // namespace
}  // namespace efm

using namespace efm;

static int bh_launch_mining(const BhLayout& L, const void* E_all, uint8_t* ws, long B_all, long row_offset, long B_local,
                            long D, cudaStream_t st) {
  auto* kpos = reinterpret_cast<unsigned long long*>(ws + L.kpos);
  auto* kneg = reinterpret_cast<unsigned long long*>(ws + L.kneg);
  float* sq = reinterpret_cast<float*>(ws + L.sq);
  int32_t* lab = reinterpret_cast<int32_t*>(ws + L.lab);
  auto* xm_bf16 = reinterpret_cast<__nv_bfloat16*>(ws + L.xm);
  float* xm_f32 = reinterpret_cast<float*>(ws + L.xm);
  if (L.use_tc) {
    tc::GramParams g{};
    g.M = (int)B_local; g.N = (int)B_all; g.nk = L.Dp / tc::BK;
    const int grid = tc::plan_gram(g, 0);
    MineEpi::Params ep{sq, lab, kpos, kneg, (int)row_offset};
    int rc = tc::launch_gram<MineEpi>(xm_bf16 + row_offset * L.Dp, xm_bf16, g, ep, grid, st);
    return rc;
  } else {
    const float* X = L.xm_f32 ? xm_f32 : static_cast<const float*>(E_all);
    const int row_blocks = (int)cdiv(B_local, FM);
    int splits = (int)cdiv(4L * num_sms(), row_blocks);
    const int CT = (int)cdiv(B_all, FN);
    if (splits > CT) splits = CT;
    if (splits < 1) splits = 1;
    bh_mine_f32_kernel<<<dim3(splits, row_blocks), 256, 0, st>>>(X, sq, lab, (int)B_all, (int)row_offset, (int)B_local,
                                                                  (int)D, kpos, kneg);
    EFM_LAUNCH_CHECK("bh_mine_f32");
  }

  return EFM_OK;
}

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
  EFM_REQUIRE(B_all < (1L << 31) - 512 && D < (1 << 20), EFM_E_ARG, "efm_bh_triplet_fwd_bwd: shape too large");
  EFM_REQUIRE(E_all && labels_all && loss && idx_p && idx_n && workspace, EFM_E_ARG, "efm_bh_triplet_fwd_bwd: null pointer");
  EFM_REQUIRE(dtype == EFM_F32 || dtype == EFM_BF16 || dtype == EFM_F16, EFM_E_ARG, "efm_bh_triplet_fwd_bwd: dtype %d", dtype);
  EFM_REQUIRE(aligned(workspace, 256), EFM_E_ALIGN, "efm_bh_triplet_fwd_bwd: workspace must be 256-byte aligned");
  const BhLayout L = bh_layout(B_local, B_all, D, dtype, flags);
  EFM_REQUIRE(ws_bytes >= L.total, EFM_E_WORKSPACE, "efm_bh_triplet_fwd_bwd: workspace %zu < required %zu", ws_bytes, L.total);
  if (B_local == 0) return EFM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto* kpos = reinterpret_cast<unsigned long long*>(ws + L.kpos);
  auto* kneg = reinterpret_cast<unsigned long long*>(ws + L.kneg);
  float* sq = reinterpret_cast<float*>(ws + L.sq);
  float* nrm = reinterpret_cast<float*>(ws + L.nrm);
  int32_t* lab = reinterpret_cast<int32_t*>(ws + L.lab);
  auto* xm_bf16 = reinterpret_cast<__nv_bfloat16*>(ws + L.xm);
  float* xm_f32 = reinterpret_cast<float*>(ws + L.xm);
  const int normalize = (flags & EFM_BH_NORMALIZE) ? 1 : 0;
  const int detach = (flags & EFM_BH_DETACH_GATHERED) ? 1 : 0;
  const unsigned grid_all = (unsigned)cdiv(B_all, WARPS), grid_loc = (unsigned)cdiv(B_local, WARPS);

  EFM_DISPATCH_DTYPE(dtype, T, {
    bh_prep_kernel<T><<<grid_all, WARPS * 32, 0, st>>>((const T*)E_all, labels_all, B_all, (int)D, L.Dp, normalize,
                                                       L.use_tc ? 1 : 0, L.xm_f32 ? 1 : 0, row_offset, B_local, sq, nrm,
                                                       lab, xm_bf16, xm_f32, kpos, kneg, dE_all);
  });
  EFM_LAUNCH_CHECK("bh_prep");

  {
    int rc = bh_launch_mining(L, E_all, ws, B_all, row_offset, B_local, D, st);
    if (rc) return rc;
  }

  EFM_DISPATCH_DTYPE(dtype, T, {
    bh_finalize_kernel<T><<<grid_loc, WARPS * 32, 0, st>>>((const T*)E_all, nrm, normalize, B_all, (int)D, row_offset,
                                                           B_local, margin, head_grad, detach, kpos, kneg, loss, idx_p,
                                                           idx_n, dE_all);
  });
  EFM_LAUNCH_CHECK("bh_finalize");
  if (dE_all && normalize) {
    EFM_DISPATCH_DTYPE(dtype, T, {
      bh_norm_bwd_kernel<T><<<grid_all, WARPS * 32, 0, st>>>((const T*)E_all, nrm, B_all, (int)D, dE_all);
    });
    EFM_LAUNCH_CHECK("bh_norm_bwd");
  }
  return EFM_OK;
}

// Backward for given mined indices (head_grad per anchor): used when the loss vector was returned
// un-reduced and autograd later supplies d(total)/d(loss_i).  workspace: B_all floats (row norms).
extern "C" size_t efm_bh_backward_workspace_bytes(long B_all) { return align_up((size_t)(B_all > 0 ? B_all : 0) * 4, 256) * 2 + 1024; }

extern "C" int efm_bh_backward(const void* E_all, long B_all, long row_offset, long B_local, long D, int dtype,
                               float margin, const float* head_grad, int flags, const int32_t* idx_p,
                               const int32_t* idx_n, float* dE_all, void* workspace, size_t ws_bytes, void* stream) {
  EFM_REQUIRE(B_all > 0 && B_local >= 0 && D > 0 && row_offset >= 0 && row_offset + B_local <= B_all, EFM_E_ARG,
              "efm_bh_backward: bad shape B_all=%ld row_offset=%ld B_local=%ld D=%ld", B_all, row_offset, B_local, D);
  EFM_REQUIRE(E_all && idx_p && idx_n && dE_all && workspace, EFM_E_ARG, "efm_bh_backward: null pointer");
  EFM_REQUIRE(aligned(workspace, 256), EFM_E_ALIGN, "efm_bh_backward: workspace must be 256-byte aligned");
  EFM_REQUIRE(ws_bytes >= efm_bh_backward_workspace_bytes(B_all), EFM_E_WORKSPACE, "efm_bh_backward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* nrm = static_cast<float*>(workspace);
  float* sq = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + align_up((size_t)B_all * 4, 256));
  const int normalize = (flags & EFM_BH_NORMALIZE) ? 1 : 0;
  const int detach = (flags & EFM_BH_DETACH_GATHERED) ? 1 : 0;
  const unsigned grid_all = (unsigned)cdiv(B_all, WARPS), grid_loc = (unsigned)cdiv(B_local, WARPS);
  EFM_CUDA(cudaMemsetAsync(dE_all, 0, (size_t)B_all * D * 4, st));
  if (normalize) {
    EFM_DISPATCH_DTYPE(dtype, T, {
      // reuse the row-norm pass of prep without touching keys / mining rows: a (B_all, D) l2norm stats pass
      bh_prep_kernel<T><<<grid_all, WARPS * 32, 0, st>>>((const T*)E_all, nullptr, B_all, (int)D, (int)D, 1, 0, 0, 0, 0, sq, nrm,
                                                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    });
    EFM_LAUNCH_CHECK("bh_prep(stats)");
  }
  if (B_local > 0) {
    EFM_DISPATCH_DTYPE(dtype, T, {
      bh_finalize_kernel<T><<<grid_loc, WARPS * 32, 0, st>>>((const T*)E_all, nrm, normalize, B_all, (int)D, row_offset,
                                                             B_local, margin, head_grad, detach, nullptr, nullptr, nullptr,
                                                             const_cast<int32_t*>(idx_p), const_cast<int32_t*>(idx_n), dE_all);
    });
    EFM_LAUNCH_CHECK("bh_finalize(bwd)");
  }
  if (normalize) {
    EFM_DISPATCH_DTYPE(dtype, T, {
      bh_norm_bwd_kernel<T><<<grid_all, WARPS * 32, 0, st>>>((const T*)E_all, nrm, B_all, (int)D, dE_all);
    });
    EFM_LAUNCH_CHECK("bh_norm_bwd");
  }
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
  return bh_launch_mining(L, E_all, static_cast<uint8_t*>(workspace), B_all, row_offset, B_local, D,
                          static_cast<cudaStream_t>(stream));
}
