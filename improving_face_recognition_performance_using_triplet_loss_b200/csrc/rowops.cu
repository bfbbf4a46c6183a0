// rowops.cu -- HBM-bound row-wise kernels: normalisation (a3), explicit TripletLoss fwd/bwd (a1),
// cosine pairs (a4), reductions, histogram (a5).  One warp per row, 128-bit loads when the row
// stride allows, warp-shuffle reductions; nothing here is GEMM-shaped.
#include "common.cuh"

namespace efm {
namespace {

constexpr int WARPS = 8;

template <typename T, int VEC> struct VecIO {
  static_assert(VEC == 1 || VEC * sizeof(T) == 16, "VEC is 1 or a 128-bit vector");
  static __device__ __forceinline__ void ld(const T* p, float (&o)[VEC]) {
    if constexpr (VEC == 1) {
      o[0] = El<T>::to_f(*p);
    } else {
      union { uint4 raw; T v[VEC]; } u;
      u.raw = __ldg(reinterpret_cast<const uint4*>(p));
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = El<T>::to_f(u.v[k]);
    }
  }
  static __device__ __forceinline__ void st(T* p, const float (&o)[VEC]) {
    if constexpr (VEC == 1) {
      *p = El<T>::from_f(o[0]);
    } else {
      union { uint4 raw; T v[VEC]; } u;
#pragma unroll
      for (int k = 0; k < VEC; ++k) u.v[k] = El<T>::from_f(o[k]);
      *reinterpret_cast<uint4*>(p) = u.raw;
    }
  }
};

// ---------------------------------------------------------------------------------- l2norm
template <typename T, int VEC>
__global__ void __launch_bounds__(WARPS * 32) l2norm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                 float* __restrict__ row_norm, long B, int D) {
  const long row = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const T* xr = x + row * D;
  float ss = 0.f;
  for (int d = lane * VEC; d < D; d += 32 * VEC) {
    float v[VEC];
    VecIO<T, VEC>::ld(xr + d, v);
#pragma unroll
    for (int k = 0; k < VEC; ++k) ss = fmaf(v[k], v[k], ss);
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  if (row_norm && lane == 0) row_norm[row] = nrm;
  if (y) {
    T* yr = y + row * D;
    for (int d = lane * VEC; d < D; d += 32 * VEC) {
      float v[VEC];
      VecIO<T, VEC>::ld(xr + d, v);
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = v[k] / nrm;
      VecIO<T, VEC>::st(yr + d, v);
    }
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(WARPS * 32) l2norm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                 T* __restrict__ dx, long B, int D) {
  const long row = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const T* xr = x + row * D;
  const T* gr = dy + row * D;
  float ss = 0.f, xg = 0.f;
  for (int d = lane * VEC; d < D; d += 32 * VEC) {
    float v[VEC], g[VEC];
    VecIO<T, VEC>::ld(xr + d, v);
    VecIO<T, VEC>::ld(gr + d, g);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { ss = fmaf(v[k], v[k], ss); xg = fmaf(v[k], g[k], xg); }
  }
  ss = warp_sum(ss); xg = warp_sum(xg);
  const float nrm = sqrtf(ss);
  const float ydg = xg / nrm;                  // y . dy
  T* dr = dx + row * D;
  for (int d = lane * VEC; d < D; d += 32 * VEC) {
    float v[VEC], g[VEC];
    VecIO<T, VEC>::ld(xr + d, v);
    VecIO<T, VEC>::ld(gr + d, g);
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = (g[k] - (v[k] / nrm) * ydg) / nrm;
    VecIO<T, VEC>::st(dr + d, v);
  }
}

// ---------------------------------------------------------------------------------- reductions
template <typename T, bool DOT>
__global__ void __launch_bounds__(256) reduce_kernel(const T* __restrict__ x, const T* __restrict__ y, long n,
                                                      float* __restrict__ out) {
  float acc = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float a = El<T>::to_f(x[i]);
    const float b = DOT ? El<T>::to_f(y[i]) : a;
    acc = fmaf(a, b, acc);
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) frob_fwd_kernel(const T* __restrict__ x, const float* __restrict__ sumsq,
                                                        T* __restrict__ y, long n) {
  const float nrm = sqrtf(*sumsq);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = El<T>::from_f(El<T>::to_f(x[i]) / nrm);
}

template <typename T>
__global__ void __launch_bounds__(256) frob_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                        const float* __restrict__ sumsq,
                                                        const float* __restrict__ dot_x_dy, T* __restrict__ dx, long n) {
  const float nrm = sqrtf(*sumsq);
  const float ydg = *dot_x_dy / nrm;           // sum(y * dy), y = x / nrm
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dx[i] = El<T>::from_f((El<T>::to_f(dy[i]) - (El<T>::to_f(x[i]) / nrm) * ydg) / nrm);
}

// ---------------------------------------------------------------------------------- TripletLoss (a1)
template <typename T, int VEC>
__device__ __forceinline__ float triplet_row_sum(const T* ar, const T* pr, const T* nr, int D, int lane) {
  float s = 0.f;
  for (int d = lane * VEC; d < D; d += 32 * VEC) {
    float a[VEC], p[VEC], n[VEC];
    VecIO<T, VEC>::ld(ar + d, a);
    VecIO<T, VEC>::ld(pr + d, p);
    VecIO<T, VEC>::ld(nr + d, n);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float dp = p[k] - a[k], dn = n[k] - a[k];
      s += dp * dp - dn * dn;
    }
  }
  return warp_sum(s);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(WARPS * 32) triplet_fwd_kernel(const T* __restrict__ a, const T* __restrict__ p,
                                                                  const T* __restrict__ n, long B, int D, float margin,
                                                                  float* __restrict__ loss) {
  const long row = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const float s = triplet_row_sum<T, VEC>(a + row * D, p + row * D, n + row * D, D, lane);
  if (lane == 0) loss[row] = fmaxf(s + margin, 0.f);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(WARPS * 32) triplet_bwd_kernel(const T* __restrict__ a, const T* __restrict__ p,
                                                                  const T* __restrict__ n,
                                                                  const float* __restrict__ head_grad, long B, int D,
                                                                  float margin, T* __restrict__ da, T* __restrict__ dp,
                                                                  T* __restrict__ dn) {
  const long row = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const T* ar = a + row * D; const T* pr = p + row * D; const T* nr = n + row * D;
  const float s = triplet_row_sum<T, VEC>(ar, pr, nr, D, lane);
  const bool active = (s + margin) > 0.f;                         // relu'(0) = 0
  const float w = active ? (head_grad ? head_grad[row] : 1.f) : 0.f;
  const float w2 = 2.f * w;
  for (int d = lane * VEC; d < D; d += 32 * VEC) {
    float av[VEC], pv[VEC], nv[VEC], o[VEC];
    VecIO<T, VEC>::ld(ar + d, av);
    VecIO<T, VEC>::ld(pr + d, pv);
    VecIO<T, VEC>::ld(nr + d, nv);
    if (da) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = w2 * (nv[k] - pv[k]);
      VecIO<T, VEC>::st(da + row * D + d, o);
    }
    if (dp) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = w2 * (pv[k] - av[k]);
      VecIO<T, VEC>::st(dp + row * D + d, o);
    }
    if (dn) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = w2 * (av[k] - nv[k]);
      VecIO<T, VEC>::st(dn + row * D + d, o);
    }
  }
}

// ---------------------------------------------------------------------------------- cosine pairs (a4)
template <typename T, int VEC>
__global__ void __launch_bounds__(WARPS * 32) cosine_pairs_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                                   const int32_t* __restrict__ ia,
                                                                   const int32_t* __restrict__ ib, long P, int D,
                                                                   float* __restrict__ out) {
  const long row = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= P) return;
  const int lane = threadIdx.x & 31;
  const T* ar = a + (ia ? (long)ia[row] : row) * D;
  const T* br = b + (ib ? (long)ib[row] : row) * D;
  float dot = 0.f, sa = 0.f, sb = 0.f;
  for (int d = lane * VEC; d < D; d += 32 * VEC) {
    float x[VEC], y[VEC];
    VecIO<T, VEC>::ld(ar + d, x);
    VecIO<T, VEC>::ld(br + d, y);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { dot = fmaf(x[k], y[k], dot); sa = fmaf(x[k], x[k], sa); sb = fmaf(y[k], y[k], sb); }
  }
  dot = warp_sum(dot); sa = warp_sum(sa); sb = warp_sum(sb);
  if (lane == 0) out[row] = dot / (sqrtf(sa) * sqrtf(sb));
}

// ---------------------------------------------------------------------------------- histogram (a5)
// numpy.histogram's uniform-bin fast path in float64, including its edge corrections.
__global__ void __launch_bounds__(256) histogram_kernel(const float* __restrict__ x, long n,
                                                         const double* __restrict__ edges, int bins,
                                                         unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned int sh[];
  for (int b = threadIdx.x; b < bins; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const double first = edges[0], last = edges[bins];
  const double denom = last - first;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    if (!(v >= first) || !(v <= last)) continue;
    long idx = (long)(((v - first) / denom) * (double)bins);
    if (idx == bins) idx -= 1;
    if (v < edges[idx]) idx -= 1;
    if (v >= edges[idx + 1] && idx != bins - 1) idx += 1;
    atomicAdd(&sh[idx], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += blockDim.x)
    if (sh[b]) atomicAdd(&counts[b], (unsigned long long)sh[b]);
}

__global__ void gather_pid_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ pids, long nq,
                                  int32_t* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) out[i] = idx[i] < 0 ? -1 : pids[idx[i]];
}

template <typename T> bool can_vec(long D, std::initializer_list<const void*> ptrs) {
  constexpr int V = 16 / sizeof(T);
  if (D % V) return false;
  for (const void* p : ptrs)
    if (p && !aligned(p, 16)) return false;
  return true;
}

inline unsigned flat_grid(long n) {
  long g = cdiv(n, 256);
  const long cap = 16L * num_sms();
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace efm

using namespace efm;

#define EFM_ROW_CHECK(name, B, D)                                                                     \
  EFM_REQUIRE((B) >= 0 && (D) > 0 && (D) < (1L << 30), EFM_E_ARG, name ": bad shape B=%ld D=%ld", (long)(B), (long)(D))

extern "C" int efm_l2norm_fwd(const void* x, void* y, float* row_norm, long B, long D, int dtype, void* stream) {
  EFM_ROW_CHECK("efm_l2norm_fwd", B, D);
  if (B == 0) return EFM_OK;
  EFM_REQUIRE(x && (y || row_norm), EFM_E_ARG, "efm_l2norm_fwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)cdiv(B, WARPS);
  EFM_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = 16 / sizeof(T);
    if (can_vec<T>(D, {x, y}))
      l2norm_fwd_kernel<T, V><<<grid, WARPS * 32, 0, st>>>((const T*)x, (T*)y, row_norm, B, (int)D);
    else
      l2norm_fwd_kernel<T, 1><<<grid, WARPS * 32, 0, st>>>((const T*)x, (T*)y, row_norm, B, (int)D);
  });
  EFM_LAUNCH_CHECK("l2norm_fwd");
  return EFM_OK;
}

extern "C" int efm_l2norm_bwd(const void* x, const void* dy, void* dx, long B, long D, int dtype, void* stream) {
  EFM_ROW_CHECK("efm_l2norm_bwd", B, D);
  if (B == 0) return EFM_OK;
  EFM_REQUIRE(x && dy && dx, EFM_E_ARG, "efm_l2norm_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)cdiv(B, WARPS);
  EFM_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = 16 / sizeof(T);
    if (can_vec<T>(D, {x, dy, dx}))
      l2norm_bwd_kernel<T, V><<<grid, WARPS * 32, 0, st>>>((const T*)x, (const T*)dy, (T*)dx, B, (int)D);
    else
      l2norm_bwd_kernel<T, 1><<<grid, WARPS * 32, 0, st>>>((const T*)x, (const T*)dy, (T*)dx, B, (int)D);
  });
  EFM_LAUNCH_CHECK("l2norm_bwd");
  return EFM_OK;
}

extern "C" int efm_sumsq(const void* x, long n, int dtype, float* out, void* stream) {
  EFM_REQUIRE(n >= 0 && out, EFM_E_ARG, "efm_sumsq: bad arguments");
  if (n == 0) return EFM_OK;
  EFM_REQUIRE(x, EFM_E_ARG, "efm_sumsq: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EFM_DISPATCH_DTYPE(dtype, T, { reduce_kernel<T, false><<<flat_grid(n), 256, 0, st>>>((const T*)x, nullptr, n, out); });
  EFM_LAUNCH_CHECK("sumsq");
  return EFM_OK;
}

extern "C" int efm_dot(const void* x, const void* y, long n, int dtype, float* out, void* stream) {
  EFM_REQUIRE(n >= 0 && out, EFM_E_ARG, "efm_dot: bad arguments");
  if (n == 0) return EFM_OK;
  EFM_REQUIRE(x && y, EFM_E_ARG, "efm_dot: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EFM_DISPATCH_DTYPE(dtype, T, { reduce_kernel<T, true><<<flat_grid(n), 256, 0, st>>>((const T*)x, (const T*)y, n, out); });
  EFM_LAUNCH_CHECK("dot");
  return EFM_OK;
}

extern "C" int efm_frob_fwd(const void* x, const float* sumsq, void* y, long n, int dtype, void* stream) {
  EFM_REQUIRE(n >= 0, EFM_E_ARG, "efm_frob_fwd: bad n");
  if (n == 0) return EFM_OK;
  EFM_REQUIRE(x && sumsq && y, EFM_E_ARG, "efm_frob_fwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EFM_DISPATCH_DTYPE(dtype, T, { frob_fwd_kernel<T><<<flat_grid(n), 256, 0, st>>>((const T*)x, sumsq, (T*)y, n); });
  EFM_LAUNCH_CHECK("frob_fwd");
  return EFM_OK;
}

extern "C" int efm_frob_bwd(const void* x, const void* dy, const float* sumsq, const float* dot_x_dy, void* dx, long n,
                            int dtype, void* stream) {
  EFM_REQUIRE(n >= 0, EFM_E_ARG, "efm_frob_bwd: bad n");
  if (n == 0) return EFM_OK;
  EFM_REQUIRE(x && dy && sumsq && dot_x_dy && dx, EFM_E_ARG, "efm_frob_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EFM_DISPATCH_DTYPE(dtype, T, {
    frob_bwd_kernel<T><<<flat_grid(n), 256, 0, st>>>((const T*)x, (const T*)dy, sumsq, dot_x_dy, (T*)dx, n);
  });
  EFM_LAUNCH_CHECK("frob_bwd");
  return EFM_OK;
}

extern "C" int efm_triplet_fwd(const void* a, const void* p, const void* n, long B, long D, int dtype, float margin,
                               float* loss, void* stream) {
  EFM_ROW_CHECK("efm_triplet_fwd", B, D);
  if (B == 0) return EFM_OK;
  EFM_REQUIRE(a && p && n && loss, EFM_E_ARG, "efm_triplet_fwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)cdiv(B, WARPS);
  EFM_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = 16 / sizeof(T);
    if (can_vec<T>(D, {a, p, n}))
      triplet_fwd_kernel<T, V><<<grid, WARPS * 32, 0, st>>>((const T*)a, (const T*)p, (const T*)n, B, (int)D, margin, loss);
    else
      triplet_fwd_kernel<T, 1><<<grid, WARPS * 32, 0, st>>>((const T*)a, (const T*)p, (const T*)n, B, (int)D, margin, loss);
  });
  EFM_LAUNCH_CHECK("triplet_fwd");
  return EFM_OK;
}

extern "C" int efm_triplet_bwd(const void* a, const void* p, const void* n, const float* head_grad, long B, long D,
                               int dtype, float margin, void* da, void* dp, void* dn, void* stream) {
  EFM_ROW_CHECK("efm_triplet_bwd", B, D);
  if (B == 0) return EFM_OK;
  EFM_REQUIRE(a && p && n, EFM_E_ARG, "efm_triplet_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)cdiv(B, WARPS);
  EFM_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = 16 / sizeof(T);
    if (can_vec<T>(D, {a, p, n, da, dp, dn}))
      triplet_bwd_kernel<T, V><<<grid, WARPS * 32, 0, st>>>((const T*)a, (const T*)p, (const T*)n, head_grad, B, (int)D,
                                                             margin, (T*)da, (T*)dp, (T*)dn);
    else
      triplet_bwd_kernel<T, 1><<<grid, WARPS * 32, 0, st>>>((const T*)a, (const T*)p, (const T*)n, head_grad, B, (int)D,
                                                             margin, (T*)da, (T*)dp, (T*)dn);
  });
  EFM_LAUNCH_CHECK("triplet_bwd");
  return EFM_OK;
}

static int cosine_pairs_impl(const void* a, const void* b, const int32_t* ia, const int32_t* ib, long P, long D,
                             int dtype, float* out, void* stream) {
  EFM_ROW_CHECK("efm_cosine_pairs", P, D);
  if (P == 0) return EFM_OK;
  EFM_REQUIRE(a && b && out, EFM_E_ARG, "efm_cosine_pairs: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)cdiv(P, WARPS);
  EFM_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = 16 / sizeof(T);
    if (can_vec<T>(D, {a, b}))
      cosine_pairs_kernel<T, V><<<grid, WARPS * 32, 0, st>>>((const T*)a, (const T*)b, ia, ib, P, (int)D, out);
    else
      cosine_pairs_kernel<T, 1><<<grid, WARPS * 32, 0, st>>>((const T*)a, (const T*)b, ia, ib, P, (int)D, out);
  });
  EFM_LAUNCH_CHECK("cosine_pairs");
  return EFM_OK;
}

extern "C" int efm_cosine_pairs(const void* a, const void* b, long P, long D, int dtype, float* out, void* stream) {
  return cosine_pairs_impl(a, b, nullptr, nullptr, P, D, dtype, out, stream);
}

extern "C" int efm_cosine_pairs_indexed(const void* E, const int32_t* ia, const int32_t* ib, long P, long D, int dtype,
                                        float* out, void* stream) {
  EFM_REQUIRE(ia && ib, EFM_E_ARG, "efm_cosine_pairs_indexed: null index pointer");
  return cosine_pairs_impl(E, E, ia, ib, P, D, dtype, out, stream);
}

extern "C" int efm_histogram(const float* x, long n, const double* edges, int bins, unsigned long long* counts,
                             void* stream) {
  EFM_REQUIRE(n >= 0 && bins > 0 && bins <= 4096, EFM_E_ARG, "efm_histogram: bad n=%ld bins=%d", n, bins);
  if (n == 0) return EFM_OK;
  EFM_REQUIRE(x && edges && counts, EFM_E_ARG, "efm_histogram: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  histogram_kernel<<<flat_grid(n), 256, bins * sizeof(unsigned int), st>>>(x, n, edges, bins, counts);
  EFM_LAUNCH_CHECK("histogram");
  return EFM_OK;
}

extern "C" int efm_gather_pid(const int32_t* idx, const int32_t* person_ids, long nq, int32_t* out_pid, void* stream) {
  EFM_REQUIRE(nq >= 0, EFM_E_ARG, "efm_gather_pid: bad nq");
  if (nq == 0) return EFM_OK;
  EFM_REQUIRE(idx && person_ids && out_pid, EFM_E_ARG, "efm_gather_pid: null pointer");
  gather_pid_kernel<<<(unsigned)cdiv(nq, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(idx, person_ids, nq, out_pid);
  EFM_LAUNCH_CHECK("gather_pid");
  return EFM_OK;
}
