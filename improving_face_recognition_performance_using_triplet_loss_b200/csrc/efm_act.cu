// efm_act.cu -- EFM-3 / MFM-2 max-out activation, forward and backward (HBM-bound).
//
// Replaces the 6-op MXNet sequence SliceChannel(3) -> maximum, maximum, minimum, minimum -> Concat of
// lightcnn.py:32-37 (and :22-27,53-58,61-66,123-128; efm_symbol.py:63-77) by one pass:
// per sample the op is three parallel contiguous spans of S = (C/3)*H*W elements
//     y[n, 0:S]  = max3(x[n,0:S], x[n,S:2S], x[n,2S:3S]),  y[n, S:2S] = min3(...)
// so the kernel never maps work per channel plane; it streams 128-bit vectors along the span.
// Algorithmic bytes: fwd (1 + 2/3) * |x|, bwd (1 + 2/3 + 1) * |x|  (MFM-2: 1.5 / 2.5).
#include "common.cuh"

namespace efm {
namespace {

template <int BYTES> struct Raw;
template <> struct Raw<16> { using type = uint4; };
template <> struct Raw<8> { using type = uint2; };
template <> struct Raw<4> { using type = uint32_t; };
template <> struct Raw<2> { using type = uint16_t; };

template <typename T, int VEC> struct Pack {
  using raw_t = typename Raw<sizeof(T) * VEC>::type;
  union { raw_t raw; T v[VEC]; };
  __device__ __forceinline__ Pack() {}
};

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_stream(const T* p) {
  Pack<T, VEC> r;
  r.raw = __ldcs(reinterpret_cast<const typename Pack<T, VEC>::raw_t*>(p));
  return r;
}
template <typename T, int VEC>
__device__ __forceinline__ void st_stream(T* p, const Pack<T, VEC>& v) {
  __stcs(reinterpret_cast<typename Pack<T, VEC>::raw_t*>(p), v.raw);
}
// uint16_t has no __ldcs/__stcs overload
template <> __device__ __forceinline__ Pack<__nv_bfloat16, 1> ld_stream<__nv_bfloat16, 1>(const __nv_bfloat16* p) {
  Pack<__nv_bfloat16, 1> r; r.v[0] = *p; return r;
}
template <> __device__ __forceinline__ Pack<__half, 1> ld_stream<__half, 1>(const __half* p) {
  Pack<__half, 1> r; r.v[0] = *p; return r;
}
template <> __device__ __forceinline__ void st_stream<__nv_bfloat16, 1>(__nv_bfloat16* p, const Pack<__nv_bfloat16, 1>& v) { *p = v.v[0]; }
template <> __device__ __forceinline__ void st_stream<__half, 1>(__half* p, const Pack<__half, 1>& v) { *p = v.v[0]; }

// MXNet mshadow_op::maximum / minimum: IsNan(a) ? a : (a > b ? a : b) -- selection on the stored value so
// NaN payloads and signed zeros come out bit-identical.
template <typename T> __device__ __forceinline__ T mx_max(T a, T b) {
  float fa = El<T>::to_f(a), fb = El<T>::to_f(b);
  return (fa != fa) ? a : (fa > fb ? a : b);
}
template <typename T> __device__ __forceinline__ T mx_min(T a, T b) {
  float fa = El<T>::to_f(a), fb = El<T>::to_f(b);
  return (fa != fa) ? a : (fa < fb ? a : b);
}

template <typename T, int VEC, int MODE>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long S, long N) {
  constexpr int OUTS = (MODE == 3) ? 2 : 1;
  const long SV = S / VEC;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long n = blockIdx.y; n < N; n += gridDim.y) {
    const T* xn = x + n * MODE * S;
    T* yn = y + n * OUTS * S;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < SV; i += stride) {
      const long e = i * VEC;
      Pack<T, VEC> a = ld_stream<T, VEC>(xn + e);
      Pack<T, VEC> b = ld_stream<T, VEC>(xn + S + e);
      if (MODE == 3) {
        Pack<T, VEC> c = ld_stream<T, VEC>(xn + 2 * S + e);
        Pack<T, VEC> hi, lo;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          hi.v[k] = mx_max(mx_max(a.v[k], b.v[k]), c.v[k]);
          lo.v[k] = mx_min(mx_min(a.v[k], b.v[k]), c.v[k]);
        }
        st_stream<T, VEC>(yn + e, hi);
        st_stream<T, VEC>(yn + S + e, lo);
      } else {
        Pack<T, VEC> hi;
#pragma unroll
        for (int k = 0; k < VEC; ++k) hi.v[k] = mx_max(a.v[k], b.v[k]);
        st_stream<T, VEC>(yn + e, hi);
      }
    }
  }
}

// MXNet backward: maximum -> lhs where lhs >= rhs, rhs where lhs < rhs; minimum -> lhs where lhs <= rhs,
// rhs where lhs > rhs (all comparisons false on NaN -> no gradient).
template <typename T, int VEC, int MODE>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                       T* __restrict__ dx, long S, long N) {
  constexpr int OUTS = (MODE == 3) ? 2 : 1;
  const long SV = S / VEC;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long n = blockIdx.y; n < N; n += gridDim.y) {
    const T* xn = x + n * MODE * S;
    const T* gn = dy + n * OUTS * S;
    T* dn = dx + n * MODE * S;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < SV; i += stride) {
      const long e = i * VEC;
      Pack<T, VEC> a = ld_stream<T, VEC>(xn + e);
      Pack<T, VEC> b = ld_stream<T, VEC>(xn + S + e);
      Pack<T, VEC> gmax = ld_stream<T, VEC>(gn + e);
      if (MODE == 3) {
        Pack<T, VEC> c = ld_stream<T, VEC>(xn + 2 * S + e);
        Pack<T, VEC> gmin = ld_stream<T, VEC>(gn + S + e);
        Pack<T, VEC> d0, d1, d2;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float s0 = El<T>::to_f(a.v[k]), s1 = El<T>::to_f(b.v[k]), s2 = El<T>::to_f(c.v[k]);
          const float gx = El<T>::to_f(gmax.v[k]), gm = El<T>::to_f(gmin.v[k]);
          const float m1 = El<T>::to_f(mx_max(a.v[k], b.v[k]));
          const float n1 = El<T>::to_f(mx_min(a.v[k], b.v[k]));
          const float g_m1 = (m1 >= s2) ? gx : 0.f, g2 = (m1 < s2) ? gx : 0.f;
          const float g0 = (s0 >= s1) ? g_m1 : 0.f, g1 = (s0 < s1) ? g_m1 : 0.f;
          const float h_n1 = (n1 <= s2) ? gm : 0.f, h2 = (n1 > s2) ? gm : 0.f;
          const float h0 = (s0 <= s1) ? h_n1 : 0.f, h1 = (s0 > s1) ? h_n1 : 0.f;
          d0.v[k] = El<T>::from_f(g0 + h0);
          d1.v[k] = El<T>::from_f(g1 + h1);
          d2.v[k] = El<T>::from_f(g2 + h2);
        }
        st_stream<T, VEC>(dn + e, d0);
        st_stream<T, VEC>(dn + S + e, d1);
        st_stream<T, VEC>(dn + 2 * S + e, d2);
      } else {
        Pack<T, VEC> d0, d1;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float s0 = El<T>::to_f(a.v[k]), s1 = El<T>::to_f(b.v[k]);
          const float gx = El<T>::to_f(gmax.v[k]);
          d0.v[k] = El<T>::from_f((s0 >= s1) ? gx : 0.f);
          d1.v[k] = El<T>::from_f((s0 < s1) ? gx : 0.f);
        }
        st_stream<T, VEC>(dn + e, d0);
        st_stream<T, VEC>(dn + S + e, d1);
      }
    }
  }
}

template <int VEC> dim3 act_grid(long S, long N) {
  const long SV = S / VEC;
  long gx = cdiv(SV, 256);
  const long cap = 32L * num_sms();           // plenty of CTAs in flight; grid-stride covers the rest
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  long gy = N < 65535 ? N : 65535;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

template <typename T, int MODE, int VEC>
int launch_act_vec(bool bwd, const T* xp, const T* gp, T* op, long N, long S, cudaStream_t st) {
  const size_t bytes = VEC * sizeof(T);
  const bool ok = (S % VEC == 0) && aligned(xp, bytes) && aligned(op, bytes) && (!bwd || aligned(gp, bytes));
  if (!ok) {
    if constexpr (VEC > 1) return launch_act_vec<T, MODE, VEC / 2>(bwd, xp, gp, op, N, S, st);
    set_error("efm_act: pointer not aligned to the element size");
    return EFM_E_ALIGN;
  } else {
    dim3 g = act_grid<VEC>(S, N);
    if (bwd) act_bwd_kernel<T, VEC, MODE><<<g, 256, 0, st>>>(xp, gp, op, S, N);
    else act_fwd_kernel<T, VEC, MODE><<<g, 256, 0, st>>>(xp, op, S, N);
    EFM_LAUNCH_CHECK("efm_act");
    return EFM_OK;
  }
}

template <typename T, int MODE>
int launch_act(bool bwd, const void* x, const void* dy, void* out, long N, long S, cudaStream_t st) {
  constexpr int V16 = 16 / sizeof(T);
  return launch_act_vec<T, MODE, V16>(bwd, static_cast<const T*>(x), static_cast<const T*>(dy),
                                      static_cast<T*>(out), N, S, st);
}

int act_common(bool bwd, int mode, const void* x, const void* dy, void* out, long N, long C, long HW, int dtype,
               void* stream) {
  EFM_REQUIRE(mode == EFM_ACT_EFM3 || mode == EFM_ACT_MFM2, EFM_E_ARG, "efm_act: mode must be 3 (EFM) or 2 (MFM), got %d", mode);
  EFM_REQUIRE(N >= 0 && C > 0 && HW > 0, EFM_E_ARG, "efm_act: bad shape N=%ld C=%ld HW=%ld", N, C, HW);
  EFM_REQUIRE(C % mode == 0, EFM_E_ARG, "efm_act: C=%ld is not divisible by %d (SliceChannel would fail)", C, mode);
  if (N == 0) return EFM_OK;
  EFM_REQUIRE(x && out && (!bwd || dy), EFM_E_ARG, "efm_act: null pointer");
  const long S = (C / mode) * HW;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EFM_DISPATCH_DTYPE(dtype, T, {
    if (mode == 3) return launch_act<T, 3>(bwd, x, dy, out, N, S, st);
    return launch_act<T, 2>(bwd, x, dy, out, N, S, st);
  });
  return EFM_OK;
}

}  // namespace
}  // namespace efm

extern "C" int efm_act_fwd(int mode, const void* x, void* y, long N, long C, long HW, int dtype, void* stream) {
  return efm::act_common(false, mode, x, nullptr, y, N, C, HW, dtype, stream);
}
extern "C" int efm_act_bwd(int mode, const void* x, const void* dy, void* dx, long N, long C, long HW, int dtype,
                           void* stream) {
  return efm::act_common(true, mode, x, dy, dx, N, C, HW, dtype, stream);
}
