// common.cuh -- shared helpers for libefmtl (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/efmtl.h"

namespace efm {

// ---- error reporting (thread-local message, negative return codes) --------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define EFM_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::efm::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

#define EFM_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::efm::cuda_fail(_e, #expr); \
  } while (0)

#define EFM_LAUNCH_CHECK(name)                                        \
  do {                                                                \
    cudaError_t _e = cudaGetLastError();                              \
    if (_e != cudaSuccess) return ::efm::cuda_fail(_e, "launch " name); \
  } while (0)

inline int num_sms() {
  static int cache[64] = {};                       // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// The kernels of one batch-hard step are launched with programmatic stream serialization: a kernel's CTAs may be
// scheduled while the previous kernel of the stream is still draining (it calls pdl_trigger() at its top), and block in
// pdl_wait() until that kernel has completed and its memory is visible.  Both are no-ops for plain launches.
// EFM_PDL=0 switches the launch attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
struct LaunchCfg {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute at[2];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, int cluster = 1) {
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    unsigned n = 0;
    if (cluster > 1) {
      at[n].id = cudaLaunchAttributeClusterDimension;
      at[n].val.clusterDim.x = (unsigned)cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
      ++n;
    }
    if (pdl && pdl_enabled()) {
      at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    }
    cfg.attrs = at; cfg.numAttrs = n;
  }
};
// kernel names with template commas go in parentheses: EFM_LAUNCH((k<T, V>), grid, block, smem, st, pdl, args...)
#define EFM_LAUNCH(kernel, grid, block, smem, st, pdl, ...)                  \
  do {                                                                       \
    ::efm::LaunchCfg _lc(dim3(grid), dim3(block), (smem), (st), (pdl));      \
    EFM_CUDA(cudaLaunchKernelEx(&_lc.cfg, kernel, __VA_ARGS__));             \
  } while (0)

// ---- storage-type traits ------------------------------------------------------------------------
template <typename T> struct El;
template <> struct El<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct El<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct El<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

// dispatch a generic lambda on the dtype code
#define EFM_DISPATCH_DTYPE(dtype, T, ...)                                   \
  switch (dtype) {                                                          \
    case EFM_F32: { using T = float; __VA_ARGS__; break; }                  \
    case EFM_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }         \
    case EFM_F16: { using T = __half; __VA_ARGS__; break; }                 \
    default: ::efm::set_error("unsupported dtype code %d", (int)(dtype));   \
      return EFM_E_ARG;                                                     \
  }

inline size_t dtype_size(int dtype) { return dtype == EFM_F32 ? 4 : 2; }

// ---- warp helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// monotone float -> uint32 map (a < b  <=>  ord(a) < ord(b) for non-NaN)
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// mined-candidate keys: max-key for hardest positive (largest value, lowest index on ties),
// min-key for hardest negative (smallest value, lowest index on ties)
constexpr unsigned long long KEY_POS_EMPTY = 0ull;
constexpr unsigned long long KEY_NEG_EMPTY = ~0ull;
__device__ __forceinline__ unsigned long long key_pos(float v, int j) {
  return ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)j);
}
__device__ __forceinline__ unsigned long long key_neg(float v, int j) {
  return ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(uint32_t)j;
}

}  // namespace efm
