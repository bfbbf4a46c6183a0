// tmap.cu -- TMA tensor-map construction through the driver entry point (no link-time libcuda dependency).
#include <cudaTypedefs.h>

#include <mutex>

#include "tc_gram.cuh"

namespace efm {
namespace tc {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (cuda error %d, query %d)", (int)e, (int)q);
      return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, long rows, long Kp, int box_rows) {
  EFM_REQUIRE(aligned(base, 16), EFM_E_ALIGN, "TMA operand base %p is not 16-byte aligned", base);
  EFM_REQUIRE(Kp % BK == 0 && rows > 0, EFM_E_ARG, "TMA operand shape rows=%ld Kp=%ld", rows, Kp);
  auto enc = get_encode();
  if (!enc) return EFM_E_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EFM_REQUIRE(r == CUDA_SUCCESS, EFM_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return EFM_OK;
}

}  // namespace tc
}  // namespace efm
