// Host side of tma.cuh: tensor-map encoding through the driver entry point (the library does
// not link libcuda; cuTensorMapEncodeTiled is a pure host-side encoder).
#include "tma.cuh"

#include <mutex>

#include "common.cuh"

namespace nlam {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_row_map_bf16(CUtensorMap* map, const void* base, long long rows, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  NLAM_CHECK(enc, "TMA: cuTensorMapEncodeTiled is not available from this driver");
  NLAM_CHECK(((uintptr_t)base) % 16 == 0 && rows > 0, "TMA: bf16 shadow must be 16-byte aligned");
  cuuint64_t dims[2] = {64, (cuuint64_t)rows}, strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NLAM_CHECK(r == CUDA_SUCCESS, "TMA: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace nlam
