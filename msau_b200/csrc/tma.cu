#include "tma.cuh"

#include <mutex>

#include "common.cuh"

namespace msau {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool make_tmap_nhwc_f32(CUtensorMap* out, const float* base, int pitch, int W, int H, int B, int box_c, int box_w, int box_h, bool swizzle32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("tma: cuTensorMapEncodeTiled is not available from this driver"); return false; }
  const cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)W * pitch * 4, (cuuint64_t)H * W * pitch * 4};   // bytes, dims 1..3
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tma: cuTensorMapEncodeTiled failed (%d) for pitch %d, %dx%dx%d, box %dx%dx%d", (int)r, pitch, W, H, B, box_c, box_w, box_h);
    return false;
  }
  return true;
}

}  // namespace msau
