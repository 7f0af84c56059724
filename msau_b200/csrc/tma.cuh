// TMA tensor maps for NHWC fp32 activations (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link).
// A map describes a [B, H, W, pitch] fp32 tensor as the 4-D tensor {channel, x, y, page}; a box {8 ch, box_w px, box_h rows, 1}
// lands in shared memory as the dense halo plane [box_h][box_w][8 ch] the conv kernels convert from.  Out-of-image coordinates
// (negative included) are zero-filled by the hardware -- that IS the SAME padding of model/layers/utils.py:5-28.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace msau {
// false (with set_error) when the driver entry point is missing or the encode call rejects the shape
// swizzle32: CU_TENSOR_MAP_SWIZZLE_32B (byte address bit 7 XORed into bit 4 inside shared memory; box_c * 4 must be 32 bytes)
bool make_tmap_nhwc_f32(CUtensorMap* out, const float* base, int pitch, int W, int H, int B, int box_c, int box_w, int box_h,
                        bool swizzle32 = false);
}  // namespace msau
