// Tail of KVModel._extract_value (inference/kv_model.py:178-261) on the device: after the per-class closing + labelling
// (morph.cu) the host picks a few components per class from the bounding-box summary; what it then needs from the full-size
// maps is (a) which text lines each picked component touches -- np.unique(line_mask[labels == k]) (:208, :212) --, (b) the
// new_pred_mask planes (:213, :221) and (c), for lines claimed by more than one field, the range of character indices under
// the field's mask inside the line box (:236-241).  These kernels return exactly those few bytes, so neither the probability
// map nor the label maps cross PCIe.  Integer work, bit-exact.
#include <limits.h>

#include "../../include/msau_b200.h"
#include "common.cuh"

namespace msau {

// one thread per pixel of every map; slot_of[m][k] = global slot of component k of map m, or -1
__global__ void __launch_bounds__(256) kv_select_kernel(const int32_t* __restrict__ labels, const uint16_t* __restrict__ line_mask,
                                                         long npix, long total, const int32_t* __restrict__ slot_of, int max_labels,
                                                         int n_lines, uint8_t* __restrict__ presence, uint8_t* __restrict__ new_mask) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = labels[idx];
  uint8_t hit = 0;
  if (k > 0 && k <= max_labels) {
    const long m = idx / npix;
    const int slot = __ldg(slot_of + m * (max_labels + 1) + k);
    if (slot >= 0) {
      hit = 1;
      const int line = line_mask[idx - m * npix];
      if (line <= n_lines) presence[(long)slot * (n_lines + 1) + line] = 1;   // same value from every writer: no atomics needed
    }
  }
  new_mask[idx] = hit;
}

// one block per query box
__global__ void __launch_bounds__(256) kv_char_range_kernel(const uint16_t* __restrict__ char_mask, const uint8_t* __restrict__ new_mask,
                                                             int H, int W, const int32_t* __restrict__ queries, int32_t* __restrict__ out) {
  const int32_t* q = queries + (long)blockIdx.x * 5;
  const int m = q[0], x1 = q[1], y1 = q[2], x2 = q[3], y2 = q[4];
  const int bw = x2 - x1, bh = y2 - y1;
  int lo = INT_MAX, hi = 0;
  if (bw > 0 && bh > 0) {
    const uint8_t* nm = new_mask + (long)m * H * W;
    for (int e = threadIdx.x; e < bw * bh; e += 256) {
      const int y = y1 + e / bw, x = x1 + e % bw;
      const int c = char_mask[(long)y * W + x];
      if (c > 0 && nm[(long)y * W + x]) { lo = min(lo, c); hi = max(hi, c); }
    }
  }
  __shared__ int slo[8], shi[8];
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
    out[blockIdx.x * 2] = lo;
    out[blockIdx.x * 2 + 1] = hi;
  }
}

}  // namespace msau

using namespace msau;

extern "C" int msau_kv_select_components(const int32_t* labels, const uint16_t* line_mask, int n_maps, int height, int width,
                                         const int32_t* slot_of, int max_labels, int n_slots, int n_lines, uint8_t* presence,
                                         uint8_t* new_mask, void* stream) {
  MSAU_CHECK_ARG(labels && line_mask && slot_of && presence && new_mask, "kv_select_components: null argument");
  MSAU_CHECK_ARG(n_maps >= 1 && height >= 1 && width >= 1 && max_labels >= 1 && n_slots >= 1 && n_lines >= 0,
                 "kv_select_components: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const long npix = (long)height * width, total = npix * n_maps;
  MSAU_CUDA_TRY(cudaMemsetAsync(presence, 0, (size_t)n_slots * (n_lines + 1), st));
  count_launch(1);
  kv_select_kernel<<<cdiv(total, 256), 256, 0, st>>>(labels, line_mask, npix, total, slot_of, max_labels, n_lines, presence, new_mask);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_kv_char_range(const uint16_t* char_mask, const uint8_t* new_mask, int height, int width, const int32_t* queries,
                                  int n_queries, int32_t* out, void* stream) {
  MSAU_CHECK_ARG(char_mask && new_mask && queries && out && height >= 1 && width >= 1 && n_queries >= 1, "kv_char_range: bad argument");
  count_launch(1);
  kv_char_range_kernel<<<n_queries, 256, 0, (cudaStream_t)stream>>>(char_mask, new_mask, height, width, queries, out);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}
