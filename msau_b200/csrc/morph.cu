// Post-process kernels: rectangular max/min filters and 4-connected component labelling with SciPy's
// label numbering, bit-exact on label maps.
//
// Reference: inference/morph_util.py:65-84 (r_dilation / r_erosion / r_opening / r_closing =
// scipy.ndimage.maximum_filter / minimum_filter, mode='constant'), :13-22 (connected_components =
// scipy.ndimage.label + find_objects); call site inference/kv_model.py:174-177.
//
// SciPy semantics restated (see oracle/morph.py): window rows for output row i are
// [i - size//2 - origin, i - size//2 - origin + size - 1], samples outside the image are 0; label() is
// 4-connected and numbers components by the raster position of their first pixel.
//
// CCL = union-find with atomicMin linking (the root of a component is its smallest linear index, which
// is exactly SciPy's ordering key) -> flatten -> rank the roots with a ballot/popc prefix count ->
// relabel + bounding boxes.
#include <limits.h>

#include "../../include/msau_b200.h"
#include "common.cuh"

namespace msau {

__global__ void __launch_bounds__(256) rect_filter_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W,
                                                           long total, int sh, int sw, int r0, int c0, int is_max) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long npix = (long)H * W;
  const long m = idx / npix;
  const int p = (int)(idx - m * npix);
  const int y = p / W, x = p - y * W;
  const uint8_t* src = in + m * npix;
  int v = is_max ? 0 : 255;
  for (int dy = 0; dy < sh; ++dy) {
    const int yy = y + r0 + dy;
    for (int dx = 0; dx < sw; ++dx) {
      const int xx = x + c0 + dx;
      const int s = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (int)src[(long)yy * W + xx] : 0;
      v = is_max ? max(v, s) : min(v, s);
    }
  }
  out[idx] = (uint8_t)v;
}

__global__ void __launch_bounds__(256) class_equals_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long n, int cls) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] == cls ? 1 : 0;
}

// ------------------------------------------------------------------------------------------- CCL
__device__ __forceinline__ int uf_find(const int32_t* L, int a) {
  int p = L[a];
  while (p != a) { a = p; p = L[a]; }
  return a;
}

__device__ __forceinline__ void uf_union(int32_t* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }     // a < b: hang the larger root under the smaller
    const int old = atomicMin(L + b, a);
    if (old == b) return;
    b = old;
  }
}

__global__ void __launch_bounds__(256) ccl_init_kernel(const uint8_t* __restrict__ bin, int32_t* __restrict__ L, long total, long npix) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  L[idx] = bin[idx] ? (int32_t)(idx % npix) : -1;
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(const uint8_t* __restrict__ bin, int32_t* __restrict__ L, int H, int W, long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long npix = (long)H * W;
  const long m = idx / npix;
  const int p = (int)(idx - m * npix);
  const uint8_t* b = bin + m * npix;
  if (!b[p]) return;
  int32_t* Lm = L + m * npix;
  const int y = p / W, x = p - y * W;
  if (x > 0 && b[p - 1]) uf_union(Lm, p, p - 1);
  if (y > 0 && b[p - W]) uf_union(Lm, p, p - W);
}

__global__ void __launch_bounds__(256) ccl_flatten_kernel(int32_t* __restrict__ L, long total, long npix) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  if (L[idx] < 0) return;
  int32_t* Lm = L + (idx / npix) * npix;
  const int p = (int)(idx % npix);
  Lm[p] = uf_find(Lm, p);
}

// roots per 1024-pixel chunk
__global__ void __launch_bounds__(1024) ccl_count_kernel(const int32_t* __restrict__ L, long npix, int nchunks, int32_t* __restrict__ counts) {
  const int m = blockIdx.y, chunk = blockIdx.x;
  const long p = (long)chunk * 1024 + threadIdx.x;
  const bool root = p < npix && L[(long)m * npix + p] == (int32_t)p;
  const unsigned bal = __ballot_sync(0xffffffffu, root);
  __shared__ int ws[32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < 32; ++i) s += ws[i];
    counts[(long)m * (nchunks + 1) + chunk] = s;
  }
}

// exclusive scan of the chunk counts of one map (one block per map); last slot = total
__global__ void __launch_bounds__(1024) ccl_scan_kernel(int32_t* __restrict__ counts, int nchunks, int32_t* __restrict__ n_labels) {
  int32_t* c = counts + (long)blockIdx.x * (nchunks + 1);
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nchunks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nchunks ? c[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nchunks) c[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) { c[nchunks] = carry; n_labels[blockIdx.x] = carry; }
}

// labels[root] = rank + 1
__global__ void __launch_bounds__(1024) ccl_rank_kernel(const int32_t* __restrict__ L, long npix, int nchunks,
                                                         const int32_t* __restrict__ counts, int32_t* __restrict__ labels) {
  const int m = blockIdx.y, chunk = blockIdx.x;
  const long p = (long)chunk * 1024 + threadIdx.x;
  const bool root = p < npix && L[(long)m * npix + p] == (int32_t)p;
  const unsigned bal = __ballot_sync(0xffffffffu, root);
  __shared__ int ws[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) ws[wid] = __popc(bal);
  __syncthreads();
  if (root) {
    int before = counts[(long)m * (nchunks + 1) + chunk];
    for (int i = 0; i < wid; ++i) before += ws[i];
    before += __popc(bal & ((1u << lane) - 1u));
    labels[(long)m * npix + p] = before + 1;
  }
}

__global__ void __launch_bounds__(256) ccl_bbox_init_kernel(int32_t* __restrict__ bboxes, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bboxes[i] = (i & 1) ? 0 : INT_MAX;   // y0, y1, x0, x1 -> min slots start at INT_MAX, max slots at 0
}

__global__ void __launch_bounds__(256) ccl_relabel_kernel(const int32_t* __restrict__ L, int32_t* __restrict__ labels, int H, int W,
                                                           long total, int32_t* __restrict__ bboxes, int max_labels) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long npix = (long)H * W;
  const long m = idx / npix;
  const int p = (int)(idx - m * npix);
  const int r = L[idx];
  if (r < 0) { labels[idx] = 0; return; }
  const int lab = labels[m * npix + r];    // the root's own entry already holds its final value
  if (r != p) labels[idx] = lab;
  if (bboxes && lab <= max_labels) {
    int32_t* bb = bboxes + (m * max_labels + lab - 1) * 4;
    const int y = p / W, x = p - y * W;
    atomicMin(bb + 0, y); atomicMax(bb + 1, y + 1);
    atomicMin(bb + 2, x); atomicMax(bb + 3, x + 1);
  }
}

}  // namespace msau

using namespace msau;

extern "C" int msau_rect_filter(const uint8_t* in, uint8_t* out, int n_maps, int height, int width, int size_h, int size_w,
                                int origin_h, int origin_w, int is_max, void* stream) {
  MSAU_CHECK_ARG(in && out && in != out && n_maps >= 1 && height >= 1 && width >= 1, "rect_filter: bad argument");
  MSAU_CHECK_ARG(size_h >= 1 && size_w >= 1, "rect_filter: size must be >= 1");
  // SciPy rejects origins that push the window centre outside the footprint
  MSAU_CHECK_ARG(size_h / 2 + origin_h >= 0 && size_h / 2 + origin_h < size_h && size_w / 2 + origin_w >= 0 && size_w / 2 + origin_w < size_w,
                 "rect_filter: invalid origin");
  const long total = (long)n_maps * height * width;
  count_launch(1);
  rect_filter_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, height, width, total, size_h, size_w,
                                                                        -(size_h / 2) - origin_h, -(size_w / 2) - origin_w, is_max);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_class_equals(const uint8_t* class_map, uint8_t* out, long long n, int cls, void* stream) {
  MSAU_CHECK_ARG(class_map && out && n >= 1, "class_equals: bad argument");
  count_launch(1);
  class_equals_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(class_map, out, n, cls);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_ccl4(const uint8_t* binary, int n_maps, int height, int width, int32_t* labels, int32_t* n_labels,
                         int32_t* bboxes, int max_labels, int32_t* scratch, void* stream) {
  MSAU_CHECK_ARG(binary && labels && n_labels && scratch && n_maps >= 1 && height >= 1 && width >= 1, "ccl4: bad argument");
  MSAU_CHECK_ARG((long)height * width < INT_MAX, "ccl4: map too large");
  cudaStream_t st = (cudaStream_t)stream;
  const long npix = (long)height * width, total = npix * n_maps;
  const int nchunks = cdiv(npix, 1024);
  int32_t* L = scratch;
  int32_t* counts = scratch + total;
  count_launch(7);
  ccl_init_kernel<<<cdiv(total, 256), 256, 0, st>>>(binary, L, total, npix);
  ccl_merge_kernel<<<cdiv(total, 256), 256, 0, st>>>(binary, L, height, width, total);
  ccl_flatten_kernel<<<cdiv(total, 256), 256, 0, st>>>(L, total, npix);
  ccl_count_kernel<<<dim3(nchunks, n_maps), 1024, 0, st>>>(L, npix, nchunks, counts);
  ccl_scan_kernel<<<n_maps, 1024, 0, st>>>(counts, nchunks, n_labels);
  ccl_rank_kernel<<<dim3(nchunks, n_maps), 1024, 0, st>>>(L, npix, nchunks, counts, labels);
  if (bboxes && max_labels > 0) ccl_bbox_init_kernel<<<cdiv((long)n_maps * max_labels * 4, 256), 256, 0, st>>>(bboxes, (long)n_maps * max_labels * 4);
  ccl_relabel_kernel<<<cdiv(total, 256), 256, 0, st>>>(L, labels, height, width, total, (bboxes && max_labels > 0) ? bboxes : nullptr,
                                                      max_labels);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}
