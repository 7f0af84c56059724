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
// CCL: (1) every CTA labels one 64 x 64 tile in shared memory -- row runs from warp ballots, vertical links by a
// shared-memory union-find whose root is the smallest raster index (exactly SciPy's ordering key) -- and writes each
// pixel's tile-local root as a global index; (2) a small kernel links the tiles along their borders in global memory
// (atomicMin union-find over root indices only); (3) roots are ranked with a ballot/popc prefix count; (4) every pixel
// looks up the rank of its root; run starts update the bounding boxes (one set of atomics per row run, not per pixel).
#include <limits.h>

#include "../../include/msau_b200.h"
#include "common.cuh"
#include "prof.cuh"

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

// 4 outputs per thread for 1-row windows on rows whose width is a multiple of 4 (the call site's (1, 3) closing, kv_model.py:176):
// sw + 3 byte loads instead of 4 sw, one 32-bit store
__global__ void __launch_bounds__(256) rect_filter_row4_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, long total4,
                                                                int sw, int c0, int is_max) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total4) return;
  const int W4 = W >> 2;
  const long row = q / W4;
  const int x = (int)(q - row * W4) << 2;
  const uint8_t* src = in + row * W;
  int v[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) {
    const int xx = x + c0 + i;
    v[i] = (i < sw + 3 && xx >= 0 && xx < W) ? (int)__ldg(src + xx) : 0;
  }
  uint32_t packed = 0;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    int r = is_max ? 0 : 255;
#pragma unroll
    for (int d = 0; d < 8; ++d)
      if (d < sw) r = is_max ? max(r, v[o + d]) : min(r, v[o + d]);
    packed |= (uint32_t)r << (8 * o);
  }
  reinterpret_cast<uint32_t*>(out)[q] = packed;
}

__global__ void __launch_bounds__(256) class_equals_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long n, int cls) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i + 16 <= n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + i));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t o = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) o |= (((w[k] >> (8 * b)) & 0xffu) == (uint32_t)cls ? 1u : 0u) << (8 * b);
      r[k] = o;
    }
    *reinterpret_cast<uint4*>(out + i) = make_uint4(r[0], r[1], r[2], r[3]);
  } else {
    for (long k = i; k < n && k < i + 16; ++k) out[k] = in[k] == cls ? 1 : 0;
  }
}

// (class_map == cls) closed with a (1, sw) window (kv_model.py:175-176: r_closing(pred_class == c, (1, 3))) in ONE pass.
// SciPy windows (origin 0): [x - sw/2, x - sw/2 + sw - 1], zeros outside the row -- for the dilation input AND for the erosion
// input (the dilated row), which is what clears the closing's border columns.
// A thread owns 16 pixels: one 16-byte load + the words before and after, then everything is bit arithmetic on a 24-bit row
// segment (bit i <-> column x - 4 + i): sw <= 4 keeps every tap inside that segment.  Width must be a multiple of 16.
__global__ void __launch_bounds__(256) class_closing_row_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, long total16,
                                                                 int cls, int sw) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total16) return;
  const int W16 = W >> 4;
  const long row = q / W16;
  const int x = (int)(q - row * W16) << 4;
  const uint8_t* src = in + row * W + x;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
  const uint32_t before = x > 0 ? __ldg(reinterpret_cast<const uint32_t*>(src - 4)) : 0u;
  const uint32_t after = x + 16 < W ? __ldg(reinterpret_cast<const uint32_t*>(src + 16)) : 0u;
  const uint32_t words[6] = {before, v.x, v.y, v.z, v.w, after};
  const uint32_t c = (uint32_t)cls;
  uint32_t eq = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (((words[k] >> (8 * b)) & 0xffu) == c) eq |= 1u << (4 * k + b);
  // columns outside the row hold zeros (never equal to a foreground class by SciPy's constant mode: they are simply "not set")
  uint32_t valid = 0x00ffffffu;
  if (x == 0) valid &= ~0xfu;
  if (x + 16 >= W) valid &= ~0x00f00000u;
  eq &= valid;
  const int c0 = -(sw / 2);
  uint32_t dil = 0;
#pragma unroll
  for (int d = 0; d < 4; ++d)
    if (d < sw) { const int sft = c0 + d; dil |= sft >= 0 ? (eq >> sft) : (eq << -sft); }
  dil &= valid;
  uint32_t ero = 0xffffffffu;
#pragma unroll
  for (int d = 0; d < 4; ++d)
    if (d < sw) { const int sft = c0 + d; ero &= sft >= 0 ? (dil >> sft) : (dil << -sft); }
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t n4 = (ero >> (4 + 4 * k)) & 0xfu;
    o[k] = (n4 & 1u) | ((n4 & 2u) << 7) | ((n4 & 4u) << 14) | ((n4 & 8u) << 21);
  }
  *reinterpret_cast<uint4*>(out + row * W + x) = make_uint4(o[0], o[1], o[2], o[3]);
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

constexpr int CT = 64;                    // tile edge
constexpr int CT_PIX = CT * CT;           // 4096 pixels, 16 per thread

__device__ __forceinline__ int suf_find(const volatile int* L, int a) {
  int p = L[a];
  while (p != a) { a = p; p = L[a]; }
  return a;
}
__device__ __forceinline__ void suf_union(int* L, int a, int b) {
  while (true) {
    a = suf_find(L, a);
    b = suf_find(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(L + b, a);
    if (old == b) return;
    b = old;
  }
}

// (1) tile-local labelling.  grid = (tiles_x, tiles_y, n_maps), 256 threads: warp w owns tile rows w, w + 8, ..., a lane two
// adjacent 32-pixel halves of the row.  L[p] = global raster index of p's tile-local root, -1 for background.
__global__ void __launch_bounds__(256) ccl_tile_kernel(const uint8_t* __restrict__ bin, int32_t* __restrict__ L, int H, int W,
                                                        uint32_t* __restrict__ cand, int Wp) {
  __shared__ int lab[CT_PIX];
  __shared__ uint32_t rowbits[CT][2];
  const long npix = (long)H * W;
  const uint8_t* b = bin + (long)blockIdx.z * npix;
  int32_t* Lm = L + (long)blockIdx.z * npix;
  const int x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pass 1: foreground bits of every row half + run-start labels
  for (int r = warp; r < CT; r += 8) {
    const int gy = y0 + r;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gx = x0 + 32 * h + lane;
      const bool fg = gy < H && gx < W && b[(long)gy * W + gx] != 0;
      const uint32_t bits = __ballot_sync(0xffffffffu, fg);
      if (lane == 0) rowbits[r][h] = bits;
      // run start inside this 32-pixel half: one past the highest background bit below the lane
      const uint32_t z = ~bits & ((1u << lane) - 1u);
      const int start = z ? 32 - __clz(z) : 0;
      lab[r * CT + 32 * h + lane] = fg ? r * CT + 32 * h + start : -1;
    }
  }
  __syncthreads();
  // pass 2: links.  A run that crosses the middle of the row joins its halves; a pixel links to the pixel above unless its left
  // and upper-left neighbours are foreground too (then the left neighbour makes the same connection)
  for (int r = warp; r < CT; r += 8) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t cur = rowbits[r][h];
      const bool fg = (cur >> lane) & 1u;
      if (!fg) continue;
      const int i = r * CT + 32 * h + lane;
      if (h == 1 && lane == 0 && (rowbits[r][0] >> 31)) suf_union(lab, i, i - 1);
      if (r > 0) {
        const uint32_t up = rowbits[r - 1][h];
        if ((up >> lane) & 1u) {
          bool left, upleft;
          if (lane > 0) { left = (cur >> (lane - 1)) & 1u; upleft = (up >> (lane - 1)) & 1u; }
          else if (h == 1) { left = rowbits[r][0] >> 31; upleft = rowbits[r - 1][0] >> 31; }
          else { left = false; upleft = false; }
          if (!(left && upleft)) suf_union(lab, i, i - CT);
        }
      }
    }
  }
  __syncthreads();
  // pass 3: roots -> global indices; the tile-local roots are the only pixels that can still be roots after the border links:
  // their positions go into a bitmap (one word per 32 pixels of a row, rows padded to Wp words) that the ranking kernels walk
  // instead of the whole parent map
  uint32_t* cm = cand + (long)blockIdx.z * H * Wp;
  for (int r = warp; r < CT; r += 8) {
    const int gy = y0 + r;
    if (gy >= H) break;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gx = x0 + 32 * h + lane;
      const int i = r * CT + 32 * h + lane;
      int out = -1;
      bool is_root = false;
      if (gx < W && lab[i] >= 0) {
        const int root = suf_find(lab, i);
        is_root = root == i;
        out = (y0 + (root >> 6)) * W + x0 + (root & (CT - 1));
      }
      const uint32_t bits = __ballot_sync(0xffffffffu, is_root);
      if (gx < W) Lm[(long)gy * W + gx] = out;
      if (lane == 0 && x0 + 32 * h < W) cm[(long)gy * Wp + ((x0 + 32 * h) >> 5)] = bits;
    }
  }
}

// (2) links across tile borders: one thread per pixel of every tile's first column / first row
__global__ void __launch_bounds__(256) ccl_border_kernel(const uint8_t* __restrict__ bin, int32_t* __restrict__ L, int H, int W, int tiles_x,
                                                          int tiles_y) {
  const long npix = (long)H * W;
  const uint8_t* b = bin + (long)blockIdx.z * npix;
  int32_t* Lm = L + (long)blockIdx.z * npix;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;       // [0, tiles_x * H) vertical borders, then tiles_y * W horizontal ones
  const int nv = tiles_x * H;
  int x, y;
  bool vertical;
  if (t < nv) { vertical = true; x = (t / H) * CT; y = t - (t / H) * H; }
  else if (t < nv + tiles_y * W) { vertical = false; const int u = t - nv; y = (u / W) * CT; x = u - (u / W) * W; }
  else return;
  const int p = y * W + x;
  if (!b[p]) return;
  if (vertical) {
    if (x > 0 && b[p - 1]) uf_union(Lm, p, p - 1);
  } else {
    if (y > 0 && b[p - W]) {
      const bool left = x > 0 && b[p - 1], upleft = x > 0 && b[p - W - 1];
      if (!(left && upleft)) uf_union(Lm, p, p - W);
    }
  }
}

// (3) ranking.  A chunk = 32 consecutive words of the candidate bitmap (row-major, so chunk order and bit order are raster
// order); a warp per chunk, a lane per word: the lane keeps the bits whose pixel is still a root (L[p] == p).
__device__ __forceinline__ uint32_t ccl_true_roots(const int32_t* __restrict__ Lm, uint32_t bits, int y, int x0, int W) {
  uint32_t keep = 0;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const int p = y * W + x0 + b;
    if (Lm[p] == p) keep |= 1u << b;
  }
  return keep;
}

__global__ void __launch_bounds__(256) ccl_count_kernel(const int32_t* __restrict__ L, uint32_t* __restrict__ cand, long npix, int H, int W, int Wp,
                                                         int nchunks, int32_t* __restrict__ counts) {
  const int m = blockIdx.y;
  const int chunk = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (chunk >= nchunks) return;
  const long w = (long)chunk * 32 + lane;
  uint32_t keep = 0;
  if (w < (long)H * Wp) {
    const int y = (int)(w / Wp), wx = (int)(w - (long)y * Wp);
    uint32_t* cw = cand + (long)m * H * Wp + w;
    keep = ccl_true_roots(L + (long)m * npix, *cw, y, wx * 32, W);
    *cw = keep;                                   // the rank kernel reads the verified bits
  }
  int c = __popc(keep);
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[(long)m * (nchunks + 1) + chunk] = c;
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
__global__ void __launch_bounds__(256) ccl_rank_kernel(const uint32_t* __restrict__ cand, long npix, int H, int W, int Wp, int nchunks,
                                                        const int32_t* __restrict__ counts, int32_t* __restrict__ labels) {
  const int m = blockIdx.y;
  const int chunk = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (chunk >= nchunks) return;
  const long w = (long)chunk * 32 + lane;
  uint32_t bits = w < (long)H * Wp ? cand[(long)m * H * Wp + w] : 0u;
  // exclusive prefix of the per-lane counts inside the warp
  const int mine = __popc(bits);
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  int rank = counts[(long)m * (nchunks + 1) + chunk] + incl - mine;
  if (!bits) return;
  const int y = (int)(w / Wp), x0 = (int)(w - (long)y * Wp) * 32;
  int32_t* lab = labels + (long)m * npix;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    lab[y * W + x0 + b] = ++rank;
  }
}

__global__ void __launch_bounds__(256) ccl_bbox_init_kernel(int32_t* __restrict__ bboxes, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bboxes[i] = (i & 1) ? 0 : INT_MAX;   // y0, y1, x0, x1 -> min slots start at INT_MAX, max slots at 0
}

// (4) labels[p] = rank of p's root; the first pixel of every row run (within a warp's 32 pixels) updates the bounding box
__global__ void __launch_bounds__(256) ccl_relabel_kernel(const int32_t* __restrict__ L, int32_t* __restrict__ labels, int H, int W,
                                                           long total, int32_t* __restrict__ bboxes, int max_labels) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long npix = (long)H * W;
  const long m = idx < total ? idx / npix : 0;
  const int p = idx < total ? (int)(idx - m * npix) : 0;
  int lab = 0;
  int r = -1;
  if (idx < total) {
    r = L[idx];
    if (r >= 0) {
      const int32_t* Lm = L + m * npix;
      int q = Lm[r];
      while (q != r) { r = q; q = Lm[r]; }
      lab = labels[m * npix + r];      // the root's own entry already holds its final value (ccl_rank_kernel)
      if (r != p) labels[idx] = lab;
    } else {
      labels[idx] = 0;
    }
  }
  if (!bboxes) return;
  // runs: consecutive lanes with the same map, row and (non-zero) label
  const int lane = threadIdx.x & 31;
  const int y = p / W, x = p - y * W;
  const long key = lab ? (m * H + y) * (long)(INT_MAX / 2) + lab : -1 - lane;      // distinct for background lanes
  const long prev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool start = lab != 0 && (lane == 0 || prev != key || x == 0);
  const uint32_t starts = __ballot_sync(0xffffffffu, start || lab == 0);          // a background lane ends a run as well
  if (start && lab <= max_labels) {
    const uint32_t above = starts & ~((2u << lane) - 1u);                          // next boundary above this lane
    int len = (above ? __ffs(above) - 1 : 32) - lane;
    if (x + len > W) len = W - x;                                                  // the run ends with its image row
    int32_t* bb = bboxes + (m * max_labels + lab - 1) * 4;
    atomicMin(bb + 0, y); atomicMax(bb + 1, y + 1);
    atomicMin(bb + 2, x); atomicMax(bb + 3, x + len);
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
  ProfScope ps("rect_filter_kernel", 0, (double)total * 2.0, (cudaStream_t)stream);
  if (size_h == 1 && origin_h == 0 && size_w <= 8 && (width & 3) == 0 && ((uintptr_t)out & 3) == 0)
    rect_filter_row4_kernel<<<cdiv(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(in, out, width, total / 4, size_w, -(size_w / 2) - origin_w,
                                                                                   is_max);
  else
  rect_filter_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, height, width, total, size_h, size_w,
                                                                        -(size_h / 2) - origin_h, -(size_w / 2) - origin_w, is_max);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_class_closing_row(const uint8_t* class_map, uint8_t* out, int n_maps, int height, int width, int cls, int size_w,
                                      void* stream) {
  MSAU_CHECK_ARG(class_map && out && class_map != out && n_maps >= 1 && height >= 1 && width >= 1, "class_closing_row: bad argument");
  MSAU_CHECK_ARG(size_w >= 1 && size_w <= 4 && (width & 15) == 0 && (((uintptr_t)out | (uintptr_t)class_map) & 15) == 0 && cls >= 0 && cls <= 255,
                 "class_closing_row: window of 1..4 columns, width a multiple of 16, 16-byte aligned maps (use class_equals + rect_filter otherwise)");
  const long total = (long)n_maps * height * width;
  count_launch(1);
  ProfScope ps("class_closing_row_kernel", 0, (double)total * 2.0, (cudaStream_t)stream);
  class_closing_row_kernel<<<cdiv(total / 16, 256), 256, 0, (cudaStream_t)stream>>>(class_map, out, width, total / 16, cls, size_w);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_class_equals(const uint8_t* class_map, uint8_t* out, long long n, int cls, void* stream) {
  MSAU_CHECK_ARG(class_map && out && n >= 1, "class_equals: bad argument");
  count_launch(1);
  ProfScope ps("class_equals_kernel", 0, (double)n * 2.0, (cudaStream_t)stream);
  class_equals_kernel<<<cdiv(cdiv(n, 16), 256), 256, 0, (cudaStream_t)stream>>>(class_map, out, n, cls);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_ccl4(const uint8_t* binary, int n_maps, int height, int width, int32_t* labels, int32_t* n_labels,
                         int32_t* bboxes, int max_labels, int32_t* scratch, void* stream) {
  MSAU_CHECK_ARG(binary && labels && n_labels && scratch && n_maps >= 1 && height >= 1 && width >= 1, "ccl4: bad argument");
  MSAU_CHECK_ARG((long)height * width < INT_MAX, "ccl4: map too large");
  cudaStream_t st = (cudaStream_t)stream;
  const long npix = (long)height * width, total = npix * n_maps;
  // scratch: parent map [total] | candidate-root bitmap [n_maps][H][Wp] | chunk counts [n_maps][nchunks + 1]
  const int Wp = cdiv(width, 32);
  const int nchunks = cdiv((long)height * Wp, 32);
  int32_t* L = scratch;
  uint32_t* cand = reinterpret_cast<uint32_t*>(scratch + total);
  int32_t* counts = scratch + total + (long)n_maps * height * Wp;
  const int tiles_x = cdiv(width, CT), tiles_y = cdiv(height, CT);
  MSAU_CHECK_ARG(n_maps <= 65535 && tiles_y <= 65535, "ccl4: at most 65535 maps per call");
  count_launch(7);
  // algorithmic bytes of the labelling as a whole: binary map read (1 B) + int32 label map written (4 B) + the scratch
  // parent map written and read once (4 + 4 B): DESIGN.md section 3
  ProfScope ps("ccl_kernels", 0, (double)total * 13.0, st);
  ccl_tile_kernel<<<dim3(tiles_x, tiles_y, n_maps), 256, 0, st>>>(binary, L, height, width, cand, Wp);
  if (tiles_x > 1 || tiles_y > 1)
    ccl_border_kernel<<<dim3(cdiv((long)tiles_x * height + (long)tiles_y * width, 256), 1, n_maps), 256, 0, st>>>(binary, L, height, width,
                                                                                                                 tiles_x, tiles_y);
  ccl_count_kernel<<<dim3(cdiv(nchunks, 8), n_maps), 256, 0, st>>>(L, cand, npix, height, width, Wp, nchunks, counts);
  ccl_scan_kernel<<<n_maps, 1024, 0, st>>>(counts, nchunks, n_labels);
  ccl_rank_kernel<<<dim3(cdiv(nchunks, 8), n_maps), 256, 0, st>>>(cand, npix, height, width, Wp, nchunks, counts, labels);
  if (bboxes && max_labels > 0) ccl_bbox_init_kernel<<<cdiv((long)n_maps * max_labels * 4, 256), 256, 0, st>>>(bboxes, (long)n_maps * max_labels * 4);
  ccl_relabel_kernel<<<cdiv(total, 256), 256, 0, st>>>(L, labels, height, width, total, (bboxes && max_labels > 0) ? bboxes : nullptr,
                                                      max_labels);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}
