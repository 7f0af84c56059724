// Box rasterisers: ordered rectangle fill ("last writer wins") as a two-pass scatter that is bit-exact
// with the reference's Python/NumPy loops.
//
//   pass 1 (owner): every rectangle k (in the reference's write order) does atomicMax(owner[pixel], k+1)
//   pass 2 (fill) : every output pixel looks up its owner and streams the feature row / id out (coalesced)
//
// All coordinate arithmetic is IEEE fp64 with explicit round-to-nearest intrinsics (no FMA contraction)
// and C truncation toward zero == Python int():
//   R1 get_box_mask_box_label_word  data_generator_funsd_bert.py:149-186
//   R2 get_box_mask_box_label       data_generator_funsd_bert.py:64-93
//   geometry get_min_max_x_y_w_h    data_generator_funsd_bert.py:49-61
//   R3 KVModel._generate_masks_from_label  inference/kv_model.py:83-148
//   one-hot to_categorical          inference/generic_util.py:94-95 (+ transposes kv_model.py:274-278)
#include "../../include/msau_b200.h"
#include "common.cuh"
#include "prof.cuh"

namespace msau {

__device__ __forceinline__ long long pyint(double v) { return (long long)v; }   // trunc toward zero

__device__ __forceinline__ double block_reduce(double v, bool is_min, double* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) {
    if (tid < o) sh[tid] = is_min ? fmin(sh[tid], sh[tid + o]) : fmax(sh[tid], sh[tid + o]);
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------------------------------- R1/R2 geometry
__global__ void __launch_bounds__(256) geometry_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                        const double* __restrict__ w, const double* __restrict__ h,
                                                        const int32_t* __restrict__ n_chars, const int32_t* __restrict__ page_ptr,
                                                        double* __restrict__ geom) {
  __shared__ double sh[256];
  const int pg = blockIdx.x;
  const int b0 = page_ptr[pg], b1 = page_ptr[pg + 1];
  double mnx = INFINITY, mny = INFINITY, mnw = INFINITY, mnh = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
    mnx = fmin(mnx, x[i]); mny = fmin(mny, y[i]); mnw = fmin(mnw, w[i]); mnh = fmin(mnh, h[i]);
    mxx = fmax(mxx, __dadd_rn(x[i], w[i])); mxy = fmax(mxy, __dadd_rn(y[i], h[i]));
  }
  mnx = block_reduce(mnx, true, sh); mny = block_reduce(mny, true, sh);
  mnw = block_reduce(mnw, true, sh); mnh = block_reduce(mnh, true, sh);
  mxx = block_reduce(mxx, false, sh); mxy = block_reduce(mxy, false, sh);
  if (threadIdx.x == 0) {
    double min_scale = mnw;
    if (n_chars) {
      // builtin sum(): sequential fp64 accumulation in list order (dgfb.py:156-160)
      double acc = 0.0;
      for (int i = b0; i < b1; ++i) acc = __dadd_rn(acc, n_chars[i] != 0 ? __ddiv_rn(w[i], (double)n_chars[i]) : 0.0);
      const double mean = __ddiv_rn(acc, (double)(b1 - b0));
      min_scale = INFINITY;
      for (int i = b0; i < b1; ++i) {
        double r = n_chars[i] != 0 ? __ddiv_rn(w[i], (double)n_chars[i]) : 0.0;
        if (r == 0.0) r = mean;
        min_scale = fmin(min_scale, r);
      }
    }
    double* g = geom + (long)pg * 8;
    g[0] = mnx; g[1] = mny; g[2] = mnw; g[3] = mnh; g[4] = min_scale;
    g[5] = (double)(pyint(__ddiv_rn(__dsub_rn(mxy, mny), mnh)) + 1);
    g[6] = (double)(pyint(__ddiv_rn(__dsub_rn(mxx, mnx), mnw)) + 1);
    g[7] = 0.0;
  }
}

__device__ __forceinline__ int find_page(const int32_t* __restrict__ page_ptr, int n_pages, int box) {
  int lo = 0, hi = n_pages - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (page_ptr[mid] <= box) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------- R1/R2 owner pass
// one warp per box
__global__ void __launch_bounds__(256) feature_owner_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                             const double* __restrict__ w, const double* __restrict__ h,
                                                             const int32_t* __restrict__ page_ptr, int n_pages, int n_boxes,
                                                             const int32_t* __restrict__ char_ptr, const double* __restrict__ geom,
                                                             int use_min_scale, int out_h, int out_w, int32_t* __restrict__ owner) {
  const int box = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (box >= n_boxes) return;
  const int pg = find_page(page_ptr, n_pages, box);
  const double* g = geom + (long)pg * 8;
  const double sx = use_min_scale ? g[4] : g[2];
  const int Hn = min((int)g[5], out_h), Wn = min((int)g[6], out_w);
  const long long nx = pyint(__ddiv_rn(__dsub_rn(x[box], g[0]), sx));
  const long long ny = pyint(__ddiv_rn(__dsub_rn(y[box], g[1]), g[3]));
  const long long nw = max(pyint(__ddiv_rn(w[box], sx)), 1LL);
  const long long nh = max(pyint(__ddiv_rn(h[box], g[3])), 1LL);
  int32_t* own = owner + (long)pg * out_h * out_w;
  const long long y0 = max(ny, 0LL), y1 = min(ny + nh, (long long)Hn);
  if (char_ptr) {
    const int c0 = char_ptr[box], c1 = char_ptr[box + 1];
    const int len = c1 - c0;
    if (len == 0) return;                    // empty ocr text: the char loop does not run
    const long long pcw = max(pyint(__ddiv_rn((double)nw, (double)len)), 1LL);
    for (int j = 0; j < len; ++j) {
      const long long xa = max(nx + pcw * j, 0LL), xb = min(nx + pcw * (j + 1), (long long)Wn);
      if (xa >= xb) continue;
      const int cw = (int)(xb - xa);
      const long long n = (y1 - y0) * cw;
      for (long long e = lane; e < n; e += 32) {
        const long long yy = y0 + e / cw, xx = xa + e % cw;
        atomicMax(own + yy * out_w + xx, c0 + j + 1);
      }
    }
  } else {
    const long long xa = max(nx, 0LL), xb = min(nx + nw, (long long)Wn);
    if (xa >= xb || y0 >= y1) return;
    const int cw = (int)(xb - xa);
    const long long n = (y1 - y0) * cw;
    for (long long e = lane; e < n; e += 32) {
      const long long yy = y0 + e / cw, xx = xa + e % cw;
      atomicMax(own + yy * out_w + xx, box + 1);
    }
  }
}

// ------------------------------------------------------------------------------------------- R1/R2 fill pass
// layout 0: NCHW -- one thread per pixel, channel loop (each store instruction is a coalesced row segment)
__global__ void __launch_bounds__(256) feature_fill_nchw_kernel(const int32_t* __restrict__ owner, const int32_t* __restrict__ row_of,
                                                                 const double* __restrict__ table, int D, long npix_page,
                                                                 long total, float* __restrict__ grid) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long pg = idx / npix_page, p = idx - pg * npix_page;
  const int o = owner[idx];
  float* dst = grid + pg * D * npix_page + p;
  if (o > 0) {
    const double* src = table + (long)row_of[o - 1] * D;
    for (int c = 0; c < D; ++c) dst[(long)c * npix_page] = __double2float_rn(__ldg(src + c));
  } else {
    for (int c = 0; c < D; ++c) dst[(long)c * npix_page] = 0.f;
  }
}

// layout 1: NHWC (pitch Dp = round_up(D,4)) -- one thread per (pixel, channel quad)
__global__ void __launch_bounds__(256) feature_fill_nhwc_kernel(const int32_t* __restrict__ owner, const int32_t* __restrict__ row_of,
                                                                 const double* __restrict__ table, int D, int Dp, long total_px,
                                                                 float* __restrict__ grid) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = Dp >> 2;
  if (idx >= total_px * q) return;
  const long px = idx / q;
  const int c = (int)(idx - px * q) << 2;
  const int o = owner[px];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (o > 0) {
    const double* src = table + (long)row_of[o - 1] * D;
    if (c + 0 < D) v.x = __double2float_rn(__ldg(src + c));
    if (c + 1 < D) v.y = __double2float_rn(__ldg(src + c + 1));
    if (c + 2 < D) v.z = __double2float_rn(__ldg(src + c + 2));
    if (c + 3 < D) v.w = __double2float_rn(__ldg(src + c + 3));
  }
  reinterpret_cast<float4*>(grid)[idx] = v;
}

// layout 2: the feature-table ROW of the owning box / character per pixel (-1 = background) instead of the row's values.
// With a one-hot table whose row r is e_r (np.eye: the chargrid case) this is the id map the structured first layer of the
// network consumes directly (x_layout = 2 of msau_forward), so the dense [D, H, W] grid is never written.
__global__ void __launch_bounds__(256) feature_ids_kernel(const int32_t* __restrict__ owner, const int32_t* __restrict__ row_of,
                                                           long total, int16_t* __restrict__ ids) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int o = owner[idx];
  ids[idx] = o > 0 ? (int16_t)row_of[o - 1] : (int16_t)-1;
}

__global__ void __launch_bounds__(256) label_fill_kernel(const int32_t* __restrict__ owner, const int32_t* __restrict__ labels,
                                                          long total, uint8_t* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int o = owner[idx];
  out[idx] = o > 0 ? (uint8_t)(labels[o - 1] + 1) : (uint8_t)0;
}

__global__ void __launch_bounds__(256) iota_kernel(int32_t* __restrict__ a, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

// ------------------------------------------------------------------------------------------- R3
__global__ void __launch_bounds__(256) kv_geometry_kernel(const double* __restrict__ boxes, const int32_t* __restrict__ page_ptr,
                                                           double* __restrict__ geom3) {
  __shared__ double sh[256];
  __shared__ double med[2];
  const int pg = blockIdx.x;
  const int b0 = page_ptr[pg], b1 = page_ptr[pg + 1];
  const int n = b1 - b0;
  double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
    mnx = fmin(mnx, boxes[4 * i]); mny = fmin(mny, boxes[4 * i + 1]);
    mxx = fmax(mxx, boxes[4 * i + 2]); mxy = fmax(mxy, boxes[4 * i + 3]);
  }
  mnx = block_reduce(mnx, true, sh); mny = block_reduce(mny, true, sh);
  mxx = block_reduce(mxx, false, sh); mxy = block_reduce(mxy, false, sh);
  // np.median(line_heights): rank selection (sorted[k_lo], sorted[k_hi]), mean of the two middles
  const int k_lo = (n - 1) / 2, k_hi = n / 2;
  for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
    const double hi = __dsub_rn(boxes[4 * i + 3], boxes[4 * i + 1]);
    int less = 0, eq = 0;
    for (int j = b0; j < b1; ++j) {
      const double hj = __dsub_rn(boxes[4 * j + 3], boxes[4 * j + 1]);
      less += hj < hi; eq += hj == hi;
    }
    if (less <= k_lo && k_lo < less + eq) med[0] = hi;
    if (less <= k_hi && k_hi < less + eq) med[1] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double median_h = (k_lo == k_hi) ? med[0] : __ddiv_rn(__dadd_rn(med[0], med[1]), 2.0);
    const long long bg_pad = pyint(__dmul_rn(median_h, 3.0));
    const double min_x = __dsub_rn(mnx, (double)bg_pad), min_y = __dsub_rn(mny, (double)bg_pad);
    const double max_x = __dadd_rn(mxx, (double)bg_pad), max_y = __dadd_rn(mxy, (double)bg_pad);
    const double scale = __ddiv_rn(3.0, median_h);
    const double ww = __dsub_rn(max_x, min_x), hh = __dsub_rn(max_y, min_y);
    double* g = geom3 + (long)pg * 8;
    g[0] = min_x; g[1] = min_y; g[2] = scale; g[3] = (double)bg_pad;
    g[4] = (double)pyint(__dmul_rn(__dmul_rn(hh, scale), 1.0));
    g[5] = (double)pyint(__dmul_rn(__dmul_rn(ww, scale), 1.0));
    g[6] = median_h; g[7] = 0.0;
  }
}

// one warp per text line.  owner_c: last character rectangle (global char index + 1);
// owner_l: last line (global line index + 1) whose line rectangle or any character rectangle covers the pixel
__global__ void __launch_bounds__(256) kv_owner_kernel(const double* __restrict__ boxes, const int32_t* __restrict__ page_ptr,
                                                        int n_pages, int n_lines, const int32_t* __restrict__ char_ptr,
                                                        const double* __restrict__ geom3, int out_h, int out_w,
                                                        int32_t* __restrict__ owner_c, int32_t* __restrict__ owner_l,
                                                        int32_t* __restrict__ scaled_boxes) {
  const int line = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (line >= n_lines) return;
  const int pg = find_page(page_ptr, n_pages, line);
  const double* g = geom3 + (long)pg * 8;
  const double scale = g[2];
  const int Hn = min((int)g[4], out_h), Wn = min((int)g[5], out_w);
  const long long x1 = pyint(__dmul_rn(__dmul_rn(__dsub_rn(boxes[4 * line + 0], g[0]), scale), 1.0));
  const long long y1 = pyint(__dmul_rn(__dmul_rn(__dsub_rn(boxes[4 * line + 1], g[1]), scale), 1.0));
  const long long x2 = pyint(__dmul_rn(__dmul_rn(__dsub_rn(boxes[4 * line + 2], g[0]), scale), 1.0));
  const long long y2 = pyint(__dmul_rn(__dmul_rn(__dsub_rn(boxes[4 * line + 3], g[1]), scale), 1.0));
  if (lane == 0 && scaled_boxes) {
    scaled_boxes[4 * line + 0] = (int)x1; scaled_boxes[4 * line + 1] = (int)y1;
    scaled_boxes[4 * line + 2] = (int)x2; scaled_boxes[4 * line + 3] = (int)y2;
  }
  const int c0 = char_ptr[line], c1 = char_ptr[line + 1];
  const int len = c1 - c0;
  if (len == 0) return;
  const long base = (long)pg * out_h * out_w;
  const long long ya = max(y1, 0LL), yb = min(y2, (long long)Hn);
  if (ya >= yb) return;
  {  // line_id_mask[y1:y2, x1:x2] = line_idx + 1   (index within the page)
    const long long xa = max(x1, 0LL), xb = min(x2, (long long)Wn);
    if (xa < xb) {
      const int cw = (int)(xb - xa);
      const long long n = (yb - ya) * cw;
      for (long long e = lane; e < n; e += 32) atomicMax(owner_l + base + (ya + e / cw) * out_w + xa + e % cw, line + 1);
    }
  }
  const double cfw = fmax(__ddiv_rn(__dmul_rn(1.0, (double)(x2 - x1)), (double)len), 1.0);
  double cwd = fmax(__dmul_rn(0.9, cfw), 1.0);
  const double lim = (double)pyint(__dmul_rn((double)(y2 - y1), 1.2));
  if (lim < cwd) cwd = lim;                  // python min(char_w, int(...))
  for (int j = 0; j < len; ++j) {
    const double off = __dadd_rn((double)x1, __dmul_rn((double)j, cfw));
    const long long sx = pyint(off), ex = pyint(__dadd_rn(off, cwd));
    const long long xa = max(sx, 0LL), xb = min(ex, (long long)Wn);
    if (xa >= xb) continue;
    const int cw = (int)(xb - xa);
    const long long n = (yb - ya) * cw;
    for (long long e = lane; e < n; e += 32) {
      const long o = base + (ya + e / cw) * out_w + xa + e % cw;
      atomicMax(owner_c + o, c0 + j + 1);
      atomicMax(owner_l + o, line + 1);
    }
  }
}

__global__ void __launch_bounds__(256) kv_fill_kernel(const int32_t* __restrict__ owner_c, const int32_t* __restrict__ owner_l,
                                                       const int32_t* __restrict__ page_ptr, const int32_t* __restrict__ char_ptr,
                                                       const int32_t* __restrict__ char_ids, int n_lines, long npix_page, long total,
                                                       uint16_t* __restrict__ input_mask, uint16_t* __restrict__ line_mask,
                                                       uint16_t* __restrict__ char_mask) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int pg = (int)(idx / npix_page);
  const int oc = owner_c[idx], ol = owner_l[idx];
  uint16_t iv = 0, cv = 0;
  if (oc > 0) {
    const int gch = oc - 1;
    int lo = 0, hi = n_lines - 1;       // line containing this character
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (char_ptr[mid] <= gch) lo = mid; else hi = mid - 1;
    }
    iv = (uint16_t)char_ids[gch];
    cv = (uint16_t)(gch - char_ptr[lo] + 1);
  }
  input_mask[idx] = iv;
  char_mask[idx] = cv;
  line_mask[idx] = ol > 0 ? (uint16_t)(ol - 1 - page_ptr[pg] + 1) : (uint16_t)0;
}

template <int LAYOUT>
__global__ void __launch_bounds__(256) one_hot_kernel(const uint16_t* __restrict__ ids, long npix_page, long total, int n_token,
                                                       int Dp, float* __restrict__ out) {
  if (LAYOUT == 0) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long pg = idx / npix_page, p = idx - pg * npix_page;
    const int id = ids[idx];
    float* dst = out + pg * n_token * npix_page + p;
    for (int c = 0; c < n_token; ++c) dst[(long)c * npix_page] = (c == id) ? 1.f : 0.f;
  } else {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = Dp >> 2;
    if (idx >= total * q) return;
    const long px = idx / q;
    const int c = (int)(idx - px * q) << 2;
    const int id = ids[px];
    reinterpret_cast<float4*>(out)[idx] =
        make_float4(c == id ? 1.f : 0.f, c + 1 == id ? 1.f : 0.f, c + 2 == id ? 1.f : 0.f, c + 3 == id ? 1.f : 0.f);
  }
}

}  // namespace msau

using namespace msau;

extern "C" int msau_raster_geometry(const double* x, const double* y, const double* w, const double* h, const int32_t* n_chars,
                                    const int32_t* page_ptr, int n_pages, double* geom, void* stream) {
  MSAU_CHECK_ARG(x && y && w && h && page_ptr && geom && n_pages >= 1, "raster_geometry: bad argument");
  count_launch(1);
  geometry_kernel<<<n_pages, 256, 0, (cudaStream_t)stream>>>(x, y, w, h, n_chars, page_ptr, geom);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_raster_features(const double* x, const double* y, const double* w, const double* h, const int32_t* page_ptr,
                                    int n_pages, int n_boxes, const int32_t* char_ptr, const int32_t* char_feat, const int32_t* feat_row,
                                    const double* feat_table, int feat_dim, const double* geom, int use_min_scale, int out_h,
                                    int out_w, int layout, float* grid, int32_t* owner_scratch, void* stream) {
  MSAU_CHECK_ARG(x && y && w && h && page_ptr && feat_table && geom && grid && owner_scratch, "raster_features: null argument");
  MSAU_CHECK_ARG((char_ptr && char_feat) || feat_row, "raster_features: need (char_ptr, char_feat) or feat_row");
  MSAU_CHECK_ARG(n_pages >= 1 && out_h >= 1 && out_w >= 1 && feat_dim >= 1, "raster_features: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const long npix = (long)out_h * out_w, total = npix * n_pages;
  MSAU_CUDA_TRY(cudaMemsetAsync(owner_scratch, 0, sizeof(int32_t) * total, st));
  count_launch(2);
  if (n_boxes > 0)
    feature_owner_kernel<<<cdiv(n_boxes, 8), 256, 0, st>>>(x, y, w, h, page_ptr, n_pages, n_boxes, char_ptr, geom, use_min_scale,
                                                           out_h, out_w, owner_scratch);
  const int32_t* row_of = char_ptr ? char_feat : feat_row;
  // algorithmic bytes: owner map read (4 B / pixel) + the grid written in its stored dtype
  ProfScope ps(layout == 2 ? "feature_ids_kernel" : (layout == 0 ? "feature_fill_nchw_kernel" : "feature_fill_nhwc_kernel"), 0,
               (double)total * (4.0 + (layout == 2 ? 2.0 : 4.0 * (layout == 0 ? feat_dim : round_up(feat_dim, 4)))), st);
  if (layout == 2) {
    feature_ids_kernel<<<cdiv(total, 256), 256, 0, st>>>(owner_scratch, row_of, total, reinterpret_cast<int16_t*>(grid));
  } else if (layout == 0) {
    feature_fill_nchw_kernel<<<cdiv(total, 256), 256, 0, st>>>(owner_scratch, row_of, feat_table, feat_dim, npix, total, grid);
  } else {
    const int Dp = round_up(feat_dim, 4);
    feature_fill_nhwc_kernel<<<cdiv(total * (Dp / 4), 256), 256, 0, st>>>(owner_scratch, row_of, feat_table, feat_dim, Dp, total, grid);
  }
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_raster_labels(const double* x, const double* y, const double* w, const double* h, const int32_t* labels,
                                  const int32_t* page_ptr, int n_pages, int n_boxes, const double* geom, int out_h, int out_w,
                                  uint8_t* label_mask, int32_t* owner_scratch, void* stream) {
  MSAU_CHECK_ARG(x && y && w && h && labels && page_ptr && geom && label_mask && owner_scratch, "raster_labels: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long total = (long)out_h * out_w * n_pages;
  MSAU_CUDA_TRY(cudaMemsetAsync(owner_scratch, 0, sizeof(int32_t) * total, st));
  count_launch(2);
  if (n_boxes > 0)
    feature_owner_kernel<<<cdiv(n_boxes, 8), 256, 0, st>>>(x, y, w, h, page_ptr, n_pages, n_boxes, nullptr, geom, 0, out_h, out_w,
                                                           owner_scratch);
  ProfScope ps("label_fill_kernel", 0, (double)total * 5.0, st);
  label_fill_kernel<<<cdiv(total, 256), 256, 0, st>>>(owner_scratch, labels, total, label_mask);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_raster_kv_geometry(const double* boxes, const int32_t* page_ptr, int n_pages, double* geom3, void* stream) {
  MSAU_CHECK_ARG(boxes && page_ptr && geom3 && n_pages >= 1, "raster_kv_geometry: bad argument");
  count_launch(1);
  kv_geometry_kernel<<<n_pages, 256, 0, (cudaStream_t)stream>>>(boxes, page_ptr, geom3);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_raster_kv(const double* boxes, const int32_t* page_ptr, int n_pages, int n_lines, const int32_t* char_ptr,
                              const int32_t* char_ids, const double* geom3, int out_h, int out_w, uint16_t* input_mask,
                              uint16_t* line_mask, uint16_t* char_mask, int32_t* scaled_boxes, int32_t* owner_scratch, void* stream) {
  MSAU_CHECK_ARG(boxes && page_ptr && char_ptr && char_ids && geom3 && input_mask && line_mask && char_mask && owner_scratch,
                 "raster_kv: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long npix = (long)out_h * out_w, total = npix * n_pages;
  MSAU_CUDA_TRY(cudaMemsetAsync(owner_scratch, 0, sizeof(int32_t) * total * 2, st));
  count_launch(2);
  if (n_lines > 0)
    kv_owner_kernel<<<cdiv(n_lines, 8), 256, 0, st>>>(boxes, page_ptr, n_pages, n_lines, char_ptr, geom3, out_h, out_w, owner_scratch,
                                                      owner_scratch + total, scaled_boxes);
  ProfScope ps("kv_fill_kernel", 0, (double)total * (8.0 + 6.0), st);
  kv_fill_kernel<<<cdiv(total, 256), 256, 0, st>>>(owner_scratch, owner_scratch + total, page_ptr, char_ptr, char_ids, n_lines, npix,
                                                   total, input_mask, line_mask, char_mask);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

extern "C" int msau_one_hot(const uint16_t* ids, int n_pages, int height, int width, int n_token, int layout, float* out,
                            void* stream) {
  MSAU_CHECK_ARG(ids && out && n_pages >= 1 && height >= 1 && width >= 1 && n_token >= 1, "one_hot: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long npix = (long)height * width, total = npix * n_pages;
  count_launch(1);
  ProfScope ps("one_hot_kernel", 0, (double)total * (2.0 + 4.0 * (layout == 0 ? n_token : round_up(n_token, 4))), st);
  if (layout == 0) {
    one_hot_kernel<0><<<cdiv(total, 256), 256, 0, st>>>(ids, npix, total, n_token, 0, out);
  } else {
    const int Dp = round_up(n_token, 4);
    one_hot_kernel<1><<<cdiv(total * (Dp / 4), 256), 256, 0, st>>>(ids, npix, total, n_token, Dp, out);
  }
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}
