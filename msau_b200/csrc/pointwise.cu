// Pixel-wise kernels: LRN fwd/bwd, 2x2 SAME max-pool fwd/bwd, logits head (NHWC->NCHW, softmax, argmax),
// masked cross-entropy loss + dlogits.  All are HBM-bound streaming kernels: one thread per pixel (or per
// pixel x channel-quad), float4 accesses, no shared-memory reuse needed.
//
// Reference semantics:
//   LRN  : torch.nn.LocalResponseNorm(size=C) as used at model/layers/layers.py:145,161-162
//          y_c = z_c / (1 + (1e-4/C) * sum_{c' in [c-C/2, c+(C-1)/2]} z_c'^2)^0.75
//   pool : pad_2d(...,'pool2d') + MaxPool2d(2,2)   model/model.py:87-94,158-160
//   loss : MSAUWrapper.loss model/model.py:446-459 per page, averaged over pages
//   head : Softmax(dim=1) model/model.py:426-427,437 ; argmax train_chargrid_funsd_msau.py:135, kv_model.py:162
#include "common.cuh"
#include "pointwise.cuh"
#include "prof.cuh"

namespace msau {

static constexpr float kLrnAlpha = 1e-4f;
static constexpr float kLrnBeta = 0.75f;

// ------------------------------------------------------------------------------------------- LRN
// d^-0.75 = rsqrt(d) * sqrt(rsqrt(d)); d is in [1, ~1.1] so this is accurate to ~2 ulp.
__device__ __forceinline__ float pow_m075(float d) {
  const float r = rsqrtf(d);
  return r * sqrtf(r);
}

template <int C>
__global__ void __launch_bounds__(256) lrn_fwd_kernel(const float* __restrict__ z, float* __restrict__ y, long npix) {
  pdl_wait();        // PDL protocol (common.cuh)
  pdl_trigger();
  const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  float v[C];
  const float4* src = reinterpret_cast<const float4*>(z + pix * C);
#pragma unroll
  for (int i = 0; i < C / 4; ++i) {
    const float4 t = __ldg(src + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  constexpr int LO = C / 2, HI = (C - 1) / 2;
  float out[C];
  // sliding window over channels: the first half of the channels only gains terms, the second half
  // only loses them, so the running sum never cancels catastrophically.
  float s = 0.f;
#pragma unroll
  for (int k = 0; k <= HI; ++k) s = fmaf(v[k], v[k], s);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (c > 0) {
      if (c + HI < C) s = fmaf(v[c + HI < C ? c + HI : 0], v[c + HI < C ? c + HI : 0], s);
      if (c - LO - 1 >= 0) s -= v[c - LO - 1 >= 0 ? c - LO - 1 : 0] * v[c - LO - 1 >= 0 ? c - LO - 1 : 0];
    }
    const float d = fmaf(s, kLrnAlpha / C, 1.f);
    out[c] = v[c] * pow_m075(d);
  }
  float4* dst = reinterpret_cast<float4*>(y + pix * C);
#pragma unroll
  for (int i = 0; i < C / 4; ++i) dst[i] = make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
}

// dz_j = g_j d_j^-b - (2ab/C) z_j * sum_{c : j in win(c)} g_c z_c d_c^(-b-1),  {c : j in win(c)} = [j-HI, j+LO]
template <int C>
__global__ void __launch_bounds__(256) lrn_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gy,
                                                       float* __restrict__ gz, long npix) {
  pdl_wait();        // PDL protocol (common.cuh)
  pdl_trigger();
  const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  float v[C], g[C];
  const float4* sz = reinterpret_cast<const float4*>(z + pix * C);
  const float4* sg = reinterpret_cast<const float4*>(gy + pix * C);
#pragma unroll
  for (int i = 0; i < C / 4; ++i) {
    const float4 t = __ldg(sz + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    const float4 u = __ldg(sg + i);
    g[4 * i] = u.x; g[4 * i + 1] = u.y; g[4 * i + 2] = u.z; g[4 * i + 3] = u.w;
  }
  constexpr int LO = C / 2, HI = (C - 1) / 2;
  float pw[C], tt[C];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k <= HI; ++k) s = fmaf(v[k], v[k], s);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (c > 0) {
      if (c + HI < C) s = fmaf(v[c + HI < C ? c + HI : 0], v[c + HI < C ? c + HI : 0], s);
      if (c - LO - 1 >= 0) s -= v[c - LO - 1 >= 0 ? c - LO - 1 : 0] * v[c - LO - 1 >= 0 ? c - LO - 1 : 0];
    }
    const float d = fmaf(s, kLrnAlpha / C, 1.f);
    pw[c] = pow_m075(d);
    tt[c] = g[c] * v[c] * pw[c] / d;
  }
  float out[C];
  // {c : j in win(c)} = [j-HI, j+LO]: second sliding window, over tt
  float u = 0.f;
#pragma unroll
  for (int c = 0; c <= LO && c < C; ++c) u += tt[c];
#pragma unroll
  for (int j = 0; j < C; ++j) {
    if (j > 0) {
      if (j + LO < C) u += tt[j + LO < C ? j + LO : 0];
      if (j - HI - 1 >= 0) u -= tt[j - HI - 1 >= 0 ? j - HI - 1 : 0];
    }
    out[j] = g[j] * pw[j] - (2.f * kLrnAlpha * kLrnBeta / C) * v[j] * u;
  }
  float4* dst = reinterpret_cast<float4*>(gz + pix * C);
#pragma unroll
  for (int i = 0; i < C / 4; ++i) dst[i] = make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
}

// Wide levels (128 / 256 channels: the wrapper-default S=6 model, model/model.py:406): one thread per (pixel, channel),
// 256 / C pixels per block, the two channel windows as differences of an inclusive scan kept in shared memory.  These levels
// are 1/16 and 1/32 of the page resolution, so the kernels are launch-latency sized; registers would not hold a pixel.
__device__ __forceinline__ float seg_scan(float v, float* sh, int c, int C) {
  // inclusive scan of v over the C-thread segment this thread belongs to (Hillis-Steele in shared memory)
  float* seg = sh + (threadIdx.x - c);
  seg[c] = v;
  __syncthreads();
  for (int off = 1; off < C; off <<= 1) {
    const float t = c >= off ? seg[c - off] : 0.f;
    __syncthreads();
    seg[c] += t;
    __syncthreads();
  }
  return seg[c];
}

template <bool BWD>
__global__ void __launch_bounds__(256) lrn_wide_kernel(const float* __restrict__ z, const float* __restrict__ gy, float* __restrict__ out,
                                                        long npix, int C) {
  __shared__ float sh[256], sh2[256];
  const int c = threadIdx.x % C;
  const long pix = (long)blockIdx.x * (256 / C) + threadIdx.x / C;
  const bool live = pix < npix;                     // dead threads still take part in the barriers
  const int LO = C / 2, HI = (C - 1) / 2;
  const float v = live ? __ldg(z + pix * C + c) : 0.f;
  seg_scan(v * v, sh, c, C);
  const float* seg = sh + (threadIdx.x - c);
  const int hi = min(C - 1, c + HI), lo = c - LO - 1;
  const float s = seg[hi] - (lo >= 0 ? seg[lo] : 0.f);
  const float d = fmaf(s, kLrnAlpha / C, 1.f);
  const float pw = pow_m075(d);
  if (!BWD) {
    if (live) out[pix * C + c] = v * pw;
    return;
  }
  const float g = live ? __ldg(gy + pix * C + c) : 0.f;
  seg_scan(g * v * pw / d, sh2, c, C);
  const float* seg2 = sh2 + (threadIdx.x - c);
  const int hi2 = min(C - 1, c + LO), lo2 = c - HI - 1;
  const float u = seg2[hi2] - (lo2 >= 0 ? seg2[lo2] : 0.f);
  if (live) out[pix * C + c] = g * pw - (2.f * kLrnAlpha * kLrnBeta / C) * v * u;
}

// Lane-cooperative LRN for 16..128 channels: C/4 consecutive lanes own one pixel (a float4 of channels each), so a warp reads
// and writes 512 contiguous bytes per instruction whatever C is and nobody holds a whole pixel in registers (the
// thread-per-pixel kernels above need 145 / 255 registers at 32 / 64 channels and their 128 / 256-byte lane stride defeats
// L1).  Both channel windows are half-open prefix differences: with LO = C/2 and HI = C/2 - 1 a channel in the lower half of
// the pixel only needs a prefix value from the lane C/8 lanes above it and a channel in the upper half only one from the lane
// C/8 lanes below it (plus the pixel total), i.e. ONE xor-shuffle partner per lane.
template <int LP>
__device__ __forceinline__ float seg_inclusive(float v, int j) {
  constexpr unsigned full = 0xffffffffu;
#pragma unroll
  for (int off = 1; off < LP; off <<= 1) {
    const float t = __shfl_up_sync(full, v, off, LP);
    if (j >= off) v += t;
  }
  return v;
}

template <int C, bool BWD, int UNR>
__global__ void __launch_bounds__(256) lrn_coop_kernel(const float4* __restrict__ z, const float4* __restrict__ gy, float4* __restrict__ out,
                                                        long nquad) {
  pdl_wait();        // PDL protocol (common.cuh)
  pdl_trigger();
  constexpr int LP = C / 4, HALF = LP / 2;
  constexpr unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int j = lane & (LP - 1);
  const bool low = j < HALF;
  const long nwarps = (long)gridDim.x * 8;
  const long nchunk = (nquad + 32 * UNR - 1) / (32 * UNR);
  for (long ch = (long)blockIdx.x * 8 + (threadIdx.x >> 5); ch < nchunk; ch += nwarps) {
    float4 v[UNR], g[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long q = (ch * UNR + u) * 32 + lane;
      v[u] = q < nquad ? __ldg(z + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (BWD) g[u] = q < nquad ? __ldg(gy + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long q = (ch * UNR + u) * 32 + lane;
      const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      float l[4];
      l[0] = x[0] * x[0]; l[1] = fmaf(x[1], x[1], l[0]); l[2] = fmaf(x[2], x[2], l[1]); l[3] = fmaf(x[3], x[3], l[2]);
      const float inc = seg_inclusive<LP>(l[3], j);
      const float total = __shfl_sync(full, inc, LP - 1, LP);
      float e = __shfl_up_sync(full, inc, 1, LP);
      if (j == 0) e = 0.f;
      // A[k] = prefix of the squares up to channel 4j + k - 1
      const float A[4] = {e, e + l[0], e + l[1], e + l[2]};
      float pw[4], d[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float pa = __shfl_xor_sync(full, A[k], HALF, LP);
        const float sw = low ? pa : total - pa;
        d[k] = fmaf(sw, kLrnAlpha / C, 1.f);
        pw[k] = pow_m075(d[k]);
      }
      float4 r;
      if (!BWD) {
        r = make_float4(x[0] * pw[0], x[1] * pw[1], x[2] * pw[2], x[3] * pw[3]);
      } else {
        const float gg[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
        float m[4];
        m[0] = gg[0] * x[0] * pw[0] / d[0];
        m[1] = m[0] + gg[1] * x[1] * pw[1] / d[1];
        m[2] = m[1] + gg[2] * x[2] * pw[2] / d[2];
        m[3] = m[2] + gg[3] * x[3] * pw[3] / d[3];
        const float inc2 = seg_inclusive<LP>(m[3], j);
        const float total2 = __shfl_sync(full, inc2, LP - 1, LP);
        float e2 = __shfl_up_sync(full, inc2, 1, LP);
        if (j == 0) e2 = 0.f;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float pb = __shfl_xor_sync(full, e2 + m[k], HALF, LP);   // prefix up to channel 4j' + k, inclusive
          const float uu = low ? pb : total2 - pb;
          o[k] = gg[k] * pw[k] - (2.f * kLrnAlpha * kLrnBeta / C) * x[k] * uu;
        }
        r = make_float4(o[0], o[1], o[2], o[3]);
      }
      if (q < nquad) out[q] = r;
    }
  }
}

template <int C, bool BWD>
static void lrn_coop_launch(const float* z, const float* gy, float* out, long npix, cudaStream_t st) {
  constexpr int UNR = BWD ? 2 : 4;
  const long nquad = npix * (C / 4);
  const long nchunk = (nquad + 32 * UNR - 1) / (32 * UNR);
  long blocks = (nchunk + 7) / 8;
  const long cap = (long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  launch_pdl(lrn_coop_kernel<C, BWD, UNR>, dim3((unsigned)blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(z),
             reinterpret_cast<const float4*>(gy), reinterpret_cast<float4*>(out), nquad);
}

template <bool BWD>
static int lrn_dispatch(const float* z, const float* gy, float* out, long npix, int C, int coop, cudaStream_t st) {
  const int grid = cdiv(npix, 256);
  ProfScope ps(BWD ? "lrn_bwd_kernel" : "lrn_fwd_kernel", C, C, 0, 0, (int)(npix >> 10), 0, (double)npix * C * (BWD ? 12 : 6),
               (double)npix * C * 4.0 * (BWD ? 3 : 2), st);
  // measured (B = 16, 512^2 pages, profiles/README.md): the cooperative kernel wins from 16 channels up in the forward pass
  // (25 vs 29 us at 16, 15 vs 23 us at 32, 12 vs 19 us at 64) and from 32 channels up in the backward pass (35 vs 66 us at
  // 32, 21 vs 41 us at 64; 58 vs 54 us at 16); at 8 channels the thread-per-pixel kernels already stream at 5-5.9 TB/s
  const int coop_min = coop == 2 ? 8 : (BWD ? 32 : 16);
  if (coop && C >= coop_min && (C == 8 || C == 16 || C == 32 || C == 64 || C == 128)) {
    switch (C) {
      case 8: lrn_coop_launch<8, BWD>(z, gy, out, npix, st); break;
      case 16: lrn_coop_launch<16, BWD>(z, gy, out, npix, st); break;
      case 32: lrn_coop_launch<32, BWD>(z, gy, out, npix, st); break;
      case 64: lrn_coop_launch<64, BWD>(z, gy, out, npix, st); break;
      default: lrn_coop_launch<128, BWD>(z, gy, out, npix, st); break;
    }
    MSAU_CUDA_TRY(cudaGetLastError());
    return MSAU_OK;
  }
  if (C == 128 || C == 256) {
    lrn_wide_kernel<BWD><<<cdiv(npix, 256 / C), 256, 0, st>>>(z, gy, out, npix, C);
    MSAU_CUDA_TRY(cudaGetLastError());
    return MSAU_OK;
  }
#define MSAU_LRN(CV)                                                             \
  case CV:                                                                       \
    if (BWD) launch_pdl(lrn_bwd_kernel<CV>, dim3(grid), dim3(256), 0, st, z, gy, out, npix); \
    else launch_pdl(lrn_fwd_kernel<CV>, dim3(grid), dim3(256), 0, st, z, out, npix);         \
    break;
  switch (C) {
    MSAU_LRN(4) MSAU_LRN(8) MSAU_LRN(16) MSAU_LRN(32) MSAU_LRN(64)
    default:
      set_error("lrn: unsupported channel count %d (supported 4,8,16,32,64,128,256)", C);
      return MSAU_ERR_UNSUPPORTED;
  }
#undef MSAU_LRN
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

int launch_lrn_fwd(const float* z, float* y, long npix, int C, int coop, cudaStream_t st) { return lrn_dispatch<false>(z, nullptr, y, npix, C, coop, st); }
int launch_lrn_bwd(const float* z, const float* gy, float* gz, long npix, int C, int coop, cudaStream_t st) { return lrn_dispatch<true>(z, gy, gz, npix, C, coop, st); }

// ------------------------------------------------------------------------------------------- pool
__global__ void __launch_bounds__(256) pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W,
                                                        int Ho, int Wo, int C4) {
  pdl_wait();        // PDL protocol (common.cuh)
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * Ho * Wo * C4;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  long r = idx / C4;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int b = (int)(r / Ho);
  const float4* src = reinterpret_cast<const float4*>(x);
  const int y0 = 2 * oy, x0 = 2 * ox;
  float4 m = __ldg(src + (((long)b * H + y0) * W + x0) * C4 + c4);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);   // SAME zero padding takes part in the max
  const float4 v01 = (x0 + 1 < W) ? __ldg(src + (((long)b * H + y0) * W + x0 + 1) * C4 + c4) : zero;
  const float4 v10 = (y0 + 1 < H) ? __ldg(src + (((long)b * H + y0 + 1) * W + x0) * C4 + c4) : zero;
  const float4 v11 = (y0 + 1 < H && x0 + 1 < W) ? __ldg(src + (((long)b * H + y0 + 1) * W + x0 + 1) * C4 + c4) : zero;
  m.x = fmaxf(fmaxf(m.x, v01.x), fmaxf(v10.x, v11.x));
  m.y = fmaxf(fmaxf(m.y, v01.y), fmaxf(v10.y, v11.y));
  m.z = fmaxf(fmaxf(m.z, v01.z), fmaxf(v10.z, v11.z));
  m.w = fmaxf(fmaxf(m.w, v01.w), fmaxf(v10.w, v11.w));
  reinterpret_cast<float4*>(y)[idx] = m;
}

__device__ __forceinline__ void pool_route(float a, float b, float c, float d, float g, float& ga, float& gb, float& gc, float& gd) {
  // torch max_pool2d backward: the FIRST maximum in row-major window order receives the gradient.
  // b/c/d may be the SAME zero pad (then their gradient is dropped by the caller's bounds check).
  int k = 0; float m = a;
  if (b > m) { m = b; k = 1; }
  if (c > m) { m = c; k = 2; }
  if (d > m) { m = d; k = 3; }
  ga = k == 0 ? g : 0.f; gb = k == 1 ? g : 0.f; gc = k == 2 ? g : 0.f; gd = k == 3 ? g : 0.f;
}

// relu_mask: x is the output of a ReLU and this kernel is the last writer of its gradient, so the gradient through that ReLU
// (g * (x > 0), what relu_mask_kernel would do in a separate pass) is applied here, where x is in registers anyway
__global__ void __launch_bounds__(256) pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                                        int B, int H, int W, int Ho, int Wo, int C4, int accumulate, int relu_mask) {
  pdl_wait();        // PDL protocol (common.cuh)
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * Ho * Wo * C4;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  long r = idx / C4;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int b = (int)(r / Ho);
  const float4* src = reinterpret_cast<const float4*>(x);
  const int y0 = 2 * oy, x0 = 2 * ox;
  const bool hx = x0 + 1 < W, hy = y0 + 1 < H;
  const long i00 = (((long)b * H + y0) * W + x0) * C4 + c4;
  const long i01 = i00 + C4, i10 = i00 + (long)W * C4, i11 = i10 + C4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 a = __ldg(src + i00);
  const float4 v01 = hx ? __ldg(src + i01) : zero;
  const float4 v10 = hy ? __ldg(src + i10) : zero;
  const float4 v11 = (hx && hy) ? __ldg(src + i11) : zero;
  const float4 g = __ldg(reinterpret_cast<const float4*>(gy) + idx);
  float4 g00, g01, g10, g11;
  pool_route(a.x, v01.x, v10.x, v11.x, g.x, g00.x, g01.x, g10.x, g11.x);
  pool_route(a.y, v01.y, v10.y, v11.y, g.y, g00.y, g01.y, g10.y, g11.y);
  pool_route(a.z, v01.z, v10.z, v11.z, g.z, g00.z, g01.z, g10.z, g11.z);
  pool_route(a.w, v01.w, v10.w, v11.w, g.w, g00.w, g01.w, g10.w, g11.w);
  float4* dst = reinterpret_cast<float4*>(gx);
  if (accumulate) {
    // all four previous values first: interleaved load-add-store pairs on the same pointer serialise into four round trips
    const float4 o00 = dst[i00];
    const float4 o01 = hx ? dst[i01] : zero;
    const float4 o10 = hy ? dst[i10] : zero;
    const float4 o11 = (hx && hy) ? dst[i11] : zero;
    g00.x += o00.x; g00.y += o00.y; g00.z += o00.z; g00.w += o00.w;
    g01.x += o01.x; g01.y += o01.y; g01.z += o01.z; g01.w += o01.w;
    g10.x += o10.x; g10.y += o10.y; g10.z += o10.z; g10.w += o10.w;
    g11.x += o11.x; g11.y += o11.y; g11.z += o11.z; g11.w += o11.w;
  }
  if (relu_mask) {
    g00.x = a.x > 0.f ? g00.x : 0.f; g00.y = a.y > 0.f ? g00.y : 0.f; g00.z = a.z > 0.f ? g00.z : 0.f; g00.w = a.w > 0.f ? g00.w : 0.f;
    g01.x = v01.x > 0.f ? g01.x : 0.f; g01.y = v01.y > 0.f ? g01.y : 0.f; g01.z = v01.z > 0.f ? g01.z : 0.f; g01.w = v01.w > 0.f ? g01.w : 0.f;
    g10.x = v10.x > 0.f ? g10.x : 0.f; g10.y = v10.y > 0.f ? g10.y : 0.f; g10.z = v10.z > 0.f ? g10.z : 0.f; g10.w = v10.w > 0.f ? g10.w : 0.f;
    g11.x = v11.x > 0.f ? g11.x : 0.f; g11.y = v11.y > 0.f ? g11.y : 0.f; g11.z = v11.z > 0.f ? g11.z : 0.f; g11.w = v11.w > 0.f ? g11.w : 0.f;
  }
  dst[i00] = g00;
  if (hx) dst[i01] = g01;
  if (hy) dst[i10] = g10;
  if (hx && hy) dst[i11] = g11;
}

int launch_pool_fwd(const float* x, float* y, int B, int H, int W, int C, cudaStream_t st) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long total = (long)B * Ho * Wo * (C / 4);
  ProfScope ps("pool_fwd_kernel", 0, ((double)B * H * W + (double)B * Ho * Wo) * C * 4.0, st);
  launch_pdl(pool_fwd_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, x, y, B, H, W, Ho, Wo, C / 4);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

int launch_pool_bwd(const float* x, const float* gy, float* gx, int B, int H, int W, int C, int accumulate, int relu_mask, cudaStream_t st) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long total = (long)B * Ho * Wo * (C / 4);
  ProfScope ps("pool_bwd_kernel", C, C, 0, 0, W, accumulate, 0, ((double)B * H * W * (accumulate ? 3 : 2) + (double)B * Ho * Wo) * C * 4.0, st);
  launch_pdl(pool_bwd_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, x, gy, gx, B, H, W, Ho, Wo, C / 4, accumulate, relu_mask);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// ------------------------------------------------------------------------------------------- misc streaming
__global__ void __launch_bounds__(256) add_kernel(float4* __restrict__ dst, const float4* __restrict__ src, long n4, int accumulate) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = __ldg(src + i);
  if (accumulate) { const float4 o = dst[i]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
  dst[i] = v;
}

// g <- g * (y > 0): the gradient through a ReLU, materialised once for all of its consumers
__global__ void __launch_bounds__(256) relu_mask_kernel(float4* __restrict__ g, const float4* __restrict__ y, long n4) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = g[i];
  const float4 m = __ldg(y + i);
  v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
  g[i] = v;
}

int launch_relu_mask(float* g, const float* y, long n, cudaStream_t st) {
  ProfScope ps("relu_mask_kernel", 0, 0, 0, 0, (int)(n >> 20), 0, 0, (double)n * 4.0 * 3, st);
  relu_mask_kernel<<<cdiv(n / 4, 256), 256, 0, st>>>(reinterpret_cast<float4*>(g), reinterpret_cast<const float4*>(y), n / 4);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

int launch_add(float* dst, const float* src, long n, int accumulate, cudaStream_t st) {
  ProfScope ps("add_kernel", 0, (double)n * 4.0 * (accumulate ? 3 : 2), st);
  add_kernel<<<cdiv(n / 4, 256), 256, 0, st>>>(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), n / 4, accumulate);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// per-channel sum over pixels (bias gradient of the transposed conv): out[c] += sum_p g[p][c], c < c_lim
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, long npix, int C, int c_lim, float* __restrict__ out) {
  // thread t owns channel (t % C) of pixels t / C, t / C + 256 / C, ...   (C divides 256 for C in 8..64)
  const int c = threadIdx.x % C;
  const int ppb = 256 / C;
  float s = 0.f;
  for (long pix = (long)blockIdx.x * ppb + threadIdx.x / C; pix < npix; pix += (long)gridDim.x * ppb) s += __ldg(g + pix * C + c);
  __shared__ float sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int k = 0; k < ppb; ++k) t += sh[k * C + threadIdx.x];
    if (threadIdx.x < c_lim) atomicAdd(out + threadIdx.x, t);
  }
}

int launch_colsum(const float* g, long npix, int C, int c_lim, float* out, cudaStream_t st) {
  MSAU_CHECK_ARG(C >= 4 && C <= 256 && 256 % C == 0, "colsum: unsupported pitch %d", C);
  const int ppb = 256 / C;
  int grid = cdiv(npix, (long)ppb * 16);
  if (grid > 4 * sm_count()) grid = 4 * sm_count();
  if (grid < 1) grid = 1;
  ProfScope ps("colsum_kernel", 0, (double)npix * C * 4.0, st);
  colsum_kernel<<<grid, 256, 0, st>>>(g, npix, C, c_lim, out);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// ------------------------------------------------------------------------------------------- head
// logits NHWC (pitch P = 8 / 16 / 32, n_class <= P used) -> NCHW logits / softmax probabilities / uint8 argmax.
// One thread per pixel: reads P * 4 B, writes n_class strided planes (coalesced across the warp).
template <int P>
__device__ __forceinline__ void load_pixel(const float* __restrict__ src, long idx, float* v) {
  const float4* s4 = reinterpret_cast<const float4*>(src + idx * P);
#pragma unroll
  for (int i = 0; i < P / 4; ++i) {
    const float4 t = __ldg(s4 + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

template <int P>
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ lg, int n_class, long npix_per_page, int B,
                                                    float* __restrict__ logits_nchw, float* __restrict__ probs_nchw,
                                                    uint8_t* __restrict__ argmax) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix_per_page * B) return;
  const int b = (int)(idx / npix_per_page);
  const long p = idx - (long)b * npix_per_page;
  float v[P];
  load_pixel<P>(lg, idx, v);
  float m = v[0]; int am = 0;
#pragma unroll
  for (int k = 1; k < P; ++k)
    if (k < n_class && v[k] > m) { m = v[k]; am = k; }   // first maximum wins (numpy / torch argmax)
  if (argmax) argmax[idx] = (uint8_t)am;
  float* lo = logits_nchw ? logits_nchw + (long)b * n_class * npix_per_page + p : nullptr;
  if (lo) {
#pragma unroll
    for (int k = 0; k < P; ++k)
      if (k < n_class) lo[k * npix_per_page] = v[k];
  }
  if (probs_nchw) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < P; ++k) { v[k] = k < n_class ? expf(v[k] - m) : 0.f; s += v[k]; }
    const float inv = 1.f / s;
    float* po = probs_nchw + (long)b * n_class * npix_per_page + p;
#pragma unroll
    for (int k = 0; k < P; ++k)
      if (k < n_class) po[k * npix_per_page] = v[k] * inv;
  }
}

int launch_head(const float* lg, int P, int n_class, int B, long npix_per_page, float* logits_nchw, float* probs_nchw, uint8_t* argmax, cudaStream_t st) {
  MSAU_CHECK_ARG((P == 8 || P == 16 || P == 32) && n_class <= P, "head: logits pitch must be 8, 16 or 32 and n_class <= pitch (got %d, %d)", P, n_class);
  ProfScope ps("head_kernel", 0, (double)npix_per_page * B * (4.0 * P + 4.0 * n_class * ((logits_nchw ? 1 : 0) + (probs_nchw ? 1 : 0)) + (argmax ? 1 : 0)), st);
  const int grid = cdiv(npix_per_page * B, 256);
  if (P == 8) head_kernel<8><<<grid, 256, 0, st>>>(lg, n_class, npix_per_page, B, logits_nchw, probs_nchw, argmax);
  else if (P == 16) head_kernel<16><<<grid, 256, 0, st>>>(lg, n_class, npix_per_page, B, logits_nchw, probs_nchw, argmax);
  else head_kernel<32><<<grid, 256, 0, st>>>(lg, n_class, npix_per_page, B, logits_nchw, probs_nchw, argmax);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// ------------------------------------------------------------------------------------------- loss
// Two loss definitions share these kernels (LossSpec::mode):
//   0  MSAUWrapper.loss (model/model.py:446-459): CE(main) + CE(aux) over the pixels with label != 0, mean over the kept
//      pixels of a page, mean over pages;
//   1  UNetLoss (model/training/cost.py:35-65): w_main * CrossEntropyLoss(logits, tgt) + w_aux * CrossEntropyLoss(aux,
//      aux_tgt) over ALL pixels of the batch (label 0 is a class like any other), optional class weights with torch's
//      weighted-mean reduction  sum_p w[t_p] nll_p / sum_p w[t_p];  labels = argmax of the one-hot targets (:41,:52).
// stage 1: class histogram per (head, page) -- integer atomics, so the denominators are exact and deterministic;
// stage 2: per-pixel log-softmax NLL of both heads + dlogits + per-block partial sums; stage 3: fixed-order final sum.
// Out-of-range labels (negative or >= n_class; torch raises) contribute nothing and raise the plan's error flag;
// the masked accuracy of cost.py:44-50 / train...py:135-159 (argmax == label over label != 0) is counted on the way.
template <typename LT>
__global__ void __launch_bounds__(256) class_hist_kernel(const LT* __restrict__ labels, const LT* __restrict__ labels_aux, long npix_per_page,
                                                         int n_class, int* __restrict__ hist /* [2][B][32] */, int B, int* __restrict__ flags) {
  __shared__ int sh[64];
  const int b = blockIdx.y;
  if (threadIdx.x < 64) sh[threadIdx.x] = 0;
  __syncthreads();
  const LT* l0 = labels + (long)b * npix_per_page;
  const LT* l1 = labels_aux ? labels_aux + (long)b * npix_per_page : nullptr;
  bool bad = false;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npix_per_page; i += (long)gridDim.x * blockDim.x) {
    const long long a = (long long)l0[i];
    if (a < 0 || a >= n_class) bad = true; else atomicAdd(&sh[(int)a], 1);
    if (l1) {
      const long long c = (long long)l1[i];
      if (c < 0 || c >= n_class) bad = true; else atomicAdd(&sh[32 + (int)c], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64 && sh[threadIdx.x]) atomicAdd(hist + ((threadIdx.x >> 5) * B + b) * 32 + (threadIdx.x & 31), sh[threadIdx.x]);
  if (bad) atomicOr(flags, 1);
}

template <int P>
__device__ __forceinline__ float ce_head(const float* __restrict__ src, float* __restrict__ dst, long idx, int n_class, int lab, float scale,
                                         float gscale, bool active, int* correct) {
  float4* d4 = reinterpret_cast<float4*>(dst + idx * P);
  if (!active) {
#pragma unroll
    for (int i = 0; i < P / 4; ++i) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return 0.f;
  }
  float v[P];
  load_pixel<P>(src, idx, v);
  float m = v[0]; int am = 0;
#pragma unroll
  for (int k = 1; k < P; ++k) if (k < n_class && v[k] > m) { m = v[k]; am = k; }
  if (correct && lab != 0 && am == lab) ++*correct;
  float s = 0.f, vl = 0.f;
#pragma unroll
  for (int k = 0; k < P; ++k) { if (k == lab) vl = v[k]; v[k] = k < n_class ? expf(v[k] - m) : 0.f; s += v[k]; }
  const float gs = scale * gscale;
  const float inv = gs / s;
#pragma unroll
  for (int k = 0; k < P; ++k) v[k] = v[k] * inv - (k == lab ? gs : 0.f);
#pragma unroll
  for (int i = 0; i < P / 4; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  return (logf(s) + m - vl) * scale;
}

struct LossDev {
  int mode; float w_main, w_aux;
  float cw[32];      // class weights (all 1 when absent)
  int has_cw;
};

template <typename LT, int P>
__global__ void __launch_bounds__(256) ce_kernel(const float* __restrict__ lg, const float* __restrict__ la, int n_class,
                                                  const LT* __restrict__ labels, const LT* __restrict__ labels_aux, const int* __restrict__ hist,
                                                  long npix_per_page, int B, float gscale, const LossDev spec, float* __restrict__ dlg,
                                                  float* __restrict__ dla, float* __restrict__ partial, int* __restrict__ acc_counts) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  // denominators (identical in every thread: fixed-order sums over the integer histogram)
  __shared__ float den_s[2][64];   // mode 0: per page (B <= 64 cached, else recomputed); mode 1: [head][0]
  if (spec.mode == 1) {
    if (threadIdx.x < 2) {
      double d = 0.0;
      for (int b = 0; b < B; ++b)
        for (int c = 0; c < n_class; ++c) d += (double)spec.cw[c] * (double)hist[(threadIdx.x * B + b) * 32 + c];
      den_s[threadIdx.x][0] = (float)d;
    }
  } else if (threadIdx.x < 64 && threadIdx.x < B) {
    int kept = 0;
    for (int c = 1; c < n_class; ++c) kept += hist[threadIdx.x * 32 + c];
    den_s[0][threadIdx.x] = (float)kept;
  }
  __syncthreads();
  float loss = 0.f, loss_main = 0.f;
  int correct = 0, kept = 0;
  if (idx < npix_per_page * B) {
    const int b = (int)(idx / npix_per_page);
    const long long l0 = (long long)labels[idx];
    const long long l1 = labels_aux ? (long long)labels_aux[idx] : l0;
    const bool ok0 = l0 >= 0 && l0 < n_class, ok1 = l1 >= 0 && l1 < n_class;
    if (spec.mode == 0) {
      const bool act = ok0 && l0 != 0;
      float den;
      if (b < 64) den = den_s[0][b];
      else { int k = 0; for (int c = 1; c < n_class; ++c) k += hist[b * 32 + c]; den = (float)k; }
      const float scale = act ? 1.f / (den * (float)B) : 0.f;
      loss_main = ce_head<P>(lg, dlg, idx, n_class, (int)l0, scale, gscale, act, &correct);
      loss = loss_main + ce_head<P>(la, dla, idx, n_class, (int)l0, scale, gscale, act, nullptr);
      kept += act ? 1 : 0;
    } else {
      const float s0 = ok0 ? spec.w_main * spec.cw[ok0 ? (int)l0 : 0] / den_s[0][0] : 0.f;
      const float s1 = ok1 ? spec.w_aux * spec.cw[ok1 ? (int)l1 : 0] / den_s[1][0] : 0.f;
      loss_main = ce_head<P>(lg, dlg, idx, n_class, (int)l0, s0, gscale, ok0, &correct);
      loss = loss_main + ce_head<P>(la, dla, idx, n_class, (int)l1, s1, gscale, ok1 && spec.w_aux != 0.f, nullptr);
      kept += (ok0 && l0 != 0) ? 1 : 0;
    }
  }
  __shared__ float ws[8], wm[8];
  __shared__ int wc[8], wk[8];
  for (int o = 16; o; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    loss_main += __shfl_xor_sync(0xffffffffu, loss_main, o);
    correct += __shfl_xor_sync(0xffffffffu, correct, o);
    kept += __shfl_xor_sync(0xffffffffu, kept, o);
  }
  if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = loss; wm[threadIdx.x >> 5] = loss_main; wc[threadIdx.x >> 5] = correct; wk[threadIdx.x >> 5] = kept; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, sm = 0.f; int c = 0, k = 0;
    for (int i = 0; i < 8; ++i) { s += ws[i]; sm += wm[i]; c += wc[i]; k += wk[i]; }
    partial[blockIdx.x] = s;
    partial[gridDim.x + blockIdx.x] = sm;
    if (c) atomicAdd(acc_counts, c);
    if (k) atomicAdd(acc_counts + 1, k);
  }
}

// out[0] = sum of partial[0..n) (the loss), out[1] = main_scale * sum of partial[n..2n) (the main head's own term)
__global__ void __launch_bounds__(1024) final_sum_kernel(const float* __restrict__ partial, int n, float main_scale, float* __restrict__ out,
                                                         float* __restrict__ out_main) {
  __shared__ double sh[1024], shm[1024];
  double s = 0.0, sm = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) { s += (double)partial[i]; sm += (double)partial[n + i]; }
  sh[threadIdx.x] = s; shm[threadIdx.x] = sm;
  __syncthreads();
  for (int o = 512; o; o >>= 1) {
    if (threadIdx.x < o) { sh[threadIdx.x] += sh[threadIdx.x + o]; shm[threadIdx.x] += shm[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *out = (float)sh[0];
    if (out_main) *out_main = (float)(shm[0] * (double)main_scale);
  }
}

int loss_partial_count(int B, long npix_per_page) { return cdiv(npix_per_page * B, 256); }   // the partial buffer holds 2x this many floats
long loss_scratch_ints(int B) { return 2L * B * 32 + 8; }   // histogram [2][B][32] + {correct, kept}

template <typename LT>
static int loss_typed(const float* lg, const float* la, int P, int n_class, const LT* labels, const LT* labels_aux, int B, long npix_per_page,
                      float gscale, const LossDev& spec, float* dlg, float* dla, int* ints, int* flags, float* partial, float* loss_out, float* loss_main_out, cudaStream_t st) {
  const int nblk = loss_partial_count(B, npix_per_page);
  int* hist = ints;
  int* acc = ints + 2L * B * 32;
  dim3 cg(min(cdiv(npix_per_page, 256), 64), B);
  class_hist_kernel<LT><<<cg, 256, 0, st>>>(labels, spec.mode == 1 ? labels_aux : nullptr, npix_per_page, n_class, hist, B, flags);
#define MSAU_CE(PV) ce_kernel<LT, PV><<<nblk, 256, 0, st>>>(lg, la, n_class, labels, spec.mode == 1 ? labels_aux : nullptr, hist, npix_per_page, B, gscale, spec, dlg, dla, partial, acc)
  if (P == 8) MSAU_CE(8); else if (P == 16) MSAU_CE(16); else MSAU_CE(32);
#undef MSAU_CE
  final_sum_kernel<<<1, 1024, 0, st>>>(partial, nblk, (spec.mode == 1 && spec.w_main != 0.f) ? 1.f / spec.w_main : 1.f, loss_out, loss_main_out);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

int launch_loss(const float* lg, const float* la, int P, int n_class, const void* labels, const void* labels_aux, int label_dtype, int B,
                long npix_per_page, float gscale, const LossSpec& hs, float* dlg, float* dla, int* ints, int* flags, float* partial,
                float* loss_out, float* loss_main_out, cudaStream_t st) {
  MSAU_CHECK_ARG((P == 8 || P == 16 || P == 32) && n_class <= P && n_class <= 32, "loss: logits pitch 8/16/32 and n_class <= 32 supported");
  MSAU_CHECK_ARG(hs.mode == 0 || hs.mode == 1, "loss: mode must be 0 (MSAUWrapper.loss) or 1 (UNetLoss)");
  LossDev spec;
  spec.mode = hs.mode; spec.w_main = hs.w_main; spec.w_aux = hs.w_aux; spec.has_cw = hs.class_weights != nullptr;
  for (int c = 0; c < 32; ++c) spec.cw[c] = (hs.class_weights && c < n_class) ? hs.class_weights[c] : 1.f;
  // `flags` is sticky across calls: only msau_plan_error_flags clears it
  MSAU_CUDA_TRY(cudaMemsetAsync(ints, 0, sizeof(int) * (2L * B * 32 + 2), st));
  ProfScope ps("loss_kernels", 0, (double)npix_per_page * B * (4.0 * 4 * P + 2.0 * (label_dtype == 1 ? 8 : 1)), st);
  if (label_dtype == 1)
    return loss_typed<long long>(lg, la, P, n_class, (const long long*)labels, (const long long*)labels_aux, B, npix_per_page, gscale, spec, dlg,
                                 dla, ints, flags, partial, loss_out, loss_main_out, st);
  return loss_typed<uint8_t>(lg, la, P, n_class, (const uint8_t*)labels, (const uint8_t*)labels_aux, B, npix_per_page, gscale, spec, dlg, dla,
                             ints, flags, partial, loss_out, loss_main_out, st);
}

// ------------------------------------------------------------------------------------------- driver-side helpers
// tgt = torch.argmax(tgt, dim=1) on [B, C, H, W] one-hot targets (cost.py:41,52): first maximum wins
template <typename T>
__global__ void __launch_bounds__(256) onehot_argmax_kernel(const T* __restrict__ t, int C, long npix_per_page, long total, long stride_c,
                                                            long stride_p, uint8_t* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long b = idx / npix_per_page, p = idx - b * npix_per_page;
  const T* s = t + b * C * npix_per_page + p * stride_p;
  T m = s[0]; int am = 0;
  for (int c = 1; c < C; ++c) { const T v = s[(long)c * stride_c]; if (v > m) { m = v; am = c; } }
  out[idx] = (uint8_t)am;
}

int launch_onehot_argmax(const void* t, int dtype, int B, int C, long npix_per_page, int channels_last, uint8_t* out, cudaStream_t st) {
  MSAU_CHECK_ARG(C >= 1 && C <= 255, "onehot_argmax: 1..255 classes");
  const long total = (long)B * npix_per_page;
  const int grid = cdiv(total, 256);
  const long sc = channels_last ? 1 : npix_per_page, sp = channels_last ? C : 1;
  ProfScope ps("onehot_argmax_kernel", 0, (double)total * (C * (dtype == 1 ? 8.0 : dtype == 2 ? 4.0 : 1.0) + 1.0), st);
  if (dtype == 1) onehot_argmax_kernel<long long><<<grid, 256, 0, st>>>((const long long*)t, C, npix_per_page, total, sc, sp, out);
  else if (dtype == 2) onehot_argmax_kernel<float><<<grid, 256, 0, st>>>((const float*)t, C, npix_per_page, total, sc, sp, out);
  else if (dtype == 0) onehot_argmax_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)t, C, npix_per_page, total, sc, sp, out);
  else { set_error("onehot_argmax: dtype must be 0 (uint8), 1 (int64) or 2 (float32)"); return MSAU_ERR_ARG; }
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// masked accuracy + confusion counts of evaluate() (train_chargrid_funsd_msau.py:133-159): pixels with label != 0 only.
// confusion int64 [n_class][n_class] (row = label, column = prediction), accumulated (the caller zeroes it per evaluation)
template <typename LT>
__global__ void __launch_bounds__(256) confusion_kernel(const uint8_t* __restrict__ pred, const LT* __restrict__ labels, long n, int n_class,
                                                        unsigned long long* __restrict__ conf) {
  extern __shared__ unsigned int shc[];
  const int cells = n_class * n_class;
  for (int i = threadIdx.x; i < cells; i += blockDim.x) shc[i] = 0;
  __syncthreads();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long long l = (long long)labels[i];
    const int pr = pred[i];
    if (l > 0 && l < n_class && pr < n_class) atomicAdd(&shc[(int)l * n_class + pr], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += blockDim.x)
    if (shc[i]) atomicAdd(conf + i, (unsigned long long)shc[i]);
}

int launch_confusion(const uint8_t* pred, const void* labels, int label_dtype, long n, int n_class, long long* conf, cudaStream_t st) {
  MSAU_CHECK_ARG(n_class >= 1 && n_class <= 32, "confusion: n_class <= 32");
  int grid = cdiv(n, 256 * 8);
  if (grid > 8 * sm_count()) grid = 8 * sm_count();
  if (grid < 1) grid = 1;
  const size_t sh = sizeof(unsigned int) * n_class * n_class;
  ProfScope ps("confusion_kernel", 0, (double)n * (1.0 + (label_dtype == 1 ? 8 : 1)), st);
  if (label_dtype == 1) confusion_kernel<long long><<<grid, 256, sh, st>>>(pred, (const long long*)labels, n, n_class, (unsigned long long*)conf);
  else confusion_kernel<uint8_t><<<grid, 256, sh, st>>>(pred, (const uint8_t*)labels, n, n_class, (unsigned long long*)conf);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

}  // namespace msau
