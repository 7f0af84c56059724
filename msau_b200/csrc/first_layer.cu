// Structured-input first layer: a chargrid is a one-hot tensor, so the first 3x3 conv
//     z[p][co] = bias[co] + sum_{tap, ci} x[p + off(tap)][ci] * W[co][ci][tap]            (model/model.py:129-136)
// only ever adds ONE weight column per tap, W[:, id(p + off(tap)), tap], and its weight gradient is a histogram of the
// output gradients keyed by (id, tap).  The dense [B, 96, H, W] tensor is the network's largest by far (1.6 GB at
// B = 16): the generic kernels spend 2 ms per step streaming it twice.  Here it is read once:
//   onehot_scan_kernel   dense input -> int16 id map (-1 = all-zero pixel) + a device flag that stays 0 only if EVERY
//                        pixel is all-zero or exactly one 1.0 (checked value by value; no host round trip)
//   first_fwd_kernel     id-gather convolution (9 taps x 8 output channels from a 27 KB shared-memory weight table)
//   first_wgrad_kernel   scatter of dz into a per-CTA shared-memory histogram [ci][tap][co] + bias column sums
// Both structured kernels return immediately when the flag is set, and the dense kernels (conv_tc / wgrad_tc, launched
// right after with ConvArgs::skip_flag) return immediately when it is not: any dense input, e.g. a BERT grid, keeps
// working, bit-for-bit as before.  Results differ from the dense path only by fp32 summation order (no bf16 split at all).
#include "common.cuh"
#include "first_layer.cuh"
#include "prof.cuh"

namespace msau {

// ---------------------------------------------------------------------------------------------------- scan
// NCHW: a thread owns 4 consecutive pixels and walks the channel planes with 16-byte loads (coalesced along x).
__global__ void __launch_bounds__(256) onehot_scan_nchw_kernel(const float* __restrict__ x, int C, long plane, long n_quads_per_page,
                                                                int B, short* __restrict__ ids, int* __restrict__ flag) {
  const long q = (long)blockIdx.x * 256 + threadIdx.x;
  if (q >= n_quads_per_page * B) return;
  const long b = q / n_quads_per_page, r = q - b * n_quads_per_page;
  const float4* src = reinterpret_cast<const float4*>(x + b * C * plane) + r;
  const long stride4 = plane >> 2;
  int id[4] = {-1, -1, -1, -1};
  int bad = 0;
  for (int c = 0; c < C; ++c) {
    const float4 v = __ldg(src + c * stride4);
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (f[k] != 0.f) {
        if (f[k] != 1.f || id[k] >= 0) bad = 1;
        id[k] = c;
      }
    }
  }
  short* dst = ids + b * plane + r * 4;
  *reinterpret_cast<short4*>(dst) = make_short4((short)id[0], (short)id[1], (short)id[2], (short)id[3]);
  if (bad) atomicOr(flag, 1);
}

// NHWC: 8 threads per pixel, each 16-byte load covers 4 channels (coalesced along the channel axis)
__global__ void __launch_bounds__(256) onehot_scan_nhwc_kernel(const float* __restrict__ x, int C, int pitch, long npix,
                                                                short* __restrict__ ids, int* __restrict__ flag) {
  const long t = (long)blockIdx.x * 256 + threadIdx.x;
  const long p = t >> 3;
  const int sub = (int)(t & 7);
  int id = -1, bad = 0;
  if (p < npix) {
    const float4* src = reinterpret_cast<const float4*>(x + p * pitch);
    for (int c4 = sub; c4 * 4 < C; c4 += 8) {
      const float4 v = __ldg(src + c4);
      const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c4 * 4 + k < C && f[k] != 0.f) {
          if (f[k] != 1.f || id >= 0) bad = 1;
          id = c4 * 4 + k;
        }
      }
    }
  }
  // combine the 8 partial results of a pixel
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const int oid = __shfl_xor_sync(0xffffffffu, id, o);
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    if (oid >= 0) {
      if (id >= 0) bad = 1;
      id = oid;
    }
  }
  if (p < npix && sub == 0) ids[p] = (short)id;
  if (bad) atomicOr(flag, 1);
}

int launch_onehot_scan(const float* x, int layout_nchw, int C, int pitch, int B, int H, int W, short* ids, int* flag, cudaStream_t st) {
  MSAU_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
  const long plane = (long)H * W;
  const double bytes = (double)B * plane * C * 4.0 + (double)B * plane * 2.0;
  ProfScope ps("onehot_scan_kernel", 0, bytes, st);
  if (layout_nchw) {
    MSAU_CHECK_ARG(plane % 4 == 0, "onehot_scan: H*W must be a multiple of 4");
    const long quads = plane / 4 * B;
    onehot_scan_nchw_kernel<<<(unsigned)cdiv(quads, 256), 256, 0, st>>>(x, C, plane, plane / 4, B, ids, flag);
  } else {
    const long npix = plane * B;
    onehot_scan_nhwc_kernel<<<(unsigned)cdiv(npix * 8, 256), 256, 0, st>>>(x, C, pitch, npix, ids, flag);
  }
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// ---------------------------------------------------------------------------------------------------- forward
// w: packed fp32 [tap][c_in_p][8] (the layout conv.cu consumes), bias [8].  One thread = one output pixel.
__global__ void __launch_bounds__(256) first_fwd_kernel(const short* __restrict__ ids, const int* __restrict__ flag,
                                                         const float* __restrict__ w, const float* __restrict__ bias, int cin, int cinp,
                                                         int B, int H, int W, float* __restrict__ out, int po) {
  if (*flag != 0) return;
  extern __shared__ __align__(16) float wt[];                 // [tap][cin][8]
  for (int e = threadIdx.x; e < 9 * cin * 2; e += 256) {
    const int tap = e / (cin * 2), r = e - tap * cin * 2;
    reinterpret_cast<float4*>(wt)[e] = __ldg(reinterpret_cast<const float4*>(w + ((long)tap * cinp) * 8) + r);
  }
  __syncthreads();
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias)), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 1);
  const long npix = (long)B * H * W;
  for (long p = (long)blockIdx.x * 256 + threadIdx.x; p < npix; p += (long)gridDim.x * 256) {
    const int x = (int)(p % W);
    const long r = p / W;
    const int y = (int)(r % H);
    float acc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if ((unsigned)yy >= (unsigned)H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if ((unsigned)xx >= (unsigned)W) continue;
        const int id = ids[p + (long)(ky - 1) * W + (kx - 1)];
        if (id >= 0) {
          const float4* wp = reinterpret_cast<const float4*>(wt + ((ky * 3 + kx) * cin + id) * 8);
          const float4 w0 = wp[0], w1 = wp[1];
          acc[0] += w0.x; acc[1] += w0.y; acc[2] += w0.z; acc[3] += w0.w;
          acc[4] += w1.x; acc[5] += w1.y; acc[6] += w1.z; acc[7] += w1.w;
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(out + p * po);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

int launch_first_fwd(const short* ids, const int* flag, const float* w, const float* bias, int cin, int cinp, int B, int H, int W,
                     float* out, int po, cudaStream_t st) {
  const size_t smem = (size_t)9 * cin * 8 * sizeof(float);
  MSAU_CHECK_ARG(smem <= 200 * 1024, "first_fwd: weight table does not fit in shared memory (%d channels)", cin);
  static bool attr = false;
  if (!attr) {
    MSAU_CUDA_TRY(cudaFuncSetAttribute(first_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const long npix = (long)B * H * W;
  int grid = sm_count() * 4;
  if ((long)grid * 256 > npix) grid = cdiv(npix, 256);
  ProfScope ps("first_fwd_kernel", 2.0 * npix * 9 * 8, (double)npix * (2.0 + 32.0), st);
  first_fwd_kernel<<<grid, 256, smem, st>>>(ids, flag, w, bias, cin, cinp, B, H, W, out, po);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// ---------------------------------------------------------------------------------------------------- weight gradient
// dW[co][ci][ky][kx] = sum over pixels q with id(q) = ci of dz[q - (ky-1, kx-1)][co];   db[co] = sum_p dz[p][co]
__global__ void __launch_bounds__(256) first_wgrad_kernel(const short* __restrict__ ids, const int* __restrict__ flag,
                                                           const float* __restrict__ dz, int pdz, int cin, int cout, int B, int H, int W,
                                                           float* __restrict__ dW, float* __restrict__ dbias) {
  if (*flag != 0) return;
  extern __shared__ __align__(16) float hist[];               // [ci][tap][8]
  __shared__ float bsum[8];
  for (int e = threadIdx.x; e < cin * 72; e += 256) hist[e] = 0.f;
  if (threadIdx.x < 8) bsum[threadIdx.x] = 0.f;
  __syncthreads();
  const long npix = (long)B * H * W;
  float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long q = (long)blockIdx.x * 256 + threadIdx.x; q < npix; q += (long)gridDim.x * 256) {
    const float4* gp = reinterpret_cast<const float4*>(dz + q * pdz);
    const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
    bacc[0] += g0.x; bacc[1] += g0.y; bacc[2] += g0.z; bacc[3] += g0.w;
    bacc[4] += g1.x; bacc[5] += g1.y; bacc[6] += g1.z; bacc[7] += g1.w;
    const int id = ids[q];
    if (id < 0) continue;
    const int x = (int)(q % W);
    const int y = (int)((q / W) % H);
    float* h = hist + id * 72;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int py = y - (ky - 1);                            // output pixel that saw q through tap (ky, kx)
      if ((unsigned)py >= (unsigned)H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int px = x - (kx - 1);
        if ((unsigned)px >= (unsigned)W) continue;
        const float4* sp = reinterpret_cast<const float4*>(dz + (q - (long)(ky - 1) * W - (kx - 1)) * pdz);
        const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
        float* t = h + (ky * 3 + kx) * 8;
        atomicAdd(t + 0, s0.x); atomicAdd(t + 1, s0.y); atomicAdd(t + 2, s0.z); atomicAdd(t + 3, s0.w);
        atomicAdd(t + 4, s1.x); atomicAdd(t + 5, s1.y); atomicAdd(t + 6, s1.z); atomicAdd(t + 7, s1.w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float s = bacc[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&bsum[k], s);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cin * 72; e += 256) {
    const float v = hist[e];
    if (v != 0.f) {
      const int ci = e / 72, r = e - ci * 72, tap = r >> 3, co = r & 7;
      if (co < cout) atomicAdd(dW + ((long)co * cin + ci) * 9 + tap, v);
    }
  }
  if (threadIdx.x < cout) atomicAdd(dbias + threadIdx.x, bsum[threadIdx.x]);
}

int launch_first_wgrad(const short* ids, const int* flag, const float* dz, int pdz, int cin, int cout, int B, int H, int W, float* dW,
                       float* dbias, cudaStream_t st) {
  MSAU_CHECK_ARG(cout <= 8, "first_wgrad: at most 8 output channels");
  const size_t smem = (size_t)cin * 72 * sizeof(float);
  MSAU_CHECK_ARG(smem <= 200 * 1024, "first_wgrad: histogram does not fit in shared memory (%d channels)", cin);
  static bool attr = false;
  if (!attr) {
    MSAU_CUDA_TRY(cudaFuncSetAttribute(first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const long npix = (long)B * H * W;
  int grid = sm_count() * 4;
  if ((long)grid * 256 > npix) grid = cdiv(npix, 256);
  ProfScope ps("first_wgrad_kernel", 2.0 * npix * 9 * 8, (double)npix * (2.0 + 32.0), st);
  first_wgrad_kernel<<<grid, 256, smem, st>>>(ids, flag, dz, pdz, cin, cout, B, H, W, dW, dbias);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}


// ---------------------------------------------------------------------------------------------------- BERT grid
// Box-constant input (data_generator_funsd_bert.py:64-93: every pixel of a box carries that box's feature vector):
// with ids[p] = feature-table row of the owning box,
//     z[p][co] = bias[co] + sum_tap P[ids[p + off(tap)]][tap][co],     P[r][tap][co] = sum_ci table[r][ci] * W[co][ci][tap]
// i.e. the 768 -> 8 conv over a 768-channel dense grid becomes a rows x 768 x 72 projection (tiny) + the id-gather above;
// the weight gradient is the same histogram keyed by row, followed by dW[co][ci][tap] = sum_r table[r][ci] * hist[r][tap][co].

// P[r][tap][co]; w: packed fp32 [tap][cinp][8]; one block per table row, thread = (tap, co)
__global__ void __launch_bounds__(96) table_project_kernel(const float* __restrict__ table, int cin, int cinp, const float* __restrict__ w,
                                                            float* __restrict__ P) {
  const int r = blockIdx.x, t = threadIdx.x;
  if (t >= 72) return;
  const int tap = t >> 3, co = t & 7;
  const float* row = table + (long)r * cin;
  const float* wp = w + (long)tap * cinp * 8 + co;
  float acc = 0.f;
  for (int ci = 0; ci < cin; ++ci) acc = fmaf(__ldg(row + ci), __ldg(wp + ci * 8), acc);
  P[(long)r * 72 + t] = acc;
}

__global__ void __launch_bounds__(256) first_fwd_table_kernel(const short* __restrict__ ids, const float* __restrict__ P,
                                                               const float* __restrict__ bias, int B, int H, int W, float* __restrict__ out,
                                                               int po) {
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias)), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 1);
  const long npix = (long)B * H * W;
  for (long p = (long)blockIdx.x * 256 + threadIdx.x; p < npix; p += (long)gridDim.x * 256) {
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    float acc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      if ((unsigned)(y + ky - 1) >= (unsigned)H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        if ((unsigned)(x + kx - 1) >= (unsigned)W) continue;
        const int id = ids[p + (long)(ky - 1) * W + (kx - 1)];
        if (id >= 0) {
          const float4* wp = reinterpret_cast<const float4*>(P + (long)id * 72 + (ky * 3 + kx) * 8);
          const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
          acc[0] += w0.x; acc[1] += w0.y; acc[2] += w0.z; acc[3] += w0.w;
          acc[4] += w1.x; acc[5] += w1.y; acc[6] += w1.z; acc[7] += w1.w;
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(out + p * po);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// hist[r][tap][co] += dz[q - off(tap)][co] for the pixels q with ids[q] = r.  A thread walks 16 consecutive pixels of an image
// row and keeps the 72 partial sums of the current run of equal ids in registers (boxes are wide: few flushes).
__global__ void __launch_bounds__(128) first_wgrad_table_kernel(const short* __restrict__ ids, const float* __restrict__ dz, int pdz, int B,
                                                                 int H, int W, float* __restrict__ hist) {
  const int segs = (W + 15) >> 4;
  const long n_seg = (long)B * H * segs;
  for (long sg = (long)blockIdx.x * 128 + threadIdx.x; sg < n_seg; sg += (long)gridDim.x * 128) {
    const int sx = (int)(sg % segs);
    const long row = sg / segs;
    const int y = (int)(row % H);
    const long rowpix = row * W;
    float acc[72];
#pragma unroll
    for (int i = 0; i < 72; ++i) acc[i] = 0.f;
    int cur = -1;
    const int x1 = min(W, sx * 16 + 16);
    for (int x = sx * 16; x <= x1; ++x) {
      const int id = x < x1 ? (int)ids[rowpix + x] : -2;
      if (id != cur) {
        if (cur >= 0) {
          float* h = hist + (long)cur * 72;
#pragma unroll
          for (int i = 0; i < 72; ++i) {
            if (acc[i] != 0.f) atomicAdd(h + i, acc[i]);
            acc[i] = 0.f;
          }
        }
        cur = id;
      }
      if (id < 0) continue;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int py = y - (ky - 1);
        if ((unsigned)py >= (unsigned)H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int px = x - (kx - 1);
          if ((unsigned)px >= (unsigned)W) continue;
          const float4* sp = reinterpret_cast<const float4*>(dz + (rowpix + x - (long)(ky - 1) * W - (kx - 1)) * pdz);
          const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
          float* t = acc + (ky * 3 + kx) * 8;
          t[0] += s0.x; t[1] += s0.y; t[2] += s0.z; t[3] += s0.w; t[4] += s1.x; t[5] += s1.y; t[6] += s1.z; t[7] += s1.w;
        }
      }
    }
  }
}

// dW[co][ci][tap] = sum_r table[r][ci] * hist[r][tap][co]; one block per input channel, thread = (tap, co)
__global__ void __launch_bounds__(96) table_wgrad_kernel(const float* __restrict__ table, int rows, int cin, const float* __restrict__ hist,
                                                          int cout, float* __restrict__ dW) {
  const int ci = blockIdx.x, t = threadIdx.x;
  if (t >= 72) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc = fmaf(__ldg(table + (long)r * cin + ci), __ldg(hist + (long)r * 72 + t), acc);
  const int tap = t >> 3, co = t & 7;
  if (co < cout) dW[((long)co * cin + ci) * 9 + tap] += acc;
}

int launch_table_first_fwd(const short* ids, const float* table, int rows, int cin, int cinp, const float* w, const float* bias, float* P,
                           int B, int H, int W, float* out, int po, cudaStream_t st) {
  const long npix = (long)B * H * W;
  {
    ProfScope ps("table_project_kernel", 2.0 * rows * cin * 72, (double)rows * (cin + 72) * 4.0, st);
    table_project_kernel<<<rows, 96, 0, st>>>(table, cin, cinp, w, P);
  }
  int grid = sm_count() * 8;
  if ((long)grid * 256 > npix) grid = cdiv(npix, 256);
  ProfScope ps("first_fwd_kernel", 2.0 * npix * 9 * 8, (double)npix * (2.0 + 32.0), st);
  first_fwd_table_kernel<<<grid, 256, 0, st>>>(ids, P, bias, B, H, W, out, po);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

int launch_table_first_wgrad(const short* ids, const float* table, int rows, int cin, int cout, const float* dz, int pdz, float* hist, int B,
                             int H, int W, float* dW, cudaStream_t st) {
  MSAU_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(float) * 72 * rows, st));
  const long n_seg = (long)B * H * ((W + 15) / 16);
  int grid = sm_count() * 8;
  if ((long)grid * 128 > n_seg) grid = cdiv(n_seg, 128);
  {
    ProfScope ps("first_wgrad_kernel", 2.0 * B * H * W * 9 * 8, (double)B * H * W * (2.0 + 32.0), st);
    first_wgrad_table_kernel<<<grid, 128, 0, st>>>(ids, dz, pdz, B, H, W, hist);
  }
  ProfScope ps("table_wgrad_kernel", 2.0 * rows * cin * 72, (double)rows * (cin + 72) * 4.0, st);
  table_wgrad_kernel<<<cin, 96, 0, st>>>(table, rows, cin, hist, cout, dW);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

}  // namespace msau
