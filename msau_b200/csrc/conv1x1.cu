// 1x1 convolutions (coupling convs on a concat, attention f|g / h projections, and their data gradients) as a plain
// fp32 streaming kernel.  A 1x1 conv moves (C_in + C_out) * 4 bytes per pixel for C_in * C_out FMAs: at MSAU's widths
// (8..128 -> 8..64) that is 4..43 FMA per byte, so the fp32 pipes keep up with HBM and the tensor-core machinery of
// conv_tc.cu (bf16 split, operand images, TMEM round trip) only adds latency.  One thread = one pixel: it streams its
// input channels with 16-byte loads, keeps the C_out accumulators in registers and reads the weight rows from shared memory
// (warp-wide broadcast).  Exact fp32 arithmetic.
// Reference semantics: model/model.py:143-148, 246-252 (coupling conv + ReLU), model/layers/attention.py:8-24.
#include "common.cuh"
#include "conv1x1.cuh"
#include "prof.cuh"

namespace msau {

// w: packed fp32 [c1 + c2][coutp] (rows = input channels of [src1 | src2]); epilogue: + bias, ReLU, * (omask > 0), += previous
// PX pixels per thread and iteration (256 pixels apart, so every load / store instruction of a warp is still one contiguous
// run): PX x more independent 16-byte loads in flight before the FMA chain starts.
template <int CO, int PX>
__global__ void __launch_bounds__(256, CO == 8 ? 3 : (CO == 16 ? 2 : 1)) conv1x1_kernel(const ConvArgs a, long npix) {
  extern __shared__ __align__(16) float wsm[];                 // [cin][CO]
  if (a.skip_flag) { pdl_wait(); if (*a.skip_flag == 0) return; }
  const int cin = a.c1 + a.c2;
  for (int e = threadIdx.x; e < cin * CO / 4; e += 256) reinterpret_cast<float4*>(wsm)[e] = __ldg(reinterpret_cast<const float4*>(a.w) + e);
  float* bsm = wsm + cin * CO;                                 // [CO] bias row after the weights
  if (threadIdx.x < CO) bsm[threadIdx.x] = a.bias ? __ldg(a.bias + threadIdx.x) : 0.f;
  __syncthreads();
  pdl_wait();        // PDL protocol (common.cuh): only packed weights were read so far
  pdl_trigger();
  for (long p0 = (long)blockIdx.x * (256 * PX) + threadIdx.x; p0 < npix; p0 += (long)gridDim.x * (256 * PX)) {
    float acc[PX][CO];
    long pp[PX];
#pragma unroll
    for (int u = 0; u < PX; ++u) {
      pp[u] = p0 + u * 256 < npix ? p0 + u * 256 : p0;         // out-of-range slots recompute pixel p0 and are not stored
#pragma unroll
      for (int c4 = 0; c4 < CO / 4; ++c4) {
        const float4 b = *reinterpret_cast<const float4*>(bsm + c4 * 4);
        acc[u][c4 * 4] = b.x; acc[u][c4 * 4 + 1] = b.y; acc[u][c4 * 4 + 2] = b.z; acc[u][c4 * 4 + 3] = b.w;
      }
    }
    const float* wr = wsm;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const float* src = s == 0 ? a.src1 : a.src2;
      const int pitch = s == 0 ? a.p1 : a.p2;
      const int nq = (s == 0 ? a.c1 : a.c2) >> 2;
#pragma unroll 2
      for (int q = 0; q < nq; ++q, wr += 4 * CO) {
        float4 x[PX];
#pragma unroll
        for (int u = 0; u < PX; ++u) x[u] = __ldg(reinterpret_cast<const float4*>(src + pp[u] * pitch) + q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int c4 = 0; c4 < CO / 4; ++c4) {
            const float4 w = *reinterpret_cast<const float4*>(wr + k * CO + c4 * 4);
#pragma unroll
            for (int u = 0; u < PX; ++u) {
              const float xv = k == 0 ? x[u].x : k == 1 ? x[u].y : k == 2 ? x[u].z : x[u].w;
              acc[u][c4 * 4 + 0] = fmaf(xv, w.x, acc[u][c4 * 4 + 0]);
              acc[u][c4 * 4 + 1] = fmaf(xv, w.y, acc[u][c4 * 4 + 1]);
              acc[u][c4 * 4 + 2] = fmaf(xv, w.z, acc[u][c4 * 4 + 2]);
              acc[u][c4 * 4 + 3] = fmaf(xv, w.w, acc[u][c4 * 4 + 3]);
            }
          }
        }
      }
    }
    float4 prev[PX][CO / 4];
    if (a.accumulate) {
#pragma unroll
      for (int u = 0; u < PX; ++u)
#pragma unroll
        for (int c4 = 0; c4 < CO / 4; ++c4) prev[u][c4] = reinterpret_cast<const float4*>(a.out + pp[u] * a.po)[c4];
    }
#pragma unroll
    for (int u = 0; u < PX; ++u) {
      if (p0 + u * 256 >= npix) continue;
      float4* dst = reinterpret_cast<float4*>(a.out + pp[u] * a.po);
#pragma unroll
      for (int c4 = 0; c4 < CO / 4; ++c4) {
        float4 r = make_float4(acc[u][c4 * 4], acc[u][c4 * 4 + 1], acc[u][c4 * 4 + 2], acc[u][c4 * 4 + 3]);
        if (a.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
        if (a.omask) {        // gradient through the ReLU that produced this tensor (same epilogue order as conv_kernel)
          const float4 m = __ldg(reinterpret_cast<const float4*>(a.omask + pp[u] * a.pom) + c4);
          r.x = m.x > 0.f ? r.x : 0.f; r.y = m.y > 0.f ? r.y : 0.f; r.z = m.z > 0.f ? r.z : 0.f; r.w = m.w > 0.f ? r.w : 0.f;
        }
        if (a.accumulate) { const float4 o = prev[u][c4]; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
        dst[c4] = r;
      }
    }
  }
}

bool conv1x1_supported(const ConvArgs& a) {
  if (a.kh != 1 || a.kw != 1 || a.stride != 1 || a.pad_t || a.pad_l || a.osy != 1 || a.oy0 || a.ox0) return false;
  if (a.Hq != a.Hin || a.Wq != a.Win || a.Hout != a.Hin || a.Wout != a.Win) return false;
  if (a.src1_nchw || a.mask1 || a.relu1 || a.res || a.relu2 || a.add || a.addmask || a.s2d || a.d2s) return false;
  if ((a.c1 & 3) || (a.c2 & 3) || (a.p1 & 3) || (a.c2 && (a.p2 & 3)) || (a.po & 3) || (a.omask && (a.pom & 3))) return false;
  if (!(a.coutp == 8 || a.coutp == 16 || a.coutp == 32 || a.coutp == 64)) return false;
  // measured: wins where the FMA work per pixel is small (the 8/16-channel levels); the wider 1x1 convs (maps <= 128^2, few
  // pixels per SM) stay on the tensor-core kernel
  return (a.c1 + a.c2) * a.coutp <= 512;
}

int launch_conv1x1(const ConvArgs& a, cudaStream_t st) {
  MSAU_CHECK_ARG(conv1x1_supported(a), "conv1x1: unsupported shape");
  const long npix = (long)a.B * a.Hin * a.Win;
  const size_t smem = (size_t)(a.c1 + a.c2 + 1) * a.coutp * 4;
  const int px = a.coutp <= 16 ? 2 : 1;
  long blocks = (npix + 256 * px - 1) / (256 * px);
  const long cap = (long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  double bytes = (double)npix * (a.c1 + a.c2 + a.coutp * (1 + (a.accumulate ? 1 : 0) + (a.omask ? 1 : 0))) * 4.0;
  ProfScope ps("conv1x1_kernel", a.c1 + a.c2, a.coutp, 1, 1, a.Wout, a.accumulate, 2.0 * npix * (a.c1 + a.c2) * a.coutp, bytes, st);
#define MSAU_PW(CO, PX)                                                                                         \
  {                                                                                                             \
    static bool attr = false;                                                                                   \
    if (!attr) { MSAU_CUDA_TRY(cudaFuncSetAttribute(conv1x1_kernel<CO, PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; } \
    MSAU_CUDA_TRY(launch_pdl(conv1x1_kernel<CO, PX>, dim3((unsigned)blocks), dim3(256), smem, st, a, npix));       \
  }
  switch (a.coutp) {
    case 8: MSAU_PW(8, 2) break;
    case 16: MSAU_PW(16, 2) break;
    case 32: MSAU_PW(32, 1) break;
    default: MSAU_PW(64, 1) break;
  }
#undef MSAU_PW
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

}  // namespace msau
