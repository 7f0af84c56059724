// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// GEMM view of a same-size, stride-1 convolution (the forward of every 3x3 / dilated 3x3 / 1x1 / 4x4 conv of
// MSAU and, with flipped packed weights, every data-gradient):
//     D[pixel, cout] = sum_{tap, cin} A[pixel + offset(tap), cin] * W[tap, cin, cout]
//   M = 128 output pixels  = one 16-row x 8-column tile (row-group r of the MMA = image row r of the tile)
//   N = cout (padded to a multiple of 16)
//   K = 16 per instruction (see the split below)
// A is never materialised as an im2col matrix: the fp32 NHWC halo tile is converted once into a planar bf16
// image in shared memory, [8-channel plane][halo row][halo col][8 ch] (16 B per pixel per plane), which IS the
// canonical K-major no-swizzle UMMA layout (core matrix = 8 consecutive pixels x 8 channels = 128 contiguous
// bytes; SBO = halo row pitch).  A tap is just a different start address of the same image.
//
// fp32 accuracy from bf16 tensor cores (the "3-MMA split"): x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi); x*w ~= hi*whi + lo*whi + hi*wlo, fp32 accumulate.  With K = 16 = two
// 8-channel chunks per instruction this costs 1.5 instructions per (tap, 8 channels):
//     type 1:  A = [hi(tap) | lo(tap)]      (LBO = hi-plane -> lo-plane distance)   B = [Whi(tap) ; Whi(tap)]
//     type 3:  A = [hi(tap) | hi(tap')]     (LBO = offset(tap') - offset(tap))       B = [Wlo(tap) ; Wlo(tap')]
//
// One CTA = T tiles side by side (16 x 8T pixels), one TMEM accumulator of N columns per tile; input planes are
// streamed through a 2-stage shared-memory ring (mbarrier + tcgen05.commit), the epilogue reads TMEM with
// tcgen05.ld (one thread = one pixel, all couts) and applies bias / ReLU / residual / masks like conv.cu.
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "prof.cuh"

namespace msau {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// bounded wait: a protocol bug must surface as a trap, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 in, fp32 accumulate)
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split_bf16x8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct TcTile {
  int T, N, HH, HW, P, stages;      // tiles per CTA, MMA N, halo extent, input planes, ring depth
  int n1, n3;                       // type-1 / type-3 instructions per plane (= taps, ceil(taps/2))
  uint32_t in_bytes, w_bytes, stage_bytes, tmem_cols;
};

static constexpr int TC_THREADS = 256;
static constexpr int LD_U = 2;        // halo pixels whose global loads are in flight per thread

__global__ void __launch_bounds__(TC_THREADS, 3) conv_tc_kernel(const ConvArgs a, const uint16_t* __restrict__ wtc, const TcTile t) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_free[2];     // stage s may be overwritten (its MMAs retired)
  __shared__ uint64_t bar_done;        // all MMAs retired -> epilogue
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t tap_off16[16];   // smem offset of every tap, in 16-B units

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * (8 * t.T), y0 = blockIdx.y * 16;
  const int in_x0 = x0 - a.pad_l, in_y0 = y0 - a.pad_t;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(t.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 16 && tid < a.kh * a.kw) {
    const int ky = tid / a.kw, kx = tid - ky * a.kw;
    tap_off16[tid] = (uint32_t)(ky * a.dil * t.HW + kx * a.dil);
  }
  if (tid == 0) {
    mbar_init(&bar_free[0], 1);
    mbar_init(&bar_free[1], 1);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(t.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t plane_bytes = (uint32_t)t.HH * t.HW * 16;       // one of {hi, lo}
  const uint32_t row_pitch = (uint32_t)t.HW * 16;
  const int halo_px = t.HH * t.HW;
  const int planes1 = a.c1 >> 3;

  for (int p = 0; p < t.P; ++p) {
    const int s = p % t.stages;
    uint8_t* st = smem + (size_t)s * t.stage_bytes;
    if (p >= t.stages) mbar_wait(&bar_free[s], ((p / t.stages) - 1) & 1);
    // ---- load + split this 8-channel plane of the halo tile ----
    const bool from1 = p < planes1;
    const int cg = from1 ? (p << 3) : ((p - planes1) << 3);
    for (int e0 = tid; e0 < halo_px; e0 += TC_THREADS * LD_U) {
      float v[LD_U][8];
      float4 mk[LD_U][2];
      bool inb[LD_U];
#pragma unroll
      for (int u = 0; u < LD_U; ++u) {
        const int e = e0 + u * TC_THREADS;
        const int iy = e / t.HW, ix = e - iy * t.HW;
        const int gy = in_y0 + iy, gx = in_x0 + ix;
        inb[u] = e < halo_px && gy >= 0 && gy < a.Hin && gx >= 0 && gx < a.Win;
#pragma unroll
        for (int k = 0; k < 8; ++k) v[u][k] = 0.f;
        mk[u][0] = mk[u][1] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (inb[u]) {
          const long pix = ((long)b * a.Hin + gy) * a.Win + gx;
          if (from1) {
            if (a.src1_nchw) {
              const long plane = (long)a.Hin * a.Win;
              const float* sp = a.src1 + ((long)b * a.c1_logical + cg) * plane + (long)gy * a.Win + gx;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (cg + k < a.c1_logical) v[u][k] = __ldg(sp + k * plane);
            } else {
              const float4 q0 = __ldg(reinterpret_cast<const float4*>(a.src1 + pix * a.p1 + cg));
              const float4 q1 = __ldg(reinterpret_cast<const float4*>(a.src1 + pix * a.p1 + cg) + 1);
              v[u][0] = q0.x; v[u][1] = q0.y; v[u][2] = q0.z; v[u][3] = q0.w;
              v[u][4] = q1.x; v[u][5] = q1.y; v[u][6] = q1.z; v[u][7] = q1.w;
            }
            if (a.mask1) {
              mk[u][0] = __ldg(reinterpret_cast<const float4*>(a.mask1 + pix * a.pm1 + cg));
              mk[u][1] = __ldg(reinterpret_cast<const float4*>(a.mask1 + pix * a.pm1 + cg) + 1);
            }
          } else {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(a.src2 + pix * a.p2 + cg));
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(a.src2 + pix * a.p2 + cg) + 1);
            v[u][0] = q0.x; v[u][1] = q0.y; v[u][2] = q0.z; v[u][3] = q0.w;
            v[u][4] = q1.x; v[u][5] = q1.y; v[u][6] = q1.z; v[u][7] = q1.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < LD_U; ++u) {
        const int e = e0 + u * TC_THREADS;
        if (e >= halo_px) break;
        if (from1 && a.mask1) {
          const float m[8] = {mk[u][0].x, mk[u][0].y, mk[u][0].z, mk[u][0].w, mk[u][1].x, mk[u][1].y, mk[u][1].z, mk[u][1].w};
#pragma unroll
          for (int k = 0; k < 8; ++k) v[u][k] = m[k] > 0.f ? v[u][k] : 0.f;
        }
        if (from1 && a.relu1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[u][k] = fmaxf(v[u][k], 0.f);
        }
        uint4 hi, lo;
        split_bf16x8(v[u], hi, lo);
        *reinterpret_cast<uint4*>(st + (size_t)e * 16) = hi;
        *reinterpret_cast<uint4*>(st + plane_bytes + (size_t)e * 16) = lo;
      }
    }
    // ---- this plane's weight image (already bf16, already in UMMA layout) ----
    {
      const uint4* src = reinterpret_cast<const uint4*>(wtc) + (size_t)p * (t.w_bytes >> 4);
      uint4* dst = reinterpret_cast<uint4*>(st + t.in_bytes);
      for (int e = tid; e < (int)(t.w_bytes >> 4); e += TC_THREADS) dst[e] = __ldg(src + e);
    }
    fence_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    // ---- one thread issues every MMA of this plane ----
    if (tid == 0) {
      tc_fence_after();
      const uint32_t in_addr = smem_u32(st);
      const uint32_t w_addr = in_addr + t.in_bytes;
      const uint32_t wimg16 = (uint32_t)t.N * 2;          // one B image (2 K-chunks x N rows x 16 B) in 16-B units
      // descriptors differ only in their start-address field: build them once, then add (bytes >> 4)
      const uint64_t a1_base = make_desc(in_addr, plane_bytes, row_pitch);
      const uint64_t b_base = make_desc(w_addr, (uint32_t)t.N * 16, 128);
      const uint64_t a3_hi = ((uint64_t)((row_pitch >> 4) & 0x3FFF) << 32) | (1ull << 46);
      for (int tile = 0; tile < t.T; ++tile) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(tile * t.N);
        const uint32_t tile16 = (uint32_t)tile * 8;       // 8 pixels to the right, in 16-B units
        uint32_t acc = (p > 0) ? 1u : 0u;
        for (int tap = 0; tap < t.n1; ++tap) {
          tc_mma(d_tmem, a1_base + tile16 + tap_off16[tap], b_base + (uint32_t)tap * wimg16, idesc, acc);
          acc = 1u;
        }
        for (int j = 0; j < t.n3; ++j) {
          const int ta = 2 * j, tb = (2 * j + 1 < t.n1) ? 2 * j + 1 : 2 * j;
          const uint32_t offa = tap_off16[ta], offb = tap_off16[tb];
          const uint32_t lbo16 = tb == ta ? 1u : offb - offa;           // dummy second chunk multiplies zero weights
          const uint64_t ad = a3_hi | (uint64_t)(((in_addr >> 4) + tile16 + offa) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16);
          tc_mma(d_tmem, ad, b_base + (uint32_t)(t.n1 + j) * wimg16, idesc, 1u);
        }
      }
      tc_commit(&bar_free[s]);                 // arrives when every MMA issued so far has retired
      if (p == t.P - 1) tc_commit(&bar_done);
    }
  }
  // ---- epilogue: TMEM -> registers -> global ----
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  {
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int row = q * 4 + (lane >> 3);       // lane i of the accumulator = pixel (i / 8, i % 8) of the tile
    const int colx = lane & 7;
    const int oy = y0 + row;
    const int half = warp >> 2;                // warps 0-3: first half of the tiles, 4-7: second half
    const int t_lo = half ? (t.T + 1) / 2 : 0, t_hi = half ? t.T : (t.T + 1) / 2;
    for (int tile = t_lo; tile < t_hi; ++tile) {
      const int ox = x0 + tile * 8 + colx;
      const bool valid = oy < a.Hout && ox < a.Wout;
      const long pix = ((long)b * a.Hout + oy) * a.Wout + ox;
      for (int c0 = 0; c0 < a.coutp; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * t.N + c0), v);   // warp-collective
        if (!valid) continue;
        const int nq = min(4, (a.coutp - c0) >> 2);
        for (int g = 0; g < nq; ++g) {
          const int co = c0 + g * 4;
          float4 r4 = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          if (a.bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + co));
            r4.x += bv.x; r4.y += bv.y; r4.z += bv.z; r4.w += bv.w;
          }
          if (a.relu) { r4.x = fmaxf(r4.x, 0.f); r4.y = fmaxf(r4.y, 0.f); r4.z = fmaxf(r4.z, 0.f); r4.w = fmaxf(r4.w, 0.f); }
          if (a.res) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(a.res + pix * a.pr + co));
            r4.x += r.x; r4.y += r.y; r4.z += r.z; r4.w += r.w;
          }
          if (a.relu2) { r4.x = fmaxf(r4.x, 0.f); r4.y = fmaxf(r4.y, 0.f); r4.z = fmaxf(r4.z, 0.f); r4.w = fmaxf(r4.w, 0.f); }
          if (a.omask) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(a.omask + pix * a.pom + co));
            r4.x = m.x > 0.f ? r4.x : 0.f; r4.y = m.y > 0.f ? r4.y : 0.f;
            r4.z = m.z > 0.f ? r4.z : 0.f; r4.w = m.w > 0.f ? r4.w : 0.f;
          }
          if (a.add) {
            float4 r = __ldg(reinterpret_cast<const float4*>(a.add + pix * a.pa + co));
            if (a.addmask) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(a.addmask + pix * a.pam + co));
              r.x = m.x > 0.f ? r.x : 0.f; r.y = m.y > 0.f ? r.y : 0.f;
              r.z = m.z > 0.f ? r.z : 0.f; r.w = m.w > 0.f ? r.w : 0.f;
            }
            r4.x += r.x; r4.y += r.y; r4.z += r.z; r4.w += r.w;
          }
          float4* dst = reinterpret_cast<float4*>(a.out + pix * a.po + co);
          if (a.accumulate) {
            const float4 o = *dst;
            r4.x += o.x; r4.y += o.y; r4.z += o.z; r4.w += o.w;
          }
          *dst = r4;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(t.tmem_cols) : "memory");
  }
}

bool conv_tc_supported(const ConvArgs& a) {
  if (a.stride != 1 || a.osy != 1 || a.oy0 != 0 || a.ox0 != 0) return false;
  if (a.Hq != a.Hin || a.Wq != a.Win || a.Hout != a.Hin || a.Wout != a.Win) return false;
  if ((a.c1 & 7) || (a.c2 & 7) || (a.coutp & 7) || a.coutp > 128) return false;
  if (!a.src1_nchw && (a.p1 & 3)) return false;
  if (a.Win < 8) return false;
  return true;
}

int tc_weight_floats_equiv(int taps, int cin, int coutp) {
  const int N = coutp < 16 ? 16 : round_up(coutp, 16);
  const long bytes = (long)(cin / 8) * (taps + (taps + 1) / 2) * N * 32;
  return (int)((bytes + 3) / 4);
}

int launch_conv_tc(const ConvArgs& a, const uint16_t* wtc, cudaStream_t st) {
  MSAU_CHECK_ARG(conv_tc_supported(a), "conv_tc: unsupported shape");
  TcTile t;
  t.N = a.coutp < 16 ? 16 : round_up(a.coutp, 16);
  t.T = t.N <= 32 ? 8 : (t.N <= 64 ? 4 : 2);
  while (t.T > 1 && 8 * (t.T / 2) >= a.Wout) t.T /= 2;       // narrow maps: do not pay for columns that do not exist
  const int taps = a.kh * a.kw;
  t.n1 = taps; t.n3 = (taps + 1) / 2;
  t.HH = 16 + (a.kh - 1) * a.dil;
  t.HW = 8 * t.T + (a.kw - 1) * a.dil;
  t.P = (a.c1 + a.c2) / 8;
  t.in_bytes = (uint32_t)t.HH * t.HW * 32;
  t.w_bytes = (uint32_t)(t.n1 + t.n3) * t.N * 32;
  t.stage_bytes = (t.in_bytes + t.w_bytes + 127) / 128 * 128;
  t.stages = t.P >= 2 ? 2 : 1;
  int cols = t.T * t.N;
  t.tmem_cols = 32;
  while ((int)t.tmem_cols < cols) t.tmem_cols <<= 1;
  const size_t smem = (size_t)t.stage_bytes * t.stages + 1024;
  MSAU_CHECK_ARG(smem <= 220 * 1024 && t.tmem_cols <= 512, "conv_tc: tile does not fit (smem %zu B, tmem %u cols)", smem, t.tmem_cols);
  static bool attr = false;
  if (!attr) {
    MSAU_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  dim3 grid(cdiv(a.Wout, 8 * t.T), cdiv(a.Hout, 16), a.B);
  MSAU_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "conv_tc: grid too large");
  const double npix = (double)a.B * a.Hin * a.Win;
  double bytes = npix * ((a.src1_nchw ? a.c1_logical : a.c1) + a.c2 + (a.mask1 ? a.c1 : 0)) * 4.0;
  bytes += npix * a.coutp * 4.0 * (1 + (a.res ? 1 : 0) + (a.omask ? 1 : 0) + (a.add ? 1 : 0) + (a.addmask ? 1 : 0) + (a.accumulate ? 1 : 0));
  ProfScope ps("conv_tc_kernel", 2.0 * npix * taps * (a.c1 + a.c2) * a.coutp, bytes, st);
  conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a, wtc, t);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// ------------------------------------------------------------------ weight images
// src: fp32 packed [taps][cin][coutp] (the layout conv.cu consumes).  dst (bf16), per 8-channel plane p:
//   [tap]           N rows x {chunk0 = Whi[8p..8p+7], chunk1 = same}            (type 1)
//   [pair j]        N rows x {chunk0 = Wlo(tap 2j), chunk1 = Wlo(tap 2j+1) | 0}  (type 3)
// one B image = [chunk0: N x 16 B][chunk1: N x 16 B]
__global__ void __launch_bounds__(256) pack_tc_kernel(const float* __restrict__ pk, uint16_t* __restrict__ pktc,
                                                       const TcPackDesc* __restrict__ descs, int n_desc) {
  const long blk = blockIdx.x;
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].blk0 <= blk) lo = mid; else hi = mid - 1;
  }
  const TcPackDesc d = descs[lo];
  const int N = d.N, taps = d.taps, n3 = (taps + 1) / 2;
  const long per_plane = (long)(taps + n3) * N * 16;             // bf16 elements
  const long total = (long)(d.cin / 8) * per_plane;
  const long e = (blk - d.blk0) * 256 + threadIdx.x;
  if (e >= total) return;
  const int p = (int)(e / per_plane);
  long r = e - (long)p * per_plane;
  const int img = (int)(r / (N * 16)); r -= (long)img * N * 16;
  const int chunk = (int)(r / (N * 8)); r -= (long)chunk * N * 8;
  const int n = (int)(r / 8), k = (int)(r - (long)n * 8);
  float w = 0.f;
  bool want_lo = false;
  int tap = -1;
  if (img < taps) { tap = img; }
  else {
    want_lo = true;
    tap = 2 * (img - taps) + chunk;
    if (tap >= taps) tap = -1;
  }
  if (tap >= 0 && n < d.coutp) w = pk[d.src_off + ((long)tap * d.cin + 8 * p + k) * d.coutp + n];
  const __nv_bfloat16 h = __float2bfloat16_rn(w);
  const __nv_bfloat16 out = want_lo ? __float2bfloat16_rn(w - __bfloat162float(h)) : h;
  pktc[d.dst_off + e] = __bfloat16_as_ushort(out);
}

int launch_pack_tc(const float* pk, uint16_t* pktc, const TcPackDesc* d_descs, int n_desc, long total_blocks, cudaStream_t st) {
  if (n_desc == 0) return MSAU_OK;
  ProfScope ps("pack_tc_kernel", 0, (double)total_blocks * 256 * 6.0, st);
  pack_tc_kernel<<<(unsigned)total_blocks, 256, 0, st>>>(pk, pktc, d_descs, n_desc);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

}  // namespace msau
