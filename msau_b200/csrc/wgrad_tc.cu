// Weight gradient on the tensor cores (tcgen05.mma, MN-major operands, accumulators resident in TMEM).
//
//     dW[ky][kx][ci][co] = sum_pixels X[pixel + offset(ky,kx)][ci] * dY[pixel][co]
// The reduction (K) dimension is the pixel index, so both operands are "MN-major": 8 channels contiguous
// (16 B) and 8 consecutive pixels stacked at 16-B pitch = exactly the planar bf16 tile that conv_tc.cu
// builds ([8-channel plane][row][col][8 ch]).  One instruction consumes 16 consecutive pixels of one image row.
//   M = 64 = 8 row-groups of 8 input channels (one channel plane); group i reads the tile shifted by i*dil
//       pixels, so groups 0..kw-1 ARE the kx taps of one kernel row (groups kw..7 are ignored padding of the
//       minimum M: with cout <= 64 the instruction is bound by the shared-memory operand read, not the math)
//   N = cout (multiple of 8)      K = 16 pixels
//   one TMEM accumulator per kernel row ky; a CTA keeps its kh accumulators resident over ALL the pixel tiles
//   it visits and reduces them into dW with fp32 atomics once, at the end.
// Precision: operands are rounded to bf16 once (round-to-nearest), products are exact, accumulation is fp32.
// Every dW element is a sum over >= 10^4..10^6 pixels, so the zero-mean rounding errors average out
// (measured relative error of dW ~1e-4, far below the ReLU-mask noise of the data-gradient path); the hi/lo
// split that the forward convolutions need would triple the operand traffic for nothing here.
// grid = (pixel-tile ranges, input-channel planes).
//
// Four kernels share this contraction; launch_wgrad_tc() picks, in this order:
//   wgrad_tc4_kernel  >= 16 input channels (and every dilated / > 85-column layer): several X planes per CTA, warp-specialised,
//                     whole-pixel loads, all taps (or all planes, 1x1) per instruction, vector reductions into dW     [round 2]
//   wgrad_tc3_kernel  one X plane, <= 16 output channels, NHWC: cp.async raw ring + converter warps (the 512^2 level)
//   wgrad_tc2_kernel  all taps per instruction with register-staged loads (NCHW first layer, narrow maps)
//   wgrad_tc_kernel   the round-1 kernel described above (one accumulator per ky, one CTA per plane): masked dY and leftovers
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "prof.cuh"

namespace msau {

// ---- PTX wrappers (same protocol as conv_tc.cu) ----
__device__ __forceinline__ uint32_t wsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wmbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wsmem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void wmbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = wsmem_u32(bar);
  // try_wait with a suspend-time hint parks the thread in hardware instead of spinning on issue slots
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void wtc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(wsmem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wtc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool welect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wtc_mma2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// no-swizzle descriptor; for MN-major operands LBO = stride between 8-row K groups, SBO = stride between MN groups
__device__ __forceinline__ uint64_t wmake_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void wtmem_ld16(uint32_t taddr, float* v) {
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
__device__ __forceinline__ void wtmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint4 wpack8(const float* x) {
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&v);
  }
  return make_uint4(h[0], h[1], h[2], h[3]);
}

struct WgTile {
  int TR, TC, HHx, HWx, N, ny_planes;       // tile rows/cols, X halo extent, MMA N, dY channel planes
  int xq, xr, yq, yr;                       // 256 elements ahead = (xq rows, xr cols) of the X tile / (yq, yr) of the dY tile
  int tiles_x, tiles_y, n_tiles, tiles_per_cta;
  uint32_t x_plane_bytes, y_plane_bytes, stage_bytes, tmem_cols;
  int dbg;                                  // MSAU_WG_DBG (timing experiments only): 1 = skip the loads after the first two tiles, 2 = skip the MMAs
};

static constexpr int WG_THREADS = 256;
#ifndef MSAU_WG_WU
#define MSAU_WG_WU 4
#endif
static constexpr int WU = MSAU_WG_WU;   // pixels whose global loads are in flight per thread

__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_kernel(const WgradArgs a, const WgTile t) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_free[2];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int plane = blockIdx.y;               // input-channel plane (8 channels) of this CTA
  const int tile0 = blockIdx.x * t.tiles_per_cta;
  const int tile1 = min(t.n_tiles, tile0 + t.tiles_per_cta);
  const bool do_bias = a.dbias != nullptr && plane == 0;
  if (a.skip_flag) {                               // (the flag is written by an earlier kernel of this step: wait for it first)
    pdl_wait();
    if (*a.skip_flag == 0) return;                 // one-hot input: first_layer.cu produced this gradient
  }

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wsmem_u32(&tmem_base_s)), "r"(t.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    wmbar_init(&bar_free[0], a.kh);
    wmbar_init(&bar_free[1], a.kh);
    wmbar_init(&bar_done, a.kh);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 128) sbias[tid] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();        // PDL protocol (common.cuh): nothing above reads or writes activations / gradients
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_s;
  if (tile0 >= tile1) {       // nothing to do (grid rounding): still free TMEM
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(t.tmem_cols) : "memory");
    return;
  }

  // bf16 x bf16 -> fp32, A and B both MN-major, M = 64
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(t.N >> 3) << 17) |
                         ((uint32_t)(64 >> 4) << 24);
  const int x_px = t.HHx * t.HWx;
  const int y_px = t.TR * t.TC;
  const int ca0 = plane << 3;
  float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  int it = 0;
  for (int tile = tile0; tile < tile1; ++tile, ++it) {
    const int s = it & 1;
    uint8_t* st = smem + (size_t)s * t.stage_bytes;
    if (it >= 2) {
      if (tid == 0) wmbar_wait(&bar_free[s], ((it >> 1) - 1) & 1);
      __syncthreads();
    }
    const int tx = tile % t.tiles_x;
    const int rest = tile / t.tiles_x;
    const int ty = rest % t.tiles_y;
    const int b = rest / t.tiles_y;
    const int qy0 = ty * t.TR, qx0 = tx * t.TC;
    // ---- X halo tile of this plane: (HHx x HWx) pixels of 16 B.  WU pixels' loads are in flight per thread ----
    const bool dbg_skip_ld = (t.dbg & 1) && it >= 2;
    if (!dbg_skip_ld) {
      const int in_y0 = qy0 - a.pada_t, in_x0 = qx0 - a.pada_l;
      uint8_t* xh = st;
      const int plane_stride = a.Ha * a.Wa;
      const float* xsrc = a.a_nchw ? a.A + ((long)b * a.ca_logical + ca0) * plane_stride : a.A + (long)b * plane_stride * a.pa + ca0;
      const int n_valid = a.ca_logical - ca0;
      int iy = tid / t.HWx, ix = tid - iy * t.HWx;        // element tid + k * 256 of the halo tile, advanced without divisions
      for (int e0 = tid; e0 < x_px; e0 += WG_THREADS * WU) {
        float v[WU][8];
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          const int e = e0 + u * WG_THREADS;
          const int gy = in_y0 + iy, gx = in_x0 + ix;
          iy += t.xq; ix += t.xr;
          if (ix >= t.HWx) { ix -= t.HWx; ++iy; }
          const bool inb = e < x_px && (unsigned)gy < (unsigned)a.Ha && (unsigned)gx < (unsigned)a.Wa;
          const int lin = gy * a.Wa + gx;
          if (a.a_nchw) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[u][k] = (inb && k < n_valid) ? __ldg(xsrc + lin + k * plane_stride) : 0.f;
          } else {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* sp = reinterpret_cast<const float4*>(xsrc + lin * a.pa);
            const float4 q0 = inb ? __ldg(sp) : z4;
            const float4 q1 = inb ? __ldg(sp + 1) : z4;
            v[u][0] = q0.x; v[u][1] = q0.y; v[u][2] = q0.z; v[u][3] = q0.w;
            v[u][4] = q1.x; v[u][5] = q1.y; v[u][6] = q1.z; v[u][7] = q1.w;
          }
        }
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          const int e = e0 + u * WG_THREADS;
          if (e >= x_px) break;
          if (a.reluA) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[u][k] = fmaxf(v[u][k], 0.f);
          }
          *reinterpret_cast<uint4*>(xh + e * 16) = wpack8(v[u]);
        }
      }
    }
    // ---- dY tile: [planes], each plane TR x TC pixels of 16 B ----
    if (!dbg_skip_ld) {
      uint8_t* yh = st + t.x_plane_bytes;
      const int total = y_px * t.ny_planes;
      const int pl = tid % t.ny_planes;            // constant per thread (256 % ny_planes == 0)
      // b_s2d: plane pl of the virtual tensor = phase (py, px), channels [c0, c0+8) of the physical one
      const int sph = a.b_s2d ? (a.b_col0 + (pl << 3)) / a.cph : 0;
      const int spy = sph >> 1, spx = sph & 1, smul = a.b_s2d ? 2 : 1;
      const float* ysrc = a.Bm + (long)b * a.Hb * a.Wb * a.pb + (a.b_s2d ? a.b_col0 + (pl << 3) - sph * a.cph : (pl << 3));
      const float* msrc = a.maskB ? a.maskB + (long)b * a.Hb * a.Wb * a.pmb + (pl << 3) : nullptr;
      uint8_t* ydst = yh + (size_t)pl * t.y_plane_bytes;
      int yr_ = (tid / t.ny_planes) / t.TC, yc_ = (tid / t.ny_planes) - yr_ * t.TC;   // pixel of element tid + k * 256
      for (int e0 = tid; e0 < total; e0 += WG_THREADS * WU) {
        float v[WU][8];
        float4 mk[WU][2];
        int so[WU];
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          const int e = e0 + u * WG_THREADS;
          const int r = yr_, c = yc_;
          so[u] = r * t.TC + c;
          yr_ += t.yq; yc_ += t.yr;
          if (yc_ >= t.TC) { yc_ -= t.TC; ++yr_; }
          const int vy = qy0 + r, vx = qx0 + c;
          const int gy = vy * smul + spy, gx = vx * smul + spx;
          const bool inb = e < total && vy < a.Hq && vx < a.Wq && gy < a.Hb && gx < a.Wb;
          const int lin = gy * a.Wb + gx;
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4* sp = reinterpret_cast<const float4*>(ysrc + lin * a.pb);
          const float4 q0 = inb ? __ldg(sp) : z4;
          const float4 q1 = inb ? __ldg(sp + 1) : z4;
          v[u][0] = q0.x; v[u][1] = q0.y; v[u][2] = q0.z; v[u][3] = q0.w;
          v[u][4] = q1.x; v[u][5] = q1.y; v[u][6] = q1.z; v[u][7] = q1.w;
          if (msrc) {
            const float4* mp = reinterpret_cast<const float4*>(msrc + lin * a.pmb);
            mk[u][0] = inb ? __ldg(mp) : z4;
            mk[u][1] = inb ? __ldg(mp + 1) : z4;
          }
        }
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          const int e = e0 + u * WG_THREADS;
          if (e >= total) break;
          if (msrc) {
            const float m[8] = {mk[u][0].x, mk[u][0].y, mk[u][0].z, mk[u][0].w, mk[u][1].x, mk[u][1].y, mk[u][1].z, mk[u][1].w};
#pragma unroll
            for (int k = 0; k < 8; ++k) v[u][k] = m[k] > 0.f ? v[u][k] : 0.f;
          }
          if (do_bias) {
#pragma unroll
            for (int k = 0; k < 8; ++k) bacc[k] += v[u][k];
          }
          *reinterpret_cast<uint4*>(ydst + so[u] * 16) = wpack8(v[u]);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // one issuing warp per kernel row (warps 0..kh-1, one elected lane each): each owns its own TMEM accumulator
    if (warp < a.kh) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (welect_one()) {
        const int ky = warp;
        const uint32_t xh = wsmem_u32(st);
        const uint32_t yh = xh + t.x_plane_bytes;
        const int chunks = t.TC >> 4;
        const uint32_t d_tmem = tmem_base + (uint32_t)(ky * t.N);
        uint32_t first = (it == 0) ? 0u : 1u;
        const uint32_t lbo = (128u >> 4) << 16;
        const uint32_t a_hi = (((uint32_t)a.dila * 16 >> 4) & 0x3FFF) | (1u << 14);
        const uint32_t b_hi = ((t.y_plane_bytes >> 4) & 0x3FFF) | (1u << 14);
        uint32_t a_lo = (((xh >> 4) + (uint32_t)(ky * a.dila * t.HWx)) & 0x3FFF) | lbo;
        uint32_t b_lo = ((yh >> 4) & 0x3FFF) | lbo;
        const uint32_t xrow16 = (uint32_t)t.HWx - (uint32_t)chunks * 16;   // row advance after the chunks, 16-B units
        for (int r = 0; r < ((t.dbg & 2) ? 0 : t.TR); ++r) {
          if (chunks == 4) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              wtc_mma2(d_tmem, a_lo + cc * 16, a_hi, b_lo + cc * 16, b_hi, idesc, cc == 0 ? first : 1u);
            }
            first = 1u;
            a_lo += 64; b_lo += 64;
          } else {
            for (int cc = 0; cc < chunks; ++cc) {
              wtc_mma2(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first);
              first = 1u;
              a_lo += 16; b_lo += 16;                       // next 16 pixels = 256 B
            }
          }
          a_lo += xrow16;                                   // dY rows are dense: b_lo already points at the next row
        }
        wtc_commit(&bar_free[s]);
        if (tile == tile1 - 1) wtc_commit(&bar_done);
      }
      __syncwarp();
    }
  }
  // ---- bias gradient partials ----
  if (do_bias && !(t.dbg & 4)) {
    const int pl = tid % t.ny_planes;
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&sbias[pl * 8 + k], bacc[k]);
  }
  // ---- reduce the resident accumulators into dW ----
  if (tid == 0) wmbar_wait(&bar_done, 0);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 2 && !(t.dbg & 8)) {
    // M = 64 accumulator: row m lives in TMEM lane (m % 16) + 32 * (m / 16)  (cute tmem_frg, "half subpartition"
    // atom).  row m = kx * 8 + channel-in-plane -> warp 0 (lanes 0..15) holds kx 0,1 ; warp 1 (lanes 32..47) kx 2,3.
    const int kx = warp * 2 + (lane >> 3), ci = ca0 + (lane & 7);
    const bool mine = lane < 16 && kx < a.kw && ci < a.ca_lim;
    for (int ky = 0; ky < a.kh; ++ky) {
      for (int c0 = 0; c0 < t.N; c0 += 8) {
        float v[8];
        wtmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ky * t.N + c0), v);
        if (mine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int co = c0 + j;
            if (a.b_s2d) {
              // transposed conv: virtual tap (ty, tx) x phase (py, px) -> real tap: (0,0)->1, (0,1)->2, (1,1)->0, (1,0)->none
              const int vco = a.b_col0 + co, ph = vco / a.cph, c = vco - ph * a.cph;
              const int py = ph >> 1, px = ph & 1;
              const int rky = ky == 0 ? (py ? 2 : 1) : (py ? 0 : -1);
              const int rkx = kx == 0 ? (px ? 2 : 1) : (px ? 0 : -1);
              if (ph < 4 && c < a.cb_lim && rky >= 0 && rkx >= 0)
                atomicAdd(a.dW + (long)ci * a.s_ca + (long)c * a.s_cb + (rky * 3 + rkx), v[j]);
            } else if (co < a.cb_lim) atomicAdd(a.dW + (long)ci * a.s_ca + (long)co * a.s_cb + (ky * a.kw + kx), v[j]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (do_bias) {
    if (a.b_s2d) {
      if (tid < a.cb && ((a.b_col0 + tid) % a.cph) < a.cb_lim) atomicAdd(a.dbias + ((a.b_col0 + tid) % a.cph), sbias[tid]);
    } else if (tid < a.cb_lim) atomicAdd(a.dbias + tid, sbias[tid]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(t.tmem_cols) : "memory");
}


// =====================================================================================================
// dil = 1 variant: ONE instruction per 16 pixels covers all kh x kw taps.
//     dW[ky][kx][ci][co] = sum_q X[q_row][q_col + kx - pad_l][ci] * dY[q_row + pad_t - ky][q_col][co]
// (the sum re-indexed over the X pixel q instead of the output pixel).  The kx taps stay the M row-groups of
// the X operand (shift by one pixel per group); the ky taps become N column-groups of the dY operand: dY is
// staged as [row][channel plane][col][8 ch] with (kh-1) halo rows, so column-group j = (kyi, plane) sits exactly
// j * TC * 16 B after group 0 -> a plain MN-major descriptor with SBO = one staged row of one plane.
//   M = 64 (kx, ci)   N = kh * cout   K = 16 pixels   -> kh x fewer tcgen05.mma than the per-ky kernel above.
// Several warps issue concurrently (one instruction stream keeps the tensor pipe ~45 cycles/instruction busy, two or
// more reach its ~25 cycle floor for M = 64), each into its own accumulator; the copies are summed in the final reduce.
struct Wg2Tile {
  int TR, TC, HWx, N, ny_planes, TRy, n_issue;
  int tiles_x, tiles_y, n_tiles, tiles_per_cta;
  uint32_t x_bytes, stage_bytes, tmem_cols;
  int dbg;                                  // MSAU_WG_DEBUG experiments: 1 = no MMAs, 2 = no global loads
};

template <bool NCHW>
__global__ void __launch_bounds__(WG_THREADS, 3) wgrad_tc2_kernel(const WgradArgs a, const Wg2Tile t) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_free[2];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int plane = blockIdx.y;
  const int tile0 = blockIdx.x * t.tiles_per_cta;
  const int tile1 = min(t.n_tiles, tile0 + t.tiles_per_cta);
  const bool do_bias = a.dbias != nullptr && plane == 0;
  if (a.skip_flag) {                               // (the flag is written by an earlier kernel of this step: wait for it first)
    pdl_wait();
    if (*a.skip_flag == 0) return;                 // one-hot input: first_layer.cu produced this gradient
  }

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wsmem_u32(&tmem_base_s)), "r"(t.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    wmbar_init(&bar_free[0], t.n_issue);
    wmbar_init(&bar_free[1], t.n_issue);
    wmbar_init(&bar_done, t.n_issue);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 128) sbias[tid] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();        // PDL protocol (common.cuh): nothing above reads or writes activations / gradients
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_s;
  if (tile0 >= tile1) {
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(t.tmem_cols) : "memory");
    return;
  }

  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(t.N >> 3) << 17) |
                         ((uint32_t)(64 >> 4) << 24);
    const int ca0 = plane << 3;

  int it = 0;
  for (int tile = tile0; tile < tile1; ++tile, ++it) {
    const int s = it & 1;
    uint8_t* st = smem + (size_t)s * t.stage_bytes;
    if (it >= 2) {
      if (tid == 0) wmbar_wait(&bar_free[s], ((it >> 1) - 1) & 1);
      __syncthreads();
    }
    const int tx = tile % t.tiles_x;
    const int rest = tile / t.tiles_x;
    const int ty = rest % t.tiles_y;
    const int b = rest / t.tiles_y;
    const int qy0 = ty * t.TR, qx0 = tx * t.TC;
    // ---- X tile of this plane: TR rows x HWx columns (kw-1 halo columns), 16 B per pixel.  Kept lean (this kernel is
    //      bound by instruction issue, not by HBM): a warp takes rows warp, warp+8, ..., a lane the columns lane and
    //      lane+32; pointers advance by constants; the few halo columns are a separate short pass ----
    {
      const int in_x0 = qx0 - a.pada_l;
      if (NCHW) {
        const int plane_stride = a.Ha * a.Wa;
        const float* xsrc = a.A + ((long)b * a.ca_logical + ca0) * plane_stride;
        const int n_valid = a.ca_logical - ca0;
        for (int r = warp; r < t.TR; r += 8) {
          const int gy = qy0 + r;
          for (int c = lane; c < t.HWx; c += 32) {
            const int gx = in_x0 + c;
            const bool inb = gy < a.Ha && (unsigned)gx < (unsigned)a.Wa && !(t.dbg & 2);
            const float* sp = xsrc + gy * a.Wa + gx;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (inb && k < n_valid) ? __ldg(sp + k * plane_stride) : 0.f;
            *reinterpret_cast<uint4*>(st + (r * t.HWx + c) * 16) = wpack8(v);
          }
        }
      } else {
        const float* xsrc = a.A + (long)b * a.Ha * a.Wa * a.pa + ca0;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        // interior columns [0, TC): 2 rows x 2 column groups = 4 pixels in flight per thread
        const int gx0 = in_x0 + lane, gx1 = gx0 + 32;
        const bool okx0 = lane < t.TC && (unsigned)gx0 < (unsigned)a.Wa && !(t.dbg & 2), okx1 = lane + 32 < t.TC && (unsigned)gx1 < (unsigned)a.Wa && !(t.dbg & 2);
        for (int r = warp; r < t.TR; r += 16) {
          const int r2 = r + 8;
          const bool oky0 = qy0 + r < a.Ha, oky1 = r2 < t.TR && qy0 + r2 < a.Ha;
          const float* p0 = xsrc + ((qy0 + r) * a.Wa + gx0) * a.pa;
          const float* p1 = p0 + 8 * a.Wa * a.pa;
          const float4* s00 = reinterpret_cast<const float4*>(p0);
          const float4* s01 = reinterpret_cast<const float4*>(p0 + 32 * a.pa);
          const float4* s10 = reinterpret_cast<const float4*>(p1);
          const float4* s11 = reinterpret_cast<const float4*>(p1 + 32 * a.pa);
          float4 q[4][2];
          q[0][0] = (oky0 && okx0) ? __ldg(s00) : z4; q[0][1] = (oky0 && okx0) ? __ldg(s00 + 1) : z4;
          q[1][0] = (oky0 && okx1) ? __ldg(s01) : z4; q[1][1] = (oky0 && okx1) ? __ldg(s01 + 1) : z4;
          q[2][0] = (oky1 && okx0) ? __ldg(s10) : z4; q[2][1] = (oky1 && okx0) ? __ldg(s10 + 1) : z4;
          q[3][0] = (oky1 && okx1) ? __ldg(s11) : z4; q[3][1] = (oky1 && okx1) ? __ldg(s11 + 1) : z4;
          uint8_t* d0 = st + (r * t.HWx + lane) * 16;
          uint8_t* d1 = d0 + 8 * t.HWx * 16;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float v[8] = {q[u][0].x, q[u][0].y, q[u][0].z, q[u][0].w, q[u][1].x, q[u][1].y, q[u][1].z, q[u][1].w};
            if (a.reluA) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
            }
            const bool st_ok = (u < 2 || r2 < t.TR) && ((u & 1) ? lane + 32 < t.TC : lane < t.TC);
            if (st_ok) *reinterpret_cast<uint4*>((u < 2 ? d0 : d1) + (u & 1) * 512) = wpack8(v);
          }
        }
        // halo columns [TC, HWx): at most 3 per row
        const int nh = t.HWx - t.TC;
        for (int e = tid; e < t.TR * nh; e += WG_THREADS) {
          const int r = e / nh, c = t.TC + (e - r * nh);
          const int gy = qy0 + r, gx = in_x0 + c;
          const bool inb = gy < a.Ha && (unsigned)gx < (unsigned)a.Wa && !(t.dbg & 2);
          const float4* sp = reinterpret_cast<const float4*>(xsrc + (gy * a.Wa + gx) * a.pa);
          const float4 q0 = inb ? __ldg(sp) : z4, q1 = inb ? __ldg(sp + 1) : z4;
          float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
          if (a.reluA) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
          }
          *reinterpret_cast<uint4*>(st + (r * t.HWx + c) * 16) = wpack8(v);
        }
      }
    }
    // ---- dY tile: [TRy rows][planes][TC cols], rows qy0 + pad_t - (kh-1) .. qy0 + pad_t + TR - 1.
    //      A warp owns one channel plane (two when there are 16) and every (8 / planes)-th row of it ----
    {
      uint8_t* yh = st + t.x_bytes;
      const int nyp = t.ny_planes;
      const int pgrp = nyp < 8 ? nyp : 8;
      const int pl0 = warp % pgrp, rstart = warp / pgrp, rstep = 8 / pgrp;
      const int vy0 = qy0 + a.pada_t - (a.kh - 1);
      const int smul = a.b_s2d ? 2 : 1;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool two = lane + 32 < t.TC;
      for (int pl = pl0; pl < nyp; pl += 8) {
        // b_s2d: plane pl of the virtual tensor = phase (py, px), channels [c0, c0+8) of the physical one
        const int sph = a.b_s2d ? (a.b_col0 + (pl << 3)) / a.cph : 0;
        const int spy = sph >> 1, spx = sph & 1;
        const float* ysrc = a.Bm + (long)b * a.Hb * a.Wb * a.pb + (a.b_s2d ? a.b_col0 + (pl << 3) - sph * a.cph : (pl << 3));
        const int vx0 = qx0 + lane, vx1 = vx0 + 32;
        const int gx0 = vx0 * smul + spx, gx1 = vx1 * smul + spx;
        const bool okx0 = lane < t.TC && vx0 < a.Wq && gx0 < a.Wb && !(t.dbg & 2), okx1 = two && vx1 < a.Wq && gx1 < a.Wb && !(t.dbg & 2);
        float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = rstart; r < t.TRy; r += 2 * rstep) {
          const int r2 = r + rstep;
          const int vya = vy0 + r, vyb = vy0 + r2;
          const int gya = vya * smul + spy, gyb = vyb * smul + spy;
          const bool okya = (unsigned)vya < (unsigned)a.Hq && gya < a.Hb, okyb = r2 < t.TRy && (unsigned)vyb < (unsigned)a.Hq && gyb < a.Hb;
          const float4* s00 = reinterpret_cast<const float4*>(ysrc + (gya * a.Wb + gx0) * a.pb);
          const float4* s01 = reinterpret_cast<const float4*>(ysrc + (gya * a.Wb + gx1) * a.pb);
          const float4* s10 = reinterpret_cast<const float4*>(ysrc + (gyb * a.Wb + gx0) * a.pb);
          const float4* s11 = reinterpret_cast<const float4*>(ysrc + (gyb * a.Wb + gx1) * a.pb);
          float4 q[4][2];
          q[0][0] = (okya && okx0) ? __ldg(s00) : z4; q[0][1] = (okya && okx0) ? __ldg(s00 + 1) : z4;
          q[1][0] = (okya && okx1) ? __ldg(s01) : z4; q[1][1] = (okya && okx1) ? __ldg(s01 + 1) : z4;
          q[2][0] = (okyb && okx0) ? __ldg(s10) : z4; q[2][1] = (okyb && okx0) ? __ldg(s10 + 1) : z4;
          q[3][0] = (okyb && okx1) ? __ldg(s11) : z4; q[3][1] = (okyb && okx1) ? __ldg(s11 + 1) : z4;
          // bias gradient: rows of this tile only (halo rows belong to the neighbours); out-of-image pixels are zero
          const bool ca = do_bias && vya >= qy0 && vya < qy0 + t.TR, cb2 = do_bias && vyb >= qy0 && vyb < qy0 + t.TR;
          uint8_t* d0 = yh + (size_t)((r * nyp + pl) * t.TC + lane) * 16;
          uint8_t* d1 = yh + (size_t)((r2 * nyp + pl) * t.TC + lane) * 16;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float v[8] = {q[u][0].x, q[u][0].y, q[u][0].z, q[u][0].w, q[u][1].x, q[u][1].y, q[u][1].z, q[u][1].w};
            if (u < 2 ? ca : cb2) {
#pragma unroll
              for (int k = 0; k < 8; ++k) bacc[k] += v[k];
            }
            const bool st_ok = (u < 2 || r2 < t.TRy) && ((u & 1) ? two : lane < t.TC);
            if (st_ok) *reinterpret_cast<uint4*>((u < 2 ? d0 : d1) + (u & 1) * 512) = wpack8(v);
          }
        }
        if (do_bias) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float sum = bacc[k];
#pragma unroll
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) atomicAdd(&sbias[pl * 8 + k], sum);
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // issuing warps 0..n_issue-1 take the X rows r = warp, warp + n_issue, ... ; each accumulates into its own copy
    if (warp < t.n_issue) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (welect_one()) {
        const uint32_t xh = wsmem_u32(st);
        const uint32_t yh = xh + t.x_bytes;
        const int chunks = t.TC >> 4;
        const uint32_t d_tmem = tmem_base + (uint32_t)(warp * t.N);
        const uint32_t lbo = (128u >> 4) << 16;
        const uint32_t a_hi = 1u | (1u << 14);                                   // SBO = one pixel per kx group
        const uint32_t b_hi = (((uint32_t)t.TC * 16 >> 4) & 0x3FFF) | (1u << 14);  // SBO = one staged row of one plane
        const uint32_t yrow16 = (uint32_t)(t.ny_planes * t.TC);                  // staged dY row pitch, 16-B units
        uint32_t first = (it == 0) ? 0u : 1u;
        for (int r = warp; r < t.TR; r += t.n_issue) {
          uint32_t a_lo = (((xh >> 4) + (uint32_t)(r * t.HWx)) & 0x3FFF) | lbo;
          uint32_t b_lo = (((yh >> 4) + (uint32_t)r * yrow16) & 0x3FFF) | lbo;
          for (int cc = 0; cc < chunks; ++cc) {
            if (!(t.dbg & 1) || first == 0u) wtc_mma2(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first);
            first = 1u;
            a_lo += 16; b_lo += 16;
          }
        }
        wtc_commit(&bar_free[s]);
        if (tile == tile1 - 1) wtc_commit(&bar_done);
      }
      __syncwarp();
    }
  }
  // ---- reduce the resident accumulators into dW ----
  if (tid == 0) wmbar_wait(&bar_done, 0);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 2) {
    // M = 64 accumulator: row m = kx * 8 + ci lives in TMEM lane (m % 16) + 32 * (m / 16)
    const int kx = warp * 2 + (lane >> 3), ci = ca0 + (lane & 7);
    const bool mine = lane < 16 && kx < a.kw && ci < a.ca_lim;
    const int n_iss = min(t.n_issue, t.TR);       // warps that issued at least one instruction (their accumulators are defined)
    for (int c0 = 0; c0 < t.N; c0 += 8) {
      float v[8];
      wtmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      for (int i = 1; i < n_iss; ++i) {
        float w[8];
        wtmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(i * t.N + c0), w);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += w[j];
      }
      if (mine) {
        const int kyi = c0 / a.cb, cb0 = c0 - kyi * a.cb;     // column group = (kyi, 8-channel plane); ky = kh-1-kyi
        const int ky = a.kh - 1 - kyi;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int co = cb0 + j;
          if (a.b_s2d) {
            const int vco = a.b_col0 + co, ph = vco / a.cph, c = vco - ph * a.cph;
            const int py = ph >> 1, px = ph & 1;
            const int rky = ky == 0 ? (py ? 2 : 1) : (py ? 0 : -1);
            const int rkx = kx == 0 ? (px ? 2 : 1) : (px ? 0 : -1);
            if (ph < 4 && c < a.cb_lim && rky >= 0 && rkx >= 0)
              atomicAdd(a.dW + (long)ci * a.s_ca + (long)c * a.s_cb + (rky * 3 + rkx), v[j]);
          } else if (co < a.cb_lim) atomicAdd(a.dW + (long)ci * a.s_ca + (long)co * a.s_cb + (ky * a.kw + kx), v[j]);
        }
      }
    }
  }
  __syncthreads();
  if (do_bias) {
    if (a.b_s2d) {
      if (tid < a.cb && ((a.b_col0 + tid) % a.cph) < a.cb_lim) atomicAdd(a.dbias + ((a.b_col0 + tid) % a.cph), sbias[tid]);
    } else if (tid < a.cb_lim) atomicAdd(a.dbias + tid, sbias[tid]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(t.tmem_cols) : "memory");
}

static bool wgrad_tc2_config(const WgradArgs& a, Wg2Tile& t) {
  // measured (profiles/): wins for <= 16 output channels (the HBM-streaming levels); with more channels the (kh-1) halo rows
  // of the staged dY tile cost shared memory (fewer resident CTAs) and the per-ky kernel above is faster
  if (a.dila != 1 || a.cb > 16 || a.maskB) return false;
  t.N = a.kh * a.cb;
  t.ny_planes = a.cb >> 3;
  t.TC = round_up(a.Wq, 16);
  if (t.TC > 64) t.TC = 64;
  const int row_bytes = t.TC * a.cb * 2;                  // one staged dY row, all planes
  t.TR = 48 * 1024 / row_bytes - (a.kh - 1);
  if (t.TR > 16) t.TR = 16;
  if (t.TR > a.Hq) t.TR = a.Hq;
  if (t.TR < 1) return false;
  t.TRy = t.TR + a.kh - 1;
  t.HWx = t.TC + (a.kw - 1);
  const int slack_px = 7 + 16;                            // M groups kw..7 read past the useful columns
  t.x_bytes = (uint32_t)((t.TR * t.HWx + slack_px) * 16 + 127) / 128 * 128;
  const uint32_t y_bytes = (uint32_t)t.TRy * row_bytes;
  t.stage_bytes = (t.x_bytes + y_bytes + 1023) / 1024 * 1024;
  t.n_issue = 128 / t.N;
  if (t.n_issue > 4) t.n_issue = 4;
  if (t.n_issue < 1) t.n_issue = 1;
  if (t.n_issue > t.TR) t.n_issue = t.TR;
  const int cols = t.n_issue * t.N;
  t.tmem_cols = 32;
  while ((int)t.tmem_cols < cols) t.tmem_cols <<= 1;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MSAU_WG_DEBUG"); dbg = e ? atoi(e) : 0; } t.dbg = dbg; }
  t.tiles_x = cdiv(a.Wq, t.TC);
  t.tiles_y = cdiv(a.Hq, t.TR);
  t.n_tiles = t.tiles_x * t.tiles_y * a.B;
  return (size_t)t.stage_bytes * 2 + 1024 <= 200 * 1024 && t.tmem_cols <= 512;
}

static int launch_wgrad_tc2(const WgradArgs& a, const Wg2Tile& t0, cudaStream_t st) {
  Wg2Tile t = t0;
  const int planes = a.ca >> 3;
  const size_t smem = (size_t)t.stage_bytes * 2 + 1024;
  int per_sm = (int)((220 * 1024) / (smem + 1024));       // resident CTAs per SM by shared memory ...
  if (per_sm > 512 / (int)t.tmem_cols) per_sm = 512 / (int)t.tmem_cols;   // ... and by TMEM columns
  if (per_sm > 3) per_sm = 3;                              // 80 registers x 256 threads
  if (per_sm < 1) per_sm = 1;
  int ctas = (per_sm * sm_count() + planes - 1) / planes;
  if (ctas > t.n_tiles) ctas = t.n_tiles;
  if (ctas < 1) ctas = 1;
  t.tiles_per_cta = cdiv(t.n_tiles, ctas);
  ctas = cdiv(t.n_tiles, t.tiles_per_cta);
  static bool attr = false;
  if (!attr) {
    MSAU_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MSAU_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid(ctas, planes);
  const double npq = (double)a.B * a.Hq * a.Wq;
  const double wbytes = (npq * (a.a_nchw ? a.ca_logical : a.ca) + (double)a.B * a.Hb * a.Wb * (a.b_s2d ? a.cph : a.cb) * (a.maskB ? 2 : 1)) * 4.0;
  ProfScope ps(a.skip_flag ? "dense_first_layer_skippable" : "wgrad_tc2_kernel", a.ca, a.cb, a.kh, a.dila, a.Wq, a.a_nchw, a.skip_flag ? 0.0 : 2.0 * npq * a.kh * a.kw * a.ca * a.cb, a.skip_flag ? 0.0 : wbytes, st);
  if (a.a_nchw) MSAU_CUDA_TRY(launch_pdl(wgrad_tc2_kernel<true>, grid, dim3(WG_THREADS), smem, st, a, t));
  else MSAU_CUDA_TRY(launch_pdl(wgrad_tc2_kernel<false>, grid, dim3(WG_THREADS), smem, st, a, t));
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}


// =====================================================================================================
// Pipelined variant of the kernel above for the HBM-streaming levels (<= 16 output channels, NHWC operands):
// one persistent CTA per SM; 8 converter warps keep D-1 tiles of raw fp32 operands in flight with cp.async
// (thread-private shared-memory slots, zero-fill for out-of-image pixels) and turn the oldest one into the bf16
// operand images; 4 MMA warps issue the instructions (one per 16 pixels, all taps) into their own accumulators.
// Tile = TR x 64 pixels of X (+ kw-1 halo columns) and 8 staged rows of dY per channel plane (TR = 8 / planes - (kh-1)),
// so every converter warp owns exactly one staged dY (row, plane) pair and at most one X row.
struct Wg3Tile {
  int TR, HWx, N, nyp, TRy, n_issue, D, S, items;   // X rows, X row pitch (px), MMA N, dY planes, staged dY rows, issuers, raw / bf16 ring depth, slots/thread
  int ix_halo, iy0;                              // slot index of the X halo item (-1: none) and of the first dY item
  int tiles_x, tiles_y, n_tiles, tiles_per_cta;
  uint32_t x_bytes, stage_bytes, raw_bytes, tmem_cols;
};
static constexpr int WG3_CONV_WARPS = 16;  // two per tile row (32 pixels each): a warp's ~200-instruction serial chain per tile is the pace
static constexpr int WG3_ROWS = 8;          // staged dY rows per tile
static constexpr int WG3_MMA_WARPS = 4;
static constexpr int WG3_THREADS = (WG3_CONV_WARPS + WG3_MMA_WARPS) * 32;
static constexpr int WG3_TC = 64;

__device__ __forceinline__ void wcp_async16z(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void wcp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void wcp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void wmbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(wsmem_u32(bar)) : "memory");
}

// polling wait for the single lanes that wait on behalf of their warp (no suspend hint: immediate wake-up)
__device__ __forceinline__ void wmbar_wait_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = wsmem_u32(bar);
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

__global__ void __launch_bounds__(WG3_THREADS, 1) wgrad_tc3_kernel(const WgradArgs a, const Wg3Tile t) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[4];      // converters -> MMA : bf16 stage ready
  __shared__ uint64_t bar_free[4];      // MMA -> converters : the instructions reading the stage have retired
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int plane = blockIdx.y;
  const int tile0 = blockIdx.x * t.tiles_per_cta;
  const int tile1 = min(t.n_tiles, tile0 + t.tiles_per_cta);
  const int n_my = tile1 - tile0;
  const bool do_bias = a.dbias != nullptr && plane == 0;
  if (a.skip_flag) {                               // (the flag is written by an earlier kernel of this step: wait for it first)
    pdl_wait();
    if (*a.skip_flag == 0) return;                 // one-hot input: first_layer.cu produced this gradient
  }

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wsmem_u32(&tmem_base_s)), "r"(t.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int i = 0; i < 4; ++i) { wmbar_init(&bar_full[i], WG3_CONV_WARPS); wmbar_init(&bar_free[i], t.n_issue); }
    wmbar_init(&bar_done, t.n_issue);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) sbias[tid] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();        // PDL protocol (common.cuh): nothing above reads or writes activations / gradients
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_s;
  uint8_t* const stage_s = smem;                               // 2 x [X image | dY image] (bf16)
  uint8_t* const raw_s = smem + (size_t)t.S * t.stage_bytes;     // D x raw fp32 slots
  const int ca0 = plane << 3;

  if (n_my > 0 && warp < WG3_CONV_WARPS) {
    // =============================================================== converters
    // Work of this thread, the same for every tile.  Half a row (32 pixels x 32 B) is 64 chunks of 16 B; lane l copies the
    // chunks l and l+32 (pixel = chunk / 2, half = chunk & 1), so every cp.async warp instruction moves 512 contiguous
    // bytes and every 32-byte sector is requested once (cp.async.cg bypasses L1, which is a few KB next to 200 KB of
    // shared memory).  Warp w: row w / 2, columns 32 (w & 1) .. +31 of X (if row < TR) and of the staged dY row (all
    // channel planes); X halo: the first 2 * TR * (kw-1) threads take half a pixel each.
    // Items are 16 B: [X 0..1][halo][dY plane 0: 0..1][dY plane 1: 0..1].
    const uint32_t raw_u32 = wsmem_u32(raw_s) + (uint32_t)tid * 16u;
    const int nh = t.HWx - WG3_TC;
    const int wr = warp >> 1, wc = (warp & 1) * 32;              // row / first column of this warp
    const bool has_x = wr < t.TR;
    const bool has_h = t.ix_halo >= 0 && tid < 2 * t.TR * nh;
    const int hr = has_h ? (tid >> 1) / nh : 0, hc = has_h ? WG3_TC + ((tid >> 1) - hr * nh) : 0, hhalf = tid & 1;
    const int yr = wr;                                           // staged dY row of this warp
    const int half = lane & 1, px0 = wc + (lane >> 1);           // chunk l + 32 j -> pixel px0 + 16 j, same half
    constexpr uint32_t ITEM = WG3_CONV_WARPS * 32 * 16;          // bytes between this thread's consecutive raw items
    // Everything that does not depend on the tile is computed once: element offsets of this thread's items relative to the
    // tile origin (32-bit: every tensor here has < 2^31 elements), and the tile origin itself advances incrementally.
    const int xo = (wr * a.Wa + px0 - a.pada_l) * a.pa + ca0 + half * 4;        // X item 0 (item 1: + 16 pa)
    const int ho = (hr * a.Wa + hc - a.pada_l) * a.pa + ca0 + hhalf * 4;        // X halo item
    const int vrel = a.pada_t - (a.kh - 1) + yr;                                  // staged dY row relative to the tile's first row
    const int yo = (vrel * a.Wb + px0) * a.pb + half * 4;                         // dY item 0 (item 1: + 16 pb; plane 1: + 8)
    struct Cur { int tx, ty, b, xt, yt; };
    auto origin = [&](Cur& c) {
      c.xt = ((c.b * a.Ha + c.ty * t.TR) * a.Wa + c.tx * WG3_TC) * a.pa;
      c.yt = ((c.b * a.Hb + c.ty * t.TR) * a.Wb + c.tx * WG3_TC) * a.pb;
    };
    auto next = [&](Cur& c) {
      if (++c.tx == t.tiles_x) {
        c.tx = 0;
        if (++c.ty == t.tiles_y) { c.ty = 0; ++c.b; }
        origin(c);
      } else {
        c.xt += WG3_TC * a.pa;
        c.yt += WG3_TC * a.pb;
      }
    };
    Cur ahead;
    ahead.tx = tile0 % t.tiles_x; ahead.ty = (tile0 / t.tiles_x) % t.tiles_y; ahead.b = (tile0 / t.tiles_x) / t.tiles_y;
    origin(ahead);
    int n_ahead = 0;                                              // tiles issued so far
    uint32_t d_issue = 0;                                         // raw slot of the next issue, bytes
    auto issue = [&]() {
      if (n_ahead < n_my) {
        const int qy0 = ahead.ty * t.TR, qx0 = ahead.tx * WG3_TC;
        const uint32_t dst = raw_u32 + d_issue;
        if (has_x) {
          const bool oky = qy0 + wr < a.Ha;
          const int gx = qx0 - a.pada_l + px0;
          const bool in0 = oky && (unsigned)gx < (unsigned)a.Wa, in1 = oky && (unsigned)(gx + 16) < (unsigned)a.Wa;
          const float* sp = a.A + (long)(ahead.xt + xo);      // may point left of the row (halo): only dereferenced where in-bounds
          wcp_async16z(dst, in0 ? sp : a.A, in0 ? 16u : 0u);
          wcp_async16z(dst + ITEM, in1 ? sp + 16 * a.pa : a.A, in1 ? 16u : 0u);
        }
        if (has_h) {
          const bool inb = qy0 + hr < a.Ha && (unsigned)(qx0 - a.pada_l + hc) < (unsigned)a.Wa;
          wcp_async16z(dst + (uint32_t)t.ix_halo * ITEM, inb ? a.A + (long)(ahead.xt + ho) : a.A, inb ? 16u : 0u);
        }
        {
          const bool oky = (unsigned)(qy0 + vrel) < (unsigned)a.Hq;
          const int vx = qx0 + px0;
          const bool in0 = oky && vx < a.Wq, in1 = oky && vx + 16 < a.Wq;
          const float* sp = a.Bm + (long)(ahead.yt + yo);
          const uint32_t dy = dst + (uint32_t)t.iy0 * ITEM;
          wcp_async16z(dy, in0 ? sp : a.Bm, in0 ? 16u : 0u);
          wcp_async16z(dy + ITEM, in1 ? sp + 16 * a.pb : a.Bm, in1 ? 16u : 0u);
          if (t.nyp > 1) {
            wcp_async16z(dy + 2 * ITEM, in0 ? sp + 8 : a.Bm, in0 ? 16u : 0u);
            wcp_async16z(dy + 3 * ITEM, in1 ? sp + 16 * a.pb + 8 : a.Bm, in1 ? 16u : 0u);
          }
        }
        next(ahead);
      }
      ++n_ahead;
      d_issue += t.raw_bytes;
      if (d_issue == (uint32_t)t.D * t.raw_bytes) d_issue = 0;
      wcp_commit();
    };
    // 4 fp32 -> 4 bf16 (8 bytes)
    auto pack4 = [](const float4& q, bool relu) -> uint2 {
      float4 v = q;
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
      return make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    };
    for (int d = 0; d < t.D - 1; ++d) issue();
    float bacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const bool centre = do_bias && vrel >= 0 && vrel < t.TR;     // the halo rows belong to the neighbouring tiles
    const bool relu = a.reluA != 0;
    // destinations inside a bf16 stage (constant per thread)
    const uint32_t xd = (uint32_t)((wr * t.HWx + px0) * 16 + half * 8);
    const uint32_t hd = (uint32_t)((hr * t.HWx + hc) * 16 + hhalf * 8);
    const uint32_t yd = t.x_bytes + (uint32_t)((yr * t.nyp * WG3_TC + px0) * 16 + half * 8);
    int s = 0;
    uint32_t sphase = 0, d_cur = 0;
    for (int it = 0; it < n_my; ++it) {
      issue();
      if (t.D == 4) wcp_wait<3>(); else if (t.D == 3) wcp_wait<2>(); else wcp_wait<1>();
      if (it >= t.S) {
        if (lane == 0) wmbar_wait(&bar_free[s], sphase ^ 1u);
        __syncwarp();
      }
      uint8_t* st = stage_s + (size_t)s * t.stage_bytes;
      const uint8_t* rsrc = raw_s + tid * 16 + d_cur;
      if (has_x) {
        const float4 q0 = *reinterpret_cast<const float4*>(rsrc);
        const float4 q1 = *reinterpret_cast<const float4*>(rsrc + ITEM);
        *reinterpret_cast<uint2*>(st + xd) = pack4(q0, relu);
        *reinterpret_cast<uint2*>(st + xd + 256) = pack4(q1, relu);
      }
      if (has_h) {
        const float4 q = *reinterpret_cast<const float4*>(rsrc + t.ix_halo * ITEM);
        *reinterpret_cast<uint2*>(st + hd) = pack4(q, relu);
      }
      {
        const uint8_t* ry = rsrc + t.iy0 * ITEM;
        const float4 q0 = *reinterpret_cast<const float4*>(ry);
        const float4 q1 = *reinterpret_cast<const float4*>(ry + ITEM);
        if (centre) {
          bacc[0][0] += q0.x + q1.x; bacc[0][1] += q0.y + q1.y; bacc[0][2] += q0.z + q1.z; bacc[0][3] += q0.w + q1.w;
        }
        *reinterpret_cast<uint2*>(st + yd) = pack4(q0, false);
        *reinterpret_cast<uint2*>(st + yd + 256) = pack4(q1, false);
        if (t.nyp > 1) {
          const float4 p0 = *reinterpret_cast<const float4*>(ry + 2 * ITEM);
          const float4 p1 = *reinterpret_cast<const float4*>(ry + 3 * ITEM);
          if (centre) {
            bacc[1][0] += p0.x + p1.x; bacc[1][1] += p0.y + p1.y; bacc[1][2] += p0.z + p1.z; bacc[1][3] += p0.w + p1.w;
          }
          *reinterpret_cast<uint2*>(st + yd + WG3_TC * 16) = pack4(p0, false);
          *reinterpret_cast<uint2*>(st + yd + WG3_TC * 16 + 256) = pack4(p1, false);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) wmbar_arrive(&bar_full[s]);      // one arrival per warp (per-thread arrivals serialise)
      if (++s == t.S) { s = 0; sphase ^= 1u; }
      d_cur += t.raw_bytes;
      if (d_cur == (uint32_t)t.D * t.raw_bytes) d_cur = 0;
    }
    wcp_wait<0>();
    if (do_bias) {
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float sum = bacc[pl][k];
#pragma unroll
          for (int o = 16; o > 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);   // lanes of equal parity (same half)
          if (lane < 2 && pl < t.nyp) atomicAdd(&sbias[pl * 8 + half * 4 + k], sum);
        }
      }
    }
  } else if (n_my > 0 && warp - WG3_CONV_WARPS < t.n_issue) {
    // =============================================================== MMA issuers (X rows mw, mw + n_issue, ...)
    const int mw = warp - WG3_CONV_WARPS;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(t.N >> 3) << 17) |
                           ((uint32_t)(64 >> 4) << 24);
    const uint32_t d_tmem = tmem_base + (uint32_t)(mw * t.N);
    const uint32_t lbo = (128u >> 4) << 16;
    const uint32_t a_hi = 1u | (1u << 14);                                       // SBO = one pixel per kx group
    const uint32_t b_hi = (((uint32_t)WG3_TC * 16 >> 4) & 0x3FFF) | (1u << 14);  // SBO = one staged (row, plane) pair
    const uint32_t yrow16 = (uint32_t)(t.nyp * WG3_TC);                          // staged dY row pitch, 16-B units
    uint32_t first = 0u;
    for (int it = 0; it < n_my; ++it) {
      const int s = it % t.S;
      if (lane == 0) wmbar_wait(&bar_full[s], (it / t.S) & 1);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (welect_one()) {
        const uint32_t xh = wsmem_u32(stage_s + (size_t)s * t.stage_bytes);
        const uint32_t yh = xh + t.x_bytes;
        for (int r = mw; r < t.TR; r += t.n_issue) {
          uint32_t a_lo = (((xh >> 4) + (uint32_t)(r * t.HWx)) & 0x3FFF) | lbo;
          uint32_t b_lo = (((yh >> 4) + (uint32_t)r * yrow16) & 0x3FFF) | lbo;
#pragma unroll
          for (int cc = 0; cc < WG3_TC / 16; ++cc) {
            wtc_mma2(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first);
            first = 1u;
            a_lo += 16; b_lo += 16;
          }
        }
        wtc_commit(&bar_free[s]);
        if (it == n_my - 1) wtc_commit(&bar_done);
      }
      __syncwarp();
    }
  }
  // ---- reduce the resident accumulators into dW ----
  if (n_my > 0 && tid == 0) wmbar_wait(&bar_done, 0);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (n_my > 0 && warp < 2) {
    // M = 64 accumulator: row m = kx * 8 + ci lives in TMEM lane (m % 16) + 32 * (m / 16)
    const int kx = warp * 2 + (lane >> 3), ci = ca0 + (lane & 7);
    const bool mine = lane < 16 && kx < a.kw && ci < a.ca_lim;
    const int n_iss = min(t.n_issue, t.TR);
    for (int c0 = 0; c0 < t.N; c0 += 8) {
      float v[8];
      wtmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      for (int i = 1; i < n_iss; ++i) {
        float w[8];
        wtmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(i * t.N + c0), w);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += w[j];
      }
      if (mine) {
        const int kyi = c0 / a.cb, cb0 = c0 - kyi * a.cb;     // column group = (kyi, 8-channel plane); ky = kh-1-kyi
        const int ky = a.kh - 1 - kyi;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int co = cb0 + j;
          if (a.b_s2d) {
            const int vco = a.b_col0 + co, ph = vco / a.cph, c = vco - ph * a.cph;
            const int py = ph >> 1, px = ph & 1;
            const int rky = ky == 0 ? (py ? 2 : 1) : (py ? 0 : -1);
            const int rkx = kx == 0 ? (px ? 2 : 1) : (px ? 0 : -1);
            if (ph < 4 && c < a.cb_lim && rky >= 0 && rkx >= 0)
              atomicAdd(a.dW + (long)ci * a.s_ca + (long)c * a.s_cb + (rky * 3 + rkx), v[j]);
          } else if (co < a.cb_lim) atomicAdd(a.dW + (long)ci * a.s_ca + (long)co * a.s_cb + (ky * a.kw + kx), v[j]);
        }
      }
    }
  }
  __syncthreads();
  if (do_bias && n_my > 0) {
    if (a.b_s2d) {
      if (tid < a.cb && ((a.b_col0 + tid) % a.cph) < a.cb_lim) atomicAdd(a.dbias + ((a.b_col0 + tid) % a.cph), sbias[tid]);
    } else if (tid < a.cb_lim) atomicAdd(a.dbias + tid, sbias[tid]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(t.tmem_cols) : "memory");
}

static bool wgrad_tc3_config(const WgradArgs& a, Wg3Tile& t) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("MSAU_WG3_OFF"); off = e ? atoi(e) : 0; }
  if (off || a.dila != 1 || a.cb > 16 || a.maskB || a.a_nchw || a.b_s2d || a.Wq < 16) return false;
  t.nyp = a.cb >> 3;
  t.TRy = WG3_ROWS;
  t.TR = t.TRy - (a.kh - 1);
  if (t.TR < 2) return false;
  t.N = a.kh * a.cb;
  t.HWx = WG3_TC + (a.kw - 1);
  t.items = 2; t.ix_halo = -1;            // 16-byte items
  if (a.kw > 1) t.ix_halo = t.items++;
  t.iy0 = t.items; t.items += 2 * t.nyp;
  const int slack_px = 7 + 16;
  t.x_bytes = (uint32_t)((t.TR * t.HWx + slack_px) * 16 + 127) / 128 * 128;
  const uint32_t y_bytes = (uint32_t)(WG3_ROWS * t.nyp * WG3_TC * 16);
  t.stage_bytes = (t.x_bytes + y_bytes + 1023) / 1024 * 1024;
  t.raw_bytes = (uint32_t)(t.items * WG3_CONV_WARPS * 32 * 16);
  t.D = 4;
  t.S = 4;
  while ((t.D > 2 || t.S > 2) && (size_t)t.S * t.stage_bytes + (size_t)t.D * t.raw_bytes > 218 * 1024) {
    if (t.S > 2 && t.S >= t.D) --t.S; else --t.D;
  }
  t.n_issue = 128 / t.N;
  if (t.n_issue > WG3_MMA_WARPS) t.n_issue = WG3_MMA_WARPS;
  if (t.n_issue < 1) t.n_issue = 1;
  if (t.n_issue > t.TR) t.n_issue = t.TR;
  const int cols = t.n_issue * t.N;
  t.tmem_cols = 32;
  while ((int)t.tmem_cols < cols) t.tmem_cols <<= 1;
  t.tiles_x = cdiv(a.Wq, WG3_TC);
  t.tiles_y = cdiv(a.Hq, t.TR);
  t.n_tiles = t.tiles_x * t.tiles_y * a.B;
  return t.tmem_cols <= 512;
}

static int launch_wgrad_tc3(const WgradArgs& a, const Wg3Tile& t0, cudaStream_t st) {
  Wg3Tile t = t0;
  const int planes = a.ca >> 3;
  const size_t smem = (size_t)t.S * t.stage_bytes + (size_t)t.D * t.raw_bytes + 1024;
  int ctas = (sm_count() + planes - 1) / planes;
  if (ctas > t.n_tiles) ctas = t.n_tiles;
  if (ctas < 1) ctas = 1;
  t.tiles_per_cta = cdiv(t.n_tiles, ctas);
  ctas = cdiv(t.n_tiles, t.tiles_per_cta);
  static bool attr = false;
  if (!attr) {
    MSAU_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  dim3 grid(ctas, planes);
  const double npq = (double)a.B * a.Hq * a.Wq;
  const double wbytes = (npq * a.ca + (double)a.B * a.Hb * a.Wb * (a.b_s2d ? a.cph : a.cb)) * 4.0;
  ProfScope ps(a.skip_flag ? "dense_first_layer_skippable" : "wgrad_tc3_kernel", a.ca, a.cb, a.kh, a.dila, a.Wq, 2, a.skip_flag ? 0.0 : 2.0 * npq * a.kh * a.kw * a.ca * a.cb, a.skip_flag ? 0.0 : wbytes, st);
  MSAU_CUDA_TRY(launch_pdl(wgrad_tc3_kernel, grid, dim3(WG3_THREADS), smem, st, a, t));
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// tile geometry of the generic kernel; false when it does not fit shared memory / TMEM (large dilation on a wide dY)

// =====================================================================================================
// Warp-specialised kernel for the <= 128^2 levels (more than 16 output channels; dilation 1 or a single tap).
// What the per-plane kernels above cost there, measured (profiles/README.md, round 2): every 8-channel plane of X had its own
// CTAs, each re-reading ALL of dY through L2 (c32 -> c32: 640 B per pixel instead of 256, ~5.9 TB/s of L2 -> SM traffic for 30
// of the 68 us); the warps that issue the MMAs were also loaders, so the 14 us of instruction issue added to the load time
// instead of hiding under it; and 296 CTAs x 2304 scalar atomics + 2048 same-address shared atomics per CTA made a 12 us tail.
// Here one CTA per SM (576 threads) owns `ppc` planes of X at once:
//   16 loader warps   fp32 global -> bf16 operand images ([plane][row][col][8 ch] for X, [row][plane][col][8 ch] with kh-1 halo
//                     rows for dY, as in wgrad_tc2), S stages, one (plane, row) or (row, plane) item per warp, four items in flight
//    2 MMA warps      one elected lane each; warp w owns the accumulators w, w + 2: (X plane) or, for few planes, (row-interleaved copy)
//   dil = 1, k x k:   M = 64 rows = (kx, ci) of ONE plane, N = kh * cout columns = (ky, co): one instruction per 16 pixels and plane
//   1 x 1 ("gemm"):   M = 64 / 128 rows = (plane, ci) of ALL the CTA's planes, N = cout: one instruction per 16 pixels
//   epilogue          TMEM -> shared memory in dW order -> 16-byte vector reductions (red.global.add.v4.f32) into dW; the bias
//                     gradient is reduced by warp shuffles before it touches shared memory, its planes spread over the CTA groups
struct Wg4Tile {
  int TR, TC, HWx, TRy, nyp, N, M;           // X rows / columns per tile, X row pitch (px), staged dY rows, dY planes, MMA N and M
  int ppc, n_groups, gemm, S, copies;        // X planes per CTA, plane groups (grid.y), 1x1 mode, stages, accumulator copies per item
  int pmode, TRx;                            // per-ky accumulators (dilated / wide layers): X tile has the (kh-1)*dil halo rows, dY none
  int tiles_x, tiles_y, n_tiles, tiles_per_cta;
  uint32_t x_plane_bytes, x_bytes, stage_bytes, tmem_cols;
  int vec;                                   // dW layout / alignment allow 16-byte reductions (checked on the host)
  // loaders: lanes per pixel (log2) / pixels per load instruction / items per tile row, for X and dY; dY plane groups (s2d: phases)
  int lg_lpp_x, pxi_x, spr_x, lg_lpp_y, pxi_y, spr_y, ppg, ngrp, lg_ngrp;
  uint32_t y_pitch;                          // bytes between the channel planes of a staged dY row (TC pixels + bank padding)
  uint32_t m_sx, m_sy;                       // ceil(2^20 / d) for d = spr_x, spr_y: j / d = (j * m) >> 20 for the item indices (j < 4096)
  int dbg;
};
static constexpr int WG4_LOAD_WARPS = 16;
static constexpr int WG4_U = 2;                 // items whose loads are in flight per warp
static constexpr int WG4_K = 4;                 // load instructions (16 B per lane) per item
static constexpr int WG4_MMA_WARPS = 2;         // (576 threads: 112 registers each)
static constexpr int WG4_THREADS = (WG4_LOAD_WARPS + WG4_MMA_WARPS) * 32;
static constexpr int WG4_MAX_S = 4;

__device__ __forceinline__ void wred_v4(float* p, const float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(WG4_THREADS, 1) wgrad_tc4_kernel(const WgradArgs a, const Wg4Tile t) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[WG4_MAX_S];   // loaders -> MMA : operand images of the stage are written
  __shared__ uint64_t bar_free[WG4_MAX_S];   // MMA -> loaders : the instructions reading the stage have retired
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int group = blockIdx.y;
  const int pl0 = group * t.ppc;                                 // first X plane of this CTA
  const int np = min(t.ppc, (a.ca >> 3) - pl0);                  // its X planes
  const int tile0 = blockIdx.x * t.tiles_per_cta;
  const int tile1 = min(t.n_tiles, tile0 + t.tiles_per_cta);
  const int n_my = tile1 - tile0;
  const int items = t.gemm ? 1 : (t.pmode ? np * a.kh : np);     // accumulators = items x copies; copy c takes the rows r = c (mod copies)
  const int n_acc = items * t.copies;
  const int n_iss = min(WG4_MMA_WARPS, n_acc);
  if (a.skip_flag) {                               // (the flag is written by an earlier kernel of this step: wait for it first)
    pdl_wait();
    if (*a.skip_flag == 0) return;
  }

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wsmem_u32(&tmem_base_s)), "r"(t.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int i = 0; i < WG4_MAX_S; ++i) { wmbar_init(&bar_full[i], WG4_LOAD_WARPS); wmbar_init(&bar_free[i], n_iss); }
    wmbar_init(&bar_done, n_iss);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid >= 64 && tid < 192) sbias[tid - 64] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();        // PDL protocol (common.cuh): nothing above reads or writes activations / gradients
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_s;
  if (n_my <= 0) {
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(t.tmem_cols) : "memory");
    return;
  }

  if (warp < WG4_LOAD_WARPS) {
    // =============================================================== loaders
    // Every warp-level load instruction reads whole pixels: the channels of ALL the CTA's X planes (or of all the dY planes of
    // one memory-contiguous group) are consecutive in an NHWC pixel, so lane = (pixel in the instruction, 16-byte chunk of the
    // pixel) touches 32 / lanes-per-pixel cache lines per instruction instead of 32 (one lane per pixel: measured L1-wavefront
    // bound, 128 wavefronts per 2 KB).  An item = WG4_K such instructions = 2 KB: a run of pixels of one tile row.  Warp w takes
    // the items w, w + 16, ...; the loads of WG4_U items are in flight before the first conversion (deeper measured slower).
    // Each lane converts its 4 floats to 4 bf16 = one 8-byte store into the [8 ch] slot of its (plane, pixel).
    const int n_xi = t.TRx * t.spr_x;                            // X items: (row, run of pixels), all planes of the CTA
    const int n_items = n_xi + t.TRy * t.spr_y * t.ngrp;         // dY items: (staged row, run of pixels, plane group)
    const int smul = a.b_s2d ? 2 : 1;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int pxl_x = lane >> t.lg_lpp_x, ch_x = lane & ((1 << t.lg_lpp_x) - 1);
    const int pxl_y = lane >> t.lg_lpp_y, ch_y = lane & ((1 << t.lg_lpp_y) - 1);
    const bool act_x = ch_x < np * 2, act_y = ch_y < t.ppg * 2;
    const uint32_t lane_dx = (uint32_t)(ch_x >> 1) * t.x_plane_bytes + (uint32_t)(ch_x & 1) * 8u;
    const uint32_t lane_dy = (uint32_t)(ch_y >> 1) * t.y_pitch + (uint32_t)(ch_y & 1) * 8u;
    // dY items of this warp always belong to the same plane group (16 % ngrp == 0): a lane's 4 bias-gradient channels stay in
    // registers for the whole kernel
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
    int bacc_pl = -1;
    struct Item { const float* src; uint32_t dst; int step, px0, lim; bool row_ok, is_x, bias; };
    int tx = tile0 % t.tiles_x, ty = (tile0 / t.tiles_x) % t.tiles_y, b = (tile0 / t.tiles_x) / t.tiles_y;
    for (int it = 0; it < n_my; ++it) {
      const int s = it % t.S;
      if (it >= t.S) {
        if (lane == 0) wmbar_wait_spin(&bar_free[s], ((it / t.S) - 1) & 1);
        __syncwarp();
      }
      const uint32_t st = wsmem_u32(smem + (size_t)s * t.stage_bytes);
      const uint32_t yh = st + t.x_bytes;
      const int qy0 = ty * t.TR, qx0 = tx * t.TC;
      const int in_x0 = qx0 - a.pada_l;
      const int in_y0 = t.pmode ? qy0 - a.pada_t : qy0;           // the vertical taps live in the dY rows unless pmode
      const int vy0 = t.pmode ? qy0 : qy0 + a.pada_t - (a.kh - 1);
      const float* const ximg = a.A + (long)b * a.Ha * a.Wa * a.pa + (pl0 << 3);
      const float* const yimg = a.Bm + (long)b * a.Hb * a.Wb * a.pb;
      // px0 = tile column of this lane's first pixel; pixel k of the item is px0 + k * pxi (global: + k * step floats); a pixel is
      // loaded when px < lim (inside the tile and the image) and stored (zero-filled) when it is inside the tile
      auto decode = [&](int j) {
        Item I;
        I.bias = false;
        if (j < n_xi) {
          const int r = (int)(((uint32_t)j * t.m_sx) >> 20), seg = j - r * t.spr_x;
          const int gy = in_y0 + r;
          I.is_x = true;
          I.px0 = seg * WG4_K * t.pxi_x + pxl_x;
          I.row_ok = act_x && (unsigned)gy < (unsigned)a.Ha;
          I.lim = min(t.HWx, a.Wa - in_x0);                      // (the left image border: gx >= 0, checked per pixel)
          I.src = ximg + ((long)gy * a.Wa + in_x0 + I.px0) * a.pa + ch_x * 4;
          I.step = t.pxi_x * a.pa;
          I.dst = st + lane_dx + (uint32_t)(r * t.HWx + I.px0) * 16u;
        } else {
          const int jj = j - n_xi;
          const int g = jj & (t.ngrp - 1), rest = jj >> t.lg_ngrp;
          const int r = (int)(((uint32_t)rest * t.m_sy) >> 20), seg = rest - r * t.spr_y;
          // b_s2d: the planes of group g are one phase (py, px) of the physical tensor, channels [choff, choff + 8 ppg)
          const int col0 = a.b_col0 + g * t.ppg * 8;
          const int sph = a.b_s2d ? col0 / a.cph : 0;
          const int choff = a.b_s2d ? col0 - sph * a.cph : g * t.ppg * 8;
          const int spy = sph >> 1, spx = sph & 1;
          const int vy = vy0 + r, gy = vy * smul + spy;
          I.is_x = false;
          I.px0 = seg * WG4_K * t.pxi_y + pxl_y;
          I.row_ok = act_y && (unsigned)vy < (unsigned)a.Hq && gy < a.Hb;
          // vx = qx0 + px < Wq and gx = vx * smul + spx < Wb
          I.lim = min(t.TC, min(a.Wq - qx0, (a.Wb - spx + smul - 1) / smul - qx0));
          I.src = yimg + ((long)gy * a.Wb + (qx0 + I.px0) * smul + spx) * a.pb + choff + ch_y * 4;
          I.step = t.pxi_y * smul * a.pb;
          const int pl = g * t.ppg + (ch_y >> 1);
          I.dst = yh + lane_dy + (uint32_t)(r * t.nyp + g * t.ppg) * t.y_pitch + (uint32_t)I.px0 * 16u;
          // bias gradient: rows of this tile only (halo rows belong to the neighbours); the planes are spread over the groups
          I.bias = a.dbias != nullptr && act_y && (pl % t.n_groups) == group && vy >= qy0 && vy < qy0 + t.TR;
          if (I.bias) bacc_pl = pl;
        }
        return I;
      };
      const bool skip_ld = (t.dbg & 1) != 0;
      for (int j0 = warp; j0 < n_items && !skip_ld; j0 += WG4_U * WG4_LOAD_WARPS) {
        float4 q[WG4_U][WG4_K];
#pragma unroll
        for (int u = 0; u < WG4_U; ++u) {
          const int j = j0 + u * WG4_LOAD_WARPS;
          if (j < n_items) {
            const Item I = decode(j);
            const int pxi = I.is_x ? t.pxi_x : t.pxi_y;
            const int gx0 = I.is_x ? in_x0 + I.px0 : 0;         // X: global column of pixel 0 (may be negative: left padding)
#pragma unroll
            for (int k = 0; k < WG4_K; ++k) {
              const int px = I.px0 + k * pxi;
              const bool ok = I.row_ok && px < I.lim && gx0 + k * pxi >= 0;
              q[u][k] = ok ? __ldg(reinterpret_cast<const float4*>(I.src + (long)k * I.step)) : z4;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < WG4_U; ++u) {
          const int j = j0 + u * WG4_LOAD_WARPS;
          if (j < n_items) {
            const Item I = decode(j);
            const int pxi = I.is_x ? t.pxi_x : t.pxi_y;
            const int ext = I.is_x ? t.HWx : t.TC;
            const bool act = I.is_x ? act_x : act_y;
            const bool relu = I.is_x && a.reluA;
#pragma unroll
            for (int k = 0; k < WG4_K; ++k) {
              float4 v = q[u][k];
              if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              if (I.bias) { bacc[0] += v.x; bacc[1] += v.y; bacc[2] += v.z; bacc[3] += v.w; }
              if (act && I.px0 + k * pxi < ext) {
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
                const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(I.dst + (uint32_t)(k * pxi) * 16u), "r"(u0), "r"(u1) : "memory");
              }
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) wmbar_arrive(&bar_full[s]);
      if (++tx == t.tiles_x) { tx = 0; if (++ty == t.tiles_y) { ty = 0; ++b; } }
    }
    if (a.dbias != nullptr) {
      // lanes with the same chunk (= channel quadruple) hold partial sums of different pixels
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float sum = bacc[c];
        for (int o = 16; o >= (1 << t.lg_lpp_y); o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (pxl_y == 0 && bacc_pl >= 0) atomicAdd(&sbias[bacc_pl * 8 + (ch_y & 1) * 4 + c], sum);
      }
    }
  } else if (warp - WG4_LOAD_WARPS < n_iss) {
    // =============================================================== MMA issue
    const int w = warp - WG4_LOAD_WARPS;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(t.N >> 3) << 17) |
                           ((uint32_t)(t.M >> 4) << 24);
    const int chunks = t.TC >> 4;
    const uint32_t lbo = (128u >> 4) << 16;
    // row groups of the X operand: the kx taps (one pixel apart) or, 1x1, the channel planes; column groups of dY: (ky, plane)
    const uint32_t a_hi = ((t.gemm ? (t.x_plane_bytes >> 4) : (uint32_t)a.dila) & 0x3FFF) | (1u << 14);
    const uint32_t b_hi = ((t.y_pitch >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t yrow16 = (uint32_t)t.nyp * (t.y_pitch >> 4);   // staged dY row pitch, 16-B units
    for (int it = 0; it < n_my; ++it) {
      const int s = it % t.S;
      if (lane == 0) wmbar_wait_spin(&bar_full[s], (it / t.S) & 1);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (welect_one()) {
        const uint32_t xh = wsmem_u32(smem + (size_t)s * t.stage_bytes);
        const uint32_t yh = xh + t.x_bytes;
        for (int acc = w; acc < n_acc; acc += WG4_MMA_WARPS) {
          const int item = acc % items, copy = acc / items;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * t.N);
          // item = X plane, or (pmode) plane * kh + ky: the same plane image, ky * dil rows further down
          const int ipl = t.pmode ? item / a.kh : item, iky = t.pmode ? item - ipl * a.kh : 0;
          const uint32_t a16 = (xh >> 4) + (t.gemm ? 0u : (uint32_t)ipl * (t.x_plane_bytes >> 4) + (uint32_t)(iky * a.dila * t.HWx));
          uint32_t first = (it == 0) ? 0u : 1u;
          for (int r = copy; r < ((t.dbg & 2) ? 0 : t.TR); r += t.copies) {
            uint32_t a_lo = ((a16 + (uint32_t)(r * t.HWx)) & 0x3FFF) | lbo;
            uint32_t b_lo = (((yh >> 4) + (uint32_t)r * yrow16) & 0x3FFF) | lbo;
            for (int cc = 0; cc < chunks; ++cc) {
              wtc_mma2(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first);
              first = 1u;
              a_lo += 16; b_lo += 16;
            }
          }
        }
        wtc_commit(&bar_free[s]);
        if (it == n_my - 1) wtc_commit(&bar_done);
      }
      __syncwarp();
    }
  }

  // =============================================================== accumulators -> dW
  if (lane == 0) wmbar_wait_spin(&bar_done, 0);
  __syncwarp();
  __syncthreads();                                               // (the operand stages are dead: their memory becomes the dW block)
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float* const blk = reinterpret_cast<float*>(smem);             // [co][ci of this CTA][ky][kx] = dW order
  const int KK = a.kh * a.kw;
  const int seg = np * 8 * KK;
  {
    // row m of an M = 64 accumulator lives in TMEM lane (m % 16) + 32 * (m / 16), of an M = 128 one in lane m; the warps
    // w, w + 4, ... share the lane quarter w & 3 and split the (item, 8-column chunk) pairs
    const int qd = warp & 3, slot = warp >> 2;
    const int n_slots = (WG4_THREADS / 32 - qd + 3) >> 2;        // warps of this lane quarter
    const int m = t.M == 64 ? 16 * qd + lane : 32 * qd + lane;
    const int g = m >> 3, ci = m & 7;                            // row group: kx tap, or (1x1) channel plane
    const bool mine = (t.M == 64 ? lane < 16 : true) && (t.gemm ? g < np : g < a.kw);
    const int n_chunks = t.N >> 3;
    const int pairs = items * n_chunks;
    for (int k = slot; k < pairs && !(t.dbg & 8); k += n_slots) {
      const int item = k / n_chunks, c0 = (k - item * n_chunks) << 3;
      float v[8];
      wtmem_ld8(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(item * t.N + c0), v);
      for (int cp = 1; cp < t.copies; ++cp) {
        float w8[8];
        wtmem_ld8(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((cp * items + item) * t.N + c0), w8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += w8[j];
      }
      if (mine) {
        const int kyi = c0 / a.cb, cb0 = c0 - kyi * a.cb;        // column group = (kyi, 8-channel plane); ky = kh-1-kyi
        const int ipl = t.pmode ? item / a.kh : item;
        const int ky = t.pmode ? item - ipl * a.kh : a.kh - 1 - kyi;
        const int row = t.gemm ? (g * 8 + ci) : ((ipl * 8 + ci) * KK + ky * a.kw + g);
#pragma unroll
        for (int j = 0; j < 8; ++j) blk[(cb0 + j) * seg + row] = v[j];
      }
    }
  }
  __syncthreads();
  {
    const int total = a.cb * seg;
    const int ci0 = pl0 << 3;
    if (t.vec && ci0 + np * 8 <= a.ca_lim && !(t.dbg & 8)) {
      float* const base = a.dW + (long)ci0 * a.s_ca;
      for (int e = tid * 4; e < total; e += WG4_THREADS * 4) {
        const int co = e / seg, i = e - co * seg;
        wred_v4(base + (long)co * a.s_cb + i, *reinterpret_cast<const float4*>(blk + e));
      }
    } else if (!(t.dbg & 8)) {
      for (int e = tid; e < total; e += WG4_THREADS) {
        const int co = e / seg, rem = e - co * seg;
        const int cil = rem / KK, tap = rem - cil * KK;
        const int ci = ci0 + cil;
        if (ci >= a.ca_lim) continue;
        const float v = blk[e];
        if (a.b_s2d) {
          // transposed conv: virtual tap (ty, tx) x phase (py, px) -> real tap: (0,0)->1, (0,1)->2, (1,1)->0, (1,0)->none
          const int ky = tap / a.kw, kx = tap - ky * a.kw;
          const int vco = a.b_col0 + co, ph = vco / a.cph, c = vco - ph * a.cph;
          const int py = ph >> 1, px = ph & 1;
          const int rky = ky == 0 ? (py ? 2 : 1) : (py ? 0 : -1);
          const int rkx = kx == 0 ? (px ? 2 : 1) : (px ? 0 : -1);
          if (ph < 4 && c < a.cb_lim && rky >= 0 && rkx >= 0) atomicAdd(a.dW + (long)ci * a.s_ca + (long)c * a.s_cb + (rky * 3 + rkx), v);
        } else if (co < a.cb_lim) atomicAdd(a.dW + (long)ci * a.s_ca + (long)co * a.s_cb + tap, v);
      }
    }
  }
  if (a.dbias != nullptr && tid < a.cb) {
    const float v = sbias[tid];
    if (v != 0.f) {
      if (a.b_s2d) {
        if (((a.b_col0 + tid) % a.cph) < a.cb_lim) atomicAdd(a.dbias + ((a.b_col0 + tid) % a.cph), v);
      } else if (tid < a.cb_lim) atomicAdd(a.dbias + tid, v);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(t.tmem_cols) : "memory");
}

static bool wgrad_tc4_config(const WgradArgs& a, Wg4Tile& t, size_t& smem, int& ctas) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("MSAU_WG4_OFF"); off = e ? atoi(e) : 0; }
  if (off || a.no_tc4) return false;
  const bool one = a.kh == 1 && a.kw == 1;
  if (a.a_nchw || a.maskB || a.cb > 128 || (a.cb & 7) || (a.ca & 7)) return false;
  if (one && (a.pada_t != 0 || a.pada_l != 0)) return false;
  // per-ky accumulators when the taps cannot ride in the dY column groups: dilation, or kh * cout beyond one instruction
  const bool pmode = !one && (a.dila != 1 || a.kh * a.cb > 256);
  // measured (profiles/README.md): with one X plane (8 input channels, the 512^2 level) the cp.async pipeline of wgrad_tc3 is faster
  // (82 vs 110 us); from two planes on this kernel wins because every operand byte is loaded once (16 -> 16 at 256^2: 50 vs 60 us)
  static int min_ca = -1;
  if (min_ca < 0) { const char* e = getenv("MSAU_WG4_MIN_CA"); min_ca = e ? atoi(e) : 16; }
  if (!pmode && a.ca < min_ca) return false;
  if (pmode && a.b_s2d) return false;
  memset(&t, 0, sizeof(t));
  const int planes = a.ca >> 3;
  t.gemm = one;
  t.pmode = pmode;
  t.nyp = a.cb >> 3;
  t.N = pmode ? a.cb : a.kh * a.cb;
  if (t.N > 256) return false;
  int ppc_max;
  if (one) { t.M = planes > 8 ? 128 : 64; ppc_max = t.M >> 3; }
  else { t.M = 64; ppc_max = 512 / (a.kh * a.cb); if (ppc_max > 16) ppc_max = 16; }
  if (ppc_max < 1) return false;
  t.n_groups = cdiv(planes, ppc_max);
  t.ppc = cdiv(planes, t.n_groups);
  const int items = one ? 1 : (pmode ? t.ppc * a.kh : t.ppc);
  t.copies = WG4_MMA_WARPS / items;
  if (t.copies > 512 / (items * t.N)) t.copies = 512 / (items * t.N);
  if (t.copies < 1) t.copies = 1;
  t.TC = round_up(a.Wq, 16);
  if (t.TC > 64) t.TC = 64;
  t.HWx = t.TC + (a.kw - 1) * a.dila;
  t.tiles_x = cdiv(a.Wq, t.TC);
  const int xhalo = pmode ? (a.kh - 1) * a.dila : 0, yhalo = pmode ? 0 : a.kh - 1;     // halo rows of the X / dY tile
  // loader geometry: lanes per pixel = the 16-byte chunks of the pixel's channel run, rounded up to a power of two
  auto lg_pow2_at_least = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  t.lg_lpp_x = lg_pow2_at_least(t.ppc * 2);
  if (a.b_s2d) {
    // the window [b_col0, b_col0 + cb) of the phase-major virtual columns is either inside one phase or made of whole phases
    if (a.cph >= a.cb) {
      if ((a.b_col0 % a.cph) + a.cb > a.cph) return false;
      t.ngrp = 1; t.ppg = t.nyp;
    } else {
      if ((a.b_col0 % a.cph) || (a.cb % a.cph) || (a.cph & 7)) return false;
      t.ngrp = a.cb / a.cph; t.ppg = a.cph >> 3;
    }
  } else { t.ngrp = 1; t.ppg = t.nyp; }
  if (t.ngrp & (t.ngrp - 1)) return false;
  if (WG4_LOAD_WARPS % t.ngrp) return false;
  t.lg_ngrp = lg_pow2_at_least(t.ngrp);
  t.lg_lpp_y = lg_pow2_at_least(t.ppg * 2);
  if (t.lg_lpp_x > 5 || t.lg_lpp_y > 5) return false;
  t.pxi_x = 32 >> t.lg_lpp_x;
  t.pxi_y = 32 >> t.lg_lpp_y;
  t.spr_x = cdiv(cdiv(t.HWx, t.pxi_x), WG4_K);
  t.spr_y = cdiv(cdiv(t.TC, t.pxi_y), WG4_K);
  // bank padding: the planes' pxi-pixel runs of one store instruction tile the 128-byte bank window
  const uint32_t pad_x = (uint32_t)(t.pxi_x * 16) % 128u, pad_y = (uint32_t)(t.pxi_y * 16) % 128u;
  t.y_pitch = (uint32_t)t.TC * 16u + pad_y;
  int cx = sm_count() / t.n_groups;
  if (cx < 1) cx = 1;
  const size_t budget = 200 * 1024;
  const int KK = a.kh * a.kw;
  const size_t blk_bytes = (size_t)a.cb * t.ppc * 8 * KK * 4;
  // tile rows: the cheapest schedule per CTA.  A tile costs its items (2 KB of loads + conversion each) and a hand-shake; the kh-1
  // halo rows of the staged dY tile favour tall tiles, the balance over the CTAs short ones (constants in microseconds)
  auto x_plane = [&](int tr) {      // (row groups kw..7 of the instruction read up to 7 * dil + 16 pixels past the useful ones)
    return (one ? (uint32_t)(tr * t.TC * 16) : (uint32_t)(((tr + xhalo) * t.HWx + 7 * a.dila + 16) * 16 + 127) / 128 * 128) + pad_x;
  };
  double best = 1e300;
  int best_tr = 0, best_s = 0;
  for (int tr = 1; tr <= 16 && tr <= a.Hq; ++tr) {
    const uint32_t xpb = x_plane(tr);
    const size_t yb = (size_t)(tr + yhalo) * t.nyp * t.y_pitch;
    const size_t stage = ((size_t)t.ppc * xpb + yb + 1023) / 1024 * 1024;
    const size_t tail = one ? (size_t)(t.M >> 3) * xpb : 0;      // 1x1: the row groups past the CTA's planes are read (and ignored)
    int S = (int)((budget - tail) / stage);
    if (budget < tail || S < 2) continue;
    if (S > WG4_MAX_S) S = WG4_MAX_S;
    if ((size_t)S * stage + tail < blk_bytes) continue;
    const long n_tiles = (long)t.tiles_x * cdiv(a.Hq, tr) * a.B;
    const long tpc = (n_tiles + cx - 1) / cx;
    const int n_items = (tr + xhalo) * t.spr_x + (tr + yhalo) * t.spr_y * t.ngrp;
    double cost = (double)tpc * (0.05 * n_items + 0.3) * (S >= 3 ? 1.0 : 1.06);
    if (t.copies > tr) cost *= 4.0;
    if (cost < best) { best = cost; best_tr = tr; best_s = S; }
  }
  if (!best_tr) return false;
  t.TR = best_tr; t.S = best_s;
  if (t.copies > t.TR) t.copies = t.TR;
  t.TRy = t.TR + yhalo;
  t.TRx = t.TR + xhalo;
  t.x_plane_bytes = x_plane(t.TR);
  t.x_bytes = (uint32_t)t.ppc * t.x_plane_bytes;
  t.stage_bytes = (uint32_t)((t.x_bytes + (size_t)t.TRy * t.nyp * t.y_pitch + 1023) / 1024 * 1024);
  smem = (size_t)t.S * t.stage_bytes + (one ? (size_t)(t.M >> 3) * t.x_plane_bytes : 0) + 1024;
  const int cols = items * t.copies * t.N;
  t.tmem_cols = 32;
  while ((int)t.tmem_cols < cols) t.tmem_cols <<= 1;
  if (t.tmem_cols > 512) return false;
  t.tiles_y = cdiv(a.Hq, t.TR);
  t.n_tiles = t.tiles_x * t.tiles_y * a.B;
  ctas = cx < t.n_tiles ? cx : t.n_tiles;
  t.tiles_per_cta = cdiv(t.n_tiles, ctas);
  ctas = cdiv(t.n_tiles, t.tiles_per_cta);
  t.m_sx = ((1u << 20) + t.spr_x - 1) / t.spr_x;
  t.m_sy = ((1u << 20) + t.spr_y - 1) / t.spr_y;
  t.vec = !a.b_s2d && a.s_ca == KK && (a.s_cb & 3) == 0 && (((uintptr_t)a.dW) & 15) == 0 && a.cb_lim >= a.cb;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MSAU_WG_DBG"); dbg = e ? atoi(e) : 0; } t.dbg = dbg; }
  return true;
}

static int launch_wgrad_tc4(const WgradArgs& a, const Wg4Tile& t, size_t smem, int ctas, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    MSAU_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    attr = true;
  }
  dim3 grid(ctas, t.n_groups);
  const double npq = (double)a.B * a.Hq * a.Wq;
  const double wbytes = (npq * a.ca + (double)a.B * a.Hb * a.Wb * (a.b_s2d ? a.cph : a.cb)) * 4.0;
  ProfScope ps(a.skip_flag ? "dense_first_layer_skippable" : "wgrad_tc4_kernel", a.ca, a.cb, a.kh, a.dila, a.Wq, a.a_nchw,
               a.skip_flag ? 0.0 : 2.0 * npq * a.kh * a.kw * a.ca * a.cb, a.skip_flag ? 0.0 : wbytes, st);
  MSAU_CUDA_TRY(launch_pdl(wgrad_tc4_kernel, grid, dim3(WG4_THREADS), smem, st, a, t));
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

static bool wg_config(const WgradArgs& a, WgTile& t, size_t& smem, int& ctas_out) {
  t.N = a.cb;                           // M = 64 allows any multiple of 8
  t.ny_planes = a.cb >> 3;
  int px_budget = 16384 / a.cb;         // dY tile <= 32 KB of bf16
  if (px_budget > 1024) px_budget = 1024;
  t.TC = round_up(a.Wq, 16);
  if (t.TC > 64) t.TC = 64;
  while (t.TC > 16 && t.TC * 2 > px_budget * 2 && t.TC > px_budget) t.TC -= 16;
  t.TR = px_budget / t.TC;
  if (t.TR < 1) t.TR = 1;
  if (t.TR > 16) t.TR = 16;
  if (t.TR > a.Hq) t.TR = a.Hq;
  t.HHx = t.TR + (a.kh - 1) * a.dila;
  t.HWx = t.TC + (a.kw - 1) * a.dila;
  const int slack_px = 7 * a.dila + 16;                   // M groups kw..7 read past the useful columns
  t.x_plane_bytes = (uint32_t)((t.HHx * t.HWx + slack_px) * 16 + 127) / 128 * 128;
  t.y_plane_bytes = (uint32_t)(t.TR * t.TC * 16);
  const uint32_t y_bytes = (uint32_t)t.ny_planes * t.y_plane_bytes;
  t.stage_bytes = (t.x_plane_bytes + y_bytes + 1023) / 1024 * 1024;
  t.xq = WG_THREADS / t.HWx; t.xr = WG_THREADS % t.HWx;
  { const int dpx = WG_THREADS / t.ny_planes; t.yq = dpx / t.TC; t.yr = dpx % t.TC; }
  t.tiles_x = cdiv(a.Wq, t.TC);
  t.tiles_y = cdiv(a.Hq, t.TR);
  t.n_tiles = t.tiles_x * t.tiles_y * a.B;
  const int planes = a.ca >> 3;
  // CTAs per SM (x the channel planes): the tiles of a CTA run back to back with one tile of look-ahead, so several resident CTAs
  // are what hides the load -> convert -> issue latency on the small maps this kernel serves (env MSAU_WG_PER_SM for experiments)
  static int per_sm_env = -1;
  if (per_sm_env < 0) { const char* e = getenv("MSAU_WG_PER_SM"); per_sm_env = e ? atoi(e) : 0; }
  const int per_sm = per_sm_env > 0 ? per_sm_env : 2;
  int ctas = (per_sm * sm_count() + planes - 1) / planes;
  if (ctas > t.n_tiles) ctas = t.n_tiles;
  if (ctas < 1) ctas = 1;
  t.tiles_per_cta = cdiv(t.n_tiles, ctas);
  ctas = cdiv(t.n_tiles, t.tiles_per_cta);
  int cols = a.kh * t.N;
  t.tmem_cols = 32;
  while ((int)t.tmem_cols < cols) t.tmem_cols <<= 1;
  smem = (size_t)t.stage_bytes * 2 + 1024;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MSAU_WG_DBG"); dbg = e ? atoi(e) : 0; } t.dbg = dbg; }
  ctas_out = ctas;
  return smem <= 200 * 1024 && t.tmem_cols <= 512;
}

bool wgrad_tc_supported(const WgradArgs& a) {
  if (a.sa != 1 || a.sb != 1 || a.dilb != 0 || a.padb_t != 0 || a.padb_l != 0) return false;
  if (a.Ha != a.Hq || a.Wa != a.Wq) return false;
  if (a.b_s2d) {
    // (b_col0: this launch covers the virtual columns [b_col0, b_col0 + cb) of the 4 * cph phase-major columns)
    if (a.b_col0 < 0 || (a.b_col0 & 7) || a.b_col0 + a.cb > 4 * a.cph || (a.cph & 7) || a.maskB || a.kh != 2 || a.kw != 2) return false;
    if (a.Hb > 2 * a.Hq || a.Hb < 2 * a.Hq - 1 || a.Wb > 2 * a.Wq || a.Wb < 2 * a.Wq - 1) return false;
  } else if (a.Hb != a.Hq || a.Wb != a.Wq) return false;
  if ((a.ca & 7) || (a.cb & 7) || a.cb > 128 || a.kw > 4 || a.kh > 4) return false;
  const int nyp = a.cb >> 3;
  if (256 % nyp) return false;
  if (!a.a_nchw && (a.pa & 3)) return false;
  if (a.pb & 3) return false;
  if (a.dila < 1) return false;
  {
    Wg3Tile t3;
    if (wgrad_tc3_config(a, t3)) return true;
    Wg2Tile t2;
    if (wgrad_tc2_config(a, t2)) return true;
  }
  WgTile t;
  size_t smem;
  int ctas;
  return wg_config(a, t, smem, ctas);
}

int launch_wgrad_tc(const WgradArgs& a, cudaStream_t st) {
  MSAU_CHECK_ARG(wgrad_tc_supported(a), "wgrad_tc: unsupported shape");
  {
    Wg4Tile t4;
    size_t smem4 = 0;
    int ctas4 = 0;
    if (wgrad_tc4_config(a, t4, smem4, ctas4)) return launch_wgrad_tc4(a, t4, smem4, ctas4, st);
    Wg3Tile t3;
    if (wgrad_tc3_config(a, t3)) return launch_wgrad_tc3(a, t3, st);
    Wg2Tile t2;
    if (wgrad_tc2_config(a, t2)) return launch_wgrad_tc2(a, t2, st);
  }
  WgTile t;
  size_t smem = 0;
  int ctas = 0;
  const bool fits = wg_config(a, t, smem, ctas);
  const int planes = a.ca >> 3;
  MSAU_CHECK_ARG(fits, "wgrad_tc: tile does not fit (smem %zu B, tmem %u cols)", smem, t.tmem_cols);
  static bool attr = false;
  if (!attr) {
    MSAU_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid(ctas, planes);
  const double npq = (double)a.B * a.Hq * a.Wq;
  const double wbytes = (npq * (a.a_nchw ? a.ca_logical : a.ca) + npq * a.cb * (a.maskB ? 2 : 1)) * 4.0;
  ProfScope ps(a.skip_flag ? "dense_first_layer_skippable" : "wgrad_tc_kernel", a.ca, a.cb, a.kh, a.dila, a.Wq, a.a_nchw, a.skip_flag ? 0.0 : 2.0 * npq * a.kh * a.kw * a.ca * a.cb, a.skip_flag ? 0.0 : wbytes, st);
  MSAU_CUDA_TRY(launch_pdl(wgrad_tc_kernel, grid, dim3(WG_THREADS), smem, st, a, t));
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

}  // namespace msau
