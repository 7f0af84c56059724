// 3x3 (dilation 1, stride 1, SAME) convolution on tcgen05 with the kx taps folded into the N dimension.
//
// The implicit-GEMM kernel of conv_tc.cu issues one instruction per (tap, 8 input channels, term of the bf16 split):
// 14 per 128-pixel tile and channel plane.  With 8..32 output channels those instructions are bound by the 4 KB
// A-operand read from shared memory (~39 cycles each, scripts/mma_bench.cu), not by the math, so the tensor pipe --
// not HBM -- limits the 8/16-channel levels of MSAU.  Here the three kx taps ride along as extra N columns:
//     D[pixel, (kx, cout)] = sum_{ky, cin} X[pixel + ky rows, cin] * W[ky, kx, cin, cout]        (5 instructions)
//     out[pixel x, cout]   = D[x-1, (0, cout)] + D[x, (1, cout)] + D[x+1, (2, cout)]             (epilogue)
// A tile is 4 image rows x 32 columns of the halo image (M = 128: TMEM lane = 32 * row + column, so every warp of
// the epilogue owns one row segment and the horizontal shift is a warp shuffle); columns 1..30 produce outputs.
// The halo image is planar bf16 [hi | lo][row][32 cols][8 ch] with a 512-byte row pitch, so the 128 rows of the
// A operand are one contiguous 2 KB block per K chunk (SBO = 128 B) and a ky tap is a 512-byte start offset.
//
// fp32 accuracy from bf16 tensor cores as in conv_tc.cu: x*w ~= hi*whi + lo*whi + hi*wlo:
//     ky = 0,1,2 : A = [hi(ky) | lo(ky)]        B = [Whi(ky) ; Whi(ky)]
//     pair (0,1) : A = [hi(0)  | hi(1)]         B = [Wlo(0)  ; Wlo(1)]
//     pair (2,-) : A = [hi(2)  | lo(2)]         B = [Wlo(2)  ; 0]
//
// Persistent CTA per SM: 6 producer warps, 2 MMA-issuing warps (every other tile each: one instruction stream cannot
// keep the pipe busy), 8 epilogue warps.  Nothing waits on HBM with registers: the producers keep a D-deep ring of raw
// fp32 halo planes in flight with cp.async (zero-fill = SAME padding) into thread-private shared-memory slots and
// convert the oldest one to the planar bf16 hi/lo image; the epilogue's extra operands (residual, ReLU mask source,
// skip-path add, previous output) are prefetched one super-tile ahead the same way.  The weight images of all input
// planes stay resident in shared memory for the life of the CTA.
//
// Reference semantics: model/layers/layers.py:10-102 (conv), model/model.py:37-50 (residual epilogue).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "prof.cuh"
#include "tc_ptx.cuh"
#include "tma.cuh"

namespace msau {

using namespace ptx;

// MSAU_TC_DEBUG bit 32: per-role wait / work cycle counters (lane 0 of every warp), read back by msau_debug_c3_prof
__device__ unsigned long long g_c3_prof[16];

namespace {

constexpr int C3_PROD_WARPS = 6;
constexpr int C3_PROD_THREADS = C3_PROD_WARPS * 32;
#ifndef MSAU_C3_MMA_WARPS
#define MSAU_C3_MMA_WARPS 2
#endif
constexpr int C3_MMA_WARPS = MSAU_C3_MMA_WARPS;
#ifndef MSAU_C3_EPI_WARPS
#define MSAU_C3_EPI_WARPS 8       // measured: 16 epilogue warps are slower (117 vs 104 us at 8 -> 8 channels / 512^2): the shared-memory pipe, not latency, is the limit
#endif
constexpr int C3_EPI_WARPS = MSAU_C3_EPI_WARPS;      // a multiple of 4: C3_EPI_SUBS warps share each TMEM lane quarter (= tile row)
constexpr int C3_EPI_SUBS = C3_EPI_WARPS / 4;
constexpr int C3_EPI_HALF = C3_EPI_WARPS * 32 * 16;  // bytes of one 16-byte operand half over all epilogue threads
constexpr int C3_EPI_OP = 2 * C3_EPI_HALF;           // ... of one 8-channel operand
constexpr int C3_THREADS = (C3_PROD_WARPS + C3_MMA_WARPS + C3_EPI_WARPS) * 32;
constexpr int C3_MAX_STAGES = 3;
constexpr int C3_SLOTS = (8 + C3_EPI_SUBS - 1) / C3_EPI_SUBS;   // epilogue work items per thread and super-tile (a super-tile has <= 8)

struct C3Tile {
  int KS, pad, OW, NI;                // kernel size (3 or 4), SAME padding before, output columns per block (32 - KS + 1), MMAs per tile-plane
  int T, N, CP, RI, P, stages, D, A;  // tiles per super-tile, MMA N, padded cout, halo rows, input planes, image / raw / accumulator ring depth
  int blocks_x, blocks_y, n_super;
  int dgx, dgy, dgb;                  // gridDim.x decomposed in (block column, block row, page) steps
  int n_ops, ia, io, im;              // extra epilogue operands and their slot index (-1 = absent)
  int dbg;
  int tma;                            // raw halo planes arrive by TMA tensor loads (dense [row][32 px][8 ch] slots) instead of cp.async
  uint32_t plane_bytes, in_bytes, w_bytes, w_total, w_copy, raw_bytes, epi_bytes, tmem_cols;   // w_total: w_copy rounded up to 1 KB (slot alignment)
};

// position of a super-tile, advanced by gridDim.x super-tiles at a time without integer division
struct TilePos {
  int bx, by, b;
  __device__ __forceinline__ void init(int st_i, const C3Tile& t) {
    bx = st_i % t.blocks_x;
    const int rest = st_i / t.blocks_x;
    by = rest % t.blocks_y;
    b = rest / t.blocks_y;
  }
  __device__ __forceinline__ void advance(const C3Tile& t) {
    bx += t.dgx;
    if (bx >= t.blocks_x) { bx -= t.blocks_x; ++by; }
    by += t.dgy;
    if (by >= t.blocks_y) { by -= t.blocks_y; ++b; }
    b += t.dgb;
  }
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// src_bytes = 0 -> the 16 destination bytes are zero-filled (out-of-image halo pixels: SAME padding)
__device__ __forceinline__ void cp_async16z(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// polling wait for the few single lanes that wait on behalf of their warp: no suspend hint, so the wake-up is immediate
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
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

__device__ __forceinline__ void mma2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
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


// EPI: 0 = epilogue driven by run-time flags; 1 = bias; 2 = bias + ReLU; 3 = bias + residual + ReLU; 4 = ReLU mask; 5 = ReLU mask + add
// LDU: halo rows per producer warp and plane (6: T = 8 tiles, the 8-channel levels; 3: T = 4, 16 channels; 2: T <= 2, 32 channels)
// TMA: the raw fp32 halo plane of a chunk is ONE cp.async.bulk.tensor load issued by one thread (map tm1 / tm2 = src1 / src2 as
//      {channel, x, y, page}; out-of-image pixels zero-filled by the hardware = SAME padding), completion on an mbarrier
template <int EPI, int LDU, int KS, int PAD, bool TMA>
__global__ void __launch_bounds__(C3_THREADS, 1) conv3_tc_kernel(const ConvArgs a, const uint16_t* __restrict__ wtc, const C3Tile t,
                                                                 const __grid_constant__ CUtensorMap tm1,
                                                                 const __grid_constant__ CUtensorMap tm2,
                                                                 const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmO,
                                                                 const __grid_constant__ CUtensorMap tmM) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[C3_MAX_STAGES];    // producers -> MMA : stage holds one plane (image + weights)
  __shared__ uint64_t bar_empty[C3_MAX_STAGES];   // MMA -> producers : the MMAs reading the stage have retired
  __shared__ uint64_t bar_acc_full[4];            // MMA -> epilogue  : accumulator set complete
  __shared__ uint64_t bar_acc_empty[4];           // epilogue -> MMA  : accumulator set drained
  __shared__ uint64_t bar_raw[4];                 // TMA -> producers : raw plane of a chunk has landed (transaction bytes)
  __shared__ uint64_t bar_epi[C3_EPI_WARPS][C3_SLOTS];   // TMA -> one epilogue warp : extra operands of its work item k have landed
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(&tmem_base_s, t.tmem_cols);
  if (tid >= 64 && tid < 128) bias_s[tid - 64] = (a.bias && tid - 64 < t.CP) ? a.bias[tid - 64] : 0.f;
  // shared memory: [weight images of all planes][image ring: S x (hi | lo)][raw ring: D x fp32 plane][epilogue slots]
  uint8_t* const w_s = smem;
  uint8_t* const img_s = smem + t.w_total;
  uint8_t* const raw_s = img_s + (size_t)t.stages * t.in_bytes;
  uint8_t* const epi = raw_s + (size_t)t.D * t.raw_bytes;
  for (int e = tid; e < (int)(t.w_copy >> 4); e += C3_THREADS)
    reinterpret_cast<uint4*>(w_s)[e] = __ldg(reinterpret_cast<const uint4*>(wtc) + e);
  fence_async_smem();
  if (tid == 32) {
    for (int i = 0; i < C3_MAX_STAGES; ++i) { mbar_init(&bar_full[i], C3_PROD_WARPS); mbar_init(&bar_empty[i], C3_MMA_WARPS); }
    for (int i = 0; i < 4; ++i) { mbar_init(&bar_acc_full[i], C3_MMA_WARPS); mbar_init(&bar_acc_empty[i], C3_EPI_WARPS); mbar_init(&bar_raw[i], 1); }
    if (TMA)
      for (int i = 0; i < C3_EPI_WARPS * C3_SLOTS; ++i) mbar_init(&bar_epi[0][0] + i, 1);
    mbar_init_fence();
    if (TMA) { tma_prefetch_desc(&tm1); if (a.c2) tma_prefetch_desc(&tm2); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();        // everything above touched only parameters, packed weights, shared memory and TMEM (common.cuh: PDL protocol)
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_s;
  const int S = t.stages;
  const bool spin = !(t.dbg & 16);
#ifdef MSAU_C3_PROF      // role timers: build with -DMSAU_C3_PROF and run with MSAU_TC_DEBUG=32 (scripts/c3_roles.py)
  const bool prof = (t.dbg & 32) != 0;
  unsigned long long pc[4] = {0, 0, 0, 0};
  const long long t_start = prof ? clock64() : 0;
#define C3_TIMED(slot, stmt) { if (prof) { const long long _t0 = clock64(); stmt; pc[slot] += (unsigned long long)(clock64() - _t0); } else { stmt; } }
#else
#define C3_TIMED(slot, stmt) { stmt; }
#endif
  auto wait = [&](uint64_t* bar, uint32_t parity) { if (spin) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity); };
  const int RO = 4 * t.T;                                     // output rows per super-tile

  if (warp < C3_PROD_WARPS) {
    // =============================================================== producers
    // chunk = (super-tile, input plane).  Copies of chunk c + D - 1 are issued before chunk c is converted, so D - 1
    // planes (RI x 32 pixels x 32 B each) are always in flight per CTA; every thread copies and converts its own pixels
    // e = tid + u * 192, which makes the raw ring thread-private (cp.async.wait_group is the only synchronisation).
    const int planes1 = a.c1 >> 3;
    const uint32_t raw_u32 = smem_u32(raw_s);
    // Lean on purpose (the role is paced by its serial instruction chain): ring positions are counters, addresses are
    // 32-bit element offsets (every tensor here has < 2^31 elements) built from per-thread constants.
    const int lpx = lane >> 1, lhalf = (lane & 1) * 4;
    uint32_t d_issue = 0;                                          // raw slot of the next issue (bytes)
    int i_slot = 0;                                                // ... and its index (TMA: barrier of that slot)
    auto issue = [&](const TilePos& tp, int p) {
      if (TMA) {
        // one thread, one instruction: the whole RI x 32 x 8 plane; coordinates may be negative / past the edge (zero-filled)
        if (tid == 0 && tp.b < a.B && !(t.dbg & 2)) {
          const bool from1 = p < planes1;
          fence_async_smem();                                      // the slot was last read through the generic proxy
          mbar_arrive_expect_tx(&bar_raw[i_slot], t.raw_bytes);
          tma_load_4d(raw_u32 + d_issue, from1 ? &tm1 : &tm2, (from1 ? p : p - planes1) << 3, tp.bx * t.OW - t.pad, tp.by * RO - t.pad, tp.b,
                      &bar_raw[i_slot]);
        }
        d_issue += t.raw_bytes;
        if (++i_slot == t.D) { i_slot = 0; d_issue = 0; }
        return;
      }
      if (tp.b < a.B && !(t.dbg & 2)) {
        const int in_x0 = tp.bx * t.OW - t.pad, in_y0 = tp.by * RO - t.pad;
        const bool from1 = p < planes1;
        const float* base = from1 ? a.src1 : a.src2;
        const int pitch = from1 ? a.p1 : a.p2;
        const int ch = ((from1 ? p : p - planes1) << 3) + lhalf;
        // a halo row (32 pixels x 32 B) is 64 chunks of 16 B: lane l copies chunks l and l + 32 of the rows warp + 6 u, i.e.
        // half (l & 1) of the pixels l / 2 and l / 2 + 16 -- each warp instruction moves 512 contiguous bytes and every
        // 32-byte sector is requested exactly once (cp.async.cg goes straight to L2)
        const int gx0 = in_x0 + lpx;
        const bool okx0 = (unsigned)gx0 < (unsigned)a.Win, okx1 = (unsigned)(gx0 + 16) < (unsigned)a.Win;
        int gy = in_y0 + warp;
        int off = ((tp.b * a.Hin + gy) * a.Win + gx0) * pitch + ch;
        const int roff = C3_PROD_WARPS * a.Win * pitch;
        uint32_t dst = raw_u32 + d_issue + (uint32_t)tid * 16u;
#pragma unroll
        for (int u = 0; u < LDU; ++u) {
          if (warp + u * C3_PROD_WARPS < t.RI) {
            const bool oky = (unsigned)gy < (unsigned)a.Hin;
            const float* sp = base + (long)off;
            cp_async16z(dst, (oky && okx0) ? sp : base, (oky && okx0) ? 16u : 0u);
            cp_async16z(dst + C3_PROD_THREADS * 16u, (oky && okx1) ? sp + 16 * pitch : base, (oky && okx1) ? 16u : 0u);
          }
          gy += C3_PROD_WARPS; off += roff; dst += 2u * C3_PROD_THREADS * 16u;
        }
      }
      d_issue += t.raw_bytes;
      if (d_issue == (uint32_t)t.D * t.raw_bytes) d_issue = 0;
      cp_async_commit();
    };
    TilePos ahead;
    ahead.init(blockIdx.x, t);
    int p_ahead = 0;
    // (every thread keeps D - 1 groups outstanding: empty groups stand in for chunks that do not exist)
    for (int d = 0; d < t.D - 1; ++d) {
      issue(ahead, p_ahead);
      if (++p_ahead == t.P) { p_ahead = 0; ahead.advance(t); }
    }
    int s = 0;
    uint32_t sphase = 0, d_cur = 0, c = 0;
    int r_slot = 0;
    uint32_t r_phase = 0;                                                        // TMA: slot / parity of the chunk being converted
    const uint32_t dimg = (uint32_t)(lpx * 16 + (lane & 1) * 8 + warp * 512);   // this thread's first half-pixel in an image
    // raw slot addressing of this thread's half-pixels: cp.async slots are thread-private ([u][half][thread] x 16 B), a TMA slot
    // is the dense plane [row][32 px][8 ch]: row warp + 6 u, pixel lpx + 16 j, half (lane & 1) -> consecutive lanes still read
    // consecutive 16-byte chunks
    const uint32_t r_base = TMA ? (uint32_t)(warp * 1024 + lane * 16) : (uint32_t)(tid * 16);
    const uint32_t r_u = TMA ? (uint32_t)(C3_PROD_WARPS * 1024) : (uint32_t)(2 * C3_PROD_THREADS * 16);
    const uint32_t r_j = TMA ? 512u : (uint32_t)(C3_PROD_THREADS * 16);
    for (int st_i = blockIdx.x; st_i < t.n_super; st_i += gridDim.x) {
      for (int p = 0; p < t.P; ++p, ++c) {
        // TMA: the slot about to be refilled held chunk c - 1, which every producer thread has finished reading only now
        if (TMA) C3_TIMED(3, named_bar_sync(1, C3_PROD_THREADS));
        issue(ahead, p_ahead);
        if (++p_ahead == t.P) { p_ahead = 0; ahead.advance(t); }
        uint8_t* stg = img_s + (size_t)s * t.in_bytes + dimg;
        if (c >= (uint32_t)S) {
          if (lane == 0) C3_TIMED(0, wait(&bar_empty[s], sphase ^ 1u));
          __syncwarp();
        }
        if (TMA) {
          if (!(t.dbg & 2)) C3_TIMED(1, mbar_wait_short(&bar_raw[r_slot], r_phase));   // (every thread acquires the plane itself)
          if (++r_slot == t.D) { r_slot = 0; r_phase ^= 1u; }
        } else {
          // chunk c has landed once at most D - 1 newer groups are pending (D is 2..4)
          if (t.D == 4) cp_async_wait<3>(); else if (t.D == 3) cp_async_wait<2>(); else cp_async_wait<1>();
        }
        const uint8_t* rsrc = raw_s + d_cur + r_base;
        const bool relu = a.relu1 && p < planes1;
        // this thread's two half-pixels of every row: 4 channels each -> 8 bytes of the hi image + 8 bytes of the lo image
#pragma unroll
        for (int u = 0; u < LDU; ++u) {
          if (warp + u * C3_PROD_WARPS < t.RI) {
            float4 q[2];
            q[0] = *reinterpret_cast<const float4*>(rsrc + u * r_u);
            q[1] = *reinterpret_cast<const float4*>(rsrc + u * r_u + r_j);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float4 v = q[j];
              if (t.dbg & 2) v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              uint2 hi, lo;
              split_pair(v.x, v.y, hi.x, lo.x);
              split_pair(v.z, v.w, hi.y, lo.y);
              uint8_t* d = stg + u * (C3_PROD_WARPS * 512) + j * 256;
              *reinterpret_cast<uint2*>(d) = hi;
              *reinterpret_cast<uint2*>(d + t.plane_bytes) = lo;
            }
          }
        }
        // one arrival per warp: 192 per-thread arrivals on one mbarrier serialise in shared memory
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[s]);
        if (++s == S) { s = 0; sphase ^= 1u; }
        d_cur += t.raw_bytes;
        if (d_cur == (uint32_t)t.D * t.raw_bytes) d_cur = 0;
      }
    }
    if (!TMA) cp_async_wait<0>();
  } else if (warp < C3_PROD_WARPS + C3_MMA_WARPS) {
    // =============================================================== MMA issuers (tiles mw, mw + 2, ...)
    const int mw = warp - C3_PROD_WARPS;
    const uint32_t idesc = make_idesc(128, t.N, false, false);
    const uint32_t N = (uint32_t)t.N;
    const uint32_t a_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t lbo_plane = ((t.plane_bytes >> 4) & 0x3FFF) << 16;     // hi image -> lo image
    const uint32_t lbo_row = (512u >> 4) << 16;                           // next image row (next ky)
    const uint32_t b_lbo = ((N * 16 >> 4) & 0x3FFF) << 16;
    uint32_t tcount = 0, as = 0, aphase = 0, sphase = 0;
    int s = 0;
    for (int st_i = blockIdx.x; st_i < t.n_super; st_i += gridDim.x, ++tcount) {
      if (tcount >= (uint32_t)t.A) {
        if (lane == 0) C3_TIMED(1, wait(&bar_acc_empty[as], aphase ^ 1u));
        __syncwarp();
      }
      tc_fence_after();
      const uint32_t acc_base = tmem_base + as * (uint32_t)(t.T * t.N);
      for (int p = 0; p < t.P; ++p) {
        if (lane == 0) C3_TIMED(0, wait(&bar_full[s], sphase));
        __syncwarp();
        tc_fence_after();
        const uint32_t in16 = smem_u32(img_s + (size_t)s * t.in_bytes) >> 4;
        const uint32_t w16a = smem_u32(w_s + (size_t)p * t.w_bytes) >> 4;
        if (elect_one()) {
          if (!(t.dbg & 1)) {
            // instruction-major / tile-minor: consecutive instructions hit different accumulators
#pragma unroll
            for (int i = 0; i < KS + (KS + 1) / 2; ++i) {
              // i < KS: A = [hi(ky) | lo(ky)];  then pairs: A = [hi(2j) | hi(2j+1)] (next image row), the odd one out: [hi | lo] x [Wlo ; 0]
              const uint32_t ky = i < KS ? (uint32_t)i : (uint32_t)(2 * (i - KS));
              const uint32_t lbo = (i >= KS && 2 * (i - KS) + 1 < KS) ? lbo_row : lbo_plane;
              const uint32_t b_lo = ((w16a + (uint32_t)i * 2u * N) & 0x3FFF) | b_lbo;
              const uint32_t acc = (p > 0 || i > 0) ? 1u : 0u;
              for (int tile = mw; tile < t.T; tile += C3_MMA_WARPS) {
                const uint32_t a_lo = ((in16 + ((uint32_t)(4 * tile) + ky) * 32u) & 0x3FFF) | lbo;
                mma2(acc_base + (uint32_t)tile * N, a_lo, a_hi, b_lo, b_hi, idesc, acc);
              }
            }
          }
          tc_commit(&bar_empty[s]);
          if (p == t.P - 1) tc_commit(&bar_acc_full[as]);
        }
        __syncwarp();
        if (++s == S) { s = 0; sphase ^= 1u; }
      }
      if (++as == (uint32_t)t.A) { as = 0; aphase ^= 1u; }
    }
  } else {
    // =============================================================== epilogue
    // Kept lean on purpose: with 8 x 32 threads x 4 items per super-tile this role is bound by instruction issue, so the
    // super-tile position is decoded once (for the tile being drained and for the one being prefetched), item -> (tile,
    // channel chunk) is shifts and masks, and the bias comes from shared memory.
    constexpr bool G = EPI == 0;
    const int ew = warp - (C3_PROD_WARPS + C3_MMA_WARPS);   // 0..7
    const int q = warp & 3;                                  // TMEM lane quarter = tile row
    const int sub = ew >> 2;                                // 0 .. C3_EPI_SUBS-1
    const int chunks = t.CP >> 3;                            // 1, 2, 4 or 8
    const int lc = 31 - __clz(chunks);
    const int total_items = t.T << lc;                       // <= 8: item kg -> (tile kg >> lc, channel chunk kg & (chunks-1))
    const int et = ew * 32 + lane;                           // epilogue thread index 0..255
    const uint32_t epi_u32 = smem_u32(epi);
    const uint32_t slot_stride = (uint32_t)(G ? t.n_ops : ((EPI == 3 || EPI == 4) ? 1 : (EPI == 5 ? 2 : 0))) * (uint32_t)C3_EPI_OP;  // bytes between consecutive items' slots
    const uint32_t my_slot = (uint32_t)et * 16u;
    const bool lane_ok = lane >= t.pad && lane < t.pad + t.OW;
    const int tile_pix = 4 * a.Wout;                         // pixel distance between consecutive tiles (4 rows)
    // position of a super-tile for this thread: pixel index of its tile-0 output, first output row, column validity
    struct Pos { int pix; int row; bool ok; int x0; int b; };   // (32-bit: every tensor here has < 2^31 elements)
    auto decode = [&](const TilePos& tp) -> Pos {
      Pos ps;
      if (tp.b >= a.B) { ps.pix = 0; ps.row = 1 << 29; ps.ok = false; ps.x0 = 0; ps.b = -1; return ps; }
      ps.x0 = tp.bx * t.OW - t.pad;                          // column of lane 0 (warp-uniform; may be -1 / -2: zero-filled by TMA)
      ps.b = tp.b;
      const int ox = ps.x0 + lane;
      ps.row = tp.by * RO + q;
      ps.ok = lane_ok && ox < a.Wout;
      ps.pix = (tp.b * a.Hout + ps.row) * a.Wout + ox;
      return ps;
    };
    TilePos tp;
    tp.init(blockIdx.x, t);
    const bool f_relu = G ? a.relu != 0 : EPI == 2;
    const bool f_res = G ? a.res != nullptr : EPI == 3;
    const bool f_relu2 = G ? a.relu2 != 0 : EPI == 3;
    const bool f_add = G ? a.add != nullptr : EPI == 5;
    const int ia = G ? t.ia : ((EPI == 3 || EPI == 5) ? 0 : -1);
    const int io = G ? t.io : -1;
    const int im = G ? t.im : (EPI == 4 ? 0 : (EPI == 5 ? 1 : -1));
    const int n_ops = G ? t.n_ops : ((EPI == 3 || EPI == 4) ? 1 : (EPI == 5 ? 2 : 0));
    const float* pa = f_res ? a.res : a.add;
    const int ppa = f_res ? a.pr : a.pa;
    // TMA variant: the extras of (this warp's tile row, item k) are ONE 1 KB box per operand -- [32 columns][8 channels] fp32 of
    // the row, written with the 32-byte swizzle (the two 16-byte halves of the pixels 4..7 of every 8 swap places), which makes
    // the per-lane 16-byte reads below bank-conflict free -- issued by lane 0 and completed on this warp's own mbarrier.
    // (The cp.async form costs 8x the ideal shared-memory wavefronts: every lane's 16 bytes arrive as their own sector.)
    const uint32_t wslot = (uint32_t)ew * 1024u;               // TMA: this warp's 1 KB inside an operand block
    const uint32_t tslot_stride = (uint32_t)n_ops * (uint32_t)C3_EPI_OP;
    const uint32_t rd0 = (uint32_t)((lane * 32) ^ (((lane >> 2) & 1) << 4)), rd1 = rd0 ^ 16u;   // this lane's two halves in a box
    auto prefetch = [&](const Pos& ps, int k) {
      // extras of work item k of that super-tile -> this thread's slots (one cp.async group per item, possibly empty)
      const int kg = sub + C3_EPI_SUBS * k;
      const int tile = kg >> lc, ch8 = (kg & (chunks - 1)) << 3;
      if (TMA) {
        if (n_ops > 0 && ps.b >= 0 && kg < total_items && !(t.dbg & 8)) {     // warp-uniform; rows past the image are zero-filled
          __syncwarp();                                          // every lane is done with slot k
          if (lane == 0) {
            const uint32_t dst = epi_u32 + (uint32_t)k * tslot_stride + wslot;
            uint64_t* bar = &bar_epi[ew][k];
            const int row = ps.row + 4 * tile;
            fence_async_smem();
            mbar_arrive_expect_tx(bar, (uint32_t)n_ops * 1024u);
            if (ia >= 0) tma_load_4d(dst + (uint32_t)ia * (uint32_t)C3_EPI_OP, &tmA, ch8, ps.x0, row, ps.b, bar);
            if (io >= 0) tma_load_4d(dst + (uint32_t)io * (uint32_t)C3_EPI_OP, &tmO, ch8, ps.x0, row, ps.b, bar);
            if (im >= 0) tma_load_4d(dst + (uint32_t)im * (uint32_t)C3_EPI_OP, &tmM, ch8, ps.x0, row, ps.b, bar);
          }
        }
        return;
      }
      if (n_ops > 0 && ps.ok && kg < total_items && ps.row + 4 * tile < a.Hout && !(t.dbg & 8)) {
        const int pix = ps.pix + tile * tile_pix;
        const uint32_t dst = epi_u32 + (uint32_t)k * slot_stride + my_slot;
        if (ia >= 0) {
          const float* ap = pa + (long)(pix * ppa + ch8);
          cp_async16(dst + (uint32_t)ia * (uint32_t)C3_EPI_OP, ap);
          cp_async16(dst + (uint32_t)ia * (uint32_t)C3_EPI_OP + (uint32_t)C3_EPI_HALF, ap + 4);
        }
        if (io >= 0) {
          const float* op = a.out + (long)(pix * a.po + ch8);
          cp_async16(dst + (uint32_t)io * (uint32_t)C3_EPI_OP, op);
          cp_async16(dst + (uint32_t)io * (uint32_t)C3_EPI_OP + (uint32_t)C3_EPI_HALF, op + 4);
        }
        if (im >= 0) {
          const float* mp = a.omask + (long)(pix * a.pom + ch8);
          cp_async16(dst + (uint32_t)im * (uint32_t)C3_EPI_OP, mp);
          cp_async16(dst + (uint32_t)im * (uint32_t)C3_EPI_OP + (uint32_t)C3_EPI_HALF, mp + 4);
        }
      }
      cp_async_commit();
    };
    Pos cur = decode(tp);
#pragma unroll
    for (int k = 0; k < C3_SLOTS; ++k) prefetch(cur, k);

    uint32_t as = 0, aphase = 0, eph = 0;
    for (int st_i = blockIdx.x; st_i < t.n_super; st_i += gridDim.x) {
      tp.advance(t);
      const Pos nxt = decode(tp);
      const uint32_t acc_base = tmem_base + as * (uint32_t)(t.T * t.N) + ((uint32_t)(q * 32) << 16);
      if (lane == 0) C3_TIMED(0, wait(&bar_acc_full[as], aphase));
      __syncwarp();
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < C3_SLOTS; ++k) {
        const int kg = sub + C3_EPI_SUBS * k;
        if (TMA) {
          if (n_ops > 0 && kg < total_items && !(t.dbg & 8)) C3_TIMED(1, mbar_wait_short(&bar_epi[ew][k], eph));
        } else {
          cp_async_wait<C3_SLOTS - 1>();                       // this item's extras have landed
        }
        if (kg < total_items) {                                // warp-uniform
          const int tile = kg >> lc, ch8 = (kg & (chunks - 1)) << 3;
          float vk[KS][8];
          const uint32_t col = (uint32_t)(tile * t.N + ch8);
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) tmem_ld8(acc_base + col + (uint32_t)(kx * t.CP), vk[kx]);
          tmem_ld_wait();
          float r[8];
          if (KS == 3) {
            // out[x] = D[x-1][kx=0] + D[x][kx=1] + D[x+1][kx=2]
#pragma unroll
            for (int i = 0; i < 8; ++i)
              r[i] = __shfl_up_sync(0xffffffffu, vk[0][i], 1) + vk[1][i] + __shfl_down_sync(0xffffffffu, vk[KS - 1][i], 1);
          } else {
            // out[x] = sum_kx D[x + kx - pad][kx], pad = 1 (forward of the 4x4 SAME conv) or 2 (its data gradient)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float acc = 0.f;
#pragma unroll
              for (int kx = 0; kx < KS; ++kx) {
                const int sh = kx - PAD;                        // compile-time after unrolling
                if (sh < 0) acc += __shfl_up_sync(0xffffffffu, vk[kx][i], (unsigned)(-sh));
                else if (sh > 0) acc += __shfl_down_sync(0xffffffffu, vk[kx][i], (unsigned)sh);
                else acc += vk[kx][i];
              }
              r[i] = acc;
            }
          }
          if (cur.ok && cur.row + 4 * tile < a.Hout && !(t.dbg & 4)) {
            if (G || EPI <= 3) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch8);
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + ch8 + 4);
              r[0] += b0.x; r[1] += b0.y; r[2] += b0.z; r[3] += b0.w; r[4] += b1.x; r[5] += b1.y; r[6] += b1.z; r[7] += b1.w;
            }
            if (f_relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = fmaxf(r[i], 0.f);
            }
            // out = mask( relu2( relu(acc + bias) + res ) ) + add + previous
            // cp.async: [operand][half][thread] x 16 B; TMA: [operand][warp][32 px x 32 B, 32-byte swizzle]
            const uint8_t* slot = TMA ? epi + (uint32_t)k * tslot_stride + wslot + rd0 : epi + (uint32_t)k * slot_stride + my_slot;
            const int h1 = TMA ? (int)rd1 - (int)rd0 : C3_EPI_HALF;            // distance to the second half (+-16 when swizzled)
            float ev[8];
            if (ia >= 0) {
              const float4 e0 = *reinterpret_cast<const float4*>(slot + ia * C3_EPI_OP);
              const float4 e1 = *reinterpret_cast<const float4*>(slot + ia * C3_EPI_OP + h1);
              ev[0] = e0.x; ev[1] = e0.y; ev[2] = e0.z; ev[3] = e0.w; ev[4] = e1.x; ev[5] = e1.y; ev[6] = e1.z; ev[7] = e1.w;
              if (f_res) {
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] += ev[i];
              }
            }
            if (f_relu2) {
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = fmaxf(r[i], 0.f);
            }
            if (im >= 0) {
              const float4 m0 = *reinterpret_cast<const float4*>(slot + im * C3_EPI_OP);
              const float4 m1 = *reinterpret_cast<const float4*>(slot + im * C3_EPI_OP + h1);
              const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = mv[i] > 0.f ? r[i] : 0.f;
            }
            if (ia >= 0 && f_add) {
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] += ev[i];
            }
            if (io >= 0) {
              const float4 p0 = *reinterpret_cast<const float4*>(slot + io * C3_EPI_OP);
              const float4 p1 = *reinterpret_cast<const float4*>(slot + io * C3_EPI_OP + h1);
              r[0] += p0.x; r[1] += p0.y; r[2] += p0.z; r[3] += p0.w; r[4] += p1.x; r[5] += p1.y; r[6] += p1.z; r[7] += p1.w;
            }
            float4* dst = reinterpret_cast<float4*>(a.out + (long)((cur.pix + tile * tile_pix) * a.po + ch8));
            dst[0] = make_float4(r[0], r[1], r[2], r[3]);
            dst[1] = make_float4(r[4], r[5], r[6], r[7]);
          }
        }
        prefetch(nxt, k);                                      // this slot is free again: next super-tile's item k
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[as]);
      if (++as == (uint32_t)t.A) { as = 0; aphase ^= 1u; }
      cur = nxt;
      eph ^= 1u;
    }
    if (!TMA) cp_async_wait<0>();
  }
#ifdef MSAU_C3_PROF
  if (prof && lane == 0) {
    // role base: producers 0.., MMA 4.., epilogue 8..; slot 3 of each block = the warp's total resident cycles
    const int role = warp < C3_PROD_WARPS ? 0 : (warp < C3_PROD_WARPS + C3_MMA_WARPS ? 4 : 8);
    pc[role == 0 ? 2 : 3] += 0;
    const unsigned long long total = (unsigned long long)(clock64() - t_start);
    atomicAdd(&g_c3_prof[role + 0], pc[0]);
    atomicAdd(&g_c3_prof[role + 1], pc[1]);
    atomicAdd(&g_c3_prof[role + 2], role == 0 ? pc[3] : total);
    if (role == 0) atomicAdd(&g_c3_prof[3], total);
    if (warp == 0) atomicAdd(&g_c3_prof[12], 1ull);
  }
#endif
#undef C3_TIMED
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, t.tmem_cols);
}

bool c3_configure(const ConvArgs& a, C3Tile& t, bool tma = false) {
  t.CP = a.coutp;
  if (!(t.CP == 8 || t.CP == 16 || t.CP == 32 || t.CP == 64)) return false;
  t.KS = a.kh; t.pad = a.pad_l; t.OW = 32 - (t.KS - 1); t.NI = t.KS + (t.KS + 1) / 2;
  t.N = round_up(t.KS * t.CP, 16);
  t.T = 8 / (t.CP >> 3);
  { static int tmax = -1; if (tmax < 0) { const char* e = getenv("MSAU_C3_TMAX"); tmax = e ? atoi(e) : 0; } if (tmax > 0 && t.T > tmax) t.T = tmax; }
  while (t.T > 1 && 4 * (t.T / 2) >= a.Hout) t.T /= 2;       // short maps: do not pay for rows that do not exist
  t.RI = 4 * t.T + t.KS - 1;
  t.P = (a.c1 + a.c2) / 8;
  t.plane_bytes = (uint32_t)t.RI * 512;
  t.in_bytes = 2 * t.plane_bytes;
  t.tma = tma ? 1 : 0;
  // cp.async: [u][half][thread] x 16 B, rows warp + 6 u;  TMA: the dense plane [RI][32 px][8 ch] fp32
  t.raw_bytes = tma ? (uint32_t)t.RI * 1024u : (uint32_t)((t.RI > 18 ? 6 : (t.RI > 12 ? 3 : 2)) * 2 * C3_PROD_THREADS * 16);
  t.w_bytes = (uint32_t)t.NI * (uint32_t)t.N * 32;
  t.w_copy = (uint32_t)t.P * t.w_bytes;
  t.w_total = (t.w_copy + 1023u) / 1024u * 1024u;
  t.n_ops = 0; t.ia = t.io = t.im = -1;
  if (a.res || a.add) t.ia = t.n_ops++;
  if (a.accumulate) t.io = t.n_ops++;
  if (a.omask) t.im = t.n_ops++;
  const int items = (t.T * (t.CP >> 3) + C3_EPI_SUBS - 1) / C3_EPI_SUBS;   // epilogue work items per thread and super-tile
  t.epi_bytes = (uint32_t)(items * t.n_ops * C3_EPI_OP);
  t.stages = C3_MAX_STAGES; t.D = 4;
  { static int dmax = -1, smax = -1;
    if (dmax < 0) { const char* e = getenv("MSAU_C3_DMAX"); dmax = e ? atoi(e) : 0; const char* f = getenv("MSAU_C3_SMAX"); smax = f ? atoi(f) : 0; }
    if (dmax > 0 && t.D > dmax) t.D = dmax;
    if (smax > 0 && t.stages > smax) t.stages = smax; }
  auto total = [&]() { return (size_t)t.w_total + (size_t)t.stages * t.in_bytes + (size_t)t.D * t.raw_bytes + t.epi_bytes; };
  while (total() > 216 * 1024 && (t.D > 2 || t.stages > 2)) {
    if (t.D > 2 && t.D >= t.stages) --t.D; else --t.stages;
  }
  if (total() > 216 * 1024) return false;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MSAU_TC_DEBUG"); dbg = e ? atoi(e) : 0; } t.dbg = dbg; }
  t.blocks_x = cdiv(a.Wout, t.OW);
  t.blocks_y = cdiv(a.Hout, 4 * t.T);
  t.n_super = t.blocks_x * t.blocks_y * a.B;
  t.A = 512 / (t.T * t.N);
  if (t.A > 4) t.A = 4;
  if (t.A < 2) return false;
  const int cols = t.A * t.T * t.N;
  t.tmem_cols = 32;
  while ((int)t.tmem_cols < cols) t.tmem_cols <<= 1;
  return t.tmem_cols <= 512;
}

}  // namespace

bool conv3_tc_supported(const ConvArgs& a) {
  if (a.kh != a.kw || (a.kh != 3 && a.kh != 4) || a.dil != 1 || a.stride != 1 || a.pad_t != a.pad_l) return false;
  if (a.kh == 3 ? a.pad_l != 1 : (a.pad_l != 1 && a.pad_l != 2)) return false;
  if (a.kh == 4 && a.coutp != 8) return false;     // the 4x4 heads (8 -> n_class) and their data gradient
  if ((a.res && a.add) || a.addmask || a.mask1 || a.src1_nchw || a.s2d || a.d2s) return false;
  if (a.osy != 1 || a.oy0 != 0 || a.ox0 != 0) return false;
  if (a.Hq != a.Hin || a.Wq != a.Win || a.Hout != a.Hin || a.Wout != a.Win) return false;
  if ((a.c1 & 7) || (a.c2 & 7) || (a.p1 & 3) || (a.c2 && (a.p2 & 3)) || (a.po & 3)) return false;
  if ((a.res && (a.pr & 3)) || (a.add && (a.pa & 3)) || (a.omask && (a.pom & 3))) return false;
  if (a.Win < 8 || a.Hin < 2) return false;
  C3Tile t;
  return c3_configure(a, t);
}

// built with -DMSAU_C3_PROF and run with MSAU_TC_DEBUG=32: cycles summed over lane 0 of every warp of every conv3_tc launch since the last call:
// [0] producer wait stage-free  [1] producer wait raw plane  [2] producer named barrier  [3] producer warp total
// [4] MMA wait operands  [5] MMA wait accumulator  [6] MMA warp total  [8] epilogue wait accumulator  [9] epilogue wait extras
// [10] epilogue warp total  [12] CTAs
int debug_c3_prof(unsigned long long* h16) {
  MSAU_CUDA_TRY(cudaDeviceSynchronize());
  MSAU_CUDA_TRY(cudaMemcpyFromSymbol(h16, g_c3_prof, sizeof(unsigned long long) * 16));
  unsigned long long z[16] = {0};
  MSAU_CUDA_TRY(cudaMemcpyToSymbol(g_c3_prof, z, sizeof(z)));
  return MSAU_OK;
}

int launch_conv3_tc(const ConvArgs& a, const uint16_t* wtc, cudaStream_t st, int use_tma) {
  MSAU_CHECK_ARG(conv3_tc_supported(a), "conv3_tc: unsupported shape");
  C3Tile t;
  { static int env = -2; if (env == -2) { const char* e = getenv("MSAU_C3_TMA"); env = e ? atoi(e) : -1; } if (env >= 0) use_tma = env; }
  CUtensorMap tm1, tm2, tmA, tmO, tmM;
  memset(&tm1, 0, sizeof(tm1));
  memset(&tm2, 0, sizeof(tm2));
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmO, 0, sizeof(tmO));
  memset(&tmM, 0, sizeof(tmM));
  bool tma = use_tma != 0 && c3_configure(a, t, true);
  if (tma) {
    // {channel, x, y, page} maps of the two sources; box = one 8-channel halo plane of a super-tile
    tma = make_tmap_nhwc_f32(&tm1, a.src1, a.p1, a.Win, a.Hin, a.B, 8, 32, t.RI) &&
          (!a.c2 || make_tmap_nhwc_f32(&tm2, a.src2, a.p2, a.Win, a.Hin, a.B, 8, 32, t.RI));
    // epilogue operands: one image row of a tile (32 columns x 8 channels) per box, 32-byte swizzle
    const float* pa = a.res ? a.res : a.add;
    if (tma && pa) tma = make_tmap_nhwc_f32(&tmA, pa, a.res ? a.pr : a.pa, a.Wout, a.Hout, a.B, 8, 32, 1, true);
    if (tma && a.accumulate) tma = make_tmap_nhwc_f32(&tmO, a.out, a.po, a.Wout, a.Hout, a.B, 8, 32, 1, true);
    if (tma && a.omask) tma = make_tmap_nhwc_f32(&tmM, a.omask, a.pom, a.Wout, a.Hout, a.B, 8, 32, 1, true);
  }
  if (!tma) MSAU_CHECK_ARG(c3_configure(a, t, false), "conv3_tc: tile does not fit");
  const size_t smem = (size_t)t.w_total + (size_t)t.stages * t.in_bytes + (size_t)t.D * t.raw_bytes + t.epi_bytes + 1024;
  const int grid = t.n_super < sm_count() ? t.n_super : sm_count();
  t.dgx = grid % t.blocks_x;
  t.dgy = (grid / t.blocks_x) % t.blocks_y;
  t.dgb = (grid / t.blocks_x) / t.blocks_y;
  const bool general = a.res || a.omask || a.add || a.accumulate || a.relu2;
  const double npix = (double)a.B * a.Hin * a.Win;
  double bytes = npix * (a.c1 + a.c2) * 4.0;
  bytes += npix * a.coutp * 4.0 * (1 + (a.res ? 1 : 0) + (a.omask ? 1 : 0) + (a.add ? 1 : 0) + (a.accumulate ? 1 : 0));
  // (the 32-channel launches -- maps <= 128^2, latency-sized, on conv_tc until round 2 -- are booked as their own row, so that
  //  the 8/16-channel family stays the launch set round 1 reported)
  ProfScope ps(a.coutp >= 32 ? "conv3_tc_kernel_c32" : "conv3_tc_kernel", a.c1 + a.c2, a.coutp, a.kh, 1, a.Wout, (a.relu1 ? 2 : 0) + (general ? 1 : 0),
               2.0 * npix * a.kh * a.kw * (a.c1 + a.c2) * a.coutp, bytes, st);
  // epilogue specialisation (the flag-driven variant covers everything else, e.g. accumulation into a touched gradient)
  int epi = 0;
  const bool bias = a.bias != nullptr;
  if (!a.accumulate) {
    if (bias && !a.res && !a.add && !a.omask && !a.relu2) epi = a.relu ? 2 : 1;
    else if (bias && a.res && a.relu2 && !a.relu && !a.omask && !a.add) epi = 3;
    else if (!bias && a.omask && !a.res && !a.relu && !a.relu2) epi = a.add ? 5 : 4;
  }
  { static int gen = -1; if (gen < 0) { const char* e = getenv("MSAU_C3_GENERIC"); gen = e ? atoi(e) : 0; } if (gen) epi = 0; }
  const int ldu = t.RI > 18 ? 6 : (t.RI > 12 ? 3 : 2);
#define MSAU_C3_LAUNCH_T(E, L, K, PD, TM)                                                                                             \
  {                                                                                                                            \
    static bool attr = false;                                                                                                  \
    if (!attr) { MSAU_CUDA_TRY(cudaFuncSetAttribute(conv3_tc_kernel<E, L, K, PD, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; } \
    MSAU_CUDA_TRY(launch_pdl(conv3_tc_kernel<E, L, K, PD, TM>, dim3(grid), dim3(C3_THREADS), smem, st, a, wtc, t, tm1, tm2, tmA, tmO, tmM)); \
  }
#define MSAU_C3_LAUNCH(E, L, K, PD) { if (tma) MSAU_C3_LAUNCH_T(E, L, K, PD, true) else MSAU_C3_LAUNCH_T(E, L, K, PD, false) }
#define MSAU_C3_LDU(E) { if (ldu == 6) MSAU_C3_LAUNCH(E, 6, 3, 1) else if (ldu == 3) MSAU_C3_LAUNCH(E, 3, 3, 1) else MSAU_C3_LAUNCH(E, 2, 3, 1) }
#define MSAU_C3_K4(E, PD) { if (ldu == 6) MSAU_C3_LAUNCH(E, 6, 4, PD) else MSAU_C3_LAUNCH(E, 3, 4, PD) }
  if (t.KS == 4) {
    if (t.pad == 1) { if (epi == 1) MSAU_C3_K4(1, 1) else MSAU_C3_K4(0, 1) }
    else { if (epi == 1) MSAU_C3_K4(1, 2) else MSAU_C3_K4(0, 2) }
  } else
  switch (epi) {
    case 1: MSAU_C3_LDU(1) break;
    case 2: MSAU_C3_LDU(2) break;
    case 3: MSAU_C3_LDU(3) break;
    case 4: MSAU_C3_LDU(4) break;
    case 5: MSAU_C3_LDU(5) break;
    default: MSAU_C3_LDU(0) break;
  }
#undef MSAU_C3_LDU
#undef MSAU_C3_K4
#undef MSAU_C3_LAUNCH
#undef MSAU_C3_LAUNCH_T
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// ------------------------------------------------------------------ weight images
// src: fp32 packed [tap = ky*3+kx][cin][coutp].  dst (bf16) per 8-channel plane p: 5 images of [chunk0: N x 8][chunk1: N x 8],
// row n = kx * coutp + co:  images 0..2 = {Whi(ky), Whi(ky)},  image 3 = {Wlo(0), Wlo(1)},  image 4 = {Wlo(2), 0}
__global__ void __launch_bounds__(256) pack_tc3_kernel(const float* __restrict__ pk, uint16_t* __restrict__ pktc,
                                                        const TcPackDesc* __restrict__ descs, int n_desc) {
  // one thread per [8 ch] slot of an image row (16 bytes out; consecutive threads = consecutive output channels: coalesced reads)
  const int blk = (int)blockIdx.x;
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].blk0 <= blk) lo = mid; else hi = mid - 1;
  }
  const TcPackDesc d = descs[lo];
  const int N = d.N;
  const int KS = d.taps == 16 ? 4 : 3, NI = KS + (KS + 1) / 2;
  const int per_plane = NI * N * 2;                              // slots per 8-channel plane: images x 2 chunks x N rows
  const int total = (d.cin / 8) * per_plane;
  const int e = (blk - (int)d.blk0) * 256 + (int)threadIdx.x;
  if (e >= total) return;
  const int p = e / per_plane;
  int r = e - p * per_plane;
  const int img = r / (2 * N); r -= img * 2 * N;
  const int chunk = r / N, n = r - chunk * N;
  const int kx = n / d.coutp, co = n - kx * d.coutp;
  int ky = -1;
  bool want_lo = false;
  if (img < KS) ky = img;
  else {
    want_lo = true;
    ky = 2 * (img - KS) + chunk;
    if (ky >= KS) ky = -1;
  }
  uint32_t h2[4] = {0u, 0u, 0u, 0u};
  if (ky >= 0 && kx < KS) {
    const float* src = pk + d.src_off + ((long)(ky * KS + kx) * d.cin + 8 * p) * d.coutp + co;
    float w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = __ldg(src + (long)k * d.coutp);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __nv_bfloat16 h = __float2bfloat16_rn(w[k]);
      const __nv_bfloat16 out = want_lo ? __float2bfloat16_rn(w[k] - __bfloat162float(h)) : h;
      h2[k >> 1] |= (uint32_t)__bfloat16_as_ushort(out) << (16 * (k & 1));
    }
  }
  *reinterpret_cast<uint4*>(pktc + d.dst_off + (long)e * 8) = make_uint4(h2[0], h2[1], h2[2], h2[3]);     // (dst_off is a multiple of 128)
}

int launch_pack_tc3(const float* pk, uint16_t* pktc, const TcPackDesc* d_descs, int n_desc, long total_blocks, cudaStream_t st) {
  if (n_desc == 0) return MSAU_OK;
  ProfScope ps("pack_tc_kernel", 0, (double)total_blocks * 256 * 48.0, st);
  pack_tc3_kernel<<<(unsigned)total_blocks, 256, 0, st>>>(pk, pktc, d_descs, n_desc);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

}  // namespace msau
