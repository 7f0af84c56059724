// Spatial self-attention of MSAU's deepest scale on the 5th-generation tensor cores (tcgen05.mma, TMEM).
//
// Reference: SelfAttentionBlock.forward, model/layers/attention.py:152-162 (see attention.cu for the fp32
// CUDA-core version and the derivation of the backward formulas):
//     S = G F^T (N x N, never materialised)   P[i,j] = exp(S[i,j] - lse_i)   (soft-max over j)
//     O[j,:] = sum_i P[i,j] Hh[i,:]           (contraction over i)            out = x + O
//     dHh = P dO          D_i = Hh[i,:] . dHh[i,:]        dP = Hh dO^T        dS = P o (dP - D_i)
//     dG  = dS F  =  (P o dP) F - D_i (P F)               dF = dS^T G
//
// Every pass has the same shape: a CTA owns 128 "stationary" positions (the 128 TMEM lanes) and streams over
// all positions in chunks of 128, 64 columns at a time:
//     score MMAs   T[lane, col] = sum_k A[lane,k] B[col,k]    (A: stationary operand, B: streamed, both
//                  K-major bf16 images in shared memory, accumulator in TMEM)
//     element-wise 8 warps read T with tcgen05.ld, apply exp2 (and the dP product), write the result back to
//                  TMEM as packed bf16 (tcgen05.st), in place over the scores
//     output MMAs  Acc[lane, :] += E[lane, col] V[col, :]     (A operand = E in TMEM, B = MN-major image)
// fp32-level accuracy from bf16 operands by the same hi/lo split as conv_tc.cu: the K dimension of a score
// MMA carries the slots  f_hi*g_hi | f_lo*g_hi | f_hi*g_lo | 1*(-lse_i, three bf16 pieces) | pad*(-30000),
// so the accumulator IS log2(P) (f is pre-scaled by log2 e, the row log-sum-exp is folded in, padded
// positions vanish) and the element-wise stage is a bare ex2.  The forward output uses three output MMAs
// (P_hi V_hi + P_lo V_hi + P_hi V_lo); the backward passes use single bf16 terms for dP and the output
// products (each result is a sum over >= 10^3 positions of zero-mean rounding errors, as in wgrad_tc.cu).
//
// Operand images live in a caller-provided scratch buffer, layout [batch][chunk of 128][plane][128][8] bf16:
// a (chunk, all planes) block is one contiguous TMA bulk copy, a plane is both a K-major core-matrix column
// (K = the 8 elements) and an MN-major one (K = positions), so one image serves both kinds of MMA.
//   Fs  [f' hi | f' lo | f' hi | 1,1,1,pad]      score operand, F side          (f' = f log2 e)
//   Gs  [g hi  | g hi  | g lo  | -lse x3,-3e4]   score operand, G side          (lse patched in by the stats pass)
//   Hi  [h hi (C/8 planes) | -D_i x3 | 0]        V of the forward, dP operand of the backward
//   Hlo [h lo]                                   forward only
//   Oi  [dO (C/8) | 1,1,1 | f | 1]               dP operand, V of dHh / dG (the 1 columns give row sums)
//   Gv  [g | 0]                                  V of dF
#include "attention.cuh"
#include "common.cuh"
#include "prof.cuh"
#include "tc_ptx.cuh"

namespace msau {
using namespace ptx;

static constexpr int AT_CH = 128;                       // positions per chunk / stationary tile
static constexpr uint32_t AT_PLANE_B = AT_CH * 16;      // bytes of one plane of one chunk
static constexpr int AT_PLANE_E = AT_CH * 8;            // bf16 elements of one plane of one chunk
static constexpr float AT_NEG = -30000.f;
static constexpr float AT_LOG2E = 1.4426950408889634f;

enum { AT_STATS = 0, AT_OUT = 1, AT_BWD_HG = 2, AT_BWD_F = 3 };

template <int C>
struct AtDims {
  static constexpr int D = C / 8;
  static constexpr int K1 = (3 * D + 4 + 15) / 16 * 16;   // score slots (3 split terms + 3 lse pieces + pad flag)
  static constexpr int PS = K1 / 8;                         // planes of Fs / Gs
  static constexpr int CP = C / 8;                          // planes of a C-channel tensor
  static constexpr int AUGP = (3 * D) / 8, AUGE = (3 * D) % 8;   // plane / first element of the augmentation slots
  static constexpr int PHI = CP + 2, POI = CP + 3, PGV = 2;
  static constexpr int N1 = C + 16;                         // backward accumulator 0: [dHh (C) | row sums of P (8) | P F (8)]
  static_assert(D <= 8 && AUGE + 4 <= 8, "augmentation slots must share one plane");
};

struct AtImages { uint16_t *Fs, *Gs, *Hi, *Hlo, *Oi, *Gv; };

template <int C>
static size_t at_planes_total() {
  using Dm = AtDims<C>;
  return 2 * Dm::PS + Dm::PHI + Dm::CP + Dm::POI + Dm::PGV;
}

template <int C>
static AtImages at_carve(void* scratch, int B, int nch) {
  using Dm = AtDims<C>;
  uint16_t* p = reinterpret_cast<uint16_t*>(scratch);
  const size_t per = (size_t)B * nch * AT_PLANE_E;
  AtImages im;
  im.Fs = p; p += per * Dm::PS;
  im.Gs = p; p += per * Dm::PS;
  im.Hi = p; p += per * Dm::PHI;
  im.Hlo = p; p += per * Dm::CP;
  im.Oi = p; p += per * Dm::POI;
  im.Gv = p;
  return im;
}

__device__ __forceinline__ void split3(float x, uint16_t& a, uint16_t& b, uint16_t& c) {
  a = bf16_bits(x);
  float r = x - bf16_val(a);
  b = bf16_bits(r);
  r -= bf16_val(b);
  c = bf16_bits(r);
}
__device__ __forceinline__ uint4 pack_u16x8(const uint16_t* v) {
  return make_uint4((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16),
                    (uint32_t)v[4] | ((uint32_t)v[5] << 16), (uint32_t)v[6] | ((uint32_t)v[7] << 16));
}

// ------------------------------------------------------------------ operand images
// one thread per position, one block per (chunk, batch image)
template <int C, bool BWD>
__global__ void __launch_bounds__(AT_CH) attn_prep_kernel(const float* __restrict__ FG, const float* __restrict__ Hh,
                                                           const float* __restrict__ dO, const float* __restrict__ lse, int N,
                                                           AtImages im) {
  using Dm = AtDims<C>;
  constexpr int D = Dm::D;
  const int chunk = blockIdx.x, b = blockIdx.y, nch = gridDim.x, t = threadIdx.x;
  const int pos = chunk * AT_CH + t;
  const bool valid = pos < N;
  const long gp = (long)b * N + pos;
  const long cb = (long)b * nch + chunk;
  const uint16_t one = 0x3F80;
  float f[D], g[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    f[k] = valid ? __ldg(FG + gp * 2 * D + k) : 0.f;
    g[k] = valid ? __ldg(FG + gp * 2 * D + D + k) : 0.f;
  }
  {
    uint16_t fs[Dm::K1], gs[Dm::K1];
#pragma unroll
    for (int k = 0; k < Dm::K1; ++k) { fs[k] = 0; gs[k] = 0; }
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float fp = f[k] * AT_LOG2E;
      const uint16_t fh = bf16_bits(fp), fl = bf16_bits(fp - bf16_val(fh));
      const uint16_t gh = bf16_bits(g[k]), gl = bf16_bits(g[k] - bf16_val(gh));
      fs[k] = fh; gs[k] = gh;
      fs[D + k] = fl; gs[D + k] = gh;
      fs[2 * D + k] = fh; gs[2 * D + k] = gl;
    }
    fs[3 * D] = one; fs[3 * D + 1] = one; fs[3 * D + 2] = one;
    fs[3 * D + 3] = valid ? 0 : one;
    gs[3 * D + 3] = bf16_bits(AT_NEG);
    if (!valid) {
      gs[3 * D] = bf16_bits(AT_NEG);
    } else if (BWD) {
      split3(-__ldg(lse + gp), gs[3 * D], gs[3 * D + 1], gs[3 * D + 2]);
    }
#pragma unroll
    for (int p = 0; p < Dm::PS; ++p) {
      *reinterpret_cast<uint4*>(im.Fs + (cb * Dm::PS + p) * AT_PLANE_E + t * 8) = pack_u16x8(fs + 8 * p);
      *reinterpret_cast<uint4*>(im.Gs + (cb * Dm::PS + p) * AT_PLANE_E + t * 8) = pack_u16x8(gs + 8 * p);
    }
  }
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int p = 0; p < Dm::CP; ++p) {
    float v[8];
    if (valid) {
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(Hh + gp * C + 8 * p));
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(Hh + gp * C + 8 * p) + 1);
      v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    uint4 hi, lo;
    split_pair(v[0], v[1], hi.x, lo.x); split_pair(v[2], v[3], hi.y, lo.y);
    split_pair(v[4], v[5], hi.z, lo.z); split_pair(v[6], v[7], hi.w, lo.w);
    *reinterpret_cast<uint4*>(im.Hi + (cb * Dm::PHI + p) * AT_PLANE_E + t * 8) = hi;
    if (!BWD) *reinterpret_cast<uint4*>(im.Hlo + (cb * Dm::CP + p) * AT_PLANE_E + t * 8) = lo;
  }
  if (BWD) {
    *reinterpret_cast<uint4*>(im.Hi + (cb * Dm::PHI + Dm::CP) * AT_PLANE_E + t * 8) = zero4;       // -D_i: written by the dHh pass
    *reinterpret_cast<uint4*>(im.Hi + (cb * Dm::PHI + Dm::CP + 1) * AT_PLANE_E + t * 8) = zero4;
#pragma unroll
    for (int p = 0; p < Dm::CP; ++p) {
      uint4 o = zero4;
      if (valid) {
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(dO + gp * C + 8 * p));
        const float4 q1 = __ldg(reinterpret_cast<const float4*>(dO + gp * C + 8 * p) + 1);
        o = make_uint4(pack_bf16(q0.x, q0.y), pack_bf16(q0.z, q0.w), pack_bf16(q1.x, q1.y), pack_bf16(q1.z, q1.w));
      }
      *reinterpret_cast<uint4*>(im.Oi + (cb * Dm::POI + p) * AT_PLANE_E + t * 8) = o;
    }
    uint16_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = k < 3 ? one : 0;
    *reinterpret_cast<uint4*>(im.Oi + (cb * Dm::POI + Dm::CP) * AT_PLANE_E + t * 8) = pack_u16x8(w);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = k < D ? bf16_bits(f[k < D ? k : 0]) : 0;
    *reinterpret_cast<uint4*>(im.Oi + (cb * Dm::POI + Dm::CP + 1) * AT_PLANE_E + t * 8) = pack_u16x8(w);
    *reinterpret_cast<uint4*>(im.Oi + (cb * Dm::POI + Dm::CP + 2) * AT_PLANE_E + t * 8) = make_uint4(one, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = k < D ? bf16_bits(g[k < D ? k : 0]) : 0;
    *reinterpret_cast<uint4*>(im.Gv + (cb * Dm::PGV) * AT_PLANE_E + t * 8) = pack_u16x8(w);
    *reinterpret_cast<uint4*>(im.Gv + (cb * Dm::PGV + 1) * AT_PLANE_E + t * 8) = zero4;
  }
}

// ------------------------------------------------------------------ the streaming kernel
struct AtArgs {
  const uint16_t *stat1, *stat2, *str1, *str2, *str3;   // image bases
  int stat1_pt, stat2_pt, str1_pt, str2_pt, str3_pt;     // planes per chunk of each image (its full plane count)
  int N, nch;
  const float* X; float* out;                            // AT_OUT
  float* lse; uint16_t* Gs_w;                            // AT_STATS
  float* dHh; float* dFG; uint16_t* Hi_w;                // backward
};

template <int MODE, int C>
struct AtCfg {
  using Dm = AtDims<C>;
  static constexpr bool BWD = MODE == AT_BWD_HG || MODE == AT_BWD_F;
  static constexpr int STAT1_PL = Dm::PS;
  static constexpr int STAT2_PL = MODE == AT_BWD_HG ? Dm::CP : (MODE == AT_BWD_F ? Dm::CP + 2 : 0);
  static constexpr int STR1_PL = Dm::PS;
  static constexpr int STR2_PL = MODE == AT_OUT ? Dm::CP : (MODE == AT_BWD_HG ? Dm::POI : (MODE == AT_BWD_F ? Dm::PHI : 0));
  static constexpr int STR3_PL = MODE == AT_OUT ? Dm::CP : (MODE == AT_BWD_F ? Dm::PGV : 0);
  static constexpr uint32_t STAT_BYTES = (STAT1_PL + STAT2_PL) * AT_PLANE_B;
  static constexpr uint32_t STAGE_BYTES = (STR1_PL + STR2_PL + STR3_PL) * AT_PLANE_B;
  static constexpr int NS = MODE == AT_STATS ? 4 : 3;
  static constexpr int KS1 = Dm::K1 / 16;
  static constexpr int KS2 = MODE == AT_BWD_HG ? C / 16 : (MODE == AT_BWD_F ? (C + 16) / 16 : 0);
  static constexpr int SBW = BWD ? 128 : 64;             // TMEM columns of one score buffer
  static constexpr int ACC0 = 2 * SBW;
  static constexpr int ACC0_W = MODE == AT_OUT ? C : (MODE == AT_BWD_HG ? Dm::N1 : (MODE == AT_BWD_F ? 16 : 0));
  static constexpr int ACC1 = ACC0 + ACC0_W;
  static constexpr int COLS = ACC1 + (MODE == AT_BWD_HG ? 16 : 0);
  static constexpr int TMEM_COLS = COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512);
  static constexpr size_t SMEM = STAT_BYTES + (size_t)NS * STAGE_BYTES + 1024;
};

static constexpr int AT_THREADS = 320;   // 8 element-wise warps + TMA producer + MMA issuer

template <int MODE, int C>
__global__ void __launch_bounds__(AT_THREADS, 1) attn_tc_kernel(const AtArgs a) {
  using Cf = AtCfg<MODE, C>;
  using Dm = AtDims<C>;
  constexpr int D = Dm::D;
  constexpr int NS = Cf::NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_stat, bar_full[NS], bar_free[NS], bar_s[2], bar_p[2], bar_acc;
  __shared__ uint32_t tmem_base_s;
  __shared__ float exch[2][AT_CH];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, b = blockIdx.y;
  const int T = 2 * a.nch;                               // 64-column sub-chunks
  if (warp == 9) tmem_alloc(&tmem_base_s, Cf::TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar_stat, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 256); }
    mbar_init(&bar_acc, 1);
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  uint8_t* stat_s = smem;
  uint8_t* stage_s = smem + Cf::STAT_BYTES;
  const long cb0 = (long)b * a.nch;

  if (warp == 8) {
    // =============================================================== TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_stat, Cf::STAT_BYTES);
      bulk_g2s(stat_s, a.stat1 + (cb0 + tile) * a.stat1_pt * AT_PLANE_E, Cf::STAT1_PL * AT_PLANE_B, &bar_stat);
      if (Cf::STAT2_PL)
        bulk_g2s(stat_s + Cf::STAT1_PL * AT_PLANE_B, a.stat2 + (cb0 + tile) * a.stat2_pt * AT_PLANE_E, Cf::STAT2_PL * AT_PLANE_B,
                 &bar_stat);
      for (int ch = 0; ch < a.nch; ++ch) {
        const int s = ch % NS;
        if (ch >= NS) mbar_wait(&bar_free[s], ((ch / NS) - 1) & 1);
        uint8_t* dst = stage_s + (size_t)s * Cf::STAGE_BYTES;
        mbar_arrive_expect_tx(&bar_full[s], Cf::STAGE_BYTES);
        bulk_g2s(dst, a.str1 + (cb0 + ch) * a.str1_pt * AT_PLANE_E, Cf::STR1_PL * AT_PLANE_B, &bar_full[s]);
        if (Cf::STR2_PL)
          bulk_g2s(dst + Cf::STR1_PL * AT_PLANE_B, a.str2 + (cb0 + ch) * a.str2_pt * AT_PLANE_E, Cf::STR2_PL * AT_PLANE_B,
                   &bar_full[s]);
        if (Cf::STR3_PL)
          bulk_g2s(dst + (Cf::STR1_PL + Cf::STR2_PL) * AT_PLANE_B, a.str3 + (cb0 + ch) * a.str3_pt * AT_PLANE_E,
                   Cf::STR3_PL * AT_PLANE_B, &bar_full[s]);
      }
    }
  } else if (warp == 9) {
    // =============================================================== MMA issuer (whole warp walks, one lane issues)
    const uint32_t idesc_s = make_idesc(128, 64, false, false);
    const uint32_t idesc_v0 = make_idesc(128, Cf::ACC0_W > 0 ? Cf::ACC0_W : 16, false, true);
    const uint32_t idesc_v16 = make_idesc(128, 16, false, true);
    const uint32_t stat_addr = smem_u32(stat_s);
    constexpr uint32_t KSTEP16 = (2 * AT_PLANE_B) >> 4;    // descriptor increment of one K step (two planes)
    mbar_wait(&bar_stat, 0);
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        const int ch = t >> 1, hb = t & 1, s = ch % NS;
        if (hb == 0) mbar_wait(&bar_full[s], (ch / NS) & 1);
        if (MODE == AT_STATS && t >= 2) mbar_wait(&bar_p[t & 1], ((t >> 1) - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t stg = smem_u32(stage_s + (size_t)s * Cf::STAGE_BYTES);
          const uint32_t sb = tmem_base + (uint32_t)((t & 1) * Cf::SBW);
          const uint64_t da = make_desc(stat_addr, AT_PLANE_B, 128);
          const uint64_t db = make_desc(stg + hb * 64 * 16, AT_PLANE_B, 128);
#pragma unroll
          for (int ks = 0; ks < Cf::KS1; ++ks) mma_ss(sb, da + ks * KSTEP16, db + ks * KSTEP16, idesc_s, ks > 0 ? 1u : 0u);
          if (Cf::BWD) {
            const uint64_t da2 = make_desc(stat_addr + Cf::STAT1_PL * AT_PLANE_B, AT_PLANE_B, 128);
            const uint64_t db2 = make_desc(stg + Cf::STR1_PL * AT_PLANE_B + hb * 64 * 16, AT_PLANE_B, 128);
#pragma unroll
            for (int ks = 0; ks < Cf::KS2; ++ks) mma_ss(sb + 64, da2 + ks * KSTEP16, db2 + ks * KSTEP16, idesc_s, ks > 0 ? 1u : 0u);
          }
          tc_commit(&bar_s[t & 1]);
          if (MODE == AT_STATS && hb == 1) tc_commit(&bar_free[s]);
        }
        __syncwarp();
      }
      if (MODE != AT_STATS && t >= 1) {
        const int u = t - 1, ch = u >> 1, hb = u & 1, s = ch % NS;
        mbar_wait(&bar_p[u & 1], (u >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t stg = smem_u32(stage_s + (size_t)s * Cf::STAGE_BYTES);
          const uint32_t sb = tmem_base + (uint32_t)((u & 1) * Cf::SBW);
          const uint32_t acc0 = tmem_base + Cf::ACC0, acc1 = tmem_base + Cf::ACC1;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t colA = sb + 32 * (kk >> 1) + 8 * (kk & 1);
            const uint32_t k0b = (uint32_t)(hb * 64 + kk * 16) * 16;
            const uint32_t first = (u == 0 && kk == 0) ? 0u : 1u;
            if (MODE == AT_OUT) {
              const uint64_t vhi = make_desc(stg + Cf::STR1_PL * AT_PLANE_B + k0b, 128, AT_PLANE_B);
              const uint64_t vlo = make_desc(stg + (Cf::STR1_PL + Cf::STR2_PL) * AT_PLANE_B + k0b, 128, AT_PLANE_B);
              mma_ts(acc0, colA, vhi, idesc_v0, first);
              mma_ts(acc0, colA + 16, vhi, idesc_v0, 1u);
              mma_ts(acc0, colA, vlo, idesc_v0, 1u);
            } else if (MODE == AT_BWD_HG) {
              const uint64_t v1 = make_desc(stg + Cf::STR1_PL * AT_PLANE_B + k0b, 128, AT_PLANE_B);
              const uint64_t v2 = make_desc(stg + (Cf::STR1_PL + Dm::CP + 1) * AT_PLANE_B + k0b, 128, AT_PLANE_B);
              mma_ts(acc0, colA, v1, idesc_v0, first);
              mma_ts(acc1, colA + 64, v2, idesc_v16, first);
            } else {
              const uint64_t vg = make_desc(stg + (Cf::STR1_PL + Cf::STR2_PL) * AT_PLANE_B + k0b, 128, AT_PLANE_B);
              mma_ts(acc0, colA, vg, idesc_v16, first);
            }
          }
          if (hb == 1) tc_commit(&bar_free[s]);
          if (u == T - 1) tc_commit(&bar_acc);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================================================== element-wise warps
    const int q = warp & 3, hf = warp >> 2;               // TMEM lane quarter, column half
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;                        // stationary position inside the tile
    const int pos = tile * AT_CH + row;
    const bool valid = pos < a.N;
    const long gp = (long)b * a.N + pos;
    float m = -INFINITY, z = 0.f;
    float psum = 0.f, dsum = 0.f;                         // backward: fp32 row sums of P and P o dP (for D_i)
    for (int t = 0; t < T; ++t) {
      const uint32_t sb = tmem_base + lane_base + (uint32_t)((t & 1) * Cf::SBW);
      if (lane == 0) mbar_wait(&bar_s[t & 1], (t >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(sb + 32 * hf, r);
      if (MODE == AT_STATS) {
        tmem_ld_wait();
        float tm = __uint_as_float(r[0]);
#pragma unroll
        for (int c = 1; c < 32; ++c) tm = fmaxf(tm, __uint_as_float(r[c]));
        if (tm > m) { z *= ex2f(m - tm); m = tm; }
        float zz = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) zz += ex2f(__uint_as_float(r[c]) - m);
        z += zz;
      } else if (MODE == AT_OUT) {
        tmem_ld_wait();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) split_pair(ex2f(__uint_as_float(r[2 * c])), ex2f(__uint_as_float(r[2 * c + 1])), hi[c], lo[c]);
        tmem_st16(sb + 32 * hf, hi);
        tmem_st16(sb + 32 * hf + 16, lo);
        tmem_st_wait();
      } else {
        uint32_t dp[32];
        tmem_ld32(sb + 64 + 32 * hf, dp);
        tmem_ld_wait();
        uint32_t e0[16], e1[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float p0 = ex2f(__uint_as_float(r[2 * c])), p1 = ex2f(__uint_as_float(r[2 * c + 1]));
          const float d0 = p0 * __uint_as_float(dp[2 * c]), d1 = p1 * __uint_as_float(dp[2 * c + 1]);
          if (MODE == AT_BWD_HG) {
            e0[c] = pack_bf16(p0, p1); e1[c] = pack_bf16(d0, d1);
            psum += p0 + p1; dsum += d0 + d1;
          }
          else e0[c] = pack_bf16(d0, d1);
        }
        tmem_st16(sb + 32 * hf, e0);
        if (MODE == AT_BWD_HG) tmem_st16(sb + 64 + 32 * hf, e1);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&bar_p[t & 1]);
    }
    // ------------------------------------------------------------- epilogues
    if (MODE == AT_STATS) {
      if (hf == 1) { exch[0][row] = m; exch[1][row] = z; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (hf == 0 && valid) {
        const float m1 = exch[0][row], z1 = exch[1][row];
        const float mm = fmaxf(m, m1);
        const float zz = z * ex2f(m - mm) + z1 * ex2f(m1 - mm);
        const float l2 = mm + lg2f(zz);
        a.lse[gp] = l2;
        uint16_t p0, p1, p2;
        split3(-l2, p0, p1, p2);
        uint16_t* dst = a.Gs_w + ((cb0 + tile) * Dm::PS + Dm::AUGP) * AT_PLANE_E + row * 8 + Dm::AUGE;
        dst[0] = p0; dst[1] = p1; dst[2] = p2;
      }
    } else {
      if (MODE == AT_BWD_HG) {
        if (hf == 1) { exch[0][row] = psum; exch[1][row] = dsum; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (lane == 0) mbar_wait(&bar_acc, 0);
      __syncwarp();
      tc_fence_after();
      const uint32_t acc0 = tmem_base + lane_base + Cf::ACC0;
      if (MODE == AT_OUT) {
        constexpr int HC = C / 2;                          // channels per column half
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(acc0 + hf * HC + c0, v);
          tmem_ld_wait();
          if (valid) {
            const float4* xs = reinterpret_cast<const float4*>(a.X + gp * C + hf * HC + c0);
            float4* os = reinterpret_cast<float4*>(a.out + gp * C + hf * HC + c0);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 x = __ldg(xs + c4);
              os[c4] = make_float4(x.x + __uint_as_float(v[4 * c4]), x.y + __uint_as_float(v[4 * c4 + 1]),
                                   x.z + __uint_as_float(v[4 * c4 + 2]), x.w + __uint_as_float(v[4 * c4 + 3]));
            }
          }
        }
      } else if (MODE == AT_BWD_HG) {
        if (hf == 0) {
#pragma unroll
          for (int c0 = 0; c0 < C; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(acc0 + c0, v);
            tmem_ld_wait();
            if (valid) {
              float4* ds = reinterpret_cast<float4*>(a.dHh + gp * C + c0);
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4)
                ds[c4] = make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]),
                                     __uint_as_float(v[4 * c4 + 3]));
            }
          }
          uint32_t pf[16], pdf[16];
          tmem_ld16(acc0 + C, pf);                         // [row sums (8) | P F (D)]
          tmem_ld16(tmem_base + lane_base + Cf::ACC1, pdf);   // (P o dP) F
          tmem_ld_wait();
          if (valid) {
            // D_i as the P-weighted mean of dP in exactly the arithmetic of the sums it is subtracted from (the row
            // sums of the bf16-rounded P o dP and P): sum_j P (dP - D_i) F is then free of cancellation error.
            // (Hh[i,:] . dHh[i,:] is the same number up to rounding, but rounded differently.)
            const float Dt = __uint_as_float(pdf[8]) / __uint_as_float(pf[0]);
#pragma unroll
            for (int k = 0; k < D; ++k)
              a.dFG[gp * 2 * D + D + k] = __uint_as_float(pdf[k]) - Dt * __uint_as_float(pf[8 + k]);
            // the dF pass rounds P o (dP - D_i) once, after the subtraction: it gets D_i from the fp32 row sums
            const float Df = (dsum + exch[1][row]) / (psum + exch[0][row]);
            uint16_t p0, p1, p2;
            split3(-Df, p0, p1, p2);
            uint16_t* dst = a.Hi_w + ((cb0 + tile) * Dm::PHI + Dm::CP) * AT_PLANE_E + row * 8;
            dst[0] = p0; dst[1] = p1; dst[2] = p2;
          }
        }
      } else {
        if (hf == 0) {
          uint32_t v[16];
          tmem_ld16(acc0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int k = 0; k < D; ++k) a.dFG[gp * 2 * D + k] = __uint_as_float(v[k]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, Cf::TMEM_COLS);
}

// ------------------------------------------------------------------ host
bool attn_tc_supported(int C, int d) { return (C == 64 && d == 8) || (C == 32 && d == 4); }

size_t attn_tc_scratch_bytes(int B, int N, int C) {
  const size_t nch = (size_t)cdiv(N, AT_CH);
  const size_t planes = C == 64 ? at_planes_total<64>() : at_planes_total<32>();
  return (size_t)B * nch * planes * AT_PLANE_B + 256;
}

template <int MODE, int C>
static int at_launch(const AtArgs& a, int B, cudaStream_t st) {
  using Cf = AtCfg<MODE, C>;
  static bool attr = false;
  if (!attr) {
    MSAU_CUDA_TRY(cudaFuncSetAttribute(attn_tc_kernel<MODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cf::SMEM));
    attr = true;
  }
  dim3 grid(a.nch, B);
  attn_tc_kernel<MODE, C><<<grid, AT_THREADS, Cf::SMEM, st>>>(a);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

template <int C>
static int attn_tc_fwd_impl(const float* FG, const float* Hh, const float* X, int B, int N, float* lse, float* out, void* scratch,
                            cudaStream_t st) {
  using Dm = AtDims<C>;
  const int nch = cdiv(N, AT_CH);
  const AtImages im = at_carve<C>(scratch, B, nch);
  attn_prep_kernel<C, false><<<dim3(nch, B), AT_CH, 0, st>>>(FG, Hh, nullptr, nullptr, N, im);
  MSAU_CUDA_TRY(cudaGetLastError());
  AtArgs a;
  memset(&a, 0, sizeof(a));
  a.N = N; a.nch = nch;
  a.stat1 = im.Gs; a.stat1_pt = Dm::PS; a.str1 = im.Fs; a.str1_pt = Dm::PS;
  a.lse = lse; a.Gs_w = im.Gs;
  MSAU_TRY((at_launch<AT_STATS, C>(a, B, st)));
  memset(&a, 0, sizeof(a));
  a.N = N; a.nch = nch;
  a.stat1 = im.Fs; a.stat1_pt = Dm::PS;
  a.str1 = im.Gs; a.str1_pt = Dm::PS; a.str2 = im.Hi; a.str2_pt = Dm::PHI; a.str3 = im.Hlo; a.str3_pt = Dm::CP;
  a.X = X; a.out = out;
  MSAU_TRY((at_launch<AT_OUT, C>(a, B, st)));
  return MSAU_OK;
}

template <int C>
static int attn_tc_bwd_impl(const float* FG, const float* Hh, const float* dO, const float* lse, int B, int N, float* dFG, float* dHh,
                            void* scratch, cudaStream_t st) {
  using Dm = AtDims<C>;
  const int nch = cdiv(N, AT_CH);
  const AtImages im = at_carve<C>(scratch, B, nch);
  attn_prep_kernel<C, true><<<dim3(nch, B), AT_CH, 0, st>>>(FG, Hh, dO, lse, N, im);
  MSAU_CUDA_TRY(cudaGetLastError());
  AtArgs a;
  memset(&a, 0, sizeof(a));
  a.N = N; a.nch = nch;
  a.stat1 = im.Gs; a.stat1_pt = Dm::PS; a.stat2 = im.Hi; a.stat2_pt = Dm::PHI;
  a.str1 = im.Fs; a.str1_pt = Dm::PS; a.str2 = im.Oi; a.str2_pt = Dm::POI;
  a.dHh = dHh; a.dFG = dFG; a.Hi_w = im.Hi;
  MSAU_TRY((at_launch<AT_BWD_HG, C>(a, B, st)));
  memset(&a, 0, sizeof(a));
  a.N = N; a.nch = nch;
  a.stat1 = im.Fs; a.stat1_pt = Dm::PS; a.stat2 = im.Oi; a.stat2_pt = Dm::POI;
  a.str1 = im.Gs; a.str1_pt = Dm::PS; a.str2 = im.Hi; a.str2_pt = Dm::PHI; a.str3 = im.Gv; a.str3_pt = Dm::PGV;
  a.dFG = dFG;
  MSAU_TRY((at_launch<AT_BWD_F, C>(a, B, st)));
  return MSAU_OK;
}

int launch_attn_tc_fwd(const float* FG, const float* Hh, const float* X, int B, int N, int C, int d, float* lse, float* out,
                       void* scratch, cudaStream_t st) {
  MSAU_CHECK_ARG(attn_tc_supported(C, d), "attention (tensor core): unsupported (C=%d, d=%d); supported (64,8) and (32,4)", C, d);
  MSAU_CHECK_ARG(((uintptr_t)scratch & 127) == 0, "attention scratch must be 128-byte aligned");
  // prep + row statistics sweep (d MACs) + output sweep (d + C MACs); 1 exp per entry per sweep
  ProfScope ps("attn_fwd_kernels", 2.0 * B * (double)N * N * (2 * d + C), (double)B * N * (2 * d + 3 * C + 2) * 4.0, st);
  return C == 64 ? attn_tc_fwd_impl<64>(FG, Hh, X, B, N, lse, out, scratch, st)
                 : attn_tc_fwd_impl<32>(FG, Hh, X, B, N, lse, out, scratch, st);
}

int launch_attn_tc_bwd(const float* FG, const float* Hh, const float* dO, const float* lse, int B, int N, int C, int d, float* dFG,
                       float* dHh, void* scratch, cudaStream_t st) {
  MSAU_CHECK_ARG(attn_tc_supported(C, d), "attention (tensor core): unsupported (C=%d, d=%d); supported (64,8) and (32,4)", C, d);
  MSAU_CHECK_ARG(((uintptr_t)scratch & 127) == 0, "attention scratch must be 128-byte aligned");
  ProfScope ps("attn_bwd_kernels", 2.0 * B * (double)N * N * (5 * d + 3 * C), (double)B * N * (4 * d + 3 * C + 3) * 4.0, st);
  return C == 64 ? attn_tc_bwd_impl<64>(FG, Hh, dO, lse, B, N, dFG, dHh, scratch, st)
                 : attn_tc_bwd_impl<32>(FG, Hh, dO, lse, B, N, dFG, dHh, scratch, st);
}

}  // namespace msau
