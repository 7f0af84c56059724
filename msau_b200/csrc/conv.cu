// Generic fp32 NHWC convolution kernels (CUDA-core path).
//
//  * conv_kernel  : direct convolution with smem-staged input tile + weight chunk.  One kernel covers
//                   the forward of every conv flavour of MSAU (3x3, dilated 3x3, 1x1 on a concat, 4x4 with
//                   asymmetric SAME padding, the 4 sub-pixel phases of the transposed conv) and, with
//                   flipped/transposed packed weights, every data-gradient (incl. the stride-2 gather that
//                   is the transposed conv's dgrad).  Epilogue: bias -> relu -> +res -> relu -> *mask ->
//                   +add*mask -> (+=).
//  * wgrad_kernel : weight gradient as a split-K outer-product reduction read straight through L1
//                   (pixels are the K dimension), register-blocked 4(ca) x 4(cb) x KW taps per thread.
//
// Reference semantics: model/layers/layers.py:10-164,207-260 (conv / dilated conv / transposed conv),
// model/layers/utils.py:5-28 (SAME padding -> pad_t/pad_l here), model/model.py:37-50 (residual epilogue).
#include "common.cuh"
#include "prof.cuh"

namespace msau {

struct ConvTile {
  int ncb, CB, NG, RP, TH, TIH, TIW, GK;
};

template <int CO_T, int PIX_T>
__global__ void __launch_bounds__(256) conv_kernel(const ConvArgs a, const ConvTile t) {
  extern __shared__ float4 smem4[];
  if (a.skip_flag && *a.skip_flag == 0) return;   // one-hot input: first_layer.cu produced this output
  const int tid = threadIdx.x;
  const int lane = tid & 31, wy = tid >> 5;
  const int cog = wy % t.NG, rip = wy / t.NG;
  const bool active = rip < t.RP;
  const int b = blockIdx.z / t.ncb;
  const int cout0 = (blockIdx.z % t.ncb) * t.CB;
  const int qx0 = blockIdx.x * 32, qy0 = blockIdx.y * t.TH;
  const int in_x0 = qx0 * a.stride - a.pad_l, in_y0 = qy0 * a.stride - a.pad_t;
  const int taps = a.kh * a.kw;
  const int cin = a.c1 + a.c2;
  const int G = cin >> 2;
  const int tile_px = t.TIH * t.TIW;
  float4* in_s = smem4;
  float* w_s = reinterpret_cast<float*>(smem4 + t.GK * tile_px);

  float acc[PIX_T][CO_T];
#pragma unroll
  for (int p = 0; p < PIX_T; ++p)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[p][c] = 0.f;

  for (int g0 = 0; g0 < G; g0 += t.GK) {
    const int ng = min(t.GK, G - g0);
    __syncthreads();
    // ---- stage the input tile (ng channel-quads) ----
    for (int idx = tid; idx < ng * tile_px; idx += 256) {
      const int g = idx / tile_px;
      const int rem = idx - g * tile_px;
      const int iy = rem / t.TIW, ix = rem - iy * t.TIW;
      const int gy = in_y0 + iy, gx = in_x0 + ix;
      const int cg = (g0 + g) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < a.Hin && gx >= 0 && gx < a.Win) {
        const long pix = ((long)b * a.Hin + gy) * a.Win + gx;
        if (cg < a.c1) {
          if (a.src1_nchw) {
            const long plane = (long)a.Hin * a.Win;
            const float* s = a.src1 + ((long)b * a.c1_logical + cg) * plane + (long)gy * a.Win + gx;
            if (cg + 0 < a.c1_logical) v.x = __ldg(s);
            if (cg + 1 < a.c1_logical) v.y = __ldg(s + plane);
            if (cg + 2 < a.c1_logical) v.z = __ldg(s + 2 * plane);
            if (cg + 3 < a.c1_logical) v.w = __ldg(s + 3 * plane);
          } else {
            v = __ldg(reinterpret_cast<const float4*>(a.src1 + pix * a.p1 + cg));
          }
          if (a.mask1) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(a.mask1 + pix * a.pm1 + cg));
            v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
            v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
          }
          if (a.relu1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        } else {
          v = __ldg(reinterpret_cast<const float4*>(a.src2 + pix * a.p2 + (cg - a.c1)));
        }
      }
      in_s[idx] = v;
    }
    // ---- stage the weight chunk: w_s[tap][c (4*ng)][CB] ----
    {
      const int rows = ng * 4;
      const int cb4 = t.CB >> 2;
      const int n4 = taps * rows * cb4;
      for (int idx = tid; idx < n4; idx += 256) {
        const int co4 = idx % cb4;
        const int r = idx / cb4;
        const int c = r % rows, tap = r / rows;
        const float4 wv = __ldg(reinterpret_cast<const float4*>(
            a.w + ((long)tap * cin + (g0 << 2) + c) * a.coutp + cout0 + (co4 << 2)));
        reinterpret_cast<float4*>(w_s)[idx] = wv;
      }
    }
    __syncthreads();
    if (active) {
      const int rows = ng * 4;
      for (int ky = 0; ky < a.kh; ++ky) {
        for (int kx = 0; kx < a.kw; ++kx) {
          const int tap = ky * a.kw + kx;
          const int col = lane * a.stride + kx * a.dil;
          for (int g = 0; g < ng; ++g) {
            float4 xv[PIX_T];
#pragma unroll
            for (int p = 0; p < PIX_T; ++p) {
              const int row = (rip + p * t.RP) * a.stride + ky * a.dil;
              xv[p] = in_s[g * tile_px + row * t.TIW + col];
            }
            const float* wrow = w_s + ((tap * rows + g * 4) * t.CB) + cog * CO_T;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
              for (int q = 0; q < CO_T / 4; ++q) {
                const float4 wv = *reinterpret_cast<const float4*>(wrow + c * t.CB + q * 4);
#pragma unroll
                for (int p = 0; p < PIX_T; ++p) {
                  const float xs = c == 0 ? xv[p].x : (c == 1 ? xv[p].y : (c == 2 ? xv[p].z : xv[p].w));
                  acc[p][q * 4 + 0] = fmaf(xs, wv.x, acc[p][q * 4 + 0]);
                  acc[p][q * 4 + 1] = fmaf(xs, wv.y, acc[p][q * 4 + 1]);
                  acc[p][q * 4 + 2] = fmaf(xs, wv.z, acc[p][q * 4 + 2]);
                  acc[p][q * 4 + 3] = fmaf(xs, wv.w, acc[p][q * 4 + 3]);
                }
              }
            }
          }
        }
      }
    }
  }
  if (!active) return;
  // ---- epilogue ----
  const int qx = qx0 + lane;
  if (qx >= a.Wq) return;
  const int ox = qx * a.osy + a.ox0;
  const int cbase = cout0 + cog * CO_T;
#pragma unroll
  for (int p = 0; p < PIX_T; ++p) {
    const int qy = qy0 + rip + p * t.RP;
    if (qy >= a.Hq) continue;
    const int oy = qy * a.osy + a.oy0;
    if (oy >= a.Hout || ox >= a.Wout) continue;
    const long pix = ((long)b * a.Hout + oy) * a.Wout + ox;
#pragma unroll
    for (int q = 0; q < CO_T / 4; ++q) {
      const int co = cbase + q * 4;
      float4 v = make_float4(acc[p][q * 4 + 0], acc[p][q * 4 + 1], acc[p][q * 4 + 2], acc[p][q * 4 + 3]);
      if (a.bias) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + co));
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      }
      if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (a.res) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(a.res + pix * a.pr + co));
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      if (a.relu2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (a.omask) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(a.omask + pix * a.pom + co));
        v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
        v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
      }
      if (a.add) {
        float4 r = __ldg(reinterpret_cast<const float4*>(a.add + pix * a.pa + co));
        if (a.addmask) {
          const float4 m = __ldg(reinterpret_cast<const float4*>(a.addmask + pix * a.pam + co));
          r.x = m.x > 0.f ? r.x : 0.f; r.y = m.y > 0.f ? r.y : 0.f;
          r.z = m.z > 0.f ? r.z : 0.f; r.w = m.w > 0.f ? r.w : 0.f;
        }
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      float4* dst = reinterpret_cast<float4*>(a.out + pix * a.po + co);
      if (a.accumulate) {
        const float4 o = *dst;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *dst = v;
    }
  }
}

int launch_conv(const ConvArgs& a, cudaStream_t st) {
  MSAU_CHECK_ARG(a.c1 % 4 == 0 && a.c2 % 4 == 0 && a.coutp % 8 == 0, "conv: channel counts must be padded (c1=%d c2=%d cout=%d)", a.c1, a.c2, a.coutp);
  MSAU_CHECK_ARG(a.src1_nchw || a.p1 % 4 == 0, "conv: src pitch must be a multiple of 4");
  ConvTile t;
  const bool co16 = (a.coutp % 16 == 0);
  const int CO_T = co16 ? 16 : 8;
  t.CB = a.coutp;
  while (t.CB > 64 || (t.CB / CO_T) > 8) {
    if (t.CB % 2) break;
    t.CB /= 2;
  }
  MSAU_CHECK_ARG(t.CB % CO_T == 0 && a.coutp % t.CB == 0 && t.CB / CO_T <= 8, "conv: unsupported cout %d", a.coutp);
  t.ncb = a.coutp / t.CB;
  t.NG = t.CB / CO_T;
  t.RP = 8 / t.NG;
  const int PIX_T = 4;
  t.TH = t.RP * PIX_T;
  t.TIH = (t.TH - 1) * a.stride + (a.kh - 1) * a.dil + 1;
  t.TIW = 31 * a.stride + (a.kw - 1) * a.dil + 1;
  const int G = (a.c1 + a.c2) / 4;
  t.GK = G >= 2 ? 2 : 1;
  size_t smem = (size_t)t.GK * t.TIH * t.TIW * 16 + (size_t)a.kh * a.kw * t.GK * 4 * t.CB * 4;
  if (smem > 200 * 1024) {
    t.GK = 1;
    smem = (size_t)t.GK * t.TIH * t.TIW * 16 + (size_t)a.kh * a.kw * t.GK * 4 * t.CB * 4;
  }
  MSAU_CHECK_ARG(smem <= 220 * 1024, "conv: tile does not fit shared memory (%zu B)", smem);
  dim3 grid(cdiv(a.Wq, 32), cdiv(a.Hq, t.TH), a.B * t.ncb);
  MSAU_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "conv: grid too large");
  // algorithmic work: every operand tensor crosses HBM once; 2 flops per multiply-add
  const double npix_in = (double)a.B * a.Hin * a.Win, npix_out = (double)a.B * a.Hq * a.Wq;
  double bytes = npix_in * ((a.src1_nchw ? a.c1_logical : a.c1) + a.c2 + (a.mask1 ? a.c1 : 0)) * 4.0;
  bytes += npix_out * a.coutp * 4.0 * (1 + (a.res ? 1 : 0) + (a.omask ? 1 : 0) + (a.add ? 1 : 0) + (a.addmask ? 1 : 0) + (a.accumulate ? 1 : 0));
  const double flops = 2.0 * npix_out * a.kh * a.kw * (a.c1 + a.c2) * a.coutp;
  ProfScope ps(co16 ? "conv_kernel<16,4>" : "conv_kernel<8,4>", flops, bytes, st);
  if (co16) {
    static bool attr = false;
    if (!attr) { MSAU_CUDA_TRY(cudaFuncSetAttribute(conv_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    conv_kernel<16, 4><<<grid, 256, smem, st>>>(a, t);
  } else {
    static bool attr = false;
    if (!attr) { MSAU_CUDA_TRY(cudaFuncSetAttribute(conv_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    conv_kernel<8, 4><<<grid, 256, smem, st>>>(a, t);
  }
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

// =====================================================================================================
// wgrad
// =====================================================================================================
struct WgradTile {
  int MB, NB, TM, TN, TT, KS, n_mb, n_nb, rows_per_block;
};

template <int KW, bool BTAP>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a, const WgradTile t) {
  extern __shared__ float red[];   // [KW][MB][NB] + [NB]
  if (a.skip_flag && *a.skip_flag == 0) return;   // one-hot input: first_layer.cu produced this gradient
  const int tid = threadIdx.x;
  const int tile = tid % t.TT, ks = tid / t.TT;
  const bool active = ks < t.KS;
  const int tm = tile % t.TM, tn = tile / t.TM;
  // blockIdx.x enumerates (ky, m-block, n-block); blockIdx.y the pixel-row range
  int bx = blockIdx.x;
  const int nbk = bx % t.n_nb; bx /= t.n_nb;
  const int mbk = bx % t.n_mb; bx /= t.n_mb;
  const int ky = bx;
  const int ca0 = mbk * t.MB + tm * 4, cb0 = nbk * t.NB + tn * 4;
  const long total_rows = (long)a.B * a.Hq;
  const long r0 = (long)blockIdx.y * t.rows_per_block;
  const long r1 = min(total_rows, r0 + t.rows_per_block);

  float acc[KW][4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < KW; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[k][i][j] = 0.f;
  const bool do_bias = (a.dbias != nullptr) && ky == 0 && mbk == 0 && tm == 0;

  if (active) {
    const long planeA = (long)a.Ha * a.Wa;
    for (long r = r0; r < r1; ++r) {
      const int b = (int)(r / a.Hq);
      const int qy = (int)(r - (long)b * a.Hq);
      const int ya = a.sa * qy + ky * a.dila - a.pada_t;
      const int yb = a.sb * qy + ky * a.dilb - a.padb_t;
      const bool ya_ok = ya >= 0 && ya < a.Ha;
      const bool yb_ok = yb >= 0 && yb < a.Hb;
      if (!BTAP && !yb_ok) continue;
      if (BTAP && !ya_ok) continue;
      if (!BTAP && !ya_ok && !do_bias) continue;
      for (int qx = ks; qx < a.Wq; qx += t.KS) {
        float4 av[BTAP ? 1 : KW];
        float4 bv[BTAP ? KW : 1];
#pragma unroll
        for (int k = 0; k < (BTAP ? 1 : KW); ++k) {
          const int xa = a.sa * qx + k * a.dila - a.pada_l;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ya_ok && xa >= 0 && xa < a.Wa) {
            if (a.a_nchw) {
              const float* s = a.A + ((long)b * a.ca_logical + ca0) * planeA + (long)ya * a.Wa + xa;
              if (ca0 + 0 < a.ca_logical) v.x = __ldg(s);
              if (ca0 + 1 < a.ca_logical) v.y = __ldg(s + planeA);
              if (ca0 + 2 < a.ca_logical) v.z = __ldg(s + 2 * planeA);
              if (ca0 + 3 < a.ca_logical) v.w = __ldg(s + 3 * planeA);
            } else {
              v = __ldg(reinterpret_cast<const float4*>(a.A + (((long)b * a.Ha + ya) * a.Wa + xa) * a.pa + ca0));
            }
            if (a.reluA) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          }
          av[k] = v;
        }
#pragma unroll
        for (int k = 0; k < (BTAP ? KW : 1); ++k) {
          const int xb = a.sb * qx + k * a.dilb - a.padb_l;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (yb_ok && xb >= 0 && xb < a.Wb) {
            const long pix = ((long)b * a.Hb + yb) * a.Wb + xb;
            v = __ldg(reinterpret_cast<const float4*>(a.Bm + pix * a.pb + cb0));
            if (a.maskB) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(a.maskB + pix * a.pmb + cb0));
              v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
              v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
            }
          }
          bv[k] = v;
        }
        if (do_bias) { bsum[0] += bv[0].x; bsum[1] += bv[0].y; bsum[2] += bv[0].z; bsum[3] += bv[0].w; }
#pragma unroll
        for (int k = 0; k < KW; ++k) {
          const float4 x = av[BTAP ? 0 : k];
          const float4 y = bv[BTAP ? k : 0];
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[k][i][0] = fmaf(xs[i], y.x, acc[k][i][0]);
            acc[k][i][1] = fmaf(xs[i], y.y, acc[k][i][1]);
            acc[k][i][2] = fmaf(xs[i], y.z, acc[k][i][2]);
            acc[k][i][3] = fmaf(xs[i], y.w, acc[k][i][3]);
          }
        }
      }
    }
  }
  // ---- reduce over the KS pixel-splits in shared memory (fixed order => deterministic per block) ----
  const int nred = KW * t.MB * t.NB;
  for (int i = tid; i < nred + t.NB; i += 256) red[i] = 0.f;
  __syncthreads();
  for (int s = 0; s < t.KS; ++s) {
    if (active && ks == s) {
#pragma unroll
      for (int k = 0; k < KW; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) red[(k * t.MB + tm * 4 + i) * t.NB + tn * 4 + j] += acc[k][i][j];
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) red[nred + tn * 4 + j] += bsum[j];
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < nred; i += 256) {
    const int n = i % t.NB;
    const int m = (i / t.NB) % t.MB;
    const int k = i / (t.NB * t.MB);
    const int ca = mbk * t.MB + m, cb = nbk * t.NB + n;
    if (ca < a.ca_lim && cb < a.cb_lim) atomicAdd(a.dW + ca * a.s_ca + cb * a.s_cb + (ky * a.kw + k), red[i]);
  }
  if (a.dbias != nullptr && ky == 0 && mbk == 0) {
    for (int n = tid; n < t.NB; n += 256) {
      const int cb = nbk * t.NB + n;
      if (cb < a.cb_lim) atomicAdd(a.dbias + cb, red[nred + n]);
    }
  }
}

int launch_wgrad(const WgradArgs& a, cudaStream_t st) {
  MSAU_CHECK_ARG(a.ca % 4 == 0 && a.cb % 4 == 0, "wgrad: channels must be multiples of 4");
  const bool btap = (a.dilb != 0);
  MSAU_CHECK_ARG(!(btap && a.dila != 0), "wgrad: only one operand may depend on the tap");
  MSAU_CHECK_ARG(!(btap && a.dbias), "wgrad: bias gradient needs a tap-independent B");
  WgradTile t;
  auto pick = [](int c) { for (int m = 64; m >= 4; m -= 4) if (c % m == 0) return m; return 4; };
  t.MB = pick(a.ca);
  t.NB = pick(a.cb);
  MSAU_CHECK_ARG(a.ca % t.MB == 0 && a.cb % t.NB == 0, "wgrad: unsupported channel counts %d x %d", a.ca, a.cb);
  t.TM = t.MB / 4; t.TN = t.NB / 4; t.TT = t.TM * t.TN;
  t.KS = 256 / t.TT;
  if (t.KS > a.Wq) t.KS = a.Wq;
  t.n_mb = a.ca / t.MB; t.n_nb = a.cb / t.NB;
  const int outblocks = a.kh * t.n_mb * t.n_nb;
  const long rows = (long)a.B * a.Hq;
  int want = (8 * sm_count() + outblocks - 1) / outblocks;
  if (want < 1) want = 1;
  if (want > rows) want = (int)rows;
  t.rows_per_block = (int)((rows + want - 1) / want);
  const int ny = (int)((rows + t.rows_per_block - 1) / t.rows_per_block);
  dim3 grid(outblocks, ny);
  const size_t smem = ((size_t)a.kw * t.MB * t.NB + t.NB) * 4;
  const double npq = (double)a.B * a.Hq * a.Wq;
  const double wbytes = ((double)a.B * a.Ha * a.Wa * (a.a_nchw ? a.ca_logical : a.ca) + (double)a.B * a.Hb * a.Wb * a.cb * (a.maskB ? 2 : 1)) * 4.0;
  const double wflops = 2.0 * npq * a.kh * a.kw * a.ca * a.cb;
  ProfScope ps(btap ? "wgrad_kernel<3,btap>" : (a.kw == 1 ? "wgrad_kernel<1>" : (a.kw == 3 ? "wgrad_kernel<3>" : "wgrad_kernel<4>")), wflops, wbytes, st);
#define MSAU_WG(KWV, BT)                                                                                    \
  {                                                                                                         \
    static bool attr = false;                                                                               \
    if (!attr) { MSAU_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel<KWV, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; } \
    wgrad_kernel<KWV, BT><<<grid, 256, smem, st>>>(a, t);                                                   \
  }
  if (btap) {
    MSAU_CHECK_ARG(a.kw == 3, "wgrad: tap-dependent B supports kw=3 only");
    MSAU_WG(3, true)
  } else if (a.kw == 1) MSAU_WG(1, false)
  else if (a.kw == 3) MSAU_WG(3, false)
  else if (a.kw == 4) MSAU_WG(4, false)
  else { set_error("wgrad: unsupported kernel width %d", a.kw); return MSAU_ERR_UNSUPPORTED; }
#undef MSAU_WG
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

}  // namespace msau
