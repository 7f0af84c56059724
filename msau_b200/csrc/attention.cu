// Fused spatial self-attention of MSAU's deepest scale, forward and backward, without ever
// materialising the N x N relation map (N = H*W positions of the deepest feature map).
//
// Reference: SelfAttentionBlock.forward, model/layers/attention.py:152-162
//     s[i,j]   = sum_c g[c,i] f[c,j]                (f,g : d = C/8 channels, 1x1 convs of x)
//     beta[i,:] = softmax_j s[i,:]                   (normalised over j ...)
//     o[c,j]   = sum_i h[c,i] beta[i,j]              (... but contracted over i)
//     out      = x + o
// Layout here: FG [B,N,2d] (f = channels 0..d-1, g = d..2d-1), Hh / x / out [B,N,C], all fp32 NHWC.
//
// forward  : attn_stats  (row i : m_i = max_j s, zinv_i = 1/sum_j exp(s-m_i))          i-stationary
//            attn_out    (col j : out[j,:] = x[j,:] + sum_i p_ij Hh[i,:])               j-stationary
// backward : attn_bwd_h  (row i : dHh[i,:] = sum_j p_ij dO[j,:] ; D_i = Hh[i,:].dHh[i,:]) i-stationary
//            attn_bwd_g  (row i : dG[i,:] = sum_j p_ij (Hh[i,:].dO[j,:] - D_i) F[j,:])    i-stationary
//            attn_bwd_f  (col j : dF[j,:] = sum_i p_ij (Hh[i,:].dO[j,:] - D_i) G[i,:])    j-stationary
// with p_ij = exp(s_ij - m_i) * zinv_i.  The head dimension is tiny (d = 8), so these kernels are
// exp/FMA-bound on the CUDA cores, not tensor-bound.
#include "common.cuh"
#include "attention.cuh"
#include "prof.cuh"

namespace msau {

static constexpr int TR = 128;   // stationary rows per block (one per thread)
static constexpr int TT = 32;    // streamed positions per shared-memory tile

template <int D>
__device__ __forceinline__ float dotD(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) s = fmaf(a[k], b[k], s);
  return s;
}

// ------------------------------------------------------------------ forward: row statistics
template <int D>
__global__ void __launch_bounds__(TR) attn_stats_kernel(const float* __restrict__ FG, int N, float* __restrict__ mrow,
                                                         float* __restrict__ zinv) {
  __shared__ float Fs[TT][D];
  const int b = blockIdx.y;
  const int i = blockIdx.x * TR + threadIdx.x;
  const float* fg = FG + (long)b * N * 2 * D;
  float g[D];
#pragma unroll
  for (int k = 0; k < D; ++k) g[k] = i < N ? __ldg(fg + (long)i * 2 * D + D + k) : 0.f;
  float m = -INFINITY, z = 0.f;
  for (int j0 = 0; j0 < N; j0 += TT) {
    const int cnt = min(TT, N - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < TT * D; e += TR) {
      const int jj = e / D, k = e - jj * D;
      Fs[jj][k] = jj < cnt ? __ldg(fg + (long)(j0 + jj) * 2 * D + k) : 0.f;
    }
    __syncthreads();
    float s[TT];
    float tm = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < TT; ++jj) {
      s[jj] = dotD<D>(g, Fs[jj]);
      if (jj < cnt) tm = fmaxf(tm, s[jj]);
    }
    if (tm > m) { z *= __expf(m - tm); m = tm; }
#pragma unroll
    for (int jj = 0; jj < TT; ++jj)
      if (jj < cnt) z += __expf(s[jj] - m);
  }
  if (i < N) {
    mrow[(long)b * N + i] = m;
    zinv[(long)b * N + i] = 1.f / z;
  }
}

// ------------------------------------------------------------------ forward: output (j-stationary)
// C = channel chunk handled by this CTA (blockIdx.z selects it), CT = channel pitch of Hh / X / out
template <int D, int C>
__global__ void __launch_bounds__(TR) attn_out_kernel(const float* __restrict__ FG, const float* __restrict__ Hh,
                                                       const float* __restrict__ X, const float* __restrict__ mrow,
                                                       const float* __restrict__ zinv, int N, int CT, float* __restrict__ out) {
  __shared__ __align__(16) float Gs[TT][D];
  __shared__ __align__(16) float Hs[TT][C];
  __shared__ float Ms[TT], Zs[TT];
  const int b = blockIdx.y;
  const int j = blockIdx.x * TR + threadIdx.x;
  const float* fg = FG + (long)b * N * 2 * D;
  const float* hh = Hh + (long)b * N * CT + blockIdx.z * C;
  float f[D];
#pragma unroll
  for (int k = 0; k < D; ++k) f[k] = j < N ? __ldg(fg + (long)j * 2 * D + k) : 0.f;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (int i0 = 0; i0 < N; i0 += TT) {
    const int cnt = min(TT, N - i0);
    __syncthreads();
    for (int e = threadIdx.x; e < TT * D; e += TR) {
      const int ii = e / D, k = e - ii * D;
      Gs[ii][k] = ii < cnt ? __ldg(fg + (long)(i0 + ii) * 2 * D + D + k) : 0.f;
    }
    for (int e = threadIdx.x; e < TT * C / 4; e += TR) {
      const int ii = e / (C / 4), c4 = e - ii * (C / 4);
      reinterpret_cast<float4*>(&Hs[ii][0])[c4] =
          ii < cnt ? __ldg(reinterpret_cast<const float4*>(hh + (long)(i0 + ii) * CT) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x < TT) {
      const int ii = threadIdx.x;
      Ms[ii] = ii < cnt ? mrow[(long)b * N + i0 + ii] : 0.f;
      Zs[ii] = ii < cnt ? zinv[(long)b * N + i0 + ii] : 0.f;   // zinv = 0 kills padded rows
    }
    __syncthreads();
#pragma unroll 4
    for (int ii = 0; ii < TT; ++ii) {
      const float p = __expf(dotD<D>(Gs[ii], f) - Ms[ii]) * Zs[ii];
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 h = reinterpret_cast<const float4*>(&Hs[ii][0])[c4];
        acc[4 * c4 + 0] = fmaf(p, h.x, acc[4 * c4 + 0]);
        acc[4 * c4 + 1] = fmaf(p, h.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(p, h.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(p, h.w, acc[4 * c4 + 3]);
      }
    }
  }
  if (j < N) {
    const float4* xs = reinterpret_cast<const float4*>(X + ((long)b * N + j) * CT + blockIdx.z * C);
    float4* os = reinterpret_cast<float4*>(out + ((long)b * N + j) * CT + blockIdx.z * C);
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 x = __ldg(xs + c4);
      os[c4] = make_float4(x.x + acc[4 * c4], x.y + acc[4 * c4 + 1], x.z + acc[4 * c4 + 2], x.w + acc[4 * c4 + 3]);
    }
  }
}

// ------------------------------------------------------------------ backward: dHh and D (i-stationary)
template <int D, int C>
__global__ void __launch_bounds__(TR) attn_bwd_h_kernel(const float* __restrict__ FG, const float* __restrict__ Hh,
                                                         const float* __restrict__ dO, const float* __restrict__ mrow,
                                                         const float* __restrict__ zinv, int N, int CT, float* __restrict__ dHh,
                                                         float* __restrict__ Dvec) {
  __shared__ __align__(16) float Fs[TT][D];
  __shared__ __align__(16) float Os[TT][C];
  const int b = blockIdx.y;
  const int i = blockIdx.x * TR + threadIdx.x;
  const float* fg = FG + (long)b * N * 2 * D;
  const float* go = dO + (long)b * N * CT + blockIdx.z * C;
  float g[D];
#pragma unroll
  for (int k = 0; k < D; ++k) g[k] = i < N ? __ldg(fg + (long)i * 2 * D + D + k) : 0.f;
  const float m = i < N ? mrow[(long)b * N + i] : 0.f;
  const float zi = i < N ? zinv[(long)b * N + i] : 0.f;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (int j0 = 0; j0 < N; j0 += TT) {
    const int cnt = min(TT, N - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < TT * D; e += TR) {
      const int jj = e / D, k = e - jj * D;
      Fs[jj][k] = jj < cnt ? __ldg(fg + (long)(j0 + jj) * 2 * D + k) : 0.f;
    }
    for (int e = threadIdx.x; e < TT * C / 4; e += TR) {
      const int jj = e / (C / 4), c4 = e - jj * (C / 4);
      reinterpret_cast<float4*>(&Os[jj][0])[c4] =
          jj < cnt ? __ldg(reinterpret_cast<const float4*>(go + (long)(j0 + jj) * CT) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < TT; ++jj) {
      const float p = __expf(dotD<D>(g, Fs[jj]) - m) * zi;   // padded columns have dO = 0
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 o = reinterpret_cast<const float4*>(&Os[jj][0])[c4];
        acc[4 * c4 + 0] = fmaf(p, o.x, acc[4 * c4 + 0]);
        acc[4 * c4 + 1] = fmaf(p, o.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(p, o.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(p, o.w, acc[4 * c4 + 3]);
      }
    }
  }
  if (i < N) {
    const float4* hs = reinterpret_cast<const float4*>(Hh + ((long)b * N + i) * CT + blockIdx.z * C);
    float4* ds = reinterpret_cast<float4*>(dHh + ((long)b * N + i) * CT + blockIdx.z * C);
    float dsum = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 h = __ldg(hs + c4);
      dsum = fmaf(h.x, acc[4 * c4], dsum); dsum = fmaf(h.y, acc[4 * c4 + 1], dsum);
      dsum = fmaf(h.z, acc[4 * c4 + 2], dsum); dsum = fmaf(h.w, acc[4 * c4 + 3], dsum);
      ds[c4] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
    }
    if (gridDim.z == 1) Dvec[(long)b * N + i] = dsum;
    else atomicAdd(Dvec + (long)b * N + i, dsum);       // channel chunks: Dvec zeroed by the launcher
  }
}

// ------------------------------------------------------------------ backward: dG (i-stationary)
template <int D, int C>
__global__ void __launch_bounds__(TR) attn_bwd_g_kernel(const float* __restrict__ FG, const float* __restrict__ Hh,
                                                         const float* __restrict__ dO, const float* __restrict__ mrow,
                                                         const float* __restrict__ zinv, const float* __restrict__ Dvec, int N,
                                                         int CT, float* __restrict__ dFG) {
  __shared__ __align__(16) float Fs[TT][D];
  __shared__ __align__(16) float Os[TT][C];
  const int b = blockIdx.y;
  const int i = blockIdx.x * TR + threadIdx.x;
  const float* fg = FG + (long)b * N * 2 * D;
  const float* go = dO + (long)b * N * CT + blockIdx.z * C;
  float g[D], h[C], acc[D];
#pragma unroll
  for (int k = 0; k < D; ++k) { g[k] = i < N ? __ldg(fg + (long)i * 2 * D + D + k) : 0.f; acc[k] = 0.f; }
#pragma unroll
  for (int c = 0; c < C; ++c) h[c] = i < N ? __ldg(Hh + ((long)b * N + i) * CT + blockIdx.z * C + c) : 0.f;
  const float m = i < N ? mrow[(long)b * N + i] : 0.f;
  const float zi = i < N ? zinv[(long)b * N + i] : 0.f;
  // channel chunks (gridDim.z > 1): sum_j p (db - D) F is linear in db, so every chunk adds its partial dot product and
  // chunk 0 alone subtracts D; the partial results meet in dFG with atomics (zeroed by the launcher)
  const float Di = (i < N && blockIdx.z == 0) ? Dvec[(long)b * N + i] : 0.f;
  for (int j0 = 0; j0 < N; j0 += TT) {
    const int cnt = min(TT, N - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < TT * D; e += TR) {
      const int jj = e / D, k = e - jj * D;
      Fs[jj][k] = jj < cnt ? __ldg(fg + (long)(j0 + jj) * 2 * D + k) : 0.f;
    }
    for (int e = threadIdx.x; e < TT * C / 4; e += TR) {
      const int jj = e / (C / 4), c4 = e - jj * (C / 4);
      reinterpret_cast<float4*>(&Os[jj][0])[c4] =
          jj < cnt ? __ldg(reinterpret_cast<const float4*>(go + (long)(j0 + jj) * CT) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
#pragma unroll 2
    for (int jj = 0; jj < TT; ++jj) {
      if (jj >= cnt) break;
      const float p = __expf(dotD<D>(g, Fs[jj]) - m) * zi;
      float db = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 o = reinterpret_cast<const float4*>(&Os[jj][0])[c4];
        db = fmaf(h[4 * c4], o.x, db); db = fmaf(h[4 * c4 + 1], o.y, db);
        db = fmaf(h[4 * c4 + 2], o.z, db); db = fmaf(h[4 * c4 + 3], o.w, db);
      }
      const float ds = p * (db - Di);
#pragma unroll
      for (int k = 0; k < D; ++k) acc[k] = fmaf(ds, Fs[jj][k], acc[k]);
    }
  }
  if (i < N) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
      if (gridDim.z == 1) dFG[((long)b * N + i) * 2 * D + D + k] = acc[k];
      else atomicAdd(dFG + ((long)b * N + i) * 2 * D + D + k, acc[k]);
    }
  }
}

// ------------------------------------------------------------------ backward: dF (j-stationary)
template <int D, int C>
__global__ void __launch_bounds__(TR) attn_bwd_f_kernel(const float* __restrict__ FG, const float* __restrict__ Hh,
                                                         const float* __restrict__ dO, const float* __restrict__ mrow,
                                                         const float* __restrict__ zinv, const float* __restrict__ Dvec, int N,
                                                         int CT, float* __restrict__ dFG) {
  __shared__ __align__(16) float Gs[TT][D];
  __shared__ __align__(16) float Hs[TT][C];
  __shared__ float Ms[TT], Zs[TT], Ds[TT];
  const int b = blockIdx.y;
  const int j = blockIdx.x * TR + threadIdx.x;
  const float* fg = FG + (long)b * N * 2 * D;
  const float* hh = Hh + (long)b * N * CT + blockIdx.z * C;
  float f[D], o[C], acc[D];
#pragma unroll
  for (int k = 0; k < D; ++k) { f[k] = j < N ? __ldg(fg + (long)j * 2 * D + k) : 0.f; acc[k] = 0.f; }
#pragma unroll
  for (int c = 0; c < C; ++c) o[c] = j < N ? __ldg(dO + ((long)b * N + j) * CT + blockIdx.z * C + c) : 0.f;
  for (int i0 = 0; i0 < N; i0 += TT) {
    const int cnt = min(TT, N - i0);
    __syncthreads();
    for (int e = threadIdx.x; e < TT * D; e += TR) {
      const int ii = e / D, k = e - ii * D;
      Gs[ii][k] = ii < cnt ? __ldg(fg + (long)(i0 + ii) * 2 * D + D + k) : 0.f;
    }
    for (int e = threadIdx.x; e < TT * C / 4; e += TR) {
      const int ii = e / (C / 4), c4 = e - ii * (C / 4);
      reinterpret_cast<float4*>(&Hs[ii][0])[c4] =
          ii < cnt ? __ldg(reinterpret_cast<const float4*>(hh + (long)(i0 + ii) * CT) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x < TT) {
      const int ii = threadIdx.x;
      Ms[ii] = ii < cnt ? mrow[(long)b * N + i0 + ii] : 0.f;
      Zs[ii] = ii < cnt ? zinv[(long)b * N + i0 + ii] : 0.f;
      Ds[ii] = (ii < cnt && blockIdx.z == 0) ? Dvec[(long)b * N + i0 + ii] : 0.f;   // see attn_bwd_g_kernel
    }
    __syncthreads();
#pragma unroll 2
    for (int ii = 0; ii < TT; ++ii) {
      if (ii >= cnt) break;
      const float p = __expf(dotD<D>(Gs[ii], f) - Ms[ii]) * Zs[ii];
      float db = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 h = reinterpret_cast<const float4*>(&Hs[ii][0])[c4];
        db = fmaf(h.x, o[4 * c4], db); db = fmaf(h.y, o[4 * c4 + 1], db);
        db = fmaf(h.z, o[4 * c4 + 2], db); db = fmaf(h.w, o[4 * c4 + 3], db);
      }
      const float ds = p * (db - Ds[ii]);
#pragma unroll
      for (int k = 0; k < D; ++k) acc[k] = fmaf(ds, Gs[ii][k], acc[k]);
    }
  }
  if (j < N) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
      if (gridDim.z == 1) dFG[((long)b * N + j) * 2 * D + k] = acc[k];
      else atomicAdd(dFG + ((long)b * N + j) * 2 * D + k, acc[k]);
    }
  }
}

// ------------------------------------------------------------------ host
bool attn_supported(int C, int d) { return (C == 32 || C == 64 || C == 128 || C == 256) && d == C / 8; }

// channel chunk per CTA: the per-thread accumulators / operand rows hold at most 64 channels
template <int D, int CC>
static int attn_fwd_t(const float* FG, const float* Hh, const float* X, int B, int N, int C, float* mrow, float* zinv, float* out,
                      cudaStream_t st) {
  dim3 grid(cdiv(N, TR), B), gridc(cdiv(N, TR), B, C / CC);
  attn_stats_kernel<D><<<grid, TR, 0, st>>>(FG, N, mrow, zinv);
  attn_out_kernel<D, CC><<<gridc, TR, 0, st>>>(FG, Hh, X, mrow, zinv, N, C, out);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

template <int D, int CC>
static int attn_bwd_t(const float* FG, const float* Hh, const float* dO, const float* mrow, const float* zinv, int B, int N, int C,
                      float* Dvec, float* dFG, float* dHh, cudaStream_t st) {
  dim3 gridc(cdiv(N, TR), B, C / CC);
  if (C / CC > 1) {
    MSAU_CUDA_TRY(cudaMemsetAsync(Dvec, 0, sizeof(float) * B * N, st));
    MSAU_CUDA_TRY(cudaMemsetAsync(dFG, 0, sizeof(float) * B * N * 2 * D, st));
  }
  attn_bwd_h_kernel<D, CC><<<gridc, TR, 0, st>>>(FG, Hh, dO, mrow, zinv, N, C, dHh, Dvec);
  attn_bwd_g_kernel<D, CC><<<gridc, TR, 0, st>>>(FG, Hh, dO, mrow, zinv, Dvec, N, C, dFG);
  attn_bwd_f_kernel<D, CC><<<gridc, TR, 0, st>>>(FG, Hh, dO, mrow, zinv, Dvec, N, C, dFG);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

int launch_attn_fwd(const float* FG, const float* Hh, const float* X, int B, int N, int C, int d, float* mrow, float* zinv,
                    float* out, cudaStream_t st) {
  // two sweeps over the N x N relation map: stats (d MACs) + output (d + C MACs); 1 exp per entry per sweep
  ProfScope ps("attn_fwd_kernels", 2.0 * B * (double)N * N * (2 * d + C), (double)B * N * (2 * d + 3 * C + 2) * 4.0, st);
  if (!attn_supported(C, d)) {
    set_error("attention: unsupported (C=%d, d=%d); supported C in {32,64,128,256} with d = C/8", C, d);
    return MSAU_ERR_UNSUPPORTED;
  }
  switch (C) {
    case 32: return attn_fwd_t<4, 32>(FG, Hh, X, B, N, C, mrow, zinv, out, st);
    case 64: return attn_fwd_t<8, 64>(FG, Hh, X, B, N, C, mrow, zinv, out, st);
    case 128: return attn_fwd_t<16, 64>(FG, Hh, X, B, N, C, mrow, zinv, out, st);
    default: return attn_fwd_t<32, 64>(FG, Hh, X, B, N, C, mrow, zinv, out, st);
  }
}

int launch_attn_bwd(const float* FG, const float* Hh, const float* dO, const float* mrow, const float* zinv, int B, int N, int C,
                    int d, float* Dvec, float* dFG, float* dHh, cudaStream_t st) {
  ProfScope ps("attn_bwd_kernels", 2.0 * B * (double)N * N * (5 * d + 3 * C), (double)B * N * (4 * d + 3 * C + 3) * 4.0, st);
  if (!attn_supported(C, d)) {
    set_error("attention: unsupported (C=%d, d=%d); supported C in {32,64,128,256} with d = C/8", C, d);
    return MSAU_ERR_UNSUPPORTED;
  }
  switch (C) {
    case 32: return attn_bwd_t<4, 32>(FG, Hh, dO, mrow, zinv, B, N, C, Dvec, dFG, dHh, st);
    case 64: return attn_bwd_t<8, 64>(FG, Hh, dO, mrow, zinv, B, N, C, Dvec, dFG, dHh, st);
    case 128: return attn_bwd_t<16, 64>(FG, Hh, dO, mrow, zinv, B, N, C, Dvec, dFG, dHh, st);
    default: return attn_bwd_t<32, 64>(FG, Hh, dO, mrow, zinv, B, N, C, Dvec, dFG, dHh, st);
  }
}

}  // namespace msau
