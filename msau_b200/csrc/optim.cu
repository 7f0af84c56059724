// Weight repacking (reference state_dict layout -> kernel layouts) and the fused global-norm clip + Adam.
//
// Reference: train_chargrid_funsd_msau.py:24-26 (Adam lr=1e-4, torch defaults b1=.9 b2=.999 eps=1e-8),
// :58 clip_grad_norm(params, args.clip=True -> max_norm 1.0), :59 optimizer.step().
// torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to <= 1.
// torch.optim.Adam (no amsgrad, no weight decay): m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
// Parameters whose grad is None in the reference (the dead last-block attention) have an all-zero
// gradient here; with zero-initialised moments Adam's update for them is exactly 0, as in torch's skip.
#include "optim.cuh"
#include "prof.cuh"

namespace msau {

__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ params, float* __restrict__ packed,
                                                    const PackDesc* __restrict__ descs, int n_desc) {
  const long blk = blockIdx.x;
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].blk0 <= blk) lo = mid; else hi = mid - 1;
  }
  const PackDesc d = descs[lo];
  const long e = (blk - d.blk0) * 256 + threadIdx.x;
  const long total = (long)d.TH * d.TW * d.I_log * d.O_log;
  if (e >= total) return;
  const int o = (int)(e % d.O_log);
  long r = e / d.O_log;
  const int i = (int)(r % d.I_log); r /= d.I_log;
  const int tx = (int)(r % d.TW);
  const int ty = (int)(r / d.TW);
  const int ky = d.ky0 + d.kys * ty, kx = d.kx0 + d.kxs * tx;
  const float v = params[d.src_off + (long)(i + d.i_off) * d.s_i + (long)(o + d.o_off) * d.s_o + ky * d.KW + kx];
  packed[d.dst_off + ((long)((ty + d.ty_d0) * d.TWd + tx + d.tx_d0) * d.I + d.i_dst0 + i) * d.O + d.o_dst0 + o] = v;
}

int launch_pack(const float* params, float* packed, const PackDesc* d_descs, int n_desc, long total_blocks, cudaStream_t st) {
  ProfScope ps("pack_kernel", 0, (double)total_blocks * 256 * 8.0, st);
  pack_kernel<<<(unsigned)total_blocks, 256, 0, st>>>(params, packed, d_descs, n_desc);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

static constexpr int kNormBlocks = 296;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n, float* __restrict__ partial) {
  double s = 0.0;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const float x = g[i];
    s += (double)x * (double)x;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = (float)sh[0];
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long n, float step_size, float inv_sqrt_bc2, float b1,
                                                    float b2, float eps, float max_norm, const float* __restrict__ partial,
                                                    int n_partial, float* __restrict__ total_norm_out) {
  // every block re-derives the same global norm from the fixed-order partials => identical on all ranks
  __shared__ float coef_s;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_partial; ++i) s += (double)partial[i];
    const float total = (float)sqrt(s);
    float c = max_norm / (total + 1e-6f);
    coef_s = c > 1.f ? 1.f : c;
    if (blockIdx.x == 0 && total_norm_out) *total_norm_out = total;
  }
  __syncthreads();
  const float coef = coef_s;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const float gi = g[i] * coef;
    g[i] = gi;   // clip_grad_norm_ scales .grad in place
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] -= step_size * (mi / denom);
  }
}

int launch_clip_adam(float* params, float* grads, float* m, float* v, long n, int step, float lr, float b1, float b2, float eps,
                     float max_norm, float* partial, float* total_norm_out, cudaStream_t st) {
  MSAU_CHECK_ARG(step >= 1, "adam: step must be >= 1");
  ProfScope ps("clip_adam_kernels", 0, (double)n * 4.0 * 8, st);
  sumsq_kernel<<<kNormBlocks, 256, 0, st>>>(grads, n, partial);
  const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  adam_kernel<<<kNormBlocks, 256, 0, st>>>(params, grads, m, v, n, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), b1, b2, eps, max_norm,
                                           partial, kNormBlocks, total_norm_out);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

}  // namespace msau
