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
  // (32-bit index arithmetic: a layer has < 2^31 weights; the 64-bit divisions were most of this kernel's time)
  const int e = (int)(blk - d.blk0) * 256 + (int)threadIdx.x;
  const int total = d.TH * d.TW * d.I_log * d.O_log;
  if (e >= total) return;
  int r = e / d.O_log;
  const int o = e - r * d.O_log;
  const int r2 = r / d.I_log;
  const int i = r - r2 * d.I_log;
  const int ty = r2 / d.TW;
  const int tx = r2 - ty * d.TW;
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

// block 0 also advances the device-side step counter (graph-captured steps cannot take the step count as an argument)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n, float* __restrict__ partial, int* __restrict__ step_dev) {
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
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = (float)sh[0];
    if (blockIdx.x == 0 && step_dev) *step_dev += 1;
  }
}

// kind 0: torch.optim.Adam            m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// kind 1: torch.optim.RMSprop         v = a v + (1-a) g^2 ; p -= lr * g / (sqrt(v) + eps)          (momentum 0, not centered; a = b1)
// kind 2: torch.optim.SGD(momentum)   buf = mu buf + g (buf = g at step 1) ; p -= lr * buf          (dampening 0; mu = b1)
// weight decay (all kinds, torch semantics): g += wd * p before the moment updates.  max_norm > 0: clip_grad_norm_ first.
__global__ void __launch_bounds__(256) optim_kernel(int kind, float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                     float* __restrict__ v, long n, int step_host, const int* __restrict__ step_dev, float lr,
                                                     float b1, float b2, float eps, float wd, float max_norm,
                                                     const float* __restrict__ partial, int n_partial, float* __restrict__ total_norm_out) {
  // every block re-derives the same global norm from the fixed-order partials => identical on all ranks
  __shared__ float coef_s, step_size_s, inv_sqrt_bc2_s;
  __shared__ int step_s;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_partial; ++i) s += (double)partial[i];
    const float total = (float)sqrt(s);
    float c = max_norm > 0.f ? max_norm / (total + 1e-6f) : 1.f;
    coef_s = c > 1.f ? 1.f : c;
    if (blockIdx.x == 0 && total_norm_out) *total_norm_out = total;
    const int t = step_dev ? *step_dev : step_host;
    step_s = t;
    if (kind == 0) {
      const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
      step_size_s = (float)((double)lr / bc1);
      inv_sqrt_bc2_s = (float)(1.0 / sqrt(bc2));
    } else {
      step_size_s = lr; inv_sqrt_bc2_s = 1.f;
    }
  }
  __syncthreads();
  const float coef = coef_s, step_size = step_size_s, inv_sqrt_bc2 = inv_sqrt_bc2_s;
  const bool first = step_s <= 1;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    float gi = g[i] * coef;
    if (max_norm > 0.f) g[i] = gi;   // clip_grad_norm_ scales .grad in place
    float pi = p[i];
    if (wd != 0.f) gi += wd * pi;
    if (kind == 0) {
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi; v[i] = vi;
      pi -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    } else if (kind == 1) {
      const float vi = b1 * v[i] + (1.f - b1) * gi * gi;
      v[i] = vi;
      pi -= step_size * (gi / (sqrtf(vi) + eps));
    } else {
      const float bi = first ? gi : b1 * m[i] + gi;
      m[i] = bi;
      pi -= step_size * bi;
    }
    p[i] = pi;
  }
}

int launch_optim(int kind, float* params, float* grads, float* m, float* v, long n, int step, int* step_dev, float lr, float b1, float b2,
                 float eps, float wd, float max_norm, float* partial, float* total_norm_out, cudaStream_t st) {
  MSAU_CHECK_ARG(kind >= 0 && kind <= 2, "optimizer: kind must be 0 (Adam), 1 (RMSprop) or 2 (SGD with momentum)");
  MSAU_CHECK_ARG(step_dev || step >= 1, "optimizer: step must be >= 1 (or pass a device step counter)");
  ProfScope ps("clip_adam_kernels", 0, (double)n * 4.0 * (kind == 0 ? 8 : 6), st);
  sumsq_kernel<<<kNormBlocks, 256, 0, st>>>(grads, n, partial, step_dev);
  optim_kernel<<<kNormBlocks, 256, 0, st>>>(kind, params, grads, m, v, n, step, step_dev, lr, b1, b2, eps, wd, max_norm, partial, kNormBlocks,
                                            total_norm_out);
  MSAU_CUDA_TRY(cudaGetLastError());
  return MSAU_OK;
}

int launch_clip_adam(float* params, float* grads, float* m, float* v, long n, int step, float lr, float b1, float b2, float eps,
                     float max_norm, float* partial, float* total_norm_out, cudaStream_t st) {
  return launch_optim(0, params, grads, m, v, n, step, nullptr, lr, b1, b2, eps, 0.f, max_norm, partial, total_norm_out, st);
}

}  // namespace msau
