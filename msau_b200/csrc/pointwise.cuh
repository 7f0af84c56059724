#pragma once
#include "common.cuh"
namespace msau {
// coop: engine option "lrn_coop" (0 = thread-per-pixel kernels only, 1 = lane-cooperative where it wins, 2 = from 8 channels up)
int launch_lrn_fwd(const float* z, float* y, long npix, int C, int coop, cudaStream_t st);
int launch_lrn_bwd(const float* z, const float* gy, float* gz, long npix, int C, int coop, cudaStream_t st);
int launch_pool_fwd(const float* x, float* y, int B, int H, int W, int C, cudaStream_t st);
int launch_pool_bwd(const float* x, const float* gy, float* gx, int B, int H, int W, int C, int accumulate, int relu_mask, cudaStream_t st);
int launch_relu_mask(float* g, const float* y, long n, cudaStream_t st);
int launch_add(float* dst, const float* src, long n, int accumulate, cudaStream_t st);
int launch_head(const float* lg, int P, int n_class, int B, long npix_per_page, float* logits_nchw, float* probs_nchw, uint8_t* argmax, cudaStream_t st);
int loss_partial_count(int B, long npix_per_page);
long loss_scratch_ints(int B);
int launch_colsum(const float* g, long npix, int C, int c_lim, float* out, cudaStream_t st);
// mode 0: MSAUWrapper.loss (model/model.py:446-459); mode 1: UNetLoss (model/training/cost.py:35-65)
struct LossSpec {
  int mode = 0;
  float w_main = 1.f, w_aux = 1.f;
  const float* class_weights = nullptr;   // host, n_class floats, or null
};
// ints: >= loss_scratch_ints(B) ints = class histogram [2][B][32] | correct | kept;  flags: 1 int, bit 0 = label out of range (sticky)
int launch_loss(const float* lg, const float* la, int P, int n_class, const void* labels, const void* labels_aux, int label_dtype, int B,
                long npix_per_page, float gscale, const LossSpec& spec, float* dlg, float* dla, int* ints, int* flags, float* partial /* 2 x loss_partial_count floats */,
                float* loss_out, float* loss_main_out /* or null */, cudaStream_t st);
int launch_onehot_argmax(const void* t, int dtype, int B, int C, long npix_per_page, int channels_last, uint8_t* out, cudaStream_t st);
int launch_confusion(const uint8_t* pred, const void* labels, int label_dtype, long n, int n_class, long long* conf, cudaStream_t st);
}  // namespace msau
