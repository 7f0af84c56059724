#pragma once
#include "common.cuh"
namespace msau {
extern int g_lrn_coop;   // engine option "lrn_coop"
int launch_lrn_fwd(const float* z, float* y, long npix, int C, cudaStream_t st);
int launch_lrn_bwd(const float* z, const float* gy, float* gz, long npix, int C, cudaStream_t st);
int launch_pool_fwd(const float* x, float* y, int B, int H, int W, int C, cudaStream_t st);
int launch_pool_bwd(const float* x, const float* gy, float* gx, int B, int H, int W, int C, int accumulate, int relu_mask, cudaStream_t st);
int launch_relu_mask(float* g, const float* y, long n, cudaStream_t st);
int launch_add(float* dst, const float* src, long n, int accumulate, cudaStream_t st);
int launch_head(const float* lg, int P, int n_class, int B, long npix_per_page, float* logits_nchw, float* probs_nchw, uint8_t* argmax, cudaStream_t st);
int loss_partial_count(int B, long npix_per_page);
int launch_colsum(const float* g, long npix, int C, int c_lim, float* out, cudaStream_t st);
int launch_loss(const float* lg, const float* la, int n_class, const void* labels, int label_is_i64, int B, long npix_per_page,
                float gscale, float* dlg, float* dla, int* counts, float* partial, float* loss_out, cudaStream_t st);
}  // namespace msau
