#pragma once
#include "common.cuh"
namespace msau {
// Structured (one-hot) first layer, see first_layer.cu.  `flag` is a device int: 0 after the scan iff the input is one-hot.
int launch_onehot_scan(const float* x, int layout_nchw, int C, int pitch, int B, int H, int W, short* ids, int* flag, cudaStream_t st);
int launch_first_fwd(const short* ids, const int* flag, const float* w, const float* bias, int cin, int cinp, int B, int H, int W,
                     float* out, int po, cudaStream_t st);
int launch_first_wgrad(const short* ids, const int* flag, const float* dz, int pdz, int cin, int cout, int B, int H, int W, float* dW,
                       float* dbias, cudaStream_t st);
}  // namespace msau
