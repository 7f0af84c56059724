#pragma once
#include "common.cuh"
namespace msau {
// Structured (one-hot) first layer, see first_layer.cu.  `flag` is a device int: 0 after the scan iff the input is one-hot.
int launch_onehot_scan(const float* x, int layout_nchw, int C, int pitch, int B, int H, int W, short* ids, int* flag, cudaStream_t st);
int launch_first_fwd(const short* ids, const int* flag, const float* w, const float* bias, int cin, int cinp, int B, int H, int W,
                     float* out, int po, cudaStream_t st);
int launch_first_wgrad(const short* ids, const int* flag, const float* dz, int pdz, int cin, int cout, int B, int H, int W, float* dW,
                       float* dbias, cudaStream_t st);
// Box-constant (BERT-grid) input given as (row-id map, feature table [rows, cin] fp32): P = table x W projection + id-gather
// forward; histogram by row + table^T x hist for the weight gradient.  P: rows*72 floats, hist: rows*72 floats (caller scratch).
int launch_table_first_fwd(const short* ids, const float* table, int rows, int cin, int cinp, const float* w, const float* bias, float* P,
                           int B, int H, int W, float* out, int po, cudaStream_t st);
int launch_table_first_wgrad(const short* ids, const float* table, int rows, int cin, int cout, const float* dz, int pdz, float* hist, int B,
                             int H, int W, float* dW, cudaStream_t st);
}  // namespace msau
