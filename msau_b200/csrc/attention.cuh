#pragma once
#include "common.cuh"
namespace msau {
bool attn_supported(int C, int d);
int launch_attn_fwd(const float* FG, const float* Hh, const float* X, int B, int N, int C, int d, float* mrow, float* zinv,
                    float* out, cudaStream_t st);
int launch_attn_bwd(const float* FG, const float* Hh, const float* dO, const float* mrow, const float* zinv, int B, int N, int C,
                    int d, float* Dvec, float* dFG, float* dHh, cudaStream_t st);
}  // namespace msau
