#pragma once
#include "common.cuh"
namespace msau {
bool attn_supported(int C, int d);
int launch_attn_fwd(const float* FG, const float* Hh, const float* X, int B, int N, int C, int d, float* mrow, float* zinv,
                    float* out, cudaStream_t st);
int launch_attn_bwd(const float* FG, const float* Hh, const float* dO, const float* mrow, const float* zinv, int B, int N, int C,
                    int d, float* Dvec, float* dFG, float* dHh, cudaStream_t st);
// tcgen05 path (attn_tc.cu): `lse` [B,N] = log2-domain row log-sum-exp saved by the forward for the backward;
// `scratch` >= attn_tc_scratch_bytes, 128-byte aligned (operand images, rebuilt by every call)
bool attn_tc_supported(int C, int d);
size_t attn_tc_scratch_bytes(int B, int N, int C);
int launch_attn_tc_fwd(const float* FG, const float* Hh, const float* X, int B, int N, int C, int d, float* lse, float* out,
                       void* scratch, cudaStream_t st);
int launch_attn_tc_bwd(const float* FG, const float* Hh, const float* dO, const float* lse, int B, int N, int C, int d, float* dFG,
                       float* dHh, void* scratch, cudaStream_t st);
}  // namespace msau
