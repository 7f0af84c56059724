#pragma once
#include "common.cuh"
namespace msau {
// One window of a packed-weight array: dst[ty][tx][I][O] (I = rows / contraction channels, O = columns /
// produced channels, both possibly zero-padded; the packed buffer is zero-filled before every pack).
// The window [i_dst0, i_dst0+I_log) x [o_dst0, o_dst0+O_log) is gathered from a reference-layout
// parameter tensor: src[src_off + (i+i_off)*s_i + (o+o_off)*s_o + ky*KW + kx], (ky,kx) = (ky0+kys*ty, kx0+kxs*tx).
struct PackDesc {
  long dst_off, src_off;
  int TH, TW, I, O, i_dst0, o_dst0, I_log, O_log, i_off, o_off;
  long s_i, s_o;
  int ky0, kys, kx0, kxs, KW;
  int TWd, ty_d0, tx_d0;   // destination tap grid: width and origin of this descriptor's taps inside it
  long blk0;
};
int launch_pack(const float* params, float* packed, const PackDesc* d_descs, int n_desc, long total_blocks, cudaStream_t st);
int launch_clip_adam(float* params, float* grads, float* m, float* v, long n, int step, float lr, float b1, float b2, float eps,
                     float max_norm, float* partial /*>= 1024 floats*/, float* total_norm_out, cudaStream_t st);
// kind 0 Adam / 1 RMSprop (b1 = alpha) / 2 SGD with momentum (b1 = momentum); step_dev: optional device step counter (incremented by the
// call; for CUDA-graph capture), else `step` (1-based) from the host; max_norm <= 0: no gradient clipping
int launch_optim(int kind, float* params, float* grads, float* m, float* v, long n, int step, int* step_dev, float lr, float b1, float b2,
                 float eps, float wd, float max_norm, float* partial, float* total_norm_out, cudaStream_t st);
}  // namespace msau
