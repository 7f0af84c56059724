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
}  // namespace msau
