#pragma once
#include "common.cuh"
namespace msau {
struct TcPackDesc {
  long src_off;   // floats, into the fp32 packed buffer: [taps][cin][src_pitch]
  long dst_off;   // bf16 elements, into the tensor-core weight buffer
  int taps, cin, coutp, N;
  long blk0;
  int src_pitch, col0;   // the image covers columns col0 .. col0 + coutp - 1 of the source (256-column layers = two 128-column images)
};
// bf16 elements of one 128-column weight image (second half of a 256-column layer sits this far behind the first)
long tc_half_elems(int taps, int cin);
bool conv_tc_supported(const ConvArgs& a);
int tc_weight_floats_equiv(int taps, int cin, int coutp);
int launch_conv_tc(const ConvArgs& a, const uint16_t* wtc, cudaStream_t st);
int launch_pack_tc(const float* pk, uint16_t* pktc, const TcPackDesc* d_descs, int n_desc, long total_blocks, cudaStream_t st);
// 3x3 / dilation-1 convolutions with the kx taps folded into the MMA N dimension (conv3_tc.cu)
bool conv3_tc_supported(const ConvArgs& a);
// use_tma: raw halo planes by TMA tensor loads (engine option "conv3_tma"; env MSAU_C3_TMA overrides)
int launch_conv3_tc(const ConvArgs& a, const uint16_t* wtc, cudaStream_t st, int use_tma = 1);
int debug_c3_prof(unsigned long long* h16);
int launch_pack_tc3(const float* pk, uint16_t* pktc, const TcPackDesc* d_descs, int n_desc, long total_blocks, cudaStream_t st);
bool wgrad_tc_supported(const WgradArgs& a);
int launch_wgrad_tc(const WgradArgs& a, cudaStream_t st);
}  // namespace msau
