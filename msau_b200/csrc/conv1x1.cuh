#pragma once
#include "common.cuh"
namespace msau {
bool conv1x1_supported(const ConvArgs& a);
int launch_conv1x1(const ConvArgs& a, cudaStream_t st);
}  // namespace msau
