// Micro-benchmark: achieved HBM read bandwidth of a [B, H, W, 8] fp32 tensor as a function of the tile shape a CTA reads
// per step (rows x cols pixels of 32 B), with plain LDG.128 pairs, lane-pair mapping, and cp.async.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tile_read_bench scripts/tile_read_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

struct Cfg { int B, H, W, TR, TC, mode, ctas_per_sm; };

__global__ void __launch_bounds__(256) k(const float* __restrict__ x, float* out, Cfg c) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tiles_x = c.W / c.TC, tiles_y = c.H / c.TR;
  const int n_tiles = tiles_x * tiles_y * c.B;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(n_tiles, t0 + per);
  float acc = 0.f;
  const int px = c.TR * c.TC;
  for (int t = t0; t < t1; ++t) {
    const int tx = t % tiles_x, r = t / tiles_x, ty = r % tiles_y, b = r / tiles_y;
    const float* base = x + (((long)b * c.H + ty * c.TR) * c.W + tx * c.TC) * 8;
    if (c.mode == 0) {            // one pixel (2 x 16 B) per thread
      for (int e = threadIdx.x; e < px; e += 256) {
        const int row = e / c.TC, col = e - row * c.TC;
        const float4* p = reinterpret_cast<const float4*>(base + ((long)row * c.W + col) * 8);
        const float4 a = __ldg(p), d = __ldg(p + 1);
        acc += a.x + a.w + d.y + d.z;
      }
    } else if (c.mode == 1) {     // half pixel (16 B) per thread: a warp instruction reads 512 contiguous bytes
      for (int e = threadIdx.x; e < 2 * px; e += 256) {
        const int pe = e >> 1, row = pe / c.TC, col = pe - row * c.TC;
        const float4* p = reinterpret_cast<const float4*>(base + ((long)row * c.W + col) * 8) + (e & 1);
        const float4 a = __ldg(p);
        acc += a.x + a.w;
      }
    } else {                      // cp.async 16 B per thread-half-pixel into shared memory, double buffered
      uint32_t s = (uint32_t)__cvta_generic_to_shared(smem) + (t & 1) * (px * 32);
      for (int e = threadIdx.x; e < 2 * px; e += 256) {
        const int pe = e >> 1, row = pe / c.TC, col = pe - row * c.TC;
        const float* p = base + ((long)row * c.W + col) * 8 + (e & 1) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s + e * 16), "l"(p) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  const int B = 16, H = 512, W = 512;
  const size_t n = (size_t)B * H * W * 8;
  float *x, *out;
  cudaMalloc(&x, n * 4); cudaMalloc(&out, 4);
  cudaMemset(x, 0, n * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int shapes[][2] = {{8, 64}, {4, 128}, {2, 256}, {1, 512}, {16, 64}, {4, 32}, {8, 512}};
  for (int mode = 0; mode < 3; ++mode)
    for (int cps = 1; cps <= 4; cps *= 2)
      for (auto& s : shapes) {
        Cfg c{B, H, W, s[0], s[1], mode, cps};
        if (mode == 2 && (size_t)s[0] * s[1] * 64 > 64 * 1024) continue;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const size_t smem = mode == 2 ? (size_t)s[0] * s[1] * 64 : 0;
        k<<<148 * cps, 256, smem>>>(x, out, c);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) k<<<148 * cps, 256, smem>>>(x, out, c);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("mode %d ctas/SM %d tile %2dx%3d : %7.1f us  %6.0f GB/s (%s)\n", mode, cps, s[0], s[1], ms / 5 * 1e3, n * 4 / (ms / 5 * 1e-3) / 1e9,
               cudaGetErrorString(e));
      }
  return 0;
}
