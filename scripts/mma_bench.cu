// Micro-benchmark: cycles per tcgen05.mma (kind::f16, cta_group::1) as a function of the instruction shape and
// of the shared-memory operand layout (no-swizzle descriptors with various LBO/SBO/alignment).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk(uint32_t a, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
struct Cfg { int M, N, amaj, bmaj; uint32_t a_off, a_lbo, a_sbo, b_lbo, b_sbo, layout; int n_mma; int step16; int smem_kb; int b_off_kb; int d_rot; int d_stride; int fill; int b_rot; };

__global__ void k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  for (int i = threadIdx.x; i < c.smem_kb * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = c.fill ? (0x3f803f80u ^ ((i * 2654435761u) & 0x007f007fu)) : 0;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tb)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.amaj << 15) | ((uint32_t)c.bmaj << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    uint64_t ad = mk(s32(smem) + c.a_off, c.a_lbo, c.a_sbo, c.layout);
    uint64_t bd = mk(s32(smem) + c.b_off_kb * 1024, c.b_lbo, c.b_sbo, c.layout);
    long long t0 = clock64();
    for (int i = 0; i < c.n_mma; ++i) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb + (uint32_t)((i % c.d_rot) * c.d_stride)),
                   "l"(ad + (uint64_t)((i & 7) * c.step16)), "l"(bd + (uint64_t)((i / 8 % c.b_rot) * 32)), "r"(idesc), "r"(i >= c.d_rot ? 1u : 0u)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
    }
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128));
}

int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long* d; cudaMalloc(&d, 16);
  struct { const char* name; Cfg c; } tests[] = {
    {"conv-like N16 zeros, 1 acc",         {128, 16, 0, 0, 0, 19008, 1056, 256, 128, 0, 512, 8, 44, 40, 1, 16, 0, 1}},
    {"conv-like N16 data,  1 acc",         {128, 16, 0, 0, 0, 19008, 1056, 256, 128, 0, 512, 8, 44, 40, 1, 16, 1, 1}},
    {"conv-like N16 data,  8 acc rot",     {128, 16, 0, 0, 0, 19008, 1056, 256, 128, 0, 512, 8, 44, 40, 8, 16, 1, 1}},
    {"conv-like N16 data,  8 acc, B rot",  {128, 16, 0, 0, 0, 19008, 1056, 256, 128, 0, 512, 8, 44, 40, 8, 16, 1, 14}},
    {"conv-like N16 data,  1 acc, B rot",  {128, 16, 0, 0, 0, 19008, 1056, 256, 128, 0, 512, 8, 44, 40, 1, 16, 1, 14}},
  };
  for (auto& t : tests) {
    for (int per_sm = 1; per_sm <= 1; ++per_sm) {
      const int ctas = 148 * per_sm;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k<<<ctas, 128, t.c.smem_kb * 1024>>>(t.c, d);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      k<<<ctas, 128, t.c.smem_kb * 1024>>>(t.c, d);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long h[2] = {0, 0};
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-36s CTAs/SM %d: issue %6.1f cyc/mma  total %6.1f cyc/mma ; kernel %.1f us -> %.1f cyc per mma per SM (%s)\n", t.name, per_sm,
             (double)h[0] / t.c.n_mma, (double)h[1] / t.c.n_mma, ms * 1e3, ms * 1e-3 * 1.9e9 / (t.c.n_mma * per_sm), cudaGetErrorString(e));
    }
  }
  return 0;
}
