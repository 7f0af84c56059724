// Micro-benchmark: steady-state cycles per tcgen05.mma (kind::f16, cta_group::1) for the operand shapes / layouts the
// MSAU kernels use, with the issue loop kept trivially cheap (unrolled, constant descriptor increments), optionally with
// other warps hammering shared memory (the producers' STS traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_bench scripts/mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk(uint32_t a, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
struct Cfg { int M, N, amaj, bmaj; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int iters; int sts_warps; int same_acc; int issuers; };

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}

__global__ void k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tb;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f803f80u ^ ((i * 2654435761u) & 0x007f007fu);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tb)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    stop = 0;
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < c.issuers) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.amaj << 15) | ((uint32_t)c.bmaj << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint64_t ad = mk(s32(smem), c.a_lbo, c.a_sbo);
    const uint64_t bd = mk(s32(smem) + 100 * 1024, c.b_lbo, c.b_sbo);
    const uint32_t dstep = c.same_acc ? 0 : (uint32_t)c.N;
    long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) mma(tb + (c.same_acc ? warp * c.N : 0) + j * dstep, ad + (uint64_t)(j * 8), bd + (uint64_t)((i & 7) * 64), idesc, i > 0 ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[warp])) : "memory");
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar[warp])), "r"(0) : "memory");
    }
    long long t2 = clock64();
    stop = 1;
    if (blockIdx.x == 0 && warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp >= 4 && warp < 4 + c.sts_warps) {
    // background shared-memory writers (16-B stores, conflict-free), like the producers' image stores
    uint4* dst = reinterpret_cast<uint4*>(smem + 120 * 1024) + threadIdx.x;
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    while (!stop) {
#pragma unroll
      for (int r = 0; r < 8; ++r) dst[r * 256] = v;
      v.x++;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long* d; cudaMalloc(&d, 16);
  struct { const char* name; Cfg c; } tests[] = {
    {"conv  M128 N16, 1 issuer",                 {128, 16, 0, 0, 19008, 1056, 256, 128, 2000, 0, 1, 1}},
    {"conv  M128 N16, 2 issuers",                {128, 16, 0, 0, 19008, 1056, 256, 128, 2000, 0, 1, 2}},
    {"conv  M128 N16, 4 issuers",                {128, 16, 0, 0, 19008, 1056, 256, 128, 2000, 0, 1, 4}},
    {"conv  M128 N16, 4 issuers + 4 STS warps",  {128, 16, 0, 0, 19008, 1056, 256, 128, 2000, 4, 1, 4}},
    {"conv  M128 N64, 2 issuers",                {128, 64, 0, 0, 19008, 1056, 1024, 128, 2000, 0, 1, 2}},
    {"conv  M128 N64, 4 issuers",                {128, 64, 0, 0, 19008, 1056, 1024, 128, 2000, 0, 1, 4}},
    {"wgrad M64  N8  MN-major, 1 issuer",        {64, 8, 1, 1, 128, 16, 128, 16384, 2000, 0, 1, 1}},
    {"wgrad M64  N8  MN-major, 2 issuers",       {64, 8, 1, 1, 128, 16, 128, 16384, 2000, 0, 1, 2}},
    {"wgrad M64  N8  MN-major, 4 issuers",       {64, 8, 1, 1, 128, 16, 128, 16384, 2000, 0, 1, 4}},
    {"wgrad M128 N8  MN-major, 4 issuers",       {128, 8, 1, 1, 128, 16, 128, 16384, 2000, 0, 1, 4}},
    {"wgrad M64  N64 MN-major, 4 issuers",       {64, 64, 1, 1, 128, 16, 128, 2048, 2000, 0, 1, 4}},
  };
  for (auto& t : tests) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 256, 180 * 1024>>>(t.c, d);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<148, 256, 180 * 1024>>>(t.c, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double n = t.c.iters * 8.0;
    printf("%-44s issue %6.1f total %6.1f cyc/mma per issuer -> %5.1f cyc per mma per SM ; kernel %.1f us (%s)\n", t.name, (double)h[0] / n, (double)h[1] / n, (double)h[1] / n / t.c.issuers, ms * 1e3,
           cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
