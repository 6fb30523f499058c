// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (128 x N x 8) and kind::f16 (128 x N x 16) from shared-memory
// operands (K-major, SWIZZLE_128B), one CTA per SM, one issuing thread, all MMAs into one TMEM accumulator.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a umma_tf32_rate.cu -o umma_tf32_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int KIND>   // 0: tf32, 1: bf16
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* cycles, int iters, int n, int distinct) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t smem = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  for (uint32_t o = threadIdx.x * 16; o < 65536; o += 128 * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(smem + o), "r"(0x3f800000u) : "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = KIND == 0 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24))
                                     : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24));
    const uint64_t da = desc_sw128(smem), db = desc_sw128(smem + 16384);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t a = da + (uint64_t)(2 * k) + (distinct ? (uint64_t)((it & 1) * (32768 >> 4)) : 0);
        const uint64_t b = db + (uint64_t)(2 * k) + (distinct ? (uint64_t)((it & 1) * (32768 >> 4)) : 0);
        if (KIND == 0)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 65536 * 2 + 1024;
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int kind = 0; kind < 2; ++kind)
    for (int n : {32, 64, 128, 256})
      for (int ctas : {1, 148}) {
        if (kind == 0) rate_kernel<0><<<ctas, 128, smem>>>(d, iters, n, 0);
        else rate_kernel<1><<<ctas, 128, smem>>>(d, iters, n, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        const double per = (double)h[0] / (iters * 4.0);
        const double kk = kind == 0 ? 8.0 : 16.0;
        printf("%s 128 x %3d x %2d  ctas %3d: %7.1f cycles / MMA  => %.0f FLOP/clk/SM  (%s)\n", kind == 0 ? "tf32" : "bf16", n, (int)kk,
               ctas, per, 2.0 * 128 * n * kk / per, cudaGetErrorString(e));
      }
  return 0;
}
