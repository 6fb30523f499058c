// Micro-benchmark: legacy warp-level mma.sync rates on sm_100a (TF32 m16n8k8, BF16 m16n8k16).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void mma_tf32(float* out, int iters) {
  float c[8][4];
  unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f000000u, 0x3f000000u}, b[2] = {0x3f800000u, 0x3e800000u};
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void mma_bf16(float* out, int iters) {
  float c[8][4];
  unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f003f00u, 0x3f003f00u}, b[2] = {0x3f803f80u, 0x3e803e80u};
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  const int iters = 20000;
  float* d; cudaMalloc(&d, (size_t)148 * 8 * 1024 * 4);
  for (int threads : {128, 256, 512}) {
    const int blocks = 148 * (1024 / threads);
    float ms = time_it([&] { mma_tf32<<<blocks, threads>>>(d, iters); });
    const double warps = (double)blocks * threads / 32;
    printf("threads/CTA %4d: mma.sync tf32 m16n8k8  %8.1f TFLOP/s  (%.2f clk per mma per SM at 1.965 GHz)\n", threads,
           2.0 * 16 * 8 * 8 * 8 * iters * warps / ms * 1e-9, ms * 1e-3 * 1.965e9 / (8.0 * iters * warps / 148));
    ms = time_it([&] { mma_bf16<<<blocks, threads>>>(d, iters); });
    printf("threads/CTA %4d: mma.sync bf16 m16n8k16 %8.1f TFLOP/s\n", threads, 2.0 * 16 * 8 * 16 * 8 * iters * warps / ms * 1e-9);
  }
  return 0;
}
