// Micro-benchmark: DFMA (CUDA cores), DMMA (mma.sync.m8n8k4.f64) and FFMA2 issue rates on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
  const double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2];
  double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma2_kernel(float* out, int iters) {
  float2 a[16];
  const float2 x = make_float2(1.0000001f, 0.9999f), y = make_float2(0.5f, 0.25f);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = make_float2(i + threadIdx.x, i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("{ .reg .b64 ra, rb, rc;\n mov.b64 ra, {%0,%1};\n mov.b64 rb, {%2,%3};\n mov.b64 rc, {%4,%5};\n"
                   " fma.rn.f32x2 ra, ra, rb, rc;\n mov.b64 {%0,%1}, ra; }"
                   : "+f"(a[i].x), "+f"(a[i].y) : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
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
  const int blocks = 148 * 4, threads = 512, iters = 20000;
  double* d; cudaMalloc(&d, (size_t)blocks * threads * 8);
  float ms = time_it([&] { dfma_kernel<<<blocks, threads>>>(d, iters); });
  printf("DFMA : %.2f TFLOP/s\n", 2.0 * 16 * iters * (double)blocks * threads / ms * 1e-9);
  ms = time_it([&] { dmma_kernel<<<blocks, threads>>>(d, iters); });
  printf("DMMA : %.2f TFLOP/s (m8n8k4)\n", 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * threads / 32 / ms * 1e-9);
  ms = time_it([&] { ffma2_kernel<<<blocks, threads>>>((float*)d, iters); });
  printf("FFMA2: %.2f TFLOP/s\n", 2.0 * 2 * 16 * iters * (double)blocks * threads / ms * 1e-9);
  return 0;
}
