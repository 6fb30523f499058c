// Dependent-chain latencies of the fp64 pipe and friends on B200 (one warp, clock64 around N dependent ops).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 2048;

template <int OP>
__global__ void chain(double a, double b, double* out, long long* cyc) {
  double x = a + threadIdx.x * 1e-9, y = b;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, a);                       // DFMA
    if (OP == 1) x = x + y;                              // DADD
    if (OP == 2) x = x * y;                              // DMUL
    if (OP == 3) x = a / x + b;                          // IEEE division + DADD
    if (OP == 4) x = sqrt(x) + b;                        // IEEE sqrt + DADD
    if (OP == 5) x = rsqrt(x) + b;                       // rsqrt + DADD
    if (OP == 6) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + b; }   // MUFU.RCP64H + DADD
    if (OP == 7) x = x + __shfl_xor_sync(0xffffffffu, x, 1);   // shuffle (2 x 32 bit) + DADD
    if (OP == 8) x = (x < y) ? x + a : x - a;            // DSETP + select + DADD
    if (OP == 9) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = (double)f; }   // cvt round trip + FFMA
    if (OP == 10) x = __drcp_rn(x) + b;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 32 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  chain<OP><<<1, 32>>>(1.0000001, 0.999999, out, cyc);
  chain<OP><<<1, 32>>>(1.0000001, 0.999999, out, cyc);
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s %.1f cycles per iteration\n", name, (double)h / N);
}

// throughput: W warps each running 8 independent DFMA chains
__global__ void tput(double a, double b, double* out, long long* cyc) {
  double x[8];
  for (int q = 0; q < 8; ++q) x[q] = a + q + threadIdx.x * 1e-9;
  const long long t0 = clock64();
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = fma(x[q], b, a);
  const long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < 8; ++q) s += x[q];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  run<0>("DFMA chain");
  run<1>("DADD chain");
  run<2>("DMUL chain");
  run<3>("a / x + b (IEEE div + DADD)");
  run<4>("sqrt(x) + b");
  run<5>("rsqrt(x) + b");
  run<6>("rcp.approx.ftz.f64 + DADD");
  run<7>("shfl_xor(double) + DADD");
  run<8>("DSETP + select + DADD");
  run<9>("f64->f32 FFMA f32->f64");
  run<10>("__drcp_rn + DADD");
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  for (int w : {1, 2, 4, 8, 16, 32}) {
    tput<<<1, 32 * w>>>(1.0000001, 0.999999, out, cyc);
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("DFMA throughput, %2d warps x 8 chains: %.2f cycles per warp-DFMA per SM\n", w, (double)h / (N * 8.0 * w));
  }
  return 0;
}
