// Micro-benchmark of the rotate phase of jacobi_gra_kernel (eig_gra_device.cuh): clocks per rotation step for
// one CTA of 512 threads, and for stripped variants that show where the time goes.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../include -I../../dnn-compression-tensor-admm_b200/csrc
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#include "eig_gra_device.cuh"

namespace tta {
void set_error(const char*, ...) {}
void count_launch() {}
}

using namespace tta;

// MODE 0: full step; 1: no V threads; 2: no pivot record (reuses stale records); 3: barrier only
template <int MODE>
__global__ void __launch_bounds__(512, 1) rotate_kernel(const float* c_init, long long* cycles, int rounds, float* dump) {
  __shared__ __align__(16) float Cb[2 * kGraCsz];
  __shared__ __align__(16) float Vm[32 * kGraVs];
  __shared__ float4 rec[32];
  int fin = 0;
  const int tid = threadIdx.x;
  int nrot = 0;
  float maxrel2 = 0.f;
  long long total = 0;
  for (int r = 0; r < rounds; ++r) {
    for (int e = tid; e < 1024; e += 512) {
      Cb[(e >> 5) * kGraCs + (e & 31)] = c_init[e];
      Vm[(e >> 5) * kGraVs + (e & 31)] = ((e >> 5) == (e & 31)) ? 1.f : 0.f;
    }
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 0) {
      fin = gra_rotate_cross(Cb, Vm, rec, tid, 2.5e-13f, 1e-30f, nrot, maxrel2);
    } else {
      if (tid < 16) rec[tid] = gra_record(Cb[tid * kGraCs + tid], Cb[(16 + tid) * kGraCs + 16 + tid], Cb[tid * kGraCs + 16 + tid], 2.5e-13f, 1e-30f, nrot, maxrel2);
      __syncthreads();
      int cur = 0;
#pragma unroll 1
      for (int s = 0; s < 16; ++s) {
        const float* Cc = Cb + cur * kGraCsz;
        float* Cn = Cb + (cur ^ 1) * kGraCsz;
        const float4* rc = rec + (MODE == 2 ? 0 : cur * 16);
        if (MODE != 3) {
          if (tid < 256) {
            const int dd = tid >> 4, a = tid & 15, b = (a + dd) & 15;
            const int qa = 16 + ((a + s) & 15), qb = 16 + ((b + s) & 15);
            const float4 ra = rc[a], rb = rc[b];
            const float n01 = gra_block(Cc, Cn, a, qa, b, qb, ra, rb, dd == 0);
            if (MODE != 2 && dd == 1 && s < 15) rec[(cur ^ 1) * 16 + a] = gra_record(ra.z, rb.w, n01, 2.5e-13f, 1e-30f, nrot, maxrel2);
          } else if (MODE != 1) {
            const int u = tid - 256, b = u & 15;
            gra_vrot(Vm, (u >> 4) * 2, b, 16 + ((b + s) & 15), rc[b]);
          }
        }
        cur ^= 1;
        __syncthreads();
      }
    }
    total += clock64() - t0;
  }
  if (tid == 0) cycles[0] = total + (nrot & 1) * 0 + (maxrel2 > 1e30f);
  if (dump) {
    __syncthreads();
    for (int e = tid; e < 1024; e += 512) {
      dump[e] = Cb[fin * kGraCsz + (e >> 5) * kGraCs + (e & 31)];
      dump[1024 + e] = Vm[(e >> 5) * kGraVs + (e & 31)];
    }
  }
}

int main() {
  std::vector<float> c(1024);
  // a symmetric positive definite 32 x 32 matrix with sizeable off-diagonals
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) c[i * 32 + j] = (i == j ? 40.f + i : 0.f) + 1.f / (1 + abs(i - j)) + 0.01f * ((i * 7 + j * 3) % 5);
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < i; ++j) c[i * 32 + j] = c[j * 32 + i];
  float* dc; long long* dt;
  cudaMalloc(&dc, 4096); cudaMalloc(&dt, 8);
  cudaMemcpy(dc, c.data(), 4096, cudaMemcpyHostToDevice);
  const int rounds = 2000;
  const char* names[4] = {"full step (kernel code)", "no V threads", "no pivot records", "barrier only"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) rotate_kernel<0><<<1, 512>>>(dc, dt, rounds, nullptr);
      if (mode == 1) rotate_kernel<1><<<1, 512>>>(dc, dt, rounds, nullptr);
      if (mode == 2) rotate_kernel<2><<<1, 512>>>(dc, dt, rounds, nullptr);
      if (mode == 3) rotate_kernel<3><<<1, 512>>>(dc, dt, rounds, nullptr);
      cudaDeviceSynchronize();
    }
    long long t; cudaMemcpy(&t, dt, 8, cudaMemcpyDeviceToHost);
    printf("%-26s %7.1f clk per rotation step (%s)\n", names[mode], (double)t / rounds / 16, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
