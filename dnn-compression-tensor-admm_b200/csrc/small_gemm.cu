// Skinny contractions of the decomposed-layer forwards (TTLinear.py:79-88, TTConv.py:133-147): the
// outer cores of a TT chain contract K = s*r <= 96 inputs into N <= 96 outputs for millions of rows --
// HBM-bound, far too small for a UMMA tile.  One thread owns one row: the (N x K) core sits in shared
// memory as fp32, the row is read with 16-byte loads, the N results go out through a split-row address
// map  dest = (i / m_inner) * s_outer + (i % m_inner) * s_inner + j * s_col  so that the last step of a
// chain can write the layer output in its final (token, out_feature) order, bias included.
#include <cuda_bf16.h>

#include "tta_common.cuh"

namespace tta {

constexpr int kSgMaxK = 96, kSgMaxN = 96;
constexpr int kSgThreads = 256;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ void store_out(T* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kSgThreads) small_gemm_kernel(const TIn* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                                TOut* __restrict__ C, const float* __restrict__ bias,
                                                                int64_t M, int N, int K, int64_t a_inner,
                                                                int64_t a_outer, int64_t m_inner,
                                                                int64_t s_outer, int64_t s_inner, int64_t s_col,
                                                                int64_t bias_inner, int64_t bias_col) {
  __shared__ float Bs[kSgMaxN * kSgMaxK];
  for (int e = threadIdx.x; e < N * K; e += kSgThreads) Bs[e] = __bfloat162float(B[e]);
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kSgThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kSgThreads) {
    float a[kSgMaxK];
    const int64_t ao = i / a_inner;
    const TIn* arow = A + ao * a_outer + (i - ao * a_inner) * K;
#pragma unroll 8
    for (int k = 0; k < K; ++k) a[k] = to_f32<TIn>(arow[k]);
    const int64_t io = i / m_inner, ii = i - io * m_inner;
    TOut* crow = C + io * s_outer + ii * s_inner;
    for (int j = 0; j < N; ++j) {
      const float* b = Bs + j * K;
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < K; ++k) acc = fmaf(a[k], b[k], acc);
      if (bias) acc += __ldg(bias + ii * bias_inner + j * bias_col);
      store_out<TOut>(crow + (int64_t)j * s_col, acc);
    }
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                       int64_t n) {
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}

template <typename TIn, typename TOut>
static int launch_small(const void* a, const void* b, void* c, const float* bias, int64_t M, int N, int K,
                        int64_t a_inner, int64_t a_outer, int64_t m_inner, int64_t s_outer, int64_t s_inner, int64_t s_col, int64_t bias_inner,
                        int64_t bias_col, cudaStream_t st) {
  int64_t blocks = (M + kSgThreads - 1) / kSgThreads;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  small_gemm_kernel<TIn, TOut><<<(int)blocks, kSgThreads, 0, st>>>(
      reinterpret_cast<const TIn*>(a), reinterpret_cast<const __nv_bfloat16*>(b), reinterpret_cast<TOut*>(c), bias, M, N,
      K, a_inner, a_outer, m_inner, s_outer, s_inner, s_col, bias_inner, bias_col);
  TTA_CHECK_LAUNCH("small_gemm launch");
  return TTA_OK;
}

}  // namespace tta

extern "C" {

int tta_small_gemm(const void* a, int a_is_f32, const void* b_bf16, void* c, int c_is_f32, const float* bias, int64_t M,
                   int N, int K, int64_t a_inner, int64_t a_outer, int64_t m_inner, int64_t s_outer, int64_t s_inner,
                   int64_t s_col, int64_t bias_inner, int64_t bias_col, void* stream) {
  using namespace tta;
  if (M <= 0) return TTA_OK;
  if (!a || !b_bf16 || !c || N <= 0 || K <= 0 || N > kSgMaxN || K > kSgMaxK || m_inner <= 0 || a_inner <= 0) {
    set_error("small_gemm: needs 0 < K <= %d, 0 < N <= %d (got N=%d K=%d)", kSgMaxK, kSgMaxN, N, K);
    return TTA_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (a_is_f32 && c_is_f32)
    return launch_small<float, float>(a, b_bf16, c, bias, M, N, K, a_inner, a_outer, m_inner, s_outer, s_inner, s_col, bias_inner, bias_col, st);
  if (a_is_f32)
    return launch_small<float, __nv_bfloat16>(a, b_bf16, c, bias, M, N, K, a_inner, a_outer, m_inner, s_outer, s_inner, s_col, bias_inner, bias_col, st);
  if (c_is_f32)
    return launch_small<__nv_bfloat16, float>(a, b_bf16, c, bias, M, N, K, a_inner, a_outer, m_inner, s_outer, s_inner, s_col, bias_inner, bias_col, st);
  return launch_small<__nv_bfloat16, __nv_bfloat16>(a, b_bf16, c, bias, M, N, K, a_inner, a_outer, m_inner, s_outer, s_inner, s_col, bias_inner, bias_col, st);
}

int tta_cast_bf16(const float* x, void* y, int64_t n, void* stream) {
  using namespace tta;
  if (n <= 0) return TTA_OK;
  if (!x || !y || ((uintptr_t)x & 15) || ((uintptr_t)y & 7)) {
    set_error("cast_bf16: null or misaligned pointer");
    return TTA_E_INVALID;
  }
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  cast_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n);
  TTA_CHECK_LAUNCH("cast_bf16 launch");
  return TTA_OK;
}
}
