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

// 16-byte row loads (rows are contiguous runs of K elements; alignment is checked on the host)
template <typename TIn, int KMAX>
__device__ __forceinline__ void load_row(const TIn* __restrict__ arow, int K, bool vec, float (&a)[KMAX]) {
  constexpr int kPerVec = 16 / sizeof(TIn);
  if (vec) {
#pragma unroll
    for (int v = 0; v < KMAX / kPerVec; ++v) {
      if (v * kPerVec < K) {
        const uint4 raw = *reinterpret_cast<const uint4*>(arow + v * kPerVec);
        const TIn* e = reinterpret_cast<const TIn*>(&raw);
#pragma unroll
        for (int q = 0; q < kPerVec; ++q) a[v * kPerVec + q] = to_f32<TIn>(e[q]);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) a[k] = k < K ? to_f32<TIn>(arow[k]) : 0.f;
  }
}

// KMAX: compile-time bound of K (row held in registers, loops fully unrolled and guarded).
template <typename TIn, typename TOut, int KMAX>
__global__ void __launch_bounds__(kSgThreads) small_gemm_kernel(const TIn* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                                TOut* __restrict__ C, const float* __restrict__ bias,
                                                                int64_t M, int N, int K, int64_t a_inner,
                                                                int64_t a_outer, int64_t m_inner,
                                                                int64_t s_outer, int64_t s_inner, int64_t s_col,
                                                                int64_t bias_inner, int64_t bias_col, int vec) {
  __shared__ float Bs[kSgMaxN * KMAX];       // row j at Bs + j*KMAX, zero padded to KMAX
  for (int e = threadIdx.x; e < N * KMAX; e += kSgThreads) {
    const int j = e / KMAX, k = e - j * KMAX;
    Bs[e] = k < K ? __bfloat162float(B[j * K + k]) : 0.f;
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kSgThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kSgThreads) {
    float a[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) a[k] = 0.f;
    const int64_t ao = i / a_inner;
    load_row<TIn, KMAX>(A + ao * a_outer + (i - ao * a_inner) * K, K, vec != 0, a);
    const int64_t io = i / m_inner, ii = i - io * m_inner;
    TOut* crow = C + io * s_outer + ii * s_inner;
    for (int j = 0; j < N; ++j) {
      const float4* b4 = reinterpret_cast<const float4*>(Bs + j * KMAX);
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX / 4; ++k) {
        const float4 bv = b4[k];            // warp-uniform address: shared-memory broadcast
        acc0 = fmaf(a[4 * k], bv.x, acc0);
        acc1 = fmaf(a[4 * k + 1], bv.y, acc1);
        acc0 = fmaf(a[4 * k + 2], bv.z, acc0);
        acc1 = fmaf(a[4 * k + 3], bv.w, acc1);
      }
      float acc = acc0 + acc1;
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

template <typename TIn, typename TOut, int KMAX>
static int launch_small_k(const void* a, const void* b, void* c, const float* bias, int64_t M, int N, int K,
                          int64_t a_inner, int64_t a_outer, int64_t m_inner, int64_t s_outer, int64_t s_inner,
                          int64_t s_col, int64_t bias_inner, int64_t bias_col, cudaStream_t st) {
  int64_t blocks = (M + kSgThreads - 1) / kSgThreads;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  // 16-byte loads need every row start aligned: base, row length and the outer stride
  const int per = 16 / (int)sizeof(TIn);
  const int vec = (((uintptr_t)a & 15) == 0) && (K % per == 0) && (a_outer % per == 0);
  small_gemm_kernel<TIn, TOut, KMAX><<<(int)blocks, kSgThreads, 0, st>>>(
      reinterpret_cast<const TIn*>(a), reinterpret_cast<const __nv_bfloat16*>(b), reinterpret_cast<TOut*>(c), bias, M, N,
      K, a_inner, a_outer, m_inner, s_outer, s_inner, s_col, bias_inner, bias_col, vec);
  TTA_CHECK_LAUNCH("small_gemm launch");
  return TTA_OK;
}

template <typename TIn, typename TOut>
static int launch_small(const void* a, const void* b, void* c, const float* bias, int64_t M, int N, int K,
                        int64_t a_inner, int64_t a_outer, int64_t m_inner, int64_t s_outer, int64_t s_inner,
                        int64_t s_col, int64_t bias_inner, int64_t bias_col, cudaStream_t st) {
#define TTA_SG(KM)                                                                                                  \
  return launch_small_k<TIn, TOut, KM>(a, b, c, bias, M, N, K, a_inner, a_outer, m_inner, s_outer, s_inner, s_col, \
                                       bias_inner, bias_col, st)
  if (K <= 8) TTA_SG(8);
  if (K <= 16) TTA_SG(16);
  if (K <= 32) TTA_SG(32);
  if (K <= 48) TTA_SG(48);
  if (K <= 64) TTA_SG(64);
  TTA_SG(96);
#undef TTA_SG
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
