// Activation layout kernels of the decomposed conv forwards (TTConv.py:132,137,142,149; TKConv.py:205-222):
// the TT / Tucker chains contract the channel index, so activations are kept pixel-major (NHWC, bf16);
// the k x k core convolution (TTConv.py:139, TKConv.py:95) is an implicit GEMM over im2col rows.
#include <cuda_bf16.h>

#include "tta_common.cuh"

namespace tta {

// x (B, C, HW) fp32  ->  y (B, HW, ldc) bf16 (first C columns written)
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          int C, int HW, int ldc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* xb = x + (int64_t)b * C * HW;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < C && p < HW) ? xb[(int64_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* yb = y + (int64_t)b * HW * ldc;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < HW && c < C) yb[(int64_t)p * ldc + c] = __float2bfloat16(tile[tx][i]);
  }
}

// x (B, HW, ldc) bf16/fp32 -> y (B, C, HW) fp32 (+ bias[c])
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y,
                                                          const float* __restrict__ bias, int C, int HW, int ldc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* xb = x + (int64_t)b * HW * ldc;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) v = (float)xb[(int64_t)p * ldc + c];
    tile[i][tx] = v;
  }
  __syncthreads();
  float* yb = y + (int64_t)b * C * HW;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (c < C && p < HW) yb[(int64_t)c * HW + p] = tile[tx][i] + (bias ? __ldg(bias + c) : 0.f);
  }
}

// im2col: x (B, H, W, ldx) bf16 -> rows (B*Ho*Wo) x ldo, column (kh*KW + kw)*C + c ; zero padding
__global__ void __launch_bounds__(256) im2col_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                    int B, int H, int W, int C, int ldx, int KH, int KW, int sh, int sw,
                                                    int ph, int pw, int dh, int dw, int Ho, int Wo, int ldo) {
  const int64_t rows = (int64_t)B * Ho * Wo;
  const int kcols = KH * KW * C;
  const int64_t total = rows * ldo;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ldo;
    const int col = (int)(e - r * ldo);
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (col < kcols) {
      const int tap = col / C, c = col - tap * C;
      const int kh = tap / KW, kw = tap - kh * KW;
      const int wo = (int)(r % Wo);
      const int64_t t = r / Wo;
      const int ho = (int)(t % Ho);
      const int b = (int)(t / Ho);
      const int hi = ho * sh - ph + kh * dh, wi = wo * sw - pw + kw * dw;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = x[(((int64_t)b * H + hi) * W + wi) * ldx + c];
    }
    out[e] = v;
  }
}

}  // namespace tta

extern "C" {

int tta_nchw_to_nhwc_bf16(const float* x, void* y, int B, int C, int HW, int ldc, void* stream) {
  using namespace tta;
  if (B <= 0 || C <= 0 || HW <= 0) return TTA_OK;
  if (!x || !y || ldc < C || B > 65535) {
    set_error("nchw_to_nhwc: bad argument");
    return TTA_E_INVALID;
  }
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), C, HW, ldc);
  TTA_CHECK_LAUNCH("nchw_to_nhwc launch");
  return TTA_OK;
}

int tta_nhwc_to_nchw_f32(const void* x, int x_is_f32, float* y, const float* bias, int B, int C, int HW, int ldc,
                         void* stream) {
  using namespace tta;
  if (B <= 0 || C <= 0 || HW <= 0) return TTA_OK;
  if (!x || !y || ldc < C || B > 65535) {
    set_error("nhwc_to_nchw: bad argument");
    return TTA_E_INVALID;
  }
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  if (x_is_f32)
    nhwc_to_nchw_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(x), y, bias, C, HW, ldc);
  else
    nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), y,
                                                                              bias, C, HW, ldc);
  TTA_CHECK_LAUNCH("nhwc_to_nchw launch");
  return TTA_OK;
}

int tta_im2col_bf16(const void* x, void* out, int B, int H, int W, int C, int ldx, int KH, int KW, int sh, int sw, int ph,
                    int pw, int dh, int dw, int Ho, int Wo, int ldo, void* stream) {
  using namespace tta;
  if (B <= 0) return TTA_OK;
  if (!x || !out || ldo < KH * KW * C || ldx < C || Ho <= 0 || Wo <= 0) {
    set_error("im2col: bad argument");
    return TTA_E_INVALID;
  }
  const int64_t total = (int64_t)B * Ho * Wo * ldo;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  im2col_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C, ldx, KH,
                                                                KW, sh, sw, ph, pw, dh, dw, Ho, Wo, ldo);
  TTA_CHECK_LAUNCH("im2col launch");
  return TTA_OK;
}
}
