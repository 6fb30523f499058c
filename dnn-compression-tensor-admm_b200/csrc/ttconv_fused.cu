// Fused forward of the factorised convolutions: 1x1 (C_in -> r_a)  ->  k x k (r_a -> r_b, stride, zero
// padding)  ->  1x1 (r_b -> C_out) + bias, one kernel, intermediates in shared memory.
//
// This is the contraction of TTConv2dM.forward (TTConv.py:130-153: in-core chain, F.conv2d with
// core_kernel, out-core chain) and of TKConv2dC/M.forward (TKConv.py:93-98, 205-222: first factor, core,
// last factor).  For the TT layers the host folds the in-core chain into one (r_a x C_in) matrix and the
// out-core chain into one (C_out x r_b) matrix (weights only, cached): with the channel counts of the
// reference's tables (<= 64) the folded 1x1 maps cost no more MACs than the chains, and the three stages
// then need no intermediate tensor in HBM at all -- the reference's op chain writes and re-reads five of
// them per layer and is launch-bound.
//
// One CTA = one spatial tile of one image: the input patch (tile + halo, all C_in channels) is staged in
// shared memory, stage 1 produces the r_a-channel patch, stage 2 the r_b-channel tile, stage 3 writes the
// NCHW output.  fp32 throughout (4 x 4 register tiles, FFMA); HBM traffic = input + output activations.
#include "tta_common.cuh"

namespace tta {

constexpr int kFcThreads = 256;

struct FcDesc {
  int B, Cin, H, W, Ra, Rb, Cout, Ho, Wo, stride, pad;
  int TH, TW, PH, PW, PP;   // output tile, input patch, padded patch pixel count (multiple of 4)
  int tiles_x, tiles_y;
};

// out[r][p] (R x P, row stride ldo) = W[r][:] (row stride K) . in[:][p] (K x P, row stride ldi);  4 x 4 tiles
__device__ __forceinline__ void fc_pointwise(const float* __restrict__ Wm, const float* __restrict__ in,
                                             float* __restrict__ out, int R, int K, int P, int ldi, int ldo, int tid) {
  const int pt = (P + 3) >> 2, rtl = (R + 3) >> 2;
  for (int t = tid; t < pt * rtl; t += kFcThreads) {
    const int r0 = (t / pt) * 4, p0 = (t % pt) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(in + k * ldi + p0);
      float w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = (r0 + i < R) ? Wm[(r0 + i) * K + k] : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(w[i], v.x, acc[i][0]);
        acc[i][1] = fmaf(w[i], v.y, acc[i][1]);
        acc[i][2] = fmaf(w[i], v.z, acc[i][2]);
        acc[i][3] = fmaf(w[i], v.w, acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r0 + i < R)
        *reinterpret_cast<float4*>(out + (r0 + i) * ldo + p0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

template <int KS>
__global__ void __launch_bounds__(kFcThreads) ttconv_fused_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ a_in,
                                                                  const float* __restrict__ kern,
                                                                  const float* __restrict__ a_out,
                                                                  const float* __restrict__ bias,
                                                                  float* __restrict__ y, const FcDesc d) {
  extern __shared__ __align__(16) float fsm[];
  const int tid = threadIdx.x;
  const int TP = ((d.TH * d.TW) + 3) & ~3;
  float* wA = fsm;                                   // Ra x Cin
  const int Rbp = (d.Rb + 3) & ~3;
  float* wK = wA + ((d.Ra * d.Cin + 3) & ~3);        // [ra][ky][kx][rb], rb padded to a multiple of 4 with zeros
  float* wO = wK + d.Ra * KS * KS * Rbp;             // Cout x Rb
  float* xs = wO + ((d.Cout * d.Rb + 3) & ~3);       // Cin x PP   (later: z2, Rb x TP)
  const int xz = d.Cin * d.PP > d.Rb * TP ? d.Cin * d.PP : d.Rb * TP;
  float* z1 = xs + xz;                               // Ra x PP
  float* z2 = xs;

  for (int e = tid; e < d.Ra * d.Cin; e += kFcThreads) wA[e] = a_in[e];
  for (int e = tid; e < d.Ra * KS * KS * Rbp; e += kFcThreads) {
    const int rb = e % Rbp, t = e / Rbp;             // t = (ra*KS + ky)*KS + kx
    const int ra = t / (KS * KS), kk = t - ra * KS * KS;
    wK[e] = rb < d.Rb ? __ldg(kern + ((int64_t)rb * d.Ra + ra) * KS * KS + kk) : 0.f;
  }
  for (int e = tid; e < d.Cout * d.Rb; e += kFcThreads) wO[e] = a_out[e];

  const int b = blockIdx.y;
  const int ty = blockIdx.x / d.tiles_x, tx = blockIdx.x % d.tiles_x;
  const int oy0 = ty * d.TH, ox0 = tx * d.TW;
  const int iy0 = oy0 * d.stride - d.pad, ix0 = ox0 * d.stride - d.pad;

  // ---- input patch (zero outside the image) ----
  const float* xb = x + (int64_t)b * d.Cin * d.H * d.W;
  for (int e = tid; e < d.Cin * d.PP; e += kFcThreads) {
    const int c = e / d.PP, p = e - c * d.PP;
    const int py = p / d.PW, px = p - py * d.PW;
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (p < d.PH * d.PW && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) v = __ldg(xb + ((int64_t)c * d.H + iy) * d.W + ix);
    xs[e] = v;
  }
  __syncthreads();

  // ---- stage 1: z1 = A_in . x over the whole patch (zero input -> zero output: the padding of the k x k stage) ----
  fc_pointwise(wA, xs, z1, d.Ra, d.Cin, d.PP, d.PP, d.PP, tid);
  __syncthreads();

  // ---- stage 2: z2[rb][oy][ox] = sum_{ra,ky,kx} K[rb][ra][ky][kx] z1[ra][oy*s+ky][ox*s+kx]; 4 rb x 4 ox per thread ----
  {
    const int xt = (d.TW + 3) >> 2, rtl = (d.Rb + 3) >> 2;
    const int ntile = rtl * d.TH * xt;
    const int s = d.stride;
    for (int t = tid; t < ntile; t += kFcThreads) {
      const int r0 = (t / (d.TH * xt)) * 4;
      const int rem = t % (d.TH * xt);
      const int oy = rem / xt, oxl = (rem % xt) * 4;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int ra = 0; ra < d.Ra; ++ra) {
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          // rows of the patch are 16-byte aligned (PW % 4 == 0) and oxl * s is a multiple of 4
          const float4* zr = reinterpret_cast<const float4*>(z1 + ra * d.PP + (oy * s + ky) * d.PW + oxl * s);
          float zv[12];
          {
            const float4 q0 = zr[0], q1 = zr[1];
            zv[0] = q0.x; zv[1] = q0.y; zv[2] = q0.z; zv[3] = q0.w;
            zv[4] = q1.x; zv[5] = q1.y; zv[6] = q1.z; zv[7] = q1.w;
            if (s == 2) {
              const float4 q2 = zr[2];
              zv[8] = q2.x; zv[9] = q2.y; zv[10] = q2.z; zv[11] = q2.w;
            } else {
              zv[8] = zv[9] = zv[10] = zv[11] = 0.f;
            }
          }
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            const float4 w4 = *reinterpret_cast<const float4*>(wK + ((ra * KS + ky) * KS + kx) * Rbp + r0);
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float z = (s == 1) ? zv[j + kx] : zv[2 * j + kx];
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(w[i], z, acc[i][j]);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (r0 + i < d.Rb)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (oxl + j < d.TW) z2[(r0 + i) * TP + oy * d.TW + oxl + j] = acc[i][j];
    }
  }
  __syncthreads();

  // ---- stage 3: y = A_out . z2 + bias, NCHW ----
  {
    const int npix = d.TH * d.TW;
    const int pt = (npix + 3) >> 2, rtl = (d.Cout + 3) >> 2;
    float* yb = y + (int64_t)b * d.Cout * d.Ho * d.Wo;
    for (int t = tid; t < pt * rtl; t += kFcThreads) {
      const int r0 = (t / pt) * 4, p0 = (t % pt) * 4;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int k = 0; k < d.Rb; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(z2 + k * TP + p0);
        float w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = (r0 + i < d.Cout) ? wO[(r0 + i) * d.Rb + k] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][0] = fmaf(w[i], v.x, acc[i][0]);
          acc[i][1] = fmaf(w[i], v.y, acc[i][1]);
          acc[i][2] = fmaf(w[i], v.z, acc[i][2]);
          acc[i][3] = fmaf(w[i], v.w, acc[i][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = r0 + i;
        if (co >= d.Cout) continue;
        const float bv = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = p0 + j;
          if (p >= npix) continue;
          const int oy = oy0 + p / d.TW, ox = ox0 + p % d.TW;
          if (oy < d.Ho && ox < d.Wo) yb[((int64_t)co * d.Ho + oy) * d.Wo + ox] = acc[i][j] + bv;
        }
      }
    }
  }
}

static size_t fc_smem_floats(const FcDesc& d, int KS) {
  const int TP = ((d.TH * d.TW) + 3) & ~3;
  const size_t xz = (size_t)d.Cin * d.PP > (size_t)d.Rb * TP ? (size_t)d.Cin * d.PP : (size_t)d.Rb * TP;
  return (size_t)((d.Ra * d.Cin + 3) & ~3) + (size_t)d.Ra * KS * KS * ((d.Rb + 3) & ~3) + ((d.Cout * d.Rb + 3) & ~3) + xz +
         (size_t)d.Ra * d.PP + 16;
}

}  // namespace tta

extern "C" int tta_ttconv_fused_fwd(const float* x, const float* a_in, const float* kern, const float* a_out,
                                    const float* bias, float* y, int B, int Cin, int H, int W, int Ra, int Rb,
                                    int Cout, int KS, int stride, int pad, void* stream) {
  using namespace tta;
  if (!x || !a_in || !kern || !a_out || !y || B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || Ra <= 0 || Rb <= 0 ||
      Cout <= 0) {
    set_error("ttconv_fused: bad argument");
    return TTA_E_INVALID;
  }
  if ((KS != 1 && KS != 3) || stride < 1 || stride > 2 || pad < 0 || pad > KS) {
    set_error("ttconv_fused: unsupported geometry (kernel %d, stride %d, pad %d)", KS, stride, pad);
    return TTA_E_INVALID;
  }
  FcDesc d;
  d.B = B; d.Cin = Cin; d.H = H; d.W = W; d.Ra = Ra; d.Rb = Rb; d.Cout = Cout;
  d.stride = stride; d.pad = pad;
  d.Ho = (H + 2 * pad - KS) / stride + 1;
  d.Wo = (W + 2 * pad - KS) / stride + 1;
  if (d.Ho <= 0 || d.Wo <= 0) {
    set_error("ttconv_fused: empty output");
    return TTA_E_INVALID;
  }
  // largest tile (<= 16 x 16) whose working set leaves room for two CTAs per SM
  d.TH = d.Ho < 16 ? d.Ho : 16;
  d.TW = d.Wo < 16 ? d.Wo : 16;
  const size_t budget = 100 * 1024;
  for (;;) {
    d.PH = (d.TH - 1) * stride + KS;
    d.PW = ((d.TW - 1) * stride + KS + 3) & ~3;        // patch rows padded to 16 bytes
    d.PP = d.PH * d.PW + 16;                           // slack: stage 2 reads whole float4 groups past a row end
    if (fc_smem_floats(d, KS) * sizeof(float) <= budget || (d.TH == 1 && d.TW <= 4)) break;
    if (d.TH >= d.TW && d.TH > 1) d.TH = (d.TH + 1) / 2; else d.TW = (d.TW + 1) / 2;
  }
  const size_t smem = fc_smem_floats(d, KS) * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("ttconv_fused: working set %zu B does not fit shared memory", smem);
    return TTA_E_INVALID;
  }
  d.tiles_x = (d.Wo + d.TW - 1) / d.TW;
  d.tiles_y = (d.Ho + d.TH - 1) / d.TH;
  cudaStream_t st = (cudaStream_t)stream;
  static size_t smem_set[2] = {0, 0};
  const int which = KS == 3 ? 1 : 0;
  if (smem > 48 * 1024 && smem > smem_set[which]) {
    int rc = which ? check_cuda(cudaFuncSetAttribute(ttconv_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ttconv_fused smem attribute")
                   : check_cuda(cudaFuncSetAttribute(ttconv_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ttconv_fused smem attribute");
    if (rc) return rc;
    smem_set[which] = smem;
  }
  dim3 grid(d.tiles_x * d.tiles_y, B);
  if (KS == 3)
    ttconv_fused_kernel<3><<<grid, kFcThreads, smem, st>>>(x, a_in, kern, a_out, bias, y, d);
  else
    ttconv_fused_kernel<1><<<grid, kFcThreads, smem, st>>>(x, a_in, kern, a_out, bias, y, d);
  TTA_CHECK_LAUNCH("ttconv_fused launch");
  return TTA_OK;
}
