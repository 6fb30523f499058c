// Fused forward of the factorised 3x3 convolutions on the 5th-generation tensor cores (bf16 tcgen05, fp32 TMEM
// accumulators):  1x1 (C_in -> r_a)  ->  3x3 (r_a -> r_b, stride 1 or 2, zero padding 1)  ->  1x1 (r_b -> C_out) + bias,
// ONE kernel, both intermediates stay in shared memory as bf16.
//
// This is the contraction of TTConv2dM.forward (TTConv.py:130-153: in-core chain, F.conv2d with core_kernel, out-core
// chain; the host folds the in / out chains into one matrix each) and of TKConv2dC/M.forward (TKConv.py:93-98, 205-222:
// first factor, core, last factor).  csrc/ttconv_fused.cu is the fp32 CUDA-core version of the same contraction (and
// still serves strides, kernel sizes and paddings this kernel does not).
//
// Pixel-major implicit GEMM without im2col.  The batch is laid out as one flat sequence of zero-PADDED grids,
// position g = b * Hp * Wp + (y + 1) * Wp + (x + 1) with Hp = H + 2, Wp = W + 2.  In that space the tap (ky, kx) of the
// 3x3 stage is a constant offset (ky - 1) * Wp + (kx - 1), and the pad positions -- whose stage-1 output is exactly
// zero because stage 1 has no bias and reads zero -- isolate rows and images from each other.  A CTA takes a chunk of
// T x 128 positions plus a halo of Wp + 1 positions on either side:
//   load    x (NCHW fp32) -> bf16, eight channels = 16 bytes per position and channel group, "planes" [group][position]
//   stage 1 Z1[pos, ra] = X[pos, :] . A_in[ra, :]          tcgen05.mma M = 128 positions, N = r_a, K = C_in
//   stage 2 Z2[pos, rb] = sum_taps Z1[pos + tap, :] . K_tap[rb, :]     nine MMAs groups on ROW-SHIFTED views of Z1
//   stage 3 Y[pos, co]  = Z2[pos, :] . A_out[co, :] + bias  -> NCHW fp32, pad positions dropped
// The operands use the un-swizzled K-major canonical layout (8 rows x 16 bytes core matrices stored contiguously,
// SBO = 128 bytes): rows are 16 bytes apart inside a plane, so a view shifted by any number of positions is just a
// different descriptor start address -- this is what makes the shifted-window trick possible without copies.
// HBM traffic = x (+ halo) + y; every weight is read from L2 once per CTA.
#include <cstdlib>

#include "tc_common.cuh"

namespace tta {

constexpr int kTcThreads = 256;      // pack kernel

struct TcConvDesc {
  int B, Cin, H, W, Ra, Rb, Cout;
  int stride, Ho, Wo;             // stride 2: the stride-1 result is formed at every position, the odd ones are dropped
  int Cinp, Rap, Rbp, Coutp;      // padded to multiples of 16 (MMA K step / N granularity)
  int Wp, Gp, halo;               // padded row, padded grid size, Wp + 1
  int T, TE;                      // output tiles per chunk, stage-1 tiles per chunk (chunk + both halos)
  long long total;                // B * Gp
  // shared-memory byte offsets
  int o_wa, o_wk, o_wo, o_bias, o_x, o_z1, smem;
  int wbytes;                     // size of the weight image (= offset of the activation planes)
  int xbytes;                     // one X buffer (two are laid out)
  int nb;                         // columns of one TMEM accumulator: 32 or 64 (two are allocated)
};

// un-swizzled K-major operand: core matrices (8 rows x 16 B) contiguous, `lbo` bytes between the two 16-byte K chunks
// of one MMA, 128 bytes between 8-row groups
__device__ __forceinline__ uint64_t tcv_desc(uint32_t addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint32_t tcv_pack(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ void tcv_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// weight matrix w[n][k] (n < N, k < K, element at w + n*sn + k*sk) -> planes [k / 8][Np rows][16 bytes], zero padded
__device__ __forceinline__ void tcv_pack_weight(uint8_t* dst, const float* __restrict__ w, int N, int K, int Np, int Kp,
                                                int64_t sn, int64_t sk, int tid) {
  const int groups = Kp >> 3;
  for (int it = tid; it < groups * Np; it += kTcThreads) {
    const int n = it % Np, kg = it / Np;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kg * 8 + e;
      v[e] = (n < N && k < K) ? __ldg(w + (int64_t)n * sn + (int64_t)k * sk) : 0.f;
    }
    *reinterpret_cast<uint4*>(dst + (size_t)(kg * Np + n) * 16) =
        make_uint4(tcv_pack(v[0], v[1]), tcv_pack(v[2], v[3]), tcv_pack(v[4], v[5]), tcv_pack(v[6], v[7]));
  }
}

// Weights -> the shared-memory image of the forward kernel (bf16 planes of A_in, the nine K_tap, A_out; fp32 bias): done
// once per weight change by the host (fwd_common.FoldedConv), so that a CTA fetches its weights with ONE bulk copy.
__global__ void __launch_bounds__(kTcThreads) ttconv_tc_pack_kernel(const float* __restrict__ a_in, const float* __restrict__ kern,
                                                                   const float* __restrict__ a_out,
                                                                   const float* __restrict__ bias, uint8_t* __restrict__ blob,
                                                                   const __grid_constant__ TcConvDesc d) {
  const int tid = threadIdx.x;
  tcv_pack_weight(blob + d.o_wa, a_in, d.Ra, d.Cin, d.Rap, d.Cinp, d.Cin, 1, tid);
  for (int tap = 0; tap < 9; ++tap)
    tcv_pack_weight(blob + d.o_wk + (size_t)tap * (size_t)(d.Rap * d.Rbp * 2), kern + tap, d.Rb, d.Ra, d.Rbp, d.Rap,
                    (int64_t)d.Ra * 9, 9, tid);
  tcv_pack_weight(blob + d.o_wo, a_out, d.Cout, d.Rb, d.Coutp, d.Rbp, d.Rb, 1, tid);
  float* bias_b = reinterpret_cast<float*>(blob + d.o_bias);
  for (int c = tid; c < d.Coutp; c += kTcThreads) bias_b[c] = (bias && c < d.Cout) ? __ldg(bias + c) : 0.f;
}

// Persistent, warp-specialised: a CTA fetches the weight image once and walks over chunks blockIdx.x, blockIdx.x + gridDim.x, ...
//   warps 0-7   loaders: x of chunk i + 1 -> bf16 planes in the other X buffer while chunk i is being multiplied
//   warp 8      tcgen05.mma issuer (stage 1 -> stage 2 -> stage 3 of a chunk, two TMEM accumulators used alternately)
//   warps 9-12  drain: TMEM -> bf16 -> Z1 / Z2 planes, and TMEM + bias -> y (warp & 3 selects the TMEM lane quadrant)
constexpr int kTcvLoaders = 256;
constexpr int kTcvThreads = kTcvLoaders + 32 + 128;

__global__ void __launch_bounds__(kTcvThreads, 2) ttconv_tc_kernel(const float* __restrict__ x, const uint8_t* __restrict__ blob,
                                                                  float* __restrict__ y, const __grid_constant__ TcConvDesc d) {
  extern __shared__ __align__(128) uint8_t tcv_smem_raw[];
  const uint32_t smem = (tc::smem_u32(tcv_smem_raw) + 127u) & ~127u;
  __shared__ uint64_t bars[11];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t bar0 = tc::smem_u32(bars);
  const uint32_t afull0 = bar0, aempty0 = bar0 + 16, xfull0 = bar0 + 32, xempty0 = bar0 + 48, z1done = bar0 + 64, z2done = bar0 + 72,
                 wbar = bar0 + 80;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s2 = 0; s2 < 2; ++s2) {
      tc::mbar_init(afull0 + 8 * s2, 1);
      tc::mbar_init(aempty0 + 8 * s2, 4);
      tc::mbar_init(xfull0 + 8 * s2, kTcvLoaders / 32);
      tc::mbar_init(xempty0 + 8 * s2, 1);
    }
    tc::mbar_init(z1done, 4);
    tc::mbar_init(z2done, 4);
    tc::mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the weight image: one bulk copy per CTA, lands while the first chunk of activations is being converted
    tc::mbar_expect_tx(wbar, (uint32_t)d.wbytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem), "l"(blob),
                 "r"((uint32_t)d.wbytes), "r"(wbar)
                 : "memory");
  }
  if (warp == kTcvLoaders / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)(2 * d.nb))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  const uint32_t wa = smem + d.o_wa, wk = smem + d.o_wk, wo = smem + d.o_wo, z1 = smem + d.o_z1;
  const float* bias_s = reinterpret_cast<const float*>(tcv_smem_raw + ((smem - tc::smem_u32(tcv_smem_raw)) + d.o_bias));
  const int NP = d.TE * 128;                                 // rows of the X and Z1 planes
  const int PE = d.T * 128 + 2 * d.halo;                     // rows of them that are ever read by stage 2
  const uint32_t plx = (uint32_t)NP * 16u, plz2 = (uint32_t)d.T * 128u * 16u;
  const uint32_t xbytes = (uint32_t)d.xbytes;                // one X buffer (Z2 of the same chunk aliases it)
  const long long nchunks = (d.total + (long long)d.T * 128 - 1) / ((long long)d.T * 128);

  if (warp < kTcvLoaders / 32) {
    // =============================== loaders ===============================
    const int groups = d.Cinp >> 3;
    const int64_t hw = (int64_t)d.H * d.W;
    int it_ = 0;
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it_) {
      const int xb = it_ & 1;
      if (it_ >= 2) tc::mbar_wait(xempty0 + 8 * xb, (uint32_t)(((it_ >> 1) - 1) & 1));
      const uint32_t xs = smem + d.o_x + (uint32_t)xb * xbytes;
      const int e0 = (int)(ch * d.T * 128) - d.halo;           // position of row 0 of the planes
      // a thread owns two rows (positions) per pass: the position is decoded once, then every channel group of both rows
      // is fetched with 16 independent loads in flight before the first conversion
      const bool full8 = (d.Cin & 7) == 0;
      for (int r0 = tid; r0 < PE; r0 += 2 * kTcvLoaders) {
        const float* src[2];
        int rowi[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int row = r0 + u * kTcvLoaders;
          rowi[u] = row;
          src[u] = nullptr;
          const int g = e0 + row;
          if (row < PE && g >= 0 && g < (int)d.total) {
            const int b = g / d.Gp;
            const int rem = g - b * d.Gp;
            const int py = rem / d.Wp, px = rem - py * d.Wp;
            if (py >= 1 && py <= d.H && px >= 1 && px <= d.W)
              src[u] = x + (int64_t)b * d.Cin * hw + (int64_t)(py - 1) * d.W + (px - 1);
          }
        }
        for (int kg = 0; kg < groups; ++kg) {
          float v[2][8];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float* p = src[u] + (int64_t)kg * 8 * hw;
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[u][e] = (src[u] != nullptr && (full8 || kg * 8 + e < d.Cin)) ? __ldg(p + e * hw) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (rowi[u] < PE)
              tc::sts128(xs + (uint32_t)kg * plx + (uint32_t)rowi[u] * 16u, tcv_pack(v[u][0], v[u][1]), tcv_pack(v[u][2], v[u][3]),
                         tcv_pack(v[u][4], v[u][5]), tcv_pack(v[u][6], v[u][7]));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy (UMMA)
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(xfull0 + 8 * xb);
    }
  } else if (warp == kTcvLoaders / 32) {
    // =============================== MMA issuer ===============================
    tc::mbar_wait(wbar, 0);                                             // the weight image has landed
    const uint32_t id1 = tc::umma_idesc_bf16(d.Rap), id2 = tc::umma_idesc_bf16(d.Rbp), id3 = tc::umma_idesc_bf16(d.Coutp);
    const uint32_t pla = (uint32_t)d.Rap * 16u, plk = (uint32_t)d.Rbp * 16u, plo = (uint32_t)d.Coutp * 16u;
    const uint32_t tapb = (uint32_t)(d.Rap * d.Rbp * 2);
    int tile_ctr = 0, it_ = 0;
    auto acquire = [&]() {      // accumulator of the next tile
      const int buf = tile_ctr & 1;
      if (tile_ctr >= 2) tc::mbar_wait(aempty0 + 8 * buf, (uint32_t)(((tile_ctr >> 1) - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      return buf;
    };
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it_) {
      const int xb = it_ & 1;
      const uint32_t xs = smem + d.o_x + (uint32_t)xb * xbytes, z2 = xs;
      tc::mbar_wait(xfull0 + 8 * xb, (uint32_t)((it_ >> 1) & 1));
      // ---- stage 1: Z1 = X . A_in^T over the chunk and its halos ----
      for (int e = 0; e < d.TE; ++e, ++tile_ctr) {
        const int buf = acquire();
        if (tc::elect_one()) {
          for (int k2 = 0; k2 < (d.Cinp >> 4); ++k2)
            tc::umma_bf16(tmem_base + (uint32_t)(buf * d.nb), tcv_desc(xs + (uint32_t)(2 * k2) * plx + (uint32_t)e * 2048u, plx),
                          tcv_desc(wa + (uint32_t)(2 * k2) * pla, pla), id1, k2 ? 1u : 0u);
          tc::umma_commit(afull0 + 8 * buf);
        }
        __syncwarp();
      }
      // ---- stage 2: Z2 = sum over taps of shifted Z1 . K_tap^T ----
      tc::mbar_wait(z1done, (uint32_t)(it_ & 1));
      for (int t = 0; t < d.T; ++t, ++tile_ctr) {
        const int buf = acquire();
        if (tc::elect_one()) {
          uint32_t acc = 0;
          for (int tap = 0; tap < 9; ++tap) {
            const int shift = (tap / 3 - 1) * d.Wp + (tap % 3 - 1);
            const uint32_t arow = (uint32_t)(t * 128 + d.halo + shift) * 16u;
            for (int k2 = 0; k2 < (d.Rap >> 4); ++k2) {
              tc::umma_bf16(tmem_base + (uint32_t)(buf * d.nb), tcv_desc(z1 + (uint32_t)(2 * k2) * plx + arow, plx),
                            tcv_desc(wk + (uint32_t)tap * tapb + (uint32_t)(2 * k2) * plk, plk), id2, acc);
              acc = 1;
            }
          }
          tc::umma_commit(afull0 + 8 * buf);
        }
        __syncwarp();
      }
      // ---- stage 3: Y = Z2 . A_out^T ----
      tc::mbar_wait(z2done, (uint32_t)(it_ & 1));
      for (int t = 0; t < d.T; ++t, ++tile_ctr) {
        const int buf = acquire();
        if (tc::elect_one()) {
          for (int k2 = 0; k2 < (d.Rbp >> 4); ++k2)
            tc::umma_bf16(tmem_base + (uint32_t)(buf * d.nb), tcv_desc(z2 + (uint32_t)(2 * k2) * plz2 + (uint32_t)t * 2048u, plz2),
                          tcv_desc(wo + (uint32_t)(2 * k2) * plo, plo), id3, k2 ? 1u : 0u);
          tc::umma_commit(afull0 + 8 * buf);
          if (t == d.T - 1) tc::umma_commit(xempty0 + 8 * xb);        // X / Z2 of this chunk may be overwritten
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== drain ===============================
    const int quad = warp & 3;                                          // warps 9..12 -> quadrants 1, 2, 3, 0
    const int64_t hw = (int64_t)d.Ho * d.Wo;                            // output plane
    tc::mbar_wait(wbar, 0);                                             // bias
    int tile_ctr = 0, it_ = 0;
    auto pack_rows = [&](uint32_t dst_plane0, uint32_t plane_bytes, int row, int ngroups16, int buf) {
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * d.nb);
      for (int cg = 0; cg < ngroups16; ++cg) {
        uint32_t v[16];
        tcv_ld16(taddr + (uint32_t)cg * 16u, v);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tc::sts128(dst_plane0 + (uint32_t)(2 * cg + h) * plane_bytes + (uint32_t)row * 16u,
                     tcv_pack(__uint_as_float(v[8 * h]), __uint_as_float(v[8 * h + 1])),
                     tcv_pack(__uint_as_float(v[8 * h + 2]), __uint_as_float(v[8 * h + 3])),
                     tcv_pack(__uint_as_float(v[8 * h + 4]), __uint_as_float(v[8 * h + 5])),
                     tcv_pack(__uint_as_float(v[8 * h + 6]), __uint_as_float(v[8 * h + 7])));
      }
    };
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it_) {
      const int xb = it_ & 1;
      const uint32_t z2 = smem + d.o_x + (uint32_t)xb * xbytes;
      const int c0 = (int)(ch * d.T * 128);
      for (int e = 0; e < d.TE; ++e, ++tile_ctr) {                      // stage 1 -> Z1
        const int buf = tile_ctr & 1;
        tc::mbar_wait(afull0 + 8 * buf, (uint32_t)((tile_ctr >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        pack_rows(z1, plx, e * 128 + quad * 32 + lane, d.Rap >> 4, buf);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (e == d.TE - 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(aempty0 + 8 * buf);
          if (e == d.TE - 1) tc::mbar_arrive(z1done);
        }
      }
      for (int t = 0; t < d.T; ++t, ++tile_ctr) {                       // stage 2 -> Z2
        const int buf = tile_ctr & 1;
        tc::mbar_wait(afull0 + 8 * buf, (uint32_t)((tile_ctr >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        pack_rows(z2, plz2, t * 128 + quad * 32 + lane, d.Rbp >> 4, buf);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (t == d.T - 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(aempty0 + 8 * buf);
          if (t == d.T - 1) tc::mbar_arrive(z2done);
        }
      }
      for (int t = 0; t < d.T; ++t, ++tile_ctr) {                       // stage 3 -> y
        const int buf = tile_ctr & 1;
        tc::mbar_wait(afull0 + 8 * buf, (uint32_t)((tile_ctr >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * d.nb);
        const int g = c0 + t * 128 + quad * 32 + lane;
        bool valid = g < (int)d.total;
        float* yp = y;
        if (valid) {
          const int b = g / d.Gp;
          const int rem = g - b * d.Gp;
          const int py = rem / d.Wp, px = rem - py * d.Wp;
          valid = py >= 1 && py <= d.H && px >= 1 && px <= d.W;
          int oy = py - 1, ox = px - 1;
          if (d.stride == 2) {
            valid = valid && !(oy & 1) && !(ox & 1);
            oy >>= 1;
            ox >>= 1;
          }
          yp = y + (int64_t)b * d.Cout * hw + (int64_t)oy * d.Wo + ox;
        }
        for (int cg = 0; cg < (d.Coutp >> 4); ++cg) {
          uint32_t v[16];
          tcv_ld16(taddr + (uint32_t)cg * 16u, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = cg * 16 + j;
              if (co < d.Cout) yp[(int64_t)co * hw] = __uint_as_float(v[j]) + bias_s[co];
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(aempty0 + 8 * buf);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kTcvLoaders / 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * d.nb)) : "memory");
  }
}

static int tcv_round16(int v) { return (v + 15) & ~15; }

// geometry this kernel serves: 3x3, stride 1 or 2, padding 1, channel counts and ranks <= 64
bool ttconv_tc_supported(int Cin, int Ra, int Rb, int Cout, int KS, int stride, int pad) {
  return KS == 3 && (stride == 1 || stride == 2) && pad == 1 && Cin <= 64 && Ra <= 64 && Rb <= 64 && Cout <= 64;
}

static void tcv_layout(TcConvDesc& d, int T) {
  d.T = T;
  d.TE = (T * 128 + 2 * d.halo + 127) / 128;
  int o = 0;
  d.o_wa = o; o += d.Cinp * d.Rap * 2;
  d.o_wk = o; o += 9 * d.Rap * d.Rbp * 2;
  d.o_wo = o; o += d.Rbp * d.Coutp * 2;
  d.o_bias = o; o += d.Coutp * 4;
  o = (o + 127) & ~127;
  d.wbytes = o;
  const int xb = d.Cinp * d.TE * 128 * 2, z2b = d.Rbp * T * 128 * 2;
  d.xbytes = xb > z2b ? xb : z2b;
  d.o_x = o; o += 2 * d.xbytes;
  d.o_z1 = o; o += d.Rap * d.TE * 128 * 2;
  d.smem = o + 128;
}

}  // namespace tta

namespace tta {
static int tcv_prepare(TcConvDesc& d, int B, int Cin, int H, int W, int Ra, int Rb, int Cout, int stride = 1) {
  d.stride = stride;
  d.Ho = (H + 2 - 3) / stride + 1;
  d.Wo = (W + 2 - 3) / stride + 1;
  d.B = B; d.Cin = Cin; d.H = H; d.W = W; d.Ra = Ra; d.Rb = Rb; d.Cout = Cout;
  d.Cinp = tcv_round16(Cin); d.Rap = tcv_round16(Ra); d.Rbp = tcv_round16(Rb); d.Coutp = tcv_round16(Cout);
  d.Wp = W + 2;
  d.Gp = (H + 2) * d.Wp;
  d.halo = d.Wp + 1;
  d.total = (long long)B * d.Gp;
  d.nb = (d.Rap > 32 || d.Rbp > 32 || d.Coutp > 32) ? 64 : 32;
  // chunk size: as many 128-position tiles per CTA as fit in shared memory (fewer halo positions are recomputed), as long
  // as every SM still gets a chunk
  // chunk size: tiles per chunk.  Small chunks keep the stage chain of a chunk short (the loads of the next chunk overlap
  // it); every chunk recomputes its two halos in stage 1.  TTA_TTCONV_T overrides (measurements).
  static const int forced = [] { const char* e = getenv("TTA_TTCONV_T"); return e ? atoi(e) : 0; }();
  int T = forced > 0 ? forced : 2;
  for (; T > 1; T >>= 1) {
    tcv_layout(d, T);
    if (d.smem <= 220 * 1024) break;
  }
  tcv_layout(d, T);
  return d.smem <= 227 * 1024 ? TTA_OK : TTA_E_INVALID;
}
}  // namespace tta

extern "C" int tta_ttconv_tc_supported(int Cin, int Ra, int Rb, int Cout, int KS, int stride, int pad) {
  return tta::ttconv_tc_supported(Cin, Ra, Rb, Cout, KS, stride, pad) ? 1 : 0;
}

extern "C" int64_t tta_ttconv_tc_blob_bytes(int Cin, int Ra, int Rb, int Cout) {
  tta::TcConvDesc d;
  tta::tcv_prepare(d, 1, Cin, 8, 8, Ra, Rb, Cout);
  return d.wbytes;
}

extern "C" int tta_ttconv_tc_pack(const float* a_in, const float* kern, const float* a_out, const float* bias, void* blob,
                                  int Cin, int Ra, int Rb, int Cout, void* stream) {
  using namespace tta;
  if (!a_in || !kern || !a_out || !blob || !ttconv_tc_supported(Cin, Ra, Rb, Cout, 3, 1, 1)) {
    set_error("ttconv_tc_pack: bad argument");
    return TTA_E_INVALID;
  }
  TcConvDesc d;
  tcv_prepare(d, 1, Cin, 8, 8, Ra, Rb, Cout);
  ttconv_tc_pack_kernel<<<1, kTcThreads, 0, (cudaStream_t)stream>>>(a_in, kern, a_out, bias, reinterpret_cast<uint8_t*>(blob), d);
  TTA_CHECK_LAUNCH("ttconv_tc_pack launch");
  return TTA_OK;
}

extern "C" int tta_ttconv_tc_fwd(const float* x, const void* blob, float* y, int B, int Cin, int H, int W, int Ra, int Rb,
                                 int Cout, int KS, int stride, int pad, void* stream) {
  using namespace tta;
  if (!x || !blob || !y || B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || Ra <= 0 || Rb <= 0 || Cout <= 0) {
    set_error("ttconv_tc: bad argument");
    return TTA_E_INVALID;
  }
  if (!ttconv_tc_supported(Cin, Ra, Rb, Cout, KS, stride, pad)) {
    set_error("ttconv_tc: unsupported geometry (kernel %d, stride %d, pad %d, channels %d/%d/%d/%d): use tta_ttconv_fused_fwd",
              KS, stride, pad, Cin, Ra, Rb, Cout);
    return TTA_E_INVALID;
  }
  TcConvDesc d;
  if (tcv_prepare(d, B, Cin, H, W, Ra, Rb, Cout, stride) != TTA_OK) {
    set_error("ttconv_tc: working set %d B does not fit shared memory", d.smem);
    return TTA_E_INVALID;
  }
  if (d.total >= (1ll << 31) - 4096) {
    set_error("ttconv_tc: %lld padded positions exceed the 32-bit position space", d.total);
    return TTA_E_INVALID;
  }
  const long long chunks = (d.total + (long long)d.T * 128 - 1) / ((long long)d.T * 128);
  static int smem_set = 0;
  if (d.smem > 48 * 1024 && d.smem > smem_set) {
    int rc = check_cuda(cudaFuncSetAttribute(ttconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem),
                        "ttconv_tc smem attribute");
    if (rc) return rc;
    smem_set = d.smem;
  }
  // persistent: as many CTAs as fit at once (shared memory, 512 TMEM columns, 2048 threads per SM), each walks over chunks
  int per_sm = (227 * 1024) / (d.smem + 1024);
  if (per_sm > 512 / (2 * d.nb)) per_sm = 512 / (2 * d.nb);
  if (per_sm > 2) per_sm = 2;                      // registers: __launch_bounds__(416, 2)
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)kNumSMs * per_sm;
  if (grid > chunks) grid = chunks;
  ttconv_tc_kernel<<<(unsigned)grid, kTcvThreads, d.smem, (cudaStream_t)stream>>>(x, reinterpret_cast<const uint8_t*>(blob), y, d);
  TTA_CHECK_LAUNCH("ttconv_tc launch");
  return TTA_OK;
}
