// Fused forward of a two-factor (TT-matrix / Tucker / SVD) linear layer on the 5th-generation tensor
// cores:   y[M, N2] = bf16( x[M, K1] * W1[N1, K1]^T ) * W2[N2, N1]^T + bias
// One persistent kernel; the rank-N1 intermediate never leaves the SM.
//
// This is the contraction of TTLinearM.forward (TTLinear.py:75-93): for the 4-core TT-matrices of the
// reference's tables (tt_deit_small_patch16_224_hp.py) the two input-side cores fold into W1 (r2 x in)
// and the two output-side cores into W2 (out x r2) with no more MACs than the four-step chain
// (weights-only, cached by the host), so the layer is  in -> r2 -> out  with r2 = 256 / 320.
// Also TKLinearM (TKLinear.py:60-75: first factor, core, last factor with the core folded into one side).
//
// CTA = 128 rows of x, persistent over row tiles (grid = min(tiles, 148)).  Warp roles (192 threads):
//   warp 0     TMA producer (one elected lane): x k-blocks into a 2-stage ring, W1 / W2 blocks (<= 128 rows x
//              64 k, SWIZZLE_128B) into a ring of 16 KB stages; mbarrier expect_tx / complete_tx.
//   warp 1     TMEM allocator (512 columns) + tcgen05.mma issuer (one elected lane).
//              GEMM 1: acc1[128 x N1] (TMEM columns 0..)  += x-block * W1-block^T
//              GEMM 2: acc2[buf][128 x BN2]               = V * W2-chunk^T, V = bf16(acc1): packed in place over
//                      acc1 in TMEM and used as the TMEM A operand (default), or staged in shared memory in the
//                      UMMA K-major layout (TTA_LR2_TS=0)
//   warps 2-5  epilogue (one TMEM lane quadrant each): acc1 -> bf16 -> V (UMMA K-major SWIZZLE_128B layout,
//              written by hand) ; acc2 chunks -> +bias -> swizzled staging -> TMA store of y (fp32 or bf16).  acc2 is double buffered, so
//              the chunk epilogue overlaps the next chunk's MMAs, and GEMM 1 of the next row tile overlaps the
//              last chunk epilogues of this one.
// TMEM budget: acc1 = round_up(N1, 32) columns, acc2 = 2 x BN2 with BN2 = min(128, (512 - acc1) / 2 rounded
// down to 32)  =>  N1 <= 384.
#include <stdlib.h>

#include "tc_common.cuh"

namespace tta {

namespace lr2 {

using namespace tta::tc;

constexpr int kXStages = 2;
constexpr int kMaxWStages = 12;
constexpr int kThreads = 192;
constexpr int kMaxN1 = 384;


struct Params {
  int M, K1, N1, N2;
  int nk1;        // k-blocks of GEMM 1
  int nb1;        // 128-row blocks of W1
  int n1p16;      // N1 rounded up to 16 (MMA N granularity)
  int nkv;        // k-blocks of GEMM 2 = ceil(N1 / 64)
  int bn2;        // GEMM-2 chunk width
  int nchunks;    // ceil(N2 / bn2)
  int acc2_col;   // first TMEM column of acc2
  int wstages;
  int ntiles;
  int out_f32;
  int ts;         // GEMM 2 takes its A operand (V, bf16) from TMEM instead of shared memory
  int bias_vec;   // bias is 16-byte aligned: 128-bit loads
  int64_t ldy;
};

// Barrier slots (8 bytes each) inside one shared array; all barrier / buffer addresses are formed once as 32-bit
// shared-window addresses (taking the address of a __shared__ variable inside the loops costs an S2UR + shifts
// every time).
constexpr int kBarXFull = 0;
constexpr int kBarXEmpty = kBarXFull + kXStages;
constexpr int kBarWFull = kBarXEmpty + kXStages;
constexpr int kBarWEmpty = kBarWFull + kMaxWStages;
constexpr int kBarAcc1Full = kBarWEmpty + kMaxWStages;
constexpr int kBarVReady = kBarAcc1Full + 1;
constexpr int kBarAcc2Full = kBarVReady + kMaxN1 / 64;
constexpr int kBarAcc2Empty = kBarAcc2Full + 2;
constexpr int kNumBars = kBarAcc2Empty + 2;


template <bool PROF>
__global__ void __launch_bounds__(kThreads, 1)
    lowrank2_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                        const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_y,
                        const float* __restrict__ bias, const Params p,
                        unsigned long long* __restrict__ prof) {
  const long long t_kernel0 = PROF ? clock64() : 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[kNumBars];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t Vs = smem0;                                         // nkv x 16 KB
  const uint32_t Xs = Vs + (uint32_t)(p.ts ? 0 : p.nkv) * kStageBytes;   // kXStages x 16 KB (V lives in TMEM in TS mode)
  const uint32_t Ws = Xs + (uint32_t)kXStages * kStageBytes;         // wstages x 16 KB
  const uint32_t St = Ws + (uint32_t)p.wstages * kStageBytes;        // 4 warps x 2 x 4 KB output staging
  const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kXStages; ++s) {
      mbar_init(BAR(kBarXFull + s), 1);
      mbar_init(BAR(kBarXEmpty + s), 1);
    }
    for (int s = 0; s < kMaxWStages; ++s) {
      mbar_init(BAR(kBarWFull + s), 1);
      mbar_init(BAR(kBarWEmpty + s), 1);
    }
    mbar_init(BAR(kBarAcc1Full), 1);
    for (int b = 0; b < kMaxN1 / 64; ++b) mbar_init(BAR(kBarVReady + b), 128);
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR(kBarAcc2Full + b), 1);
      mbar_init(BAR(kBarAcc2Empty + b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  // TTA_LR2_PROF=1: cycles spent by CTA 0's roles in each wait / work phase (diagnostics)
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tp = 0;
  const bool profiling = PROF && prof != nullptr && blockIdx.x == 0;
#define LR2_T0() if constexpr (PROF) { if (profiling) tp = clock64(); }
#define LR2_ACC(i) if constexpr (PROF) { if (profiling) { const long long t1_ = clock64(); pc[i] += t1_ - tp; tp = t1_; } }

  // The producer and the MMA issuer run their loops with the whole warp (barrier waits are warp-wide polls) and
  // issue through ONE elected lane: code under `elect.sync` stays on the uniform datapath, whereas a role
  // wrapped in `if (lane == 0)` makes the compiler serialise every TMA / UMMA instruction in an election loop.
  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    const bool leader = elect_one();
    if (leader) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w2)) : "memory");
    }
    int xs = 0, ws = 0;
    uint32_t xph = 0, wph = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int m0 = tile * kBM;
      for (int kb = 0; kb < p.nk1; ++kb) {
        mbar_wait(BAR(kBarXEmpty + xs), xph ^ 1);
        if (leader) {
          mbar_expect_tx(BAR(kBarXFull + xs), kStageBytes);
          tma_load_2d(Xs + (uint32_t)xs * kStageBytes, &tm_x, BAR(kBarXFull + xs), kb * kBK, m0);
        }
        if (++xs == kXStages) { xs = 0; xph ^= 1; }
        for (int j = 0; j < p.nb1; ++j) {
          LR2_T0() mbar_wait(BAR(kBarWEmpty + ws), wph ^ 1); LR2_ACC(0)
          if (leader) {
            mbar_expect_tx(BAR(kBarWFull + ws), kStageBytes);
            tma_load_2d(Ws + (uint32_t)ws * kStageBytes, &tm_w1, BAR(kBarWFull + ws), kb * kBK, j * 128);
          }
          if (++ws == p.wstages) { ws = 0; wph ^= 1; }
        }
      }
      for (int c = 0; c < p.nchunks; ++c)
        for (int kb = 0; kb < p.nkv; ++kb) {
          LR2_T0() mbar_wait(BAR(kBarWEmpty + ws), wph ^ 1); LR2_ACC(1)
          if (leader) {
            mbar_expect_tx(BAR(kBarWFull + ws), (uint32_t)p.bn2 * 128u);
            tma_load_2d(Ws + (uint32_t)ws * kStageBytes, &tm_w2, BAR(kBarWFull + ws), kb * kBK, c * p.bn2);
          }
          if (++ws == p.wstages) { ws = 0; wph ^= 1; }
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const long long t_role0 = PROF ? clock64() : 0;
    const bool leader = elect_one();
    int xs = 0, ws = 0;
    uint32_t xph = 0, wph = 0;
    uint32_t it = 0, g = 0;   // row tiles / GEMM-2 chunks processed by this CTA
    const uint32_t idesc2 = umma_idesc_bf16(p.bn2);
    const uint64_t descX = umma_desc_sw128(Xs), descW = umma_desc_sw128(Ws), descV = umma_desc_sw128(Vs);
    constexpr uint64_t kStageDesc = kStageBytes >> 4;    // descriptor address units are 16 bytes
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      // ---- GEMM 1 ----
      for (int kb = 0; kb < p.nk1; ++kb) {
        LR2_T0() mbar_wait(BAR(kBarXFull + xs), xph); LR2_ACC(2)
        const uint64_t da = descX + (uint64_t)xs * kStageDesc;
        for (int j = 0; j < p.nb1; ++j) {
          LR2_T0() mbar_wait(BAR(kBarWFull + ws), wph); LR2_ACC(3)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (leader) {
            const uint64_t db = descW + (uint64_t)ws * kStageDesc;
            const int nj = (p.n1p16 - j * 128) < 128 ? (p.n1p16 - j * 128) : 128;
            const uint32_t idesc1 = umma_idesc_bf16(nj);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16(tmem_base + (uint32_t)(j * 128), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1,
                        (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(BAR(kBarWEmpty + ws));
          }
          __syncwarp();
          LR2_ACC(0)
          if (++ws == p.wstages) { ws = 0; wph ^= 1; }
        }
        if (leader) umma_commit(BAR(kBarXEmpty + xs));
        __syncwarp();
        if (++xs == kXStages) { xs = 0; xph ^= 1; }
      }
      if (leader) umma_commit(BAR(kBarAcc1Full));
      __syncwarp();
      // ---- GEMM 2 ----
      for (int c = 0; c < p.nchunks; ++c, ++g) {
        const uint32_t buf = g & 1;
        LR2_T0() mbar_wait(BAR(kBarAcc2Empty + buf), ((g >> 1) & 1) ^ 1); LR2_ACC(4)
        const uint32_t d_tmem = tmem_base + (uint32_t)(p.acc2_col + (int)buf * p.bn2);
        for (int kb = 0; kb < p.nkv; ++kb) {
          LR2_T0() if (c == 0) mbar_wait(BAR(kBarVReady + kb), it & 1); LR2_ACC(5)
          mbar_wait(BAR(kBarWFull + ws), wph); LR2_ACC(6)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (leader) {
            const uint64_t db = descW + (uint64_t)ws * kStageDesc;
            if (p.ts) {
              // A = V from TMEM: bf16 pairs packed in 32-bit cells, 16 k-elements = 8 columns per MMA
              const uint32_t a_tmem = tmem_base + (uint32_t)(kb * 32);
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                umma_bf16_ts(d_tmem, a_tmem + (uint32_t)(8 * k), db + (uint64_t)(2 * k), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
            } else {
              const uint64_t da = descV + (uint64_t)kb * kStageDesc;
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(BAR(kBarWEmpty + ws));
          }
          __syncwarp();
          LR2_ACC(7)
          if (++ws == p.wstages) { ws = 0; wph ^= 1; }
        }
        if (leader) umma_commit(BAR(kBarAcc2Full + buf));
        __syncwarp();
      }
    }
    if constexpr (PROF) {
      if (profiling && lane == 0) {
        prof[24] = (unsigned long long)(t_role0 - t_kernel0);
        prof[25] = (unsigned long long)(clock64() - t_role0);
      }
    }
  } else {
    // ------------------------------ epilogue warps ------------------------------
    const int q = warp & 3;                  // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;           // row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stage = St + (uint32_t)q * 8192;   // 2 x (32 rows x 128 B) per warp
    uint32_t it = 0, g = 0, nst = 0;
    // Bias of a block's 32 columns: eight 128-bit loads (every lane reads the same addresses: one broadcast
    // transaction each) on the aligned interior, predicated scalars on a ragged edge.  With 224 KB of the SM's
    // memory carved out as shared memory the loads are served by L2 (~600 cycles), so the bias of block i + 1
    // is requested while block i is processed.
    auto load_bias = [&](int gn0, float (&dst)[32]) {
      if (bias == nullptr || gn0 >= p.N2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = 0.f;
      } else if (gn0 + 32 <= p.N2 && p.bias_vec) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + gn0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(b4 + j);
          dst[4 * j] = t.x; dst[4 * j + 1] = t.y; dst[4 * j + 2] = t.z; dst[4 * j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = (gn0 + j < p.N2) ? __ldg(bias + gn0 + j) : 0.f;
      }
    };
    float bnext[32];
    load_bias(0, bnext);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int tile0 = tile * kBM;
      // ---- acc1 -> bf16 -> V ----
      LR2_T0() mbar_wait(BAR(kBarAcc1Full), it & 1); LR2_ACC(0)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c0 = 0; c0 < p.nkv * 64; c0 += 32) {
        uint32_t v[32];
        if (c0 < p.n1p16) {
          tmem_ld32(lane_addr + (uint32_t)c0, v);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = c0 + 2 * j;
          const float lo = col < p.N1 ? __uint_as_float(v[2 * j]) : 0.f;
          const float hi = col + 1 < p.N1 ? __uint_as_float(v[2 * j + 1]) : 0.f;
          __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (p.ts) {
          // V overwrites acc1 in place: the pairs of columns c0 .. c0+31 (already in registers) go to the 16
          // cells c0/2 .. c0/2+15 of this lane, all of which lie in the part of acc1 that has been read
          tmem_st16(lane_addr + (uint32_t)(c0 >> 1), pk);
          if ((c0 & 63) == 32) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(BAR(kBarVReady + (c0 >> 6)));
          }
        } else {
          const uint32_t vrow = Vs + (uint32_t)(c0 >> 6) * kStageBytes + row * 128;
          const int ch0 = (c0 & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(vrow + (((ch0 + i) ^ (row & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          if ((c0 & 63) == 32) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes of V -> async proxy (UMMA)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(BAR(kBarVReady + (c0 >> 6)));
          }
        }
      }
      LR2_ACC(1)
      // ---- acc2 chunks -> y ----
      for (int c = 0; c < p.nchunks; ++c, ++g) {
        const uint32_t buf = g & 1;
        LR2_T0() mbar_wait(BAR(kBarAcc2Full + buf), (g >> 1) & 1); LR2_ACC(2)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < p.bn2; c0 += 32) {
          uint32_t v[32];
          LR2_T0()
          tmem_ld32(lane_addr + (uint32_t)(p.acc2_col + (int)buf * p.bn2 + c0), v);
          LR2_ACC(4)
          if (c0 + 32 >= p.bn2) {
            // the last block of this accumulator is in registers: hand the buffer back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(BAR(kBarAcc2Empty + buf));
          }
          const int gn0 = c * p.bn2 + c0;
          float bv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) bv[j] = bnext[j];
          load_bias(c0 + 32 < p.bn2 ? gn0 + 32 : (c + 1 < p.nchunks ? (c + 1) * p.bn2 : 0), bnext);
          if (gn0 >= p.N2) continue;                                   // warp-uniform
          // The accumulator arrives one row per lane; a row-per-lane store would touch 32 lines per
          // instruction.  The 32 x 32 block goes (bias added, converted) into this warp's staging buffer in
          // the SWIZZLE_128B (fp32) / SWIZZLE_64B (bf16) pattern of the output tensor map, and one lane hands
          // it to the TMA store engine, which writes whole lines and clips rows >= M / columns >= N2 (16-byte
          // granules).  Two staging buffers per warp: the store of block i drains while block i + 1 is staged.
          const uint32_t sbuf = stage + (nst & 1) * 4096;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffer of block i - 2 is free
          __syncwarp();
          LR2_ACC(5)
          if (p.out_f32) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              sts128(sbuf + lane * 128 + ((ch ^ (lane & 7)) << 4), __float_as_uint(__uint_as_float(v[4 * ch]) + bv[4 * ch]),
                     __float_as_uint(__uint_as_float(v[4 * ch + 1]) + bv[4 * ch + 1]),
                     __float_as_uint(__uint_as_float(v[4 * ch + 2]) + bv[4 * ch + 2]),
                     __float_as_uint(__uint_as_float(v[4 * ch + 3]) + bv[4 * ch + 3]));
          } else {
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[8 * ch + 2 * e]) + bv[8 * ch + 2 * e],
                                                         __uint_as_float(v[8 * ch + 2 * e + 1]) + bv[8 * ch + 2 * e + 1]);
                pk[e] = *reinterpret_cast<uint32_t*>(&h);
              }
              sts128(sbuf + lane * 64 + ((ch ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
          }
          LR2_ACC(6)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged block -> async proxy (TMA)
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tm_y)),
                         "r"(sbuf), "r"(gn0), "r"(tile0 + q * 32)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++nst;
          LR2_ACC(7)
        }
        LR2_ACC(3)
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores done before the CTA retires
    __syncwarp();
  }

  if constexpr (PROF) {
    if (profiling && lane == 0 && warp <= 2)
      for (int i = 0; i < 8; ++i) prof[warp * 8 + i] = (unsigned long long)pc[i];
  }
#undef LR2_T0
#undef LR2_ACC
#undef BAR
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (PROF) {
    if (profiling && tid == 0) prof[26] = (unsigned long long)(clock64() - t_kernel0);
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


}  // namespace lr2
}  // namespace tta

extern "C" int tta_lowrank2_fwd(const void* x, int64_t ldx, const void* w1, int64_t ld1, const void* w2, int64_t ld2,
                                const float* bias, void* y, int64_t ldy, int out_fp32, int64_t M, int K1, int N1,
                                int N2, void* stream) {
  using namespace tta;
  using namespace tta::lr2;
  if (M <= 0 || N2 <= 0) return TTA_OK;
  if (!x || !w1 || !w2 || !y || K1 <= 0 || N1 <= 0) {
    set_error("lowrank2_fwd: bad argument");
    return TTA_E_INVALID;
  }
  if (N1 > kMaxN1) {
    set_error("lowrank2_fwd: inner width %d exceeds %d (TMEM budget)", N1, kMaxN1);
    return TTA_E_INVALID;
  }
  if ((ldx & 7) || (ld1 & 7) || (ld2 & 7) || ((uintptr_t)x & 15) || ((uintptr_t)w1 & 15) || ((uintptr_t)w2 & 15) ||
      ((uintptr_t)y & 15) || (ldy & (out_fp32 ? 3 : 7)) || M > (int64_t)kBM * 0x7fffff) {
    set_error("lowrank2_fwd: operands must be 16-byte aligned with leading dimensions multiples of 8 (ldx %lld ld1 %lld ld2 %lld ldy %lld)",
              (long long)ldx, (long long)ld1, (long long)ld2, (long long)ldy);
    return TTA_E_INVALID;
  }
  Params p;
  p.M = (int)M; p.K1 = K1; p.N1 = N1; p.N2 = N2;
  p.nk1 = (K1 + kBK - 1) / kBK;
  p.nb1 = (N1 + 127) / 128;
  p.n1p16 = (N1 + 15) & ~15;
  p.nkv = (N1 + 63) / 64;
  p.acc2_col = (p.n1p16 + 31) & ~31;
  int bn2 = ((512 - p.acc2_col) / 2) & ~31;
  if (bn2 > 128) bn2 = 128;
  // no wider than the output needs
  while (bn2 > 32 && bn2 - 32 >= N2) bn2 -= 32;
  p.bn2 = bn2;
  p.nchunks = (N2 + bn2 - 1) / bn2;
  p.ntiles = (int)((M + kBM - 1) / kBM);
  p.out_f32 = out_fp32 ? 1 : 0;
  p.bias_vec = (bias && ((uintptr_t)bias & 15) == 0) ? 1 : 0;
  p.ldy = ldy;
  const size_t budget = 227 * 1024 - 1024 - 512;   // dynamic shared memory minus alignment slack and the barriers
  static int ts_mode = -1;
  if (ts_mode < 0) {
    const char* e = getenv("TTA_LR2_TS");
    ts_mode = (e && e[0] == '0') ? 0 : 1;
  }
  p.ts = ts_mode;
  const size_t fixed = (size_t)((p.ts ? 0 : p.nkv) + kXStages + 2) * kStageBytes;   // (V,) x ring, output staging (8 x 4 KB)
  int wst = (int)((budget - fixed) / kStageBytes);
  if (wst > kMaxWStages) wst = kMaxWStages;
  if (wst < 2) {
    set_error("lowrank2_fwd: shared memory exhausted (N1 = %d)", N1);
    return TTA_E_INVALID;
  }
  p.wstages = wst;
  const size_t smem = fixed + (size_t)wst * kStageBytes + 1024;

  CUtensorMap tm_x, tm_w1, tm_w2, tm_y;
  int rc = make_map(&tm_x, x, M, K1, ldx, kBM);
  if (rc) return rc;
  rc = make_map(&tm_w1, w1, N1, K1, ld1, 128);
  if (rc) return rc;
  rc = make_map(&tm_w2, w2, N2, N1, ld2, bn2);
  if (rc) return rc;
  // output: 32 x 32 blocks, 128-byte (fp32) or 64-byte (bf16) rows in the staging buffers
  // The TMA store works in 16-byte granules: the map covers N2 rounded up to 4 fp32 / 8 bf16 columns (<= ldy by the
  // alignment rule above); the pad columns of a ragged N2 receive zeros.
  const int n2m = out_fp32 ? ((N2 + 3) & ~3) : ((N2 + 7) & ~7);
  rc = out_fp32 ? make_map(&tm_y, y, M, n2m, ldy, 32, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map(&tm_y, y, M, n2m, ldy, 32, 32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;

  static size_t smem_set = 0;
  if (smem > smem_set) {
    rc = check_cuda(cudaFuncSetAttribute(lowrank2_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "lowrank2_fwd smem attribute");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(lowrank2_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "lowrank2_fwd smem attribute");
    if (rc) return rc;
    smem_set = smem;
  }
  const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
  static unsigned long long* prof = nullptr;
  static bool prof_init = false;
  if (!prof_init) {
    prof_init = true;
    const char* e = getenv("TTA_LR2_PROF");
    if (e && e[0] == '1' && cudaMalloc(&prof, 32 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(prof, 0, 32 * sizeof(unsigned long long));
  }
  if (prof) {
    unsigned long long h[32];
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    if (h[8] | h[16])
      fprintf(stderr,
              "[lr2 prof, CTA 0, cycles] producer: wait w_empty g1 %llu g2 %llu | mma: wait x_full %llu w_full(g1) %llu acc2_empty %llu "
              "v_ready %llu w_full(g2) %llu issue(g2) %llu issue(g1) %llu | epilogue warp: wait acc1 %llu convert V %llu wait acc2 %llu store y (rest) %llu "
              "[tmem ld %llu, wait staging %llu, bias+convert+sts %llu, fence+tma %llu] | setup %llu mma role %llu whole CTA %llu\n",
              h[0], h[1], h[10], h[11], h[12], h[13], h[14], h[15], h[8], h[16], h[17], h[18], h[19], h[20], h[21], h[22], h[23], h[24], h[25], h[26]);
    cudaMemset(prof, 0, sizeof(h));
  }
  if (prof)
    lowrank2_fwd_kernel<true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tm_x, tm_w1, tm_w2, tm_y, bias, p, prof);
  else
    lowrank2_fwd_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tm_x, tm_w1, tm_w2, tm_y, bias, p, nullptr);
  TTA_CHECK_LAUNCH("lowrank2_fwd launch");
  return TTA_OK;
}
