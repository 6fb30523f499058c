// Persistent TMA-fed bf16 GEMM on the 5th-generation tensor cores:
//   C[M,N] (bf16 or fp32) = A[M,K] (bf16, K-major) * B[N,K]^T (bf16, K-major) (+ bias[N])
// Serves `tta_gemm_bf16_tc` (the single contractions of the decomposed-layer forwards: TTLinearR / TKLinearR
// with the rebuilt weight, TTLinear.py:96-160, TKLinear.py:98-122; the im2col products of TTConv2dR / TKConv2dR,
// TTConv.py:313-333) whenever the operands meet TMA's alignment rules; gemm_tc.cu (cp.async producers, one tile
// per CTA) remains for the others.
//
// Grid = min(tiles, 148) persistent CTAs; a tile is 128 rows x BN columns, BN = the N chunk (multiple of 32,
// <= 256, chosen so that the chunks cover N with the least padding), tiles ordered with the chunk index
// fastest so that the CTAs working at the same time share the A rows in L2.  Warp roles (192 threads):
//   warp 0     TMA producer (one elected lane): per k-block one A box (128 x 64) and one B box (BN x 64),
//              SWIZZLE_128B, into a ring of stages; mbarrier expect_tx / complete_tx.
//   warp 1     TMEM allocator (512 columns = two accumulators) + tcgen05.mma issuer (one elected lane).
//   warps 2-5  epilogue (one TMEM lane quadrant each): accumulator -> +bias -> swizzled staging -> TMA store.
// The accumulator is double buffered: the epilogue of tile i overlaps the MMAs of tile i + 1.
#include <stdlib.h>

#include "tc_common.cuh"

namespace tta {
namespace gt {

using namespace tta::tc;

constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct Params {
  int M, N, K;
  int nkb;          // k-blocks
  int bn;           // chunk width (multiple of 32, <= 256)
  int nchunks;
  int ntiles;       // m tiles * nchunks
  int stages;
  int stage_bytes;  // 16 KB (A) + BN * 128 rounded up to 1 KB (B)
  int out_f32;
  int bias_vec;
};

constexpr int kBarFull = 0;
constexpr int kBarEmpty = kBarFull + kMaxStages;
constexpr int kBarAccFull = kBarEmpty + kMaxStages;
constexpr int kBarAccEmpty = kBarAccFull + 2;
constexpr int kNumBars = kBarAccEmpty + 2;

__global__ void __launch_bounds__(kThreads, 1)
    gemm_tma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ CUtensorMap tm_c, const float* __restrict__ bias, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[kNumBars];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t Rs = smem0;                                            // stages x (A 16 KB | B)
  const uint32_t St = Rs + (uint32_t)p.stages * (uint32_t)p.stage_bytes;   // 4 warps x 2 x 4 KB output staging
  const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(BAR(kBarFull + s), 1);
      mbar_init(BAR(kBarEmpty + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR(kBarAccFull + b), 1);
      mbar_init(BAR(kBarAccEmpty + b), 128);
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

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    const bool leader = elect_one();
    if (leader) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
    }
    int st = 0;
    uint32_t ph = 0;
    const uint32_t tx_bytes = (uint32_t)kStageBytes + (uint32_t)p.bn * 128u;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int m0 = (tile / p.nchunks) * kBM, n0 = (tile % p.nchunks) * p.bn;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(BAR(kBarEmpty + st), ph ^ 1);
        if (leader) {
          const uint32_t sa = Rs + (uint32_t)st * (uint32_t)p.stage_bytes;
          mbar_expect_tx(BAR(kBarFull + st), tx_bytes);
          tma_load_2d(sa, &tm_a, BAR(kBarFull + st), kb * kBK, m0);
          tma_load_2d(sa + kStageBytes, &tm_b, BAR(kBarFull + st), kb * kBK, n0);
        }
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const bool leader = elect_one();
    int st = 0;
    uint32_t ph = 0, g = 0;
    const uint32_t idesc = umma_idesc_bf16(p.bn);
    const uint64_t desc0 = umma_desc_sw128(Rs);
    const uint64_t stage_desc = (uint64_t)(p.stage_bytes >> 4);   // descriptor address units are 16 bytes
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++g) {
      const uint32_t buf = g & 1;
      mbar_wait(BAR(kBarAccEmpty + buf), ((g >> 1) & 1) ^ 1);
      const uint32_t d_tmem = tmem_base + buf * 256u;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(BAR(kBarFull + st), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) {
          const uint64_t da = desc0 + (uint64_t)st * stage_desc;
          const uint64_t db = da + (uint64_t)(kStageBytes >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(BAR(kBarEmpty + st));
          if (kb == p.nkb - 1) umma_commit(BAR(kBarAccFull + buf));
        }
        __syncwarp();
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue warps ------------------------------
    const int q = warp & 3;                  // TMEM lane quadrant this warp may read
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stage = St + (uint32_t)q * 8192;   // 2 x (32 rows x 128 B) per warp
    uint32_t g = 0, nst = 0;
    // bias of a block's 32 columns, requested one block ahead (served by L2 with this shared-memory carve-out)
    auto load_bias = [&](int gn0, float (&dst)[32]) {
      if (bias == nullptr || gn0 >= p.N) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = 0.f;
      } else if (gn0 + 32 <= p.N && p.bias_vec) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + gn0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(b4 + j);
          dst[4 * j] = t.x; dst[4 * j + 1] = t.y; dst[4 * j + 2] = t.z; dst[4 * j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = (gn0 + j < p.N) ? __ldg(bias + gn0 + j) : 0.f;
      }
    };
    float bnext[32];
    load_bias(((int)blockIdx.x % p.nchunks) * p.bn, bnext);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++g) {
      const int m0 = (tile / p.nchunks) * kBM, n0 = (tile % p.nchunks) * p.bn;
      const int next_tile = tile + (int)gridDim.x;
      const int next_n0 = next_tile < p.ntiles ? (next_tile % p.nchunks) * p.bn : 0;
      const uint32_t buf = g & 1;
      mbar_wait(BAR(kBarAccFull + buf), (g >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * 256u + (uint32_t)c0, v);
        if (c0 + 32 >= p.bn) {
          // the last block of this accumulator is in registers: hand the buffer back to the MMA issuer
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(BAR(kBarAccEmpty + buf));
        }
        const int gn0 = n0 + c0;
        float bv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) bv[j] = bnext[j];
        load_bias(c0 + 32 < p.bn ? gn0 + 32 : next_n0, bnext);
        if (gn0 >= p.N) continue;                                   // warp-uniform
        const uint32_t sbuf = stage + (nst & 1) * 4096;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffer of block i - 2 is free
        __syncwarp();
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
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged block -> async proxy (TMA)
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tm_c)),
                       "r"(sbuf), "r"(gn0), "r"(m0 + q * 32)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++nst;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores done before the CTA retires
    __syncwarp();
  }
#undef BAR
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace gt

// true when the operands meet the TMA rules of gemm_tma_kernel (16-byte aligned bases and row pitches)
bool gemm_tma_eligible(const void* a, int64_t lda, const void* b, int64_t ldb, const void* c, int64_t ldc, int M, int N,
                       int K, int out_fp32) {
  if (M < 1 || N < 1 || K < 8 || (K & 7)) return false;
  if ((lda & 7) || (ldb & 7) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15) || ((uintptr_t)c & 15)) return false;
  if (ldc & (out_fp32 ? 3 : 7)) return false;
  if ((int64_t)M > (int64_t)gt::kBM * 0x3fffff) return false;
  return true;
}

// Returns TTA_OK after enqueueing; the pad columns of C up to a 16-byte granule (inside ldc) receive zeros.
int gemm_tma_launch(const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int M, int N, int K,
                    const float* bias, int out_fp32, cudaStream_t st) {
  using namespace gt;
  Params p;
  p.M = M; p.N = N; p.K = K;
  p.nkb = (K + kBK - 1) / kBK;
  const int nch = (N + 255) / 256;
  int bn = ((N + nch - 1) / nch + 31) & ~31;
  if (bn > 256) bn = 256;
  p.bn = bn;
  p.nchunks = (N + bn - 1) / bn;
  const int64_t mt = ((int64_t)M + kBM - 1) / kBM;
  p.ntiles = (int)(mt * p.nchunks);
  p.stage_bytes = kStageBytes + ((bn * 128 + 1023) & ~1023);
  const size_t budget = 227 * 1024 - 1024 - 512 - 32 * 1024;   // minus alignment slack, barriers, output staging
  int stages = (int)(budget / (size_t)p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > p.nkb * 2 && stages > 2) stages = p.nkb * 2 > 2 ? p.nkb * 2 : 2;
  p.stages = stages;
  p.out_f32 = out_fp32 ? 1 : 0;
  p.bias_vec = (bias && ((uintptr_t)bias & 15) == 0) ? 1 : 0;
  const size_t smem = (size_t)stages * p.stage_bytes + 32 * 1024 + 1024;

  CUtensorMap tm_a, tm_b, tm_c;
  int rc = make_map(&tm_a, a, M, K, lda, kBM);
  if (rc) return rc;
  rc = make_map(&tm_b, b, N, K, ldb, bn);
  if (rc) return rc;
  const int ncols = out_fp32 ? ((N + 3) & ~3) : ((N + 7) & ~7);
  rc = out_fp32 ? make_map(&tm_c, c, M, ncols, ldc, 32, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map(&tm_c, c, M, ncols, ldc, 32, 32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;

  static size_t smem_set = 0;
  if (smem > smem_set) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "gemm_tma smem attribute");
    if (rc) return rc;
    smem_set = smem;
  }
  const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
  gemm_tma_kernel<<<grid, kThreads, smem, st>>>(tm_a, tm_b, tm_c, bias, p);
  TTA_CHECK_LAUNCH("gemm_tma launch");
  return TTA_OK;
}

}  // namespace tta
