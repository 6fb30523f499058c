// Weight-gradient GEMM on the 5th-generation tensor cores: C[M,N] (fp32) = A^T B with A (K x M) and B (K x N) bf16
// row-major -- the reduction index (the tokens of a batch, K ~ 50 000) is the SLOW index of both operands.
// This is dW2 = dY^T V and dW1 = dV^T X of the fused two-factor linear layer (fwd_common.LowRank2Fn.backward: the
// training path of TTLinearM / TKLinearM, engines.py:293-320 drives it), which used torch.mm (cuBLAS) in round 1.
//
// Both operands are MN-major for the MMA: a TMA box of 64 reduction rows x 64 contiguous elements (128 bytes,
// SWIZZLE_128B) is exactly the canonical MN-major SW128 atom sequence of the UMMA descriptors (8 reduction rows x 128
// bytes per atom, SBO = 1024 bytes between atoms along K, LBO = one box = 8 KB between 64-element blocks along M / N),
// so no transposition happens anywhere: TMA -> shared memory -> tcgen05.mma with the a_major / b_major bits set.
// The output is small (M, N <= 1536) and the reduction long: split-K over CTAs, fp32 partial tiles in a workspace,
// summed in a fixed order by a second kernel (deterministic).
//   warp 0     TMA producer (one elected lane): four boxes per 64-row k-block, 4-stage ring, mbarrier expect_tx.
//   warp 1     TMEM allocator (256 columns) + tcgen05.mma issuer (kind::f16, bf16 operands, fp32 accumulation).
//   warps 2-5  epilogue: tcgen05.ld -> partial tile.
#include <cstdlib>

#include "tc_common.cuh"

namespace tta {
namespace gtn {

using namespace tta::tc;

constexpr int kThreads = 192;
constexpr int kStages = 4;
constexpr int kBoxBytes = 64 * 128;            // 64 reduction rows x 128 bytes
constexpr int kMaxStageBytes = 6 * kBoxBytes;  // A: two 64-column boxes, B: two (BN = 128) or four (BN = 256)
constexpr int kSmem = kStages * kMaxStageBytes + 1024;

struct Params {
  int M, N, K;
  int nsplit, kb_per_split, nkb;
  int tiles_m, tiles_n;
  int bn;         // tile width: 128 or 256 (each CTA re-reads its A and B rows from L2: wider tiles, less traffic)
  float* ws;      // partials [nsplit][M][N]
};

// MN-major operand, SWIZZLE_128B: LBO = bytes between 64-element blocks along M / N, SBO = bytes between 8-row groups
// along K
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(kBoxBytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
    gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Params p) {
  extern __shared__ __align__(1024) uint8_t gtn_smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kStages + 1];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t smem0 = (smem_u32(gtn_smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * kStages, accbar = bar0 + 16 * kStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int item = blockIdx.x;
  const int tn = item % p.tiles_n;
  item /= p.tiles_n;
  const int tm = item % p.tiles_m;
  const int split = item / p.tiles_m;
  const int m0 = tm * 128, n0 = tn * p.bn;
  const int nbb = p.bn >> 6;                                  // B boxes per stage
  const uint32_t stage_bytes = (uint32_t)(2 + nbb) * kBoxBytes;
  const int kb0 = split * p.kb_per_split;
  int kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.nkb) kb1 = p.nkb;
  const int nk = kb1 > kb0 ? kb1 - kb0 : 0;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
      for (int i = 0; i < nk; ++i) {
        const int s = i % kStages;
        if (i >= kStages) mbar_wait(empty0 + 8 * s, (uint32_t)(((i / kStages) - 1) & 1));
        const uint32_t bar = full0 + 8 * s;
        mbar_expect_tx(bar, stage_bytes);
        const uint32_t base = smem0 + (uint32_t)s * stage_bytes;
        const int kr = (kb0 + i) * 64;
        tma_load_2d(base, &tm_a, bar, m0, kr);
        tma_load_2d(base + kBoxBytes, &tm_a, bar, m0 + 64, kr);
        for (int bb = 0; bb < nbb; ++bb) tma_load_2d(base + (uint32_t)(2 + bb) * kBoxBytes, &tm_b, bar, n0 + 64 * bb, kr);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // D = fp32, A = B = bf16, both MN-major (bits 15 / 16), M = 128, N = bn
    const uint32_t idesc = umma_idesc_bf16(p.bn) | (1u << 15) | (1u << 16);
    for (int i = 0; i < nk; ++i) {
      const int s = i % kStages;
      mbar_wait(full0 + 8 * s, (uint32_t)((i / kStages) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t base = smem0 + (uint32_t)s * stage_bytes;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)       // 16 reduction rows = two 8-row atoms = 2048 bytes
          umma_bf16(tmem_base, desc_mn_sw128(base + (uint32_t)ks * 2048u), desc_mn_sw128(base + 2 * kBoxBytes + (uint32_t)ks * 2048u),
                    idesc, (i | ks) ? 1u : 0u);
        umma_commit(empty0 + 8 * s);
        if (i == nk - 1) umma_commit(accbar);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int gm = m0 + quad * 32 + lane;
    if (nk > 0) {
      mbar_wait(accbar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    float* prow = p.ws + ((int64_t)split * p.M + gm) * p.N + n0;
    const bool vec = (p.N & 3) == 0;
    for (int c0 = 0; c0 < p.bn; c0 += 32) {
      uint32_t v[32];
      if (nk > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (gm < p.M) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int gn = n0 + c0 + j;
          if (vec && gn + 4 <= p.N) {
            *reinterpret_cast<float4*>(prow + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (gn + e < p.N) prow[c0 + j + e] = __uint_as_float(v[j + e]);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// C[m][n] = sum over splits of the partial tiles (fixed order)
__global__ void __launch_bounds__(256) gemm_tn_reduce(const float* __restrict__ ws, float* __restrict__ c, int M, int N, int64_t ldc,
                                                     int nsplit) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += ws[(int64_t)sp * total + e];
    const int64_t m = e / N;
    c[m * ldc + (e - m * N)] = s;
  }
}

static void plan(int M, int N, int K, int& tiles_m, int& tiles_n, int& nkb, int& nsplit, int& kb_per_split, int& bn) {
  static const int env_bn = [] { const char* e = getenv("TTA_TN_BN"); return e ? atoi(e) : 0; }();
  static const int env_ctas = [] { const char* e = getenv("TTA_TN_CTAS"); return e ? atoi(e) : 0; }();
  bn = env_bn ? env_bn : (N > 128 ? 256 : 128);
  tiles_m = (M + 127) / 128;
  tiles_n = (N + bn - 1) / bn;
  nkb = (K + 63) / 64;
  const int tiles = tiles_m * tiles_n;
  // about 1.5 CTAs per SM in total (measured best of 1 / 1.5 / 2 / 3 on the DeiT-small shapes: the kernel is bound by the
  // L2 -> shared-memory traffic of the re-read operand rows, 6-7 TB/s, not by the split)
  nsplit = ((env_ctas ? env_ctas : 3 * kNumSMs / 2) + tiles - 1) / tiles;
  if (nsplit > nkb) nsplit = nkb;
  if (nsplit > 64) nsplit = 64;
  if (nsplit < 1) nsplit = 1;
  kb_per_split = (nkb + nsplit - 1) / nsplit;
  nsplit = (nkb + kb_per_split - 1) / kb_per_split;
  if (nsplit < 1) nsplit = 1;
}

}  // namespace gtn
}  // namespace tta

extern "C" int64_t tta_gemm_bf16_tn_workspace_bytes(int M, int N, int K) {
  int tm, tn, nkb, ns, per, bn;
  tta::gtn::plan(M, N, K, tm, tn, nkb, ns, per, bn);
  return (int64_t)ns * M * N * 4;
}

extern "C" int tta_gemm_bf16_tn(const void* a, int64_t lda, const void* b, int64_t ldb, float* c, int64_t ldc, int M, int N, int K,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace tta;
  using namespace tta::gtn;
  if (!a || !b || !c || !workspace || M <= 0 || N <= 0 || K <= 0) {
    set_error("gemm_bf16_tn: bad argument");
    return TTA_E_INVALID;
  }
  if ((lda & 7) || (ldb & 7) || lda < M || ldb < N || (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) {
    set_error("gemm_bf16_tn: operands need 16-byte aligned bases and row pitches that are multiples of 8 elements");
    return TTA_E_INVALID;
  }
  Params p;
  p.M = M; p.N = N; p.K = K;
  plan(M, N, K, p.tiles_m, p.tiles_n, p.nkb, p.nsplit, p.kb_per_split, p.bn);
  if (workspace_bytes < (int64_t)p.nsplit * M * N * 4) {
    set_error("gemm_bf16_tn: workspace of %lld bytes is too small", (long long)workspace_bytes);
    return TTA_E_INVALID;
  }
  p.ws = reinterpret_cast<float*>(workspace);
  CUtensorMap tm_a, tm_b;
  int rc = tc::make_map(&tm_a, a, K, M, lda, 64, 64);        // rows = reduction index, box 64 rows x 64 elements (128 B)
  if (rc) return rc;
  rc = tc::make_map(&tm_b, b, K, N, ldb, 64, 64);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem), "gemm_tn smem attribute");
    if (rc) return rc;
    attr_set = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  gemm_tn_kernel<<<p.tiles_m * p.tiles_n * p.nsplit, kThreads, kSmem, st>>>(tm_a, tm_b, p);
  TTA_CHECK_LAUNCH("gemm_tn launch");
  int gx = (int)(((int64_t)M * N + 255) / 256);
  if (gx > kNumSMs * 8) gx = kNumSMs * 8;
  gemm_tn_reduce<<<gx, 256, 0, st>>>(p.ws, c, M, N, ldc, p.nsplit);
  TTA_CHECK_LAUNCH("gemm_tn reduce launch");
  return TTA_OK;
}
