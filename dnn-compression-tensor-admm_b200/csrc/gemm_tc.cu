// bf16 tensor-core GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), hand-written PTX.
//   C[M,N] (bf16 or fp32) = A[M,K] (bf16, K-major) * B[N,K]^T (bf16, K-major) (+ bias[N])
// Used by the decomposed-layer forwards (TTLinear / TTConv core contractions, TTLinear.py:79-88):
// the big middle contractions of the TT chain are token-major GEMMs with K, N in the hundreds.
//
// CTA = one 128 x BN output tile.  Warp roles (160 threads):
//   warps 0-3  producers: global -> shared with cp.async (16 B, zero-fill at the M/N/K edges) into a
//              4-stage ring, operands laid out as 128-byte rows with the SWIZZLE_128B pattern the UMMA
//              shared-memory descriptor expects; afterwards the same warps are the epilogue (each
//              owns one TMEM lane quadrant): tcgen05.ld -> bias -> convert -> global.
//   warp 4     TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases ring
//              slots / publishes the accumulator through mbarriers.
// Accumulator: 128 lanes x BN fp32 columns of TMEM.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "tta_common.cuh"

namespace tta {

// gemm_tma.cu
bool gemm_tma_eligible(const void* a, int64_t lda, const void* b, int64_t ldb, const void* c, int64_t ldc, int M, int N,
                       int K, int out_fp32);
int gemm_tma_launch(const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int M, int N, int K,
                    const float* bias, int out_fp32, cudaStream_t st);

constexpr int kTcBM = 128;
constexpr int kTcBK = 64;        // 64 bf16 = one 128-byte swizzle row
constexpr int kTcStages = 4;
constexpr int kTcLag = 2;        // cp.async groups in flight before a stage is published
constexpr int kTcThreads = 160;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, 16-byte units
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}

// instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128, N = bn
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}

template <int BN, bool OUT_F32>
__global__ void __launch_bounds__(kTcThreads, 1)
    gemm_bf16_tc_kernel(const __nv_bfloat16* __restrict__ A, int64_t lda, const __nv_bfloat16* __restrict__ B,
                        int64_t ldb, void* __restrict__ Cout, int64_t ldc, int M, int N, int K,
                        const float* __restrict__ bias) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stage][A 128x128B][B BNx128B], 1024-byte aligned
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kABytes = kTcBM * 128;
  constexpr int kBBytes = BN * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  __shared__ uint64_t full_bar[kTcStages];
  __shared__ uint64_t empty_bar[kTcStages];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTcBM;
  const int n0 = blockIdx.y * BN;
  const int nkb = (K + kTcBK - 1) / kTcBK;

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < 4) {
    // ------------------------------ producers ------------------------------
    for (int kb = 0; kb < nkb + kTcLag; ++kb) {
      if (kb < nkb) {
        const int s = kb % kTcStages;
        if (kb >= kTcStages) mbar_wait(&empty_bar[s], ((kb / kTcStages) - 1) & 1);
        uint8_t* sa = smem + (size_t)s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        const int k0 = kb * kTcBK;
#pragma unroll
        for (int i = 0; i < (kTcBM * 8) / 128; ++i) {
          const int c = i * 128 + tid;
          const int row = c >> 3, ch = c & 7;
          const int gk = k0 + ch * 8;
          const int gm = m0 + row;
          const bool ok = (gm < M) && (gk < K);
          const __nv_bfloat16* src = ok ? (A + (int64_t)gm * lda + gk) : A;
          cp_async16(smem_u32(sa + row * 128 + ((ch ^ (row & 7)) << 4)), src, ok ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < (BN * 8) / 128; ++i) {
          const int c = i * 128 + tid;
          const int row = c >> 3, ch = c & 7;
          const int gk = k0 + ch * 8;
          const int gn = n0 + row;
          const bool ok = (gn < N) && (gk < K);
          const __nv_bfloat16* src = ok ? (B + (int64_t)gn * ldb + gk) : B;
          cp_async16(smem_u32(sb + row * 128 + ((ch ^ (row & 7)) << 4)), src, ok ? 16 : 0);
        }
      }
      cp_async_commit();
      if (kb >= kTcLag) {
        cp_async_wait<kTcLag>();                               // k-block kb - kTcLag has landed
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
        mbar_arrive(&full_bar[(kb - kTcLag) % kTcStages]);
      }
    }
    // ------------------------------ epilogue ------------------------------
    mbar_wait(&accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int gm = m0 + warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(lane_addr + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (gm < M) {
        const int gn0 = n0 + c0;
        if (OUT_F32) {
          float* crow = reinterpret_cast<float*>(Cout) + (int64_t)gm * ldc;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int gn = gn0 + j;
            if (gn < N) crow[gn] = __uint_as_float(v[j]) + (bias ? __ldg(bias + gn) : 0.f);
          }
        } else {
          __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(Cout) + (int64_t)gm * ldc;
          if (gn0 + 16 <= N && ((ldc & 7) == 0) && ((gn0 & 7) == 0)) {
            uint32_t packed[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float lo = __uint_as_float(v[2 * j]) + (bias ? __ldg(bias + gn0 + 2 * j) : 0.f);
              const float hi = __uint_as_float(v[2 * j + 1]) + (bias ? __ldg(bias + gn0 + 2 * j + 1) : 0.f);
              __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
              packed[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            uint4* dst = reinterpret_cast<uint4*>(crow + gn0);
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int gn = gn0 + j;
              if (gn < N) crow[gn] = __float2bfloat16(__uint_as_float(v[j]) + (bias ? __ldg(bias + gn) : 0.f));
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ------------------------------ MMA issuer ------------------------------
    // The whole warp runs the loop (barrier waits are warp-wide polls) and ONE elected lane issues: code under
    // `elect.sync` stays on the uniform datapath, whereas `if (lane == 0)` makes the compiler wrap every
    // UMMA / commit in an election loop.
    uint32_t leader;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(leader));
    constexpr uint32_t idesc = umma_idesc_bf16(BN);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]), accum_addr = smem_u32(&accum_bar);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kTcStages;
      mbar_wait(&full_bar[s], (kb / kTcStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        const uint32_t sa = smem_base + (uint32_t)s * kStageBytes;
        const uint32_t sb = sa + kABytes;
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < kTcBK / 16; ++k) {
          const uint32_t accumulate = (kb > 0 || k > 0) ? 1u : 0u;
          // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
              "}\n" ::"r"(tmem_base),
              "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(accumulate)
              : "memory");
        }
        // frees the ring slot once the MMAs above have consumed it (implies fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8u * s)
                     : "memory");
        if (kb == nkb - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(accum_addr)
                       : "memory");
      }
      __syncwarp();
    }
    (void)full0;
  }

  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

template <int BN, bool OUT_F32>
static int launch_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M, int N, int K,
                     const float* bias, cudaStream_t st) {
  const size_t smem = (size_t)kTcStages * (kTcBM * 128 + BN * 128) + 1024;
  auto kern = gemm_bf16_tc_kernel<BN, OUT_F32>;
  int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "gemm_tc smem attribute");
  if (rc) return rc;
  dim3 grid((M + kTcBM - 1) / kTcBM, (N + BN - 1) / BN);
  kern<<<grid, kTcThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(A), lda,
                                        reinterpret_cast<const __nv_bfloat16*>(B), ldb, C, ldc, M, N, K, bias);
  TTA_CHECK_LAUNCH("gemm_tc launch");
  return TTA_OK;
}

}  // namespace tta

extern "C" int tta_gemm_bf16_tc(const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int M,
                                int N, int K, const float* bias, int out_fp32, void* stream) {
  using namespace tta;
  if (M <= 0 || N <= 0) return TTA_OK;
  if (!a || !b || !c || K <= 0 || (K & 7) || (lda & 7) || (ldb & 7) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15)) {
    set_error("gemm_bf16_tc: needs K, lda, ldb multiples of 8 and 16-byte aligned operands (M=%d N=%d K=%d)", M, N, K);
    return TTA_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // TMA-fed persistent kernel (gemm_tma.cu) whenever the operands meet TMA's alignment rules; TTA_GEMM_TMA=0
  // keeps the cp.async kernel below (comparison / debugging)
  static int use_tma = -1;
  if (use_tma < 0) {
    const char* e = getenv("TTA_GEMM_TMA");
    use_tma = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_tma && gemm_tma_eligible(a, lda, b, ldb, c, ldc, M, N, K, out_fp32))
    return gemm_tma_launch(a, lda, b, ldb, c, ldc, M, N, K, bias, out_fp32, st);
  // tile width: 64 when it wastes fewer padded columns than 128
  const int waste128 = ((N + 127) / 128) * 128 - N;
  const int waste64 = ((N + 63) / 64) * 64 - N;
  const bool use64 = waste64 < waste128;
  if (use64)
    return out_fp32 ? launch_tc<64, true>(a, lda, b, ldb, c, ldc, M, N, K, bias, st)
                    : launch_tc<64, false>(a, lda, b, ldb, c, ldc, M, N, K, bias, st);
  return out_fp32 ? launch_tc<128, true>(a, lda, b, ldb, c, ldc, M, N, K, bias, st)
                  : launch_tc<128, false>(a, lda, b, ldb, c, ldc, M, N, K, bias, st);
}
