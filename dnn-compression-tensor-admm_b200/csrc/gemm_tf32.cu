// Batched fp32-accurate GEMM on the 5th-generation tensor cores: 3xTF32 split on tcgen05 + TMEM.
//   C[M,N] = (A[M,K] * B[K,N]) .* colscale[N],  A(i,k) at a + i*sai + k*sak,  B(k,j) at b + k*sbk + j*sbj
// (the tta_gemm_task operator of include/tta.h: projection side of every truncated SVD, the tt2ten
// reconstruction chain, the Tucker mode products).
//
// Every fp32 operand value v is split as v = hi + lo with hi = tf32(v), lo = tf32(v - hi); the product
// is accumulated as  A_hi B_hi + A_lo B_hi + A_hi B_lo  in the fp32 TMEM accumulator (the lo*lo term is
// below fp32 rounding), which recovers ~22 bits of each product where a single TF32 pass keeps 11.
//
// CTA = one 128 x BN output tile of one task (task tiles are concatenated as in gemm.cu).  160 threads:
//   warps 0-3  producers: global (any stride combination; 128-bit loads when the reduction index is the
//              fast one) -> registers -> hi / lo split -> shared memory, rows of 32 fp32 = 128 bytes in the
//              SWIZZLE_128B pattern of the UMMA descriptors, 3-stage ring; then the epilogue (each warp owns
//              one TMEM lane quadrant): tcgen05.ld -> colscale -> global.
//   warp 4     TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, 12 MMAs per k-block);
//              tcgen05.commit releases ring slots / publishes the accumulator through mbarriers.
#include "tta_common.cuh"

namespace tta {

constexpr int kT3BM = 128;
constexpr int kT3BK = 32;          // 32 fp32 = one 128-byte swizzle row
constexpr int kT3Stages = 3;
constexpr int kT3Threads = 160;
constexpr int kT3MaxTasks = 192;

struct T3Table {
  int n_tasks;
  int total;
  int start[kT3MaxTasks + 1];
};

__device__ __forceinline__ uint32_t t3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t3_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t3_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void t3_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t3_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void t3_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "T3_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni T3_WAIT_DONE;\n"
      "bra.uni T3_WAIT_LOOP;\n"
      "T3_WAIT_DONE:\n"
      "}\n" ::"r"(t3_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t t3_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor: D = fp32, A = B = tf32, both K-major, M = 128, N = bn
__host__ __device__ constexpr uint32_t t3_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kT3BM >> 4) << 24);
}

__device__ __forceinline__ float t3_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// One 128-byte row chunk (4 fp32 of reduction indices k0+4*ch .. +3 of tile row `row`) -> hi / lo tiles.
__device__ __forceinline__ void t3_store_chunk(uint8_t* hi_tile, uint8_t* lo_tile, int row, int ch, float4 v) {
  const float4 h = make_float4(t3_tf32(v.x), t3_tf32(v.y), t3_tf32(v.z), t3_tf32(v.w));
  const float4 l = make_float4(t3_tf32(v.x - h.x), t3_tf32(v.y - h.y), t3_tf32(v.z - h.z), t3_tf32(v.w - h.w));
  const int off = row * 128 + ((ch ^ (row & 7)) << 4);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// Stage `rows` x 32 values of an operand whose element (r, k) lives at base + r*sr + k*sk.
template <int ROWS>
__device__ __forceinline__ void t3_produce(uint8_t* hi_tile, uint8_t* lo_tile, const float* __restrict__ base,
                                           int64_t sr, int64_t sk, int r0, int rmax, int k0, int K, int tid) {
  if (sk == 1) {
    // reduction index contiguous: 8 lanes cover one 128-byte row
    const bool vec = ((sr & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
#pragma unroll
    for (int i = 0; i < (ROWS * 8) / 128; ++i) {
      const int c = i * 128 + tid;
      const int row = c >> 3, ch = c & 7;
      const int gr = r0 + row, gk = k0 + ch * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rmax && gk < K) {
        const float* p = base + (int64_t)gr * sr + gk;
        if (vec && gk + 3 < K) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          v.x = __ldg(p);
          if (gk + 1 < K) v.y = __ldg(p + 1);
          if (gk + 2 < K) v.z = __ldg(p + 2);
          if (gk + 3 < K) v.w = __ldg(p + 3);
        }
      }
      t3_store_chunk(hi_tile, lo_tile, row, ch, v);
    }
  } else {
    // tile-row index varies fastest in memory (or generic strides): consecutive lanes take consecutive rows
#pragma unroll
    for (int i = 0; i < (ROWS * 8) / 128; ++i) {
      const int c = i * 128 + tid;
      const int row = c % ROWS, ch = c / ROWS;
      const int gr = r0 + row, gk = k0 + ch * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rmax) {
        const float* p = base + (int64_t)gr * sr + (int64_t)gk * sk;
        if (gk < K) v.x = __ldg(p);
        if (gk + 1 < K) v.y = __ldg(p + sk);
        if (gk + 2 < K) v.z = __ldg(p + 2 * sk);
        if (gk + 3 < K) v.w = __ldg(p + 3 * sk);
      }
      t3_store_chunk(hi_tile, lo_tile, row, ch, v);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kT3Threads, 1)
    gemm_tf32x3_kernel(const tta_gemm_task* __restrict__ tasks, const __grid_constant__ T3Table tab) {
  extern __shared__ __align__(1024) uint8_t t3_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(t3_smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kABytes = kT3BM * 128;
  constexpr int kBBytes = BN * 128;
  constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;     // A_hi, A_lo, B_hi, B_lo
  __shared__ uint64_t full_bar[kT3Stages];
  __shared__ uint64_t empty_bar[kT3Stages];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // which tile of which task
  int lo = 0, hi = tab.n_tasks;
  const int item = blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.start[mid] <= item) lo = mid; else hi = mid;
  }
  const tta_gemm_task tk = tasks[lo];
  const int tiles_n = (tk.N + BN - 1) / BN;
  const int local = item - tab.start[lo];
  const int m0 = (local / tiles_n) * kT3BM;
  const int n0 = (local % tiles_n) * BN;
  const int nkb = (tk.K + kT3BK - 1) / kT3BK;

  if (tid == 0) {
    for (int s = 0; s < kT3Stages; ++s) {
      t3_mbar_init(&full_bar[s], 128);
      t3_mbar_init(&empty_bar[s], 1);
    }
    t3_mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t3_smem_u32(&tmem_base_smem)),
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
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kT3Stages;
      if (kb >= kT3Stages) t3_mbar_wait(&empty_bar[s], ((kb / kT3Stages) - 1) & 1);
      uint8_t* a_hi = smem + (size_t)s * kStageBytes;
      uint8_t* a_lo = a_hi + kABytes;
      uint8_t* b_hi = a_lo + kABytes;
      uint8_t* b_lo = b_hi + kBBytes;
      const int k0 = kb * kT3BK;
      t3_produce<kT3BM>(a_hi, a_lo, tk.a, tk.sai, tk.sak, m0, tk.M, k0, tk.K, tid);
      t3_produce<BN>(b_hi, b_lo, tk.b, tk.sbj, tk.sbk, n0, tk.N, k0, tk.K, tid);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
      t3_mbar_arrive(&full_bar[s]);
    }
    // ------------------------------ epilogue ------------------------------
    t3_mbar_wait(&accum_bar, 0);
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
      if (gm < tk.M) {
        float* crow = tk.c + (int64_t)gm * tk.ldc;
        const int gn0 = n0 + c0;
        if (gn0 + 16 <= tk.N && ((tk.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(tk.c) & 15) == 0) && ((gn0 & 3) == 0)) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            if (tk.colscale) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(tk.colscale + gn0 + j));
              o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w;
            }
            *reinterpret_cast<float4*>(crow + gn0 + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int gn = gn0 + j;
            if (gn < tk.N) crow[gn] = __uint_as_float(v[j]) * (tk.colscale ? __ldg(tk.colscale + gn) : 1.f);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = t3_idesc(BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kT3Stages;
        t3_mbar_wait(&full_bar[s], (kb / kT3Stages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa_hi = t3_smem_u32(smem + (size_t)s * kStageBytes);
        const uint32_t sa_lo = sa_hi + kABytes;
        const uint32_t sb_hi = sa_lo + kABytes;
        const uint32_t sb_lo = sb_hi + kBBytes;
        const uint64_t da_hi = t3_desc_sw128(sa_hi), da_lo = t3_desc_sw128(sa_lo);
        const uint64_t db_hi = t3_desc_sw128(sb_hi), db_lo = t3_desc_sw128(sb_lo);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          // small terms first: A_lo B_hi, A_hi B_lo, then A_hi B_hi
          const uint64_t da = pass == 0 ? da_lo : da_hi;
          const uint64_t db = pass == 1 ? db_lo : db_hi;
#pragma unroll
          for (int k = 0; k < kT3BK / 8; ++k) {
            const uint32_t accumulate = (kb > 0 || pass > 0 || k > 0) ? 1u : 0u;
            // advance 8 tf32 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
                "}\n" ::"r"(tmem_base),
                "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(accumulate)
                : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         t3_smem_u32(&empty_bar[s]))
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       t3_smem_u32(&accum_bar))
                   : "memory");
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// A task goes to the tensor cores when its tile grid is not mostly padding.
bool gemm_tf32x3_eligible(const tta_gemm_task& tk) { return tk.M >= 48 && tk.N >= 24 && tk.K >= 8; }
static int t3_bn(const tta_gemm_task& tk) {
  const int waste128 = ((tk.N + 127) / 128) * 128 - tk.N;
  const int waste64 = ((tk.N + 63) / 64) * 64 - tk.N;
  return waste64 < waste128 ? 64 : 128;
}

template <int BN>
static int t3_launch(const tta_gemm_task* tasks_dev, const T3Table& tab, cudaStream_t st) {
  constexpr size_t smem = (size_t)kT3Stages * (2 * kT3BM * 128 + 2 * BN * 128) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "gemm_tf32x3 smem attribute");
    if (rc) return rc;
    attr_set = true;
  }
  gemm_tf32x3_kernel<BN><<<tab.total, kT3Threads, smem, st>>>(tasks_dev, tab);
  TTA_CHECK_LAUNCH("gemm_tf32x3 launch");
  return TTA_OK;
}

// Enqueue the tensor-core tiles of tasks [first, first + cnt): two launches (BN = 64 and BN = 128), a task
// contributes tiles to the launch whose tile width wastes fewer columns.  Ineligible tasks contribute none.
int gemm_tf32x3_run(const tta_gemm_task* tasks_dev, const tta_gemm_task* tasks_host, int cnt, cudaStream_t st) {
  for (int pass = 0; pass < 2; ++pass) {
    const int bn = pass ? 128 : 64;
    T3Table tab;
    tab.n_tasks = cnt;
    int64_t total = 0;
    for (int t = 0; t < cnt; ++t) {
      const tta_gemm_task& tk = tasks_host[t];
      tab.start[t] = (int)total;
      if (!gemm_tf32x3_eligible(tk) || t3_bn(tk) != bn) continue;
      total += (int64_t)((tk.M + kT3BM - 1) / kT3BM) * ((tk.N + bn - 1) / bn);
      if (total > 0x7fffffff) {
        set_error("gemm_tf32x3: too many tiles");
        return TTA_E_INVALID;
      }
    }
    tab.start[cnt] = (int)total;
    tab.total = (int)total;
    if (total == 0) continue;
    const int rc = pass ? t3_launch<128>(tasks_dev, tab, st) : t3_launch<64>(tasks_dev, tab, st);
    if (rc) return rc;
  }
  return TTA_OK;
}

}  // namespace tta
