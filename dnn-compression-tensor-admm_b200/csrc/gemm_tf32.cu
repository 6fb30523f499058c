// Batched fp32-accurate GEMM on the 5th-generation tensor cores: 3xTF32 split on tcgen05 + TMEM with a fresh
// accumulator for every 32 reduction indices.
//   C[M,N] = (A[M,K] * B[K,N]) .* colscale[N],  A(i,k) at a + i*sai + k*sak,  B(k,j) at b + k*sbk + j*sbj
// (the tta_gemm_task operator of include/tta.h: projection side of every truncated SVD -- ttd.py:21-26 --, the tt2ten
// reconstruction chain -- ttd.py:39-40 --, the Tucker mode products behind admm.py:116-117).
//
// Every fp32 operand value v is split as v = hi + lo with hi = tf32(v), lo = tf32(v - hi); a k-block contributes
// A_lo B_hi + A_hi B_lo + A_hi B_hi (12 MMAs of 8 reduction indices) to a ZEROED TMEM accumulator.  The tensor core's
// fp32 accumulation truncates: its error grows linearly with the accumulation length (scripts/ubench/
// tf32_accum_error.py: 1.3e-7 after 32 indices = CUDA-core fp32 grade, 3.5e-6 after 512), which the small-gap TT
// projections amplify past the 1e-4 parity bar -- so the running sum lives in fp32 registers of the drain warps
// (round to nearest), not in TMEM.
//
// CTA = one 128 x (<= 128) output tile of one task; the operand with the longer index takes the 128 MMA rows (a
// projection carry = E^T A has M = r <= 130 rows and N up to 73 728 columns: it is computed as C^T = B^T A^T).
// 512 threads (16 warps = 4 per scheduler):
//   warps 0-5   producers: cp.async (16 bytes when the reduction index is contiguous and aligned, 4 bytes for any other
//               stride combination; zero fill past the edges) into a 3-stage ring of raw fp32 tiles, two k-blocks
//               ahead; then hi / lo split (tf32_split.cuh) of the oldest stage into one of two operand slots.
//   warp 7      TMEM allocator + tcgen05.mma issuer (kind::tf32), two 128-column accumulators used alternately.
//   warps 8-15  drain: tcgen05.ld -> FADD into 64 registers per thread (warp & 3 = TMEM lane quadrant); epilogue:
//               colscale, store (coalesced across lanes in the transposed orientation).
#include "tc_common.cuh"
#include "tf32_split.cuh"

namespace tta {

constexpr int kT3BM = 128;
constexpr int kT3BK = 32;          // 32 fp32 = one 128-byte swizzle row
constexpr int kT3Threads = 512;
constexpr int kT3Producers = 192;
constexpr int kT3MaxTasks = 192;
constexpr int kT3Stages = 3;
constexpr int kT3TileBytes = kT3BM * 128;              // 16 KB: 128 rows x 128 B
constexpr int kT3SlotBytes = 4 * kT3TileBytes;         // R_hi, R_lo, S_hi, S_lo
constexpr int kT3RawBytes = 2 * kT3TileBytes;          // raw R, raw S
constexpr int kT3Smem = 2 * kT3SlotBytes + kT3Stages * kT3RawBytes + 1024;
constexpr int kT3Chunks = (kT3BM * 8 + kT3Producers - 1) / kT3Producers;     // 6 chunks of 16 bytes per thread and tile

struct T3Table {
  int n_tasks;
  int total;
  int start[kT3MaxTasks + 1];
  unsigned char transposed[kT3MaxTasks];     // 1: the columns of C take the 128 MMA rows
};

__device__ __forceinline__ void t3_cp16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void t3_cp4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// Issue the copies of rows [0, nrows) x 32 reduction indices of an operand whose element (r, k) lives at
// base + r*sr + k*sk into a raw tile (rows >= rmax - r0 and indices >= K are zero filled).
//   sk == 1 (reduction index contiguous) or generic strides: raw element (row, k) at row * 128 + k * 4
//   sr == 1 (operand index contiguous):                       raw element (row, k) at k * 512 + row * 4
__device__ __forceinline__ void t3_issue(uint32_t raw, const float* __restrict__ base, int64_t sr, int64_t sk, int r0, int rmax,
                                         int nrows, int k0, int K, int tid) {
  const int items = nrows * 8;
  const bool rowfast = (sr == 1 && sk != 1);
  if (!rowfast) {
    const bool vec = (sk == 1) && ((sr & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
#pragma unroll
    for (int i = 0; i < kT3Chunks; ++i) {
      const int c = tid + i * kT3Producers;
      if (c < items) {
        const int row = c >> 3, ch = c & 7;
        const int gr = r0 + row, gk = k0 + ch * 4;
        const bool rok = gr < rmax;
        const uint32_t dst = raw + (uint32_t)row * 128u + (uint32_t)ch * 16u;
        int left = rok ? K - gk : 0;                       // reduction indices of this chunk that exist
        left = left < 0 ? 0 : (left > 4 ? 4 : left);
        const float* p = left ? base + (int64_t)gr * sr + (int64_t)gk * sk : base;
        if (vec) {
          t3_cp16(dst, p, (uint32_t)left * 4u);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) t3_cp4(dst + 4u * e, e < left ? p + e * sk : base, e < left ? 4u : 0u);
        }
      }
    }
  } else if (((sk & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
    // operand index contiguous and aligned: 16 bytes = four consecutive rows at one reduction index
    const int quads = nrows >> 2;
#pragma unroll
    for (int i = 0; i < kT3Chunks; ++i) {
      const int c = tid + i * kT3Producers;
      if (c < items) {
        const int rq = c % quads, kk = c / quads;
        const int gr = r0 + rq * 4, gk = k0 + kk;
        int left = gk < K ? rmax - gr : 0;                 // rows of this chunk that exist
        left = left < 0 ? 0 : (left > 4 ? 4 : left);
        t3_cp16(raw + (uint32_t)kk * 512u + (uint32_t)rq * 16u, left ? base + gr + (int64_t)gk * sk : base, (uint32_t)left * 4u);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < kT3Chunks; ++i) {
      const int c = tid + i * kT3Producers;
      if (c < items) {
        const int row = c % nrows, ch = c / nrows;
        const int gr = r0 + row, gk = k0 + ch * 4;
        const bool rok = gr < rmax;
        const uint32_t dst = raw + (uint32_t)(ch * 4) * 512u + (uint32_t)row * 4u;
        const float* p = base + gr + (int64_t)gk * sk;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = rok && gk + e < K;
          t3_cp4(dst + 512u * e, ok ? p + e * sk : base, ok ? 4u : 0u);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kT3Threads, 1)
    gemm_tf32x3_kernel(const tta_gemm_task* __restrict__ tasks, const __grid_constant__ T3Table tab) {
  extern __shared__ __align__(1024) uint8_t t3_smem_raw[];
  const uint32_t smem = (tc::smem_u32(t3_smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t bar0 = tc::smem_u32(bars);
  const uint32_t op_full0 = bar0, op_empty0 = bar0 + 16, acc_full0 = bar0 + 32, acc_empty0 = bar0 + 48;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // which tile of which task
  int lo = 0, hi = tab.n_tasks;
  const int item = blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.start[mid] <= item) lo = mid; else hi = mid;
  }
  const tta_gemm_task tk = tasks[lo];
  const bool tr = tab.transposed[lo] != 0;
  // R = the operand on the 128 MMA rows, S = the operand on the MMA columns
  const float* rbase = tr ? tk.b : tk.a;
  const float* sbase = tr ? tk.a : tk.b;
  const int64_t r_sr = tr ? tk.sbj : tk.sai, r_sk = tr ? tk.sbk : tk.sak;
  const int64_t s_sr = tr ? tk.sai : tk.sbj, s_sk = tr ? tk.sak : tk.sbk;
  const int rcount = tr ? tk.N : tk.M, scount = tr ? tk.M : tk.N;
  const int tiles_s = (scount + kT3BM - 1) / kT3BM;
  const int local = item - tab.start[lo];
  const int r0 = (local / tiles_s) * kT3BM;
  const int s0 = (local % tiles_s) * kT3BM;
  const int nq = (tk.K + kT3BK - 1) / kT3BK;
  int ncols = scount - s0;
  if (ncols > kT3BM) ncols = kT3BM;
  const int N = (ncols + 15) & ~15;                      // MMA N

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(op_full0 + 8 * s, 6);
      tc::mbar_init(op_empty0 + 8 * s, 1);
      tc::mbar_init(acc_full0 + 8 * s, 1);
      tc::mbar_init(acc_empty0 + 8 * s, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base_smem)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t ring = smem + 2 * kT3SlotBytes;

  if (warp < 6) {
    // ------------------------------ producers ------------------------------
    const bool r_rowfast = (r_sr == 1 && r_sk != 1), s_rowfast = (s_sr == 1 && s_sk != 1);
    auto issue = [&](int kb) {
      const uint32_t raw = ring + (uint32_t)(kb % kT3Stages) * kT3RawBytes;
      t3_issue(raw, rbase, r_sr, r_sk, r0, rcount, kT3BM, kb * kT3BK, tk.K, tid);
      t3_issue(raw + kT3TileBytes, sbase, s_sr, s_sk, s0, scount, N, kb * kT3BK, tk.K, tid);
    };
    for (int kb = 0; kb < kT3Stages - 1; ++kb) {
      if (kb < nq) issue(kb);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int q = 0; q < nq; ++q) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kT3Stages - 2) : "memory");       // this thread's copies of k-block q
      asm volatile("bar.sync 1, %0;" ::"n"(kT3Producers) : "memory");                // everybody's; and the split of q - 1 is over
      if (q + kT3Stages - 1 < nq) issue(q + kT3Stages - 1);                          // into the stage k-block q - 1 used
      asm volatile("cp.async.commit_group;" ::: "memory");
      const int slot = q & 1;
      if (q >= 2) tc::mbar_wait(op_empty0 + 8 * slot, (uint32_t)(((q >> 1) - 1) & 1));
      const uint32_t raw = ring + (uint32_t)(q % kT3Stages) * kT3RawBytes;
      const uint32_t ops = smem + (uint32_t)slot * kT3SlotBytes;
      tf32::transform<kT3Producers>(raw, 0, false, ops, ops + kT3TileBytes, kT3BM, r_rowfast ? 512u : 128u, r_rowfast, tid);
      tf32::transform<kT3Producers>(raw + kT3TileBytes, 0, false, ops + 2 * kT3TileBytes, ops + 3 * kT3TileBytes, N,
                                    s_rowfast ? 512u : 128u, s_rowfast, tid);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(op_full0 + 8 * slot);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == 7) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = tf32::idesc(N);
    for (int q = 0; q < nq; ++q) {
      const int buf = q & 1;
      const uint32_t ops = smem + (uint32_t)buf * kT3SlotBytes;
      const uint64_t dr_hi = tc::umma_desc_sw128(ops), dr_lo = tc::umma_desc_sw128(ops + kT3TileBytes);
      const uint64_t ds_hi = tc::umma_desc_sw128(ops + 2 * kT3TileBytes), ds_lo = tc::umma_desc_sw128(ops + 3 * kT3TileBytes);
      tc::mbar_wait(op_full0 + 8 * buf, (uint32_t)((q >> 1) & 1));
      if (q >= 2) tc::mbar_wait(acc_empty0 + 8 * buf, (uint32_t)(((q >> 1) - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tc::elect_one()) {
        tf32::mma_kblock(tmem_base + (uint32_t)buf * kT3BM, dr_hi, dr_lo, ds_hi, ds_lo, idesc);
        tc::umma_commit(op_empty0 + 8 * buf);
        tc::umma_commit(acc_full0 + 8 * buf);
      }
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ------------------------------ drain + epilogue ------------------------------
    const int quad = warp & 3, half = (warp - 8) >> 2;
    const int cbase = half * 64;
    float acc[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = 0.f;
    for (int q = 0; q < nq; ++q) {
      const int buf = q & 1;
      tc::mbar_wait(acc_full0 + 8 * buf, (uint32_t)((q >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kT3BM + cbase);
      tf32::drain64(taddr, cbase, N, acc);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc_empty0 + 8 * buf);
    }
    const int gr = r0 + quad * 32 + lane;                 // index on the R side
    const int nc = ncols - cbase;                         // columns (S side) of this warp that exist
    if (gr < rcount && nc > 0) {
      if (!tr) {
        // C[gr][s0 + cbase + c]: 64 consecutive columns per thread
        float* crow = tk.c + (int64_t)gr * tk.ldc + s0 + cbase;
        const float* cs = tk.colscale ? tk.colscale + s0 + cbase : nullptr;
        const bool vec = ((tk.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(tk.c) & 15) == 0);
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          if (c < nc) {
            float o[4] = {acc[c], acc[c + 1], acc[c + 2], acc[c + 3]};
            if (cs) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (c + j < nc) o[j] *= __ldg(cs + c + j);
            }
            if (vec && c + 4 <= nc) {
              *reinterpret_cast<float4*>(crow + c) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (c + j < nc) crow[c + j] = o[j];
            }
          }
        }
      } else {
        // transposed tile: C[s0 + cbase + c][gr], lanes = consecutive columns of C
        const float sc = tk.colscale ? __ldg(tk.colscale + gr) : 1.f;
        float* ccol = tk.c + (int64_t)(s0 + cbase) * tk.ldc + gr;
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c < nc) ccol[(int64_t)c * tk.ldc] = acc[c] * sc;
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 7) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// A task goes to the tensor cores when it is big enough to pay: measured on B200 (scripts/bench_gemm_shapes.py), one task
// per launch, L2 cold: 0.15 GFLOP (2304 x 128 x 256) 25 us against 21 us on the CUDA cores, 0.46 GFLOP (the step-2
// projection of a ResNet-50 layer4 3x3 convolution, 105 x 4608 x 480) 41 against 33 us, 1.2 GFLOP (4608 x 256 x 512) 38
// against 50 us, 9.7 GFLOP (9216 x 512 x 1024) 130 against 325 us, 77 GFLOP (18432 x 1024 x 2048) 0.90 against 2.48 ms
// (85 against 31 TFLOP/s).  Below ~0.6 GFLOP a task is a handful of k-blocks on a fraction of the SMs and the fixed
// costs of the pipeline (TMEM allocation, ring fill, 12 instruction-bound staging chunks per thread and k-block) decide.
// mode 1: by size; mode 2 (tests): every task the kernel can serve.
static int g_t3_mode = 1;
void gemm_tf32x3_set_mode(int mode) { g_t3_mode = mode; }
int gemm_tf32x3_mode() { return g_t3_mode; }
bool gemm_tf32x3_eligible(const tta_gemm_task& tk) {
  if (g_t3_mode == 0 || tk.M < 1 || tk.N < 1 || tk.K < 1) return false;
  if (g_t3_mode == 2) return true;
  const int big = tk.M > tk.N ? tk.M : tk.N, small = tk.M > tk.N ? tk.N : tk.M;
  return big >= 128 && small >= 16 && tk.K >= 64 && 2.0 * tk.M * tk.N * tk.K >= 6e8;
}

// The longer of the two output indices takes the 128 MMA rows (the other one is padded to a multiple of 16 only).
static bool t3_transposed(const tta_gemm_task& tk) {
  auto cost = [](int r, int s) { return (int64_t)((r + 127) / 128 * 128) * ((s + 15) / 16 * 16); };
  return cost(tk.N, tk.M) < cost(tk.M, tk.N);
}

// Enqueue the tensor-core tiles of tasks [0, cnt); ineligible tasks contribute none.
int gemm_tf32x3_run(const tta_gemm_task* tasks_dev, const tta_gemm_task* tasks_host, int cnt, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT3Smem),
                        "gemm_tf32x3 smem attribute");
    if (rc) return rc;
    attr_set = true;
  }
  T3Table tab;
  tab.n_tasks = cnt;
  int64_t total = 0;
  for (int t = 0; t < cnt; ++t) {
    const tta_gemm_task& tk = tasks_host[t];
    tab.start[t] = (int)total;
    tab.transposed[t] = 0;
    if (!gemm_tf32x3_eligible(tk)) continue;
    const bool tr = t3_transposed(tk);
    tab.transposed[t] = tr ? 1 : 0;
    const int rc_ = tr ? tk.N : tk.M, sc_ = tr ? tk.M : tk.N;
    total += (int64_t)((rc_ + kT3BM - 1) / kT3BM) * ((sc_ + kT3BM - 1) / kT3BM);
    if (total > 0x7fffffff) {
      set_error("gemm_tf32x3: too many tiles");
      return TTA_E_INVALID;
    }
  }
  for (int t = cnt; t < kT3MaxTasks; ++t) tab.transposed[t] = 0;
  tab.start[cnt] = (int)total;
  tab.total = (int)total;
  if (total == 0) return TTA_OK;
  gemm_tf32x3_kernel<<<(int)total, kT3Threads, kT3Smem, st>>>(tasks_dev, tab);
  TTA_CHECK_LAUNCH("gemm_tf32x3 launch");
  return TTA_OK;
}

}  // namespace tta
