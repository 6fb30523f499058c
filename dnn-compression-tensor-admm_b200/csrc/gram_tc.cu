// Gram matrices on the 5th-generation tensor cores: TMA-fed tcgen05 3xTF32 with a per-k-block accumulator flush.
//   G(i,j) = sum_c v(i,c) v(j,c),   v = a (+ a2)          (tta_gram_task, nb == 1)
// Replaces the input side of ttd.py:16-17 (numpy.linalg.svd of an unfolding) and, through `a2`, the
// V = W + U of admm.py:45 that feeds it: the first TT step reads W and U IN PLACE (two tensor maps over the
// parameter / dual tensors, summed in shared memory) -- the order of the reduction index does not matter to a
// Gram matrix, so the (O, I, KK) -> (O, KK, I) permute of admm.py:96 is not needed to form it.
//
// Accuracy (scripts/ubench/tf32_accum_error.py, B200): the TMEM accumulator truncates, the error of a 3xTF32 product
// grows linearly with the reduction length (1.3e-7 at K = 32, 3.5e-6 at K = 512, 5.7e-5 at K = 8192), and the
// small-gap projections amplify it past the 1e-4 parity bar.  So every k-block of 32 reduction indices (12 MMAs)
// gets a fresh accumulator; the drain warps pull it out of TMEM and add it to fp32 registers with round-to-nearest;
// a CTA covers a slice of the reduction range, gram_finish sums the slices in fp64.
//
// One CTA = one 128 x 128 lower-triangular tile x one slice.  512 threads (16 warps = 4 per scheduler, 128 registers):
//   warp 6      TMA producer: boxes of `rb` operand rows x 4096 / rb reduction indices (16 KB) of a and a2 for the
//               row block and (off the diagonal) the column block, 2-stage ring, mbarrier expect_tx.
//   warps 0-5   transform: t = a + a2, hi = tf32(t), lo = tf32(t - hi) -> K-major SWIZZLE_128B operand tiles
//               (transposing when the operand index is the contiguous one: column Grams A^T A).
//   warp 7      TMEM allocator + tcgen05.mma issuer (kind::tf32; lo*hi, hi*lo, hi*hi per 8 indices), two 128-column
//               accumulators used alternately; tcgen05.commit hands the operand slot back and publishes the block.
//   warps 8-15  drain: tcgen05.ld -> FADD into 64 registers per thread (warp & 3 = TMEM lane quadrant, two warps per
//               quadrant share the 128 columns); at the end the fp32 partial tile goes to part[split][k][k].
#include <cstring>
#include <mutex>
#include <vector>

#include "tc_common.cuh"
#include "tf32_split.cuh"

namespace tta {

constexpr int kGtThreads = 512;
constexpr int kGtXformThreads = 192;
constexpr int kGtMaxTasks = 40;
constexpr int kGtTile = 128;
constexpr int kGtStageElems = 4096;                 // fp32 per (source, operand) box: 16 KB
constexpr int kGtBoxBytes = kGtStageElems * 4;
constexpr int kGtOpBytes = 4 * kGtBoxBytes;         // A_hi, A_lo, B_hi, B_lo: 128 rows x 128 B each
constexpr int kGtStageBytes = 4 * kGtBoxBytes;      // A_a, A_a2, B_a, B_a2
constexpr int kGtSmem = kGtOpBytes + 2 * kGtStageBytes + 1024;

struct GtTask {
  float* part;      // fp32 partials [nsplit][k][k]
  int k, red;       // matrix size, reduction length
  int nsplit;       // slices of the reduction range
  int slice;        // reduction indices per slice (multiple of 128)
  int rb;           // operand rows per box: 32 / 64 / 128
  int colmajor;     // 1: the operand index is the contiguous one (A^T A of a row-major A)
  int nsrc;         // 1 or 2 (a2 given)
  int ntile;        // tiles per side
};

struct GtParams {
  int n_tasks;
  int total;
  int start[kGtMaxTasks + 1];
  GtTask task[kGtMaxTasks];
  CUtensorMap maps[kGtMaxTasks][2];
};

__global__ void __launch_bounds__(kGtThreads, 1) gram_tc_kernel(const __grid_constant__ GtParams P) {
  extern __shared__ __align__(1024) uint8_t gt_smem_raw[];
  const uint32_t smem = (tc::smem_u32(gt_smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[10];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t bar0 = tc::smem_u32(bars);
  // barrier indices
  const uint32_t st_full0 = bar0, st_empty0 = bar0 + 16, op_full = bar0 + 32, op_empty = bar0 + 40, acc_full0 = bar0 + 48,
                 acc_empty0 = bar0 + 64;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // which tile / slice of which task
  int t = 0;
  {
    int lo = 0, hi = P.n_tasks;
    const int item = blockIdx.x;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (P.start[mid] <= item) lo = mid; else hi = mid;
    }
    t = lo;
  }
  const GtTask tk = P.task[t];
  const int npairs = tk.ntile * (tk.ntile + 1) / 2;
  const int local = blockIdx.x - P.start[t];
  const int split = local / npairs;
  const int pr = local - split * npairs;
  int ti = 0;
  while ((ti + 1) * (ti + 2) / 2 <= pr) ++ti;            // ti >= tj, pair index = ti (ti + 1) / 2 + tj
  const int tj = pr - ti * (ti + 1) / 2;
  const int i0 = ti * kGtTile, j0 = tj * kGtTile;
  const bool diag = (ti == tj);
  const int rb = tk.rb, kcols = kGtStageElems / rb, SUB = kcols / 32;
  const int kbeg = split * tk.slice;
  int kend = kbeg + tk.slice;
  if (kend > tk.red) kend = tk.red;
  const int nq = kend > kbeg ? (kend - kbeg + 31) / 32 : 0;      // k-blocks of this slice
  const int nst = (nq + SUB - 1) / SUB;                         // TMA stages
  int ncols = tk.k - j0;                                        // MMA N: columns of the tile that exist, in 16s
  if (ncols > kGtTile) ncols = kGtTile;
  const int N = (ncols + 15) & ~15;
  const bool has2 = tk.nsrc == 2;
  const bool colmajor = tk.colmajor != 0;

  if (tid == 0) {
    tc::mbar_init(st_full0, 1);
    tc::mbar_init(st_full0 + 8, 1);
    tc::mbar_init(st_empty0, 6);
    tc::mbar_init(st_empty0 + 8, 6);
    tc::mbar_init(op_full, 6);
    tc::mbar_init(op_empty, 1);
    tc::mbar_init(acc_full0, 1);
    tc::mbar_init(acc_full0 + 8, 1);
    tc::mbar_init(acc_empty0, 8);
    tc::mbar_init(acc_empty0 + 8, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base_smem)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < 6 && rb < kGtTile) {
    // operand rows past the box are never written by the transform: they must read as zero
    for (uint32_t o = (uint32_t)tid * 16u; o < (uint32_t)kGtOpBytes; o += (uint32_t)kGtXformThreads * 16u) tc::sts128(smem + o, 0u, 0u, 0u, 0u);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  const uint32_t a_hi = smem, a_lo = smem + kGtBoxBytes, b_hi = smem + 2 * kGtBoxBytes, b_lo = smem + 3 * kGtBoxBytes;
  const uint32_t stage0 = smem + kGtOpBytes;

  if (warp == 6) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      const CUtensorMap* m0 = &P.maps[t][0];
      const CUtensorMap* m1 = &P.maps[t][1];
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m0)) : "memory");
      if (has2) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m1)) : "memory");
      const uint32_t bytes = (uint32_t)kGtBoxBytes * (uint32_t)tk.nsrc * (diag ? 1u : 2u);
      for (int s = 0; s < nst; ++s) {
        const int slot = s & 1;
        if (s >= 2) tc::mbar_wait(st_empty0 + 8 * slot, (uint32_t)(((s >> 1) - 1) & 1));
        const uint32_t bar = st_full0 + 8 * slot;
        tc::mbar_expect_tx(bar, bytes);
        const uint32_t base = stage0 + (uint32_t)slot * kGtStageBytes;
        const int kc = kbeg + s * kcols;
        if (!colmajor) {
          tc::tma_load_2d(base, m0, bar, kc, i0);
          if (has2) tc::tma_load_2d(base + kGtBoxBytes, m1, bar, kc, i0);
          if (!diag) {
            tc::tma_load_2d(base + 2 * kGtBoxBytes, m0, bar, kc, j0);
            if (has2) tc::tma_load_2d(base + 3 * kGtBoxBytes, m1, bar, kc, j0);
          }
        } else {
          tc::tma_load_2d(base, m0, bar, i0, kc);
          if (has2) tc::tma_load_2d(base + kGtBoxBytes, m1, bar, i0, kc);
          if (!diag) {
            tc::tma_load_2d(base + 2 * kGtBoxBytes, m0, bar, j0, kc);
            if (has2) tc::tma_load_2d(base + 3 * kGtBoxBytes, m1, bar, j0, kc);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ------------------------------ transform ------------------------------
    for (int q = 0; q < nq; ++q) {
      const int s = q / SUB, sub = q - s * SUB, slot = s & 1;
      if (sub == 0) tc::mbar_wait(st_full0 + 8 * slot, (uint32_t)((s >> 1) & 1));
      if (q >= 1) tc::mbar_wait(op_empty, (uint32_t)((q - 1) & 1));
      // box = rb rows x kcols reduction indices (pitch kcols * 4 bytes), or kcols reduction rows x rb operand indices
      // (pitch rb * 4 bytes, transposed while splitting); `sub` = which 32 reduction indices of the box
      const uint32_t pitch = (uint32_t)(colmajor ? rb : kcols) * 4u;
      const uint32_t base = stage0 + (uint32_t)slot * kGtStageBytes + (colmajor ? (uint32_t)(sub * 32) * pitch : (uint32_t)sub * 128u);
      tf32::transform<kGtXformThreads>(base, base + kGtBoxBytes, has2, a_hi, a_lo, rb, pitch, colmajor, tid);
      if (!diag)
        tf32::transform<kGtXformThreads>(base + 2 * kGtBoxBytes, base + 3 * kGtBoxBytes, has2, b_hi, b_lo, rb, pitch, colmajor, tid);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy (UMMA)
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(op_full);
        if (sub == SUB - 1 || q == nq - 1) tc::mbar_arrive(st_empty0 + 8 * slot);
      }
    }
  } else if (warp == 7) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = tf32::idesc(N);
    const uint64_t da_hi = tc::umma_desc_sw128(a_hi), da_lo = tc::umma_desc_sw128(a_lo);
    const uint64_t db_hi = diag ? da_hi : tc::umma_desc_sw128(b_hi), db_lo = diag ? da_lo : tc::umma_desc_sw128(b_lo);
    for (int q = 0; q < nq; ++q) {
      const int buf = q & 1;
      tc::mbar_wait(op_full, (uint32_t)(q & 1));
      if (q >= 2) tc::mbar_wait(acc_empty0 + 8 * buf, (uint32_t)(((q >> 1) - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tc::elect_one()) {
        tf32::mma_kblock(tmem_base + (uint32_t)buf * kGtTile, da_hi, da_lo, db_hi, db_lo, idesc);
        tc::umma_commit(op_empty);
        tc::umma_commit(acc_full0 + 8 * buf);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------ drain ------------------------------
    const int quad = warp & 3, half = (warp - 8) >> 2;
    const int cbase = half * 64;                                   // this warp's columns of the tile
    float acc[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = 0.f;
    for (int q = 0; q < nq; ++q) {
      const int buf = q & 1;
      tc::mbar_wait(acc_full0 + 8 * buf, (uint32_t)((q >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kGtTile + cbase);
      tf32::drain64(taddr, cbase, N, acc);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc_empty0 + 8 * buf);
    }
    // partial tile -> part[split][gi][j0 + cbase ..]
    const int gi = i0 + quad * 32 + lane;
    if (gi < tk.k) {
      float* prow = tk.part + ((int64_t)split * tk.k + gi) * tk.k + j0 + cbase;
      const int nc = ncols - cbase;                                // columns of this warp that exist
      const bool vec = ((tk.k & 3) == 0);
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        if (c < nc) {
          if (vec && c + 4 <= nc) {
            *reinterpret_cast<float4*>(prow + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c + j < nc) prow[c + j] = acc[c + j];
          }
        }
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

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static bool g_gram_tc = true;

bool gram_tc_enabled() { return g_gram_tc; }
void gram_tc_enable(bool on) { g_gram_tc = on; }

// A task goes to the tensor cores when TMA can address its operand(s): one reduction run (nb == 1), one of the two
// indices contiguous, 16-byte aligned base and row pitch.  Same predicate on the device (gram_finish reads fp32
// partials for these tasks).
__host__ __device__ bool gram_tc_eligible(const tta_gram_task& tk) {
  if (tk.nb != 1 || tk.k < 2 || tk.nc < 32) return false;
  if ((reinterpret_cast<uintptr_t>(tk.a) & 15) || (reinterpret_cast<uintptr_t>(tk.a2) & 15)) return false;
  if (tk.sc == 1 && tk.si != 1) return (tk.si & 3) == 0 && tk.si >= tk.nc;       // row Gram A A^T
  if (tk.si == 1 && tk.sc != 1) return (tk.sc & 3) == 0 && tk.sc >= tk.k;        // column Gram A^T A
  return false;
}

static int gt_rows_per_box(int k) { return k <= 32 ? 32 : k <= 64 ? 64 : 128; }

namespace {
struct GtCacheEntry {
  std::vector<tta_gram_task> key;
  GtParams params;
};
std::mutex g_gt_mutex;
std::vector<GtCacheEntry*> g_gt_cache;

int gt_build(const tta_gram_task* host, int cnt, GtParams* out) {
  GtParams& P = *out;
  memset(&P, 0, sizeof(P));
  P.n_tasks = cnt;
  int64_t total = 0;
  for (int t = 0; t < cnt; ++t) {
    const tta_gram_task& tk = host[t];
    P.start[t] = (int)total;
    GtTask& g = P.task[t];
    g.part = reinterpret_cast<float*>(tk.part);
    g.k = tk.k;
    g.red = tk.nc;
    g.nsplit = tk.nsplit;
    int per = (tk.nc + tk.nsplit - 1) / tk.nsplit;
    g.slice = (per + 127) / 128 * 128;
    g.rb = gt_rows_per_box(tk.k);
    g.colmajor = (tk.si == 1) ? 1 : 0;
    g.nsrc = tk.a2 ? 2 : 1;
    g.ntile = (tk.k + kGtTile - 1) / kGtTile;
    const int kcols = kGtStageElems / g.rb;
    for (int s = 0; s < g.nsrc; ++s) {
      const float* base = s ? tk.a2 : tk.a;
      int rc;
      if (!g.colmajor)   // rows = operand index (k), cols = reduction index, pitch si
        rc = tc::make_map(&P.maps[t][s], base, tk.k, tk.nc, tk.si, g.rb, kcols, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                          CU_TENSOR_MAP_SWIZZLE_NONE);
      else               // rows = reduction index, cols = operand index (k), pitch sc
        rc = tc::make_map(&P.maps[t][s], base, tk.nc, tk.k, tk.sc, kcols, g.rb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                          CU_TENSOR_MAP_SWIZZLE_NONE);
      if (rc) return rc;
    }
    total += (int64_t)(g.ntile * (g.ntile + 1) / 2) * g.nsplit;
    if (total > 0x7fffffff) {
      set_error("gram_tc: too many tiles");
      return TTA_E_INVALID;
    }
  }
  P.start[cnt] = (int)total;
  P.total = (int)total;
  return TTA_OK;
}
}  // namespace

// Enqueue the tensor-core partial pass for `cnt` eligible tasks (host copies, contiguous).  Tensor maps are encoded
// once per distinct task list (the plans bake raw pointers into their tables and reuse them every update).
int gram_tc_run(const tta_gram_task* host, int cnt, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGtSmem),
                        "gram_tc smem attribute");
    if (rc) return rc;
    attr_set = true;
  }
  for (int first = 0; first < cnt; first += kGtMaxTasks) {
    const int n = (cnt - first) < kGtMaxTasks ? (cnt - first) : kGtMaxTasks;
    const GtParams* P = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_gt_mutex);
      for (GtCacheEntry* e : g_gt_cache)
        if ((int)e->key.size() == n && memcmp(e->key.data(), host + first, sizeof(tta_gram_task) * n) == 0) {
          P = &e->params;
          break;
        }
      if (!P) {
        if (g_gt_cache.size() >= 1024) {
          for (GtCacheEntry* e : g_gt_cache) delete e;
          g_gt_cache.clear();
        }
        GtCacheEntry* e = new GtCacheEntry;
        e->key.assign(host + first, host + first + n);
        const int rc = gt_build(host + first, n, &e->params);
        if (rc) {
          delete e;
          return rc;
        }
        g_gt_cache.push_back(e);
        P = &e->params;
      }
    }
    if (P->total == 0) continue;
    gram_tc_kernel<<<P->total, kGtThreads, kGtSmem, st>>>(*P);
    TTA_CHECK_LAUNCH("gram_tc launch");
  }
  return TTA_OK;
}

}  // namespace tta

extern "C" void tta_gram_enable_tc(int on) { tta::gram_tc_enable(on != 0); }
