// Gram matrices with fp64 accumulation:  G(i,j) = sum_{b,c} a(i,b,c) a(j,b,c).
// The truncated SVDs of ttd.py:17 / admm.py:131,143 and tensorly's partial_svd only need the
// dominant left (or right) singular subspace of an m x n unfolding; that subspace is the dominant
// eigenspace of the k x k Gram matrix, k = min(m, n).  Products of two fp32 values are exact in
// fp64, so G carries no rounding beyond the final fp32 store.
//
// Pass 1 (gram_partial): 64x64 lower-triangular output tiles x `nsplit` slices of the reduction
//   range; DFMA on 4x4 register tiles, operands staged through shared memory as doubles.
// Pass 2 (gram_finish): deterministic sum over slices, mirror, store as the eigensolver's column
//   state X (zero padded).
#include "tta_common.cuh"

namespace tta {

constexpr int kGramT = 64;   // output tile
constexpr int kGramBK = 16;  // reduction chunk
constexpr int kGramThreads = 256;
constexpr int kGramMaxTasks = 192;

struct GramTable {
  int n_tasks;
  int total;
  int start[kGramMaxTasks + 1];
};

// tile class of a problem: 0 = 32 x 32 (k <= 32), 1 = 64 x 64, 2 = 128 x 128 (k > 64 and a reduction long
// enough that the coarser tiling still yields enough work items)
__host__ __device__ inline int gram_class(const tta_gram_task& tk) {
  if (tk.k <= 32) return 0;
  if (tk.k <= kGramT) return 1;
  return ((int64_t)tk.nb * tk.nc >= 1536) ? 2 : 1;
}
__host__ __device__ inline int gram_tile(int cls) { return cls == 0 ? 32 : cls == 1 ? 64 : 128; }

__device__ __forceinline__ int64_t gram_off(const tta_gram_task& tk, int64_t rho) {
  if (tk.nb == 1) return rho * tk.sc;
  const int64_t b = rho / tk.nc;
  return b * tk.sb + (rho - b * tk.nc) * tk.sc;
}

// operand element: a (+ a2)
__device__ __forceinline__ float gram_ld(const tta_gram_task& tk, int64_t off) {
  const float v = __ldg(tk.a + off);
  return tk.a2 ? v + __ldg(tk.a2 + off) : v;
}

// tensor-core route (gram_tc.cu)
__host__ __device__ bool gram_tc_eligible(const tta_gram_task& tk);
bool gram_tc_enabled();
int gram_tc_run(const tta_gram_task* host, int cnt, cudaStream_t st);

struct GramFinishFlags {
  unsigned char tc[kGramMaxTasks];      // 1: `part` holds fp32 partials of 128 x 128 tiles (tensor-core pass)
};

// TS x TS output tiles, TT x TT outputs per thread: <64, 4> for 32 < k <= 64, <32, 2> for k <= 32 (the
// first and last TT steps of a k x k convolution have k = r <= 32 and a reduction length of up to
// 73 728: a 64-wide tile would spend 3/4 .. 15/16 of its DFMAs on padding).
template <int TS, int TT>
__global__ void __launch_bounds__(kGramThreads) gram_partial(const tta_gram_task* __restrict__ tasks,
                                                            const __grid_constant__ GramTable tab) {
  __shared__ __align__(16) double As[kGramBK][TS + 2];
  __shared__ __align__(16) double Bs[kGramBK][TS + 2];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int r0 = (warp & 1) * (TS / 2) + (lane >> 2) * TT;  // rows of this thread inside the tile
  const int c0 = (warp >> 1) * (TS / 4) + (lane & 3) * TT;  // cols
  constexpr int kLoads = TS * kGramBK / kGramThreads;
  constexpr int kRowBits = TS == 64 ? 6 : 5;

  for (int item = blockIdx.x; item < tab.total; item += gridDim.x) {
    int lo = 0, hi = tab.n_tasks;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tab.start[mid] <= item) lo = mid; else hi = mid;
    }
    const tta_gram_task tk = tasks[lo];
    const int T = (tk.k + TS - 1) / TS;
    const int npairs = T * (T + 1) / 2;
    const int local = item - tab.start[lo];
    const int split = local / npairs;
    int pr = local - split * npairs;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= pr) ++ti;  // ti >= tj, pair index = ti(ti+1)/2 + tj
    const int tj = pr - ti * (ti + 1) / 2;
    const int i0 = ti * TS, j0 = tj * TS;
    const bool diag = (ti == tj);

    const int64_t R = (int64_t)tk.nb * tk.nc;
    const int64_t per = (R + tk.nsplit - 1) / tk.nsplit;
    const int64_t rbeg = (int64_t)split * per;
    const int64_t rend = (rbeg + per) < R ? (rbeg + per) : R;
    const bool rowfast = (tk.si == 1);

    double acc[TT][TT];
#pragma unroll
    for (int i = 0; i < TT; ++i)
#pragma unroll
      for (int j = 0; j < TT; ++j) acc[i][j] = 0.0;

    // the global loads of chunk c + 1 are issued into registers before the DFMAs of chunk c (otherwise every
    // chunk pays a full load -> barrier -> compute -> barrier round trip: the kernel was latency bound at
    // 1/30 of the bytes it could stream)
    float ra[kLoads], rb[kLoads];
    auto fetch = [&](int64_t rr) {
#pragma unroll
      for (int q = 0; q < kLoads; ++q) {
        const int e = q * kGramThreads + tid;
        int row, kk;
        if (rowfast) { kk = e >> kRowBits; row = e & (TS - 1); } else { row = e >> 4; kk = e & 15; }
        const int64_t rho = rr + kk;
        const bool rok = rho < rend;
        const int64_t off = rok ? gram_off(tk, rho) : 0;
        ra[q] = (rok && (i0 + row) < tk.k) ? gram_ld(tk, (int64_t)(i0 + row) * tk.si + off) : 0.f;
        rb[q] = (!diag && rok && (j0 + row) < tk.k) ? gram_ld(tk, (int64_t)(j0 + row) * tk.si + off) : 0.f;
      }
    };
    fetch(rbeg);
    for (int64_t rr = rbeg; rr < rend; rr += kGramBK) {
#pragma unroll
      for (int q = 0; q < kLoads; ++q) {
        const int e = q * kGramThreads + tid;
        int row, kk;
        if (rowfast) { kk = e >> kRowBits; row = e & (TS - 1); } else { row = e >> 4; kk = e & 15; }
        As[kk][row] = (double)ra[q];
        if (!diag) Bs[kk][row] = (double)rb[q];
      }
      __syncthreads();
      if (rr + kGramBK < rend) fetch(rr + kGramBK);
      const double(*Bp)[TS + 2] = diag ? As : Bs;
#pragma unroll
      for (int kk = 0; kk < kGramBK; ++kk) {
        double a[TT], b[TT];
#pragma unroll
        for (int h = 0; h < TT / 2; ++h) {
          const double2 av = *reinterpret_cast<const double2*>(&As[kk][r0 + 2 * h]);
          const double2 bv = *reinterpret_cast<const double2*>(&Bp[kk][c0 + 2 * h]);
          a[2 * h] = av.x; a[2 * h + 1] = av.y;
          b[2 * h] = bv.x; b[2 * h + 1] = bv.y;
        }
#pragma unroll
        for (int i = 0; i < TT; ++i)
#pragma unroll
          for (int j = 0; j < TT; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }

    double* P = tk.part + (int64_t)split * tk.k * tk.k;
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      const int gi = i0 + r0 + i;
      if (gi >= tk.k) continue;
#pragma unroll
      for (int j = 0; j < TT; ++j) {
        const int gj = j0 + c0 + j;
        if (gj < tk.k) P[(int64_t)gi * tk.k + gj] = acc[i][j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k > 64: 128 x 128 lower-triangular tiles, 8 x 8 DFMA register tiles, double-buffered operands.
// Per reduction index a thread reads 8 + 8 doubles from shared memory for 64 DFMAs (the 4 x 4 tiling
// of the small-k kernel reads 4 + 4 for 16 and is shared-memory bound); the global loads of chunk
// c+1 are in flight while chunk c is being multiplied.
// ---------------------------------------------------------------------------------------------
constexpr int kGramT2 = 128;
constexpr int kGramBK2 = 16;
constexpr int kGramLd2 = kGramT2 + 2;
constexpr size_t kGram2Smem = (size_t)2 * 2 * kGramBK2 * kGramLd2 * sizeof(double);

__global__ void __launch_bounds__(kGramThreads, 1) gram_partial_big(const tta_gram_task* __restrict__ tasks,
                                                                   const __grid_constant__ GramTable tab) {
  extern __shared__ __align__(16) double gsm[];
  // [buffer][operand][kk][row]
  auto tile = [&](int buf, int op) { return gsm + ((size_t)(buf * 2 + op) * kGramBK2) * kGramLd2; };
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;      // 16 x 16 threads, 8 x 8 outputs each (rows ty*8.., cols tx*8..)

  for (int item = blockIdx.x; item < tab.total; item += gridDim.x) {
    int lo = 0, hi = tab.n_tasks;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tab.start[mid] <= item) lo = mid; else hi = mid;
    }
    const tta_gram_task tk = tasks[lo];
    const int T = (tk.k + kGramT2 - 1) / kGramT2;
    const int npairs = T * (T + 1) / 2;
    const int local = item - tab.start[lo];
    const int split = local / npairs;
    int pr = local - split * npairs;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= pr) ++ti;
    const int tj = pr - ti * (ti + 1) / 2;
    const int i0 = ti * kGramT2, j0 = tj * kGramT2;
    const bool diag = (ti == tj);

    const int64_t R = (int64_t)tk.nb * tk.nc;
    const int64_t per = (R + tk.nsplit - 1) / tk.nsplit;
    const int64_t rbeg = (int64_t)split * per;
    const int64_t rend = (rbeg + per) < R ? (rbeg + per) : R;
    const bool rowfast = (tk.si == 1);

    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    float ra[8], rb[8];
    auto fetch = [&](int64_t rr) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = q * kGramThreads + tid;     // 0..2047 = 128 rows x 16 reduction indices
        int row, kk;
        if (rowfast) { kk = e >> 7; row = e & 127; } else { row = e >> 4; kk = e & 15; }
        const int64_t rho = rr + kk;
        const bool rok = rho < rend;
        const int64_t off = rok ? gram_off(tk, rho) : 0;
        ra[q] = (rok && (i0 + row) < tk.k) ? gram_ld(tk, (int64_t)(i0 + row) * tk.si + off) : 0.f;
        rb[q] = (!diag && rok && (j0 + row) < tk.k) ? gram_ld(tk, (int64_t)(j0 + row) * tk.si + off) : 0.f;
      }
    };
    auto stash = [&](int buf) {
      double* As = tile(buf, 0);
      double* Bs = tile(buf, 1);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = q * kGramThreads + tid;
        int row, kk;
        if (rowfast) { kk = e >> 7; row = e & 127; } else { row = e >> 4; kk = e & 15; }
        As[kk * kGramLd2 + row] = (double)ra[q];
        if (!diag) Bs[kk * kGramLd2 + row] = (double)rb[q];
      }
    };

    __syncthreads();          // previous item's readers are done with both buffers
    fetch(rbeg);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int64_t rr = rbeg; rr < rend; rr += kGramBK2) {
      const bool more = rr + kGramBK2 < rend;
      if (more) fetch(rr + kGramBK2);
      const double* As = tile(buf, 0);
      const double* Bs = diag ? As : tile(buf, 1);
#pragma unroll
      for (int kk = 0; kk < kGramBK2; ++kk) {
        double a[8], b[8];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const double2 av = *reinterpret_cast<const double2*>(As + kk * kGramLd2 + ty * 8 + 2 * h);
          const double2 bv = *reinterpret_cast<const double2*>(Bs + kk * kGramLd2 + tx * 8 + 2 * h);
          a[2 * h] = av.x; a[2 * h + 1] = av.y;
          b[2 * h] = bv.x; b[2 * h + 1] = bv.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      if (more) stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }

    double* P = tk.part + (int64_t)split * tk.k * tk.k;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gi = i0 + ty * 8 + i;
      if (gi >= tk.k) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gj = j0 + tx * 8 + j;
        if (gj < tk.k) P[(int64_t)gi * tk.k + gj] = acc[i][j];
      }
    }
  }
}

// G = sum over slices of the partial tiles (fp64, fixed order), mirrored; stored as the eigensolver's column state
// x[col*ld + row] (fp32, zero padded to ld x kpad) and / or as the k x k fp64 matrix g64.
// One or eight lanes per element of the lower triangle (i >= j, consecutive lanes / lane groups take consecutive j: the
// partial rows are read contiguously); eight lanes split the slices between them; both partial kernels store the tiles that
// hold the lower triangle, and the tensor-core tiles on the diagonal are not bitwise symmetric (lo*hi and hi*lo swap
// roles across the diagonal), so the lower triangle is the one that is taken.
__global__ void __launch_bounds__(256) gram_finish(const tta_gram_task* __restrict__ tasks,
                                                   const __grid_constant__ GramFinishFlags flags) {
  const tta_gram_task tk = tasks[blockIdx.y];
  const bool tcp = flags.tc[blockIdx.y] != 0;
  const int64_t kk2 = (int64_t)tk.k * tk.k;
  // lanes per element: 8 when there are many slices to add up (k <= 32 problems with reductions of 73 728: 144 slices),
  // 1 otherwise (consecutive threads then read consecutive elements of a slice: coalesced)
  const int lpe = tk.nsplit >= 24 ? 8 : 1;
  const int lsh = lpe == 8 ? 3 : 0;
  const int sub = threadIdx.x & (lpe - 1);
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int grp = (threadIdx.x & 31) >> lsh;
  const int per_warp = 32 >> lsh;
  for (int64_t e0 = (gtid >> 5) * per_warp; e0 < kk2; e0 += nthr >> lsh) {       // warp-uniform trip count (full-mask shuffles)
    const int64_t e = e0 + grp;
    const int i = (int)(e / tk.k);
    const int j = (int)(e - (int64_t)i * tk.k);
    const bool lower = e < kk2 && j <= i;
    double s = 0.0;
    if (lower) {
      if (tcp) {
        const float* p = reinterpret_cast<const float*>(tk.part) + e;
        for (int sidx = sub; sidx < tk.nsplit; sidx += lpe) s += (double)p[(int64_t)sidx * kk2];
      } else {
        const double* p = tk.part + e;
        for (int sidx = sub; sidx < tk.nsplit; sidx += lpe) s += p[(int64_t)sidx * kk2];
      }
    }
    if (lpe == 8) {
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
    }
    if (lower && sub == 0) {
      if (tk.g64) {
        tk.g64[(int64_t)i * tk.k + j] = s;
        tk.g64[(int64_t)j * tk.k + i] = s;
      }
      if (tk.x) {
        tk.x[(int64_t)j * tk.ld + i] = (float)s;
        tk.x[(int64_t)i * tk.ld + j] = (float)s;
      }
    }
  }
  if (tk.x) {   // zero padding of the column state beyond k
    const int64_t total = (int64_t)tk.ld * tk.kpad;
    for (int64_t e = gtid; e < total; e += nthr) {
      const int col = (int)(e / tk.ld);
      const int row = (int)(e - (int64_t)col * tk.ld);
      if (row >= tk.k || col >= tk.k) tk.x[e] = 0.f;
    }
  }
}

}  // namespace tta

extern "C" int tta_gram_batched(const tta_gram_task* tasks_dev, const tta_gram_task* tasks_host, int n_tasks,
                                void* stream) {
  using namespace tta;
  if (n_tasks < 0 || (n_tasks > 0 && (!tasks_dev || !tasks_host))) {
    set_error("gram: bad task table");
    return TTA_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(gram_partial_big, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kGram2Smem), "gram smem attribute");
    if (rc) return rc;
    attr_set = true;
  }
  for (int first = 0; first < n_tasks; first += kGramMaxTasks) {
    const int cnt = (n_tasks - first) < kGramMaxTasks ? (n_tasks - first) : kGramMaxTasks;
    int64_t max_elems = 0;
    for (int t = 0; t < cnt; ++t) {
      const tta_gram_task& tk = tasks_host[first + t];
      if (tk.k <= 0 || tk.nb <= 0 || tk.nc <= 0 || tk.nsplit <= 0 || tk.ld < tk.k || tk.kpad < tk.k ||
          (tk.ld & 3) || !tk.a || !tk.part || (!tk.x && !tk.g64)) {
        set_error("gram: task %d invalid (k=%d nb=%d nc=%d nsplit=%d ld=%d kpad=%d)", first + t, tk.k, tk.nb,
                  tk.nc, tk.nsplit, tk.ld, tk.kpad);
        return TTA_E_INVALID;
      }
      const int64_t el = (int64_t)tk.k * tk.k * (tk.nsplit >= 24 ? 8 : 1);     // lanes per element (gram_finish)
      max_elems = el > max_elems ? el : max_elems;
    }
    // tensor-core route: TMA-addressable tasks are computed by gram_tc_kernel (fp32 partials)
    GramFinishFlags flags;
    bool any_cc = false;
    {
      tta_gram_task tcs[kGramMaxTasks];
      int ntc = 0;
      const bool on = gram_tc_enabled();
      for (int t = 0; t < cnt; ++t) {
        const tta_gram_task& tk = tasks_host[first + t];
        flags.tc[t] = (on && gram_tc_eligible(tk)) ? 1 : 0;
        if (flags.tc[t]) tcs[ntc++] = tk; else any_cc = true;
      }
      for (int t = cnt; t < kGramMaxTasks; ++t) flags.tc[t] = 0;
      if (ntc) {
        const int rc = gram_tc_run(tcs, ntc, st);
        if (rc) return rc;
      }
    }
    // pass 0: k <= 32 on 32 x 32 tiles; pass 1: k <= 64 on 64 x 64 tiles; pass 2: larger k on 128 x 128
    // tiles.  A task that does not belong to a pass contributes zero tiles to its table.
    for (int pass = 0; any_cc && pass < 3; ++pass) {
      GramTable tab;
      tab.n_tasks = cnt;
      int64_t total = 0;
      const int tsz = pass == 0 ? 32 : pass == 1 ? kGramT : kGramT2;
      for (int t = 0; t < cnt; ++t) {
        const tta_gram_task& tk = tasks_host[first + t];
        tab.start[t] = (int)total;
        if (flags.tc[t] || gram_class(tk) != pass) continue;
        const int T = (tk.k + tsz - 1) / tsz;
        total += (int64_t)(T * (T + 1) / 2) * tk.nsplit;
        if (total > 0x7fffffff) {
          set_error("gram: too many tiles");
          return TTA_E_INVALID;
        }
      }
      tab.start[cnt] = (int)total;
      tab.total = (int)total;
      if (total == 0) continue;
      if (pass < 2) {
        const int grid = total < (int64_t)kNumSMs * 8 ? (int)total : kNumSMs * 8;
        if (pass == 0)
          gram_partial<32, 2><<<grid, kGramThreads, 0, st>>>(tasks_dev + first, tab);
        else
          gram_partial<64, 4><<<grid, kGramThreads, 0, st>>>(tasks_dev + first, tab);
      } else {
        const int grid = total < (int64_t)kNumSMs ? (int)total : kNumSMs;
        gram_partial_big<<<grid, kGramThreads, kGram2Smem, st>>>(tasks_dev + first, tab);
      }
      TTA_CHECK_LAUNCH("gram_partial launch");
    }
    int gx = (int)((max_elems + 255) / 256);
    if (gx > kNumSMs * 16) gx = kNumSMs * 16;
    if (gx < 1) gx = 1;
    gram_finish<<<dim3(gx, cnt), 256, 0, st>>>(tasks_dev + first, flags);
    TTA_CHECK_LAUNCH("gram_finish launch");
  }
  return TTA_OK;
}
