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

__device__ __forceinline__ int64_t gram_off(const tta_gram_task& tk, int64_t rho) {
  if (tk.nb == 1) return rho * tk.sc;
  const int64_t b = rho / tk.nc;
  return b * tk.sb + (rho - b * tk.nc) * tk.sc;
}

__global__ void __launch_bounds__(kGramThreads) gram_partial(const tta_gram_task* __restrict__ tasks,
                                                            const __grid_constant__ GramTable tab) {
  __shared__ __align__(16) double As[kGramBK][kGramT + 2];
  __shared__ __align__(16) double Bs[kGramBK][kGramT + 2];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int r0 = (warp & 1) * 32 + (lane >> 2) * 4;  // rows of this thread inside the tile
  const int c0 = (warp >> 1) * 16 + (lane & 3) * 4;  // cols

  for (int item = blockIdx.x; item < tab.total; item += gridDim.x) {
    int lo = 0, hi = tab.n_tasks;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tab.start[mid] <= item) lo = mid; else hi = mid;
    }
    const tta_gram_task tk = tasks[lo];
    const int T = (tk.k + kGramT - 1) / kGramT;
    const int npairs = T * (T + 1) / 2;
    const int local = item - tab.start[lo];
    const int split = local / npairs;
    int pr = local - split * npairs;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= pr) ++ti;  // ti >= tj, pair index = ti(ti+1)/2 + tj
    const int tj = pr - ti * (ti + 1) / 2;
    const int i0 = ti * kGramT, j0 = tj * kGramT;
    const bool diag = (ti == tj);

    const int64_t R = (int64_t)tk.nb * tk.nc;
    const int64_t per = (R + tk.nsplit - 1) / tk.nsplit;
    const int64_t rbeg = (int64_t)split * per;
    const int64_t rend = (rbeg + per) < R ? (rbeg + per) : R;
    const bool rowfast = (tk.si == 1);

    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int64_t rr = rbeg; rr < rend; rr += kGramBK) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = q * kGramThreads + tid;
        int row, kk;
        if (rowfast) { kk = e >> 6; row = e & 63; } else { row = e >> 4; kk = e & 15; }
        const int64_t rho = rr + kk;
        const bool rok = rho < rend;
        const int64_t off = rok ? gram_off(tk, rho) : 0;
        float va = 0.f, vb = 0.f;
        if (rok && (i0 + row) < tk.k) va = __ldg(tk.a + (int64_t)(i0 + row) * tk.si + off);
        As[kk][row] = (double)va;
        if (!diag) {
          if (rok && (j0 + row) < tk.k) vb = __ldg(tk.a + (int64_t)(j0 + row) * tk.si + off);
          Bs[kk][row] = (double)vb;
        }
      }
      __syncthreads();
      const double(*Bp)[kGramT + 2] = diag ? As : Bs;
#pragma unroll
      for (int kk = 0; kk < kGramBK; ++kk) {
        const double2 a01 = *reinterpret_cast<const double2*>(&As[kk][r0]);
        const double2 a23 = *reinterpret_cast<const double2*>(&As[kk][r0 + 2]);
        const double2 b01 = *reinterpret_cast<const double2*>(&Bp[kk][c0]);
        const double2 b23 = *reinterpret_cast<const double2*>(&Bp[kk][c0 + 2]);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }

    double* P = tk.part + (int64_t)split * tk.k * tk.k;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = i0 + r0 + i;
      if (gi >= tk.k) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = j0 + c0 + j;
        if (gj < tk.k) P[(int64_t)gi * tk.k + gj] = acc[i][j];
      }
    }
  }
}

// x[col*ld + row] = sum_s part[s][max-tile-ordered (row,col)]; zero padding beyond k.
__global__ void __launch_bounds__(256) gram_finish(const tta_gram_task* __restrict__ tasks) {
  const tta_gram_task tk = tasks[blockIdx.y];
  const int64_t total = (int64_t)tk.ld * tk.kpad;
  const int64_t kk2 = (int64_t)tk.k * tk.k;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(e / tk.ld);
    const int row = (int)(e - (int64_t)col * tk.ld);
    float v = 0.f;
    if (row < tk.k && col < tk.k) {
      int i = row, j = col;
      if ((i / kGramT) < (j / kGramT)) { i = col; j = row; }  // stored tiles have ti >= tj
      const double* p = tk.part + (int64_t)i * tk.k + j;
      double s = 0.0;
      for (int sidx = 0; sidx < tk.nsplit; ++sidx) s += p[(int64_t)sidx * kk2];
      v = (float)s;
      if (tk.g64) tk.g64[(int64_t)row * tk.k + col] = s;
    }
    tk.x[e] = v;
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
  for (int first = 0; first < n_tasks; first += kGramMaxTasks) {
    const int cnt = (n_tasks - first) < kGramMaxTasks ? (n_tasks - first) : kGramMaxTasks;
    GramTable tab;
    tab.n_tasks = cnt;
    int64_t total = 0, max_elems = 0;
    for (int t = 0; t < cnt; ++t) {
      const tta_gram_task& tk = tasks_host[first + t];
      if (tk.k <= 0 || tk.nb <= 0 || tk.nc <= 0 || tk.nsplit <= 0 || tk.ld < tk.k || tk.kpad < tk.k ||
          (tk.ld & 3) || !tk.a || !tk.part || !tk.x) {
        set_error("gram: task %d invalid (k=%d nb=%d nc=%d nsplit=%d ld=%d kpad=%d)", first + t, tk.k, tk.nb,
                  tk.nc, tk.nsplit, tk.ld, tk.kpad);
        return TTA_E_INVALID;
      }
      const int T = (tk.k + kGramT - 1) / kGramT;
      tab.start[t] = (int)total;
      total += (int64_t)(T * (T + 1) / 2) * tk.nsplit;
      const int64_t el = (int64_t)tk.ld * tk.kpad;
      max_elems = el > max_elems ? el : max_elems;
      if (total > 0x7fffffff) {
        set_error("gram: too many tiles");
        return TTA_E_INVALID;
      }
    }
    tab.start[cnt] = (int)total;
    tab.total = (int)total;
    const int grid = total < (int64_t)kNumSMs * 8 ? (int)total : kNumSMs * 8;
    gram_partial<<<grid, kGramThreads, 0, st>>>(tasks_dev + first, tab);
    TTA_CHECK_LAUNCH("gram_partial launch");
    int gx = (int)((max_elems + 255) / 256);
    if (gx > kNumSMs * 4) gx = kNumSMs * 4;
    if (gx < 1) gx = 1;
    gram_finish<<<dim3(gx, cnt), 256, 0, st>>>(tasks_dev + first);
    TTA_CHECK_LAUNCH("gram_finish launch");
  }
  return TTA_OK;
}
