// Unfold (V = W + U fused with the (O,I,KK)->(O,KK,I) permute of admm.py:45,96) and fold
// (admm.py:99) as per-output-channel slab transposes staged through shared memory: both the
// global read and the global write are contiguous runs.
#include "tta_common.cuh"

namespace tta {

constexpr int kFoldThreads = 256;
constexpr int kFoldSmemFloats = 4096;
constexpr int kFoldMaxTasks = 256;

struct FoldTable {
  int n_tasks;
  int total;
  int start[kFoldMaxTasks + 1];
};

__host__ __device__ inline int fold_ich(int KK) {
  int ich = kFoldSmemFloats / (KK > 0 ? KK : 1);
  if (ich >= 32) ich &= ~31;
  return ich < 1 ? 1 : ich;
}

template <bool FOLD>
__global__ void __launch_bounds__(kFoldThreads) fold_kernel(const tta_fold_task* __restrict__ tasks,
                                                           const __grid_constant__ FoldTable tab) {
  __shared__ float s[kFoldSmemFloats];
  const int tid = threadIdx.x;
  for (int item = blockIdx.x; item < tab.total; item += gridDim.x) {
    int lo = 0, hi = tab.n_tasks;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tab.start[mid] <= item) lo = mid; else hi = mid;
    }
    const tta_fold_task tk = tasks[lo];
    const int KK = tk.KK, I = tk.I;
    const int ich = fold_ich(KK);
    const int per_o = (I + ich - 1) / ich;
    const int local = item - tab.start[lo];
    const int o = local / per_o;
    const int i0 = (local - o * per_o) * ich;
    const int ni = (I - i0) < ich ? (I - i0) : ich;
    const int cnt = ni * KK;
    const int64_t nat = ((int64_t)o * I + i0) * KK;   // natural (O,I,KK) offset of this chunk
    const int64_t perm = (int64_t)o * KK * I + i0;    // permuted (O,KK,I) offset: + kk*I + ii
    if (!FOLD) {
      const float* w = tk.w + nat;
      const float* u = tk.u ? tk.u + nat : nullptr;
      for (int e = tid; e < cnt; e += kFoldThreads) s[e] = w[e] + (u ? u[e] : 0.f);
      __syncthreads();
      float* t = tk.t + perm;
      for (int e = tid; e < cnt; e += kFoldThreads) {
        const int kk = e / ni, ii = e - kk * ni;
        t[(int64_t)kk * I + ii] = s[ii * KK + kk];
      }
    } else {
      const float* t = tk.t + perm;
      for (int e = tid; e < cnt; e += kFoldThreads) {
        const int kk = e / ni, ii = e - kk * ni;
        s[ii * KK + kk] = t[(int64_t)kk * I + ii];
      }
      __syncthreads();
      float* z = tk.z + nat;
      for (int e = tid; e < cnt; e += kFoldThreads) z[e] = s[e];
    }
    __syncthreads();
  }
}

template <bool FOLD>
static int launch_fold(const tta_fold_task* tasks_dev, const tta_fold_task* tasks_host, int n_tasks,
                       void* stream) {
  if (n_tasks < 0 || (n_tasks > 0 && (!tasks_dev || !tasks_host))) {
    set_error("fold: bad task table");
    return TTA_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int first = 0; first < n_tasks; first += kFoldMaxTasks) {
    const int cnt = (n_tasks - first) < kFoldMaxTasks ? (n_tasks - first) : kFoldMaxTasks;
    FoldTable tab;
    tab.n_tasks = cnt;
    int64_t total = 0;
    for (int t = 0; t < cnt; ++t) {
      const tta_fold_task& tk = tasks_host[first + t];
      if (tk.O <= 0 || tk.I <= 0 || tk.KK <= 0 || tk.KK > kFoldSmemFloats || !tk.t ||
          (FOLD ? !tk.z : !tk.w)) {
        set_error("fold: task %d invalid (O=%d I=%d KK=%d)", first + t, tk.O, tk.I, tk.KK);
        return TTA_E_INVALID;
      }
      tab.start[t] = (int)total;
      const int ich = fold_ich(tk.KK);
      total += (int64_t)tk.O * ((tk.I + ich - 1) / ich);
      if (total > 0x7fffffff) {
        set_error("fold: too many chunks");
        return TTA_E_INVALID;
      }
    }
    tab.start[cnt] = (int)total;
    tab.total = (int)total;
    if (total == 0) continue;
    int grid = total < kNumSMs * 8 ? (int)total : kNumSMs * 8;
    fold_kernel<FOLD><<<grid, kFoldThreads, 0, st>>>(tasks_dev + first, tab);
    TTA_CHECK_LAUNCH("fold launch");
  }
  return TTA_OK;
}

}  // namespace tta

extern "C" {
int tta_unfold_add_batched(const tta_fold_task* tasks_dev, const tta_fold_task* tasks_host, int n_tasks,
                           void* stream) {
  return tta::launch_fold<false>(tasks_dev, tasks_host, n_tasks, stream);
}
int tta_fold_store_batched(const tta_fold_task* tasks_dev, const tta_fold_task* tasks_host, int n_tasks,
                           void* stream) {
  return tta::launch_fold<true>(tasks_dev, tasks_host, n_tasks, stream);
}
}
