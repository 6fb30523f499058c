// Batched fp32 GEMM with general operand strides: C = (A * B) .* colscale.
// Used for the projection side of every truncated SVD (carry = U_r^T A, core = A V_r / sigma:
// ttd.py:21-26), the tt2ten reconstruction chain (ttd.py:39-40) and the Tucker mode products
// (tensorly multi_mode_dot behind admm.py:116-117).  One launch covers all layers of a step: the
// tile list of every task is concatenated and grid-strided.
#include "tta_common.cuh"

namespace tta {

constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16;
constexpr int kGemmThreads = 256;
constexpr int kGemmMaxTasks = 192;

struct GemmTable {
  int n_tasks;
  int total;
  int start[kGemmMaxTasks + 1];
};

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
};
template <>
struct Vec4<double> {
  using type = double4;
};

// T = float: the fp32 path.  T = double: same tiling with DFMA, used by the eigenvector refinement
// (all pointers of the task then address doubles).
template <typename T>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(const tta_gemm_task* __restrict__ tasks,
                                                           const __grid_constant__ GemmTable tab) {
  __shared__ __align__(32) T As[kGemmBK][kGemmBM + 4];
  __shared__ __align__(32) T Bs[kGemmBK][kGemmBN + 4];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;

  for (int item = blockIdx.x; item < tab.total; item += gridDim.x) {
    int lo = 0, hi = tab.n_tasks;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tab.start[mid] <= item) lo = mid; else hi = mid;
    }
    const tta_gemm_task tk = tasks[lo];
    const int tiles_n = (tk.N + kGemmBN - 1) / kGemmBN;
    const int local = item - tab.start[lo];
    const int m0 = (local / tiles_n) * kGemmBM;
    const int n0 = (local % tiles_n) * kGemmBN;
    const bool a_kfast = (tk.sak == 1);
    const bool b_jfast = (tk.sbj == 1);

    const T* __restrict__ pa = reinterpret_cast<const T*>(tk.a);
    const T* __restrict__ pb = reinterpret_cast<const T*>(tk.b);
    T* __restrict__ pc = reinterpret_cast<T*>(tk.c);
    const bool guarded = sizeof(T) == 8 && (tk.flags & TTA_GEMM_GUARD);
    if (guarded && *reinterpret_cast<const double*>(tk.colscale) == 0.0) continue;   // CTA-uniform
    const T* __restrict__ pcs = guarded ? nullptr : reinterpret_cast<const T*>(tk.colscale);
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

    // The global loads of k-tile t + 1 are issued into registers before the FMAs of tile t, so their latency
    // overlaps the arithmetic (with few tiles per launch -- the refinement / projection GEMMs of a three-layer
    // group -- the kernel is otherwise a chain of load -> barrier -> compute -> barrier round trips).
    T ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = q * kGemmThreads + tid;  // 0..1023
        int ar, ak;
        if (a_kfast) { ar = e >> 4; ak = e & 15; } else { ak = e >> 6; ar = e & 63; }
        const int gi = m0 + ar, gk = k0 + ak;
        ra[q] = (gi < tk.M && gk < tk.K) ? __ldg(pa + (int64_t)gi * tk.sai + (int64_t)gk * tk.sak) : T(0);
        int bk, bj;
        if (b_jfast) { bk = e >> 6; bj = e & 63; } else { bj = e >> 4; bk = e & 15; }
        const int gj = n0 + bj, gkb = k0 + bk;
        rb[q] = (gj < tk.N && gkb < tk.K) ? __ldg(pb + (int64_t)gkb * tk.sbk + (int64_t)gj * tk.sbj) : T(0);
      }
    };
    auto stash = [&]() {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = q * kGemmThreads + tid;
        int ar, ak;
        if (a_kfast) { ar = e >> 4; ak = e & 15; } else { ak = e >> 6; ar = e & 63; }
        As[ak][ar] = ra[q];
        int bk, bj;
        if (b_jfast) { bk = e >> 6; bj = e & 63; } else { bj = e >> 4; bk = e & 15; }
        Bs[bk][bj] = rb[q];
      }
    };
    fetch(0);
    for (int k0 = 0; k0 < tk.K; k0 += kGemmBK) {
      stash();
      __syncthreads();
      if (k0 + kGemmBK < tk.K) fetch(k0 + kGemmBK);
#pragma unroll
      for (int kk = 0; kk < kGemmBK; ++kk) {
        using V4 = typename Vec4<T>::type;
        const V4 av = *reinterpret_cast<const V4*>(&As[kk][ty * 4]);
        const V4 bv = *reinterpret_cast<const V4*>(&Bs[kk][tx * 4]);
        const T a[4] = {av.x, av.y, av.z, av.w};
        const T b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = m0 + ty * 4 + i;
      if (gi >= tk.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = n0 + tx * 4 + j;
        if (gj >= tk.N) continue;
        T v = acc[i][j];
        if (pcs) v *= __ldg(pcs + gj);
        if (sizeof(T) == 8 && (tk.flags & TTA_GEMM_STORE_F32))
          reinterpret_cast<float*>(tk.c)[(int64_t)gi * tk.ldc + gj] = (float)v;
        else
          pc[(int64_t)gi * tk.ldc + gj] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) sqnorm_kernel(const tta_sqnorm_task* __restrict__ tasks,
                                                    double* __restrict__ out) {
  __shared__ double s_part[8];
  const tta_sqnorm_task tk = tasks[blockIdx.y];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tk.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)tk.x[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < 8 ? s_part[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, t);
  }
}

}  // namespace tta

namespace tta {
// tensor-core path for fp32 tasks (gemm_tf32.cu)
bool gemm_tf32x3_eligible(const tta_gemm_task& tk);
int gemm_tf32x3_run(const tta_gemm_task* tasks_dev, const tta_gemm_task* tasks_host, int cnt, cudaStream_t st);
void gemm_tf32x3_set_mode(int mode);
int gemm_tf32x3_mode();

template <typename T>
static int launch_gemm(const tta_gemm_task* tasks_dev, const tta_gemm_task* tasks_host, int n_tasks, void* stream) {
  if (n_tasks < 0 || (n_tasks > 0 && (!tasks_dev || !tasks_host))) {
    set_error("gemm: bad task table");
    return TTA_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int first = 0; first < n_tasks; first += kGemmMaxTasks) {
    const int cnt = (n_tasks - first) < kGemmMaxTasks ? (n_tasks - first) : kGemmMaxTasks;
    GemmTable tab;
    tab.n_tasks = cnt;
    int64_t total = 0;
    for (int t = 0; t < cnt; ++t) {
      const tta_gemm_task& tk = tasks_host[first + t];
      if (tk.M < 0 || tk.N < 0 || tk.K < 0 || (tk.M > 0 && tk.N > 0 && (!tk.c || (tk.K > 0 && (!tk.a || !tk.b))))) {
        set_error("gemm: task %d invalid (M=%d N=%d K=%d)", first + t, tk.M, tk.N, tk.K);
        return TTA_E_INVALID;
      }
      tab.start[t] = (int)total;
      if (sizeof(T) == 4 && gemm_tf32x3_eligible(tk)) continue;   // served by the tcgen05 kernel below
      total += (int64_t)((tk.M + kGemmBM - 1) / kGemmBM) * ((tk.N + kGemmBN - 1) / kGemmBN);
      if (total > 0x7fffffff) {
        set_error("gemm: too many tiles");
        return TTA_E_INVALID;
      }
    }
    tab.start[cnt] = (int)total;
    tab.total = (int)total;
    if (sizeof(T) == 4 && gemm_tf32x3_mode()) {
      const int rc = gemm_tf32x3_run(tasks_dev + first, tasks_host + first, cnt, st);
      if (rc) return rc;
    }
    if (total == 0) continue;
    const int grid = total < (int64_t)kNumSMs * 16 ? (int)total : kNumSMs * 16;
    gemm_kernel<T><<<grid, kGemmThreads, 0, st>>>(tasks_dev + first, tab);
    TTA_CHECK_LAUNCH("gemm launch");
  }
  return TTA_OK;
}
}  // namespace tta

extern "C" {

void tta_gemm_enable_tc(int on) { tta::gemm_tf32x3_set_mode(on < 0 ? 0 : on > 2 ? 2 : on); }

int tta_gemm_batched(const tta_gemm_task* tasks_dev, const tta_gemm_task* tasks_host, int n_tasks,
                     void* stream) {
  return tta::launch_gemm<float>(tasks_dev, tasks_host, n_tasks, stream);
}

int tta_gemm_f64_batched(const tta_gemm_task* tasks_dev, const tta_gemm_task* tasks_host, int n_tasks,
                         void* stream) {
  return tta::launch_gemm<double>(tasks_dev, tasks_host, n_tasks, stream);
}

int tta_sqnorm_batched(const tta_sqnorm_task* tasks_dev, const tta_sqnorm_task* tasks_host, int n_tasks,
                       double* out_dev, void* stream) {
  using namespace tta;
  if (n_tasks <= 0) return TTA_OK;
  if (!tasks_dev || !tasks_host || !out_dev) {
    set_error("sqnorm: null argument");
    return TTA_E_INVALID;
  }
  int64_t mx = 0;
  for (int t = 0; t < n_tasks; ++t) mx = tasks_host[t].n > mx ? tasks_host[t].n : mx;
  int gx = (int)((mx + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  if (gx > kNumSMs * 4) gx = kNumSMs * 4;
  dim3 grid(gx, n_tasks);
  sqnorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tasks_dev, out_dev);
  TTA_CHECK_LAUNCH("sqnorm launch");
  return TTA_OK;
}
}
