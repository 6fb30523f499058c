// fp64 refinement of the fp32 Jacobi eigenvectors + dominant-r selection (see include/tta.h).
//
// With Q the (nearly orthonormal) fp32 eigenvector estimate, S = Q^T G Q, T = Q^T Q, R = I - T:
//   lambda_j = S_jj / T_jj
//   E_ij     = (S_ij + lambda_j R_ij) / (lambda_j - lambda_i)      i != j, gap resolved
//   E_ij     = R_ij / 2                                            unresolved cluster (span-neutral)
//   E_jj     = R_jj / 2
//   q_j'     = q_j + sum_i q_i E_ij                                (Ogita & Aishima 2018, one step)
// Only the r dominant columns j are formed (the truncation of ttd.py:21-23 / admm.py:132-134), so only
// the rows j of S and T that can be selected are needed: refine_prepare sorts the vectors by their
// fp32 eigenvalue estimate ||x_j|| (descending) and the caller's GEMMs form the first `wnd` rows of S
// and T only (wnd = r + a safety margin far wider than the fp32 ordering error).  Eigenvalues of
// vectors behind the window enter the gap denominators as their fp32 estimates.
#include "tta_common.cuh"

namespace tta {

constexpr int kRefThreads = 512;
constexpr double kRefCutRel = 4e-7;     // eigenvalues below cut_rel * lambda_max count as zero
constexpr double kRefGapRel = 1e-6;     // |lambda_i - lambda_j| below gap_rel * lambda_max: cluster
constexpr double kRefMaxCorr = 0.05;    // first-order correction must stay small

constexpr int kRefSplit = 8;            // CTAs per problem in the prepare / coeff kernels

// c[0..k) (scratch, overwritten later by refine_coeff) = squared column norms of X
__global__ void __launch_bounds__(kRefThreads) refine_norms_kernel(const tta_refine_task* __restrict__ tasks) {
  const tta_refine_task tk = tasks[blockIdx.x];
  const int k = tk.k;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarp = kRefThreads / 32;
  for (int j = blockIdx.y * nwarp + warp; j < k; j += nwarp * kRefSplit) {
    const float* x = tk.x + (int64_t)j * tk.ld;
    double a = 0.0;
    for (int e = lane; e < k; e += 32) a = fma((double)x[e], (double)x[e], a);
    a = warp_sum(a);
    if (lane == 0) tk.c[j] = a;
  }
}

// qt row p = the unit vector of the p-th largest column of X; lam0[p] = its norm (0 for null columns)
__global__ void __launch_bounds__(kRefThreads) refine_prepare_kernel(const tta_refine_task* __restrict__ tasks) {
  extern __shared__ double s_nrm[];   // k squared norms
  const tta_refine_task tk = tasks[blockIdx.x];
  const int k = tk.k;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarp = kRefThreads / 32;
  for (int j = tid; j < k; j += kRefThreads) s_nrm[j] = tk.c[j];
  __syncthreads();
  double mx = 0.0;
  for (int j = 0; j < k; ++j) mx = fmax(mx, s_nrm[j]);
  const double cut = mx * (kRefCutRel * kRefCutRel);
  for (int j = blockIdx.y * nwarp + warp; j < k; j += nwarp * kRefSplit) {
    const double aj = s_nrm[j];
    int pos = 0;
    for (int i = lane; i < k; i += 32) {
      const double ai = s_nrm[i];
      pos += (ai > aj) || (ai == aj && i < j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
    const float* x = tk.x + (int64_t)j * tk.ld;
    const bool live = aj > cut && aj > 0.0;
    const double inv = live ? 1.0 / sqrt(aj) : 0.0;
    double* q = tk.qt + (int64_t)pos * k;
    for (int e = lane; e < k; e += 32) q[e] = (double)x[e] * inv;
    if (lane == 0) tk.lam0[pos] = live ? sqrt(aj) : 0.0;
  }
}

// s, t: the first wnd rows of S = Qt G Qt^T and T = Qt Qt^T (wnd x k, row-major)
__global__ void __launch_bounds__(kRefThreads) refine_coeff_kernel(const tta_refine_task* __restrict__ tasks) {
  extern __shared__ double s_lam[];  // k eigenvalues, then r ints (selection)
  const tta_refine_task tk = tasks[blockIdx.x];
  const int k = tk.k, wnd = tk.wnd;
  int* s_sel = reinterpret_cast<int*>(s_lam + k);   // s_sel[p] = row holding rank p
  const int tid = threadIdx.x;
  for (int j = tid; j < k; j += kRefThreads) {
    if (j < wnd) {
      const double tjj = tk.t[(int64_t)j * k + j];
      s_lam[j] = tjj > 0.5 ? tk.s[(int64_t)j * k + j] / tjj : 0.0;
    } else {
      s_lam[j] = tk.lam0[j];
    }
  }
  __syncthreads();
  double lmax = 0.0;
  for (int j = 0; j < k; ++j) lmax = fmax(lmax, s_lam[j]);
  // descending rank of the window rows (rows behind the window are smaller by construction)
  for (int j = tid; j < wnd; j += kRefThreads) {
    const double lj = s_lam[j];
    int pos = 0;
    for (int i = 0; i < wnd; ++i) {
      const double li = s_lam[i];
      pos += (li > lj) || (li == lj && i < j);
    }
    if (pos < tk.r) s_sel[pos] = j;
  }
  __syncthreads();
  const double gap_min = kRefGapRel * lmax;
  const int total = tk.r * k;
  for (int idx = blockIdx.y * kRefThreads + tid; idx < total; idx += kRefThreads * kRefSplit) {
    const int p = idx / k, i = idx - p * k;
    const int j = s_sel[p];
    const double lj = s_lam[j], li = s_lam[i];
    const double tij = tk.t[(int64_t)j * k + i];
    const bool live_i = (i < wnd) ? (tk.t[(int64_t)i * k + i] > 0.5) : (li > 0.0);
    double c;
    if (i == j) {
      c = 1.0 + 0.5 * (1.0 - tij);
    } else if (!live_i) {
      c = 0.0;                                   // null direction: nothing to mix in
    } else {
      const double rij = -tij;
      const double gap = lj - li;
      c = 0.5 * rij;
      if (fabs(gap) > gap_min) {
        const double e = (tk.s[(int64_t)j * k + i] + lj * rij) / gap;
        if (fabs(e) <= kRefMaxCorr) c = e;
      }
    }
    tk.c[idx] = c;
  }
  if (blockIdx.y == 0)
    for (int p = tid; p < tk.r; p += kRefThreads) tk.lam[p] = s_lam[s_sel[p]];
}

// One warp per selected row: renormalise the corrected vector in fp64 (the first-order update leaves
// ||q'||^2 = 1 + O(theta^2)), then emit the fp32 outputs.
__global__ void __launch_bounds__(256) refine_finalize_kernel(const tta_refine_task* __restrict__ tasks) {
  const tta_refine_task tk = tasks[blockIdx.y];
  const int k = tk.k, r = tk.r;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double lmax = tk.lam[0];
  for (int p = blockIdx.x * 8 + warp; p < r; p += gridDim.x * 8) {
    const double lam = tk.lam[p];
    const bool live = lam > kRefCutRel * lmax && lam > 0.0;
    const double* src = tk.e64 + (int64_t)p * k;
    double nrm = 0.0;
    for (int e = lane; e < k; e += 32) nrm = fma(src[e], src[e], nrm);
    nrm = warp_sum(nrm);
    const double inv = (live && nrm > 0.0) ? 1.0 / sqrt(nrm) : 0.0;
    const float sg = live ? (float)sqrt(lam) : 0.f;
    for (int e = lane; e < k; e += 32) {
      const float v = (float)(src[e] * inv);
      tk.e[(int64_t)p * k + e] = v;
      if (tk.et) tk.et[(int64_t)e * r + p] = v;
      if (tk.se) tk.se[(int64_t)p * k + e] = v * sg;
    }
    if (lane == 0) {
      if (tk.sigma) tk.sigma[p] = sg;
      if (tk.isigma) tk.isigma[p] = live ? (float)(1.0 / sqrt(lam)) : 0.f;
    }
  }
}

static int validate(const tta_refine_task* th, int n, const char* what, bool finalize_only = false) {
  if (n < 0 || (n > 0 && !th)) {
    set_error("%s: bad task table", what);
    return TTA_E_INVALID;
  }
  for (int t = 0; t < n; ++t) {
    const tta_refine_task& tk = th[t];
    if (finalize_only) {   // also the output stage of tta_symeig_top_batched: only e64 / lam / outputs are read
      if (tk.k <= 0 || tk.r <= 0 || tk.r > tk.k || !tk.lam || !tk.e64 || !tk.e) {
        set_error("%s: task %d invalid (k=%d r=%d)", what, t, tk.k, tk.r);
        return TTA_E_INVALID;
      }
      continue;
    }
    if (tk.k <= 0 || tk.r <= 0 || tk.r > tk.k || tk.wnd < tk.r || tk.wnd > tk.k || tk.ld < tk.k || !tk.x || !tk.qt ||
        !tk.s || !tk.t || !tk.c || !tk.lam || !tk.lam0 || !tk.e64 || !tk.e) {
      set_error("%s: task %d invalid (k=%d r=%d wnd=%d ld=%d)", what, t, tk.k, tk.r, tk.wnd, tk.ld);
      return TTA_E_INVALID;
    }
  }
  return TTA_OK;
}

}  // namespace tta

extern "C" {

int tta_refine_prepare_batched(const tta_refine_task* tasks_dev, const tta_refine_task* tasks_host, int n_tasks,
                               void* stream) {
  using namespace tta;
  int rc = validate(tasks_host, n_tasks, "refine_prepare");
  if (rc || n_tasks == 0) return rc;
  size_t smem = 0;
  for (int t = 0; t < n_tasks; ++t) {
    const size_t need = (size_t)tasks_host[t].k * 8;
    smem = need > smem ? need : smem;
  }
  if (smem > 48 * 1024) {
    rc = check_cuda(cudaFuncSetAttribute(refine_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "refine_prepare smem attribute");
    if (rc) return rc;
  }
  refine_norms_kernel<<<dim3(n_tasks, kRefSplit), kRefThreads, 0, (cudaStream_t)stream>>>(tasks_dev);
  TTA_CHECK_LAUNCH("refine_norms launch");
  refine_prepare_kernel<<<dim3(n_tasks, kRefSplit), kRefThreads, smem, (cudaStream_t)stream>>>(tasks_dev);
  TTA_CHECK_LAUNCH("refine_prepare launch");
  return TTA_OK;
}

int tta_refine_coeff_batched(const tta_refine_task* tasks_dev, const tta_refine_task* tasks_host, int n_tasks,
                             void* stream) {
  using namespace tta;
  int rc = validate(tasks_host, n_tasks, "refine_coeff");
  if (rc || n_tasks == 0) return rc;
  size_t smem = 0;
  for (int t = 0; t < n_tasks; ++t) {
    const size_t need = (size_t)tasks_host[t].k * 8 + (size_t)tasks_host[t].r * 4;
    smem = need > smem ? need : smem;
  }
  if (smem > 48 * 1024) {
    rc = check_cuda(cudaFuncSetAttribute(refine_coeff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "refine_coeff smem attribute");
    if (rc) return rc;
  }
  refine_coeff_kernel<<<dim3(n_tasks, kRefSplit), kRefThreads, smem, (cudaStream_t)stream>>>(tasks_dev);
  TTA_CHECK_LAUNCH("refine_coeff launch");
  return TTA_OK;
}

int tta_refine_finalize_batched(const tta_refine_task* tasks_dev, const tta_refine_task* tasks_host, int n_tasks,
                                void* stream) {
  using namespace tta;
  int rc = validate(tasks_host, n_tasks, "refine_finalize", true);
  if (rc || n_tasks == 0) return rc;
  int mx = 0;
  for (int t = 0; t < n_tasks; ++t) mx = tasks_host[t].r > mx ? tasks_host[t].r : mx;
  int gx = (mx + 7) / 8;
  if (gx > kNumSMs) gx = kNumSMs;
  if (gx < 1) gx = 1;
  refine_finalize_kernel<<<dim3(gx, n_tasks), 256, 0, (cudaStream_t)stream>>>(tasks_dev);
  TTA_CHECK_LAUNCH("refine_finalize launch");
  return TTA_OK;
}
}
