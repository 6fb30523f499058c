// Batched symmetric eigensolver: one-sided (Hestenes) block Jacobi on the columns of X = G, and the
// selection of the dominant r eigenpairs.  Replaces LAPACK gesdd behind numpy.linalg.svd
// (ttd.py:17, admm.py:131,143) and tensorly's partial_svd (admm.py:116,124).
//
// Why Jacobi: random-init weights have a flat (Marchenko-Pastur) spectrum, so subspace / randomised
// iterations do not reach the 1e-4 parity bound; a full-accuracy decomposition of the k x k Gram
// matrix (k up to 512 for ResNet-50, 2048 for the Tucker sweep) is needed.  Rotating the columns of
// X = G until they are mutually orthogonal gives X = V diag(lambda): no separate eigenvector
// accumulation.
//
// Parallel structure.  The kpad columns are cut into blocks of `bw` columns.  One sweep is
//   launch 0        : all pairs inside each block       (CTA = two blocks, round-robin inside each)
//   launch 1..nb-1  : all pairs between block A and B   (CTA = one block pair; block pairs of one
//                     launch form one round of a round-robin tournament, so they are disjoint)
// Every CTA stages its 2*bw columns in shared memory, one warp owns one column pair per step
// (float4 column slices, warp-shuffle reductions for the three inner products) and writes the block
// back.  All problems of a network step share the same launches.  Convergence (no rotation in a
// sweep) is tested on the host once per sweep.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tta_common.cuh"

namespace tta {

struct JacItem {
  int32_t prob;
  int32_t blk_a;
  int32_t blk_b;  // -1: single block (intra launch, odd block count)
  int32_t kind;   // 0 intra, 1 cross
};

constexpr int kJacMaxBw = 16;
constexpr int kJacThreads = kJacMaxBw * 32;
constexpr float kJacFloorRel = 1e-7f;  // columns below floor_rel * max column norm are numerically zero

// round `r` of the circle-method tournament on n (even) players; pair q in [0, n/2)
__host__ __device__ inline void rr_pair(int n, int r, int q, int& p0, int& p1) {
  if (q == 0) {
    p0 = n - 1;
    p1 = r;
  } else {
    p0 = (r + q) % (n - 1);
    p1 = (r - q + (n - 1)) % (n - 1);
  }
}

// floor2[p] = (floor_rel * max_j ||x_j||)^2
__global__ void __launch_bounds__(256) jacobi_floor_kernel(const tta_eig_task* __restrict__ tasks,
                                                          float* __restrict__ floor2) {
  __shared__ float s_max[8];
  const tta_eig_task tk = tasks[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = 0.f;
  for (int j = warp; j < tk.k; j += 8) {
    const float* x = tk.x + (int64_t)j * tk.ld;
    float a = 0.f;
    for (int e = lane; e < tk.k; e += 32) a = fmaf(x[e], x[e], a);
    a = warp_sum(a);
    mx = fmaxf(mx, a);
  }
  if (lane == 0) s_max[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, s_max[i]);
    floor2[blockIdx.x] = m * (kJacFloorRel * kJacFloorRel);
  }
}

__device__ __forceinline__ int jacobi_rotate_pair(float* __restrict__ x, float* __restrict__ y, int ld, int lane,
                                                  float tol, float floor2) {
  float a = 0.f, b = 0.f, c = 0.f;
  for (int e = lane * 4; e < ld; e += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(x + e);
    const float4 yv = *reinterpret_cast<const float4*>(y + e);
    a = fmaf(xv.x, xv.x, a); a = fmaf(xv.y, xv.y, a); a = fmaf(xv.z, xv.z, a); a = fmaf(xv.w, xv.w, a);
    b = fmaf(yv.x, yv.x, b); b = fmaf(yv.y, yv.y, b); b = fmaf(yv.z, yv.z, b); b = fmaf(yv.w, yv.w, b);
    c = fmaf(xv.x, yv.x, c); c = fmaf(xv.y, yv.y, c); c = fmaf(xv.z, yv.z, c); c = fmaf(xv.w, yv.w, c);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  if (!(a > floor2) || !(b > floor2)) return 0;
  if (!(fabsf(c) > tol * (sqrtf(a) * sqrtf(b)))) return 0;
  const float zeta = (b - a) / (2.f * c);
  const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
  const float cs = 1.f / sqrtf(1.f + t * t);
  const float sn = cs * t;
  // x' = x - sn*(y + tau*x), y' = y + sn*(x - tau*y) with tau = sn/(1+cs)  (== cs*x - sn*y, sn*x + cs*y).
  // Late rotations have cs == 1.0f after rounding; applying the 1-cs part explicitly keeps every
  // rotation norm-preserving to rounding error instead of inflating the columns by t^2/2 each time.
  const float tau = sn / (1.f + cs);
  for (int e = lane * 4; e < ld; e += 128) {
    float4 xv = *reinterpret_cast<const float4*>(x + e);
    float4 yv = *reinterpret_cast<const float4*>(y + e);
    float4 xn, yn;
    xn.x = fmaf(-sn, fmaf(tau, xv.x, yv.x), xv.x); yn.x = fmaf(sn, fmaf(-tau, yv.x, xv.x), yv.x);
    xn.y = fmaf(-sn, fmaf(tau, xv.y, yv.y), xv.y); yn.y = fmaf(sn, fmaf(-tau, yv.y, xv.y), yv.y);
    xn.z = fmaf(-sn, fmaf(tau, xv.z, yv.z), xv.z); yn.z = fmaf(sn, fmaf(-tau, yv.z, xv.z), yv.z);
    xn.w = fmaf(-sn, fmaf(tau, xv.w, yv.w), xv.w); yn.w = fmaf(sn, fmaf(-tau, yv.w, xv.w), yv.w);
    *reinterpret_cast<float4*>(x + e) = xn;
    *reinterpret_cast<float4*>(y + e) = yn;
  }
  return 1;
}

__global__ void __launch_bounds__(kJacThreads) jacobi_step_kernel(const JacItem* __restrict__ items,
                                                                  const tta_eig_task* __restrict__ tasks,
                                                                  int32_t* __restrict__ counts,
                                                                  const int32_t* __restrict__ done,
                                                                  const float* __restrict__ floor2, float tol) {
  extern __shared__ __align__(16) float cols[];
  __shared__ int s_rot;
  const JacItem it = items[blockIdx.x];
  if (done[it.prob]) return;
  const tta_eig_task tk = tasks[it.prob];
  const int bw = tk.bw, ld = tk.ld;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = it.blk_b >= 0 ? 2 : 1;
  const int ld4 = ld >> 2;
  if (tid == 0) s_rot = 0;

  // stage the blocks: column c of block h -> cols[(h*bw + c)*ld ...]
  for (int h = 0; h < nblk; ++h) {
    const int blk = h == 0 ? it.blk_a : it.blk_b;
    const float4* src = reinterpret_cast<const float4*>(tk.x + (int64_t)blk * bw * ld);
    float4* dst = reinterpret_cast<float4*>(cols + (int64_t)h * bw * ld);
    for (int e = tid; e < bw * ld4; e += kJacThreads) dst[e] = src[e];
  }
  __syncthreads();

  const float fl = floor2[it.prob];
  int nrot = 0;
  if (it.kind == 1) {
    // cross pairs: step s pairs column w of A with column (w+s)%bw of B
    for (int s = 0; s < bw; ++s) {
      if (warp < bw) {
        float* x = cols + (int64_t)warp * ld;
        float* y = cols + (int64_t)(bw + (warp + s) % bw) * ld;
        nrot += jacobi_rotate_pair(x, y, ld, lane, tol, fl);
      }
      __syncthreads();
    }
  } else {
    const int half = bw >> 1;
    for (int r = 0; r < bw - 1; ++r) {
      if (warp < half * nblk) {
        const int h = warp / half, q = warp - h * half;
        int p0, p1;
        rr_pair(bw, r, q, p0, p1);
        float* x = cols + (int64_t)(h * bw + (p0 < p1 ? p0 : p1)) * ld;
        float* y = cols + (int64_t)(h * bw + (p0 < p1 ? p1 : p0)) * ld;
        nrot += jacobi_rotate_pair(x, y, ld, lane, tol, fl);
      }
      __syncthreads();
    }
  }
  if (lane == 0 && nrot) atomicAdd(&s_rot, nrot);

  for (int h = 0; h < nblk; ++h) {
    const int blk = h == 0 ? it.blk_a : it.blk_b;
    float4* dst = reinterpret_cast<float4*>(tk.x + (int64_t)blk * bw * ld);
    const float4* src = reinterpret_cast<const float4*>(cols + (int64_t)h * bw * ld);
    for (int e = tid; e < bw * ld4; e += kJacThreads) dst[e] = src[e];
  }
  __syncthreads();
  if (tid == 0 && s_rot) atomicAdd(counts + it.prob, s_rot);
}

// done[p] |= (counts[p] == 0); counts[p] = 0
__global__ void jacobi_sweep_end_kernel(int32_t* counts, int32_t* done, int32_t* sweeps, int n) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (!done[p]) {
    sweeps[p] += 1;
    if (counts[p] == 0) done[p] = 1;
  }
  counts[p] = 0;
}

// ---------------------------------------------------------------------------------------------
// dominant-r selection
// ---------------------------------------------------------------------------------------------
constexpr int kSelThreads = 512;

__global__ void __launch_bounds__(kSelThreads) select_kernel(const tta_select_task* __restrict__ tasks) {
  extern __shared__ float s_lam[];  // k norms, then k ints (rank position)
  const tta_select_task tk = tasks[blockIdx.x];
  int* s_pos = reinterpret_cast<int*>(s_lam + tk.k);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarp = kSelThreads / 32;

  for (int j = warp; j < tk.k; j += nwarp) {
    const float* x = tk.x + (int64_t)j * tk.ld;
    float a = 0.f;
    for (int e = lane; e < tk.k; e += 32) a = fmaf(x[e], x[e], a);
    a = warp_sum(a);
    if (lane == 0) s_lam[j] = sqrtf(a);
  }
  __syncthreads();
  float lmax = 0.f;
  for (int j = 0; j < tk.k; ++j) lmax = fmaxf(lmax, s_lam[j]);
  // descending rank with index tie-break
  for (int j = tid; j < tk.k; j += kSelThreads) {
    const float lj = s_lam[j];
    int pos = 0;
    for (int i = 0; i < tk.k; ++i) {
      const float li = s_lam[i];
      pos += (li > lj) || (li == lj && i < j);
    }
    s_pos[j] = pos;
  }
  __syncthreads();
  const float cut = lmax * (kJacFloorRel * 4.f);
  for (int j = warp; j < tk.k; j += nwarp) {
    const int pos = s_pos[j];
    if (pos >= tk.r) continue;
    const float lam = s_lam[j];
    const bool live = lam > cut && lam > 0.f;
    const float inv = live ? 1.f / lam : 0.f;
    const float sg = live ? sqrtf(lam) : 0.f;
    const float* x = tk.x + (int64_t)j * tk.ld;
    for (int e = lane; e < tk.k; e += 32) {
      const float v = x[e] * inv;
      tk.e[(int64_t)pos * tk.k + e] = v;
      if (tk.et) tk.et[(int64_t)e * tk.r + pos] = v;
      if (tk.se) tk.se[(int64_t)pos * tk.k + e] = v * sg;
    }
    if (lane == 0) {
      if (tk.sigma) tk.sigma[pos] = sg;
      if (tk.isigma) tk.isigma[pos] = live ? 1.f / sg : 0.f;
    }
  }
}

static bool g_prof_on = false;
static double g_prof_ms = 0.0;
static unsigned long long g_prof_launches = 0;

static void build_schedule(const tta_eig_task* th, int n, std::vector<std::vector<JacItem>>& launches) {
  launches.clear();
  for (int p = 0; p < n; ++p) {
    const int bw = th[p].bw;
    const int nb = th[p].kpad / bw;
    const int nlaunch = nb <= 1 ? 1 : 1 + ((nb & 1) ? nb : nb - 1);
    if ((int)launches.size() < nlaunch) launches.resize(nlaunch);
    for (int b = 0; b < nb; b += 2) launches[0].push_back({p, b, (b + 1 < nb) ? b + 1 : -1, 0});
    if (nb > 1) {
      const int ne = (nb & 1) ? nb + 1 : nb;
      for (int r = 0; r < ne - 1; ++r)
        for (int q = 0; q < ne / 2; ++q) {
          int p0, p1;
          rr_pair(ne, r, q, p0, p1);
          if (p0 >= nb || p1 >= nb) continue;
          launches[1 + r].push_back({p, p0 < p1 ? p0 : p1, p0 < p1 ? p1 : p0, 1});
        }
    }
  }
}

}  // namespace tta

extern "C" {

void tta_jacobi_profile_enable(int on) { tta::g_prof_on = on != 0; }
void tta_jacobi_profile_read(double* step_ms, unsigned long long* step_launches) {
  if (step_ms) *step_ms = tta::g_prof_ms;
  if (step_launches) *step_launches = tta::g_prof_launches;
  tta::g_prof_ms = 0.0;
  tta::g_prof_launches = 0;
}

size_t tta_jacobi_scratch_bytes(const tta_eig_task* tasks_host, int n_tasks) {
  using namespace tta;
  if (!tasks_host || n_tasks <= 0) return 0;
  std::vector<std::vector<JacItem>> launches;
  build_schedule(tasks_host, n_tasks, launches);
  size_t items = 0;
  for (auto& l : launches) items += l.size();
  // counts, done, sweeps (int32 each), floor2 (float), items
  return (size_t)n_tasks * 4 * sizeof(int32_t) + items * sizeof(JacItem) + 64;
}

int tta_jacobi_eigh_batched(const tta_eig_task* tasks_dev, const tta_eig_task* tasks_host, int n_tasks, float tol,
                            int max_sweeps, int32_t* scratch_dev, size_t scratch_bytes, int32_t* sweeps_out,
                            void* stream) {
  using namespace tta;
  if (n_tasks <= 0) return TTA_OK;
  if (!tasks_dev || !tasks_host || !scratch_dev || max_sweeps <= 0 || !(tol > 0.f)) {
    set_error("jacobi: bad argument");
    return TTA_E_INVALID;
  }
  size_t smem = 0;
  for (int p = 0; p < n_tasks; ++p) {
    const tta_eig_task& tk = tasks_host[p];
    if ((tk.bw != 8 && tk.bw != 16) || tk.k <= 0 || tk.ld < tk.k || (tk.ld & 3) || tk.kpad < tk.k ||
        (tk.kpad % tk.bw) || !tk.x) {
      set_error("jacobi: task %d invalid (k=%d ld=%d kpad=%d bw=%d)", p, tk.k, tk.ld, tk.kpad, tk.bw);
      return TTA_E_INVALID;
    }
    const size_t need = (size_t)2 * tk.bw * tk.ld * sizeof(float);
    smem = need > smem ? need : smem;
  }
  if (smem > 227 * 1024) {
    set_error("jacobi: column blocks need %zu B of shared memory (> 227 KB); use bw=8", smem);
    return TTA_E_INVALID;
  }
  if (scratch_bytes < tta_jacobi_scratch_bytes(tasks_host, n_tasks)) {
    set_error("jacobi: scratch too small");
    return TTA_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<std::vector<JacItem>> launches;
  build_schedule(tasks_host, n_tasks, launches);

  int32_t* counts = scratch_dev;
  int32_t* done = scratch_dev + n_tasks;
  int32_t* sweeps = scratch_dev + 2 * n_tasks;
  float* floor2 = reinterpret_cast<float*>(scratch_dev + 3 * n_tasks);
  JacItem* items_dev = reinterpret_cast<JacItem*>(scratch_dev + 4 * n_tasks);

  std::vector<JacItem> flat;
  std::vector<size_t> offs;
  for (auto& l : launches) {
    offs.push_back(flat.size());
    flat.insert(flat.end(), l.begin(), l.end());
  }
  int rc = check_cuda(cudaMemsetAsync(scratch_dev, 0, (size_t)n_tasks * 4 * sizeof(int32_t), st), "jacobi memset");
  if (rc) return rc;
  rc = check_cuda(cudaMemcpyAsync(items_dev, flat.data(), flat.size() * sizeof(JacItem), cudaMemcpyHostToDevice, st),
                  "jacobi schedule upload");
  if (rc) return rc;
  rc = check_cuda(cudaFuncSetAttribute(jacobi_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "jacobi smem attribute");
  if (rc) return rc;

  jacobi_floor_kernel<<<n_tasks, 256, 0, st>>>(tasks_dev, floor2);
  TTA_CHECK_LAUNCH("jacobi floor launch");

  std::vector<int32_t> done_h(n_tasks, 0), sweeps_h(n_tasks, 0);
  // launches needed by each problem
  std::vector<int> nl(n_tasks);
  for (int p = 0; p < n_tasks; ++p) {
    const int nb = tasks_host[p].kpad / tasks_host[p].bw;
    nl[p] = nb <= 1 ? 1 : 1 + ((nb & 1) ? nb : nb - 1);
  }
  bool all_done = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_prof_on) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
  }
  for (int sweep = 0; sweep < max_sweeps && !all_done; ++sweep) {
    int need = 0;
    for (int p = 0; p < n_tasks; ++p)
      if (!done_h[p]) need = nl[p] > need ? nl[p] : need;
    if (ev0) cudaEventRecord(ev0, st);
    for (int l = 0; l < need; ++l) {
      const int cnt = (int)launches[l].size();
      if (cnt == 0) continue;
      jacobi_step_kernel<<<cnt, kJacThreads, smem, st>>>(items_dev + offs[l], tasks_dev, counts, done, floor2, tol);
      TTA_CHECK_LAUNCH("jacobi step launch");
      if (ev0) ++g_prof_launches;
    }
    if (ev1) cudaEventRecord(ev1, st);
    jacobi_sweep_end_kernel<<<(n_tasks + 127) / 128, 128, 0, st>>>(counts, done, sweeps, n_tasks);
    TTA_CHECK_LAUNCH("jacobi sweep_end launch");
    rc = check_cuda(cudaMemcpyAsync(done_h.data(), done, n_tasks * sizeof(int32_t), cudaMemcpyDeviceToHost, st),
                    "jacobi done readback");
    if (rc) return rc;
    rc = check_cuda(cudaStreamSynchronize(st), "jacobi sweep sync");
    if (rc) return rc;
    if (ev0) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) g_prof_ms += ms;
    }
    all_done = true;
    for (int p = 0; p < n_tasks; ++p) all_done = all_done && done_h[p];
  }
  if (ev0) {
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  }
  rc = check_cuda(cudaMemcpyAsync(sweeps_h.data(), sweeps, n_tasks * sizeof(int32_t), cudaMemcpyDeviceToHost, st),
                  "jacobi sweeps readback");
  if (rc) return rc;
  rc = check_cuda(cudaStreamSynchronize(st), "jacobi final sync");
  if (rc) return rc;
  if (sweeps_out) memcpy(sweeps_out, sweeps_h.data(), n_tasks * sizeof(int32_t));
  if (!all_done) {
    int bad = 0;
    for (int p = 0; p < n_tasks; ++p)
      if (!done_h[p]) bad = p;
    set_error("jacobi: problem %d (k=%d) not converged after %d sweeps", bad, tasks_host[bad].k, max_sweeps);
    return TTA_E_NOCONV;
  }
  return TTA_OK;
}

int tta_select_batched(const tta_select_task* tasks_dev, const tta_select_task* tasks_host, int n_tasks,
                       void* stream) {
  using namespace tta;
  if (n_tasks <= 0) return TTA_OK;
  if (!tasks_dev || !tasks_host) {
    set_error("select: null table");
    return TTA_E_INVALID;
  }
  size_t smem = 0;
  for (int t = 0; t < n_tasks; ++t) {
    const tta_select_task& tk = tasks_host[t];
    if (tk.k <= 0 || tk.r <= 0 || tk.r > tk.k || tk.ld < tk.k || !tk.x || !tk.e) {
      set_error("select: task %d invalid (k=%d r=%d ld=%d)", t, tk.k, tk.r, tk.ld);
      return TTA_E_INVALID;
    }
    const size_t need = (size_t)tk.k * 8;
    smem = need > smem ? need : smem;
  }
  if (smem > 48 * 1024) {
    int rc = check_cuda(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "select smem attribute");
    if (rc) return rc;
  }
  select_kernel<<<n_tasks, kSelThreads, smem, (cudaStream_t)stream>>>(tasks_dev);
  TTA_CHECK_LAUNCH("select launch");
  return TTA_OK;
}
}
