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

#include "eig_device.cuh"

namespace tta {

struct JacItem {
  int32_t prob;
  int32_t blk_a;
  int32_t blk_b;  // -1: single block (intra launch, odd block count)
  int32_t kind;   // 0 intra, 1 cross
};

constexpr int kJacMaxBw = 16;
constexpr int kJacThreads = kJacMaxBw * 32;

// floor2[p] = max_j ||x_j||^2 (raw bits of a non-negative float, so atomicMax on the integer view is the
// float maximum); the solvers scale it by floor_rel^2 when they read it.  kFloorSplit CTAs per problem.
constexpr int kFloorSplit = 8;
__global__ void __launch_bounds__(256) jacobi_floor_kernel(const tta_eig_task* __restrict__ tasks,
                                                          float* __restrict__ floor2) {
  __shared__ float s_max[8];
  const tta_eig_task tk = tasks[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = 0.f;
  for (int j = blockIdx.y * 8 + warp; j < tk.k; j += 8 * kFloorSplit) {
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
    atomicMax(reinterpret_cast<unsigned*>(floor2) + blockIdx.x, __float_as_uint(m));
  }
}

__global__ void __launch_bounds__(kJacThreads) jacobi_step_kernel(const JacItem* __restrict__ items,
                                                                  const tta_eig_task* __restrict__ tasks,
                                                                  int32_t* __restrict__ counts,
                                                                  const int32_t* __restrict__ done,
                                                                  const float* __restrict__ floor2, float tol2) {
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

  const float fl = floor2[it.prob] * (kJacFloorRel * kJacFloorRel);
  float* nrm = cols + 2 * bw * ld;     // cached squared norms (2*bw floats behind the columns)
  int nrot;
  const int nv = (ld + 127) >> 7;
  switch (nv) {
    case 1: nrot = jacobi_block<1>(cols, nrm, it.kind, nblk, bw, ld, warp, lane, tol2, fl); break;
    case 2: nrot = jacobi_block<2>(cols, nrm, it.kind, nblk, bw, ld, warp, lane, tol2, fl); break;
    case 3: nrot = jacobi_block<3>(cols, nrm, it.kind, nblk, bw, ld, warp, lane, tol2, fl); break;
    case 4: nrot = jacobi_block<4>(cols, nrm, it.kind, nblk, bw, ld, warp, lane, tol2, fl); break;
    default: nrot = jacobi_block<0>(cols, nrm, it.kind, nblk, bw, ld, warp, lane, tol2, fl); break;
  }
  if (lane == 0 && nrot) atomicAdd(&s_rot, nrot);
  __syncthreads();
  const int total_rot = s_rot;
  if (total_rot == 0) return;   // nothing changed: skip the write-back (uniform across the CTA)

  for (int h = 0; h < nblk; ++h) {
    const int blk = h == 0 ? it.blk_a : it.blk_b;
    float4* dst = reinterpret_cast<float4*>(tk.x + (int64_t)blk * bw * ld);
    const float4* src = reinterpret_cast<const float4*>(cols + (int64_t)h * bw * ld);
    for (int e = tid; e < bw * ld4; e += kJacThreads) dst[e] = src[e];
  }
  if (tid == 0) atomicAdd(counts + it.prob, total_rot);
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

static bool g_force_legacy = false;   // test hook: route every problem through the multi-launch solver

static bool g_allow_gra = true;       // test hook: 0 keeps every cluster problem on the column-rotation kernel
static float g_stop_rel = 3e-4f;      // gram-rotate-apply solver: stop after a sweep whose largest relative
                                      // off-diagonal (seen before rotating) is below this

static bool use_cluster(const tta_eig_task& tk) {
  if (g_force_legacy) return false;
  return (g_allow_gra && jacobi_gra_eligible(tk)) || jacobi_cluster_eligible(tk);
}

static void build_schedule(const tta_eig_task* th, int n, std::vector<std::vector<JacItem>>& launches) {
  launches.clear();
  for (int p = 0; p < n; ++p) {
    if (use_cluster(th[p])) continue;
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
void tta_jacobi_force_multilaunch(int on) { tta::g_force_legacy = on != 0; }
void tta_jacobi_enable_gra(int on) { tta::g_allow_gra = on != 0; }
void tta_jacobi_set_stop_rel(float stop_rel) { tta::g_stop_rel = stop_rel > 0.f ? stop_rel : 0.f; }
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
  // counts, done, sweeps (int32 each), floor2 (float), cluster ids, cluster status, items
  return (size_t)n_tasks * 6 * sizeof(int32_t) + items * sizeof(JacItem) + 64;
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
  std::vector<int> cl_probs;
  int n_legacy = 0;
  for (int p = 0; p < n_tasks; ++p) {
    const tta_eig_task& tk = tasks_host[p];
    if (tk.bw < 2 || tk.bw > 32 || (tk.bw & 1) || tk.k <= 0 || tk.ld < tk.k || (tk.ld & 3) || tk.kpad < tk.k ||
        (tk.kpad % tk.bw) || !tk.x) {
      set_error("jacobi: task %d invalid (k=%d ld=%d kpad=%d bw=%d)", p, tk.k, tk.ld, tk.kpad, tk.bw);
      return TTA_E_INVALID;
    }
    if (use_cluster(tk)) {
      cl_probs.push_back(p);
      continue;
    }
    if (tk.bw > kJacMaxBw) {
      set_error("jacobi: task %d (k=%d) needs the multi-launch solver, which supports bw <= %d", p, tk.k, kJacMaxBw);
      return TTA_E_INVALID;
    }
    ++n_legacy;
    const size_t need = (size_t)2 * tk.bw * (tk.ld + 1) * sizeof(float);
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
  int32_t* cl_ids = scratch_dev + 4 * n_tasks;
  int32_t* cl_status = scratch_dev + 5 * n_tasks;
  JacItem* items_dev = reinterpret_cast<JacItem*>(scratch_dev + 6 * n_tasks);

  std::vector<JacItem> flat;
  std::vector<size_t> offs;
  for (auto& l : launches) {
    offs.push_back(flat.size());
    flat.insert(flat.end(), l.begin(), l.end());
  }
  int rc = check_cuda(cudaMemsetAsync(scratch_dev, 0, (size_t)n_tasks * 6 * sizeof(int32_t), st), "jacobi memset");
  if (rc) return rc;
  std::vector<int32_t> done_h(n_tasks, 0), sweeps_h(n_tasks, 0);
  for (int p : cl_probs) done_h[p] = 1;     // owned by the cluster solver: the step kernels skip them
  if (n_legacy) {
    rc = check_cuda(cudaMemcpyAsync(done, done_h.data(), n_tasks * sizeof(int32_t), cudaMemcpyHostToDevice, st),
                    "jacobi done upload");
    if (rc) return rc;
    rc = check_cuda(cudaMemcpyAsync(items_dev, flat.data(), flat.size() * sizeof(JacItem), cudaMemcpyHostToDevice, st),
                    "jacobi schedule upload");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(jacobi_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "jacobi smem attribute");
    if (rc) return rc;
  }

  jacobi_floor_kernel<<<dim3(n_tasks, kFloorSplit), 256, 0, st>>>(tasks_dev, floor2);
  TTA_CHECK_LAUNCH("jacobi floor launch");

  cudaEvent_t cev0 = nullptr, cev1 = nullptr;
  if (g_prof_on && !cl_probs.empty()) {
    cudaEventCreate(&cev0);
    cudaEventCreate(&cev1);
    cudaEventRecord(cev0, st);
  }
  rc = jacobi_cluster_run(tasks_dev, tasks_host, cl_probs, tol * tol, g_stop_rel * g_stop_rel, max_sweeps, g_allow_gra,
                          cl_ids, sweeps, cl_status, floor2, st);
  if (rc) return rc;
  if (cev0) cudaEventRecord(cev1, st);

  // launches needed by each problem
  std::vector<int> nl(n_tasks);
  for (int p = 0; p < n_tasks; ++p) {
    const int nb = tasks_host[p].kpad / tasks_host[p].bw;
    nl[p] = nb <= 1 ? 1 : 1 + ((nb & 1) ? nb : nb - 1);
  }
  bool all_done = (n_legacy == 0);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_prof_on && n_legacy) {
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
      jacobi_step_kernel<<<cnt, kJacThreads, smem, st>>>(items_dev + offs[l], tasks_dev, counts, done, floor2,
                                                         tol * tol);
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
  if (!sweeps_out && n_legacy == 0 && !cev0) {
    // asynchronous mode: nothing is read back here; the caller inspects the scratch buffer later
    // (tta_jacobi_read_results) -- a network step then has no host synchronisation between waves.
    return TTA_OK;
  }
  std::vector<int32_t> status_h(n_tasks, 0);
  rc = check_cuda(cudaMemcpyAsync(sweeps_h.data(), sweeps, n_tasks * sizeof(int32_t), cudaMemcpyDeviceToHost, st),
                  "jacobi sweeps readback");
  if (rc) return rc;
  rc = check_cuda(cudaMemcpyAsync(status_h.data(), cl_status, n_tasks * sizeof(int32_t), cudaMemcpyDeviceToHost, st),
                  "jacobi status readback");
  if (rc) return rc;
  rc = check_cuda(cudaStreamSynchronize(st), "jacobi final sync");
  if (rc) return rc;
  if (cev0) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, cev0, cev1) == cudaSuccess) g_prof_ms += ms;
    g_prof_launches += 1;
    cudaEventDestroy(cev0);
    cudaEventDestroy(cev1);
  }
  for (int p : cl_probs) {
    done_h[p] = status_h[p];
    all_done = all_done && status_h[p];
  }
  if (n_legacy) {
    // the multi-launch solver tracks convergence in done[]; mirror it into the status slots, which is what
    // tta_jacobi_read_results (the plans' collect step) inspects
    rc = check_cuda(cudaMemcpyAsync(cl_status, done_h.data(), n_tasks * sizeof(int32_t), cudaMemcpyHostToDevice, st),
                    "jacobi status upload");
    if (rc) return rc;
    rc = check_cuda(cudaStreamSynchronize(st), "jacobi status sync");
    if (rc) return rc;
  }
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

int tta_jacobi_read_results(const int32_t* scratch_host, const tta_eig_task* tasks_host, int n_tasks, int max_sweeps,
                            int32_t* sweeps_out) {
  using namespace tta;
  if (n_tasks <= 0) return TTA_OK;
  if (!scratch_host || !tasks_host) {
    set_error("jacobi results: bad argument");
    return TTA_E_INVALID;
  }
  const int32_t* sweeps = scratch_host + 2 * n_tasks;
  const int32_t* status = scratch_host + 5 * n_tasks;
  int bad = -1;
  for (int p = 0; p < n_tasks; ++p) {
    if (sweeps_out) sweeps_out[p] = sweeps[p];
    if (!status[p]) bad = p;
  }
  if (bad >= 0) {
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
