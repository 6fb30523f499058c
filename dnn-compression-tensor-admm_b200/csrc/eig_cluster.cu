// Persistent thread-block-cluster Jacobi eigensolver (k <= 512): one cluster of P CTAs owns one
// eigenproblem for ALL of its sweeps.
//
// The multi-launch solver in eig.cu pays one kernel launch (~13 us of fixed cost on B200 for a
// 512-thread / 64 KB CTA) per block round, runs all problems of a wave in lock step and needs a host
// synchronisation per sweep.  Here instead:
//   * the kpad = 2*P*bw columns live in the shared memory of the P CTAs of one cluster (2*bw columns
//     each; every warp keeps its own column in registers during a block round);
//   * one sweep = the pairs inside each block + a (2P-1)-round circle tournament over the 2P blocks;
//     between rounds the blocks rotate between CTAs through L2 (global X) bracketed by two hardware
//     cluster barriers -- no kernel boundary, no host involvement;
//   * convergence ("no rotation in a full sweep") is decided inside the cluster through distributed
//     shared memory, so every problem stops by itself and small problems free their SMs early;
//   * independent problems are independent clusters of the same launch (problems with different P
//     go to concurrent internal streams).
#include <cooperative_groups.h>

#include <string.h>

#include <algorithm>
#include <map>
#include <utility>
#include <vector>

#include "eig_device.cuh"
#include "stream_pool.cuh"

namespace cg = cooperative_groups;

namespace tta {


// Cluster-wide barrier with release/acquire semantics.  Written as inline PTX with "memory" clobbers:
// the compiler must not move shared/global accesses across either half (with the cooperative-groups
// builtins ptxas was seen hoisting a global load between arrive and wait).  The leading
// __syncthreads() orders this CTA's own shared-memory traffic first.
__device__ __forceinline__ void cluster_sync_all() {
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void cl_copy_block(float* __restrict__ dst, const float* __restrict__ src, int n4, int tid,
                                              int nthreads, bool from_global) {
  const float4* s = reinterpret_cast<const float4*>(src);
  float4* d = reinterpret_cast<float4*>(dst);
  if (from_global) {
    for (int e = tid; e < n4; e += nthreads) d[e] = __ldcg(s + e);   // L2-coherent: written by a peer CTA
  } else {
    for (int e = tid; e < n4; e += nthreads) __stcg(d + e, s[e]);
  }
}

template <int NV>
__device__ __forceinline__ void cluster_body(cg::cluster_group& cluster, float* __restrict__ cols, int* s_rot,
                                             int* s_total, int* s_counts, const tta_eig_task& tk, int prob,
                                             int32_t* __restrict__ sweeps_out, int32_t* __restrict__ status_out,
                                             float fl, float tol2, float stop2, int max_sweeps) {
  const int P = (int)cluster.num_blocks();
  const int c = (int)cluster.block_rank();
  const int bw = tk.bw, ld = tk.ld;
  const int tid = threadIdx.x, nthreads = blockDim.x, warp = tid >> 5, lane = tid & 31;
  const int blk4 = bw * (ld >> 2);                 // float4 per block
  const int64_t blk = (int64_t)bw * ld;            // floats per block
  float* top = cols;
  float* bot = cols + blk;
  float* nrm = cols + 2 * blk;                     // cached squared norms of the 2*bw resident columns

  // block slots in global X: top row 0..P-1, bottom row P..2P-1
  cl_copy_block(top, tk.x + (int64_t)c * blk, blk4, tid, nthreads, true);
  cl_copy_block(bot, tk.x + (int64_t)(P + c) * blk, blk4, tid, nthreads, true);
  if (tid == 0) *s_rot = 0;
  __syncthreads();

  // circle-method rotation of the block slots: t0 fixed; t_c -> t_{c+1}; t_{P-1} -> b_{P-1};
  // b_c -> b_{c-1}; b_0 -> t_1.
  const int top_dst = (c == 0) ? 0 : (c == P - 1 ? (2 * P - 1) : c + 1);
  const int bot_dst = (c == 0) ? 1 : (P + c - 1);

  int sweep = 0;
  int converged = 0;
#ifdef TTA_DEBUG_CLUSTER
  if (tid == 0) printf("[cl] blk %d P %d c %d prob %d bw %d ld %d top_dst %d bot_dst %d\n", (int)blockIdx.x, P, c, prob, bw, ld, top_dst, bot_dst);
#endif
  while (sweep < max_sweeps) {
    int nrot = jacobi_block<NV>(cols, nrm, 0, 2, bw, ld, warp, lane, tol2, fl, stop2);     // pairs inside both blocks
    for (int round = 0; round < 2 * P - 1; ++round) {
      nrot += jacobi_block<NV>(cols, nrm, 1, 2, bw, ld, warp, lane, tol2, fl, stop2);      // top x bottom pairs
      if (P > 1) {
        // jacobi_block ends with __syncthreads(): shared columns are final for this round
        if (c != 0) cl_copy_block(tk.x + (int64_t)top_dst * blk, top, blk4, tid, nthreads, false);
        cl_copy_block(tk.x + (int64_t)bot_dst * blk, bot, blk4, tid, nthreads, false);
        cluster_sync_all();                          // release/acquire at cluster scope: blocks visible in L2
        if (c != 0) cl_copy_block(top, tk.x + (int64_t)c * blk, blk4, tid, nthreads, true);
        cl_copy_block(bot, tk.x + (int64_t)(P + c) * blk, blk4, tid, nthreads, true);
        cluster_sync_all();                          // nobody overwrites a slot that is still being read
      }
    }
    ++sweep;
    // ---- convergence: total number of rotations of this sweep over the cluster ----
    if (lane == 0 && nrot) atomicAdd(s_rot, nrot);
    __syncthreads();
    int total;
    if (P == 1) {
      total = *s_rot;
      __syncthreads();
      if (tid == 0) *s_rot = 0;
    } else {
      if (tid == 0) {
        int* remote = cluster.map_shared_rank(s_counts, 0);
        remote[c] = *s_rot;
        *s_rot = 0;
      }
      cluster_sync_all();
      if (tid == 0) {
        const int* remote = cluster.map_shared_rank(s_counts, 0);
        int t = 0;
        for (int i = 0; i < P; ++i) t += remote[i];
        *s_total = t;
      }
      __syncthreads();
      total = *s_total;
      cluster_sync_all();                            // s_counts of CTA 0 may be rewritten after this point
    }
#ifdef TTA_DEBUG_CLUSTER
    if (tid == 0 && sweep <= 12) printf("[cl] c %d sweep %d total %d\n", c, sweep, total);
#endif
    // `total` = rotations of the sweep + 65536 per rotation of a pair that was further from orthogonal than the
    // stop threshold (seen before rotating).  No such pair => the sweep left off-diagonals of order stop^2
    // (quadratic convergence), the same rule as the gram-rotate-apply solver: stop instead of spending two or
    // three more sweeps at the fp32 noise floor waiting for a sweep without any rotation.
    if (total == 0 || (total >> 16) == 0) {
      converged = 1;
      break;
    }
  }

  // final state back to global X (column order is irrelevant to the selection stage)
  cl_copy_block(tk.x + (int64_t)c * blk, top, blk4, tid, nthreads, false);
  cl_copy_block(tk.x + (int64_t)(P + c) * blk, bot, blk4, tid, nthreads, false);
  if (c == 0 && tid == 0) {
    sweeps_out[prob] = sweep;
    status_out[prob] = converged;
  }
}

// THREADS = 512: bw <= 16, two CTAs per SM (a network step has more cluster CTAs than the 148 SMs, and
// a cluster that cannot be co-scheduled waits for a whole eigensolve); THREADS = 1024: bw <= 32.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 1)
    jacobi_cluster_kernel(const tta_eig_task* __restrict__ tasks, const int32_t* __restrict__ prob_ids,
                          int32_t* __restrict__ sweeps_out, int32_t* __restrict__ status_out,
                          const float* __restrict__ floor2, float tol2, float stop2, int max_sweeps) {
  extern __shared__ __align__(16) float cols[];
  __shared__ int s_rot;
  __shared__ int s_total;
  __shared__ int s_counts[16];
  cg::cluster_group cluster = cg::this_cluster();
  const int prob = prob_ids[blockIdx.x / cluster.num_blocks()];
  const tta_eig_task tk = tasks[prob];
  const float fl = floor2[prob] * (kJacFloorRel * kJacFloorRel);
  switch ((tk.ld + 127) >> 7) {
    case 1: cluster_body<1>(cluster, cols, &s_rot, &s_total, s_counts, tk, prob, sweeps_out, status_out, fl, tol2, stop2, max_sweeps); break;
    case 2: cluster_body<2>(cluster, cols, &s_rot, &s_total, s_counts, tk, prob, sweeps_out, status_out, fl, tol2, stop2, max_sweeps); break;
    case 3: cluster_body<3>(cluster, cols, &s_rot, &s_total, s_counts, tk, prob, sweeps_out, status_out, fl, tol2, stop2, max_sweeps); break;
    default: cluster_body<4>(cluster, cols, &s_rot, &s_total, s_counts, tk, prob, sweeps_out, status_out, fl, tol2, stop2, max_sweeps); break;
  }
}

// One pool per (device, calling stream): independent callers (the layer groups of admm.ADMM.update run on
// their own streams) must not share internal streams, or the launches of one caller would queue behind
// the other's in stream order although the problems are independent.
StreamPool* pool_for(int dev, cudaStream_t caller) {
  static std::map<std::pair<int, cudaStream_t>, StreamPool*> pools;
  const auto key = std::make_pair(dev, caller);
  auto it = pools.find(key);
  if (it != pools.end()) return it->second;
  StreamPool* p = new StreamPool();
  // Stream i serves the i-th largest cluster size of a call: descending priority, so that when SMs
  // free up the block scheduler places the big clusters (the critical path of a wave) first and the
  // small problems fill the SMs that are left.  A caller on a default-priority stream starts one level
  // lower than a caller on a high-priority stream (the critical layer group).  Streams are created on
  // first use: every stream occupies one of the device's hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS),
  // and streams that share a queue serialise.
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // hi is numerically the smallest
  int caller_prio = 0;
  if (caller == nullptr || cudaStreamGetPriority(caller, &caller_prio) != cudaSuccess) caller_prio = 0;
  p->base_prio = prio_hi + (caller_prio < 0 ? 0 : 1);
  p->prio_lo = prio_lo;
  if (cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  pools[key] = p;
  return p;
}

bool jacobi_cluster_eligible(const tta_eig_task& tk) {
  if (tk.ld > 512 || tk.bw < 2 || (tk.bw & 1) || tk.bw > 32) return false;
  if (tk.kpad % (2 * tk.bw)) return false;
  const int P = tk.kpad / (2 * tk.bw);
  return P == 1 || P == 2 || P == 4 || P == 8 || P == 16;
}

// One launch (on `gs`) of the column-rotation cluster kernel for problems that all use cluster size P.
static int cluster_enqueue(const tta_eig_task* tasks_dev, const tta_eig_task* th, const std::vector<int>& probs, int P,
                           float tol2, float stop2, int max_sweeps, const int32_t* ids_dev, int32_t* sweeps_dev,
                           int32_t* status_dev, const float* floor2, cudaStream_t gs) {
  static bool attr_set = false;
  static size_t smem_set[2] = {0, 0};
  int rc;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(jacobi_cluster_kernel<512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                    "jacobi cluster non-portable attribute");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(jacobi_cluster_kernel<1024>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                    "jacobi cluster non-portable attribute");
    if (rc) return rc;
    attr_set = true;
  }
  size_t smem = 0;
  int bwmax = 2;
  for (int p : probs) {
    const size_t need = (size_t)2 * th[p].bw * (th[p].ld + 1) * sizeof(float);
    smem = need > smem ? need : smem;
    bwmax = th[p].bw > bwmax ? th[p].bw : bwmax;
  }
  const int big = bwmax > 16 ? 1 : 0;
  if (smem > smem_set[big]) {
    rc = big ? check_cuda(cudaFuncSetAttribute(jacobi_cluster_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem), "jacobi cluster smem attribute")
             : check_cuda(cudaFuncSetAttribute(jacobi_cluster_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem), "jacobi cluster smem attribute");
    if (rc) return rc;
    smem_set[big] = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(probs.size() * P), 1, 1);
  cfg.blockDim = dim3((unsigned)(bwmax * 32), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = gs;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!big)
    rc = check_cuda(cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel<512>, tasks_dev, ids_dev, sweeps_dev, status_dev,
                                       floor2, tol2, stop2, max_sweeps),
                    "jacobi cluster launch");
  else
    rc = check_cuda(cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel<1024>, tasks_dev, ids_dev, sweeps_dev, status_dev,
                                       floor2, tol2, stop2, max_sweeps),
                    "jacobi cluster launch");
  if (rc) return rc;
  count_launch();
  return TTA_OK;
}

// Enqueue the persistent solvers for the problems listed in `probs` (indices into the task table):
// problems are grouped by (solver, cluster size); every group is one launch on an internal stream,
// forked from / joined back into `st`, largest clusters first (they are the critical path).
// `ids_dev` must hold probs.size() int32.
int jacobi_cluster_run(const tta_eig_task* tasks_dev, const tta_eig_task* th, const std::vector<int>& probs,
                       float tol2, float stop2, int max_sweeps, bool allow_gra, int32_t* ids_dev, int32_t* sweeps_dev,
                       int32_t* status_dev, const float* floor2, cudaStream_t st) {
  if (probs.empty()) return TTA_OK;
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  StreamPool* pool = pool_for(dev, st);
  if (!pool) {
    set_error("jacobi cluster: cannot create internal streams");
    return TTA_E_CUDA;
  }
  // key = P (column-rotation kernel) or 100 + P (gram-rotate-apply kernel); iterate descending P
  std::map<int, std::vector<int>> groups;
  for (int p : probs) {
    if (allow_gra && jacobi_gra_eligible(th[p]))
      groups[100 + th[p].kpad / 32].push_back(p);
    else
      groups[th[p].kpad / (2 * th[p].bw)].push_back(p);
  }
  std::vector<int> order;
  for (auto& kv : groups) order.push_back(kv.first);
  std::sort(order.begin(), order.end(), [](int a, int b) { return (a % 100) > (b % 100) || ((a % 100) == (b % 100) && a > b); });
  std::vector<int32_t> flat;
  std::map<int, int> offs;
  for (int key : order) {
    offs[key] = (int)flat.size();
    flat.insert(flat.end(), groups[key].begin(), groups[key].end());
  }
  const void* ids_src = flat.data();
  if (void* slot = pinned_stage(flat.size() * sizeof(int32_t))) {   // pinned source: the copy only enqueues
    memcpy(slot, flat.data(), flat.size() * sizeof(int32_t));
    ids_src = slot;
  }
  rc = check_cuda(cudaMemcpyAsync(ids_dev, ids_src, flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st),
                  "jacobi cluster ids upload");
  if (rc) return rc;
  rc = check_cuda(cudaEventRecord(pool->fork, st), "jacobi cluster fork");
  if (rc) return rc;
  bool used[kPoolStreams] = {};
  int gi = 0;
  for (int key : order) {
    const int si = gi++ % kPoolStreams;
    cudaStream_t gs = pool->get(si);
    if (!gs) {
      set_error("jacobi cluster: cannot create an internal stream");
      return TTA_E_CUDA;
    }
    if (!used[si]) {
      rc = check_cuda(cudaStreamWaitEvent(gs, pool->fork, 0), "jacobi cluster stream wait");
      if (rc) return rc;
      used[si] = true;
    }
    const std::vector<int>& g = groups[key];
    const int32_t* ids = ids_dev + offs[key];
    if (key >= 100)
      rc = jacobi_gra_enqueue(tasks_dev, th, g, key - 100, tol2, stop2, max_sweeps, ids, sweeps_dev, status_dev, floor2, gs);
    else
      rc = cluster_enqueue(tasks_dev, th, g, key, tol2, stop2, max_sweeps, ids, sweeps_dev, status_dev, floor2, gs);
    if (rc) return rc;
  }
  for (int si = 0; si < kPoolStreams; ++si) {
    if (!used[si]) continue;
    rc = check_cuda(cudaEventRecord(pool->join[si], pool->s[si]), "jacobi cluster join record");
    if (rc) return rc;
    rc = check_cuda(cudaStreamWaitEvent(st, pool->join[si], 0), "jacobi cluster join wait");
    if (rc) return rc;
  }
  return TTA_OK;
}

}  // namespace tta
