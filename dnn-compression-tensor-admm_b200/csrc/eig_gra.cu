// Gram-rotate-apply cluster Jacobi eigensolver (32 < k <= 512): the one-sided Jacobi sweep of
// eig_cluster.cu with its bulk work recast as small dense contractions.
//
// One cluster of P = kpad/32 CTAs owns one eigenproblem for all of its sweeps; every CTA holds two
// blocks of 16 columns (column-major, padded stride `lds`) in shared memory.  One round of the block
// tournament used to be 16 lock-step steps of "load y column, dot, rotate, store y column" -- 64 KB
// of shared-memory traffic and a ~1 us dependent chain per step.  Here a round is
//   1. Gram    C = [T B]^T [T B]   (only the 16 x 16 cross block T^T B is contracted over the k rows;
//                                   the diagonal blocks T^T T, B^T B travel with their blocks and are
//                                   refreshed once per sweep)                       FFMA2, k*256 FMA
//   2. rotate  the same 16 x 16 cross pairs in the same order, but as a two-sided Jacobi on the
//              32 x 32 matrix C in shared memory (one thread per 2 x 2 block, one barrier per step),
//              accumulating the rotations in V (32 x 32)                            latency, tiny
//   3. apply   [T' B'] = [T B] V, each lane two rows in registers, written STRAIGHT into the shared
//              memory of the CTAs that own the blocks in the next round (distributed shared memory,
//              other-parity buffer), then ONE cluster barrier                       FFMA2, k*1024 FMA
// so the column state is read twice and written once per round instead of 16 times, and the
// arithmetic is identical to the rotation sequence of the column-wise solver (rotating the columns of
// M by J is the congruence J^T C J of its Gram matrix).
//
// Stopping rule: a sweep whose largest |c_pq| / sqrt(c_pp c_qq) (seen before rotating) is below
// `stop_rel` leaves off-diagonals of order stop_rel^2 (quadratic convergence), so the solver stops
// after that sweep instead of spending one or two more sweeps to see zero rotations; the fp64
// refinement (refine.cu) works on the result either way.
#include <cooperative_groups.h>

#include <stdlib.h>

#include <vector>

#include "eig_gra_device.cuh"

namespace cg = cooperative_groups;

namespace tta {

template <int RL>
__global__ void __launch_bounds__(kGraThreads, RL == 4 ? 2 : 1)
    jacobi_gra_kernel(const tta_eig_task* __restrict__ tasks, const int32_t* __restrict__ prob_ids,
                      int32_t* __restrict__ sweeps_out, int32_t* __restrict__ status_out,
                      const float* __restrict__ floor2, float tol2, float stop2, int max_sweeps,
                      unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(16) float smem[];
  long long pc[4] = {0, 0, 0, 0};   // phase cycles (gram, rotate, apply, barrier) when `prof` is given
  long long t0 = 0;
#define GRA_TICK(i)                         \
  if (prof) {                               \
    const long long t1 = clock64();         \
    pc[i] += t1 - t0;                       \
    t0 = t1;                                \
  }
  __shared__ int s_rot;
  __shared__ unsigned s_max;
  __shared__ int s_counts[16];
  __shared__ unsigned s_maxes[16];
  __shared__ int s_total;
  __shared__ unsigned s_maxall;

  cg::cluster_group cluster = cg::this_cluster();
  const int P = (int)cluster.num_blocks();
  const int c = (int)cluster.block_rank();
  const int prob = prob_ids[blockIdx.x / P];
  const tta_eig_task tk = tasks[prob];
  const float fl = floor2[prob] * (kJacFloorRel * kJacFloorRel);
  const int ld = tk.ld, ld4 = ld >> 2, lds = gra_lds(ld);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bufsz = gra_buf_floats(ld);
  float* part = smem + 2 * bufsz;       // 16 x 256 Gram partials
  float* Cb = part + 4096;              // 2 x (32 x 33)
  float* Vm = Cb + 2 * kGraCsz;         // 32 x 32 (stride kGraVs)
  float4* rec = reinterpret_cast<float4*>(Vm + 32 * kGraVs);   // 2 x 16 rotation records
  int* pq = reinterpret_cast<int*>(rec + 32);                  // 16 pairs (intra steps)
  const int blkf = 16 * lds;            // floats per block of columns

  // initial state: block slot c -> top, slot P + c -> bottom of parity 0
  {
    float* buf = smem;
    for (int e = tid; e < 32 * ld4; e += kGraThreads) {
      const int col = e / ld4, r4 = e - col * ld4;
      const int slot = col < 16 ? c : P + c;
      const float4 v = __ldcg(reinterpret_cast<const float4*>(tk.x + ((int64_t)slot * 16 + (col & 15)) * ld) + r4);
      *reinterpret_cast<float4*>(buf + col * lds + 4 * r4) = v;
    }
  }
  if (tid == 0) {
    s_rot = 0;
    s_max = 0u;
  }
  __syncthreads();

  // circle-method rotation of the block slots: t0 fixed; t_c -> t_{c+1}; t_{P-1} -> b_{P-1};
  // b_c -> b_{c-1}; b_0 -> t_1   (slot index < P: top of CTA index; >= P: bottom of CTA index - P)
  const int top_dst = (c == 0) ? 0 : (c == P - 1 ? (2 * P - 1) : c + 1);
  const int bot_dst = (c == 0) ? 1 : (P + c - 1);

  int par = 0;
  int sweep = 0, converged = 0;
  int nrot = 0;
  float maxrel2 = 0.f;
  while (sweep < max_sweeps) {
    for (int round = -1; round < 2 * P - 1; ++round) {
      float* buf = smem + par * bufsz;
      float* nbuf = smem + (par ^ 1) * bufsz;
      float* C0 = Cb;   // gram results go to buffer 0 of C
      if (prof) t0 = clock64();
      // ---- 1. Gram ----
      if (round < 0) {
        // pairs inside T and inside B: refresh both diagonal blocks from the columns
        gra_gram16(buf, buf, part, ld4, lds, warp, lane);
        __syncthreads();
        if (tid < 256) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < 16; ++w) v += part[w * 256 + tid];
          const int i = tid >> 4, j = tid & 15;
          C0[i * kGraCs + j] = v;
          C0[i * kGraCs + 16 + j] = 0.f;
          C0[(16 + i) * kGraCs + j] = 0.f;
        }
        __syncthreads();
        gra_gram16(buf + blkf, buf + blkf, part, ld4, lds, warp, lane);
        __syncthreads();
        if (tid < 256) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < 16; ++w) v += part[w * 256 + tid];
          const int i = tid >> 4, j = tid & 15;
          C0[(16 + i) * kGraCs + 16 + j] = v;
        }
      } else {
        gra_gram16(buf, buf + blkf, part, ld4, lds, warp, lane);
        __syncthreads();
        if (tid < 256) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < 16; ++w) v += part[w * 256 + tid];
          const int i = tid >> 4, j = tid & 15;
          C0[i * kGraCs + 16 + j] = v;
          C0[(16 + j) * kGraCs + i] = v;
        } else {
          // carried diagonal blocks
          const int u = tid - 256, i = u >> 4, j = u & 15;
          C0[i * kGraCs + j] = buf[32 * lds + u];
          C0[(16 + i) * kGraCs + 16 + j] = buf[32 * lds + 256 + u];
        }
      }
      for (int e = tid; e < 1024; e += kGraThreads) Vm[(e >> 5) * kGraVs + (e & 31)] = ((e >> 5) == (e & 31)) ? 1.f : 0.f;
      __syncthreads();

      GRA_TICK(0)
      // ---- 2. rotate ----
      const int cur = (round < 0) ? gra_rotate_intra(Cb, Vm, rec, pq, tid, tol2, fl, nrot, maxrel2)
                                  : gra_rotate_cross(Cb, Vm, rec, tid, tol2, fl, nrot, maxrel2);
      const float* Cf = Cb + cur * kGraCsz;
      // Renormalise the columns of V.  cs and sn are rounded to fp32, so every plane rotation scales its
      // two columns by 1 + O(6e-8) -- coherently over all k rows, unlike the per-element rounding of a
      // column-wise rotation.  Without this the column norms (the eigenvalue estimates) drift by ~1e-4
      // over the ~300 rounds of a solve.
      if (tid < 32) {
        float ss = 0.f;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) ss = fmaf(Vm[i * kGraVs + tid], Vm[i * kGraVs + tid], ss);
        float rn = mufu_rsqrt(ss);
        rn = rn * fmaf(-0.5f * ss, rn * rn, 1.5f);
#pragma unroll 8
        for (int i = 0; i < 32; ++i) Vm[i * kGraVs + tid] *= rn;
      }
      __syncthreads();
      GRA_TICK(1)

      // ---- 3. apply, straight into the next owners' other-parity buffers ----
      float* tbuf = nbuf;   // destination buffers (base of the other-parity buffer of the next owner)
      float* bbuf = nbuf;
      int tslot = 0, bslot = 1;   // 0: top half, 1: bottom half of that buffer
      if (round >= 0 && P > 1) {
        tslot = top_dst < P ? 0 : 1;
        bslot = bot_dst < P ? 0 : 1;
        tbuf = cluster.map_shared_rank(nbuf, top_dst - tslot * P);
        bbuf = cluster.map_shared_rank(nbuf, bot_dst - bslot * P);
      }
      gra_apply<RL>(buf, Vm, tbuf + tslot * blkf, bbuf + bslot * blkf, ld, lds, warp, lane);
      // the diagonal Gram blocks travel with their columns
      if (tid < 256) {
        const int i = tid >> 4, j = tid & 15;
        tbuf[32 * lds + tslot * 256 + tid] = Cf[i * kGraCs + j];
      } else {
        const int u = tid - 256, i = u >> 4, j = u & 15;
        bbuf[32 * lds + bslot * 256 + u] = Cf[(16 + i) * kGraCs + 16 + j];
      }
      __syncthreads();
      GRA_TICK(2)
      gra_cluster_sync();
      GRA_TICK(3)
      par ^= 1;
    }
    ++sweep;
    // ---- convergence vote over the cluster ----
    if (nrot) atomicAdd(&s_rot, nrot);
    if (maxrel2 > 0.f) atomicMax(&s_max, __float_as_uint(maxrel2));
    nrot = 0;
    maxrel2 = 0.f;
    __syncthreads();
    int total;
    unsigned mx;
    if (P == 1) {
      total = s_rot;
      mx = s_max;
      __syncthreads();
      if (tid == 0) {
        s_rot = 0;
        s_max = 0u;
      }
    } else {
      if (tid == 0) {
        int* rc = cluster.map_shared_rank(s_counts, 0);
        unsigned* rm = cluster.map_shared_rank(s_maxes, 0);
        rc[c] = s_rot;
        rm[c] = s_max;
        s_rot = 0;
        s_max = 0u;
      }
      gra_cluster_sync();
      if (tid == 0) {
        const int* rc = cluster.map_shared_rank(s_counts, 0);
        const unsigned* rm = cluster.map_shared_rank(s_maxes, 0);
        int t = 0;
        unsigned m = 0u;
        for (int i = 0; i < P; ++i) {
          t += rc[i];
          m = rm[i] > m ? rm[i] : m;
        }
        s_total = t;
        s_maxall = m;
      }
      __syncthreads();
      total = s_total;
      mx = s_maxall;
      gra_cluster_sync();   // CTA 0's vote arrays may be rewritten after this point
    }
    if (total == 0 || __uint_as_float(mx) < stop2) {
      converged = 1;
      break;
    }
  }

  // final state back to global X (column order is irrelevant to the selection stage)
  {
    const float* buf = smem + par * bufsz;
    for (int e = tid; e < 32 * ld4; e += kGraThreads) {
      const int col = e / ld4, r4 = e - col * ld4;
      const int slot = col < 16 ? c : P + c;
      const float4 v = *reinterpret_cast<const float4*>(buf + col * lds + 4 * r4);
      __stcg(reinterpret_cast<float4*>(tk.x + ((int64_t)slot * 16 + (col & 15)) * ld) + r4, v);
    }
  }
  if (c == 0 && tid == 0) {
    sweeps_out[prob] = sweep;
    status_out[prob] = converged;
  }
  if (prof && tid == 0 && c == P - 1)
    for (int i = 0; i < 4; ++i) atomicAdd(prof + i, (unsigned long long)pc[i]);
#undef GRA_TICK
}

bool jacobi_gra_eligible(const tta_eig_task& tk) {
  if (tk.bw != 16 || tk.ld > 512 || tk.ld < 4 || (tk.ld & 3)) return false;
  if (tk.kpad % 32) return false;
  const int P = tk.kpad / 32;
  return P >= 2 && P <= 16;
}

static int gra_rl(int ld) { return ld <= 256 ? 4 : 8; }

template <int RL>
static int gra_launch(int P, int nprob, size_t smem, cudaStream_t gs, const tta_eig_task* tasks_dev, const int32_t* ids,
                      int32_t* sweeps_dev, int32_t* status_dev, const float* floor2, float tol2, float stop2,
                      int max_sweeps, unsigned long long* prof) {
  static size_t smem_set = 0;
  static bool np_set = false;
  int rc;
  if (!np_set) {
    rc = check_cuda(cudaFuncSetAttribute(jacobi_gra_kernel<RL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                    "jacobi gra non-portable attribute");
    if (rc) return rc;
    np_set = true;
  }
  if (smem > smem_set) {
    rc = check_cuda(cudaFuncSetAttribute(jacobi_gra_kernel<RL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "jacobi gra smem attribute");
    if (rc) return rc;
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nprob * P), 1, 1);
  cfg.blockDim = dim3(kGraThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = gs;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  rc = check_cuda(cudaLaunchKernelEx(&cfg, jacobi_gra_kernel<RL>, tasks_dev, ids, sweeps_dev, status_dev, floor2, tol2,
                                     stop2, max_sweeps, prof),
                  "jacobi gra launch");
  if (rc) return rc;
  count_launch();
  return TTA_OK;
}

// Enqueue one launch (on `gs`) for the problems `ids` (device array, all with the same cluster size P
// and the same output-column split).
int jacobi_gra_enqueue(const tta_eig_task* tasks_dev, const tta_eig_task* th, const std::vector<int>& probs, int P,
                       float tol2, float stop2, int max_sweeps, const int32_t* ids_dev, int32_t* sweeps_dev,
                       int32_t* status_dev, const float* floor2, cudaStream_t gs) {
  if (probs.empty()) return TTA_OK;
  int ldmax = 0;
  for (int p : probs) ldmax = th[p].ld > ldmax ? th[p].ld : ldmax;
  const int rl = gra_rl(ldmax);
  for (int p : probs)
    if (gra_rl(th[p].ld) != rl) {
      set_error("jacobi gra: mixed column lengths in one launch group");
      return TTA_E_INVALID;
    }
  const size_t smem = gra_smem_bytes(ldmax);
  const int n = (int)probs.size();
  // TTA_GRA_PROF=1: per-phase cycle counts of the last CTA of every cluster, printed by the next call
  static unsigned long long* prof = nullptr;
  static bool prof_init = false;
  if (!prof_init) {
    prof_init = true;
    const char* e = getenv("TTA_GRA_PROF");
    if (e && e[0] == '1' && cudaMalloc(&prof, 4 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(prof, 0, 4 * sizeof(unsigned long long));
  }
  if (prof) {
    unsigned long long h[4];
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    if (h[0] | h[1] | h[2] | h[3])
      fprintf(stderr, "[gra prof] cycles gram %llu rotate %llu apply %llu barrier %llu\n", h[0], h[1], h[2], h[3]);
    cudaMemset(prof, 0, sizeof(h));
  }
  if (rl == 4)
    return gra_launch<4>(P, n, smem, gs, tasks_dev, ids_dev, sweeps_dev, status_dev, floor2, tol2, stop2, max_sweeps, prof);
  return gra_launch<8>(P, n, smem, gs, tasks_dev, ids_dev, sweeps_dev, status_dev, floor2, tol2, stop2, max_sweeps, prof);
}

}  // namespace tta
