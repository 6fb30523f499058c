// Dominant-r symmetric eigensolver in fp64 (32 < k <= 608): replaces numpy.linalg.svd / LAPACK gesdd of
// ttd.py:17 and admm.py:131,143 (and tensorly's partial_svd, admm.py:116,124) on the Gram matrix of an unfolding.
//
//   A. trd_reduce_kernel       Householder tridiagonalisation G = Q T Q^T.  One thread-block cluster per problem,
//                              G distributed by columns (cyclic) over the shared memory of the P CTAs, ONE
//                              exchange + cluster barrier per Householder step (k - 2 steps).
//   B. trd_eigval_kernel       the r largest eigenvalues of T by multisection on the Sturm count.
//   C. trd_eigvec_kernel       their eigenvectors from the twisted factorisation of T - lambda I.
//   D. trd_backtransform_kernel  x = H_0 H_1 ... H_{k-3} z, one warp per pair of vectors, reflectors in pairs.
// tta_refine_finalize_batched then turns (e64, lam) into the fp32 outputs of the projection step.
//
// Why not Jacobi (eig_gra.cu): a k = 512 Jacobi solve is ~5000 dependent rotation steps in ~300 cluster
// rounds (4.1 ms) and needs an fp64 refinement afterwards; the tridiagonalisation is k - 2 = 510 dependent
// steps, everything after it is embarrassingly parallel, and all of it is fp64, so no refinement.
//
// Step j of the reduction (LAPACK dsytd2 'L' arithmetic), A = current trailing matrix, v_j the reflector:
//     p = tau_j A v_j,   w = p - (tau_j/2)(p.v_j) v_j,   A <- A - v_j w^T - w v_j^T
// Column i lives in CTA i mod P.  A is symmetric, so the CTA that owns column i also owns element i of every
// ROW: after its part of the update it holds element i of the next column j + 1 = row j + 1.  Per step a
// CTA therefore sends, for each of its columns i, the pair (p_i, A[j+1][i]) to every CTA (remote shared-memory
// stores), one cluster barrier follows, and every WARP of every CTA rebuilds w_j, column j + 1 and the next
// reflector v_{j+1} redundantly in registers (two warp reductions, no block barrier, no broadcast from an
// "owner").  The rank-2 update of step j is fused with the dot products of step j + 1: each column element is
// read and written once per step.
#include <cooperative_groups.h>

#include <algorithm>
#include <map>
#include <vector>

#include "stream_pool.cuh"
#include "trd_device.cuh"
#include "tta_common.cuh"

namespace cg = cooperative_groups;

namespace tta {

constexpr int kTrdMaxK = 608;      // 16 CTAs x 38 columns x 640 rows x 8 B = 194.5 KB of shared memory
constexpr int kTrdMinK = 3;
constexpr int kTrdMaxP = 16;

__host__ __device__ inline int trd_cluster_size(int k) {
  const int p = (k + 31) / 32;
  return p < 1 ? 1 : (p > kTrdMaxP ? kTrdMaxP : p);
}
// rows per lane (template parameter of the kernels): k <= 32 * NR
__host__ __device__ inline int trd_nr(int k) {
  return k <= 64 ? 2 : k <= 128 ? 4 : k <= 256 ? 8 : k <= 384 ? 12 : k <= 512 ? 16 : 20;
}

__device__ __forceinline__ void trd_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ double trd_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// a[t] for a warp-uniform runtime t (register arrays need static indices)
template <int NR>
__device__ __forceinline__ double trd_pick(const double (&a)[NR], int t) {
  double r = 0.0;
#pragma unroll
  for (int q = 0; q < NR; ++q)
    if (q == t) r = a[q];
  return r;
}

// ------------------------------------------------------------------------------------------------------
// A. tridiagonalisation
// ------------------------------------------------------------------------------------------------------
template <int NR, int NT, int MC>
__global__ void __launch_bounds__(NT, 1) trd_reduce_kernel(const tta_symeig_task* __restrict__ tasks, int first) {
  extern __shared__ __align__(16) double trd_sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int P = (int)cluster.num_blocks();
  const int c = (int)cluster.block_rank();
  const tta_symeig_task tk = tasks[first + blockIdx.x / P];
  const int k = tk.k;
  constexpr int KP = NR * 32;
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nclmax = (k + P - 1) / P;
  const int ncl = (k - c + P - 1) / P;          // own columns: global index i = c + P * li
  double* A = trd_sm;                           // nclmax columns of KP rows
  double* pbuf = A + (size_t)nclmax * KP;       // [2][KP]  p_i of the step, by parity
  double* sbuf = pbuf + 2 * KP;                 // [2][KP]  row j + 1 of the matrix before the update of step j
  const TrdLayout L = trd_layout(k, tk.r);
  double* gd = tk.work + L.d;
  double* ge = tk.work + L.e;
  double* gtau = tk.work + L.tau;
  double* gv = tk.work + L.v;

  for (int li = warp; li < ncl; li += NW) {
    const double* src = tk.g + (int64_t)(c + P * li) * k;   // row i == column i
    double* dst = A + (size_t)li * KP;
    for (int r = lane; r < KP; r += 32) dst[r] = r < k ? src[r] : 0.0;
  }
  for (int e = tid; e < 2 * KP; e += NT) {
    pbuf[e] = 0.0;
    sbuf[e] = 0.0;
  }
  if (c == 0 && tid == 0) *tk.status = 0;
  __syncthreads();
  // virtual step -1 (parity 1): no update, "row 0 before the update" = column 0 of G
  for (int r = tid; r < k; r += NT) sbuf[KP + r] = tk.g[r];
  __syncthreads();
  trd_cluster_sync();   // every CTA's exchange buffers are initialised before the first remote store

  double vprev[NR], wprev[NR], vcur[NR];
#pragma unroll
  for (int t = 0; t < NR; ++t) vprev[t] = wprev[t] = vcur[t] = 0.0;
  double tau_cur = 0.0;
  double vcolp[MC], wcolp[MC];   // v_{j-1}, w_{j-1} at this warp's own column indices
#pragma unroll
  for (int m = 0; m < MC; ++m) vcolp[m] = wcolp[m] = 0.0;

  for (int j = -1; j <= k - 3; ++j) {
    if (j >= 0) {
      // ---- fused pass: update of step j - 1, dot products of step j, row j + 1 ----
      const int t0 = (j + 1) >> 5;
      double dot[MC], sv[MC];
#pragma unroll
      for (int m = 0; m < MC; ++m) {
        dot[m] = 0.0;
        sv[m] = 0.0;
        const int li = warp + NW * m;
        const int i = c + P * li;
        if (li < ncl && i > j) {   // warp-uniform
          double* colp = A + (size_t)li * KP + lane;
          const double wc = wcolp[m], vc = vcolp[m];
#pragma unroll
          for (int t = 0; t < NR; ++t) {
            if (t >= t0) {
              double a = colp[32 * t];
              a = fma(-vprev[t], wc, a);
              a = fma(-wprev[t], vc, a);
              colp[32 * t] = a;
              dot[m] = fma(a, vcur[t], dot[m]);
              if (lane + 32 * t == j + 1) sv[m] = a;
            }
          }
        }
      }
#pragma unroll
      for (int m = 0; m < MC; ++m) {
        dot[m] = warp_sum(dot[m]) * tau_cur;
        sv[m] = trd_bcast(sv[m], (j + 1) & 31);
      }
      if (lane < P) {
        double* rp = cluster.map_shared_rank(pbuf, lane) + (j & 1) * KP;
        double* rs = cluster.map_shared_rank(sbuf, lane) + (j & 1) * KP;
#pragma unroll
        for (int m = 0; m < MC; ++m) {
          const int li = warp + NW * m;
          const int i = c + P * li;
          if (li < ncl && i > j) {
            rp[i] = dot[m];
            rs[i] = sv[m];
          }
        }
      }
      trd_cluster_sync();
    }

    // ---- scalar phase of step j (every warp, redundantly): w_j, column j + 1, reflector v_{j+1} ----
    const double* pb = pbuf + (j & 1) * KP;
    const double* sb = sbuf + (j & 1) * KP;
    double w[NR], col[NR];
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      const int idx = lane + 32 * t;
      w[t] = pb[idx];
      col[t] = sb[idx];
      acc = fma(w[t], vcur[t], acc);
    }
    acc = warp_sum(acc);
    const double K = 0.5 * tau_cur * acc;
#pragma unroll
    for (int t = 0; t < NR; ++t) w[t] = (lane + 32 * t > j) ? fma(-K, vcur[t], w[t]) : 0.0;
    const int j1 = j + 1;
    const double wj1 = trd_bcast(trd_pick<NR>(w, j1 >> 5), j1 & 31);
#pragma unroll
    for (int t = 0; t < NR; ++t) col[t] = (col[t] - w[t]) - wj1 * vcur[t];   // column j1 after update j, entries >= j1
    const double dj1 = trd_bcast(trd_pick<NR>(col, j1 >> 5), j1 & 31);
    double tau_n = 0.0, beta = 0.0;
    double vnew[NR];
    if (j1 <= k - 3) {
      const int j2 = j1 + 1;
      const double alpha = trd_bcast(trd_pick<NR>(col, j2 >> 5), j2 & 31);
      double sg = 0.0;
#pragma unroll
      for (int t = 0; t < NR; ++t)
        if (lane + 32 * t > j2) sg = fma(col[t], col[t], sg);
      sg = warp_sum(sg);
      double inv = 0.0;
      beta = alpha;
      if (sg > 0.0) {
        beta = -copysign(sqrt(fma(alpha, alpha, sg)), alpha);
        tau_n = (beta - alpha) / beta;
        inv = 1.0 / (alpha - beta);
      }
#pragma unroll
      for (int t = 0; t < NR; ++t) {
        const int idx = lane + 32 * t;
        vnew[t] = idx > j2 ? col[t] * inv : (idx == j2 ? 1.0 : 0.0);
      }
      if (c == (j1 % P) && warp == NW - 1) {
        double* row = gv + (int64_t)j1 * L.kp;
#pragma unroll
        for (int t = 0; t < NR; ++t)
          if (lane + 32 * t < L.kp) row[lane + 32 * t] = vnew[t];
      }
    } else {   // j1 == k - 2: the last off-diagonal element, no reflector
      beta = trd_bcast(trd_pick<NR>(col, (k - 1) >> 5), (k - 1) & 31);
#pragma unroll
      for (int t = 0; t < NR; ++t) vnew[t] = 0.0;
    }
    if (c == 0 && tid == 0) {
      gd[j1] = dj1;
      ge[j1] = beta;
      gtau[j1] = tau_n;
    }
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      vprev[t] = vcur[t];
      wprev[t] = w[t];
      vcur[t] = vnew[t];
    }
    tau_cur = tau_n;
#pragma unroll
    for (int m = 0; m < MC; ++m) {
      const int i = c + P * (warp + NW * m);
      const int ii = i < KP ? i : 0;
      vcolp[m] = trd_bcast(trd_pick<NR>(vprev, ii >> 5), ii & 31);
      wcolp[m] = trd_bcast(trd_pick<NR>(wprev, ii >> 5), ii & 31);
    }
  }

  // d[k-1]: element (k-1, k-1) after the pending update of step k - 3
#pragma unroll
  for (int m = 0; m < MC; ++m) {
    const int li = warp + NW * m;
    const int i = c + P * li;
    if (li < ncl && i == k - 1 && lane == ((k - 1) & 31)) {
      gd[k - 1] = A[(size_t)li * KP + (k - 1)] - 2.0 * vcolp[m] * wcolp[m];
      ge[k - 1] = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// B. eigenvalues: 128 shifts per eigenvalue and pass (7 bits), 8 passes
// ------------------------------------------------------------------------------------------------------
constexpr int kTrdEvalThreads = 512;
constexpr int kTrdEvalPerBlock = 4;
constexpr int kTrdEvalPasses = 8;

__global__ void __launch_bounds__(kTrdEvalThreads) trd_eigval_kernel(const tta_symeig_task* __restrict__ tasks) {
  extern __shared__ __align__(16) double ev_sm[];
  __shared__ double s_red[kTrdEvalThreads / 32];
  __shared__ int s_cnt[kTrdEvalPerBlock][4];
  const tta_symeig_task tk = tasks[blockIdx.y];
  const int k = tk.k, r = tk.r;
  if ((int)blockIdx.x * kTrdEvalPerBlock >= r) return;
  const TrdLayout L = trd_layout(k, r);
  double* dd = ev_sm;
  double* ee2 = ev_sm + L.kp;
  const double* gd = tk.work + L.d;
  const double* ge = tk.work + L.e;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // |T|_inf bound -> power-of-two scale (exact), so that every eigenvalue lies in [-1, 1]
  double mx = 0.0;
  for (int i = tid; i < k; i += kTrdEvalThreads) {
    const double el = i > 0 ? fabs(ge[i - 1]) : 0.0, er = i < k - 1 ? fabs(ge[i]) : 0.0;
    mx = fmax(mx, fabs(gd[i]) + el + er);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = 0.0;
  for (int q = 0; q < kTrdEvalThreads / 32; ++q) mx = fmax(mx, s_red[q]);
  int ex = mx > 0.0 ? trd_exponent(mx) + 1 : 0;
  ex = ex > 1000 ? 1000 : (ex < -1000 ? -1000 : ex);
  const double sc = trd_pow2(-ex);
  for (int i = tid; i < k; i += kTrdEvalThreads) {
    dd[i] = gd[i] * sc;
    const double e = i > 0 ? ge[i - 1] * sc : 0.0;
    ee2[i] = e * e;
  }
  __syncthreads();

  const int g = tid >> 7, s = tid & 127;
  const int p = blockIdx.x * kTrdEvalPerBlock + g;
  const bool live = p < r;
  const int m = k - 1 - p;            // ascending index of the p-th largest eigenvalue
  double lo = -1.0 - 1e-9, hi = 1.0 + 1e-9;
  for (int pass = 0; pass < kTrdEvalPasses; ++pass) {
    const double step = (hi - lo) * (1.0 / 129.0);
    const double x = fma(step, (double)(s + 1), lo);
    const int cnt = live ? trd_sturm_count(dd, ee2, k, x) : 0;
    const unsigned b = __ballot_sync(0xffffffffu, live && cnt <= m);   // x <= lambda_m
    if (lane == 0) s_cnt[g][warp & 3] = __popc(b);
    __syncthreads();
    const int nle = s_cnt[g][0] + s_cnt[g][1] + s_cnt[g][2] + s_cnt[g][3];
    __syncthreads();
    const double nlo = fma(step, (double)nle, lo);
    const double nhi = nle < 128 ? fma(step, (double)(nle + 1), lo) : hi;
    lo = nlo;
    hi = nhi;
  }
  if (live && s == 0) {
    const double lam = 0.5 * (lo + hi);
    tk.work[L.lams + p] = lam;
    tk.lam[p] = lam * trd_pow2(ex);
  }
  if (blockIdx.x == 0 && tid == 0) tk.work[L.hdr] = sc;
}

// ------------------------------------------------------------------------------------------------------
// C. eigenvectors of T: twisted factorisation, two threads (forward / backward sweep) per vector
// ------------------------------------------------------------------------------------------------------
constexpr int kTrdEvecThreads = 128;
constexpr double kTrdClusterGap = 1e-9;    // scaled units: closer eigenvalue pairs are reported (status 1)
constexpr double kTrdLiveRel = 4e-7;       // same cut as refine.cu: smaller eigenvalues count as zero

__global__ void __launch_bounds__(kTrdEvecThreads) trd_eigvec_kernel(const tta_symeig_task* __restrict__ tasks) {
  extern __shared__ __align__(16) double evc_sm[];
  const tta_symeig_task tk = tasks[blockIdx.y];
  const int k = tk.k, r = tk.r;
  if ((int)blockIdx.x * (kTrdEvecThreads / 2) >= r) return;
  const TrdLayout L = trd_layout(k, r);
  double* dd = evc_sm;
  double* ee = evc_sm + L.kp;
  const double sc = tk.work[L.hdr];
  const int tid = threadIdx.x;
  for (int i = tid; i < k; i += kTrdEvecThreads) {
    dd[i] = tk.work[L.d + i] * sc;
    ee[i] = i < k - 1 ? tk.work[L.e + i] * sc : 0.0;
  }
  __syncthreads();
  const int p = blockIdx.x * (kTrdEvecThreads / 2) + (tid >> 1);
  const int dir = tid & 1;
  const bool live = p < r;
  const int pc = live ? p : r - 1;
  const double lam = tk.work[L.lams + pc];
  double* S = tk.work + L.s + pc;
  double* Pv = tk.work + L.p + pc;
  if (live) {
    if (dir == 0) {
      double q = trd_guard(dd[0] - lam);
      S[0] = q;
      for (int i = 0; i < k - 1; ++i) {
        q = trd_guard((dd[i + 1] - lam) - ee[i] * ee[i] / q);
        S[(int64_t)(i + 1) * r] = q;
      }
    } else {
      double q = trd_guard(dd[k - 1] - lam);
      Pv[(int64_t)(k - 1) * r] = q;
      for (int i = k - 2; i >= 0; --i) {
        q = trd_guard((dd[i] - lam) - ee[i] * ee[i] / q);
        Pv[(int64_t)i * r] = q;
      }
    }
  }
  __syncwarp();
  // twist index: smallest |gamma_i|, gamma_i = s_i + p_i - (d_i - lambda)
  const int half = k >> 1;
  const int ibeg = dir == 0 ? 0 : half, iend = dir == 0 ? half : k;
  double best = 1e300;
  int bi = ibeg;
  if (live) {
    for (int i = ibeg; i < iend; ++i) {
      const double gm = fabs(S[(int64_t)i * r] + Pv[(int64_t)i * r] - (dd[i] - lam));
      if (gm < best) {
        best = gm;
        bi = i;
      }
    }
  }
  const double ob = __shfl_xor_sync(0xffffffffu, best, 1);
  const int oi = __shfl_xor_sync(0xffffffffu, bi, 1);
  const int tw = (ob < best || (ob == best && oi < bi)) ? oi : bi;
  if (live) {
    double* out = tk.e64 + (int64_t)p * k;
    double z = 1.0;
    if (dir == 0) {
      out[tw] = 1.0;
      for (int i = tw - 1; i >= 0; --i) {
        z = -(ee[i] / S[(int64_t)i * r]) * z;
        out[i] = z;
      }
      if (p + 1 < r) {
        const double l0 = tk.work[L.lams], ln = tk.work[L.lams + p + 1];
        if (lam > kTrdLiveRel * l0 && lam - ln < kTrdClusterGap) atomicOr(tk.status, 1);
      }
    } else {
      for (int i = tw; i < k - 1; ++i) {
        z = -(ee[i] / Pv[(int64_t)(i + 1) * r]) * z;
        out[i + 1] = z;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// D. back-transformation, in place on e64: one warp per two vectors, two reflectors per reduction
// ------------------------------------------------------------------------------------------------------
constexpr int kTrdBtWarps = 2;

template <int NR>
__global__ void __launch_bounds__(32 * kTrdBtWarps) trd_backtransform_kernel(const tta_symeig_task* __restrict__ tasks,
                                                                            int first) {
  const tta_symeig_task tk = tasks[first + blockIdx.y];
  const int k = tk.k, r = tk.r;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = 2 * (blockIdx.x * kTrdBtWarps + warp), p1 = p0 + 1;
  if (p0 >= r) return;
  const TrdLayout L = trd_layout(k, r);
  const double* gv = tk.work + L.v;
  const double* gtau = tk.work + L.tau;
  double x0[NR], x1[NR];
#pragma unroll
  for (int t = 0; t < NR; ++t) {
    const int idx = lane + 32 * t;
    x0[t] = idx < k ? tk.e64[(int64_t)p0 * k + idx] : 0.0;
    x1[t] = (idx < k && p1 < r) ? tk.e64[(int64_t)p1 * k + idx] : 0.0;
  }
  double an[NR], bn[NR];   // prefetched reflector pair
  auto fetch = [&](int j) {
    const int t0 = (j > 0 ? j : 0) >> 5;   // row j - 1 is non-zero from index j on
    const double* ra = gv + (int64_t)j * L.kp + lane;
    const double* rb = gv + (int64_t)(j > 0 ? j - 1 : 0) * L.kp + lane;
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      const bool in = t >= t0 && 32 * t < L.kp;
      an[t] = in ? __ldg(ra + 32 * t) : 0.0;
      bn[t] = (in && j > 0) ? __ldg(rb + 32 * t) : 0.0;
    }
  };
  int j = k - 3;
  if (j >= 0) fetch(j);
  for (; j >= 0; j -= 2) {
    double a[NR], b[NR];
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      a[t] = an[t];
      b[t] = bn[t];
    }
    const double ta = gtau[j], tb = j > 0 ? gtau[j - 1] : 0.0;
    if (j - 2 >= 0) fetch(j - 2);
    double ax0 = 0.0, ax1 = 0.0, bx0 = 0.0, bx1 = 0.0, ba = 0.0;
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      ax0 = fma(a[t], x0[t], ax0);
      ax1 = fma(a[t], x1[t], ax1);
      bx0 = fma(b[t], x0[t], bx0);
      bx1 = fma(b[t], x1[t], bx1);
      ba = fma(b[t], a[t], ba);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ax0 += __shfl_xor_sync(0xffffffffu, ax0, o);
      ax1 += __shfl_xor_sync(0xffffffffu, ax1, o);
      bx0 += __shfl_xor_sync(0xffffffffu, bx0, o);
      bx1 += __shfl_xor_sync(0xffffffffu, bx1, o);
      ba += __shfl_xor_sync(0xffffffffu, ba, o);
    }
    // x <- (I - tb b b^T)(I - ta a a^T) x
    const double ca0 = ta * ax0, ca1 = ta * ax1;
    const double cb0 = tb * fma(-ca0, ba, bx0), cb1 = tb * fma(-ca1, ba, bx1);
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      x0[t] = fma(-cb0, b[t], fma(-ca0, a[t], x0[t]));
      x1[t] = fma(-cb1, b[t], fma(-ca1, a[t], x1[t]));
    }
  }
#pragma unroll
  for (int t = 0; t < NR; ++t) {
    const int idx = lane + 32 * t;
    if (idx < k) {
      tk.e64[(int64_t)p0 * k + idx] = x0[t];
      if (p1 < r) tk.e64[(int64_t)p1 * k + idx] = x1[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static size_t trd_reduce_smem(int k, int P, int nr) {
  const int nclmax = (k + P - 1) / P;
  return ((size_t)nclmax * nr * 32 + 4 * (size_t)nr * 32) * sizeof(double);
}

template <int NR, int NT, int MC>
static int trd_launch_reduce(const tta_symeig_task* tasks_dev, int first, int count, int P, size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;
  static bool np_set = false;
  int rc;
  if (!np_set) {
    rc = check_cuda(cudaFuncSetAttribute(trd_reduce_kernel<NR, NT, MC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                    "symeig non-portable cluster attribute");
    if (rc) return rc;
    np_set = true;
  }
  if (smem > smem_set) {
    rc = check_cuda(cudaFuncSetAttribute(trd_reduce_kernel<NR, NT, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "symeig smem attribute");
    if (rc) return rc;
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(count * P), 1, 1);
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  rc = check_cuda(cudaLaunchKernelEx(&cfg, trd_reduce_kernel<NR, NT, MC>, tasks_dev, first), "symeig reduce launch");
  if (rc) return rc;
  count_launch();
  return TTA_OK;
}

static int trd_reduce_dispatch(int nr, const tta_symeig_task* tasks_dev, int first, int count, int P, size_t smem,
                               cudaStream_t st) {
  switch (nr) {
    case 2: return trd_launch_reduce<2, 512, 2>(tasks_dev, first, count, P, smem, st);
    case 4: return trd_launch_reduce<4, 512, 2>(tasks_dev, first, count, P, smem, st);
    case 8: return trd_launch_reduce<8, 512, 2>(tasks_dev, first, count, P, smem, st);
    case 12: return trd_launch_reduce<12, 256, 4>(tasks_dev, first, count, P, smem, st);
    case 16: return trd_launch_reduce<16, 256, 4>(tasks_dev, first, count, P, smem, st);
    default: return trd_launch_reduce<20, 256, 5>(tasks_dev, first, count, P, smem, st);
  }
}

static int trd_backtransform_dispatch(int nr, const tta_symeig_task* tasks_dev, int first, int count, int rmax,
                                      cudaStream_t st) {
  const dim3 grid((unsigned)((rmax + 2 * kTrdBtWarps - 1) / (2 * kTrdBtWarps)), (unsigned)count);
  const int nt = 32 * kTrdBtWarps;
  switch (nr) {
    case 2: trd_backtransform_kernel<2><<<grid, nt, 0, st>>>(tasks_dev, first); break;
    case 4: trd_backtransform_kernel<4><<<grid, nt, 0, st>>>(tasks_dev, first); break;
    case 8: trd_backtransform_kernel<8><<<grid, nt, 0, st>>>(tasks_dev, first); break;
    case 12: trd_backtransform_kernel<12><<<grid, nt, 0, st>>>(tasks_dev, first); break;
    case 16: trd_backtransform_kernel<16><<<grid, nt, 0, st>>>(tasks_dev, first); break;
    default: trd_backtransform_kernel<20><<<grid, nt, 0, st>>>(tasks_dev, first); break;
  }
  TTA_CHECK_LAUNCH("symeig back-transform launch");
  return TTA_OK;
}

}  // namespace tta

extern "C" {

size_t tta_symeig_work_doubles(int k, int r) {
  if (k <= 0 || r <= 0) return 0;
  return (size_t)tta::trd_layout(k, r).total;
}

int tta_symeig_max_k(void) { return tta::kTrdMaxK; }

int tta_symeig_top_batched(const tta_symeig_task* tasks_dev, const tta_symeig_task* tasks_host, int n_tasks,
                           void* stream) {
  using namespace tta;
  if (n_tasks < 0 || (n_tasks > 0 && (!tasks_dev || !tasks_host))) {
    set_error("symeig: bad task table");
    return TTA_E_INVALID;
  }
  if (n_tasks == 0) return TTA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int kmax = 0, rmax = 0;
  for (int t = 0; t < n_tasks; ++t) {
    const tta_symeig_task& tk = tasks_host[t];
    if (tk.k < kTrdMinK || tk.k > kTrdMaxK || tk.r <= 0 || tk.r > tk.k || !tk.g || !tk.work || !tk.lam || !tk.e64 ||
        !tk.status) {
      set_error("symeig: task %d invalid (k=%d r=%d; %d <= k <= %d)", t, tk.k, tk.r, kTrdMinK, kTrdMaxK);
      return TTA_E_INVALID;
    }
    kmax = tk.k > kmax ? tk.k : kmax;
    rmax = tk.r > rmax ? tk.r : rmax;
  }
  // runs of consecutive tasks with the same (cluster size, rows per lane): one launch each, on internal
  // streams so that the clusters of different runs share the GPU (callers sort by descending k)
  struct Run { int first, count, P, nr, kmax, rmax; };
  std::vector<Run> runs;
  for (int t = 0; t < n_tasks; ++t) {
    const int P = trd_cluster_size(tasks_host[t].k), nr = trd_nr(tasks_host[t].k);
    if (runs.empty() || runs.back().P != P || runs.back().nr != nr) runs.push_back({t, 0, P, nr, 0, 0});
    Run& rn = runs.back();
    rn.count++;
    rn.kmax = std::max(rn.kmax, tasks_host[t].k);
    rn.rmax = std::max(rn.rmax, tasks_host[t].r);
  }
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  StreamPool* pool = runs.size() > 1 ? pool_for(dev, st) : nullptr;
  if (runs.size() > 1 && !pool) {
    set_error("symeig: cannot create internal streams");
    return TTA_E_CUDA;
  }
  if (pool) {
    rc = check_cuda(cudaEventRecord(pool->fork, st), "symeig fork");
    if (rc) return rc;
  }
  bool used[kPoolStreams] = {};
  for (size_t ri = 0; ri < runs.size(); ++ri) {
    const Run& rn = runs[ri];
    cudaStream_t gs = st;
    if (pool) {
      const int si = (int)(ri % kPoolStreams);
      gs = pool->get(si);
      if (!gs) {
        set_error("symeig: cannot create an internal stream");
        return TTA_E_CUDA;
      }
      if (!used[si]) {
        rc = check_cuda(cudaStreamWaitEvent(gs, pool->fork, 0), "symeig stream wait");
        if (rc) return rc;
        used[si] = true;
      }
    }
    rc = trd_reduce_dispatch(rn.nr, tasks_dev, rn.first, rn.count, rn.P, trd_reduce_smem(rn.kmax, rn.P, rn.nr), gs);
    if (rc) return rc;
  }
  if (pool) {
    for (int si = 0; si < kPoolStreams; ++si) {
      if (!used[si]) continue;
      rc = check_cuda(cudaEventRecord(pool->join[si], pool->s[si]), "symeig join record");
      if (rc) return rc;
      rc = check_cuda(cudaStreamWaitEvent(st, pool->join[si], 0), "symeig join wait");
      if (rc) return rc;
    }
  }
  const size_t smem = (size_t)2 * ((kmax + 31) & ~31) * sizeof(double);
  trd_eigval_kernel<<<dim3((unsigned)((rmax + kTrdEvalPerBlock - 1) / kTrdEvalPerBlock), (unsigned)n_tasks),
                      kTrdEvalThreads, smem, st>>>(tasks_dev);
  TTA_CHECK_LAUNCH("symeig eigenvalue launch");
  trd_eigvec_kernel<<<dim3((unsigned)((rmax + kTrdEvecThreads / 2 - 1) / (kTrdEvecThreads / 2)), (unsigned)n_tasks),
                      kTrdEvecThreads, smem, st>>>(tasks_dev);
  TTA_CHECK_LAUNCH("symeig eigenvector launch");
  for (const Run& rn : runs) {
    rc = trd_backtransform_dispatch(rn.nr, tasks_dev, rn.first, rn.count, rn.rmax, st);
    if (rc) return rc;
  }
  return TTA_OK;
}
}
