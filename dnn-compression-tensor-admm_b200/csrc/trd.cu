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

// ---- mbarrier / distributed-shared-memory helpers (PTX: no CUDA C++ spelling for st.async) ----
__device__ __forceinline__ uint32_t trd_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t trd_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void trd_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void trd_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void trd_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TRD_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TRD_DONE_%=;\n"
      "bra TRD_WAIT_%=;\n"
      "TRD_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 16-byte store into another CTA's shared memory that signals that CTA's mbarrier (complete_tx has release
// semantics at cluster scope): data and notification travel together, no fence, no cluster barrier
__device__ __forceinline__ void trd_st_async_pair(uint32_t remote_addr, double a, double b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr),
               "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(remote_bar)
               : "memory");
}

__device__ __forceinline__ double trd_rcp(double x) {   // 1/x to ~1 ulp for normal x: MUFU seed + 2 Newton steps
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// ------------------------------------------------------------------------------------------------------
// A. tridiagonalisation
//
// Shared memory of a CTA: its columns A (nclmax x KP doubles), the exchange buffers ps[2][KP] of (p_i, s_i)
// pairs (s_i = element i of row j + 1 before the update of step j), vsh[2][KP] = the current reflector
// (so that scalars of it are plain broadcast loads), two mbarriers.
// Step j:  F_j  every warp updates its columns with (v_{j-1}, w_{j-1}), takes their dot products with v_j and
//               sends (p_i, s_i) to every CTA with st.async (the receiving CTA's mbarrier counts the bytes);
//          wait on the own mbarrier of parity j & 1: all k - 1 - j pairs of the step have landed;
//          S_j  every warp rebuilds w_j, column j + 1 and v_{j+1} in registers, redundantly.
// Buffers alternate by step parity.  A CTA can only send step j + 2 after it has received all of step j + 1,
// which every warp of every CTA sends after it is done reading step j: two buffers suffice.
//
// Register vectors live in a frame RELATIVE to the first live chunk of 32 rows: element [u] of a lane is row
// 32 (tb + u) + lane, tb = (j + 1) / 32.  The two special entries of a step (rows j + 1 and j + 2) are then
// always in chunks 0 / 1, so the only per-lane masks are there, and the dead rows are chunks that have been
// shifted out (one register shift every 32 steps) instead of predicates on every element -- the ALU pipe
// (16 lanes per scheduler) was the bottleneck of the version that masked every chunk.  The chunks past the
// end of the matrix are skipped by picking, per step, the smallest of four unrolled variants that covers the
// live chunks; only its last NR / 4 chunks carry a (uniform) predicate.
// ------------------------------------------------------------------------------------------------------
template <int N>
struct TrdInt { static constexpr int value = N; };

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
  constexpr int QU = NR >= 4 ? NR / 4 : 1;       // chunk granularity of the unrolled variants
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nclmax = (k + P - 1) / P;
  const int ncl = (k - c + P - 1) / P;          // own columns: global index i = c + P * li
  double* A = trd_sm;                                                   // nclmax columns of KP rows
  double2* ps = reinterpret_cast<double2*>(A + (size_t)nclmax * KP);    // [2][KP]
  double* vsh = reinterpret_cast<double*>(ps + 2 * KP);                 // [2][KP]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(vsh + 2 * KP);   // [2]
  const TrdLayout L = trd_layout(k, tk.r);
  double* gd = tk.work + L.d;
  double* ge = tk.work + L.e;
  double* gtau = tk.work + L.tau;
  double* gv = tk.work + L.v;
  const int kchunks = L.kp >> 5;                // chunks that exist in the global reflector rows

  for (int li = warp; li < ncl; li += NW) {
    const double* src = tk.g + (int64_t)(c + P * li) * k;   // row i == column i
    double* dst = A + (size_t)li * KP;
    for (int r = lane; r < KP; r += 32) dst[r] = r < k ? src[r] : 0.0;
  }
  // virtual step -1 (parity 1): no update, p = 0, "row 0 before the update" = column 0 of G
  for (int e = tid; e < KP; e += NT) {
    ps[e] = make_double2(0.0, 0.0);
    ps[KP + e] = make_double2(0.0, e < k ? tk.g[e] : 0.0);
    vsh[e] = 0.0;
    vsh[KP + e] = 0.0;
  }
  if (tid == 0) {
    trd_mbar_init(trd_smem_u32(bars), 1);
    trd_mbar_init(trd_smem_u32(bars + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (c == 0) *tk.status = 0;
  }
  __syncthreads();
  trd_cluster_sync();   // every CTA's buffers and barriers exist before the first remote store

  const uint32_t bar_local = trd_smem_u32(bars);
  uint32_t rps = 0, rbar = 0;     // lane < P: exchange buffer and barriers of CTA `lane`
  if (lane < P) {
    rps = trd_mapa(trd_smem_u32(ps), (uint32_t)lane);
    rbar = trd_mapa(bar_local, (uint32_t)lane);
  }

  double va[NR], vb[NR], wq[NR];
#pragma unroll
  for (int u = 0; u < NR; ++u) va[u] = vb[u] = wq[u] = 0.0;
  double tau_cur = 0.0;
  double vcolp[MC], wcolp[MC];   // v_{j-1}, w_{j-1} at this warp's own column indices
#pragma unroll
  for (int m = 0; m < MC; ++m) vcolp[m] = wcolp[m] = 0.0;

  // ---- F_j + wait ----   vc = v_j, vo = v_{j-1}, wq = w_{j-1};  U = unrolled chunk count >= live chunks
  auto fused = [&](auto utag, const int j, const double(&vc)[NR], const double(&vo)[NR]) {
    constexpr int U = decltype(utag)::value;
    const int par = j & 1;
    if (tid == 0) trd_mbar_expect_tx(bar_local + 8 * par, 16u * (uint32_t)(k - 1 - j));
    const int j1 = j + 1;
    const int tb = j1 >> 5;
    const int nlive = NR - tb;
    double dot[MC], sv[MC];
    double* colp[MC];
    bool on[MC];
#pragma unroll
    for (int m = 0; m < MC; ++m) {
      const int li = warp + NW * m;
      on[m] = li < ncl && (c + P * li) > j;   // warp-uniform
      colp[m] = A + (size_t)(on[m] ? li : 0) * KP + 32 * tb + lane;
      dot[m] = 0.0;
    }
    double an[MC];
#pragma unroll
    for (int m = 0; m < MC; ++m) an[m] = on[m] ? colp[m][0] : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool lv = (u < U - QU) || (u < nlive);           // compile-time true except in the last QU chunks
      const bool lvn = (u + 1 < U - QU) || (u + 1 < nlive);
      double ac[MC];
#pragma unroll
      for (int m = 0; m < MC; ++m) ac[m] = an[m];
      if (u + 1 < U) {
#pragma unroll
        for (int m = 0; m < MC; ++m)
          if (on[m] && lvn) an[m] = colp[m][32 * (u + 1)];    // next chunk's loads before this chunk's stores
      }
#pragma unroll
      for (int m = 0; m < MC; ++m) {
        double a = ac[m];
        a = fma(-vo[u], wcolp[m], a);
        a = fma(-wq[u], vcolp[m], a);
        if (on[m] && lv) colp[m][32 * u] = a;
        dot[m] = fma(a, vc[u], dot[m]);
        if (u == 0) sv[m] = a;
      }
    }
#pragma unroll
    for (int m = 0; m < MC; ++m) {
      dot[m] = warp_sum(dot[m]) * tau_cur;
      sv[m] = trd_bcast(sv[m], j1 & 31);
    }
    if (lane < P) {
#pragma unroll
      for (int m = 0; m < MC; ++m) {
        const int i = c + P * (warp + NW * m);
        if (on[m]) trd_st_async_pair(rps + (uint32_t)(par * KP + i) * 16u, dot[m], sv[m], rbar + 8 * par);
      }
    }
    trd_mbar_wait(bar_local + 8 * par, (uint32_t)((j >> 1) & 1));
  };

  // ---- S_j ----   in: vc = v_j;  out: wq = w_j, vo = v_{j+1}, tau_cur = tau_{j+1}, own-column scalars of (v_j, w_j)
  auto scalar = [&](auto utag, const int j, const double(&vc)[NR], double(&vo)[NR]) {
    constexpr int U = decltype(utag)::value;
    const int par = j & 1;
    const int j1 = j + 1, j2 = j + 2;
    const int tb = j1 >> 5;
    const int nlive = NR - tb;
    const double2* psb = ps + par * KP;
    const double* vs = vsh + par * KP;
    const double2* psl = psb + 32 * tb + lane;
    const int lim = j2 - 32 * tb;            // rows of chunk 0 (and lane 0 of chunk 1 when lim == 32) up to j2
    // One pass, one shuffle tree.  With c1 = 1 + v[j1], A_i = (s_i - p_i) - p_j1 v_i and K = tau/2 sum p_i v_i:
    //   w_i = p_i - K v_i,   column j1 after the update: col_i = A_i + K c1 v_i,
    //   sigma = sum_{i > j2} col_i^2 = S_AA + 2 K c1 S_Av + (K c1)^2 S_vv      (sums over i > j2)
    // so the norm of the next reflector does not wait for a second reduction behind K.
    const double2 q1 = psb[j1];
    const double v1 = vs[j1];                       // 1 (0 in the virtual step)
    const double c1 = 1.0 + v1;
    double s_pv = 0.0, s_aa = 0.0, s_av = 0.0, s_vv = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool lv = (u < U - QU) || (u < nlive);
      double2 q = make_double2(0.0, 0.0);
      if (lv) q = psl[32 * u];
      wq[u] = q.x;     // stale for rows <= j, where v_j is zero
      double am = fma(-q1.x, vc[u], q.y - q.x);
      double vm = vc[u];
      if (u == 0 && lane <= lim) am = vm = 0.0;              // rows <= j2 do not enter the next reflector
      if (u == 1 && lane <= lim - 32) am = vm = 0.0;
      s_pv = fma(q.x, vc[u], s_pv);
      s_aa = fma(am, am, s_aa);
      s_av = fma(am, vm, s_av);
      s_vv = fma(vm, vm, s_vv);
      vo[u] = am;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_pv += __shfl_xor_sync(0xffffffffu, s_pv, o);
      s_aa += __shfl_xor_sync(0xffffffffu, s_aa, o);
      s_av += __shfl_xor_sync(0xffffffffu, s_av, o);
      s_vv += __shfl_xor_sync(0xffffffffu, s_vv, o);
    }
    const double K = 0.5 * tau_cur * s_pv;
    const double kc = K * c1;
    const double wj1 = fma(-K, v1, q1.x);
    const double dj1 = (q1.y - wj1) - wj1 * v1;
#pragma unroll
    for (int u = 0; u < U; ++u) wq[u] = fma(-K, vc[u], wq[u]);   // w_j (garbage only in rows where v stays zero)
    double tau_n = 0.0, beta;
    if (j1 <= k - 3) {
      const double2 q2 = psb[j2];
      const double v2 = vs[j2];
      const double w2 = fma(-K, v2, q2.x);
      const double alpha = (q2.y - w2) - wj1 * v2;
      const double sg = fma(kc, fma(kc, s_vv, 2.0 * s_av), s_aa);
      double inv = 0.0;
      beta = alpha;
      if (sg > 0.0) {
        const double x = fma(alpha, alpha, sg);
        const double rs = rsqrt(x);
        beta = -copysign(x * rs, alpha);
        tau_n = (beta - alpha) * -copysign(rs, alpha);     // (beta - alpha) / beta
        inv = trd_rcp(alpha - beta);
      }
      const double kci = kc * inv;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double vm = vc[u];
        if (u == 0 && lane <= lim) vm = 0.0;
        if (u == 1 && lane <= lim - 32) vm = 0.0;
        vo[u] = fma(kci, vm, vo[u] * inv);                 // (A_i + K c1 v_i) / (alpha - beta)
      }
      if (lane == lim) vo[0] = 1.0;
      if (U > 1 && lane == lim - 32) vo[1] = 1.0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool lv = (u < U - QU) || (u < nlive);
        if ((u % NW) == warp && lv) vsh[(par ^ 1) * KP + 32 * (tb + u) + lane] = vo[u];
      }
      if (c == (j1 % P) && warp == NW - 1) {
        double* row = gv + (int64_t)j1 * L.kp + 32 * tb + lane;
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (tb + u < kchunks) row[32 * u] = vo[u];
      }
    } else {   // j1 == k - 2: last off-diagonal element and d[k-1]; no reflector
      const int kl = k - 1;
      const double2 ql = psb[kl];
      const double vl = vs[kl];
      const double wl = fma(-K, vl, ql.x);
      beta = (ql.y - wl) - wj1 * vl;
      if (c == kl % P && warp == (kl / P) % NW && lane == 0) {
        gd[kl] = A[(size_t)(kl / P) * KP + kl] - 2.0 * vl * wl;   // pending update of step k - 3
        ge[kl] = 0.0;
      }
    }
    if (c == 0 && tid == 0) {
      gd[j1] = dj1;
      ge[j1] = beta;
      gtau[j1] = tau_n;
    }
#pragma unroll
    for (int m = 0; m < MC; ++m) {
      const int i = c + P * (warp + NW * m);
      const int ii = i < KP ? i : 0;
      const double vi = vs[ii];
      vcolp[m] = vi;
      wcolp[m] = fma(-K, vi, psb[ii].x);
    }
    tau_cur = tau_n;
    __syncwarp();   // the lanes that send in F_{j+1} release this warp's vsh stores with them
    if ((j2 & 31) == 0) {   // row j + 2 opens a new chunk: the frame moves up by one chunk
#pragma unroll
      for (int u = 0; u + 1 < NR; ++u) {
        va[u] = va[u + 1];
        vb[u] = vb[u + 1];
        wq[u] = wq[u + 1];
      }
      va[NR - 1] = vb[NR - 1] = wq[NR - 1] = 0.0;
    }
  };

  // one step with the smallest unrolled variant that covers the live chunks
  auto step = [&](const int j, double(&vc)[NR], double(&vo)[NR]) {
    const int nlive = NR - ((j + 1) >> 5);
    if (nlive > NR - QU) {
      if (j >= 0) fused(TrdInt<NR>{}, j, vc, vo);
      scalar(TrdInt<NR>{}, j, vc, vo);
    } else if (NR - QU >= 1 && nlive > NR - 2 * QU) {
      fused(TrdInt<(NR - QU >= 1 ? NR - QU : 1)>{}, j, vc, vo);
      scalar(TrdInt<(NR - QU >= 1 ? NR - QU : 1)>{}, j, vc, vo);
    } else if (NR - 2 * QU >= 1 && nlive > NR - 3 * QU) {
      fused(TrdInt<(NR - 2 * QU >= 1 ? NR - 2 * QU : 1)>{}, j, vc, vo);
      scalar(TrdInt<(NR - 2 * QU >= 1 ? NR - 2 * QU : 1)>{}, j, vc, vo);
    } else {
      fused(TrdInt<(NR - 3 * QU >= 1 ? NR - 3 * QU : 1)>{}, j, vc, vo);
      scalar(TrdInt<(NR - 3 * QU >= 1 ? NR - 3 * QU : 1)>{}, j, vc, vo);
    }
  };

  step(-1, vb, va);                      // virtual step: v_0 -> va
  for (int j = 0; j <= k - 3; j += 2) {
    step(j, va, vb);                     // v_{j+1} -> vb
    if (j + 1 <= k - 3) step(j + 1, vb, va);   // v_{j+2} -> va
  }
  trd_cluster_sync();   // no CTA leaves while stores into its shared memory may still be in flight
}

// ------------------------------------------------------------------------------------------------------
// B. eigenvalues: one CTA (4 warps) per eigenvalue, 128 shifts per pass (7 bits), 8 passes.  The fp64 pipe
// issues one warp instruction per 2 cycles per scheduler, so few warps per SM is what keeps the 3-instruction
// step of the Sturm recurrence at its 8-cycle DFMA latency.  The same launch carries one extra CTA per group
// of four reflectors: the six inner products the back-transformation needs (they are the same for every
// vector).
// ------------------------------------------------------------------------------------------------------
constexpr int kTrdEvalThreads = 128;
constexpr int kTrdEvalPasses = 8;

// Sturm count without the pivmin guard (device hot loop): zeros only arise for exactly decoupled blocks with a
// shift equal to a diagonal entry; the count is then off by one for that single shift.  Renormalised every
// eighth step (signs are all that matters), branch-free; |T| <= 1 bounds the growth by 3^8 in between.
__device__ __forceinline__ int trd_sturm_fast(const double2* __restrict__ de, int k, double x) {
  double pp = 1.0;
  double p = de[0].x - x;
  int cnt = (int)((unsigned)__double2hiint(p) >> 31);
  auto one = [&](int i) {
    const double2 q = de[i];                                  // (d_i, e_{i-1}^2): one 16-byte broadcast load
    const double pn = fma(q.x - x, p, -(q.y * pp));
    cnt += (int)((unsigned)(__double2hiint(pn) ^ __double2hiint(p)) >> 31);   // sign change (integer pipe)
    pp = p;
    p = pn;
  };
  int i = 1;
  for (; i + 8 <= k; i += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) one(i + q);
    const int ex = trd_exponent(p);
    const double sc = (ex > 64 || (ex < -64 && ex > -1023)) ? trd_pow2(-ex) : 1.0;
    p *= sc;
    pp *= sc;
  }
  for (; i < k; ++i) one(i);
  return cnt;
}

__global__ void __launch_bounds__(kTrdEvalThreads) trd_eigval_kernel(const tta_symeig_task* __restrict__ tasks) {
  extern __shared__ __align__(16) double ev_sm[];
  __shared__ double s_red[kTrdEvalThreads / 32];
  __shared__ int s_cnt[kTrdEvalThreads / 32];
  const tta_symeig_task tk = tasks[blockIdx.y];
  const int k = tk.k, r = tk.r;
  const int p = blockIdx.x;
  const TrdLayout L = trd_layout(k, r);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p >= r) {
    // ---- inner products of the reflectors of group g (rows j0, j0-1, j0-2, j0-3) ----
    const int g = p - r;
    const int j0 = k - 3 - 4 * g;
    if (j0 < 0) return;
    const double* gv = tk.work + L.v;
    double* out = tk.work + L.cross + 8 * (int64_t)g;
    // pair list: (b,a) (c,a) (c,b) (d,a) (d,b) (d,c) -> rows (j0 - x, j0 - y)
    for (int q = warp; q < 6; q += kTrdEvalThreads / 32) {
      const int xr = q == 0 ? 1 : (q < 3 ? 2 : 3);
      const int yr = q == 0 ? 0 : (q == 1 ? 0 : (q == 2 ? 1 : q - 3));
      double acc = 0.0;
      if (j0 - xr >= 0) {
        const double* rx = gv + (int64_t)(j0 - xr) * L.kp;
        const double* ry = gv + (int64_t)(j0 - yr) * L.kp;
        for (int idx = (j0 & ~31) + lane; idx < L.kp; idx += 32) acc = fma(rx[idx], ry[idx], acc);   // both zero below j0
      }
      acc = warp_sum(acc);
      if (lane == 0) out[q] = acc;
    }
    return;
  }
  double2* de = reinterpret_cast<double2*>(ev_sm);
  const double* gd = tk.work + L.d;
  const double* ge = tk.work + L.e;

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
    const double e = i > 0 ? ge[i - 1] * sc : 0.0;
    de[i] = make_double2(gd[i] * sc, e * e);
  }
  __syncthreads();

  const int m = k - 1 - p;            // ascending index of the p-th largest eigenvalue
  double lo = -1.0 - 1e-9, hi = 1.0 + 1e-9;
  for (int pass = 0; pass < kTrdEvalPasses; ++pass) {
    const double step = (hi - lo) * (1.0 / (kTrdEvalThreads + 1));
    const double x = fma(step, (double)(tid + 1), lo);
    const int cnt = trd_sturm_fast(de, k, x);
    const unsigned b = __ballot_sync(0xffffffffu, cnt <= m);   // x <= lambda_m
    if (lane == 0) s_cnt[warp] = __popc(b);
    __syncthreads();
    int nle = 0;
#pragma unroll
    for (int q = 0; q < kTrdEvalThreads / 32; ++q) nle += s_cnt[q];
    __syncthreads();
    const double nlo = fma(step, (double)nle, lo);
    const double nhi = nle < kTrdEvalThreads ? fma(step, (double)(nle + 1), lo) : hi;
    lo = nlo;
    hi = nhi;
  }
  if (tid == 0) {
    const double lam = 0.5 * (lo + hi);
    tk.work[L.lams + p] = lam;
    tk.lam[p] = lam * trd_pow2(ex);
    if (p == 0) tk.work[L.hdr] = sc;
  }
}

// ------------------------------------------------------------------------------------------------------
// C. eigenvectors of T: twisted factorisation, two threads (forward / backward sweep) per vector, one warp
// per CTA (16 vectors): the sweeps are chains of k dependent reciprocals, and a lone warp keeps them at the
// latency of the fp64 pipe.  The pivots stay in shared memory (element i of vector v at [i][v]).
// ------------------------------------------------------------------------------------------------------
constexpr int kTrdEvecThreads = 32;
constexpr int kTrdEvecPer = kTrdEvecThreads / 2;
constexpr double kTrdClusterGap = 1e-9;    // scaled units: closer eigenvalue pairs are reported (status 1)
constexpr double kTrdLiveRel = 4e-7;       // same cut as refine.cu: smaller eigenvalues count as zero

__global__ void __launch_bounds__(kTrdEvecThreads) trd_eigvec_kernel(const tta_symeig_task* __restrict__ tasks) {
  extern __shared__ __align__(16) double evc_sm[];
  const tta_symeig_task tk = tasks[blockIdx.y];
  const int k = tk.k, r = tk.r;
  if ((int)blockIdx.x * kTrdEvecPer >= r) return;
  const TrdLayout L = trd_layout(k, r);
  double* dd = evc_sm;
  double* ee = evc_sm + L.kp;
  double* e2 = evc_sm + 2 * L.kp;
  double* Ssm = evc_sm + 3 * L.kp;                       // [k][16] forward pivots
  double* Psm = Ssm + (size_t)k * kTrdEvecPer;           // [k][16] backward pivots
  const double sc = tk.work[L.hdr];
  const int tid = threadIdx.x;
  for (int i = tid; i < k; i += kTrdEvecThreads) {
    dd[i] = tk.work[L.d + i] * sc;
    const double e = i < k - 1 ? tk.work[L.e + i] * sc : 0.0;
    ee[i] = e;
    e2[i] = e * e;
  }
  __syncwarp();
  const int vloc = tid >> 1;
  const int p = blockIdx.x * kTrdEvecPer + vloc;
  const int dir = tid & 1;
  const bool live = p < r;
  const double lam = tk.work[L.lams + (live ? p : r - 1)];
  double* S = Ssm + vloc;
  double* Pv = Psm + vloc;
  if (dir == 0) {
    double q = trd_guard(dd[0] - lam);
    S[0] = q;
    for (int i = 0; i < k - 1; ++i) {
      q = trd_guard(fma(-e2[i], trd_rcp(q), dd[i + 1] - lam));
      S[(i + 1) * kTrdEvecPer] = q;
    }
  } else {
    double q = trd_guard(dd[k - 1] - lam);
    Pv[(k - 1) * kTrdEvecPer] = q;
    for (int i = k - 2; i >= 0; --i) {
      q = trd_guard(fma(-e2[i], trd_rcp(q), dd[i] - lam));
      Pv[i * kTrdEvecPer] = q;
    }
  }
  __syncwarp();
  // twist index: smallest |gamma_i|, gamma_i = s_i + p_i - (d_i - lambda)
  const int half = k >> 1;
  const int ibeg = dir == 0 ? 0 : half, iend = dir == 0 ? half : k;
  double best = 1e300;
  int bi = ibeg;
#pragma unroll 4
  for (int i = ibeg; i < iend; ++i) {
    const double gm = fabs(S[i * kTrdEvecPer] + Pv[i * kTrdEvecPer] - (dd[i] - lam));
    if (gm < best) {
      best = gm;
      bi = i;
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
#pragma unroll 4
      for (int i = tw - 1; i >= 0; --i) {
        z = -(ee[i] * trd_rcp(S[i * kTrdEvecPer])) * z;
        out[i] = z;
      }
      if (p + 1 < r) {
        const double l0 = tk.work[L.lams], ln = tk.work[L.lams + p + 1];
        if (lam > kTrdLiveRel * l0 && lam - ln < kTrdClusterGap) atomicOr(tk.status, 1);
      }
    } else {
#pragma unroll 4
      for (int i = tw; i < k - 1; ++i) {
        z = -(ee[i] * trd_rcp(Pv[(i + 1) * kTrdEvecPer])) * z;
        out[i + 1] = z;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// D. back-transformation x = H_0 ... H_{k-3} z, in place on e64.  One warp per vector (x in registers,
// lane-strided), four reflectors per reduction round: with a, b, c, d = v_j, v_{j-1}, v_{j-2}, v_{j-3}
//     fa = ta (a.x);  fb = tb (b.x - fa b.a);  fc = tc (c.x - fa c.a - fb c.b);  fd = td (d.x - fa d.a - fb d.b - fc d.c)
//     x <- x - fa a - fb b - fc c - fd d
// (the inner products of the reflectors come precomputed from the eigenvalue launch).  The four rows are staged
// through shared memory with cp.async, two groups ahead, shared by the warps of the CTA; the staging buffers
// start zeroed and only the non-zero chunks of a row are ever copied, so no element needs a predicate.
// ------------------------------------------------------------------------------------------------------
constexpr int kTrdBtWarps = 4;
constexpr int kTrdBtStages = 3;

template <int NR>
__global__ void __launch_bounds__(32 * kTrdBtWarps) trd_backtransform_kernel(const tta_symeig_task* __restrict__ tasks,
                                                                            int first) {
  extern __shared__ __align__(16) double bt_sm[];     // [stages][4 rows of KP | 4 taus | 6 inner products, pad]
  constexpr int KP = NR * 32;
  constexpr int kStage = 4 * KP + 16;
  const tta_symeig_task tk = tasks[first + blockIdx.y];
  const int k = tk.k, r = tk.r;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((int)blockIdx.x * kTrdBtWarps >= r) return;
  const int p = blockIdx.x * kTrdBtWarps + warp;
  const bool live = p < r;
  const TrdLayout L = trd_layout(k, r);
  const double* gv = tk.work + L.v;
  const double* gtau = tk.work + L.tau;
  const double* gcross = tk.work + L.cross;
  for (int e = tid; e < kTrdBtStages * kStage; e += 32 * kTrdBtWarps) bt_sm[e] = 0.0;
  double x[NR];
#pragma unroll
  for (int t = 0; t < NR; ++t) {
    const int idx = lane + 32 * t;
    x[t] = (live && idx < k) ? tk.e64[(int64_t)p * k + idx] : 0.0;
  }
  __syncthreads();
  const int nref = k - 2;                       // reflectors 0 .. k-3
  const int ngroups = (nref + 3) / 4;           // group g holds rows j0, j0-1, j0-2, j0-3 with j0 = k-3-4g
  // stage loader: rows below 0 do not exist (their slots keep the zeros / stale rows, tau = 0 for them)
  auto load = [&](int g) {
    if (g < ngroups) {
      double* dst = bt_sm + (size_t)(g % kTrdBtStages) * kStage;
      const int j0 = k - 3 - 4 * g;
      const int nch = L.kp / 2;
      if (tid < 4) {           // tau of rows j0 - tid (slots of rows below 0 keep an older, harmless tau: their
        if (j0 - tid >= 0) {   // rows are stale as well, and the scalar below is forced to zero)
          const uint32_t d = trd_smem_u32(dst + 4 * KP + tid);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gtau + j0 - tid) : "memory");
        }
      } else if (tid < 7) {
        const uint32_t d = trd_smem_u32(dst + 4 * KP + 4 + 2 * (tid - 4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gcross + 8 * g + 2 * (tid - 4)) : "memory");
      }
      for (int q = 0; q < 4; ++q) {
        const int row = j0 - q;
        if (row < 0) break;
        const double* src = gv + (int64_t)row * L.kp;
        const int c0 = (row >> 5) * 16;   // the reduction wrote row `row` from the 32-row chunk of index `row` on
        for (int ch = c0 + tid; ch < nch; ch += 32 * kTrdBtWarps) {
          const uint32_t d = trd_smem_u32(dst + q * KP + 2 * ch);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 2 * ch) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load(0);
  load(1);
  for (int g = 0; g < ngroups; ++g) {
    load(g + 2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncthreads();
    const double* sb = bt_sm + (size_t)(g % kTrdBtStages) * kStage;
    const double* st = sb + lane;
    const int j0 = k - 3 - 4 * g;
    const double ta = sb[4 * KP], tb = j0 >= 1 ? sb[4 * KP + 1] : 0.0, tc = j0 >= 2 ? sb[4 * KP + 2] : 0.0,
                 td = j0 >= 3 ? sb[4 * KP + 3] : 0.0;
    const double2 c01 = *reinterpret_cast<const double2*>(sb + 4 * KP + 4);
    const double2 c23 = *reinterpret_cast<const double2*>(sb + 4 * KP + 6);
    const double2 c45 = *reinterpret_cast<const double2*>(sb + 4 * KP + 8);
    double a[NR], b[NR], cc_[NR], d[NR];
    double ax = 0, bx = 0, cx = 0, dx = 0;
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      a[t] = st[32 * t];
      b[t] = st[KP + 32 * t];
      cc_[t] = st[2 * KP + 32 * t];
      d[t] = st[3 * KP + 32 * t];
      ax = fma(a[t], x[t], ax);
      bx = fma(b[t], x[t], bx);
      cx = fma(cc_[t], x[t], cx);
      dx = fma(d[t], x[t], dx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ax += __shfl_xor_sync(0xffffffffu, ax, o);
      bx += __shfl_xor_sync(0xffffffffu, bx, o);
      cx += __shfl_xor_sync(0xffffffffu, cx, o);
      dx += __shfl_xor_sync(0xffffffffu, dx, o);
    }
    const double fa = ta * ax;
    const double fb = tb * fma(-fa, c01.x, bx);
    const double fc = tc * fma(-fb, c23.x, fma(-fa, c01.y, cx));
    const double fd = td * fma(-fc, c45.y, fma(-fb, c45.x, fma(-fa, c23.y, dx)));
#pragma unroll
    for (int t = 0; t < NR; ++t) x[t] = fma(-fd, d[t], fma(-fc, cc_[t], fma(-fb, b[t], fma(-fa, a[t], x[t]))));
    __syncthreads();   // the stage is free for the load of group g + 3
  }
  if (live) {
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      const int idx = lane + 32 * t;
      if (idx < k) tk.e64[(int64_t)p * k + idx] = x[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static size_t trd_reduce_smem(int k, int P, int nr) {
  const int nclmax = (k + P - 1) / P;
  // columns + ps[2][KP] (16 B) + vsh[2][KP] + two mbarriers
  return ((size_t)nclmax * nr * 32 + 6 * (size_t)nr * 32 + 2) * sizeof(double);
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

// bench.py profiling aid: when enabled every trd_reduce launch is bracketed by CUDA events on the stream it runs on
static bool g_trd_prof = false;
struct TrdProfRec { cudaEvent_t a, b; };
static std::vector<TrdProfRec> g_trd_recs;
// mode 2 (scripts/trace_symeig_stages.py): five events per tta_symeig_top_batched call on the caller's stream --
// before the reduction, after it, after the eigenvalues, after the vectors, after the back-transformation
static int g_trd_stage_prof = 0;
struct TrdStageRec { cudaEvent_t e[5]; int kmax, n_tasks; };
static std::vector<TrdStageRec> g_trd_stage_recs;

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

template <int NR>
static int trd_launch_backtransform(const tta_symeig_task* tasks_dev, int first, int count, int rmax, cudaStream_t st) {
  const size_t smem = (size_t)kTrdBtStages * (4 * NR * 32 + 16) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    int rc = check_cuda(cudaFuncSetAttribute(trd_backtransform_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem), "symeig back-transform smem attribute");
    if (rc) return rc;
    attr_set = true;
  }
  const dim3 grid((unsigned)((rmax + kTrdBtWarps - 1) / kTrdBtWarps), (unsigned)count);
  trd_backtransform_kernel<NR><<<grid, 32 * kTrdBtWarps, smem, st>>>(tasks_dev, first);
  TTA_CHECK_LAUNCH("symeig back-transform launch");
  return TTA_OK;
}

static int trd_backtransform_dispatch(int nr, const tta_symeig_task* tasks_dev, int first, int count, int rmax,
                                      cudaStream_t st) {
  switch (nr) {
    case 2: return trd_launch_backtransform<2>(tasks_dev, first, count, rmax, st);
    case 4: return trd_launch_backtransform<4>(tasks_dev, first, count, rmax, st);
    case 8: return trd_launch_backtransform<8>(tasks_dev, first, count, rmax, st);
    case 12: return trd_launch_backtransform<12>(tasks_dev, first, count, rmax, st);
    case 16: return trd_launch_backtransform<16>(tasks_dev, first, count, rmax, st);
    default: return trd_launch_backtransform<20>(tasks_dev, first, count, rmax, st);
  }
}

}  // namespace tta

extern "C" {

size_t tta_symeig_work_doubles(int k, int r) {
  if (k <= 0 || r <= 0) return 0;
  return (size_t)tta::trd_layout(k, r).total;
}

int tta_symeig_max_k(void) { return tta::kTrdMaxK; }

void tta_symeig_profile_enable(int on) {
  tta::g_trd_prof = on == 1;
  tta::g_trd_stage_prof = on == 2 ? 1 : 0;
}

/* mode 2 readback: per call (in call order) kmax, n_tasks, and the milliseconds of the four stages relative to a common
 * origin event `origin` recorded by the caller (cudaEvent_t handle) -- start of reduce, end of reduce, end of eigenvalues,
 * end of vectors, end of back-transformation.  Returns the number of records written (7 doubles each). */
int tta_symeig_stage_profile_read(void* origin, double* out, int max_records) {
  using namespace tta;
  int n = 0;
  for (TrdStageRec& r : g_trd_stage_recs) {
    if (n < max_records) {
      out[7 * n] = r.kmax;
      out[7 * n + 1] = r.n_tasks;
      for (int i = 0; i < 5; ++i) {
        float t = 0.f;
        cudaEventSynchronize(r.e[i]);
        cudaEventElapsedTime(&t, (cudaEvent_t)origin, r.e[i]);
        out[7 * n + 2 + i] = t;
      }
      ++n;
    }
    for (int i = 0; i < 5; ++i) cudaEventDestroy(r.e[i]);
  }
  g_trd_stage_recs.clear();
  return n;
}

void tta_symeig_profile_read(double* reduce_ms, unsigned long long* reduce_launches) {
  using namespace tta;
  double ms = 0.0;
  unsigned long long n = 0;
  for (TrdProfRec& r : g_trd_recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms += t;
      ++n;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_trd_recs.clear();
  if (reduce_ms) *reduce_ms = ms;
  if (reduce_launches) *reduce_launches = n;
}

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
  TrdStageRec srec;
  const bool stage_prof = g_trd_stage_prof != 0;
  auto stage_mark = [&](int i) {
    if (stage_prof && cudaEventCreate(&srec.e[i]) == cudaSuccess) cudaEventRecord(srec.e[i], st);
  };
  srec.kmax = kmax;
  srec.n_tasks = n_tasks;
  stage_mark(0);
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
    TrdProfRec rec = {nullptr, nullptr};
    if (g_trd_prof && cudaEventCreate(&rec.a) == cudaSuccess && cudaEventCreate(&rec.b) == cudaSuccess)
      cudaEventRecord(rec.a, gs);
    rc = trd_reduce_dispatch(rn.nr, tasks_dev, rn.first, rn.count, rn.P, trd_reduce_smem(rn.kmax, rn.P, rn.nr), gs);
    if (rc) return rc;
    if (rec.b) {
      cudaEventRecord(rec.b, gs);
      g_trd_recs.push_back(rec);
    }
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
  stage_mark(1);
  const int kpmax = (kmax + 31) & ~31;
  const size_t smem = (size_t)2 * kpmax * sizeof(double);
  // grid.x = eigenvalues of the widest task + one CTA per group of four reflectors of the largest task
  trd_eigval_kernel<<<dim3((unsigned)(rmax + (kmax + 1) / 4), (unsigned)n_tasks), kTrdEvalThreads, smem, st>>>(tasks_dev);
  TTA_CHECK_LAUNCH("symeig eigenvalue launch");
  stage_mark(2);
  const size_t smem_vec = ((size_t)3 * kpmax + (size_t)2 * kmax * kTrdEvecPer) * sizeof(double);
  static size_t smem_vec_set = 48 * 1024;
  if (smem_vec > smem_vec_set) {
    rc = check_cuda(cudaFuncSetAttribute(trd_eigvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_vec),
                    "symeig eigenvector smem attribute");
    if (rc) return rc;
    smem_vec_set = smem_vec;
  }
  trd_eigvec_kernel<<<dim3((unsigned)((rmax + kTrdEvecPer - 1) / kTrdEvecPer), (unsigned)n_tasks), kTrdEvecThreads,
                      smem_vec, st>>>(tasks_dev);
  TTA_CHECK_LAUNCH("symeig eigenvector launch");
  stage_mark(3);
  for (const Run& rn : runs) {
    rc = trd_backtransform_dispatch(rn.nr, tasks_dev, rn.first, rn.count, rn.rmax, st);
    if (rc) return rc;
  }
  stage_mark(4);
  if (stage_prof) g_trd_stage_recs.push_back(srec);
  return TTA_OK;
}
}
