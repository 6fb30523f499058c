// Device-side building blocks of the one-sided Jacobi solvers (shared by eig.cu and eig_cluster.cu).
#pragma once
#include <vector>

#include "tta_common.cuh"

namespace tta {

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

// Rotation parameters from the three inner products a = x.x, b = y.y, c = x.y.
// Fast-math intrinsics on purpose: whatever t comes out, (sn, tau) derived from that one t define an
// exact plane rotation up to rounding, so approximate division / rsqrt only perturbs the annihilation
// angle by O(ulp) (a slightly slower convergence), never the orthogonality of the transform.  The
// serial latency of this scalar chain is what bounds a Jacobi step, hence no IEEE div / sqrt here.
__device__ __forceinline__ float mufu_rcp(float x) {   // single MUFU.RCP, no range fix-up code
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_rsqrt(float x) {  // single MUFU.RSQ
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ bool jacobi_params(float a, float b, float c, float tol2, float floor2, float& sn,
                                              float& tau) {
  if (!(a > floor2) || !(b > floor2)) return false;
  if (!(c * c > (tol2 * a) * b)) return false;
  const float zeta = (b - a) * mufu_rcp(2.f * c);
  const float h = fmaf(zeta, zeta, 1.f);
  // |zeta| < ~1e15 here (c passed the threshold test), so h is finite and rsqrt is safe
  float t = mufu_rcp(fabsf(zeta) + h * mufu_rsqrt(h));
  t = copysignf(t, zeta);
  const float cs = mufu_rsqrt(fmaf(t, t, 1.f));
  sn = cs * t;
  tau = sn * mufu_rcp(1.f + cs);
  return true;
}

// x' = x - sn*(y + tau*x), y' = y + sn*(x - tau*y) with tau = sn/(1+cs)  (== cs*x - sn*y, sn*x + cs*y).
// Late rotations have cs == 1.0f after rounding; applying the 1-cs part explicitly keeps every
// rotation norm-preserving to rounding error instead of inflating the columns by t^2/2 each time.
__device__ __forceinline__ void rot2(float sn, float tau, float x, float y, float& xn, float& yn) {
  xn = fmaf(-sn, fmaf(tau, x, y), x);
  yn = fmaf(sn, fmaf(-tau, y, x), y);
}

template <int NV>
__device__ __forceinline__ void load_col(const float* __restrict__ col, int ld, int lane, float4 (&v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int e = (j * 32 + lane) * 4;
    v[j] = e < ld ? *reinterpret_cast<const float4*>(col + e) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NV>
__device__ __forceinline__ void store_col(float* __restrict__ col, int ld, int lane, const float4 (&v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int e = (j * 32 + lane) * 4;
    if (e < ld) *reinterpret_cast<float4*>(col + e) = v[j];
  }
}

// One column pair held in registers (NV float4 per lane per column).  Returns 1 if rotated.
// (No de Rijk column swap: with the parallel tournament orderings used here, moving the larger
// column to `x` makes pairs chase each other across blocks and the sweep count explodes.)
template <int NV>
__device__ __forceinline__ int rotate_regs(float4 (&x)[NV], float4 (&y)[NV], float tol2, float floor2) {
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    a0 = fmaf(x[j].x, x[j].x, a0); a1 = fmaf(x[j].y, x[j].y, a1);
    b0 = fmaf(y[j].x, y[j].x, b0); b1 = fmaf(y[j].y, y[j].y, b1);
    c0 = fmaf(x[j].x, y[j].x, c0); c1 = fmaf(x[j].y, y[j].y, c1);
    a0 = fmaf(x[j].z, x[j].z, a0); a1 = fmaf(x[j].w, x[j].w, a1);
    b0 = fmaf(y[j].z, y[j].z, b0); b1 = fmaf(y[j].w, y[j].w, b1);
    c0 = fmaf(x[j].z, y[j].z, c0); c1 = fmaf(x[j].w, y[j].w, c1);
  }
  float a = a0 + a1, b = b0 + b1, c = c0 + c1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {   // three interleaved butterflies
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  float sn, tau;
  if (!jacobi_params(a, b, c, tol2, floor2, sn, tau)) return 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float4 xn, yn;
    rot2(sn, tau, x[j].x, y[j].x, xn.x, yn.x);
    rot2(sn, tau, x[j].y, y[j].y, xn.y, yn.y);
    rot2(sn, tau, x[j].z, y[j].z, xn.z, yn.z);
    rot2(sn, tau, x[j].w, y[j].w, xn.w, yn.w);
    x[j] = xn;
    y[j] = yn;
  }
  return 1;
}

// Generic fallback for long columns (ld > 512): two passes over shared memory.
__device__ __forceinline__ int rotate_smem(float* __restrict__ x, float* __restrict__ y, int ld, int lane, float tol2,
                                           float floor2) {
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
  float sn, tau;
  if (!jacobi_params(a, b, c, tol2, floor2, sn, tau)) return 0;
  for (int e = lane * 4; e < ld; e += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(x + e);
    const float4 yv = *reinterpret_cast<const float4*>(y + e);
    float4 xn, yn;
    rot2(sn, tau, xv.x, yv.x, xn.x, yn.x);
    rot2(sn, tau, xv.y, yv.y, xn.y, yn.y);
    rot2(sn, tau, xv.z, yv.z, xn.z, yn.z);
    rot2(sn, tau, xv.w, yv.w, xn.w, yn.w);
    *reinterpret_cast<float4*>(x + e) = xn;
    *reinterpret_cast<float4*>(y + e) = yn;
  }
  return 1;
}

// All pair rotations of one staged block pair.  NV > 0: register-resident columns (ld <= 128*NV).
template <int NV>
__device__ __forceinline__ int jacobi_block(float* __restrict__ cols, int kind, int nblk, int bw, int ld, int warp,
                                            int lane, float tol2, float fl) {
  constexpr int N = NV > 0 ? NV : 1;
  int nrot = 0;
  if (kind == 1) {
    // cross pairs: warp w keeps column w of block A in registers for the whole block step and meets
    // column (w+s)%bw of block B at step s (disjoint pairs within a step).
    float4 x[N];
    float* xcol = cols + warp * ld;
    if (NV > 0 && warp < bw) load_col<N>(xcol, ld, lane, x);
    int yslot = warp;                       // (warp + s) % bw without the integer division
    for (int s = 0; s < bw; ++s, ++yslot) {
      if (warp < bw) {
        if (yslot >= bw) yslot -= bw;
        float* ycol = cols + (bw + yslot) * ld;
        if (NV > 0) {
          float4 y[N];
          load_col<N>(ycol, ld, lane, y);
          if (rotate_regs<N>(x, y, tol2, fl)) {
            store_col<N>(ycol, ld, lane, y);
            ++nrot;
          }
        } else {
          nrot += rotate_smem(xcol, ycol, ld, lane, tol2, fl);
        }
      }
      __syncthreads();
    }
    if (NV > 0 && warp < bw && nrot) store_col<N>(xcol, ld, lane, x);
    __syncthreads();   // the block in shared memory is complete for every reader after this point
  } else {
    const int half = bw >> 1;
    for (int r = 0; r < bw - 1; ++r) {
      if (warp < half * nblk) {
        const int h = warp / half, q = warp - h * half;
        int p0, p1;
        rr_pair(bw, r, q, p0, p1);
        float* xcol = cols + (h * bw + (p0 < p1 ? p0 : p1)) * ld;
        float* ycol = cols + (h * bw + (p0 < p1 ? p1 : p0)) * ld;
        if (NV > 0) {
          float4 x[N], y[N];
          load_col<N>(xcol, ld, lane, x);
          load_col<N>(ycol, ld, lane, y);
          if (rotate_regs<N>(x, y, tol2, fl)) {
            store_col<N>(xcol, ld, lane, x);
            store_col<N>(ycol, ld, lane, y);
            ++nrot;
          }
        } else {
          nrot += rotate_smem(xcol, ycol, ld, lane, tol2, fl);
        }
      }
      __syncthreads();
    }
  }
  return nrot;
}


// host-side entry of the cluster solver (eig_cluster.cu)
bool jacobi_cluster_eligible(const tta_eig_task& tk);
int jacobi_cluster_run(const tta_eig_task* tasks_dev, const tta_eig_task* th, const std::vector<int>& probs,
                       float tol2, int max_sweeps, int32_t* ids_dev, int32_t* sweeps_dev, int32_t* status_dev,
                       const float* floor2, cudaStream_t st);

}  // namespace tta
