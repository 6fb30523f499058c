// Device-side building blocks of the one-sided Jacobi solvers (shared by eig.cu and eig_cluster.cu).
#pragma once
#include <vector>

#include "tta_common.cuh"

namespace tta {

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

// t = tan(theta) is returned too: the squared norms after the rotation are a - t*c and b + t*c.
__device__ __forceinline__ bool jacobi_params(float a, float b, float c, float tol2, float floor2, float& sn,
                                              float& tau, float& t) {
  if (!(a > floor2) || !(b > floor2)) return false;
  if (!(c * c > (tol2 * a) * b)) return false;
  const float zeta = (b - a) * mufu_rcp(2.f * c);
  const float h = fmaf(zeta, zeta, 1.f);
  // |zeta| < ~1e15 here (c passed the threshold test), so h is finite and rsqrt is safe
  t = copysignf(mufu_rcp(fabsf(zeta) + h * mufu_rsqrt(h)), zeta);
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

// Blackwell packed fp32 FMA (FFMA2): two lanes of d = a*b + c per instruction.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
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

// warp-wide inner product of two register-resident columns (every lane gets the result)
template <int NV>
__device__ __forceinline__ float dot_regs(const float4 (&x)[NV], const float4 (&y)[NV]) {
  float2 p = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    p = ffma2(make_float2(x[j].x, x[j].y), make_float2(y[j].x, y[j].y), p);
    q = ffma2(make_float2(x[j].z, x[j].w), make_float2(y[j].z, y[j].w), q);
  }
  float c = (p.x + p.y) + (q.x + q.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  return c;
}

// One column pair held in registers (NV float4 per lane per column) with cached squared norms a, b
// (recomputed exactly at the start of every block round, then updated as a - t*c / b + t*c): only the
// cross product needs a reduction in the step's critical path.  Returns 1 if rotated.
// (No de Rijk column swap: with the parallel tournament orderings used here, moving the larger
// column to `x` makes pairs chase each other across blocks and the sweep count explodes.)
template <int NV>
__device__ __forceinline__ int rotate_regs(float4 (&x)[NV], float4 (&y)[NV], float& a, float& b, float tol2,
                                           float floor2, float stop2 = 0.f) {
  const float c = dot_regs<NV>(x, y);
  float sn, tau, t;
  if (!jacobi_params(a, b, c, tol2, floor2, sn, tau, t)) return 0;
  // 1 per rotation; + 65536 when the pair was further from orthogonal than the stop threshold (see cluster_body)
  const int ret = 1 + ((c * c > (stop2 * a) * b) ? 65536 : 0);
  const float2 sp = make_float2(sn, sn), sm = make_float2(-sn, -sn);
  const float2 tp = make_float2(tau, tau), tm = make_float2(-tau, -tau);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float2 x0 = make_float2(x[j].x, x[j].y), x1 = make_float2(x[j].z, x[j].w);
    const float2 y0 = make_float2(y[j].x, y[j].y), y1 = make_float2(y[j].z, y[j].w);
    const float2 xn0 = ffma2(sm, ffma2(tp, x0, y0), x0), xn1 = ffma2(sm, ffma2(tp, x1, y1), x1);
    const float2 yn0 = ffma2(sp, ffma2(tm, y0, x0), y0), yn1 = ffma2(sp, ffma2(tm, y1, x1), y1);
    x[j] = make_float4(xn0.x, xn0.y, xn1.x, xn1.y);
    y[j] = make_float4(yn0.x, yn0.y, yn1.x, yn1.y);
  }
  const float d = t * c;
  a = fmaxf(a - d, 0.f);
  b = fmaxf(b + d, 0.f);
  return ret;
}

// Generic fallback for long columns (ld > 512): two passes over shared memory.
__device__ __forceinline__ int rotate_smem(float* __restrict__ x, float* __restrict__ y, int ld, int lane, float tol2,
                                           float floor2, float stop2 = 0.f) {
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
  float sn, tau, t;
  if (!jacobi_params(a, b, c, tol2, floor2, sn, tau, t)) return 0;
  const int ret = 1 + ((c * c > (stop2 * a) * b) ? 65536 : 0);
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
  return ret;
}

// All pair rotations of one staged block pair.  NV > 0: register-resident columns (ld <= 128*NV) with
// cached squared norms in `nrm` (2*bw floats of shared memory behind the columns).
template <int NV>
__device__ __forceinline__ int jacobi_block(float* __restrict__ cols, float* __restrict__ nrm, int kind, int nblk,
                                            int bw, int ld, int warp, int lane, float tol2, float fl,
                                            float stop2 = 0.f) {
  constexpr int N = NV > 0 ? NV : 1;
  int nrot = 0;
  if (kind == 1) {
    // cross pairs: warp w keeps column w of block A in registers for the whole block step and meets
    // column (w+s)%bw of block B at step s (disjoint pairs within a step).
    float4 x[N];
    float xa = 0.f;
    float* xcol = cols + warp * ld;
    if (NV > 0 && warp < bw) {
      load_col<N>(xcol, ld, lane, x);
      xa = dot_regs<N>(x, x);
      float4 y[N];
      load_col<N>(cols + (bw + warp) * ld, ld, lane, y);
      const float yb = dot_regs<N>(y, y);
      if (lane == 0) nrm[bw + warp] = yb;
    }
    if (NV > 0) __syncthreads();
    int yslot = warp;                       // (warp + s) % bw without the integer division
    for (int s = 0; s < bw; ++s, ++yslot) {
      if (warp < bw) {
        if (yslot >= bw) yslot -= bw;
        float* ycol = cols + (bw + yslot) * ld;
        if (NV > 0) {
          float4 y[N];
          load_col<N>(ycol, ld, lane, y);
          float yb = nrm[bw + yslot];
          if (const int rr = rotate_regs<N>(x, y, xa, yb, tol2, fl, stop2)) {
            store_col<N>(ycol, ld, lane, y);
            if (lane == 0) nrm[bw + yslot] = yb;
            nrot += rr;
          }
        } else {
          nrot += rotate_smem(xcol, ycol, ld, lane, tol2, fl, stop2);
        }
      }
      __syncthreads();
    }
    if (NV > 0 && warp < bw && nrot) store_col<N>(xcol, ld, lane, x);
    __syncthreads();   // the block in shared memory is complete for every reader after this point
  } else {
    const int half = bw >> 1;
    if (NV > 0) {
      for (int col = warp; col < nblk * bw; col += (int)(blockDim.x >> 5)) {
        float4 v[N];
        load_col<N>(cols + col * ld, ld, lane, v);
        const float vv = dot_regs<N>(v, v);
        if (lane == 0) nrm[col] = vv;
      }
      __syncthreads();
    }
    for (int r = 0; r < bw - 1; ++r) {
      if (warp < half * nblk) {
        const int h = warp / half, q = warp - h * half;
        int p0, p1;
        rr_pair(bw, r, q, p0, p1);
        const int xi = h * bw + (p0 < p1 ? p0 : p1), yi = h * bw + (p0 < p1 ? p1 : p0);
        float* xcol = cols + xi * ld;
        float* ycol = cols + yi * ld;
        if (NV > 0) {
          float4 x[N], y[N];
          load_col<N>(xcol, ld, lane, x);
          load_col<N>(ycol, ld, lane, y);
          float xa = nrm[xi], yb = nrm[yi];
          if (const int rr = rotate_regs<N>(x, y, xa, yb, tol2, fl, stop2)) {
            store_col<N>(xcol, ld, lane, x);
            store_col<N>(ycol, ld, lane, y);
            if (lane == 0) {
              nrm[xi] = xa;
              nrm[yi] = yb;
            }
            nrot += rr;
          }
        } else {
          nrot += rotate_smem(xcol, ycol, ld, lane, tol2, fl, stop2);
        }
      }
      __syncthreads();
    }
  }
  return nrot;
}

// host-side entries of the persistent cluster solvers (eig_cluster.cu, eig_gra.cu)
bool jacobi_cluster_eligible(const tta_eig_task& tk);
bool jacobi_gra_eligible(const tta_eig_task& tk);
int jacobi_gra_enqueue(const tta_eig_task* tasks_dev, const tta_eig_task* th, const std::vector<int>& probs, int P,
                       float tol2, float stop2, int max_sweeps, const int32_t* ids_dev, int32_t* sweeps_dev,
                       int32_t* status_dev, const float* floor2, cudaStream_t gs);
int jacobi_cluster_run(const tta_eig_task* tasks_dev, const tta_eig_task* th, const std::vector<int>& probs,
                       float tol2, float stop2, int max_sweeps, bool allow_gra, int32_t* ids_dev, int32_t* sweeps_dev,
                       int32_t* status_dev, const float* floor2, cudaStream_t st);

}  // namespace tta
