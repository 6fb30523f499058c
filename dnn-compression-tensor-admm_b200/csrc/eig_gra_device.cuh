// Device-side building blocks of the gram-rotate-apply Jacobi kernel (eig_gra.cu); also included by the
// micro-benchmarks under scripts/ubench/.
#pragma once
#include "eig_device.cuh"

namespace tta {

constexpr int kGraThreads = 512;
constexpr int kGraCs = 48;                 // row stride of C in shared memory (rows r, r+1 in disjoint bank halves)
constexpr int kGraCsz = 32 * kGraCs;
constexpr int kGraVs = 40;                 // row stride of V (rows r, r+2 in disjoint bank halves; 16-byte aligned)

// padded column stride: lds/4 odd, so that 8 columns' float4 at one row offset hit 8 distinct
// 16-byte bank groups
__host__ __device__ inline int gra_lds(int ld) { return (((ld >> 2) + 1) | 1) << 2; }
__host__ __device__ inline int gra_buf_floats(int ld) { return 32 * gra_lds(ld) + 512; }
inline size_t gra_smem_bytes(int ld) { return (size_t)(2 * gra_buf_floats(ld) + 4096 + 2 * kGraCsz + 32 * kGraVs + 128 + 32) * sizeof(float); }

__device__ __forceinline__ void gra_cluster_sync() {
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Rotation of the column pair (x, y) with x.x = a, y.y = b, x.y = c:
//   x' = cs x - sn y,  y' = sn x + cs y,  new squared norms a - t c and b + t c.
__device__ __forceinline__ bool gra_params(float a, float b, float c, float tol2, float fl, float& cs, float& sn,
                                           float& t) {
  cs = 1.f;
  sn = 0.f;
  t = 0.f;
  if (!(a > fl) || !(b > fl)) return false;
  if (!(c * c > (tol2 * a) * b)) return false;
  // With al = (b - a)/2 and hyp = sqrt(al^2 + c^2):  cs^2 = (1 + |al|/hyp)/2,  sn = sign(al) c / (2 hyp cs),
  // t = sn/cs -- the same inner rotation as t = sign(z)/(|z| + sqrt(1 + z^2)), z = al/c, but with two
  // dependent MUFU ops instead of four.  One Newton step on each rsqrt keeps cs^2 + sn^2 = 1 to rounding.
  const float al = 0.5f * (b - a);
  const float h2 = fmaf(al, al, c * c);
  float ih = mufu_rsqrt(h2);
  ih = ih * fmaf(-0.5f * h2, ih * ih, 1.5f);
  const float cs2 = fmaf(0.5f * fabsf(al), ih, 0.5f);
  float rc = mufu_rsqrt(cs2);
  rc = rc * fmaf(-0.5f * cs2, rc * rc, 1.5f);
  cs = cs2 * rc;
  sn = copysignf(0.5f * c * ih * rc, al * c);
  t = sn * rc;
  return true;
}

// part[warp][i*16 + j] = sum over this warp's rows of x_i * y_j   (i, j in 0..15)
__device__ __forceinline__ void gra_gram16(const float* __restrict__ xb, const float* __restrict__ yb,
                                           float* __restrict__ part, int ld4, int lds, int warp, int lane) {
  const int li = lane & 7, lj = lane >> 3;
  float2 acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[a][m] = make_float2(0.f, 0.f);
  const float* x0p = xb + li * lds;
  const float* x1p = xb + (li + 8) * lds;
  const float* yp = yb + lj * lds;
  for (int r4 = warp; r4 < ld4; r4 += 16) {
    const float4 x0 = *reinterpret_cast<const float4*>(x0p + 4 * r4);
    const float4 x1 = *reinterpret_cast<const float4*>(x1p + 4 * r4);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float4 y = *reinterpret_cast<const float4*>(yp + 4 * m * lds + 4 * r4);
      const float2 ylo = make_float2(y.x, y.y), yhi = make_float2(y.z, y.w);
      acc[0][m] = ffma2(make_float2(x0.x, x0.y), ylo, acc[0][m]);
      acc[1][m] = ffma2(make_float2(x1.x, x1.y), ylo, acc[1][m]);
      acc[0][m] = ffma2(make_float2(x0.z, x0.w), yhi, acc[0][m]);
      acc[1][m] = ffma2(make_float2(x1.z, x1.w), yhi, acc[1][m]);
    }
  }
  float* out = part + warp * 256;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    out[li * 16 + lj + 4 * m] = acc[0][m].x + acc[0][m].y;
    out[(li + 8) * 16 + lj + 4 * m] = acc[1][m].x + acc[1][m].y;
  }
}

// [T' B'] = [T B] V.  Lane tile RL rows x 8 columns (the balanced shared-memory / FFMA2 shape: per
// contraction index 2 (1) 128-bit loads of M and 2 of V feed 8*RL FMAs), warp tile 8*RL rows x all 32
// columns, accumulators packed along row pairs so that M pairs and the float4 stores are natural.
// T' goes to dstT, B' to dstB (possibly another CTA's shared memory).
template <int RL>
__device__ __forceinline__ void gra_apply(const float* __restrict__ src, const float* __restrict__ Vm,
                                          float* __restrict__ dstT, float* __restrict__ dstB, int ld, int lds,
                                          int warp, int lane) {
  // lane (cg, rg): columns 8cg..8cg+7; rows row0..row0+3 and (RL == 8) row0+32..row0+35, so that the
  // 8 lanes of a quarter-warp touch 128 contiguous bytes of a column in every 128-bit access
  const int cg = lane & 3, rg = lane >> 2;
  const int row0 = warp * (8 * RL) + rg * 4;
  if (row0 >= ld) return;
  float2 acc[8][RL / 2];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < RL / 2; ++q) acc[j][q] = make_float2(0.f, 0.f);
  const float* mp = src + row0;
  const float* vp = Vm + 8 * cg;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    float2 m[RL / 2];
    {
      const float4 t = *reinterpret_cast<const float4*>(mp + i * lds);
      m[0] = make_float2(t.x, t.y);
      m[1] = make_float2(t.z, t.w);
      if constexpr (RL == 8) {
        const float4 u = *reinterpret_cast<const float4*>(mp + i * lds + 32);
        m[2] = make_float2(u.x, u.y);
        m[3] = make_float2(u.z, u.w);
      }
    }
    const float4 v0 = *reinterpret_cast<const float4*>(vp + i * kGraVs);
    const float4 v1 = *reinterpret_cast<const float4*>(vp + i * kGraVs + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 vv = make_float2(v[j], v[j]);
#pragma unroll
      for (int q = 0; q < RL / 2; ++q) acc[j][q] = ffma2(m[q], vv, acc[j][q]);
    }
  }
  float* dst = (cg < 2 ? dstT + (8 * cg) * lds : dstB + (8 * cg - 16) * lds) + row0;
  const bool hi = (RL == 8) && (row0 + 32 < ld);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<float4*>(dst + j * lds) = make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y);
    if constexpr (RL == 8) {
      if (hi) *reinterpret_cast<float4*>(dst + j * lds + 32) = make_float4(acc[j][2].x, acc[j][2].y, acc[j][3].x, acc[j][3].y);
    }
  }
}

// rotation record of one pair for one step: cs, sn, and the diagonal entries after the rotation
// (unchanged values and sn == 0 when the pair is skipped)
__device__ __forceinline__ float4 gra_record(float a, float b, float c, float tol2, float fl, int& nrot,
                                             float& maxrel2) {
  float cs, sn, t;
  if (a > fl && b > fl) maxrel2 = fmaxf(maxrel2, c * c * mufu_rcp(a * b));
  if (gra_params(a, b, c, tol2, fl, cs, sn, t)) {
    ++nrot;
    const float d = t * c;
    return make_float4(cs, sn, fmaxf(a - d, 0.f), fmaxf(b + d, 0.f));
  }
  return make_float4(1.f, 0.f, a, b);
}

// One step for the thread that owns the 2 x 2 block (rows pa, qa) x (columns pb, qb) of C.
// Returns the new C[pa][qb].
__device__ __forceinline__ float gra_block(const float* __restrict__ Cc, float* __restrict__ Cn, int pa, int qa,
                                           int pb, int qb, const float4 ra, const float4 rb, bool diag) {
  const float b00 = Cc[pa * kGraCs + pb], b01 = Cc[pa * kGraCs + qb];
  const float b10 = Cc[qa * kGraCs + pb], b11 = Cc[qa * kGraCs + qb];
  float n00, n01, n10, n11;
  if (diag) {
    const bool rot = ra.y != 0.f;
    n00 = ra.z;
    n11 = ra.w;
    n01 = rot ? 0.f : b01;
    n10 = rot ? 0.f : b10;
  } else {
    const float t00 = fmaf(rb.x, b00, -rb.y * b01), t01 = fmaf(rb.y, b00, rb.x * b01);
    const float t10 = fmaf(rb.x, b10, -rb.y * b11), t11 = fmaf(rb.y, b10, rb.x * b11);
    n00 = fmaf(ra.x, t00, -ra.y * t10);
    n01 = fmaf(ra.x, t01, -ra.y * t11);
    n10 = fmaf(ra.y, t00, ra.x * t10);
    n11 = fmaf(ra.y, t01, ra.x * t11);
  }
  Cn[pa * kGraCs + pb] = n00;
  Cn[pa * kGraCs + qb] = n01;
  Cn[qa * kGraCs + pb] = n10;
  Cn[qa * kGraCs + qb] = n11;
  return n01;
}

__device__ __forceinline__ void gra_vrot(float* __restrict__ Vm, int i0, int pb, int qb, const float4 rb) {
  if (rb.y == 0.f) return;
#pragma unroll
  for (int i = i0; i < i0 + 2; ++i) {
    const float x = Vm[i * kGraVs + pb], y = Vm[i * kGraVs + qb];
    Vm[i * kGraVs + pb] = fmaf(rb.x, x, -rb.y * y);
    Vm[i * kGraVs + qb] = fmaf(rb.y, x, rb.x * y);
  }
}

// The 16 cross steps (pair w of step s = column w of T against column (w+s)%16 of B) on C (double
// buffered) and V, one barrier per step: the thread that produces the new C[w][16+(w+s+1)%16] --
// the pivot of pair w in the next step -- also computes that pair's rotation record, so the
// parameter chain of step s+1 is off the barrier path.  Returns the buffer index of the final C.
__device__ __forceinline__ int gra_rotate_cross(float* __restrict__ Cb, float* __restrict__ Vm, float4* __restrict__ rec,
                                                int tid, float tol2, float fl, int& nrot, float& maxrel2) {
  if (tid < 16) {
    const int p = tid, q = 16 + tid;
    rec[tid] = gra_record(Cb[p * kGraCs + p], Cb[q * kGraCs + q], Cb[p * kGraCs + q], tol2, fl, nrot, maxrel2);
  }
  __syncthreads();
  int cur = 0;
#pragma unroll 1
  for (int s = 0; s < 16; ++s) {
    const float* Cc = Cb + cur * kGraCsz;
    float* Cn = Cb + (cur ^ 1) * kGraCsz;
    const float4* rc = rec + cur * 16;
    if (tid < 256) {
      // thread -> block (a, b) with b = a + dd: warp 0 owns the 16 diagonal blocks (dd == 0) and the 16
      // blocks that hold the next step's pivots (dd == 1), so the rotation-parameter chain is issued by one
      // warp instead of diverging two lanes in each of the eight
      const int dd = tid >> 4, a = tid & 15, b = (a + dd) & 15;
      const int qa = 16 + ((a + s) & 15), qb = 16 + ((b + s) & 15);
      const float4 ra = rc[a], rb = rc[b];
      const float n01 = gra_block(Cc, Cn, a, qa, b, qb, ra, rb, dd == 0);
      if (dd == 1 && s < 15)
        rec[(cur ^ 1) * 16 + a] = gra_record(ra.z, rb.w, n01, tol2, fl, nrot, maxrel2);
    } else {
      const int u = tid - 256, b = u & 15;
      gra_vrot(Vm, (u >> 4) * 2, b, 16 + ((b + s) & 15), rc[b]);
    }
    cur ^= 1;
    __syncthreads();
  }
  return cur;
}

// The 15 tournament steps inside T and inside B (once per sweep): two barriers per step.
__device__ __forceinline__ int gra_rotate_intra(float* __restrict__ Cb, float* __restrict__ Vm, float4* __restrict__ rec,
                                                int* __restrict__ pq, int tid, float tol2, float fl, int& nrot,
                                                float& maxrel2) {
  int cur = 0;
#pragma unroll 1
  for (int s = 0; s < 15; ++s) {
    const float* Cc = Cb + cur * kGraCsz;
    float* Cn = Cb + (cur ^ 1) * kGraCsz;
    if (tid < 16) {
      int p0, p1;
      rr_pair(16, s, tid & 7, p0, p1);
      const int base = (tid >> 3) << 4;
      const int p = base + (p0 < p1 ? p0 : p1), q = base + (p0 < p1 ? p1 : p0);
      pq[2 * tid] = p;
      pq[2 * tid + 1] = q;
      rec[tid] = gra_record(Cc[p * kGraCs + p], Cc[q * kGraCs + q], Cc[p * kGraCs + q], tol2, fl, nrot, maxrel2);
    }
    __syncthreads();
    if (tid < 256) {
      const int a = tid >> 4, b = tid & 15;
      gra_block(Cc, Cn, pq[2 * a], pq[2 * a + 1], pq[2 * b], pq[2 * b + 1], rec[a], rec[b], a == b);
    } else {
      const int u = tid - 256, b = u & 15;
      gra_vrot(Vm, (u >> 4) * 2, pq[2 * b], pq[2 * b + 1], rec[b]);
    }
    cur ^= 1;
    __syncthreads();
  }
  return cur;
}

}  // namespace tta
