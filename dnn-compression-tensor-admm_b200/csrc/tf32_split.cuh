// 3xTF32 operand staging shared by the tensor-core Gram and GEMM kernels (gram_tc.cu, gemm_tf32.cu):
// raw fp32 tile in shared memory (landed by TMA or cp.async) -> hi = tf32(v), lo = tf32(v - hi) in the K-major
// SWIZZLE_128B layout of the UMMA descriptors (rows of 32 fp32 = 128 bytes, 16-byte chunk index XOR row & 7).
#pragma once
#include "tc_common.cuh"

namespace tta {
namespace tf32 {

// round-to-nearest (ties away) to the 10-bit TF32 mantissa: what cvt.rna.tf32.f32 does for finite inputs, without its
// Inf / NaN guard (3 instructions per value; the operands here are finite weights).  The tensor core ignores the 13 low
// mantissa bits, so `hi` needs them cleared only because `lo` is computed from it.
__device__ __forceinline__ float round_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float round_lo(float v) { return __uint_as_float(__float_as_uint(v) + 0x1000u); }

// One 16-byte chunk (reduction indices 4*ch .. 4*ch+3 of tile row `row`) -> hi / lo tiles.
__device__ __forceinline__ void store_chunk(uint32_t hi_tile, uint32_t lo_tile, int row, int ch, float4 v) {
  const float4 h = make_float4(round_hi(v.x), round_hi(v.y), round_hi(v.z), round_hi(v.w));
  const float4 l = make_float4(round_lo(v.x - h.x), round_lo(v.y - h.y), round_lo(v.z - h.z), round_lo(v.w - h.w));
  const uint32_t off = (uint32_t)(row * 128 + ((ch ^ (row & 7)) << 4));
  tc::sts128(hi_tile + off, __float_as_uint(h.x), __float_as_uint(h.y), __float_as_uint(h.z), __float_as_uint(h.w));
  tc::sts128(lo_tile + off, __float_as_uint(l.x), __float_as_uint(l.y), __float_as_uint(l.z), __float_as_uint(l.w));
}

__device__ __forceinline__ float lds32(uint32_t addr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
  return r;
}

// `rows` x 32 reduction indices of a raw tile (+ the same indices of a second raw tile, summed) -> hi / lo tiles.
//   rowfast == false: element (row, k) at src + row * pitch + k * 4          (reduction index contiguous)
//   rowfast == true:  element (row, k) at src + k * pitch + row * 4          (operand index contiguous: transposes)
// THREADS threads cooperate; a thread owns up to MAXC 16-byte chunks and issues ALL its shared-memory loads before the
// first split (the loop is latency bound otherwise).  rows <= 128, a multiple of 8.
template <int THREADS>
__device__ __forceinline__ void transform(uint32_t src, uint32_t src2, bool has2, uint32_t hi_tile, uint32_t lo_tile, int rows,
                                          uint32_t pitch, bool rowfast, int tid) {
  constexpr int MAXC = (128 * 8 + THREADS - 1) / THREADS;
  const int items = rows * 8;
  float4 v[MAXC];
  if (!rowfast) {
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = tid + i * THREADS;
      if (c < items) {
        const uint32_t o = (uint32_t)(c >> 3) * pitch + (uint32_t)(c & 7) * 16u;
        v[i] = tc::lds128(src + o);
        if (has2) {
          const float4 w = tc::lds128(src2 + o);
          v[i].x += w.x; v[i].y += w.y; v[i].z += w.z; v[i].w += w.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = tid + i * THREADS;
      if (c < items) store_chunk(hi_tile, lo_tile, c >> 3, c & 7, v[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = tid + i * THREADS;
      if (c < items) {
        const int row = c % rows, ch = c / rows;
        const uint32_t o = (uint32_t)(ch * 4) * pitch + (uint32_t)row * 4u;
        v[i] = make_float4(lds32(src + o), lds32(src + o + pitch), lds32(src + o + 2 * pitch), lds32(src + o + 3 * pitch));
        if (has2) {
          v[i].x += lds32(src2 + o);
          v[i].y += lds32(src2 + o + pitch);
          v[i].z += lds32(src2 + o + 2 * pitch);
          v[i].w += lds32(src2 + o + 3 * pitch);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = tid + i * THREADS;
      if (c < items) store_chunk(hi_tile, lo_tile, c % rows, c / rows, v[i]);
    }
  }
}

// instruction descriptor: D = fp32, A = B = tf32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc_, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc_), "r"(acc)
      : "memory");
}

// one k-block (32 reduction indices) of the 3xTF32 product into a zeroed accumulator: small terms first
// (A_lo B_hi, A_hi B_lo, then A_hi B_hi); 8 tf32 = 32 bytes inside the 128-byte swizzle row = +2 in 16-byte units
__device__ __forceinline__ void mma_kblock(uint32_t d_tmem, uint64_t da_hi, uint64_t da_lo, uint64_t db_hi, uint64_t db_lo,
                                           uint32_t idesc_) {
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint64_t da = pass == 0 ? da_lo : da_hi;
    const uint64_t db = pass == 1 ? db_lo : db_hi;
#pragma unroll
    for (int k8 = 0; k8 < 4; ++k8) mma(d_tmem, da + (uint64_t)(2 * k8), db + (uint64_t)(2 * k8), idesc_, (pass | k8) ? 1u : 0u);
  }
}

__device__ __forceinline__ void ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr));
}

// drain one 64-column half of a 128-lane accumulator into fp32 registers (round to nearest): columns >= n are skipped
__device__ __forceinline__ void drain64(uint32_t taddr, int cbase, int n, float (&acc)[64]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t v[2][16];
#pragma unroll
    for (int c = 0; c < 2; ++c)
      if (cbase + (h * 2 + c) * 16 < n) ld16(taddr + (uint32_t)((h * 2 + c) * 16), v[c]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 2; ++c)
      if (cbase + (h * 2 + c) * 16 < n) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[(h * 2 + c) * 16 + j] += __uint_as_float(v[c][j]);
      }
  }
}

}  // namespace tf32
}  // namespace tta
