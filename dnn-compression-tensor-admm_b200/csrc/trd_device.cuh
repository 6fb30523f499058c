// Scalar building blocks of the tridiagonal eigensolver (trd.cu).  Plain C++ so that tests/ can compile the
// very same recurrences for the host (oracle-side check of the Sturm count and the twisted factorisation).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TRD_HD __host__ __device__ __forceinline__
#else
#define TRD_HD inline
#endif

namespace tta {

// Layout of tta_symeig_task.work (offsets in doubles).  kp = k rounded up to a multiple of 32.
struct TrdLayout {
  int kp;
  int64_t hdr;    // 8 doubles: [0] power-of-two scale applied to T by the eigenvalue stage
  int64_t d;      // kp  diagonal of T
  int64_t e;      // kp  off-diagonal (k - 1 used)
  int64_t tau;    // kp  Householder scalars (k - 2 used)
  int64_t lams;   // kp  the r dominant eigenvalues of the SCALED T, descending
  int64_t v;      // k x kp: row j = reflector j (zero up to index j, 1 at j + 1)
  int64_t cross;  // 8 doubles per group of four reflectors (rows j0, j0-1, j0-2, j0-3; j0 = k-3-4g):
                  // b.a, c.a, c.b, d.a, d.b, d.c (back-transformation, trd.cu)
  int64_t total;
};

TRD_HD TrdLayout trd_layout(int k, int r) {
  TrdLayout L;
  L.kp = (k + 31) & ~31;
  L.hdr = 0;
  L.d = 8;
  L.e = L.d + L.kp;
  L.tau = L.e + L.kp;
  L.lams = L.tau + L.kp;
  L.v = L.lams + L.kp;
  L.cross = L.v + (int64_t)k * L.kp;
  L.total = L.cross + 8 * (int64_t)((k + 1) / 4 + 1);
  (void)r;
  return L;
}

constexpr double kTrdPivMin = 7.888609052210118e-31;   // 2^-100 on the scaled problem (|T| <= 1)

TRD_HD double trd_pow2(int ex) {   // 2^ex, -1022 <= ex <= 1023
  const uint64_t bits = (uint64_t)(ex + 1023) << 52;
  double r;
#if defined(__CUDA_ARCH__)
  r = __longlong_as_double((long long)bits);
#else
  memcpy(&r, &bits, sizeof(r));
#endif
  return r;
}

TRD_HD int trd_exponent(double x) {   // unbiased exponent field (-1023 for zero / denormals)
  uint64_t bits;
#if defined(__CUDA_ARCH__)
  bits = (uint64_t)__double_as_longlong(x);
#else
  memcpy(&bits, &x, sizeof(bits));
#endif
  return (int)((bits >> 52) & 0x7ff) - 1023;
}

// Number of eigenvalues of the symmetric tridiagonal T (diagonal dd[0..k), squared off-diagonals
// ee2[i] = e_{i-1}^2, ee2[0] = 0; scaled so that |T| <= 1) that are smaller than x.
// Sturm sequence in product form p_i = (d_{i-1} - x) p_{i-1} - e_{i-2}^2 p_{i-2}: one dependent FMA per
// step instead of a division.  |p_i| < pivmin |p_{i-1}| is replaced by -pivmin p_{i-1} (the quotient
// form's guard, LAPACK dlaebz), the pair is renormalised by a power of two whenever it leaves 2^+-300.
TRD_HD int trd_sturm_count(const double* dd, const double* ee2, int k, double x) {
  double pp = 1.0;
  double p = dd[0] - x;
  if (fabs(p) < kTrdPivMin) p = -kTrdPivMin;
  int cnt = p < 0.0 ? 1 : 0;
  for (int i = 1; i < k; ++i) {
    double pn = fma(dd[i] - x, p, -(ee2[i] * pp));
    if (fabs(pn) < kTrdPivMin * fabs(p)) pn = -kTrdPivMin * p;
    cnt += ((pn < 0.0) != (p < 0.0)) ? 1 : 0;
    pp = p;
    p = pn;
    const int ex = trd_exponent(p);
    if (ex > 300 || ex < -300) {
      const double sc = trd_pow2(-ex);
      p *= sc;
      pp *= sc;
    }
  }
  return cnt;
}

// Quotient-form pivot with the same guard.
TRD_HD double trd_guard(double q) { return fabs(q) < kTrdPivMin ? -kTrdPivMin : q; }

}  // namespace tta
