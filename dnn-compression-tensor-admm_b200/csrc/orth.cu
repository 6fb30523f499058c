// Orthogonality regulariser of the decomposed layers' factor matrices (orthogonal.py:9-20, used by the
// fine-tune loop engines.py:290-291,297-298):  loss += rho/2 * || F F^T - I ||_F^2  (F with fewer rows than
// columns) or  rho/2 * || F^T F - I ||_F^2  (otherwise), and its gradient  2 rho (F F^T - I) F  /  2 rho F (F^T F - I).
// Both orientations are one problem: n vectors x_i of length len (element t of vector i at p + i*si + t*st),
//     R = X X^T - I   (n x n),     loss += rho/2 sum R^2,     grad x_i += g * 2 rho * sum_j R_ij x_j.
// The reference spends a torch.mm, eye, sub, norm, pow, mul, add (+ their autograd twins) per factor and step;
// here all factors of a model go through two batched launches (forward keeps R for the backward).
#include "tta_common.cuh"

namespace tta {

constexpr int kOrthTile = 32;

__device__ __forceinline__ float orth_x(const tta_orth_task& tk, int vec, int pos) {
  return (vec < tk.n && pos < tk.len) ? tk.p[(int64_t)vec * tk.si + (int64_t)pos * tk.st] : 0.f;
}

// grid (tiles_j, tiles_i, task); block 32 x 8
__global__ void __launch_bounds__(256) orth_fwd_kernel(const tta_orth_task* __restrict__ tasks, float half_rho,
                                                       double* __restrict__ loss) {
  __shared__ float A[kOrthTile][kOrthTile + 1], B[kOrthTile][kOrthTile + 1];
  __shared__ double s_part[8];
  const tta_orth_task tk = tasks[blockIdx.z];
  const int i0 = blockIdx.y * kOrthTile, j0 = blockIdx.x * kOrthTile;
  if (i0 >= tk.n || j0 >= tk.n) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool pos_fast = tk.st == 1;     // consecutive threads walk the contiguous index
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t0 = 0; t0 < tk.len; t0 += kOrthTile) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int a = ty + 8 * q;
      const int vec = pos_fast ? a : tx, pos = pos_fast ? tx : a;
      A[vec][pos] = orth_x(tk, i0 + vec, t0 + pos);
      B[vec][pos] = orth_x(tk, j0 + vec, t0 + pos);
    }
    __syncthreads();
#pragma unroll 8
    for (int pos = 0; pos < kOrthTile; ++pos) {
      const float b = B[tx][pos];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(A[ty + 8 * q][pos], b, acc[q]);
    }
    __syncthreads();
  }
  double sq = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + ty + 8 * q, j = j0 + tx;
    if (i < tk.n && j < tk.n) {
      const float r = acc[q] - (i == j ? 1.f : 0.f);
      tk.r[(int64_t)i * tk.n + j] = r;
      sq += (double)r * (double)r;
    }
  }
  sq = warp_sum(sq);
  if (tx == 0) s_part[ty] = sq;
  __syncthreads();
  if (tx == 0 && ty == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += s_part[w];
    atomicAdd(loss, (double)half_rho * s);
  }
}

// grad x_i[t] (+)= scale * sum_j R[i][j] x_j[t];  grid (tiles_t, tiles_i, task); block 32 x 8
__global__ void __launch_bounds__(256) orth_bwd_kernel(const tta_orth_task* __restrict__ tasks, float two_rho,
                                                       const float* __restrict__ grad_scale, int accumulate) {
  __shared__ float Rt[kOrthTile][kOrthTile + 1], X[kOrthTile][kOrthTile + 1];   // Rt[i][j], X[j][t]
  const tta_orth_task tk = tasks[blockIdx.z];
  const int i0 = blockIdx.y * kOrthTile, t0 = blockIdx.x * kOrthTile;
  if (i0 >= tk.n || t0 >= tk.len) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool pos_fast = tk.st == 1;
  const float scale = two_rho * (grad_scale ? grad_scale[0] : 1.f);
  // thread -> outputs: pos_fast: (vec = ty + 8q, pos = tx) else (vec = tx, pos = ty + 8q): coalesced stores
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < tk.n; j0 += kOrthTile) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int a = ty + 8 * q;
      Rt[a][tx] = (i0 + a < tk.n && j0 + tx < tk.n) ? tk.r[(int64_t)(i0 + a) * tk.n + j0 + tx] : 0.f;
      const int vec = pos_fast ? a : tx, pos = pos_fast ? tx : a;
      X[vec][pos] = orth_x(tk, j0 + vec, t0 + pos);
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < kOrthTile; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int vec = pos_fast ? ty + 8 * q : tx, pos = pos_fast ? tx : ty + 8 * q;
        acc[q] = fmaf(Rt[vec][j], X[j][pos], acc[q]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int vec = i0 + (pos_fast ? ty + 8 * q : tx), pos = t0 + (pos_fast ? tx : ty + 8 * q);
    if (vec < tk.n && pos < tk.len) {
      float* g = tk.g + (int64_t)vec * tk.si + (int64_t)pos * tk.st;
      *g = accumulate ? fmaf(scale, acc[q], *g) : scale * acc[q];
    }
  }
}

static int orth_validate(const tta_orth_task* th, int n, bool need_g, int* nmax, int* lmax) {
  if (n < 0 || (n > 0 && !th)) {
    set_error("orth: bad task table");
    return TTA_E_INVALID;
  }
  *nmax = *lmax = 0;
  for (int t = 0; t < n; ++t) {
    const tta_orth_task& tk = th[t];
    if (tk.n <= 0 || tk.len <= 0 || !tk.p || !tk.r || (need_g && !tk.g) || (tk.si != 1 && tk.st != 1)) {
      set_error("orth: task %d invalid (n=%d len=%d)", t, tk.n, tk.len);
      return TTA_E_INVALID;
    }
    *nmax = tk.n > *nmax ? tk.n : *nmax;
    *lmax = tk.len > *lmax ? tk.len : *lmax;
  }
  return TTA_OK;
}

}  // namespace tta

extern "C" {

int tta_orth_penalty_fwd_batched(const tta_orth_task* tasks_dev, const tta_orth_task* tasks_host, int n_tasks, float rho,
                                 double* loss_out, void* stream) {
  using namespace tta;
  int nmax, lmax;
  int rc = orth_validate(tasks_host, n_tasks, false, &nmax, &lmax);
  if (rc || n_tasks == 0) return rc;
  if (!tasks_dev || !loss_out) {
    set_error("orth_fwd: null argument");
    return TTA_E_INVALID;
  }
  const unsigned tiles = (unsigned)((nmax + kOrthTile - 1) / kOrthTile);
  orth_fwd_kernel<<<dim3(tiles, tiles, (unsigned)n_tasks), dim3(32, 8), 0, (cudaStream_t)stream>>>(tasks_dev, 0.5f * rho,
                                                                                                   loss_out);
  TTA_CHECK_LAUNCH("orth_fwd launch");
  return TTA_OK;
}

int tta_orth_penalty_bwd_batched(const tta_orth_task* tasks_dev, const tta_orth_task* tasks_host, int n_tasks, float rho,
                                 const float* grad_scale, int accumulate, void* stream) {
  using namespace tta;
  int nmax, lmax;
  int rc = orth_validate(tasks_host, n_tasks, true, &nmax, &lmax);
  if (rc || n_tasks == 0) return rc;
  if (!tasks_dev) {
    set_error("orth_bwd: null argument");
    return TTA_E_INVALID;
  }
  orth_bwd_kernel<<<dim3((unsigned)((lmax + kOrthTile - 1) / kOrthTile), (unsigned)((nmax + kOrthTile - 1) / kOrthTile),
                         (unsigned)n_tasks),
                    dim3(32, 8), 0, (cudaStream_t)stream>>>(tasks_dev, 2.f * rho, grad_scale, accumulate);
  TTA_CHECK_LAUNCH("orth_bwd launch");
  return TTA_OK;
}
}
