// Fused multi-tensor dual update / penalty kernels (HBM-bound, 128-bit coalesced).
//   admm.py:71-78  U += W - Z  (+ ||W - Z||^2)               16 B/elem  (read W,Z,U; write U)
//   admm.py:80-85  sum 0.5*rho*||W - Z + U||^2                12 B/elem  (read W,Z,U)
//   autograd of :83  g (+)= s*rho*(W - Z + U)                  16 / 20 B/elem
// One launch serves every listed layer: the work is cut into 4096-element chunks and the chunks are
// grid-strided over a grid that is a multiple of the SM count.
#include <stdarg.h>

#include <atomic>

#include "tta_common.cuh"

namespace tta {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Pinned staging ring for small host -> device uploads inside enqueue-only calls.  cudaMemcpyAsync from
// pageable memory first synchronises the stream ("a stream sync is performed before the copy is
// initiated"), which would stall the host once per eigensolver wave; from pinned memory it only enqueues.
// A slot is reused after kRingBytes of later uploads -- far more than a launch queue can hold in flight.
void* pinned_stage(size_t bytes) {
  constexpr size_t kRingBytes = 4u << 20;
  static char* ring = nullptr;
  static size_t head = 0;
  if (!ring) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, kRingBytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    ring = static_cast<char*>(p);
  }
  bytes = (bytes + 63) & ~(size_t)63;
  if (bytes > kRingBytes) return nullptr;
  if (head + bytes > kRingBytes) head = 0;
  void* out = ring + head;
  head += bytes;
  return out;
}

constexpr int kEwThreads = 256;
constexpr int kEwVecPerThread = 4;                                  // float4 per thread per chunk
constexpr int kEwChunk = kEwThreads * kEwVecPerThread * 4;          // 4096 elements
constexpr int kEwMaxTasks = 256;                                    // per launch (table in params)

struct EwTable {
  int n_tasks;
  int total_chunks;
  int chunk_start[kEwMaxTasks + 1];
};

enum { kDualUpdate = 0, kPenaltyFwd = 1, kPenaltyBwd = 2 };

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int MODE>
__device__ __forceinline__ float ew_apply(float w, float z, float& u, float& g, float coef, bool acc) {
  if (MODE == kDualUpdate) {
    float d = w - z;
    u += d;
    return d * d;
  } else if (MODE == kPenaltyFwd) {
    float d = w - z + u;
    return d * d;
  } else {
    float d = coef * (w - z + u);
    g = acc ? g + d : d;
    return 0.f;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(const tta_ew_task* __restrict__ tasks,
                                                        const __grid_constant__ EwTable tab, float rho,
                                                        const float* __restrict__ grad_scale,
                                                        int accumulate, double* __restrict__ out) {
  __shared__ double s_part[kEwThreads / 32];
  const int tid = threadIdx.x;
  float coef = 0.f;
  if (MODE == kPenaltyBwd) coef = rho * (grad_scale ? __ldg(grad_scale) : 1.f);
  const bool acc = accumulate != 0;

  for (int chunk = blockIdx.x; chunk < tab.total_chunks; chunk += gridDim.x) {
    // locate the tensor owning this chunk (uniform across the CTA)
    int lo = 0, hi = tab.n_tasks;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tab.chunk_start[mid] <= chunk) lo = mid; else hi = mid;
    }
    const tta_ew_task tk = tasks[lo];
    const int64_t base = (int64_t)(chunk - tab.chunk_start[lo]) * kEwChunk;
    const int64_t rem = tk.numel - base;
    const int n = rem < kEwChunk ? (int)rem : kEwChunk;
    const float* w = tk.w + base;
    const float* z = tk.z + base;
    float* u = tk.u + base;
    float* g = (MODE == kPenaltyBwd) ? tk.g + base : nullptr;
    uintptr_t align = (uintptr_t)w | (uintptr_t)z | (uintptr_t)u | (uintptr_t)g;
    float local = 0.f;

    if ((align & 15) == 0 && n == kEwChunk) {
      float4 wv[kEwVecPerThread], zv[kEwVecPerThread], uv[kEwVecPerThread], gv[kEwVecPerThread];
#pragma unroll
      for (int j = 0; j < kEwVecPerThread; ++j) {
        const int e = (j * kEwThreads + tid) * 4;
        wv[j] = ld_stream(reinterpret_cast<const float4*>(w + e));
        zv[j] = ld_stream(reinterpret_cast<const float4*>(z + e));
        uv[j] = ld4(u + e);
        if (MODE == kPenaltyBwd && acc) gv[j] = ld4(g + e);
      }
#pragma unroll
      for (int j = 0; j < kEwVecPerThread; ++j) {
        const int e = (j * kEwThreads + tid) * 4;
        local += ew_apply<MODE>(wv[j].x, zv[j].x, uv[j].x, gv[j].x, coef, acc);
        local += ew_apply<MODE>(wv[j].y, zv[j].y, uv[j].y, gv[j].y, coef, acc);
        local += ew_apply<MODE>(wv[j].z, zv[j].z, uv[j].z, gv[j].z, coef, acc);
        local += ew_apply<MODE>(wv[j].w, zv[j].w, uv[j].w, gv[j].w, coef, acc);
        if (MODE == kDualUpdate) st_stream(reinterpret_cast<float4*>(u + e), uv[j]);
        if (MODE == kPenaltyBwd) st_stream(reinterpret_cast<float4*>(g + e), gv[j]);
      }
    } else {
      for (int e = tid; e < n; e += kEwThreads) {
        float uu = u[e];
        float gg = (MODE == kPenaltyBwd && acc) ? g[e] : 0.f;
        local += ew_apply<MODE>(w[e], z[e], uu, gg, coef, acc);
        if (MODE == kDualUpdate) u[e] = uu;
        if (MODE == kPenaltyBwd) g[e] = gg;
      }
    }

    if (MODE != kPenaltyBwd && out != nullptr) {
      double v = warp_sum((double)local);
      if ((tid & 31) == 0) s_part[tid >> 5] = v;
      __syncthreads();
      if (tid < 32) {
        double t = tid < kEwThreads / 32 ? s_part[tid] : 0.0;
        t = warp_sum(t);
        if (tid == 0) {
          if (MODE == kPenaltyFwd) atomicAdd(out, 0.5 * (double)rho * t);
          else atomicAdd(out + lo, t);
        }
      }
      __syncthreads();
    }
  }
}

template <int MODE>
static int launch_ew(const tta_ew_task* tasks_dev, const tta_ew_task* tasks_host, int n_tasks, float rho,
                     const float* grad_scale, int accumulate, double* out, void* stream) {
  if (n_tasks < 0 || (n_tasks > 0 && (!tasks_dev || !tasks_host))) {
    set_error("elementwise: bad task table");
    return TTA_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int first = 0; first < n_tasks; first += kEwMaxTasks) {
    const int cnt = (n_tasks - first) < kEwMaxTasks ? (n_tasks - first) : kEwMaxTasks;
    EwTable tab;
    tab.n_tasks = cnt;
    int total = 0;
    for (int t = 0; t < cnt; ++t) {
      const tta_ew_task& tk = tasks_host[first + t];
      if (tk.numel < 0 || (tk.numel > 0 && (!tk.w || !tk.z || !tk.u)) ||
          (MODE == kPenaltyBwd && tk.numel > 0 && !tk.g)) {
        set_error("elementwise: task %d has null pointer / negative numel", first + t);
        return TTA_E_INVALID;
      }
      tab.chunk_start[t] = total;
      total += (int)((tk.numel + kEwChunk - 1) / kEwChunk);
    }
    tab.chunk_start[cnt] = total;
    tab.total_chunks = total;
    if (total == 0) continue;
    int grid = total < kNumSMs * 8 ? total : kNumSMs * 8;
    double* o = out;
    if (MODE == kDualUpdate && out) o = out + first;
    ew_kernel<MODE><<<grid, kEwThreads, 0, st>>>(tasks_dev + first, tab, rho, grad_scale, accumulate, o);
    TTA_CHECK_LAUNCH("elementwise launch");
  }
  return TTA_OK;
}

}  // namespace tta

extern "C" {

const char* tta_last_error(void) { return tta::g_err; }
int tta_version(void) { return 100; }
unsigned long long tta_launch_count(void) { return tta::g_launches.load(std::memory_order_relaxed); }

int tta_check_device(int dev) {
  cudaDeviceProp p;
  int rc = tta::check_cuda(cudaGetDeviceProperties(&p, dev), "cudaGetDeviceProperties");
  if (rc != TTA_OK) return rc;
  if (p.major != 10) {
    tta::set_error("device %d is sm_%d%d; libtta is built for sm_100a only", dev, p.major, p.minor);
    return TTA_E_ARCH;
  }
  return TTA_OK;
}

int tta_dual_update_multi(const tta_ew_task* tasks_dev, const tta_ew_task* tasks_host, int n_tasks,
                          double* sqnorm_out, void* stream) {
  return tta::launch_ew<tta::kDualUpdate>(tasks_dev, tasks_host, n_tasks, 0.f, nullptr, 0, sqnorm_out, stream);
}

int tta_penalty_fwd_multi(const tta_ew_task* tasks_dev, const tta_ew_task* tasks_host, int n_tasks,
                          float rho, double* loss_out, void* stream) {
  if (!loss_out) {
    tta::set_error("penalty_fwd: loss_out is NULL");
    return TTA_E_INVALID;
  }
  return tta::launch_ew<tta::kPenaltyFwd>(tasks_dev, tasks_host, n_tasks, rho, nullptr, 0, loss_out, stream);
}

int tta_penalty_bwd_multi(const tta_ew_task* tasks_dev, const tta_ew_task* tasks_host, int n_tasks,
                          float rho, const float* grad_scale, int accumulate, void* stream) {
  return tta::launch_ew<tta::kPenaltyBwd>(tasks_dev, tasks_host, n_tasks, rho, grad_scale, accumulate, nullptr,
                                          stream);
}

}  // extern "C"
