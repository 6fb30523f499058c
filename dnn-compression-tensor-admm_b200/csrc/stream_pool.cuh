// Internal side streams of the batched solvers (eig_cluster.cu, trd.cu): one pool per (device, calling stream).
#pragma once
#include <cuda_runtime.h>

namespace tta {

constexpr int kPoolStreams = 8;

struct StreamPool {
  cudaStream_t s[kPoolStreams] = {};
  cudaEvent_t fork = nullptr;
  cudaEvent_t join[kPoolStreams] = {};
  int base_prio = 0, prio_lo = 0;
  cudaStream_t get(int i) {
    if (!s[i]) {
      int prio = base_prio + i;
      if (prio > prio_lo) prio = prio_lo;
      if (cudaStreamCreateWithPriority(&s[i], cudaStreamNonBlocking, prio) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return s[i];
  }
};

// defined in eig_cluster.cu
StreamPool* pool_for(int dev, cudaStream_t caller);

}  // namespace tta
