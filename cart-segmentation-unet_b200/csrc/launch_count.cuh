// Process-wide count of kernels this library has launched (reported by cs_kernel_launch_count()).
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace cs {
extern std::atomic<long long> g_kernel_launches;
inline cudaError_t launched() {
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}
}  // namespace cs
