// launch_util.cuh -- opt a kernel in to the device's full dynamic shared memory, once per (kernel, device).
//
// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the FUNCTION, shared by every thread of the process.
// Setting it to the exact size of each launch (as the launchers first did) races when two host threads launch the same
// kernel with different sizes -- a mixed fit drives one layout per thread -- and the loser's launch fails with
// "invalid argument".  The limit is therefore raised to the maximum once and never lowered; the occupancy of a launch
// depends on the shared memory it actually asks for, not on this limit.
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <set>
#include <utility>

namespace xrsd {

inline cudaError_t allow_max_dynamic_smem(const void *kern) {
  static std::mutex mu;
  static std::set<std::pair<const void *, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count(std::make_pair(kern, dev))) return cudaSuccess;
  cudaFuncAttributes attr;
  e = cudaFuncGetAttributes(&attr, kern);
  if (e != cudaSuccess) return e;
  int optin = 0;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)attr.sharedSizeBytes);
  if (e == cudaSuccess) done.insert(std::make_pair(kern, dev));
  return e;
}

template <class K>
inline cudaError_t allow_max_dynamic_smem(K kern) {
  return allow_max_dynamic_smem(reinterpret_cast<const void *>(kern));
}

}  // namespace xrsd
