// cfk_common.cuh — helpers shared by the translation units of libcfk.so (cfk.cu, docfreq_stream.cu, placer.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cfk.h"

// one error string per host thread and one launch counter for the whole library (defined in cfk.cu)
extern thread_local char cfk_g_err[512];
extern long long cfk_g_launches;

namespace cfk {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint64_t EMPTY = CFK_EMPTY_KEY;

inline int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  if (e != cudaSuccess)
    snprintf(cfk_g_err, sizeof(cfk_g_err), "%s: %s", what, cudaGetErrorString(e));
  else
    snprintf(cfk_g_err, sizeof(cfk_g_err), "%s", what);
  return code;
}

#define CFK_CHECK_LAUNCH(name, n_launched)                                  \
  do {                                                                     \
    cudaError_t e_ = cudaGetLastError();                                   \
    if (e_ != cudaSuccess) return cfk::fail(CFK_ERR_CUDA, name, e_);       \
    __atomic_fetch_add(&cfk_g_launches, (long long)(n_launched), __ATOMIC_RELAXED); \
  } while (0)

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// home slot of a hashed key in a table of arbitrary capacity (multiply-shift range reduction; monotone in h)
__device__ __forceinline__ int64_t home_slot(uint64_t h, int64_t cap) {
  return (int64_t)__umul64hi(h, (uint64_t)cap);
}

inline int64_t blocks_for(int64_t n, int threads) { return (n + threads - 1) / threads; }

// the dynamic shared-memory attribute is per device: set it once per (kernel, device)
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, unsigned long long* done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
  return e;
}

}  // namespace cfk
