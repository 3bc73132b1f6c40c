// common.cuh -- shared host/device helpers for libpp_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/pp_b200.h"

namespace pp {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return PP_ERR_CUDA;
}

#define PP_CUDA(expr)                                  \
  do {                                                 \
    cudaError_t _e = (expr);                           \
    if (_e != cudaSuccess) return ::pp::cuda_fail(_e); \
  } while (0)

#define PP_LAUNCH_CHECK() PP_CUDA(cudaGetLastError())

// Every kernel launch goes through PP_KERNEL: it counts the launch (pp_launch_count) and, when
// profiling is enabled (pp_profile_enable), brackets it with CUDA events on the launching stream.
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);
#define PP_KERNEL(name, st, ...)   \
  do {                             \
    ::pp::prof_begin(name, st);    \
    __VA_ARGS__;                   \
    ::pp::prof_end(st);            \
    PP_LAUNCH_CHECK();             \
  } while (0)

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t size, used;
  bool ok;
  Arena(void* p, size_t n) : base((char*)p), size(n), used(0), ok(p != nullptr && ((uintptr_t)p % kAlign) == 0) {}
  template <class T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (!ok || used + bytes > size) { ok = false; return nullptr; }
    T* r = (T*)(base + used);
    used += bytes;
    return r;
  }
};

struct SizeCounter {
  size_t used = 0;
  template <class T>
  void take(size_t count) { used += align_up(count * sizeof(T)); }
};

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

}  // namespace pp
