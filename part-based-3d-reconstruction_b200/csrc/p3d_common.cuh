// Shared helpers for libp3d_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/p3d_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libp3d_b200 is written for sm_100a (B200) only"
#endif

#define P3D_API extern "C" __attribute__((visibility("default")))

namespace p3d {

void set_error(const char* fmt, ...);

inline cudaStream_t as_stream(p3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached per thread)
int sm_count();

}  // namespace p3d

#define P3D_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      p3d::set_error(__VA_ARGS__);      \
      return P3D_E_INVALID;             \
    }                                   \
  } while (0)

#define P3D_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      p3d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                     __LINE__);                                                         \
      return P3D_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

#define P3D_LAUNCH_CHECK() P3D_CUDA(cudaGetLastError())

static inline size_t p3d_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
