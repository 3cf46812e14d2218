// Shared helpers for the sniper_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace sn {

// thread-local last-error string, surfaced through sniper_last_error() (C-ABI mirror of
// MXGetLastError, SNIPER-mxnet/include/mxnet/c_api.h:196-204)
char* last_error_buf();
void set_error(const char* fmt, ...);

#define SN_CHECK(cond, ...)      \
  do {                           \
    if (!(cond)) {               \
      sn::set_error(__VA_ARGS__); \
      return -1;                 \
    }                            \
  } while (0)

#define SN_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) {                                                           \
      sn::set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)

#define SN_LAUNCH_CHECK() SN_CUDA(cudaGetLastError())

static inline int div_up(long a, long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;

}  // namespace sn
