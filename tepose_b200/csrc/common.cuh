// Shared host/device helpers for the tepose_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tepose_b200.h"

namespace tp {

// thread-local error text behind tp_last_error()
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define TP_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::tp::fail(TP_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

#define TP_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::tp::fail(TP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                               \
  } while (0)

#define TP_LAUNCH_CHECK()                                                                   \
  do {                                                                                      \
    ::tp::count_launch();                                                                   \
    cudaError_t _e = cudaPeekAtLastError();                                                 \
    if (_e != cudaSuccess)                                                                  \
      return ::tp::fail(TP_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();  // of the current device (cached)
void count_launch();  // bumps the counter behind tp_launch_count()

}  // namespace tp
