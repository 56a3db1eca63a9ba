// common.cuh -- shared helpers for librn_b200 (error reporting, launch checks, small device utils).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/rn_b200.h"

namespace rn {

// ---- thread-local error message -------------------------------------------------------
char* error_buffer();   // defined in abi.cu
int fail(int code, const char* fmt, ...);

#define RN_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::rn::fail(RN_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define RN_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return ::rn::fail(RN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                               \
  } while (0)

void count_launch();   // abi.cu: process-wide counter behind rn_launch_count()

#define RN_LAUNCH_CHECK(name)                                                              \
  do {                                                                                     \
    ::rn::count_launch();                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess)                                                                \
      return ::rn::fail(RN_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define RN_TRY(expr)          \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != RN_OK) return rc__; \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

int sm_count();   // cached per device (abi.cu)

// bump allocator over a caller-provided buffer (256-byte granules)
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    T* r = reinterpret_cast<T*>(base + off);
    off += round_up(count * sizeof(T), 256);
    return r;
  }
};

}  // namespace rn
