// sm_internal.h -- shared between the translation units of libshardmerge_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "../../include/shardmerge_b200.h"
#include "plan.h"

struct sm_plan {
  SmPlan p;
};

void sm_set_error(const char* fmt, ...);

#define SM_CUDA_CHECK(expr)                                                            \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      sm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -100;                                                                     \
    }                                                                                  \
  } while (0)

#define SM_LAUNCH_CHECK()                                                              \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      sm_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -101;                                                                     \
    }                                                                                  \
  } while (0)

// table layout inside the caller's buffer: [W_C : C entries][W_R : R entries] of float2
static inline size_t sm_tab_off_R(const SmPlan& p) { return (size_t)p.C * 8; }

// iteration space of the element-wise / statistics kernels over the valid half spectrum
#define SM_EW_THREADS 256
#define SM_EW_COLS (SM_EW_THREADS * 4)
