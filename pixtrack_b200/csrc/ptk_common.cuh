// Shared internals of libpixtrack_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "pixtrack_b200.h"

#define PTK_MAX_SMS 320     // CTA slots of a cooperative LM launch (2 per SM on 148 SMs = 296)

struct PtkContext {
  int device;
  int num_sms;
  // LM workspace (device): per-group partial sums and barrier counters.
  float* lm_partials;       // [PTK_MAX_SMS][2][32] floats (slot = cta, parity)
  unsigned int* lm_counters;  // [PTK_MAX_SMS + 8]
  int* lm_error;            // device flag set by a timed-out spin barrier
};

void ptk_set_error(const char* fmt, ...);

// LM workspace layout (floats then counters): [PTK_MAX_SMS][2][32] partial sums, PTK_MAX_SMS + 8 counters
#define PTK_LM_WS_PARTIAL_BYTES (sizeof(float) * PTK_MAX_SMS * 2 * 32)
#define PTK_LM_WS_BYTES (PTK_LM_WS_PARTIAL_BYTES + sizeof(unsigned int) * (PTK_MAX_SMS + 8))

// Makes ctx->device current for the scope of an entry point (the caller's device is restored on exit).
struct PtkDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit PtkDeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~PtkDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

#define PTK_CUDA_CHECK(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ptk_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PTK_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define PTK_REQUIRE(cond, msg)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      ptk_set_error("invalid argument: %s", msg);      \
      return PTK_ERR_INVALID;                          \
    }                                                  \
  } while (0)
