// Shared internals of libpixtrack_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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

// Level-0 head of the extractor, fused into the epilogue of the last decoder convolution (ptk_conv.cu, library-internal).
// The 33 x 32 weights travel as a KERNEL PARAMETER: they then sit in the constant bank and every FMA of the head takes
// its weight as a constant operand -- no load instruction and no shared-memory traffic next to the tensor core's operand
// fetch (a first version that kept them in shared memory slowed the convolution down by more than the head costs).
struct PtkHeadConst {
  float w[33 * 32];   // 32 adaptation rows + the uncertainty row, fp32 copies of the fp16 weights
  float b[33];
  float* feat;        // [H][W][32]
  float* conf;        // [H][W]
  int normalize;
};
int ptk_conv_f16_head(PtkContext* ctx, const void* in0, int32_t cin0, const void* in1, int32_t cin1, int32_t H, int32_t W,
                      int32_t in0_H, int32_t in0_W, int32_t in1_H, int32_t in1_W, const void* weights, const float* bias,
                      int32_t Cout, int32_t taps, int32_t relu, void* out, void* stream, const PtkHeadConst* head, int* fused);

// ---- programmatic dependent launch (PDL) ----
// Kernels of a chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may
// start (barrier / TMEM set-up, tensor-map and weight prefetch) while the previous kernel drains, and block in
// ptk_pdl_wait() until the previous grid has COMPLETED and its writes are visible -- before their first access to anything
// a predecessor wrote, and before their first write.  Rule: every kernel launched through ptk_launch_pdl calls
// ptk_pdl_wait() on at least one thread of every CTA before it exits (completion of a kernel then implies completion of
// everything before it in the stream), and ptk_pdl_trigger() only after the wait (at most two kernels of a chain are
// in flight).  PTK_PDL=0 launches without the attribute (the device calls are then no-ops).
#ifdef __CUDACC__
__device__ __forceinline__ void ptk_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ptk_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline bool ptk_pdl_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("PTK_PDL");
    mode = e ? atoi(e) : 1;
  }
  return mode != 0;
}

// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute (and an optional runtime cluster shape)
template <typename... KArgs, typename... Args>
static inline cudaError_t ptk_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         dim3 cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster.x;
    attr[n].val.clusterDim.y = cluster.y;
    attr[n].val.clusterDim.z = cluster.z;
    ++n;
  }
  if (ptk_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

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
