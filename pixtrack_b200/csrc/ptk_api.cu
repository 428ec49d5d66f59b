// Context management, error reporting and small layout kernels of the C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "ptk_common.cuh"

static thread_local char g_err[512] = "";

void ptk_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ptk_last_error(void) { return g_err; }
extern "C" int ptk_abi_version(void) { return PTK_ABI_VERSION; }

extern "C" int ptk_create(int device, PtkContext** out) {
  PTK_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  int count = 0;
  PTK_CUDA_CHECK(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) {
    ptk_set_error("device %d not present (%d CUDA devices); libpixtrack_b200 has no CPU fallback", device, count);
    return PTK_ERR_CUDA;
  }
  cudaDeviceProp prop;
  PTK_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    ptk_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return PTK_ERR_UNSUPPORTED;
  }
  int prev = 0;
  PTK_CUDA_CHECK(cudaGetDevice(&prev));
  PTK_CUDA_CHECK(cudaSetDevice(device));
  PtkContext* c = (PtkContext*)calloc(1, sizeof(PtkContext));
  c->device = device;
  c->num_sms = prop.multiProcessorCount < PTK_MAX_SMS ? prop.multiProcessorCount : PTK_MAX_SMS;
  cudaError_t e = cudaMalloc(&c->lm_partials, sizeof(float) * PTK_MAX_SMS * 2 * 32);
  if (e == cudaSuccess) e = cudaMalloc(&c->lm_counters, sizeof(unsigned int) * (PTK_MAX_SMS + 8));
  if (e == cudaSuccess) e = cudaMalloc(&c->lm_error, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(c->lm_counters, 0, sizeof(unsigned int) * (PTK_MAX_SMS + 8));
  if (e == cudaSuccess) e = cudaMemset(c->lm_error, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    ptk_set_error("workspace allocation failed: %s", cudaGetErrorString(e));
    ptk_destroy(c);
    return PTK_ERR_CUDA;
  }
  *out = c;
  return PTK_OK;
}

extern "C" void ptk_destroy(PtkContext* c) {
  if (c == nullptr) return;
  if (c->lm_partials) cudaFree(c->lm_partials);
  if (c->lm_counters) cudaFree(c->lm_counters);
  if (c->lm_error) cudaFree(c->lm_error);
  free(c);
}

extern "C" int ptk_num_sms(const PtkContext* c) { return c ? c->num_sms : 0; }

// Synchronising health check: 0 = no device-side abort (barrier time-out) recorded.
extern "C" int ptk_device_status(PtkContext* c) {
  PTK_REQUIRE(c != nullptr, "null context");
  PtkDeviceGuard guard(c->device);
  int flag = 0;
  PTK_CUDA_CHECK(cudaMemcpy(&flag, c->lm_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag != 0) {
    ptk_set_error("device-side abort recorded (LM barrier timed out)");
    return PTK_ERR_CUDA;
  }
  return PTK_OK;
}

// --------------------------------------------------------------------------
// [C][H*W] -> [H*W][C] through a 32-pixel shared-memory tile, optional per-pixel
// L2 normalisation over C (F.normalize(dim=0), eps = 1e-12).
// --------------------------------------------------------------------------
namespace {
constexpr int kTilePix = 32;

__global__ void __launch_bounds__(256) chw_to_hwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                                                         long long HW, int normalize) {
  extern __shared__ float tile[];  // [C][33]
  __shared__ float inv[kTilePix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pix0 = (long long)blockIdx.x * kTilePix;
  const long long pix = pix0 + lane;
  for (int c = warp; c < C; c += 8) tile[c * 33 + lane] = (pix < HW) ? src[(long long)c * HW + pix] : 0.f;
  __syncthreads();
  if (normalize) {
    for (int j = warp; j < kTilePix; j += 8) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float v = tile[c * 33 + j];
        s = fmaf(v, v, s);
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
      if (lane == 0) inv[j] = 1.f / fmaxf(sqrtf(s), 1e-12f);
    }
  } else if (threadIdx.x < kTilePix) {
    inv[threadIdx.x] = 1.f;
  }
  __syncthreads();
  const int total = kTilePix * C;
  for (int idx = threadIdx.x; idx < total; idx += 256) {
    const int j = idx / C, c = idx - j * C;
    if (pix0 + j < HW) dst[(pix0 + j) * C + c] = normalize ? tile[c * 33 + j] * inv[j] : tile[c * 33 + j];
  }
}
}  // namespace

extern "C" int ptk_chw_to_hwc(PtkContext* ctx, const float* src, float* dst, int32_t C, int32_t H, int32_t W,
                              int32_t normalize, void* stream) {
  PTK_REQUIRE(ctx && src && dst, "null argument");
  PtkDeviceGuard guard(ctx->device);
  PTK_REQUIRE(C >= 1 && C <= 1024 && H >= 1 && W >= 1, "bad shape");
  const long long HW = (long long)H * W;
  const size_t smem = (size_t)C * 33 * sizeof(float);
  if (smem > 48 * 1024) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(chw_to_hwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const unsigned blocks = (unsigned)((HW + kTilePix - 1) / kTilePix);
  chw_to_hwc_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(src, dst, C, HW, normalize);
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}

// Stream-ordered device-to-device copy (used by tests to snapshot plan-owned activations).
extern "C" int ptk_copy_d2d(void* dst, const void* src, int64_t bytes, void* stream) {
  PTK_REQUIRE(dst && src && bytes >= 0, "bad argument");
  PTK_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return PTK_OK;
}
