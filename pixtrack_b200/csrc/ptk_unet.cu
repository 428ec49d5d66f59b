// The PixLoc UNet feature extractor as a native plan: small CUDA-core kernels around the
// tcgen05 convolution (ptk_conv.cu), and the layer schedule.
//
// Replaces UNet._forward (reference pixloc/pixloc/pixlib/models/unet.py:158-190) with the PixLoc
// configuration (pixlib/configs/train_pixloc_megadepth.yaml:22-31: vgg19 encoder, decoder
// [64,64,64,32], heads at scales 0/2/4 with 32/128/128 channels + uncertainty), and the
// pre-processing of PixTrackFeatureExtractor.__call__ (pixtrack/localization/feature_extractor.py:34-59).
// Activations are channels-last fp16 (fp32 accumulation everywhere); outputs are fp32.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "ptk_common.cuh"

extern "C" int ptk_conv_f16(PtkContext* ctx, const void* in0, int32_t cin0, const void* in1, int32_t cin1, int32_t H,
                            int32_t W, int32_t in0_H, int32_t in0_W, int32_t in1_H, int32_t in1_W, const void* weights,
                            const float* bias, int32_t Cout, int32_t taps, int32_t relu, void* out, void* stream);

namespace {

// ---------------------------------------------------------------------------------------------
// E0: resize (cv2.INTER_LINEAR semantics for float images: pixel centres aligned, source index
// clamped, horizontal pass then vertical pass) + /255 + ImageNet mean/std (unet.py:159-161).
// in: [Hi][Wi][3] fp32 0..255   out: [Ho][Wo][3] fp32 normalised
// ---------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void prep_image_kernel(const TIn* __restrict__ in, int Hi, int Wi, float* __restrict__ out, int Ho, int Wo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Ho * Wo) return;
  const int y = idx / Wo, x = idx - y * Wo;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  float v[3];
  if (Hi == Ho && Wi == Wo) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (float)in[(size_t)idx * 3 + c];
  } else {
    const float sx = (float)Wi / (float)Wo, sy = (float)Hi / (float)Ho;
    float fx = ((float)x + 0.5f) * sx - 0.5f, fy = ((float)y + 0.5f) * sy - 0.5f;
    int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    fx -= (float)x0;
    fy -= (float)y0;
    if (x0 < 0) { x0 = 0; fx = 0.f; }
    if (x0 >= Wi - 1) { x0 = Wi - 1; fx = 0.f; }
    if (y0 < 0) { y0 = 0; fy = 0.f; }
    if (y0 >= Hi - 1) { y0 = Hi - 1; fy = 0.f; }
    const int x1 = min(x0 + 1, Wi - 1), y1 = min(y0 + 1, Hi - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = (float)in[((size_t)y0 * Wi + x0) * 3 + c] * (1.f - fx) + (float)in[((size_t)y0 * Wi + x1) * 3 + c] * fx;
      const float b = (float)in[((size_t)y1 * Wi + x0) * 3 + c] * (1.f - fx) + (float)in[((size_t)y1 * Wi + x1) * 3 + c] * fx;
      v[c] = a * (1.f - fy) + b * fy;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) out[(size_t)idx * 3 + c] = (v[c] / 255.f - mean[c]) / stdv[c];
}

// ---------------------------------------------------------------------------------------------
// First VGG layer: 3 -> 64 channels, 3x3, pad 1, bias, ReLU.  K = 27 is too thin for the tensor
// cores; fp32 CUDA-core direct convolution, one pixel x 64 channels per thread.
// w: [64][28] fp32 (27 taps ordered (ky, kx, c) + 1 pad), out: fp16 [H][W][64]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1_direct_kernel(const float* __restrict__ img, int H, int W,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           __half* __restrict__ out) {
  __shared__ __align__(16) float sw[64 * 28];
  __shared__ float sb[64];
  __shared__ float tile[18][18 * 3];
  for (int i = threadIdx.x; i < 64 * 28; i += 256) sw[i] = w[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < 18 * 18 * 3; i += 256) {
    const int ty = i / 54, r = i - ty * 54;
    const int tx = r / 3, c = r - tx * 3;
    const int gy = y0 + ty - 1, gx = x0 + tx - 1;
    tile[ty][r] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? img[((size_t)gy * W + gx) * 3 + c] : 0.f;
  }
  __syncthreads();
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  const int x = x0 + lx, y = y0 + ly;
  if (x >= W || y >= H) return;
  float in[28];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int j = 0; j < 9; ++j) in[ky * 9 + j] = tile[ly + ky][lx * 3 + j];
  in[27] = 0.f;
  __half* o = out + ((size_t)y * W + x) * 64;
#pragma unroll 1
  for (int c8 = 0; c8 < 64; c8 += 8) {
    uint4 pk;
    uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float acc[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float4* wr = reinterpret_cast<const float4*>(sw + (c8 + j + e) * 28);
        float a = sb[c8 + j + e];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const float4 wv = wr[q];
          a = fmaf(in[4 * q], wv.x, a);
          a = fmaf(in[4 * q + 1], wv.y, a);
          a = fmaf(in[4 * q + 2], wv.z, a);
          a = fmaf(in[4 * q + 3], wv.w, a);
        }
        acc[e] = fmaxf(a, 0.f);
      }
      const __half2 hv = __floats2half2_rn(acc[0], acc[1]);
      pw[j >> 1] = *reinterpret_cast<const uint32_t*>(&hv);
    }
    *reinterpret_cast<uint4*>(o + c8) = pk;
  }
}

// 2x2 / stride 2 max pool (floor mode), NHWC fp16, 8 channels per thread.
__global__ void maxpool2_kernel(const __half* __restrict__ in, int H, int W, int C, __half* __restrict__ out) {
  const int Ho = H >> 1, Wo = W >> 1, C8 = C >> 3;   // C8 is a power of two (C = 64 .. 512)
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)(Ho * Wo * C8)) return;
  const int sh = 31 - __clz(C8);
  const unsigned c8 = idx & (unsigned)(C8 - 1), p = idx >> sh;
  const int y = (int)(p / (unsigned)Wo), x = (int)(p - (unsigned)y * (unsigned)Wo);
  const uint4* src = reinterpret_cast<const uint4*>(in);
  const size_t base = ((size_t)(2 * y) * W + 2 * x) * C8 + c8;
  uint4 a = src[base], b = src[base + C8], c = src[base + (size_t)W * C8], d = src[base + (size_t)W * C8 + C8];
  __half2* ha = reinterpret_cast<__half2*>(&a);
  const __half2* hb = reinterpret_cast<const __half2*>(&b);
  const __half2* hc = reinterpret_cast<const __half2*>(&c);
  const __half2* hd = reinterpret_cast<const __half2*>(&d);
#pragma unroll
  for (int i = 0; i < 4; ++i) ha[i] = __hmax2(__hmax2(ha[i], hb[i]), __hmax2(hc[i], hd[i]));
  reinterpret_cast<uint4*>(out)[idx] = a;
}

// x2 bilinear upsample, align_corners=False (nn.Upsample in DecoderBlock, unet.py:19-20), NHWC fp16.
__global__ void upsample2_kernel(const __half* __restrict__ in, int H, int W, int C, __half* __restrict__ out) {
  const int Ho = 2 * H, Wo = 2 * W, C8 = C >> 3;   // C8 is a power of two (C = 64 .. 512)
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)(Ho * Wo * C8)) return;
  const int sh = 31 - __clz(C8);
  const unsigned c8 = idx & (unsigned)(C8 - 1), p = idx >> sh;
  const int y = (int)(p / (unsigned)Wo), x = (int)(p - (unsigned)y * (unsigned)Wo);
  const float sxf = fmaxf(((float)x + 0.5f) * 0.5f - 0.5f, 0.f), syf = fmaxf(((float)y + 0.5f) * 0.5f - 0.5f, 0.f);
  const int x0 = (int)sxf, y0 = (int)syf;
  const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  const float lx = sxf - (float)x0, ly = syf - (float)y0;
  const uint4* src = reinterpret_cast<const uint4*>(in);
  const uint4 v00 = src[((size_t)y0 * W + x0) * C8 + c8], v01 = src[((size_t)y0 * W + x1) * C8 + c8];
  const uint4 v10 = src[((size_t)y1 * W + x0) * C8 + c8], v11 = src[((size_t)y1 * W + x1) * C8 + c8];
  const __half2* a = reinterpret_cast<const __half2*>(&v00);
  const __half2* b = reinterpret_cast<const __half2*>(&v01);
  const __half2* c = reinterpret_cast<const __half2*>(&v10);
  const __half2* d = reinterpret_cast<const __half2*>(&v11);
  uint4 o;
  __half2* ho = reinterpret_cast<__half2*>(&o);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __half22float2(a[i]), fb = __half22float2(b[i]), fc = __half22float2(c[i]), fd = __half22float2(d[i]);
    ho[i] = __floats2half2_rn(w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x,
                              w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y);
  }
  reinterpret_cast<uint4*>(out)[idx] = o;
}

// ---------------------------------------------------------------------------------------------
// Heads (unet.py:47-50,177-188): 1x1 adaptation conv C_in -> C_out plus the 1x1 uncertainty conv
// (-> 1 channel), confidence = sigmoid(-u); optional per-pixel L2 normalisation of the descriptor
// (base_refiner.py:92-94) fused in.  fp16 activations in, fp32 out.
// w: [C_out + 1][C_in] fp32 (row C_out = uncertainty), b: [C_out + 1].
//
// A [pixels x C_in] x [C_in x (C_out+1)] GEMM on the warp-level tensor cores (mma.sync m16n8k16, fp16
// operands, fp32 accumulate): the op is ~1 GMAC in total and bound by the 78 MB it writes at level 0, far too
// small / too oddly shaped (33 and 129 output columns, K = 32) for a tcgen05 tile, and the CUDA-core version
// it replaces was shared-memory-bandwidth bound at 5-10 TFLOP/s.  A warp owns 32 pixels: A fragments come
// straight from the channels-last activation in global memory, the weights (converted to fp16) sit in shared
// memory with a conflict-free row stride, each pixel's outputs stay in the accumulator fragments of 4 lanes,
// so the squared norm is two shuffles, and rows leave as 8-byte stores (32 B contiguous per pixel and n-tile).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT n-tiles of 8 output columns: NT*8 >= C_out + 1 (5 for 32+1, 17 for 128+1)
template <int NT>
__global__ void __launch_bounds__(256) head_mma_kernel(const __half* __restrict__ x, int npix, int Cin, int Cout,
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ feat, float* __restrict__ conf,
                                                       int normalize) {
  extern __shared__ __align__(16) __half sw[];   // [NT*8][Cin + 8]
  const int ws = Cin + 8;
  const int nout = Cout + 1;
  for (int idx = threadIdx.x; idx < NT * 8 * (Cin / 4); idx += 256) {
    const int n = idx / (Cin / 4), q = idx - n * (Cin / 4);
    const float4 v = n < nout ? __ldg(reinterpret_cast<const float4*>(w + (size_t)n * Cin + q * 4))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
    __half2* d = reinterpret_cast<__half2*>(sw + n * ws + q * 4);
    d[0] = __floats2half2_rn(v.x, v.y);
    d[1] = __floats2half2_rn(v.z, v.w);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q2 = (lane & 3) * 2;
  const int groups = (npix + 31) / 32;
  for (int grp = blockIdx.x * 8 + warp; grp < groups; grp += gridDim.x * 8) {
    const int p0 = grp * 32;
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    const __half* xr[2][2];
    bool ok[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = p0 + mt * 16 + h * 8 + r;
        ok[mt][h] = p < npix;
        xr[mt][h] = x + (size_t)(ok[mt][h] ? p : 0) * Cin + q2;
      }
    auto load_a = [&](int k, uint32_t (&a)[2][4]) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        a[mt][0] = ok[mt][0] ? __ldg(reinterpret_cast<const unsigned*>(xr[mt][0] + k)) : 0u;
        a[mt][1] = ok[mt][1] ? __ldg(reinterpret_cast<const unsigned*>(xr[mt][1] + k)) : 0u;
        a[mt][2] = ok[mt][0] ? __ldg(reinterpret_cast<const unsigned*>(xr[mt][0] + k + 8)) : 0u;
        a[mt][3] = ok[mt][1] ? __ldg(reinterpret_cast<const unsigned*>(xr[mt][1] + k + 8)) : 0u;
      }
    };
    uint32_t a[2][4], an[2][4];
    load_a(0, a);
    for (int k = 0; k < Cin; k += 16) {
      if (k + 16 < Cin) load_a(k + 16, an);     // next K step's operand is in flight during this step's MMAs
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const __half* wr = sw + (nt * 8 + r) * ws + k + q2;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr), b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
        mma16816(acc[0][nt], a[0], b0, b1);
        mma16816(acc[1][nt], a[1], b0, b1);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[mt][i] = an[mt][i];
    }
    // bias, squared norm (a pixel's row lives in the 4 lanes of a quad), scale, store
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + q2;
        const float b0 = n < nout ? __ldg(b + n) : 0.f, b1 = n + 1 < nout ? __ldg(b + n + 1) : 0.f;
        acc[mt][nt][0] += b0; acc[mt][nt][1] += b1; acc[mt][nt][2] += b0; acc[mt][nt][3] += b1;
        if (n < Cout) {        // Cout is a multiple of 8: both columns are descriptor channels
          ss0 = fmaf(acc[mt][nt][0], acc[mt][nt][0], fmaf(acc[mt][nt][1], acc[mt][nt][1], ss0));
          ss1 = fmaf(acc[mt][nt][2], acc[mt][nt][2], fmaf(acc[mt][nt][3], acc[mt][nt][3], ss1));
        }
      }
      ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1); ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
      ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1); ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
      const float i0 = normalize ? 1.f / fmaxf(sqrtf(ss0), 1e-12f) : 1.f;
      const float i1 = normalize ? 1.f / fmaxf(sqrtf(ss1), 1e-12f) : 1.f;
      const int pa = p0 + mt * 16 + r, pb = pa + 8;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + q2;
        if (n < Cout) {
          if (pa < npix) *reinterpret_cast<float2*>(feat + (size_t)pa * Cout + n) = make_float2(acc[mt][nt][0] * i0, acc[mt][nt][1] * i0);
          if (pb < npix) *reinterpret_cast<float2*>(feat + (size_t)pb * Cout + n) = make_float2(acc[mt][nt][2] * i1, acc[mt][nt][3] * i1);
        } else if (n == Cout) {   // uncertainty column: confidence = sigmoid(-u)
          if (pa < npix) conf[pa] = 1.f / (1.f + expf(acc[mt][nt][0]));
          if (pb < npix) conf[pb] = 1.f / (1.f + expf(acc[mt][nt][2]));
        }
      }
    }
  }
}

template <int NT>
int launch_head(const PtkContext* ctx, const __half* x, long long npix, int Cin, int Cout, const float* w, const float* b,
                float* feat, float* conf, int normalize, cudaStream_t s) {
  const int smem = NT * 8 * (Cin + 8) * 2;
  static int configured = 0;
  if (configured < smem) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(head_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const long long groups = (npix + 31) / 32;
  long long blocks = (groups + 7) / 8;
  const long long cap = 2LL * ctx->num_sms;
  if (blocks > cap) blocks = cap;
  head_mma_kernel<NT><<<(unsigned)blocks, 256, smem, s>>>(x, (int)npix, Cin, Cout, w, b, feat, conf, normalize);
  return PTK_OK;
}

// ---------------------------------------------------------------------------------------------
// layer schedule
// ---------------------------------------------------------------------------------------------
constexpr int kNumConv = 20;   // 16 encoder + 4 decoder 3x3 convolutions
const int kEncBlocks[5][4] = {{64, 64, 0, 0}, {128, 128, 0, 0}, {256, 256, 256, 256}, {512, 512, 512, 512}, {512, 512, 512, 512}};
const int kEncCount[5] = {2, 2, 4, 4, 4};
const int kDec[4] = {64, 64, 64, 32};
const int kHeadScale[3] = {0, 2, 4};
const int kHeadDim[3] = {32, 128, 128};

}  // namespace

struct PtkExtractor {
  PtkContext* ctx;
  int H, W;                       // network input size
  PtkUnetWeights wts;
  float* img;                     // [H][W][3] normalised fp32
  __half* enc[5][4];              // conv outputs per block (the last one of each block is the skip feature)
  __half* pool[4];                // pooled input of blocks 1..4
  __half* up[4];                  // upsampled decoder inputs
  __half* dec[4];
  int eh[5], ew[5];               // spatial size of encoder block b
  int dh[4], dw[4];               // spatial size of decoder block i output
  void* arena;
  // optional per-launch timing (ptk_extractor_profile): events recorded after every launch
  cudaEvent_t* prof_ev;
  int prof_n;
};
#define PTK_MAX_LAUNCHES 48

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int ptk_extractor_create(PtkContext* ctx, const PtkUnetWeights* w, int32_t H, int32_t W,
                                    PtkExtractor** out) {
  PTK_REQUIRE(ctx && w && out, "null argument");
  PTK_REQUIRE(H >= 16 && W >= 16, "image must be at least 16x16");
  PtkExtractor* e = (PtkExtractor*)calloc(1, sizeof(PtkExtractor));
  e->ctx = ctx; e->H = H; e->W = W; e->wts = *w;
  e->eh[0] = H; e->ew[0] = W;
  for (int b = 1; b < 5; ++b) { e->eh[b] = e->eh[b - 1] / 2; e->ew[b] = e->ew[b - 1] / 2; }
  for (int i = 0; i < 4; ++i) { e->dh[i] = 2 * (i == 0 ? e->eh[4] : e->dh[i - 1]); e->dw[i] = 2 * (i == 0 ? e->ew[4] : e->dw[i - 1]); }
  // one arena for all activations
  size_t total = align_up((size_t)H * W * 3 * 4);
  for (int b = 0; b < 5; ++b)
    for (int i = 0; i < kEncCount[b]; ++i) total += align_up((size_t)e->eh[b] * e->ew[b] * kEncBlocks[b][i] * 2);
  for (int b = 1; b < 5; ++b) total += align_up((size_t)e->eh[b] * e->ew[b] * kEncBlocks[b - 1][kEncCount[b - 1] - 1] * 2);
  for (int i = 0; i < 4; ++i) {
    const int cprev = (i == 0) ? 512 : kDec[i - 1];
    total += align_up((size_t)e->dh[i] * e->dw[i] * cprev * 2) + align_up((size_t)e->dh[i] * e->dw[i] * kDec[i] * 2);
  }
  cudaError_t err = cudaMalloc(&e->arena, total);
  if (err != cudaSuccess) {
    ptk_set_error("extractor arena of %zu bytes: %s", total, cudaGetErrorString(err));
    free(e);
    return PTK_ERR_CUDA;
  }
  uint8_t* p = (uint8_t*)e->arena;
  auto take = [&](size_t bytes) { void* r = p; p += align_up(bytes); return r; };
  e->img = (float*)take((size_t)H * W * 3 * 4);
  for (int b = 0; b < 5; ++b)
    for (int i = 0; i < kEncCount[b]; ++i) e->enc[b][i] = (__half*)take((size_t)e->eh[b] * e->ew[b] * kEncBlocks[b][i] * 2);
  for (int b = 1; b < 5; ++b)
    e->pool[b - 1] = (__half*)take((size_t)e->eh[b] * e->ew[b] * kEncBlocks[b - 1][kEncCount[b - 1] - 1] * 2);
  for (int i = 0; i < 4; ++i) {
    const int cprev = (i == 0) ? 512 : kDec[i - 1];
    e->up[i] = (__half*)take((size_t)e->dh[i] * e->dw[i] * cprev * 2);
    e->dec[i] = (__half*)take((size_t)e->dh[i] * e->dw[i] * kDec[i] * 2);
  }
  *out = e;
  return PTK_OK;
}

extern "C" void ptk_extractor_destroy(PtkExtractor* e) {
  if (e == nullptr) return;
  if (e->arena) cudaFree(e->arena);
  free(e);
}

extern "C" int ptk_extractor_level_shape(const PtkExtractor* e, int32_t level, int32_t* C, int32_t* H, int32_t* W) {
  PTK_REQUIRE(e && level >= 0 && level < 3 && C && H && W, "bad argument");
  *C = kHeadDim[level];
  if (level == 0) { *H = e->dh[3]; *W = e->dw[3]; }
  else if (level == 1) { *H = e->dh[1]; *W = e->dw[1]; }
  else { *H = e->eh[4]; *W = e->ew[4]; }
  return PTK_OK;
}

// Debug / test access to intermediate activations (fp16 NHWC): kind 0 = encoder block output b,
// kind 1 = decoder block output i.
extern "C" int ptk_extractor_activation(const PtkExtractor* e, int32_t kind, int32_t index, const void** ptr, int32_t* C,
                                        int32_t* H, int32_t* W) {
  PTK_REQUIRE(e && ptr && C && H && W, "null argument");
  if (kind == 0 && index >= 0 && index < 5) {
    *ptr = e->enc[index][kEncCount[index] - 1]; *C = kEncBlocks[index][kEncCount[index] - 1]; *H = e->eh[index]; *W = e->ew[index];
    return PTK_OK;
  }
  if (kind == 1 && index >= 0 && index < 4) {
    *ptr = e->dec[index]; *C = kDec[index]; *H = e->dh[index]; *W = e->dw[index];
    return PTK_OK;
  }
  ptk_set_error("no such activation (%d, %d)", kind, index);
  return PTK_ERR_INVALID;
}

extern "C" int ptk_extractor_run(PtkExtractor* e, const void* image, int32_t img_dtype, int32_t img_h, int32_t img_w,
                                 float* const* feat, float* const* conf, int32_t normalize, void* stream) {
  PTK_REQUIRE(e && image && feat && conf, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int H = e->H, W = e->W;
  e->prof_n = 0;
  auto mark = [&]() { if (e->prof_ev != nullptr && e->prof_n < PTK_MAX_LAUNCHES) cudaEventRecord(e->prof_ev[e->prof_n++], s); };
  mark();
  PTK_REQUIRE(img_dtype == 0 || img_dtype == 1, "img_dtype must be 0 (fp32) or 1 (uint8)");
  if (img_dtype == 0) prep_image_kernel<float><<<(H * W + 255) / 256, 256, 0, s>>>((const float*)image, img_h, img_w, e->img, H, W);
  else prep_image_kernel<uint8_t><<<(H * W + 255) / 256, 256, 0, s>>>((const uint8_t*)image, img_h, img_w, e->img, H, W);
  mark();
  // ---- encoder (unet.py:163-167) ----
  int li = 0;
  conv1_direct_kernel<<<dim3((W + 15) / 16, (H + 15) / 16), 256, 0, s>>>(e->img, H, W, (const float*)e->wts.conv_w[0],
                                                                       e->wts.conv_b[0], e->enc[0][0]);
  PTK_CUDA_CHECK(cudaGetLastError());
  mark();
  li = 1;
  for (int b = 0; b < 5; ++b) {
    const int h = e->eh[b], w = e->ew[b];
    const __half* cur;
    int ccur;
    if (b == 0) {
      cur = e->enc[0][0];
      ccur = 64;
    } else {
      const int cprev = kEncBlocks[b - 1][kEncCount[b - 1] - 1];
      const long long n = (long long)h * w * (cprev / 8);
      maxpool2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(e->enc[b - 1][kEncCount[b - 1] - 1], e->eh[b - 1],
                                                                  e->ew[b - 1], cprev, e->pool[b - 1]);
      cur = e->pool[b - 1];
      ccur = cprev;
      mark();
    }
    for (int i = (b == 0 ? 1 : 0); i < kEncCount[b]; ++i) {
      const int rc = ptk_conv_f16(e->ctx, cur, ccur, nullptr, 0, h, w, h, w, 0, 0, e->wts.conv_w[li], e->wts.conv_b[li],
                                  kEncBlocks[b][i], 9, 1, e->enc[b][i], stream);
      if (rc != PTK_OK) return rc;
      mark();
      cur = e->enc[b][i];
      ccur = kEncBlocks[b][i];
      ++li;
    }
  }
  // ---- decoder (unet.py:169-173; DecoderBlock.forward :33-44) ----
  const __half* prev = e->enc[4][3];
  int cprev = 512, ph = e->eh[4], pw = e->ew[4];
  for (int i = 0; i < 4; ++i) {
    const int sb = 3 - i;   // skip feature comes from encoder block 3, 2, 1, 0
    const int cskip = kEncBlocks[sb][kEncCount[sb] - 1];
    const long long n = (long long)e->dh[i] * e->dw[i] * (cprev / 8);
    upsample2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(prev, ph, pw, cprev, e->up[i]);
    mark();
    const int rc = ptk_conv_f16(e->ctx, e->up[i], cprev, e->enc[sb][kEncCount[sb] - 1], cskip, e->dh[i], e->dw[i],
                                e->dh[i], e->dw[i], e->eh[sb], e->ew[sb], e->wts.conv_w[li], e->wts.conv_b[li], kDec[i], 9,
                                1, e->dec[i], stream);
    if (rc != PTK_OK) return rc;
    mark();
    prev = e->dec[i];
    cprev = kDec[i];
    ph = e->dh[i];
    pw = e->dw[i];
    ++li;
  }
  // ---- heads (unet.py:175-188): pre_features fine -> coarse = dec[3], dec[2], dec[1], dec[0], enc[4] ----
  for (int l = 0; l < 3; ++l) {
    const __half* src;
    int cin, h, w;
    if (kHeadScale[l] == 4) { src = e->enc[4][3]; cin = 512; h = e->eh[4]; w = e->ew[4]; }
    else { const int di = 3 - kHeadScale[l]; src = e->dec[di]; cin = kDec[di]; h = e->dh[di]; w = e->dw[di]; }
    const long long npix = (long long)h * w;
    int hrc;
    if (kHeadDim[l] == 32)
      hrc = launch_head<5>(e->ctx, src, npix, cin, 32, e->wts.head_w[l], e->wts.head_b[l], feat[l], conf[l], normalize, s);
    else
      hrc = launch_head<17>(e->ctx, src, npix, cin, 128, e->wts.head_w[l], e->wts.head_b[l], feat[l], conf[l], normalize, s);
    if (hrc != PTK_OK) return hrc;
    mark();
  }
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}

// Runs the plan once with a CUDA event after every launch and returns the per-launch durations
// (SYNCHRONISES; for benchmarks).  Launch order: prep, conv1, then per encoder block
// [pool] conv..., per decoder block upsample conv, 3 heads.  kinds[i]: 0 prep, 1 conv1 (direct),
// 2 pool, 3 tensor-core conv, 4 upsample, 5 head.  flops[i]: multiply-add count x 2 of launch i.
extern "C" int ptk_extractor_profile(PtkExtractor* e, const void* image, int32_t img_dtype, int32_t img_h, int32_t img_w,
                                     float* const* feat, float* const* conf, int32_t normalize, void* stream,
                                     int32_t max_n, float* ms, int32_t* kinds, double* flops, int32_t* n_out) {
  PTK_REQUIRE(e && ms && kinds && flops && n_out, "null argument");
  cudaEvent_t ev[PTK_MAX_LAUNCHES];
  for (int i = 0; i < PTK_MAX_LAUNCHES; ++i) PTK_CUDA_CHECK(cudaEventCreate(&ev[i]));
  e->prof_ev = ev;
  const int rc = ptk_extractor_run(e, image, img_dtype, img_h, img_w, feat, conf, normalize, stream);
  e->prof_ev = nullptr;
  if (rc != PTK_OK) return rc;
  PTK_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  const int n = e->prof_n - 1;
  // rebuild the launch list (same order as ptk_extractor_run)
  int k = 0;
  auto put = [&](int kind, double f) { if (k < max_n) { kinds[k] = kind; flops[k] = f; } ++k; };
  put(0, 0.0);
  put(1, 2.0 * e->H * e->W * 27.0 * 64.0);
  for (int b = 0; b < 5; ++b) {
    int ccur = (b == 0) ? 64 : kEncBlocks[b - 1][kEncCount[b - 1] - 1];
    if (b > 0) put(2, 0.0);
    for (int i = (b == 0 ? 1 : 0); i < kEncCount[b]; ++i) {
      put(3, 2.0 * e->eh[b] * e->ew[b] * 9.0 * ccur * kEncBlocks[b][i]);
      ccur = kEncBlocks[b][i];
    }
  }
  int cprev = 512;
  for (int i = 0; i < 4; ++i) {
    const int sb = 3 - i;
    put(4, 0.0);
    put(3, 2.0 * e->dh[i] * e->dw[i] * 9.0 * (cprev + kEncBlocks[sb][kEncCount[sb] - 1]) * kDec[i]);
    cprev = kDec[i];
  }
  for (int l = 0; l < 3; ++l) {
    const int cin = (kHeadScale[l] == 4) ? 512 : kDec[3 - kHeadScale[l]];
    const int h = (kHeadScale[l] == 4) ? e->eh[4] : e->dh[3 - kHeadScale[l]];
    const int w = (kHeadScale[l] == 4) ? e->ew[4] : e->dw[3 - kHeadScale[l]];
    put(5, 2.0 * h * w * (double)cin * (kHeadDim[l] + 1));
  }
  for (int i = 0; i < n && i < max_n; ++i) PTK_CUDA_CHECK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
  *n_out = n < k ? n : k;
  for (int i = 0; i < PTK_MAX_LAUNCHES; ++i) cudaEventDestroy(ev[i]);
  return PTK_OK;
}
