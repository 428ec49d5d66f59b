// The PixLoc UNet feature extractor as a native plan: the small kernels around the tcgen05
// convolutions of ptk_conv.cu (image prep, first layer and heads on mma.sync, upsample) and the layer schedule.
//
// Replaces UNet._forward (reference pixloc/pixloc/pixlib/models/unet.py:158-190) with the PixLoc
// configuration (pixlib/configs/train_pixloc_megadepth.yaml:22-31: vgg19 encoder, decoder
// [64,64,64,32], heads at scales 0/2/4 with 32/128/128 channels + uncertainty), and the
// pre-processing of PixTrackFeatureExtractor.__call__ (pixtrack/localization/feature_extractor.py:34-59).
// Activations are channels-last fp16 (fp32 accumulation everywhere); outputs are fp32.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "ptk_common.cuh"


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
// First VGG layer: 3 -> 64 channels, 3x3, pad 1, bias, ReLU.  K = 27 (padded to 32) is too thin for a
// tcgen05 tile; it runs on the warp-level tensor cores (mma.sync m16n8k16, fp16 operands, fp32 accumulate).
// A block owns a 16 x 16 pixel tile: the fp32 image tile + halo goes to shared memory as fp16, a warp takes two
// rows of 16 pixels (two m16 tiles) and builds its im2col A fragments by gathering 16-bit values from the tile
// (column j = ky*9 + kx*3 + c of pixel x is tile[y + ky][3*x + j - 9*ky]); the 64 x 32 weight matrix lives in
// registers as B fragments for the whole kernel (persistent blocks).  The CUDA-core version this replaces was
// shared-memory-bandwidth bound at 21 TFLOP/s (95 us at 1024 x 576).
// w: [64][28] fp32 (27 taps ordered (ky, kx, c) + 1 pad), out: fp16 [H][W][64]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 16-byte asynchronous global -> shared copy (LDGSTS); !valid zero-fills the destination (src must still be a mapped address)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

constexpr int kC1Stride = 18 * 3 + 2;   // halfs per tile row (56: keeps rows 4-byte aligned)

// ---------------------------------------------------------------------------------------------
// Quad transpose of mma.sync accumulator fragments.  In an m16n8 C fragment lane (r, q) holds columns n*8 + 2q, 2q+1
// of n-tile n: written as they are, a pixel's channels go out as 4- or 8-byte pieces from 4 lanes per n-tile, i.e. many
// small store requests (the store path, not DRAM, bounded conv1 and the level-0 head).  Two xor-shuffle rounds inside
// the quad regroup four n-tile items v[h][m] (n = 2h + m, P registers each) so that lane q ends up with the item of
// n-tile q from all four lanes: o[q'] = what lane q' held for n-tile q -- 32 contiguous bytes per lane, a full line per quad.
// ---------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void quad_transpose(const uint32_t (&v)[2][2][P], int q, uint32_t (&o)[4][P]) {
  const bool b0 = (q & 1) != 0, b1 = (q & 2) != 0;
  uint32_t x0[2][P], x1[2][P];            // items with m == b0: own / from lane q ^ 1
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int p = 0; p < P; ++p) {
      x0[h][p] = b0 ? v[h][1][p] : v[h][0][p];
      x1[h][p] = __shfl_xor_sync(0xffffffffu, b0 ? v[h][0][p] : v[h][1][p], 1);
    }
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const uint32_t k0 = b1 ? x0[1][p] : x0[0][p], k1 = b1 ? x1[1][p] : x1[0][p];                    // h == b1: from q, q ^ 1
    const uint32_t y0 = __shfl_xor_sync(0xffffffffu, b1 ? x0[0][p] : x0[1][p], 2);                // from q ^ 2
    const uint32_t y1 = __shfl_xor_sync(0xffffffffu, b1 ? x1[0][p] : x1[1][p], 2);                // from q ^ 3
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) o[sl][p] = sl == q ? k0 : (sl == (q ^ 1) ? k1 : (sl == (q ^ 2) ? y0 : y1));
  }
}
__device__ __forceinline__ void st_global_256(void* dst, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                              uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}

__global__ void __launch_bounds__(256, 2) conv1_mma_kernel(const float* __restrict__ img, int H, int W,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        __half* __restrict__ out, int tiles_x, int tiles) {
  __shared__ __half tile[18 * kC1Stride];
  __shared__ __half sw[64 * 40];           // [n][k] fp16, k padded 27 -> 32 (+8 to spread the banks)
  __shared__ float sbias[64];
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {
    const int n = i >> 5, k = i & 31;
    sw[n * 40 + k] = __float2half_rn(k < 27 ? w[n * 28 + k] : 0.f);
  }
  if (threadIdx.x < 64) sbias[threadIdx.x] = bias[threadIdx.x];
  ptk_pdl_wait();                          // the weights are constants; the image comes from the previous kernel
  ptk_pdl_trigger();
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q2 = (lane & 3) * 2;
  // B fragments of the whole weight matrix: 8 n-tiles x 2 k-steps
  uint32_t bf[8][2][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const __half* wr = sw + (nt * 8 + r) * 40 + kk * 16 + q2;
      bf[nt][kk][0] = *reinterpret_cast<const uint32_t*>(wr);
      bf[nt][kk][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
    }
  }
  // im2col offsets of this lane's 8 A columns: j = kk*16 + q2 + {0, 1, 8, 9}
  int off[2][4];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = kk * 16 + q2 + (e & 1) + (e >> 1) * 8;
      off[kk][e] = j < 27 ? (j / 9) * kC1Stride + (j % 9) : -1;
    }
  const __half hz = __float2half_rn(0.f);

  // The image tile of the NEXT tile is fetched into registers (4 floats per thread) before the MMAs of the current one,
  // so its DRAM latency hides behind them instead of sitting between two barriers.
  float nxt[4];
  auto fetch = [&](int t) {
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = threadIdx.x + j * 256;
      const int yy = i / 54, rr = i - yy * 54;
      const int xx = rr / 3, c = rr - xx * 3;
      const int gy = ty * 16 + yy - 1, gx = tx * 16 + xx - 1;
      nxt[j] = (i < 18 * 54 && gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + ((size_t)gy * W + gx) * 3 + c) : 0.f;
    }
  };
  if ((int)blockIdx.x < tiles) fetch(blockIdx.x);
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int x0 = tx * 16, y0 = ty * 16;
    __syncthreads();   // previous tile fully consumed
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = threadIdx.x + j * 256;
      if (i < 18 * 54) {
        const int yy = i / 54, rr = i - yy * 54;
        tile[yy * kC1Stride + rr] = __float2half_rn(nxt[j]);
      }
    }
    __syncthreads();
    if (t + (int)gridDim.x < tiles) fetch(t + gridDim.x);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int ly = warp * 2 + mt;                 // tile row of this m16 tile
      const __half* base0 = tile + ly * kC1Stride + r * 3;          // pixel r
      const __half* base1 = base0 + 8 * 3;                          // pixel r + 8
      uint32_t a[2][4];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        __half v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = off[kk][e] >= 0 ? base0[off[kk][e]] : hz;       // row r:     cols (q2, q2+1, q2+8, q2+9)
          v[4 + e] = off[kk][e] >= 0 ? base1[off[kk][e]] : hz;   // row r + 8
        }
        const __half2 p0 = __halves2half2(v[0], v[1]), p1 = __halves2half2(v[4], v[5]);
        const __half2 p2 = __halves2half2(v[2], v[3]), p3 = __halves2half2(v[6], v[7]);
        a[kk][0] = *reinterpret_cast<const uint32_t*>(&p0);
        a[kk][1] = *reinterpret_cast<const uint32_t*>(&p1);
        a[kk][2] = *reinterpret_cast<const uint32_t*>(&p2);
        a[kk][3] = *reinterpret_cast<const uint32_t*>(&p3);
      }
      const int y = y0 + ly;
      const int xa = x0 + r, xb = xa + 8;
      // (A quad transpose to 32-byte stores was measured here and changed nothing: this kernel is bound by how few warps
      // fit an SM, not by its store requests.)
      __half* oa = out + ((size_t)y * W + xa) * 64 + q2;
      __half* ob = out + ((size_t)y * W + xb) * 64 + q2;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 bia = *reinterpret_cast<const float2*>(sbias + nt * 8 + q2);   // (registers are what limits the CTAs per SM)
        float c[4] = {bia.x, bia.y, bia.x, bia.y};
        mma16816(c, a[0], bf[nt][0][0], bf[nt][0][1]);
        mma16816(c, a[1], bf[nt][1][0], bf[nt][1][1]);
        if (y < H) {
          if (xa < W) *reinterpret_cast<__half2*>(oa + nt * 8) = __floats2half2_rn(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f));
          if (xb < W) *reinterpret_cast<__half2*>(ob + nt * 8) = __floats2half2_rn(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f));
        }
      }
    }
  }
}

// (the 2x2 max pools of the encoder are written by the epilogue of the convolution in front of them, ptk_conv.cu)

// x2 bilinear upsample, align_corners=False (nn.Upsample in DecoderBlock, unet.py:19-20), NHWC fp16.
// For scale 2 the source position of output 2i is i - 0.25 and of 2i+1 is i + 0.25 (clamped at the borders), so
// every output is a 0.75 / 0.25 blend of input i with its left/upper or right/lower neighbour.  A thread produces the
// 2x2 output block of input pixel (y, x) for 8 channels from the 3x3 input neighbourhood (9 loads for 4 outputs,
// separable blend) instead of 4 loads and a general bilinear evaluation per output.
__global__ void upsample2_kernel(const __half* __restrict__ in, int H, int W, int C, __half* __restrict__ out) {
  ptk_pdl_wait();
  ptk_pdl_trigger();
  const int C8 = C >> 3;                           // power of two (C = 64 .. 512)
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)(H * W * C8)) return;
  const int sh = 31 - __clz(C8);
  const unsigned c8 = idx & (unsigned)(C8 - 1), p = idx >> sh;
  const int y = (int)(p / (unsigned)W), x = (int)(p - (unsigned)y * (unsigned)W);
  const int xm = max(x - 1, 0), xp = min(x + 1, W - 1), ym = max(y - 1, 0), yp = min(y + 1, H - 1);
  const uint4* src = reinterpret_cast<const uint4*>(in);
  const int rows[3] = {ym, y, yp}, cols[3] = {xm, x, xp};
  float2 v[3][3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint4 raw = __ldg(src + ((size_t)rows[r] * W + cols[c]) * C8 + c8);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[r][c][e] = __half22float2(h[e]);
    }
  // horizontal blend: left output = 0.25 * west + 0.75 * centre, right output = 0.75 * centre + 0.25 * east
  float2 hl[3][4], hr[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hl[r][e] = make_float2(0.25f * v[r][0][e].x + 0.75f * v[r][1][e].x, 0.25f * v[r][0][e].y + 0.75f * v[r][1][e].y);
      hr[r][e] = make_float2(0.75f * v[r][1][e].x + 0.25f * v[r][2][e].x, 0.75f * v[r][1][e].y + 0.25f * v[r][2][e].y);
    }
  uint4 o[2][2];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __half2* o00 = reinterpret_cast<__half2*>(&o[0][0]);
    __half2* o01 = reinterpret_cast<__half2*>(&o[0][1]);
    __half2* o10 = reinterpret_cast<__half2*>(&o[1][0]);
    __half2* o11 = reinterpret_cast<__half2*>(&o[1][1]);
    o00[e] = __floats2half2_rn(0.25f * hl[0][e].x + 0.75f * hl[1][e].x, 0.25f * hl[0][e].y + 0.75f * hl[1][e].y);
    o01[e] = __floats2half2_rn(0.25f * hr[0][e].x + 0.75f * hr[1][e].x, 0.25f * hr[0][e].y + 0.75f * hr[1][e].y);
    o10[e] = __floats2half2_rn(0.75f * hl[1][e].x + 0.25f * hl[2][e].x, 0.75f * hl[1][e].y + 0.25f * hl[2][e].y);
    o11[e] = __floats2half2_rn(0.75f * hr[1][e].x + 0.25f * hr[2][e].x, 0.75f * hr[1][e].y + 0.25f * hr[2][e].y);
  }
  uint4* dst = reinterpret_cast<uint4*>(out);
  const int Wo = 2 * W;
  const size_t base = ((size_t)(2 * y) * Wo + 2 * x) * C8 + c8;
  dst[base] = o[0][0];
  dst[base + C8] = o[0][1];
  dst[base + (size_t)Wo * C8] = o[1][0];
  dst[base + (size_t)Wo * C8 + C8] = o[1][1];
}

// ---------------------------------------------------------------------------------------------
// Heads (unet.py:47-50,177-188): 1x1 adaptation conv C_in -> C_out plus the 1x1 uncertainty conv
// (-> 1 channel), confidence = sigmoid(-u); optional per-pixel L2 normalisation of the descriptor
// (base_refiner.py:92-94) fused in.  fp16 activations in, fp32 out.
// w: [C_out + 1][C_in] fp16 (row C_out = uncertainty), b: [C_out + 1] fp32.
//
// A [pixels x C_in] x [C_in x (C_out+1)] GEMM on the warp-level tensor cores (mma.sync m16n8k16, fp16
// operands, fp32 accumulate): the op is ~1 GMAC in total and bound by the 78 MB it writes at level 0, far too
// small / too oddly shaped (33 and 129 output columns, K = 32) for a tcgen05 tile, and the CUDA-core version
// it replaces was shared-memory-bandwidth bound at 5-10 TFLOP/s.  A warp owns 32 pixels: A fragments come
// straight from the channels-last activation in global memory, the fp16 weights sit in shared
// memory with a conflict-free row stride, each pixel's outputs stay in the accumulator fragments of 4 lanes,
// so the squared norm is two shuffles, and rows leave as 8-byte stores (32 B contiguous per pixel and n-tile).
// ---------------------------------------------------------------------------------------------
// NT n-tiles of 8 output columns: NT*8 >= C_out + 1 (5 for 32+1, 17 for 128+1)
template <int NT>
__global__ void __launch_bounds__(256, NT <= 5 ? 3 : 1) head_mma_kernel(const __half* __restrict__ x, int npix, int Cin, int Cout,
                                                       const __half* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ feat, float* __restrict__ conf,
                                                       int normalize, int wpb) {
  extern __shared__ __align__(16) __half sw[];   // [NT*8][Cin + 8] weights, then 2 x [wpb * 32 pixels][Cin + 8] activations
  const int ws = Cin + 8;
  const int nout = Cout + 1;
  const int c8 = Cin >> 3;                        // 16-byte pieces per row
  const int c8_log2 = 31 - __clz(c8);             // Cin is 32, 64 or 512: a power of two
  // Staging goes through cp.async: every thread has all of its 16-byte pieces in flight at once.  (The plain
  // load -> store loop this replaces kept ONE load in flight per thread -- 34 dependent L2 round trips for the 132 KB
  // weight matrix of the coarsest head, 29 us for 0.3 GFLOP.)
  for (int idx = threadIdx.x; idx < NT * 8 * c8; idx += 256) {
    const int n = idx >> c8_log2, q = idx & (c8 - 1);
    cp_async16(sw + n * ws + q * 8, reinterpret_cast<const uint4*>(w + (size_t)(n < nout ? n : 0) * Cin) + q, n < nout);
  }
  cp_async_commit();
  ptk_pdl_wait();                                 // the weights above are constants; the activations come from the previous kernel
  ptk_pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q2 = (lane & 3) * 2;
  const int groups = (npix + 31) / 32;
  __half* sa0 = sw + NT * 8 * ws;                 // two activation buffers: round i + 1 is staged while round i is computed
  const int sa_halfs = wpb * 32 * ws;
  const int stride = gridDim.x * wpb;
  auto stage = [&](int base, int buf) {
    __half* sa = sa0 + buf * sa_halfs;
    const int pbase = base * 32;
    for (int idx = threadIdx.x; idx < wpb * 32 * c8; idx += 256) {
      const int pl = idx >> c8_log2, q = idx & (c8 - 1);
      const int p = pbase + pl;
      cp_async16(sa + pl * ws + q * 8, reinterpret_cast<const uint4*>(x + (size_t)(p < npix ? p : 0) * Cin) + q, p < npix);
    }
    cp_async_commit();
  };
  int buf = 0;
  if ((int)blockIdx.x * wpb < groups) stage(blockIdx.x * wpb, 0);
  // wpb warps of the block take a group of 32 pixels each per round; all 8 warps stage the activations
  for (int base = blockIdx.x * wpb; base < groups; base += stride, buf ^= 1) {
    if (base + stride < groups) {
      stage(base + stride, buf ^ 1);
      cp_async_wait<1>();                          // everything but the round just issued has landed (weights included)
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __half* sa = sa0 + buf * sa_halfs;
    const int grp = base + warp;
    if (warp < wpb && grp < groups) {
    const int p0 = grp * 32;
    const __half* ta = sa + warp * 32 * ws;
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    for (int k = 0; k < Cin; k += 16) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const __half* ar = ta + (mt * 16 + r) * ws + k + q2;
        a[mt][0] = *reinterpret_cast<const uint32_t*>(ar);
        a[mt][1] = *reinterpret_cast<const uint32_t*>(ar + 8 * ws);
        a[mt][2] = *reinterpret_cast<const uint32_t*>(ar + 8);
        a[mt][3] = *reinterpret_cast<const uint32_t*>(ar + 8 * ws + 8);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const __half* wr = sw + (nt * 8 + r) * ws + k + q2;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr), b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
        mma16816(acc[0][nt], a[0], b0, b1);
        mma16816(acc[1][nt], a[1], b0, b1);
      }
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
      const int q = lane & 3;
      // descriptor channels: groups of four n-tiles (32 channels) go through the quad transpose, lane q then writes
      // channels [32 g + 8 q, + 8) of its pixel as one 32-byte store (Cout is 32 or 128)
#pragma unroll
      for (int g = 0; g < (NT - 1) / 4; ++g) {
        uint32_t va[2][2][2], vb[2][2][2], oa[4][2], ob[4][2];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int nt = g * 4 + t;
          va[t >> 1][t & 1][0] = __float_as_uint(acc[mt][nt][0] * i0);
          va[t >> 1][t & 1][1] = __float_as_uint(acc[mt][nt][1] * i0);
          vb[t >> 1][t & 1][0] = __float_as_uint(acc[mt][nt][2] * i1);
          vb[t >> 1][t & 1][1] = __float_as_uint(acc[mt][nt][3] * i1);
        }
        quad_transpose<2>(va, q, oa);
        quad_transpose<2>(vb, q, ob);
        if (pa < npix)
          st_global_256(feat + (size_t)pa * Cout + g * 32 + 8 * q, oa[0][0], oa[0][1], oa[1][0], oa[1][1], oa[2][0], oa[2][1], oa[3][0],
                        oa[3][1]);
        if (pb < npix)
          st_global_256(feat + (size_t)pb * Cout + g * 32 + 8 * q, ob[0][0], ob[0][1], ob[1][0], ob[1][1], ob[2][0], ob[2][1], ob[3][0],
                        ob[3][1]);
      }
      if (q2 == 0) {   // uncertainty column (n-tile NT - 1, column 0): confidence = sigmoid(-u)
        if (pa < npix) conf[pa] = 1.f / (1.f + expf(acc[mt][NT - 1][0]));
        if (pb < npix) conf[pb] = 1.f / (1.f + expf(acc[mt][NT - 1][2]));
      }
    }
    }   // compute warps
    __syncthreads();   // this round's buffer is free before the round after next is staged into it
  }
}

template <int NT>
int launch_head(const PtkContext* ctx, const __half* x, long long npix, int Cin, int Cout, const __half* w, const float* b,
                float* feat, float* conf, int normalize, cudaStream_t s) {
  const long long groups = (npix + 31) / 32;
  long long wpb = (groups + ctx->num_sms - 1) / ctx->num_sms;
  wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);
  const int smem = (NT * 8 + 2 * (int)wpb * 32) * (Cin + 8) * 2;   // weights + two activation buffers
  static int configured = 0;
  if (configured < smem) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(head_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // the level-0 head (NT = 5) is compiled for three CTAs per SM (80 registers): ask for the shared memory to match
    PTK_CUDA_CHECK(cudaFuncSetAttribute(head_mma_kernel<NT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = smem;
  }
  long long blocks = (groups + wpb - 1) / wpb;
  const long long cap = (NT <= 5 ? 3LL : 2LL) * ctx->num_sms;
  if (blocks > cap) blocks = cap;
  PTK_CUDA_CHECK(ptk_launch_pdl(head_mma_kernel<NT>, dim3((unsigned)blocks), dim3(256), (size_t)smem, s, dim3(1, 1, 1), x, (int)npix, Cin, Cout,
                                w, b, feat, conf, normalize, (int)wpb));
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
  int head0_fused;   // last run: the level-0 head ran inside the last decoder convolution (one launch fewer)
  PtkHeadConst* head0;   // host copy of the level-0 head's weights as fp32 (they travel as kernel parameters when fused)
  // the coarse heads run next to the decoder on a side stream (forked / joined with events, so the
  // plan is still one stream-ordered unit for the caller and can be captured in a CUDA graph)
  cudaStream_t side;
  cudaEvent_t ev_fork[2], ev_join;
  // CUDA graphs of the whole plan, one per distinct (image, outputs) binding: the second call with a binding
  // captures the launches on `cap`, later calls replay the graph (one launch instead of ~30 + tensor-map encodes)
  struct PlanGraph {
    const void* image;
    int img_dtype, img_h, img_w, normalize, seen;
    float* feat[3];
    float* conf[3];
    cudaGraphExec_t exec;
    unsigned long long last_use;
  } graphs[8];
  int n_graphs;
  unsigned long long use_clock;
  cudaStream_t cap;
};
#define PTK_MAX_LAUNCHES 48

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int ptk_extractor_create(PtkContext* ctx, const PtkUnetWeights* w, int32_t H, int32_t W,
                                    PtkExtractor** out) {
  PTK_REQUIRE(ctx && w && out, "null argument");
  PTK_REQUIRE(H >= 16 && W >= 16, "image must be at least 16x16");
  PtkDeviceGuard guard(ctx->device);
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
  {   // level-0 head weights -> host fp32 (fp16 [33][32] + fp32 [33] on the device)
    e->head0 = (PtkHeadConst*)calloc(1, sizeof(PtkHeadConst));
    static_assert(sizeof(PtkHeadConst) < 8192, "kernel parameter budget");
    __half hw[33 * 32];
    PTK_CUDA_CHECK(cudaMemcpy(hw, e->wts.head_w[0], sizeof(hw), cudaMemcpyDeviceToHost));
    PTK_CUDA_CHECK(cudaMemcpy(e->head0->b, e->wts.head_b[0], 33 * sizeof(float), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 33 * 32; ++i) e->head0->w[i] = __half2float(hw[i]);
  }
  PTK_CUDA_CHECK(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) PTK_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_fork[i], cudaEventDisableTiming));
  PTK_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  *out = e;
  return PTK_OK;
}

extern "C" void ptk_extractor_destroy(PtkExtractor* e) {
  if (e == nullptr) return;
  if (e->side) cudaStreamDestroy(e->side);
  if (e->cap) cudaStreamDestroy(e->cap);
  for (int i = 0; i < e->n_graphs; ++i) if (e->graphs[i].exec) cudaGraphExecDestroy(e->graphs[i].exec);
  for (int i = 0; i < 2; ++i) if (e->ev_fork[i]) cudaEventDestroy(e->ev_fork[i]);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->arena) cudaFree(e->arena);
  free(e->head0);
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

// Kernel launches of the last run of this plan (28 with the level-0 head as its own launch, 27 with it fused).
extern "C" int ptk_extractor_launch_count(const PtkExtractor* e, int32_t* n) {
  PTK_REQUIRE(e && n, "null argument");
  *n = 28 - (e->head0_fused ? 1 : 0);
  return PTK_OK;
}

static int run_plan(PtkExtractor* e, const void* image, int32_t img_dtype, int32_t img_h, int32_t img_w,
                    float* const* feat, float* const* conf, int32_t normalize, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int H = e->H, W = e->W;
  e->prof_n = 0;
  auto mark = [&]() { if (e->prof_ev != nullptr && e->prof_n < PTK_MAX_LAUNCHES) cudaEventRecord(e->prof_ev[e->prof_n++], s); };
  mark();
  if (img_dtype == 0) prep_image_kernel<float><<<(H * W + 255) / 256, 256, 0, s>>>((const float*)image, img_h, img_w, e->img, H, W);
  else prep_image_kernel<uint8_t><<<(H * W + 255) / 256, 256, 0, s>>>((const uint8_t*)image, img_h, img_w, e->img, H, W);
  mark();
  // ---- encoder (unet.py:163-167) ----
  int li = 0;
  {
    const int tiles_x = (W + 15) / 16, tiles = tiles_x * ((H + 15) / 16);
    const int grid = tiles < 4 * e->ctx->num_sms ? tiles : 4 * e->ctx->num_sms;
    PTK_CUDA_CHECK(ptk_launch_pdl(conv1_mma_kernel, dim3(grid), dim3(256), 0, s, dim3(1, 1, 1), (const float*)e->img, H, W,
                                  (const float*)e->wts.conv_w[0], (const float*)e->wts.conv_b[0], e->enc[0][0], tiles_x, tiles));
  }
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
    } else {   // the 2x2 max pool of the previous block was written by that block's last convolution
      cur = e->pool[b - 1];
      ccur = kEncBlocks[b - 1][kEncCount[b - 1] - 1];
    }
    for (int i = (b == 0 ? 1 : 0); i < kEncCount[b]; ++i) {
      void* pool_out = (i == kEncCount[b] - 1 && b < 4) ? (void*)e->pool[b] : nullptr;
      const int rc = ptk_conv_f16_pool(e->ctx, cur, ccur, nullptr, 0, h, w, h, w, 0, 0, e->wts.conv_w[li], e->wts.conv_b[li],
                                       kEncBlocks[b][i], 9, 1, e->enc[b][i], pool_out, stream);
      if (rc != PTK_OK) return rc;
      mark();
      cur = e->enc[b][i];
      ccur = kEncBlocks[b][i];
      ++li;
    }
  }
  // ---- heads (unet.py:175-188): pre_features fine -> coarse = dec[3], dec[2], dec[1], dec[0], enc[4] ----
  auto run_head = [&](int l, cudaStream_t hs) -> int {
    const __half* src;
    int cin, h, w;
    if (kHeadScale[l] == 4) { src = e->enc[4][3]; cin = 512; h = e->eh[4]; w = e->ew[4]; }
    else { const int di = 3 - kHeadScale[l]; src = e->dec[di]; cin = kDec[di]; h = e->dh[di]; w = e->dw[di]; }
    const long long npix = (long long)h * w;
    if (kHeadDim[l] == 32)
      return launch_head<5>(e->ctx, src, npix, cin, 32, (const __half*)e->wts.head_w[l], e->wts.head_b[l], feat[l], conf[l], normalize, hs);
    return launch_head<17>(e->ctx, src, npix, cin, 128, (const __half*)e->wts.head_w[l], e->wts.head_b[l], feat[l], conf[l], normalize, hs);
  };
  // The two coarse heads only need the encoder output / the second decoder block: they run on the plan's side
  // stream next to the remaining decoder convolutions (which leave SMs idle at these map sizes).  With per-launch
  // profiling on, everything stays on the caller's stream.
  const bool fork = e->prof_ev == nullptr;
  // PTK_FUSE_HEAD=1 runs the level-0 head inside the epilogue of the last decoder convolution (27 launches instead of 28).
  // Off by default: measured on the C2 frame loop it LOSES 3 % (568 against 588 frames/s, one extraction 0.702 against
  // 0.681 ms) -- that convolution (C_out = 32) is bound by the tensor core's operand fetch from shared memory, its MMA warp
  // shares the SM's issue slots with the epilogue warps, and 2 100 more instructions per pixel row there cost more than
  // the 34 us launch they replace.  (A first version with the weights in shared memory lost 5 %.)
  static int fuse_mode = -1;
  if (fuse_mode < 0) fuse_mode = getenv("PTK_FUSE_HEAD") ? atoi(getenv("PTK_FUSE_HEAD")) : 0;
  const bool fuse_head0 = fuse_mode != 0 && e->prof_ev == nullptr && kHeadScale[0] == 0 && kHeadDim[0] == 32 && kDec[3] == 32;
  int head0_fused = 0;
  if (fork) {
    PTK_CUDA_CHECK(cudaEventRecord(e->ev_fork[0], s));
    PTK_CUDA_CHECK(cudaStreamWaitEvent(e->side, e->ev_fork[0], 0));
    const int hrc = run_head(2, e->side);
    if (hrc != PTK_OK) return hrc;
  }
  // ---- decoder (unet.py:169-173; DecoderBlock.forward :33-44) ----
  const __half* prev = e->enc[4][3];
  int cprev = 512, ph = e->eh[4], pw = e->ew[4];
  for (int i = 0; i < 4; ++i) {
    const int sb = 3 - i;   // skip feature comes from encoder block 3, 2, 1, 0
    const int cskip = kEncBlocks[sb][kEncCount[sb] - 1];
    const long long n = (long long)ph * pw * (cprev / 8);      // one thread per input pixel and 8 channels
    PTK_CUDA_CHECK(ptk_launch_pdl(upsample2_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, dim3(1, 1, 1), prev, ph, pw, cprev,
                                  e->up[i]));
    mark();
    int rc;
    if (i == 3 && fuse_head0) {   // the level-0 head rides in the epilogue of the last decoder convolution when it can
      PtkHeadConst& hf = *e->head0;
      hf.feat = feat[0];
      hf.conf = conf[0];
      hf.normalize = normalize;
      rc = ptk_conv_f16_head(e->ctx, e->up[i], cprev, e->enc[sb][kEncCount[sb] - 1], cskip, e->dh[i], e->dw[i], e->dh[i], e->dw[i],
                             e->eh[sb], e->ew[sb], e->wts.conv_w[li], e->wts.conv_b[li], kDec[i], 9, 1, e->dec[i], stream, &hf,
                             &head0_fused);
    } else {
      rc = ptk_conv_f16(e->ctx, e->up[i], cprev, e->enc[sb][kEncCount[sb] - 1], cskip, e->dh[i], e->dw[i], e->dh[i], e->dw[i],
                        e->eh[sb], e->ew[sb], e->wts.conv_w[li], e->wts.conv_b[li], kDec[i], 9, 1, e->dec[i], stream);
    }
    if (rc != PTK_OK) return rc;
    mark();
    if (fork && i == 1) {   // dec[1] feeds the level-1 head
      PTK_CUDA_CHECK(cudaEventRecord(e->ev_fork[1], s));
      PTK_CUDA_CHECK(cudaStreamWaitEvent(e->side, e->ev_fork[1], 0));
      const int hrc = run_head(1, e->side);
      if (hrc != PTK_OK) return hrc;
    }
    prev = e->dec[i];
    cprev = kDec[i];
    ph = e->dh[i];
    pw = e->dw[i];
    ++li;
  }
  if (!head0_fused) {
    const int hrc = run_head(0, s);
    if (hrc != PTK_OK) return hrc;
    mark();
  }
  e->head0_fused = head0_fused;
  if (fork) {
    PTK_CUDA_CHECK(cudaEventRecord(e->ev_join, e->side));
    PTK_CUDA_CHECK(cudaStreamWaitEvent(s, e->ev_join, 0));
  } else {
    for (int l = 1; l < 3; ++l) {
      const int hrc = run_head(l, s);
      if (hrc != PTK_OK) return hrc;
      mark();
    }
  }
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}

extern "C" int ptk_extractor_run(PtkExtractor* e, const void* image, int32_t img_dtype, int32_t img_h, int32_t img_w,
                                 float* const* feat, float* const* conf, int32_t normalize, void* stream) {
  PTK_REQUIRE(e && image && feat && conf, "null argument");
  PTK_REQUIRE(img_dtype == 0 || img_dtype == 1, "img_dtype must be 0 (fp32) or 1 (uint8)");
  PtkDeviceGuard guard(e->ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  static int graph_mode = -1;   // PTK_PLAN_GRAPH=0: always launch kernel by kernel
  if (graph_mode < 0) graph_mode = getenv("PTK_PLAN_GRAPH") ? atoi(getenv("PTK_PLAN_GRAPH")) : 1;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (graph_mode == 0 || e->prof_ev != nullptr || cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();   // (a legacy default stream reports an error for the query while another stream captures)
    return run_plan(e, image, img_dtype, img_h, img_w, feat, conf, normalize, stream);
  }
  PtkExtractor::PlanGraph* g = nullptr;
  for (int i = 0; i < e->n_graphs && g == nullptr; ++i) {
    PtkExtractor::PlanGraph& c = e->graphs[i];
    bool same = c.image == image && c.img_dtype == img_dtype && c.img_h == img_h && c.img_w == img_w && c.normalize == normalize;
    for (int l = 0; l < 3 && same; ++l) same = c.feat[l] == feat[l] && c.conf[l] == conf[l];
    if (same) g = &c;
  }
  if (g != nullptr) g->last_use = ++e->use_clock;
  if (g != nullptr && g->exec != nullptr) {
    PTK_CUDA_CHECK(cudaGraphLaunch(g->exec, s));
    return PTK_OK;
  }
  if (g == nullptr) {   // first call with this binding: run directly (also configures every kernel it uses)
    if (e->n_graphs < 8) {
      g = &e->graphs[e->n_graphs++];
    } else {            // table full: the least recently used binding makes room (callers that allocate fresh outputs
      g = &e->graphs[0];   // every call only ever occupy key slots; their graphs are never built)
      for (int i = 1; i < 8; ++i) if (e->graphs[i].last_use < g->last_use) g = &e->graphs[i];
      if (g->exec != nullptr) {
        // a replay of this graph may still be in flight on some stream: retire it before destroying it
        PTK_CUDA_CHECK(cudaDeviceSynchronize());
        cudaGraphExecDestroy(g->exec);
      }
    }
    g->image = image; g->img_dtype = img_dtype; g->img_h = img_h; g->img_w = img_w; g->normalize = normalize;
    for (int l = 0; l < 3; ++l) { g->feat[l] = feat[l]; g->conf[l] = conf[l]; }
    g->seen = 1;
    g->exec = nullptr;
    g->last_use = ++e->use_clock;
    return run_plan(e, image, img_dtype, img_h, img_w, feat, conf, normalize, stream);
  }
  // second call: capture on the plan's own stream (the caller's may be the legacy default stream), then launch
  if (e->cap == nullptr) PTK_CUDA_CHECK(cudaStreamCreateWithFlags(&e->cap, cudaStreamNonBlocking));
  PTK_CUDA_CHECK(cudaStreamBeginCapture(e->cap, cudaStreamCaptureModeThreadLocal));
  const int rc = run_plan(e, image, img_dtype, img_h, img_w, feat, conf, normalize, (void*)e->cap);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(e->cap, &graph);
  if (rc != PTK_OK || ce != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc != PTK_OK) return rc;
    g->seen = 2;   // capture not possible here: keep launching directly
    g->image = nullptr;
    return run_plan(e, image, img_dtype, img_h, img_w, feat, conf, normalize, stream);
  }
  const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    g->exec = nullptr;
    g->image = nullptr;
    cudaGetLastError();
    return run_plan(e, image, img_dtype, img_h, img_w, feat, conf, normalize, stream);
  }
  PTK_CUDA_CHECK(cudaGraphLaunch(g->exec, s));
  return PTK_OK;
}

// Runs the plan once with a CUDA event after every launch and returns the per-launch durations
// (SYNCHRONISES; for benchmarks; everything on the caller's stream).  Launch order: prep, conv1, the encoder
// convolutions (pools fused), per decoder block upsample + conv, 3 heads.  kinds[i]: 0 prep, 1 conv1 (mma.sync),
// 3 tensor-core conv, 4 upsample, 5 head.  flops[i]: multiply-add count x 2 of launch i.
extern "C" int ptk_extractor_profile(PtkExtractor* e, const void* image, int32_t img_dtype, int32_t img_h, int32_t img_w,
                                     float* const* feat, float* const* conf, int32_t normalize, void* stream,
                                     int32_t max_n, float* ms, int32_t* kinds, double* flops, int32_t* n_out) {
  PTK_REQUIRE(e && ms && kinds && flops && n_out, "null argument");
  PtkDeviceGuard guard(e->ctx->device);
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
