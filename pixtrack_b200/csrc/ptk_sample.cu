// Sparse bilinear sampling of a dense map at N pixel positions.
//
// Replaces Interpolator.__call__ / interpolate_tensor(mode='linear')
// (reference pixloc/pixloc/pixlib/geometry/interpolation.py:57-141) as used by
// PoseTrackerRefiner.interp_sparse_observations
// (pixtrack/localization/pixloc_pose_refiners.py:349-351).
// Semantics kept: coordinates are normalised by (W-1, H-1), clamped to
// [-2, 2], un-normalised again (grid_sample, align_corners=True), corners
// outside the map contribute zero; mask = pad <= p <= size-1-pad; optional
// gradient = central difference of bilinear samples one pixel apart.
//
// Works for both layouts through element strides: channels-last maps (what
// the B200 extractor emits) give fully coalesced reads -- the C channels of a
// texel are contiguous and consecutive threads take consecutive channels.
#include "ptk_common.cuh"

namespace {

__device__ __forceinline__ float bilin(const float* __restrict__ m, long long sy, long long sx, int H, int W, float ix,
                                       float iy) {
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float ax = ix - x0f, ay = iy - y0f;
  const bool xi0 = x0 >= 0 && x0 < W, xi1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool yi0 = y0 >= 0 && y0 < H, yi1 = y0 + 1 >= 0 && y0 + 1 < H;
  const float* b = m + (long long)y0 * sy + (long long)x0 * sx;
  const float v00 = (xi0 && yi0) ? __ldg(b) : 0.f;
  const float v01 = (xi1 && yi0) ? __ldg(b + sx) : 0.f;
  const float v10 = (xi0 && yi1) ? __ldg(b + sy) : 0.f;
  const float v11 = (xi1 && yi1) ? __ldg(b + sy + sx) : 0.f;
  return v00 * (1.f - ax) * (1.f - ay) + v01 * ax * (1.f - ay) + v10 * (1.f - ax) * ay + v11 * ax * ay;
}

__global__ void __launch_bounds__(256) sample_kernel(const float* __restrict__ map, long long sc, long long sy,
                                                     long long sx, int C, int H, int W,
                                                     const float* __restrict__ pts, int N, int pad,
                                                     float* __restrict__ vals, uint8_t* __restrict__ mask,
                                                     float* __restrict__ grads) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * C) return;
  const int n = (int)(idx / C), c = (int)(idx - (long long)n * C);
  const float px = pts[2 * n], py = pts[2 * n + 1];
  const float spanx = (float)(W - 1), spany = (float)(H - 1);
  // interpolation.py:66-68 then grid_sampler's unnormalize ((g + 1) / 2 * (size - 1))
  const float gx = fminf(fmaxf((px / spanx) * 2.f - 1.f, -2.f), 2.f);
  const float gy = fminf(fmaxf((py / spany) * 2.f - 1.f, -2.f), 2.f);
  const float* m = map + (long long)c * sc;
  vals[idx] = bilin(m, sy, sx, H, W, ((gx + 1.f) * 0.5f) * spanx, ((gy + 1.f) * 0.5f) * spany);
  if (c == 0 && mask != nullptr) {
    mask[n] = (px >= (float)pad) && (py >= (float)pad) && (px <= (float)(W - pad - 1)) && (py <= (float)(H - pad - 1));
  }
  if (grads != nullptr) {
    const float dx = 1.f / spanx * 2.f, dy = 1.f / spany * 2.f;   // interpolation.py:74-76
    const float iy = ((gy + 1.f) * 0.5f) * spany, ix = ((gx + 1.f) * 0.5f) * spanx;
    const float fx0 = bilin(m, sy, sx, H, W, ((gx - dx + 1.f) * 0.5f) * spanx, iy);
    const float fx1 = bilin(m, sy, sx, H, W, ((gx + dx + 1.f) * 0.5f) * spanx, iy);
    const float fy0 = bilin(m, sy, sx, H, W, ix, ((gy - dy + 1.f) * 0.5f) * spany);
    const float fy1 = bilin(m, sy, sx, H, W, ix, ((gy + dy + 1.f) * 0.5f) * spany);
    grads[2 * idx] = (fx1 - fx0) / 2.f;
    grads[2 * idx + 1] = (fy1 - fy0) / 2.f;
  }
}

}  // namespace

extern "C" int ptk_sample_points(PtkContext* ctx, const float* map, int64_t stride_c, int64_t stride_y,
                                 int64_t stride_x, int32_t C, int32_t H, int32_t W, const float* pts, int32_t N,
                                 int32_t pad, float* vals, uint8_t* mask, float* grads, void* stream) {
  PTK_REQUIRE(ctx && map && vals && (pts || N == 0), "null argument");
  PtkDeviceGuard guard(ctx->device);
  PTK_REQUIRE(C >= 1 && H >= 2 && W >= 2 && N >= 0 && pad >= 0, "bad shape");
  if (N == 0) return PTK_OK;
  const long long total = (long long)N * C;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  sample_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(map, stride_c, stride_y, stride_x, C, H, W, pts, N, pad,
                                                         vals, mask, grads);
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}
