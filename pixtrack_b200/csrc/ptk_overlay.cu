// Result overlay on the device: the NeRF render at the tracked pose blended over the camera frame, plus the pose axes.
//
// Replaces, in pixtrack/visualization/run_vis_on_poses.py (paths relative to /root/reference):
//   blend_images (:215-219)   blend = query * alpha + cvtColor(nerf, BGR2RGB) * (1 - alpha), float64, then astype(uint8)
//   draw_axes (:74-79)        three cv2.line calls (thickness t, 8-connected, colours (255,0,0), (0,255,0), (0,0,255))
//                             between the projected end points add_pose_axes (:82-112) computes on the host.
// One thread per pixel.  A thick OpenCV line is a filled quad of half-width t/2 around the segment plus filled discs of
// radius t/2 at both ends (imgproc/src/drawing.cpp ThickLine); here the same region is described by its distance to the
// segment (<= t/2 + 0.5, which is what the disc of radius t/2 covers on the pixel grid).  Pixels on the polygon's
// rasterisation boundary can differ from OpenCV's scan conversion; the blend itself is exact.
#include "ptk_common.cuh"

namespace {

struct OverlayParams {
  int H, W;
  double alpha;
  int n_seg;
  float seg[3][4];      // x0, y0, x1, y1 in pixels
  float half;           // half thickness
};

__global__ void overlay_kernel(const uint8_t* __restrict__ query, const uint8_t* __restrict__ nerf, uint8_t* __restrict__ out,
                               const OverlayParams P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.H * P.W) return;
  const int x = p % P.W, y = p / P.W;
  uint8_t px[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double q = (double)query[(size_t)p * 3 + c];
    const double n = nerf ? (double)nerf[(size_t)p * 3 + (2 - c)] : 255.0;     // cvtColor(BGR2RGB) = channel swap
    // numpy evaluates q * alpha + n * (1 - alpha) with separately rounded products: no FMA contraction here
    px[c] = (uint8_t)(int)__dadd_rn(__dmul_rn(q, P.alpha), __dmul_rn(n, 1.0 - P.alpha));
  }
  for (int s = 0; s < P.n_seg; ++s) {
    const float ax = P.seg[s][0], ay = P.seg[s][1], bx = P.seg[s][2], by = P.seg[s][3];
    const float dx = bx - ax, dy = by - ay;
    const float len2 = dx * dx + dy * dy;
    float t = len2 > 0.f ? (((float)x - ax) * dx + ((float)y - ay) * dy) / len2 : 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float ex = (float)x - (ax + t * dx), ey = (float)y - (ay + t * dy);
    if (ex * ex + ey * ey <= (P.half + 0.5f) * (P.half + 0.5f)) {
      px[0] = s == 0 ? 255 : 0;         // draw_axes passes (255,0,0), (0,255,0), (0,0,255) in the image's channel order
      px[1] = s == 1 ? 255 : 0;
      px[2] = s == 2 ? 255 : 0;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) out[(size_t)p * 3 + c] = px[c];
}

}  // namespace

extern "C" int ptk_overlay(PtkContext* ctx, const uint8_t* query, const uint8_t* nerf, int32_t H, int32_t W, double alpha,
                           const int16_t* host_axes_px, int32_t thickness, uint8_t* out, void* stream) {
  PTK_REQUIRE(ctx && query && out, "null argument");
  PTK_REQUIRE(H >= 1 && W >= 1 && (long long)H * W < 2147483647LL, "bad image size");
  PTK_REQUIRE(thickness >= 1, "thickness must be >= 1");
  PtkDeviceGuard guard(ctx->device);
  OverlayParams P;
  P.H = H; P.W = W; P.alpha = alpha;
  P.n_seg = host_axes_px ? 3 : 0;
  P.half = 0.5f * (float)thickness;
  for (int s = 0; s < 3; ++s)
    for (int k = 0; k < 4; ++k) P.seg[s][k] = host_axes_px ? (float)host_axes_px[s * 4 + k] : 0.f;
  const int n = H * W;
  overlay_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(query, nerf, out, P);
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}
