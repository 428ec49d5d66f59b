// Query-frame object mask on the device.
//
// Replaces PixLocPoseTrackerR9.get_mask + the multiply in refine()
//   (reference pixtrack/pose_trackers/pixloc_tracker_r9.py:207-214,224-225):
//     depth = get_nerf_image(testbed, nerf_pose, camera, depth=True)            # H x W x 3 uint8
//     img_erosion  = cv2.erode((depth != 0).astype(np.uint8), np.ones((5, 5)), iterations=1)
//     img_dilation = cv2.dilate(img_erosion, np.ones((5, 5)), iterations=5)
//     query_image  = query_image * mask
// cv2's default border for morphology ignores pixels outside the image (erosion pads with the maximum,
// dilation with the minimum), the anchor is the kernel centre, and five 5x5 box dilations are one 21x21 box
// dilation.  Box min / max filters are separable: four byte passes (erode rows, erode columns, dilate rows,
// dilate columns), the last one fused with the multiply into the query image.  In the reference this is a
// device -> host copy of the depth render, two OpenCV calls on the CPU and a host multiply per frame.
#include "ptk_common.cuh"

namespace {

// one separable pass over an interleaved [H][W][3] byte mask: min (kMax = false) or max over 2*R+1 taps
template <int R, bool kMax, bool kHorizontal, bool kFromDepth>
__global__ void morph_pass_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W * 3) return;
  const int c = idx % 3, p = idx / 3;
  const int x = p % W, y = p / W;
  int acc = kMax ? 0 : 1;
#pragma unroll
  for (int d = -R; d <= R; ++d) {
    const int xx = kHorizontal ? x + d : x, yy = kHorizontal ? y : y + d;
    if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;     // outside pixels do not take part
    int v = in[((size_t)yy * W + xx) * 3 + c];
    if (kFromDepth) v = v != 0;
    acc = kMax ? max(acc, v) : min(acc, v);
  }
  out[idx] = (uint8_t)acc;
}

// last pass (vertical max over 21) fused with image * mask
template <typename T>
__global__ void dilate_apply_kernel(const uint8_t* __restrict__ in, int H, int W, const T* __restrict__ img,
                                    T* __restrict__ out, uint8_t* __restrict__ mask_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W * 3) return;
  const int c = idx % 3, p = idx / 3;
  const int x = p % W, y = p / W;
  int acc = 0;
  for (int d = -10; d <= 10; ++d) {
    const int yy = y + d;
    if (yy < 0 || yy >= H) continue;
    acc |= in[((size_t)yy * W + x) * 3 + c];
  }
  if (mask_out) mask_out[idx] = (uint8_t)acc;
  if (out) out[idx] = acc ? img[idx] : (T)0;
}

}  // namespace

extern "C" int ptk_query_mask(PtkContext* ctx, const uint8_t* depth_u8, int32_t H, int32_t W, const void* image,
                              int32_t img_dtype, void* out_image, uint8_t* out_mask, uint8_t* workspace, void* stream) {
  PTK_REQUIRE(ctx && depth_u8 && workspace, "null argument");
  PtkDeviceGuard guard(ctx->device);
  PTK_REQUIRE(H >= 1 && W >= 1 && (long long)H * W * 3 < 2147483647LL, "bad image size");
  PTK_REQUIRE(img_dtype == 0 || img_dtype == 1, "img_dtype must be 0 (fp32) or 1 (uint8)");
  PTK_REQUIRE((image != nullptr) == (out_image != nullptr), "image and out_image must be given together");
  PTK_REQUIRE(out_image || out_mask, "no output");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = H * W * 3, blocks = (n + 255) / 256;
  uint8_t* a = workspace;
  uint8_t* b = workspace + (size_t)n;
  morph_pass_kernel<2, false, true, true><<<blocks, 256, 0, s>>>(depth_u8, a, H, W);   // erode, rows (reads depth != 0)
  morph_pass_kernel<2, false, false, false><<<blocks, 256, 0, s>>>(a, b, H, W);        // erode, columns
  morph_pass_kernel<10, true, true, false><<<blocks, 256, 0, s>>>(b, a, H, W);         // dilate x5, rows
  if (img_dtype == 0)
    dilate_apply_kernel<float><<<blocks, 256, 0, s>>>(a, H, W, (const float*)image, (float*)out_image, out_mask);
  else
    dilate_apply_kernel<uint8_t><<<blocks, 256, 0, s>>>(a, H, W, (const uint8_t*)image, (uint8_t*)out_image, out_mask);
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}
