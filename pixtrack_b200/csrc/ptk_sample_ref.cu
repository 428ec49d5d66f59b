// Reference-view sparse observations in one launch: project the N model points into the reference
// view, sample every pyramid level's descriptor + confidence map there, L2-normalise the
// descriptors, and AND the validity over the levels.
//
// Replaces PoseTrackerRefiner.interp_sparse_observations
//   (reference pixtrack/localization/pixloc_pose_refiners.py:327-368: per level
//    camera.scale(sc).world2image(p3d_cam) -> opt.interpolator(feats, p2d) -> mask & valid,
//    then an O(N*L) Python list-of-lists regroup)
// plus the reference-side part of BaseRefiner.refine_pose_using_features
//   (pixloc/pixloc/localization/base_refiner.py:74-84: torch.stack per level, split descriptor /
//    confidence, F.normalize(F_ref, dim=1)).
// The reference carries the model points, the pose and the camera as float64 on this path (numpy
// xyz, qvec2rotmat, Camera.from_colmap) and casts the pixel position to the map dtype only for
// the interpolation (`p2d_feat.to(feats)`); the projection here is therefore done in double with
// unfused multiplies/adds, the interpolation in float like ptk_sample.cu.
//
// One warp per point: the projection is computed redundantly by the 32 lanes (a few dozen flops),
// the lanes then split the channels of the four corner texels with 16-byte loads from the
// channels-last map, reduce the squared norm with shuffles and write the normalised descriptor
// coalesced.  Points that fall outside any level are flagged in `valid` (the LM launch takes that
// flag as its `mask`; the reference drops them from the lists instead -- same sums).
#include "ptk_common.cuh"

namespace {

struct RefParams {
  PtkRefLevel lv[PTK_MAX_LEVELS];
  int n_levels;
  int N;
  int n_cam;
  int pad;
  const double* p3d;   // [N][3]
  double cam[10];
  double T[12];
  uint8_t* valid;      // [N]
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(256) sample_ref_kernel(const RefParams P) {
  const int lane = threadIdx.x & 31;
  const int pt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pt >= P.N) return;
  const double X = P.p3d[3 * pt], Y = P.p3d[3 * pt + 1], Z = P.p3d[3 * pt + 2];
  // Pose.transform: p @ R^T + t (wrappers.py:177-185), float64
  const double* T = P.T;
  const double pcx = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X, T[0]), __dmul_rn(Y, T[1])), __dmul_rn(Z, T[2])), T[9]);
  const double pcy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X, T[3]), __dmul_rn(Y, T[4])), __dmul_rn(Z, T[5])), T[10]);
  const double pcz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X, T[6]), __dmul_rn(Y, T[7])), __dmul_rn(Z, T[8])), T[11]);
  bool ok = pcz > 1e-3;                                   // Camera.project, wrappers.py:308-314
  const double z = pcz > 1e-3 ? pcz : 1e-3;
  const double xn = pcx / z, yn = pcy / z;
  double xd = xn, yd = yn;
  if (P.n_cam > 6) {                                      // undistort_points, utils.py:36-69
    const double k1 = P.cam[6], k2 = P.cam[7];
    const double r2 = __dadd_rn(__dmul_rn(xn, xn), __dmul_rn(yn, yn));
    const double radial = __dadd_rn(__dmul_rn(k1, r2), __dmul_rn(k2, __dmul_rn(r2, r2)));
    xd = __dadd_rn(xn, __dmul_rn(xn, radial));
    yd = __dadd_rn(yn, __dmul_rn(yn, radial));
    const double disc = 9.0 * k1 * k1 - 20.0 * k2;
    const bool limited = ((k2 > 0.0) && (disc > 0.0)) || ((k2 <= 0.0) && (k1 > 0.0));
    const double limit = fabs(k2 > 0.0 ? (sqrt(disc) - 3.0 * k1) / (10.0 * k2) : 1.0 / (3.0 * k1));
    ok = ok && (!limited || (r2 < limit));
    if (P.n_cam > 8) {
      const double p1 = P.cam[8], p2 = P.cam[9];
      const double uv = __dmul_rn(xn, yn);
      xd = __dadd_rn(__dadd_rn(xd, __dmul_rn(__dmul_rn(2.0, p1), uv)),
                     __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(2.0, __dmul_rn(xn, xn)))));
      yd = __dadd_rn(__dadd_rn(yd, __dmul_rn(__dmul_rn(2.0, p2), uv)),
                     __dmul_rn(p1, __dadd_rn(r2, __dmul_rn(2.0, __dmul_rn(yn, yn)))));
    }
  }
  for (int l = 0; l < P.n_levels; ++l) {
    const PtkRefLevel& L = P.lv[l];
    // Camera.scale (wrappers.py:276-286) then denormalize / in_image (:299-306,:337-339), float64
    const double cw = P.cam[0] * L.sx, ch = P.cam[1] * L.sy;
    const double fx = P.cam[2] * L.sx, fy = P.cam[3] * L.sy;
    const double cx = __dadd_rn(__dmul_rn(__dadd_rn(P.cam[4], 0.5), L.sx), -0.5);
    const double cy = __dadd_rn(__dmul_rn(__dadd_rn(P.cam[5], 0.5), L.sy), -0.5);
    const double ud = __dadd_rn(__dmul_rn(xd, fx), cx), vd = __dadd_rn(__dmul_rn(yd, fy), cy);
    ok = ok && (ud >= 0.0) && (vd >= 0.0) && (ud <= cw - 1.0) && (vd <= ch - 1.0);
    // p2d_feat.to(feats): float from here (interpolation.py:57-95)
    const float px = (float)ud, py = (float)vd;
    const int W = L.W, H = L.H, C = L.C;
    ok = ok && (px >= (float)P.pad) && (py >= (float)P.pad) && (px <= (float)(W - P.pad - 1)) &&
         (py <= (float)(H - P.pad - 1));
    const float spanx = (float)(W - 1), spany = (float)(H - 1);
    const float gx = fminf(fmaxf((px / spanx) * 2.f - 1.f, -2.f), 2.f);
    const float gy = fminf(fmaxf((py / spany) * 2.f - 1.f, -2.f), 2.f);
    const float ix = ((gx + 1.f) * 0.5f) * spanx, iy = ((gy + 1.f) * 0.5f) * spany;
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float ax = ix - x0f, ay = iy - y0f;
    const bool xi0 = x0 >= 0 && x0 < W, xi1 = x0 + 1 >= 0 && x0 + 1 < W;
    const bool yi0 = y0 >= 0 && y0 < H, yi1 = y0 + 1 >= 0 && y0 + 1 < H;
    const float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
    const long long o00 = (long long)y0 * W + x0;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float ss = 0.f;
    float4 acc[4];   // up to C = 512
    const int C4 = C >> 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c4 = lane + 32 * j;
      acc[j] = z4;
      if (c4 < C4) {
        const float* b = L.feat + o00 * C + 4 * c4;
        const float4 v00 = (xi0 && yi0) ? ld4(b) : z4;
        const float4 v01 = (xi1 && yi0) ? ld4(b + C) : z4;
        const float4 v10 = (xi0 && yi1) ? ld4(b + (long long)W * C) : z4;
        const float4 v11 = (xi1 && yi1) ? ld4(b + (long long)W * C + C) : z4;
        float4 r;
        r.x = v00.x * w00 + v01.x * w01 + v10.x * w10 + v11.x * w11;
        r.y = v00.y * w00 + v01.y * w01 + v10.y * w10 + v11.y * w11;
        r.z = v00.z * w00 + v01.z * w01 + v10.z * w10 + v11.z * w11;
        r.w = v00.w * w00 + v01.w * w01 + v10.w * w10 + v11.w * w11;
        acc[j] = r;
        ss += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
    const float inv = L.normalize ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;     // F.normalize(dim=1)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < C4) {
        float4 r = acc[j];
        if (L.normalize) { r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv; }
        *reinterpret_cast<float4*>(L.f_out + (long long)pt * C + 4 * c4) = r;
      }
    }
    if (lane == 0 && L.w_out != nullptr) {
      const float* cb = L.conf + o00;
      const float c00 = (xi0 && yi0) ? __ldg(cb) : 0.f;
      const float c01 = (xi1 && yi0) ? __ldg(cb + 1) : 0.f;
      const float c10 = (xi0 && yi1) ? __ldg(cb + W) : 0.f;
      const float c11 = (xi1 && yi1) ? __ldg(cb + W + 1) : 0.f;
      L.w_out[pt] = c00 * w00 + c01 * w01 + c10 * w10 + c11 * w11;
    }
  }
  if (lane == 0) P.valid[pt] = ok ? 1 : 0;
}

}  // namespace

extern "C" int ptk_sample_reference(PtkContext* ctx, const PtkRefLevel* levels, int32_t n_levels, const double* p3d,
                                    int32_t N, const double* host_cam, int32_t n_cam, const double* host_T,
                                    int32_t pad, uint8_t* valid, void* stream) {
  PTK_REQUIRE(ctx && levels && host_cam && host_T && valid && (p3d || N == 0), "null argument");
  PtkDeviceGuard guard(ctx->device);
  PTK_REQUIRE(n_levels >= 1 && n_levels <= PTK_MAX_LEVELS, "1..PTK_MAX_LEVELS levels");
  PTK_REQUIRE(n_cam == 6 || n_cam == 8 || n_cam == 10, "n_cam must be 6, 8 or 10");
  PTK_REQUIRE(N >= 0 && pad >= 0, "N and pad must be >= 0");
  RefParams P;
  memset(&P, 0, sizeof(P));
  for (int l = 0; l < n_levels; ++l) {
    const PtkRefLevel& L = levels[l];
    PTK_REQUIRE(L.feat && L.f_out && (L.conf != nullptr) == (L.w_out != nullptr), "level pointers");
    PTK_REQUIRE(L.C >= 4 && L.C % 4 == 0 && L.C <= 512 && L.H >= 2 && L.W >= 2, "level shape");
    PTK_REQUIRE(((uintptr_t)L.feat % 16 == 0) && ((uintptr_t)L.f_out % 16 == 0), "16-byte alignment");
    P.lv[l] = L;
  }
  P.n_levels = n_levels; P.N = N; P.n_cam = n_cam; P.pad = pad; P.p3d = p3d; P.valid = valid;
  for (int i = 0; i < 10; ++i) P.cam[i] = i < n_cam ? host_cam[i] : 0.0;
  for (int i = 0; i < 12; ++i) P.T[i] = host_T[i];
  if (N == 0) return PTK_OK;
  sample_ref_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(P);
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}
