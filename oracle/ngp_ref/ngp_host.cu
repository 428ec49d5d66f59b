// Test infrastructure: a host-side harness over the UNMODIFIED instant-ngp headers where they lie under
// /root/reference/instant-ngp (nothing is copied).  It calls the reference's own __host__ __device__ functions
// on the CPU and prints their results as JSON; tests/golden/gen/make_nerf_goldens.py stores that output as
// tests/golden/nerf_host.json, which pins the matching pieces of oracle/nerf.py:
//   ld_random_val                  include/neural-graphics-primitives/random_val.cuh:226-289
//   srgb_to_linear / linear_to_srgb, fov_to_focal_length, pixel_to_ray
//                                  include/neural-graphics-primitives/common_device.cuh:31-61,260-307,470-472
//   BoundingBox::ray_intersect / contains   include/neural-graphics-primitives/bounding_box.cuh:163-220
//   NerfDataset::nerf_matrix_to_ngp         include/neural-graphics-primitives/nerf_loader.h:113-131
// The hash grid, SH encoding, MLPs and the marching loop are __device__-only / .cu-local in the reference and
// cannot run without a GPU; those parts of the oracle stay unpinned.
// Built by oracle/build_ref.py into oracle/_ref/ngp_host (git-ignored).
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/random_val.cuh>
#include <neural-graphics-primitives/common_device.cuh>
#include <neural-graphics-primitives/bounding_box.cuh>
#include <neural-graphics-primitives/nerf_loader.h>

#include <cstdio>
#include <vector>

using namespace ngp;
using namespace Eigen;

static void vec3(const char* key, const Vector3f& v, bool comma = true) {
  printf("\"%s\": [%.9g, %.9g, %.9g]%s", key, v.x(), v.y(), v.z(), comma ? ", " : "");
}

int main() {
  printf("{\n");
  // ---- Owen-scrambled Sobol jitter ----
  const uint32_t seeds[] = {0u, 786433u, 786433u * 7u, 786433u * 123456u, 786433u * 1020000u, 0xdeadbeefu};
  printf("\"ld_random_val\": [");
  for (int s = 0; s < 6; ++s)
    for (uint32_t i = 0; i < 16; ++i)
      printf("%s[%u, %u, %.9g]", (s || i) ? ", " : "", i, seeds[s], ld_random_val(i, seeds[s]));
  printf("],\n");
  // ---- colour transfer ----
  printf("\"srgb\": [");
  for (int i = 0; i <= 64; ++i) {
    const float x = i == 64 ? 0.04045f : (float)i / 63.f;
    printf("%s[%.9g, %.9g, %.9g]", i ? ", " : "", x, srgb_to_linear(x), linear_to_srgb(x));
  }
  printf("],\n");
  printf("\"fov_to_focal\": [");
  const float fovs[] = {20.f, 33.855026f, 45.f, 50.2f, 90.f};
  const int ress[] = {1, 40, 1008, 1920, 756};
  for (int i = 0; i < 5; ++i)
    printf("%s[%d, %.9g, %.9g]", i ? ", " : "", ress[i], fovs[i], fov_to_focal_length(ress[i], fovs[i]));
  printf("],\n");
  // ---- camera matrix conversion ----
  NerfDataset ds = {};
  ds.scale = 0.33f;
  ds.offset = {0.5f, 0.5f, 0.5f};
  ds.from_mitsuba = false;
  Matrix<float, 3, 4> nerf;
  nerf << 0.9f, -0.1f, 0.3f, 0.2f, 0.05f, 0.95f, -0.2f, 0.1f, -0.3f, 0.2f, 0.9f, 4.0f;
  const Matrix<float, 3, 4> ngpm = ds.nerf_matrix_to_ngp(nerf);
  printf("\"nerf_matrix\": [");
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) printf("%s%.9g", (r || c) ? ", " : "", nerf(r, c));
  printf("],\n\"ngp_matrix\": [");
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) printf("%s%.9g", (r || c) ? ", " : "", ngpm(r, c));
  printf("],\n");
  // ---- rays: the renderer's call (testbed_nerf.cu:1823-1849) with pixtrack's settings: snap_to_pixel_centers,
  //      screen centre 0.5, no parallax, no aperture, no distortion; then normalisation and the box test ----
  const Vector2i res(40, 28);
  const float fov = 50.f;
  const float f = fov_to_focal_length(1, fov) * (float)res.x();          // calc_focal_length, fov_axis 0, zoom 1
  const Vector2f focal(f, f);
  printf("\"rays\": {\"width\": %d, \"height\": %d, \"fov\": %.9g, \"camera\": [", res.x(), res.y(), fov);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) printf("%s%.9g", (r || c) ? ", " : "", ngpm(r, c));
  printf("], \"boxes\": [[0, 1], [-1.5, 2.5]], \"samples\": [\n");
  bool first = true;
  for (int y = 0; y < res.y(); y += 3)
    for (int x = 0; x < res.x(); x += 3)
      for (uint32_t spp = 0; spp < 2; ++spp) {
        Ray ray = pixel_to_ray(spp, {x, y}, res, focal, ngpm, Vector2f(0.5f, 0.5f), Vector3f(0.f, 0.f, 1.f), true);
        const float n = ray.d.norm();
        const Vector3f d = (1.0f / n) * ray.d;
        printf("%s{\"x\": %d, \"y\": %d, \"spp\": %u, ", first ? "" : ",\n", x, y, spp);
        first = false;
        vec3("o", ray.o);
        vec3("d_raw", ray.d);
        vec3("d", d);
        const BoundingBox b1(Vector3f::Constant(0.f), Vector3f::Constant(1.f));
        const BoundingBox b4(Vector3f::Constant(-1.5f), Vector3f::Constant(2.5f));
        const Vector2f t1 = b1.ray_intersect(ray.o, d), t4 = b4.ray_intersect(ray.o, d);
        printf("\"t_box1\": [%.9g, %.9g], \"t_box4\": [%.9g, %.9g], \"in_box4\": %d}", t1.x(), t1.y(), t4.x(), t4.y(),
               (int)b4.contains(ray.o));
      }
  printf("]}\n}\n");
  return 0;
}
