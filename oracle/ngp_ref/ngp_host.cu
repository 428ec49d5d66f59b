// Test infrastructure: a host-side harness over the UNMODIFIED instant-ngp headers where they lie under
// /root/reference/instant-ngp (nothing is copied).  It calls the reference's own __host__ __device__ functions
// on the CPU and prints their results as JSON; tests/golden/gen/make_nerf_goldens.py stores that output as
// tests/golden/nerf_host.json, which pins the matching pieces of oracle/nerf.py:
//   ld_random_val                  include/neural-graphics-primitives/random_val.cuh:226-289
//   srgb_to_linear / linear_to_srgb, fov_to_focal_length, pixel_to_ray
//                                  include/neural-graphics-primitives/common_device.cuh:31-61,260-307,470-472
//   BoundingBox::ray_intersect / contains   include/neural-graphics-primitives/bounding_box.cuh:163-220
//   NerfDataset::nerf_matrix_to_ngp         include/neural-graphics-primitives/nerf_loader.h:113-131
//   calc_dt, mip_from_pos, mip_from_dt, cascaded_grid_idx_at, density_grid_occupied_at, distance_to_next_voxel,
//   advance_to_next_voxel, warp_position / unwarp_position / warp_direction / warp_dt / unwarp_dt and the step-size
//   constants                      src/testbed_nerf.cu:46-62,77-92,185-207,261-336,443-457 -- local to that .cu and
//                                  mostly __device__: oracle/build_ref.py lifts these definitions at build time into a
//                                  temporary include with __device__ widened to __host__ __device__ (bodies untouched)
//   tcnn::morton3D / morton3D_invert / logistic    tiny-cuda-nn/include/tiny-cuda-nn/common_device.h:52-54,339-363
//   tcnn fast_hash / grid_index    tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:82-116 (lifted the same way)
//   tcnn kernel_grid (hash-grid forward: level scale, pos_fract, trilinear interpolation in fp16) and kernel_sh
//                                  encodings/grid.h:135-340, encodings/spherical_harmonics.h:46-150, common_device.h:379-431:
//                                  the __global__ kernels are lifted as host functions, thread indices via macros
//   composite_kernel_nerf + network_to_rgb / network_to_density   src/testbed_nerf.cu:209-259,754-960 (lifted as a
//                                  host function; __expf -> expf)
//   init_rays_with_payload_kernel_nerf, advance_pos_nerf, shade_kernel_nerf   src/testbed_nerf.cu:606-657,1721-1754,
//                                  1781-1890; accumulate_kernel   src/render_buffer.cu:236-271 (same lifting)
//   generate_next_nerf_network_inputs (the marching kernel)   src/testbed_nerf.cu:693-752
//   compact_kernel_nerf (atomicAdd -> a host counter)   src/testbed_nerf.cu:1756-1779; tonemap + tonemap_kernel
//                                  (surf2Dwrite -> a host array)   src/render_buffer.cu:273-345,542-569
//   THE ORCHESTRATION: `render_chain` strings the lifted kernels together exactly as the reference does --
//                                  Testbed::render_to_cpu (src/python_api.cu:127-173: one render_frame per sample),
//                                  render_frame (src/testbed.cu:2591-2749: clear, render_nerf, accumulate, tonemap),
//                                  render_nerf (src/testbed_nerf.cu:2228-2359), NerfTracer::init_rays_from_camera (:1956-2033)
//                                  and NerfTracer::trace (:2035-2157: double-buffered compaction, n_steps_between_compaction,
//                                  generate -> network -> composite rounds, final hit list) -- with an analytic stand-in for
//                                  the network (simple float arithmetic rounded to network_precision_t, restated in
//                                  tests/test_nerf_oracle.py), and prints the final image of a Shade and a Depth render.
//   NerfNetwork::set_params        include/neural-graphics-primitives/nerf_network.h:361-395: the body is lifted into a mock
//                                  class whose four sub-modules record the parameter offsets they are handed (the order of
//                                  `params_binary` in a snapshot).
// The fused MLPs (wmma fragments) cannot run without a GPU; that part of the oracle stays unpinned.
// Built by oracle/build_ref.py into oracle/_ref/ngp_host (git-ignored).
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/random_val.cuh>
#include <neural-graphics-primitives/common_device.cuh>
#include <neural-graphics-primitives/bounding_box.cuh>
#include <neural-graphics-primitives/nerf_loader.h>
#include <neural-graphics-primitives/nerf.h>
#include <neural-graphics-primitives/envmap.cuh>
#include <tiny-cuda-nn/common_device.h>
#include <tiny-cuda-nn/encodings/grid.h>

#include <cstdio>
#include <vector>

#include <vector_types.h>

using namespace Eigen;
using namespace tcnn;

namespace lifted {   // CUDA thread indices for the kernels that run here as host functions
static uint3 h_tid = {0, 0, 0}, h_bid = {0, 0, 0};
static dim3 h_bdim(1, 1, 1);
// stand-ins for the two device-only facilities the lifted kernels touch: a surface write and an atomic counter
static std::vector<float4> host_surface;
static int host_surface_w = 0;
static uint32_t host_atomic_add(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p += v; return o; }
}

NGP_NAMESPACE_BEGIN
// read_envmap is __device__-only; the ray-init kernel only calls it when an environment map is given (never here)
template <typename T>
Eigen::Array4f host_read_envmap(const T*, const Eigen::Vector2i, const Eigen::Vector3f&) { return Eigen::Array4f::Zero(); }
#define threadIdx lifted::h_tid
#define blockIdx lifted::h_bid
#define blockDim lifted::h_bdim
#define __expf expf          // the fast-math intrinsic has no host version; expf is its exact counterpart
#define read_envmap host_read_envmap
#define atomicAdd lifted::host_atomic_add
#define surf2Dwrite(v, s, xb, y) (lifted::host_surface[(size_t)(y) * lifted::host_surface_w + (xb) / sizeof(float4)] = (v))
#include "testbed_nerf_helpers.inc"
#undef surf2Dwrite
#undef atomicAdd
#undef read_envmap
#undef __expf
#undef threadIdx
#undef blockIdx
#undef blockDim
NGP_NAMESPACE_END

// tcnn's __device__ index / interpolation functions widened to __host__ __device__, and its hash-grid and SH encoding
// kernels as plain host functions (bodies untouched): renamed through macros so that they do not collide with the
// originals that grid.h declares, with the CUDA thread indices supplied as host variables.
namespace lifted {
#define threadIdx lifted::h_tid
#define blockIdx lifted::h_bid
#define blockDim lifted::h_bdim
#define fast_hash lifted_fast_hash
#define grid_index lifted_grid_index
#define pos_fract lifted_pos_fract
#define identity_fun lifted_identity_fun
#define identity_derivative lifted_identity_derivative
#define smoothstep lifted_smoothstep
#define smoothstep_derivative lifted_smoothstep_derivative
#define kernel_grid lifted_kernel_grid
#define kernel_sh lifted_kernel_sh
#include "tcnn_grid_helpers.inc"
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef fast_hash
#undef grid_index
#undef pos_fract
#undef identity_fun
#undef identity_derivative
#undef smoothstep
#undef smoothstep_derivative
#undef kernel_grid
#undef kernel_sh
}

using namespace ngp;

// NerfNetwork::set_params lifted into a mock: the sub-modules record where in the parameter vector they are pointed.
namespace paramorder {
struct Recorder {
  const char* name;
  size_t n;
  long long offset = -1;
  uint16_t* base = nullptr;
  void set_params(uint16_t* params, uint16_t*, uint16_t*, uint16_t*) { offset = (long long)(params - base); }
  size_t n_params() const { return n; }
};
struct MockNerfNetwork {
  using T = uint16_t;
  Recorder *m_density_network, *m_rgb_network, *m_pos_encoding, *m_dir_encoding;
#define override
#include "nerf_network_set_params.inc"
#undef override
};
}

static uint32_t lcg_state = 12345u;
static float rnd() {   // uniform in [0, 1)
  lcg_state = lcg_state * 1664525u + 1013904223u;
  return (float)(lcg_state >> 8) * (1.0f / 16777216.0f);
}

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
  printf("]},\n");
  // ---- marching helpers (testbed_nerf.cu) ----
  printf("\"constants\": {\"near\": %.9g, \"stepsize\": %.9g, \"min_cone_stepsize\": %.9g, \"max_cone_stepsize\": %.9g, \"cascades\": %u},\n",
         NERF_RENDERING_NEAR_DISTANCE(), STEPSIZE(), MIN_CONE_STEPSIZE(), MAX_CONE_STEPSIZE(), NERF_CASCADES());
  printf("\"calc_dt\": [");
  {
    const float ts[] = {0.05f, 0.3f, 0.5f, 1.f, 2.f, 5.f, 20.f, 200.f};
    const float cones[] = {0.f, 1.f / 256.f};
    for (int c = 0; c < 2; ++c)
      for (int i = 0; i < 8; ++i)
        printf("%s[%.9g, %.9g, %.9g]", (c || i) ? ", " : "", ts[i], cones[c], calc_dt(ts[i], cones[c]));
  }
  printf("],\n\"warp\": [");
  {
    const BoundingBox box(Vector3f(-1.5f, -1.5f, -1.5f), Vector3f(2.5f, 2.5f, 2.5f));
    for (int i = 0; i < 12; ++i) {
      const Vector3f p(rnd() * 4.f - 1.5f, rnd() * 4.f - 1.5f, rnd() * 4.f - 1.5f);
      const float dt = MIN_CONE_STEPSIZE() * (1.f + rnd() * 100.f);
      const Vector3f w = warp_position(p, box), u = unwarp_position(w, box), wd = warp_direction(p);
      printf("%s{", i ? ", " : "");
      vec3("p", p); vec3("warped", w); vec3("unwarped", u); vec3("warp_direction", wd);
      printf("\"dt\": %.9g, \"warp_dt\": %.9g, \"unwarp_dt\": %.9g}", dt, warp_dt(dt), unwarp_dt(warp_dt(dt)));
    }
  }
  printf("],\n\"morton\": [");
  for (int i = 0; i < 24; ++i) {
    const uint32_t x = (uint32_t)(rnd() * 128.f), y = (uint32_t)(rnd() * 128.f), z = (uint32_t)(rnd() * 128.f);
    const uint32_t m = tcnn::morton3D(x, y, z);
    printf("%s[%u, %u, %u, %u, %u]", i ? ", " : "", x, y, z, m, tcnn::morton3D_invert(m >> 1));
  }
  printf("],\n\"logistic\": [");
  for (int i = 0; i < 9; ++i) printf("%s[%.9g, %.9g]", i ? ", " : "", -8.f + 2.f * i, tcnn::logistic(-8.f + 2.f * i));
  printf("],\n\"grid_index\": [");
  {
    // the 16 levels of the base config at aabb_scale 1: resolution, entries of the level (hashmap_size argument)
    const uint32_t ress[] = {16, 23, 31, 43, 59, 81, 112, 154, 213, 295, 407, 562, 777, 1073, 1483, 2048};
    bool first_gi = true;
    for (int lv = 0; lv < 16; ++lv) {
      const uint64_t dense = (uint64_t)ress[lv] * ress[lv] * ress[lv];
      const uint32_t size = (uint32_t)std::min<uint64_t>((dense + 7) / 8 * 8, 1u << 19);
      for (int k = 0; k < 6; ++k) {
        uint32_t pg[3] = {(uint32_t)(rnd() * ress[lv]), (uint32_t)(rnd() * ress[lv]), (uint32_t)(rnd() * ress[lv])};
        if (k == 5) { pg[0] = ress[lv] - 1; pg[1] = ress[lv] - 1; pg[2] = ress[lv] - 1; }
        printf("%s[%u, %u, %u, %u, %u, %u, %u]", first_gi ? "" : ", ", ress[lv], size, pg[0], pg[1], pg[2],
               lifted::lifted_grid_index<3, 2>(GridType::Hash, 0, size, ress[lv], pg), lifted::lifted_fast_hash<3>(pg));
        first_gi = false;
      }
    }
  }
  // ---- hash-grid and SH encodings: the reference kernels run as host loops ----
  {
    const float pls = std::exp(std::log(2048.f * 1.f / 16.f) / (16 - 1));      // testbed.cu:2243, aabb_scale 1
    const float l2 = std::log2(pls);
    GridOffsetTable table;
    uint32_t off = 0;
    printf("],\n\"encoding\": {\"log2_per_level_scale\": %.9g, \"levels\": [", l2);
    for (uint32_t lv = 0; lv < 16; ++lv) {
      const float scale = exp2f(lv * l2) * 16 - 1.0f;
      const uint32_t res = (uint32_t)ceil(scale) + 1;
      const uint64_t dense = (uint64_t)res * res * res;
      const uint32_t size = (uint32_t)std::min<uint64_t>((dense + 7) / 8 * 8, 1u << 19);     // grid.h:898-930
      table.data[lv] = off;
      off += size;
      printf("%s[%.9g, %u, %u]", lv ? ", " : "", scale, res, size);
    }
    table.data[16] = off;
    table.size = 17;
    // table values: a reproducible pattern in [-0.5, 0.5), exactly representable before the fp16 rounding
    std::vector<__half> grid((size_t)off * 2);
    for (size_t i = 0; i < grid.size(); ++i)
      grid[i] = __float2half((float)((((uint32_t)i * 2654435761u) >> 16) & 0xFFFFu) / 65536.f - 0.5f);
    const uint32_t n = 48;
    std::vector<float> pts(n * 3), dirs(n * 3);
    for (uint32_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        pts[i * 3 + k] = rnd();
        dirs[i * 3 + k] = rnd();
      }
    pts[0] = pts[1] = pts[2] = 0.f;                               // corners and an exact vertex of level 0
    pts[3] = pts[4] = pts[5] = 0.99999f;
    pts[6] = 3.5f / 15.f; pts[7] = 4.5f / 15.f; pts[8] = 6.5f / 15.f;
    std::vector<__half> enc((size_t)32 * n), sh((size_t)16 * n);
    MatrixView<const float> pos_view(pts.data(), 1, 3), dir_view(dirs.data(), 1, 3);
    MatrixView<__half> sh_view(sh.data(), 1, 16);
    for (uint32_t i = 0; i < n; ++i) {
      lifted::h_tid.x = i;
      for (uint32_t lv = 0; lv < 16; ++lv) {
        lifted::h_bid.y = lv;
        lifted::lifted_kernel_grid<__half, 3, 2>(n, 32, table, 16, l2, 0.f, 1000.f, nullptr, InterpolationType::Linear,
                                                 GridType::Hash, grid.data(), pos_view, enc.data(), nullptr);
      }
      lifted::h_bid.y = 0;
      lifted::lifted_kernel_sh<__half>(n, 4, 0, dir_view, sh_view);
    }
    printf("], \"total_entries\": %u, \"samples\": [\n", off);
    for (uint32_t i = 0; i < n; ++i) {
      printf("%s{\"pos\": [%.9g, %.9g, %.9g], \"dir01\": [%.9g, %.9g, %.9g], \"enc\": [", i ? ",\n" : "", pts[i * 3], pts[i * 3 + 1],
             pts[i * 3 + 2], dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]);
      for (int f = 0; f < 32; ++f) printf("%s%.9g", f ? ", " : "", __half2float(enc[i + (size_t)f * n]));
      printf("], \"sh\": [");
      for (int f = 0; f < 16; ++f) printf("%s%.9g", f ? ", " : "", __half2float(sh[(size_t)i * 16 + f]));
      printf("]}");
    }
    printf("]},\n\"unused\": [");
  }
  // ---- compositing: the reference kernel over R rays x up to 6 samples, Shade and Depth modes ----
  printf("], \"composite\": [");
  for (int mode = 0; mode < 2; ++mode) {
    const uint32_t R = 24, S = 6;
    const BoundingBox aabb(Vector3f::Constant(-0.5f), Vector3f::Constant(1.5f));
    Matrix<float, 3, 4> cam;
    cam << 0.9f, -0.1f, 0.3f, 0.4f, 0.05f, 0.95f, -0.2f, -0.7f, -0.3f, 0.2f, 0.9f, 0.6f;
    std::vector<Array4f> rgba(R, Array4f::Zero());
    std::vector<float> depth(R, 0.f);
    std::vector<NerfPayload> pay(R);
    std::vector<NerfCoordinate> coords((size_t)R * S, NerfCoordinate(Vector3f::Zero(), Vector3f::Zero(), 0.f));
    std::vector<network_precision_t> out((size_t)4 * R * S);
    printf("%s{\"depth_mode\": %d, \"aabb\": [-0.5, 1.5], \"depth_scale\": %.9g, \"min_transmittance\": 1e-7, \"camera\": [", mode ? ", " : "",
           mode, 1.f / 0.33f);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) printf("%s%.9g", (r || c) ? ", " : "", cam(r, c));
    printf("], \"rays\": [\n");
    for (uint32_t i = 0; i < R; ++i) {
      pay[i].origin = cam.col(3);
      pay[i].dir = Vector3f(0.f, 0.f, 1.f);
      pay[i].t = 0.f;
      pay[i].max_weight = 0.f;
      pay[i].idx = i;
      pay[i].n_steps = (uint16_t)(1 + (i % S));
      pay[i].alive = (i % 11) != 10;
      // some rays start partly composited (a later compaction round)
      if (i % 3 == 1) rgba[i] = Array4f(0.1f * rnd(), 0.1f * rnd(), 0.1f * rnd(), 0.3f * rnd());
      printf("%s{\"n_steps\": %u, \"alive\": %d, \"rgba0\": [%.9g, %.9g, %.9g, %.9g], \"samples\": [", i ? ",\n" : "", (unsigned)pay[i].n_steps,
             (int)pay[i].alive, rgba[i].x(), rgba[i].y(), rgba[i].z(), rgba[i].w());
      for (uint32_t j = 0; j < S; ++j) {
        const Vector3f wpos(rnd(), rnd(), rnd());
        const float dt = MIN_CONE_STEPSIZE() * (1.f + 30.f * rnd());
        NerfCoordinate& c = coords[i + (size_t)j * R];
        c.pos.p = wpos;
        c.dt = warp_dt(dt);
        float raw[4] = {rnd() * 6.f - 3.f, rnd() * 6.f - 3.f, rnd() * 6.f - 3.f, (i % 4 == 3) ? 6.f + 6.f * rnd() : rnd() * 8.f - 2.f};
        for (int k = 0; k < 4; ++k) {
          out[i + (size_t)j * R + (size_t)k * R * S] = (network_precision_t)raw[k];
          raw[k] = (float)out[i + (size_t)j * R + (size_t)k * R * S];
        }
        printf("%s[%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g]", j ? ", " : "", wpos.x(), wpos.y(), wpos.z(), c.dt, raw[0], raw[1], raw[2], raw[3]);
      }
      printf("]}");
    }
    for (uint32_t i = 0; i < R; ++i) {
      lifted::h_tid.x = i;
      composite_kernel_nerf(R, R * S, 0, aabb, 0.f, 0, 0, nullptr, cam, Vector2f(100.f, 100.f), 1.f / 0.33f, rgba.data(), depth.data(),
                            pay.data(), PitchedPtr<NerfCoordinate>(coords.data(), 1), out.data(), 16, S,
                            mode ? ERenderMode::Depth : ERenderMode::Shade, nullptr, ENerfActivation::Logistic,
                            ENerfActivation::Exponential, -1, 1e-7f);
    }
    printf("], \"result\": [\n");
    for (uint32_t i = 0; i < R; ++i)
      printf("%s[%.9g, %.9g, %.9g, %.9g, %.9g, %d, %u, %.9g]", i ? ", " : "", rgba[i].x(), rgba[i].y(), rgba[i].z(), rgba[i].w(), depth[i],
             (int)pay[i].alive, (unsigned)pay[i].n_steps, pay[i].max_weight);
    printf("]}");
  }
  // occupancy bitfield with a reproducible pattern: byte i = (i * 2654435761) >> 13, 8 cascades
  std::vector<uint8_t> bits((size_t)NERF_CASCADES() * 128 * 128 * 128 / 8);
  for (size_t i = 0; i < bits.size(); ++i) bits[i] = (uint8_t)(((uint32_t)i * 2654435761u) >> 13);
  // ---- start of a ray: init kernel + jittered first advance over that occupancy pattern (two sample indices) ----
  printf("], \"ray_start\": {");
  {
    const Vector2i res(16, 10);
    const float fov = 40.f;
    const float f = fov_to_focal_length(1, fov) * (float)res.x();
    Matrix<float, 3, 4> cam;            // camera outside the unit cube, looking at its centre (NGP convention)
    cam << 0.948683f, -0.094916f, 0.301511f, -0.4f,
           0.f,        0.953463f, 0.301511f, -0.4f,
          -0.316228f, -0.284747f, 0.904534f, -2.2f;
    const BoundingBox box(Vector3f::Constant(0.f), Vector3f::Constant(1.f));
    printf("\"width\": %d, \"height\": %d, \"fov\": %.9g, \"camera\": [", res.x(), res.y(), fov);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) printf("%s%.9g", (r || c) ? ", " : "", cam(r, c));
    printf("], \"passes\": [");
    for (uint32_t spp = 0; spp < 2; ++spp) {
      const uint32_t n = (uint32_t)(res.x() * res.y());
      std::vector<NerfPayload> pay(n);
      std::vector<Array4f> fb(n, Array4f::Zero());
      std::vector<float> db(n, 0.f);
      for (int y = 0; y < res.y(); ++y)
        for (int x = 0; x < res.x(); ++x) {
          lifted::h_tid.x = (uint32_t)x; lifted::h_tid.y = (uint32_t)y;
          init_rays_with_payload_kernel_nerf(spp, pay.data(), res, Vector2f(f, f), cam, cam, Vector4f::Zero(), Vector2f(0.5f, 0.5f),
                                             Vector3f(0.f, 0.f, 1.f), true, box, Matrix3f::Identity(), 1.0f, 0.0f, CameraDistortion{}, nullptr,
                                             Vector2i::Zero(), fb.data(), db.data(), nullptr, Vector2i::Zero(), ERenderMode::Shade);
        }
      lifted::h_tid.y = 0;
      std::vector<float> t_init(n);
      std::vector<int> alive_init(n);
      for (uint32_t i = 0; i < n; ++i) { t_init[i] = pay[i].alive ? pay[i].t : 0.f; alive_init[i] = pay[i].alive; }
      for (uint32_t i = 0; i < n; ++i) {
        lifted::h_tid.x = i;
        advance_pos_nerf(n, box, Matrix3f::Identity(), cam.col(2), Vector2f(f, f), spp, pay.data(), bits.data(), 0, 0.f);
      }
      printf("%s[", spp ? ", " : "");
      for (uint32_t i = 0; i < n; ++i)
        printf("%s[%d, %.9g, %d, %.9g]", i ? ", " : "", alive_init[i], t_init[i], (int)pay[i].alive, pay[i].alive ? pay[i].t : 0.f);
      printf("]");
      if (spp == 1) {
        // the marching kernel on top of pass 1: up to 4 samples per ray (generate_next_nerf_network_inputs)
        const uint32_t S = 4;
        std::vector<NerfCoordinate> coords((size_t)n * S, NerfCoordinate(Vector3f::Zero(), Vector3f::Zero(), 0.f));
        for (uint32_t i = 0; i < n; ++i) {
          lifted::h_tid.x = i;
          generate_next_nerf_network_inputs(n, box, Matrix3f::Identity(), box, Vector2f(f, f), cam.col(2), pay.data(),
                                            PitchedPtr<NerfCoordinate>(coords.data(), 1), S, bits.data(), 0, 0.f, nullptr);
        }
        printf("], \"march_from_pass_1\": [");
        for (uint32_t i = 0; i < n; ++i) {
          printf("%s{\"alive\": %d, \"n_steps\": %u, \"t\": %.9g, \"samples\": [", i ? ", " : "", (int)pay[i].alive, (unsigned)pay[i].n_steps,
                 pay[i].alive ? pay[i].t : 0.f);
          for (uint32_t j = 0; pay[i].alive && j < pay[i].n_steps; ++j) {
            const NerfCoordinate& c = coords[i + (size_t)j * n];
            printf("%s[%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g]", j ? ", " : "", c.pos.p.x(), c.pos.p.y(), c.pos.p.z(), c.dt, c.dir.d.x(),
                   c.dir.d.y(), c.dir.d.z());
          }
          printf("]}");
        }
      }
    }
    printf("]},\n");
  }
  // ---- end of a sample pass: shade (sRGB -> linear into the frame buffer) and the running mean over samples ----
  {
    const uint32_t n = 12;
    std::vector<Array4f> rgba(n), fb(n, Array4f::Zero()), acc(n, Array4f::Zero());
    std::vector<float> dep(n), dbuf(n, 0.f);
    std::vector<NerfPayload> pay(n);
    printf("\"shade\": {\"passes\": [");
    for (uint32_t spp = 0; spp < 3; ++spp) {
      for (uint32_t i = 0; i < n; ++i) {
        const float a = (i % 4 == 0) ? 0.1f * rnd() : (i % 4 == 1 ? 1.0f : rnd());
        rgba[i] = Array4f(rnd() * a, rnd() * a, rnd() * a, a);
        dep[i] = 1.f + rnd();
        pay[i].idx = n - 1 - i;                                  // scattered write through payload.idx
        fb[i] = Array4f::Zero();
        dbuf[i] = 0.f;
      }
      printf("%s{\"rgba\": [", spp ? ", " : "");
      for (uint32_t i = 0; i < n; ++i) printf("%s[%.9g, %.9g, %.9g, %.9g, %.9g]", i ? ", " : "", rgba[i].x(), rgba[i].y(), rgba[i].z(), rgba[i].w(), dep[i]);
      for (uint32_t i = 0; i < n; ++i) {
        lifted::h_tid.x = i;
        shade_kernel_nerf(n, rgba.data(), dep.data(), pay.data(), ERenderMode::Shade, false, fb.data(), dbuf.data());
      }
      for (uint32_t i = 0; i < n; ++i) {
        lifted::h_tid.x = i; lifted::h_tid.y = 0;
        accumulate_kernel(Vector2i((int)n, 1), fb.data(), acc.data(), (float)spp, EColorSpace::Linear);
      }
      printf("], \"frame\": [");
      for (uint32_t i = 0; i < n; ++i) printf("%s[%.9g, %.9g, %.9g, %.9g, %.9g]", i ? ", " : "", fb[i].x(), fb[i].y(), fb[i].z(), fb[i].w(), dbuf[i]);
      printf("], \"accumulated\": [");
      for (uint32_t i = 0; i < n; ++i) printf("%s[%.9g, %.9g, %.9g, %.9g]", i ? ", " : "", acc[i].x(), acc[i].y(), acc[i].z(), acc[i].w());
      printf("]}");
    }
    printf("]},\n");
  }
  // ---- compaction (which finished rays reach the frame) and tonemap (background blend, exposure 0, identity curve) ----
  {
    const uint32_t n = 16;
    std::vector<Array4f> src(n), dst(n), fin(n);
    std::vector<float> sd(n, 1.f), dd(n), fd(n);
    std::vector<NerfPayload> sp(n), dp(n), fp(n);
    uint32_t counter = 0, final_counter = 0;
    printf("\"compaction\": {\"rays\": [");
    for (uint32_t i = 0; i < n; ++i) {
      const float a = (i % 4 == 0) ? 0.0005f + 0.001f * rnd() : rnd();
      src[i] = Array4f(a, a, a, (i == 5) ? 0.001f : a);
      sp[i].alive = (i % 5 == 2);
      sp[i].idx = i;
      printf("%s[%d, %.9g]", i ? ", " : "", (int)sp[i].alive, src[i].w());
    }
    for (uint32_t i = 0; i < n; ++i) {
      lifted::h_tid.x = i;
      compact_kernel_nerf(n, src.data(), sd.data(), sp.data(), dst.data(), dd.data(), dp.data(), fin.data(), fd.data(), fp.data(), &counter,
                          &final_counter);
    }
    printf("], \"still_alive\": [");
    for (uint32_t i = 0; i < counter; ++i) printf("%s%u", i ? ", " : "", dp[i].idx);
    printf("], \"final\": [");
    for (uint32_t i = 0; i < final_counter; ++i) printf("%s%u", i ? ", " : "", fp[i].idx);
    printf("]},\n\"tonemap\": [");
    const Array4f bgs[2] = {Array4f(255.f, 255.f, 255.f, 0.f), Array4f(0.2f, 0.5f, 0.8f, 1.0f)};     // pixtrack's default; an opaque one
    for (int b = 0; b < 2; ++b) {
      std::vector<Array4f> acc(n);
      for (uint32_t i = 0; i < n; ++i) {
        const float a = (i % 3 == 0) ? 1.f : rnd();
        acc[i] = Array4f(rnd() * a, rnd() * a, 3.f * rnd() * a, a);
      }
      lifted::host_surface.assign(n, float4{0, 0, 0, 0});
      lifted::host_surface_w = (int)n;
      lifted::h_tid.y = 0;
      for (uint32_t i = 0; i < n; ++i) {
        lifted::h_tid.x = i;
        tonemap_kernel(Vector2i((int)n, 1), 0.0f, bgs[b], acc.data(), EColorSpace::Linear, EColorSpace::Linear, ETonemapCurve::Identity, false, 0);
      }
      printf("%s{\"background\": [%.9g, %.9g, %.9g, %.9g], \"accumulated\": [", b ? ", " : "", bgs[b].x(), bgs[b].y(), bgs[b].z(), bgs[b].w());
      for (uint32_t i = 0; i < n; ++i) printf("%s[%.9g, %.9g, %.9g, %.9g]", i ? ", " : "", acc[i].x(), acc[i].y(), acc[i].z(), acc[i].w());
      printf("], \"out\": [");
      for (uint32_t i = 0; i < n; ++i)
        printf("%s[%.9g, %.9g, %.9g, %.9g]", i ? ", " : "", lifted::host_surface[i].x, lifted::host_surface[i].y, lifted::host_surface[i].z, lifted::host_surface[i].w);
      printf("]}");
    }
    printf("],\n\"unused2\": [");
  }
  printf("],\n");

  // ---- parameter order of a snapshot (NerfNetwork::set_params) ----
  {
    using namespace paramorder;
    std::vector<uint16_t> buf(8);
    // base.json network: density MLP 32 -> 64 -> 16 (padded), rgb MLP 32 (16 + SH16) -> 64 -> 64 -> 16 (padded); hash grid of
    // aabb_scale 1 (tests/synthetic.py::nerf_grid_size); the SH encoding has no parameters
    Recorder dn{"density_network", 64 * 32 + 16 * 64}, rn{"rgb_network", 64 * 32 + 64 * 64 + 16 * 64}, pe{"pos_encoding", 2 * 6098120},
        de{"dir_encoding", 0};
    dn.base = rn.base = pe.base = de.base = buf.data();
    MockNerfNetwork net{&dn, &rn, &pe, &de};
    net.set_params(buf.data(), buf.data(), buf.data(), buf.data());
    printf("\"set_params\": [");
    const Recorder* all[4] = {&dn, &rn, &pe, &de};
    for (int i = 0; i < 4; ++i) printf("%s{\"module\": \"%s\", \"n_params\": %zu, \"offset\": %lld}", i ? ", " : "", all[i]->name, all[i]->n, all[i]->offset);
    printf("],\n");
  }
  // ---- the whole render: python_api.cu render_to_cpu -> render_frame -> render_nerf -> NerfTracer, analytic network ----
  printf("\"render_chain\": [");
  for (int mode = 0; mode < 2; ++mode) {
    const Vector2i res(20, 14);
    const uint32_t n = (uint32_t)(res.x() * res.y());
    const int SPP = 3;
    const float fov = 34.f, min_T = 0.01f, depth_scale = 1.f / 0.33f;
    const float f = fov_to_focal_length(1, fov) * (float)res.x();
    Matrix<float, 3, 4> cam;
    cam << 0.948683f, -0.094916f, 0.301511f, -0.25f,
           0.f,        0.953463f, 0.301511f, -0.2f,
          -0.316228f, -0.284747f, 0.904534f, -1.9f;
    const BoundingBox box(Vector3f::Constant(0.f), Vector3f::Constant(1.f));     // render box = training box, aabb_scale 1
    const float cone = 0.f;                                                       // testbed_nerf.cu:2596 for aabb_scale 1
    const ERenderMode rmode = mode ? ERenderMode::Depth : ERenderMode::Shade;
    const Array4f background(255.f, 255.f, 255.f, 0.f);                           // ingp_utils.py:31
    std::vector<Array4f> acc(n, Array4f::Zero());
    std::vector<float> db(n, 0.f);
    std::vector<uint32_t> hits, rounds;
    for (int spp = 0; spp < SPP; ++spp) {
      std::vector<Array4f> fb(n, Array4f::Zero());                                // render_buffer.clear_frame
      std::fill(db.begin(), db.end(), 0.f);
      // NerfTracer::init_rays_from_camera
      struct Soa { std::vector<Array4f> rgba; std::vector<float> depth; std::vector<NerfPayload> payload; };
      Soa rays[2], hit;
      for (Soa* r : {&rays[0], &rays[1], &hit}) { r->rgba.assign(n, Array4f::Zero()); r->depth.assign(n, 0.f); r->payload.resize(n); }
      for (int y = 0; y < res.y(); ++y)
        for (int x = 0; x < res.x(); ++x) {
          lifted::h_tid.x = (uint32_t)x; lifted::h_tid.y = (uint32_t)y;
          init_rays_with_payload_kernel_nerf((uint32_t)spp, rays[0].payload.data(), res, Vector2f(f, f), cam, cam, Vector4f::Zero(),
                                             Vector2f(0.5f, 0.5f), Vector3f(0.f, 0.f, 1.f), true, box, Matrix3f::Identity(), 1.0f, 0.0f,
                                             CameraDistortion{}, nullptr, Vector2i::Zero(), fb.data(), db.data(), nullptr, Vector2i::Zero(), rmode);
        }
      lifted::h_tid.y = 0;
      for (uint32_t i = 0; i < n; ++i) {
        lifted::h_tid.x = i;
        advance_pos_nerf(n, box, Matrix3f::Identity(), cam.col(2), Vector2f(f, f), (uint32_t)spp, rays[0].payload.data(), bits.data(), 0, cone);
      }
      // NerfTracer::trace
      uint32_t hit_counter = 0, n_alive = n, step = 1, dbi = 0, n_rounds = 0;
      std::vector<NerfCoordinate> coords((size_t)n * 8, NerfCoordinate(Vector3f::Zero(), Vector3f::Zero(), 0.f));
      std::vector<network_precision_t> out;
      while (step < 10000u) {                                                       // MARCH_ITER
        Soa& cur = rays[(dbi + 1) % 2];
        Soa& tmp = rays[dbi % 2];
        ++dbi;
        uint32_t alive_counter = 0;
        for (uint32_t i = 0; i < n_alive; ++i) {
          lifted::h_tid.x = i;
          compact_kernel_nerf(n_alive, tmp.rgba.data(), tmp.depth.data(), tmp.payload.data(), cur.rgba.data(), cur.depth.data(),
                              cur.payload.data(), hit.rgba.data(), hit.depth.data(), hit.payload.data(), &alive_counter, &hit_counter);
        }
        n_alive = alive_counter;
        if (n_alive == 0) break;
        ++n_rounds;
        const uint32_t n_steps = std::min(std::max(n / n_alive, 1u), 8u);         // MIN / MAX_STEPS_INBETWEEN_COMPACTION
        std::fill(coords.begin(), coords.end(), NerfCoordinate(Vector3f::Zero(), Vector3f::Zero(), 0.f));
        for (uint32_t i = 0; i < n_alive; ++i) {
          lifted::h_tid.x = i;
          generate_next_nerf_network_inputs(n_alive, box, Matrix3f::Identity(), box, Vector2f(f, f), cam.col(2), cur.payload.data(),
                                            PitchedPtr<NerfCoordinate>(coords.data(), 1), n_steps, bits.data(), 0, cone, nullptr);
        }
        const uint32_t n_elements = ((n_alive * n_steps + 127u) / 128u) * 128u;   // next_multiple(.., batch_size_granularity)
        out.assign((size_t)16 * n_elements, (network_precision_t)0.f);
        for (uint32_t e = 0; e < n_alive * n_steps; ++e) {                          // the analytic stand-in network
          const NerfCoordinate& c = coords[e];
          const float x = c.pos.p.x(), y = c.pos.p.y(), z = c.pos.p.z();
          const float m = fmaxf(fmaxf(fabsf(x - 0.5f), fabsf(y - 0.5f)), fabsf(z - 0.5f));
          const float raw[4] = {8.f * x - 4.f, 8.f * y - 4.f, 6.f * c.dir.d.z() - 3.f, 6.f - 14.f * m};
          for (int k = 0; k < 4; ++k) out[(size_t)k * n_elements + e] = (network_precision_t)raw[k];
        }
        for (uint32_t i = 0; i < n_alive; ++i) {
          lifted::h_tid.x = i;
          composite_kernel_nerf(n_alive, n_elements, step, box, 0.f, 0, 0, nullptr, cam, Vector2f(f, f), depth_scale, cur.rgba.data(),
                                cur.depth.data(), cur.payload.data(), PitchedPtr<NerfCoordinate>(coords.data(), 1), out.data(), 16, n_steps,
                                rmode, bits.data(), ENerfActivation::Logistic, ENerfActivation::Exponential, -1, min_T);
        }
        step += n_steps;
      }
      hits.push_back(hit_counter);
      rounds.push_back(n_rounds);
      // render_nerf: shade the rays that hit; render_frame: accumulate, tonemap
      for (uint32_t i = 0; i < hit_counter; ++i) {
        lifted::h_tid.x = i;
        shade_kernel_nerf(hit_counter, hit.rgba.data(), hit.depth.data(), hit.payload.data(), rmode, false, fb.data(), db.data());
      }
      for (int y = 0; y < res.y(); ++y)
        for (int x = 0; x < res.x(); ++x) {
          lifted::h_tid.x = (uint32_t)x; lifted::h_tid.y = (uint32_t)y;
          accumulate_kernel(res, fb.data(), acc.data(), (float)spp, EColorSpace::Linear);
        }
      lifted::h_tid.y = 0;
    }
    lifted::host_surface.assign(n, float4{0, 0, 0, 0});
    lifted::host_surface_w = res.x();
    for (int y = 0; y < res.y(); ++y)
      for (int x = 0; x < res.x(); ++x) {
        lifted::h_tid.x = (uint32_t)x; lifted::h_tid.y = (uint32_t)y;
        tonemap_kernel(res, 0.0f, background, acc.data(), EColorSpace::Linear, EColorSpace::Linear, ETonemapCurve::Identity, false, 0);
      }
    lifted::h_tid.y = 0;
    printf("%s{\"depth_mode\": %d, \"width\": %d, \"height\": %d, \"spp\": %d, \"fov\": %.9g, \"min_transmittance\": %.9g, \"depth_scale\": %.9g, \"camera\": [",
           mode ? ",\n" : "", mode, res.x(), res.y(), SPP, fov, min_T, depth_scale);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) printf("%s%.9g", (r || c) ? ", " : "", cam(r, c));
    printf("], \"n_hit\": [%u, %u, %u], \"rounds\": [%u, %u, %u], \"rgba\": [", hits[0], hits[1], hits[2], rounds[0], rounds[1], rounds[2]);
    for (uint32_t i = 0; i < n; ++i)
      printf("%s[%.9g, %.9g, %.9g, %.9g]", i ? ", " : "", lifted::host_surface[i].x, lifted::host_surface[i].y, lifted::host_surface[i].z, lifted::host_surface[i].w);
    printf("], \"depth\": [");
    for (uint32_t i = 0; i < n; ++i) printf("%s%.9g", i ? ", " : "", db[i]);
    printf("]}");
  }
  printf("],\n");
  printf("\"march\": [\n");
  for (int i = 0; i < 160; ++i) {
    // positions across the cascades (cascade c covers [0.5 - 2^(c-1), 0.5 + 2^(c-1)])
    const float half = 0.5f * (float)(1 << (i % 4));
    const Vector3f pos(0.5f + (rnd() * 2.f - 1.f) * half, 0.5f + (rnd() * 2.f - 1.f) * half, 0.5f + (rnd() * 2.f - 1.f) * half);
    Vector3f dir(rnd() * 2.f - 1.f, rnd() * 2.f - 1.f, rnd() * 2.f - 1.f);
    if (i % 16 == 5) dir.x() = 0.f;              // axis-parallel rays: idir = inf
    if (i % 16 == 11) { dir.y() = 0.f; dir.z() = 0.f; dir.x() = 1.f; }
    dir = (1.0f / dir.norm()) * dir;
    const Vector3f idir = dir.cwiseInverse();
    const float t = 0.05f + rnd() * 6.f;
    const float cone = (i & 1) ? 1.f / 256.f : 0.f;
    const float dt = calc_dt(t, cone);
    const int mp = mip_from_pos(pos), md = mip_from_dt(dt, pos);
    const uint32_t res = NERF_GRIDSIZE() >> md;
    printf("%s{", i ? ",\n" : "");
    vec3("pos", pos); vec3("dir", dir);
    printf("\"t\": %.9g, \"cone\": %.9g, \"dt\": %.9g, \"mip_from_pos\": %d, \"mip_from_dt\": %d, \"grid_idx\": %u, \"occupied\": %d, "
           "\"res\": %u, \"dist\": %.9g, \"advance\": %.9g}",
           t, cone, dt, mp, md, cascaded_grid_idx_at(pos, (uint32_t)md), (int)density_grid_occupied_at(pos, bits.data(), (uint32_t)md),
           res, distance_to_next_voxel(pos, dir, idir, res), advance_to_next_voxel(t, cone, pos, dir, idir, res));
  }
  printf("]\n}\n");
  return 0;
}
