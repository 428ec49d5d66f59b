// Instant-NGP reference-view render for sm_100a: ONE persistent launch per image does ray
// generation, occupancy-grid marching, the 16-level hash-grid lookup, the two tiny MLPs, alpha
// compositing, the spp accumulation and the tone-map epilogue -- no host round trip per march
// round (the reference reads `n_alive` back after every compaction, testbed_nerf.cu:2085-2086).
//
// Replaces (paths relative to /root/reference/instant-ngp):
//   Testbed::render_to_cpu                     src/python_api.cu:127-173
//   Testbed::render_frame / render_nerf        src/testbed.cu:2591-2749, src/testbed_nerf.cu:2228-2330
//   init_rays_with_payload_kernel_nerf         src/testbed_nerf.cu:1781-1890  (+ pixel_to_ray,
//                                              include/neural-graphics-primitives/common_device.cuh:260-307)
//   advance_pos_nerf                           :606-657
//   generate_next_nerf_network_inputs          :693-752
//   NerfNetwork::inference_mixed_precision_impl include/neural-graphics-primitives/nerf_network.h:101-136
//     kernel_grid          dependencies/tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:81-116,139-275
//     kernel_sh            .../encodings/spherical_harmonics.h:47-93
//     kernel_mlp_fused     dependencies/tiny-cuda-nn/src/fully_fused_mlp.cu:55-140,501-557
//   composite_kernel_nerf                      :754-955
//   compact_kernel_nerf / shade_kernel_nerf    :1721-1779
//   accumulate_kernel / tonemap_kernel         src/render_buffer.cu:236-275,542-570
// as driven by pixtrack/visualization/run_vis_on_poses.py:28-57 with the settings of
// pixtrack/utils/ingp_utils.py:22-44 (snap_to_pixel_centers, linear output, Shade or Depth mode).
//
// Mapping to the machine
//  * Every (pixel, sample) ray is independent; the reference's march rounds + stream compaction
//    only exist to keep its separate kernels busy.  Here a LANE owns a pixel and walks its spp rays
//    one after the other; a lane whose pixel is finished takes the next pixel from a global atomic
//    counter (warp-aggregated), so warps stay full without compaction and without the host.
//  * One warp evaluates the network for its 32 current samples: each lane gathers the 16 x 8 hash
//    grid corners of its own sample (4-byte random gathers; the 24-49 MB table is L2-resident on
//    B200's 126 MB L2), the 32 x 32 feature tile goes through shared memory into mma.sync
//    m16n8k16 fragments, and the five layers are chained in registers (the accumulator fragment of
//    one layer is the A fragment of the next).  Weights (20 KB fp16) sit in shared memory.
//    The layers are 32/64 wide: too thin for a 128-row tcgen05 tile per warp, and the kernel is
//    bound by the gathers and the marching, not by the MMAs.
//  * fp16 operands, fp32 accumulation, one rounding to fp16 per layer (tiny-cuda-nn accumulates in
//    fp16 inside wmma fragments; see oracle/nerf.py header).  Hash-grid interpolation accumulates in
//    fp16 exactly like kernel_grid.
//  * The marching / compositing arithmetic is compiled with -fmad=false (see build.py) so that it
//    is the same sequence of IEEE operations as the numpy oracle.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "ptk_common.cuh"

namespace {

constexpr int kGrid = 128;                 // NERF_GRIDSIZE
constexpr int kCascades = 8;               // NERF_CASCADES
constexpr float kNear = 0.05f;             // NERF_RENDERING_NEAR_DISTANCE
constexpr float kSqrt3 = 1.73205080757f;
constexpr float kMinStep = kSqrt3 / 1024.f;                                   // MIN_CONE_STEPSIZE
constexpr float kMaxStep = kMinStep * (1 << (kCascades - 1)) * 1024.f / kGrid;  // MAX_CONE_STEPSIZE
constexpr float kWarpDtSpan = kMinStep * (1 << (kCascades - 1)) - kMinStep;   // warp_dt / unwarp_dt
constexpr int kMaxSteps = 10000;           // MARCH_ITER
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;

// padded shared-memory strides (halfs): conflict-free 32-bit fragment loads
constexpr int kS32 = 40, kS64 = 72, kSsh = 24;
constexpr int kOffWd1 = 0;                         // [64][kS32]
constexpr int kOffWd2 = kOffWd1 + 64 * kS32;       // [16][kS64]
constexpr int kOffWc1 = kOffWd2 + 16 * kS64;       // [64][kS32]
constexpr int kOffWc2 = kOffWc1 + 64 * kS32;       // [64][kS64]
constexpr int kOffWc3 = kOffWc2 + 64 * kS64;       // [8][kS64]   (rows 0..2 = r, g, b)
constexpr int kWeightHalfs = kOffWc3 + 8 * kS64;
constexpr int kWarpHalfs = 32 * kS32 + 32 * kSsh;  // feature tile + SH tile
constexpr int kSmemBytes = kWeightHalfs * 2 + kWarpsPerCta * (kWarpHalfs * 2 + 32 * 4 * 4);

struct NerfLevel {
  float scale;
  uint32_t res;
  uint32_t offset;
  uint32_t size;
  uint32_t hashed;
};

struct NerfParams {
  const __half2* grid;
  const uint8_t* bitfield;
  const __half* w[5];            // density [64][32], [16][64]; rgb [64][32], [64][64], [16][64]
  NerfLevel lv[16];
  float cam[12];                 // 3x4 row-major, NGP convention
  float tmin[3], tmax[3];        // training box (m_aabb)
  float rmin[3], rmax[3];        // render box
  float occ[kCascades][6];       // U_m: bounds of the occupied cells of cascades 0..m, padded (see ptk_nerf_create)
  int corner_mip;                // mip_from_pos of the farthest render-box corner
  float cone, focal, depth_scale, one_minus_min_T;
  float bg[4];                   // background, already linear
  int W, H, spp, depth_mode;
  float4* out_rgba;              // [H][W] or null
  uint8_t* out_u8;               // [H][W][3] or null
  float* out_depth;              // [H][W] or null
  unsigned* counter;             // pixel queue
  float* starts;                 // [spp][H*W] first sample position of every ray, < 0: the ray is dead (nerf_start_kernel)
  float4* frames;                // [spp][H*W] shaded result of every live ray (nerf_render_kernel -> nerf_resolve_kernel)
  float* frame_depth;            // [H*W] depth of the last sample-per-pixel ray
  unsigned long long* stats;     // [4] network samples, warp steps, rays marched, lanes that sat out a warp step
};

struct V3 {
  float x, y, z;
};

__device__ __forceinline__ uint32_t rev32(uint32_t x) { return __brev(x); }
__device__ __forceinline__ uint32_t laine_karras(uint32_t x, uint32_t seed) {
  x += seed;
  x ^= x * 0x6c50b47cu;
  x ^= x * 0xb82f1e52u;
  x ^= x * 0xc7afe638u;
  x ^= x * 0x8d22f6e6u;
  return x;
}
__device__ __forceinline__ uint32_t owen(uint32_t x, uint32_t seed) { return rev32(laine_karras(rev32(x), seed)); }
// random_val.cuh:284-288 with dim 0, where sobol(i, 0) is the base-2 radical inverse.
__device__ __forceinline__ float ld_random_val(uint32_t index, uint32_t seed) {
  index = owen(index, seed);
  const uint32_t hc = seed ^ (0u + (seed << 6) + (seed >> 2));
  return (float)owen(rev32(index), hc) * (1.0f / 4294967296.0f);
}

__device__ __forceinline__ float calc_dt(float t, float cone) { return fmaxf(kMinStep, fminf(kMaxStep, t * cone)); }

__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

__device__ __forceinline__ int mip_from_pos(const V3& p) {
  const float mx = fmaxf(fmaxf(fabsf(p.x - 0.5f), fabsf(p.y - 0.5f)), fabsf(p.z - 0.5f));
  int e;
  frexpf(mx, &e);
  return min(kCascades - 1, max(0, e + 1));
}

__device__ __forceinline__ int mip_from_dt(float dt, const V3& p) {
  const int mip = mip_from_pos(p);
  dt *= 2 * kGrid;
  if (dt < 1.f) return mip;
  int e;
  frexpf(dt, &e);
  return min(kCascades - 1, max(e, mip));
}

__device__ __forceinline__ bool occupied(const uint8_t* __restrict__ bits, const V3& p, int mip) {
  const float s = scalbnf(1.0f, -mip);
  const int ix = (int)(((p.x - 0.5f) * s + 0.5f) * kGrid);
  const int iy = (int)(((p.y - 0.5f) * s + 0.5f) * kGrid);
  const int iz = (int)(((p.z - 0.5f) * s + 0.5f) * kGrid);
  const uint32_t idx = expand_bits(min(max(ix, 0), kGrid - 1)) | (expand_bits(min(max(iy, 0), kGrid - 1)) << 1) |
                       (expand_bits(min(max(iz, 0), kGrid - 1)) << 2);
  return (__ldg(bits + idx / 8 + (size_t)mip * (kGrid * kGrid * kGrid / 8)) >> (idx & 7)) & 1;
}

__device__ __forceinline__ bool inside(const float* lo, const float* hi, const V3& p) {
  return p.x >= lo[0] && p.x <= hi[0] && p.y >= lo[1] && p.y <= hi[1] && p.z >= lo[2] && p.z <= hi[2];
}

// Advance t to the next sample position in an occupied cell (common loop of advance_pos_nerf and
// generate_next_nerf_network_inputs).  Returns false when the ray left the render box.
//
// [tu0, tu1] is the part of the ray inside the padded bounds of everything occupied (U_m, see
// ptk_nerf_create).  Outside it no probe can find an occupied cell, so before tu0 the ray only needs
// its t sequence advanced (t += dt(t), the same additions the voxel stepping performs) and after tu1
// it can be retired at once -- same result as marching on to the box exit.
__device__ __forceinline__ bool skip_empty(const NerfParams& P, const V3& o, const V3& d, const V3& id, float tu0,
                                           float tu1, float& t, V3& pos, float& dt) {
  while (t < tu0) t += calc_dt(t, P.cone);
  while (true) {
    if (t > tu1) return false;
    pos.x = o.x + d.x * t;
    pos.y = o.y + d.y * t;
    pos.z = o.z + d.z * t;
    if (!inside(P.rmin, P.rmax, pos)) return false;
    dt = calc_dt(t, P.cone);
    const int mip = mip_from_dt(dt, pos);
    if (occupied(P.bitfield, pos, mip)) return true;
    const float res = (float)(kGrid >> mip);
    const float px = res * pos.x, py = res * pos.y, pz = res * pos.z;
    const float tx = (floorf(px + 0.5f + 0.5f * copysignf(1.f, d.x)) - px) * id.x;
    const float ty = (floorf(py + 0.5f + 0.5f * copysignf(1.f, d.y)) - py) * id.y;
    const float tz = (floorf(pz + 0.5f + 0.5f * copysignf(1.f, d.z)) - pz) * id.z;
    const float target = t + fmaxf(fminf(fminf(tx, ty), tz) / res, 0.f);
    do {
      t += calc_dt(t, P.cone);
    } while (t < target);
  }
}

__device__ __forceinline__ float srgb_to_linear(float s) {
  return s <= 0.04045f ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f);
}

// ---- network ---------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// C[16 x N] = A[16 x K] * W[N x K]^T for one m16 tile of a warp: A as K/16 fragments, W in shared memory with row
// stride WS halfs.  (One tile at a time: the accumulators of both tiles together cost 64 registers more, which is the
// difference between two and three resident CTAs per SM -- and this kernel lives on occupancy: it is bound by the
// latency of the hash-grid gathers.)
template <int K, int N, int WS>
__device__ __forceinline__ void layer(const uint32_t (&a)[K / 16][4], const __half* __restrict__ w, int lane,
                                      float (&c)[N / 8][4]) {
  const int n_in = lane >> 2, k_in = (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < N / 8; ++nt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) c[nt][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < K / 16; ++kk) {
      const __half* wr = w + (nt * 8 + n_in) * WS + kk * 16 + k_in;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
      mma16816(c[nt], a[kk], b0, b1);
    }
  }
}

// accumulator fragments of a layer (fp32) -> A fragments of the next one (fp16), optional ReLU
template <int N, bool kRelu>
__device__ __forceinline__ void to_a(const float (&c)[N / 8][4], uint32_t (&a)[N / 16][4]) {
#pragma unroll
  for (int kk = 0; kk < N / 16; ++kk)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v0 = c[2 * kk + h][0], v1 = c[2 * kk + h][1], v2 = c[2 * kk + h][2], v3 = c[2 * kk + h][3];
      if (kRelu) {
        v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
      }
      a[kk][2 * h] = pack_h2(v0, v1);
      a[kk][2 * h + 1] = pack_h2(v2, v3);
    }
}

template <int KS, int STRIDE>
__device__ __forceinline__ void load_a(const __half* __restrict__ tile, int lane, int k0, uint32_t (&a)[4], int mt) {
  const int r = mt * 16 + (lane >> 2), c = k0 + (lane & 3) * 2;
  a[0] = *reinterpret_cast<const uint32_t*>(tile + r * STRIDE + c);
  a[1] = *reinterpret_cast<const uint32_t*>(tile + (r + 8) * STRIDE + c);
  a[2] = *reinterpret_cast<const uint32_t*>(tile + r * STRIDE + c + 8);
  a[3] = *reinterpret_cast<const uint32_t*>(tile + (r + 8) * STRIDE + c + 8);
}

// hash-grid features of one sample -> 32 halfs in the warp's feature tile row (kernel_grid).
// Per level: the 8 corner indices (dense levels: x + y*res + z*res^2, wrapped once past the end like the
// reference's `% size`; hashed levels: x ^ y*2654435761 ^ z*805459861 masked to 2^19), 8 independent 4-byte
// gathers, then the trilinear blend accumulated in fp16 exactly like kernel_grid (weight * value rounded to
// fp16, added in fp16, corner order x fastest).
__device__ __forceinline__ void hash_level(const __half2* __restrict__ g, const uint32_t (&ix)[2], const uint32_t (&iy)[2],
                                           const uint32_t (&iz)[2], bool hashed, uint32_t size, float wx, float wy,
                                           float wz, __half* __restrict__ dst) {
  // (Tried: fetching the two x-neighbours of a corner pair with one 8-byte load when they share an aligned pair of
  // entries, plus a predicated 4-byte load otherwise -- a quarter fewer L1 line look-ups, but 10-14 % SLOWER on B200:
  // profiles/r2/nerf_paired_loads.json.  Eight independent 4-byte gathers it stays.)
  __half2 v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t idx;
    if (hashed) {
      idx = (ix[c & 1] ^ iy[(c >> 1) & 1] ^ iz[c >> 2]) & (size - 1u);   // hashed levels hold 2^19 entries
    } else {
      idx = ix[c & 1] + iy[(c >> 1) & 1] + iz[c >> 2];
      idx -= idx >= size ? size : 0u;                                    // == idx % size: idx < 2 * size here
    }
    v[c] = __ldg(g + idx);
  }
  const float ux = 1.f - wx, uy = 1.f - wy, uz = 1.f - wz;
  __half2 acc = __floats2half2_rn(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w = 1.f;
    w *= (c & 1) ? wx : ux;
    w *= (c & 2) ? wy : uy;
    w *= (c & 4) ? wz : uz;
    const float2 f = __half22float2(v[c]);
    acc = __hadd2(acc, __floats2half2_rn(w * f.x, w * f.y));
  }
  *reinterpret_cast<__half2*>(dst) = acc;
}

__device__ __forceinline__ void hash_encode(const NerfParams& P, float x, float y, float z, __half* __restrict__ row) {
#pragma unroll 2
  for (int l = 0; l < 16; ++l) {
    const NerfLevel lv = P.lv[l];
    const float fx = x * lv.scale + 0.5f, fy = y * lv.scale + 0.5f, fz = z * lv.scale + 0.5f;
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const uint32_t gx = (uint32_t)(int)flx, gy = (uint32_t)(int)fly, gz = (uint32_t)(int)flz;
    uint32_t ix[2], iy[2], iz[2];
    if (lv.hashed) {   // uniform per level
      ix[0] = gx; ix[1] = gx + 1u;
      iy[0] = gy * 2654435761u; iy[1] = (gy + 1u) * 2654435761u;
      iz[0] = gz * 805459861u; iz[1] = (gz + 1u) * 805459861u;
    } else {
      const uint32_t r2 = lv.res * lv.res;
      ix[0] = gx; ix[1] = gx + 1u;
      iy[0] = gy * lv.res; iy[1] = iy[0] + lv.res;
      iz[0] = gz * r2; iz[1] = iz[0] + r2;
    }
    hash_level(P.grid + lv.offset, ix, iy, iz, lv.hashed != 0, lv.size, fx - flx, fy - fly, fz - flz, row + 2 * l);
  }
}

// degree-4 spherical harmonics of the direction (kernel_sh); the direction goes through
// warp_direction / the "* 2 - 1" of the encoder like in the reference
__device__ __forceinline__ void sh_encode(const V3& d, __half* __restrict__ row) {
  const float x = ((d.x + 1.f) * 0.5f) * 2.f - 1.f, y = ((d.y + 1.f) * 0.5f) * 2.f - 1.f, z = ((d.z + 1.f) * 0.5f) * 2.f - 1.f;
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  float o[16];
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
#pragma unroll
  for (int i = 0; i < 16; i += 2) *reinterpret_cast<__half2*>(row + i) = __floats2half2_rn(o[i], o[i + 1]);
}

__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

// Network for the warp's 32 samples.  feat / sh tiles are filled by the lanes; out[lane] =
// (raw r, raw g, raw b, raw density), each rounded to fp16 like the reference's network output.  The two m16 tiles of
// the warp go through the five layers one after the other (see `layer`).
__device__ __forceinline__ void run_network(const __half* __restrict__ wts, const __half* __restrict__ feat,
                                            const __half* __restrict__ sh, float4* __restrict__ out, int lane,
                                            bool want_rgb) {
  const int r = lane >> 2, q = lane & 3;
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    uint32_t a32[2][4];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) load_a<32, kS32>(feat, lane, kk * 16, a32[kk], mt);
    uint32_t a64[4][4];
    {
      float c[8][4];
      layer<32, 64, kS32>(a32, wts + kOffWd1, lane, c);
      to_a<64, true>(c, a64);
    }
    float cd[2][4];
    layer<64, 16, kS64>(a64, wts + kOffWd2, lane, cd);
    float4* o = out + mt * 16;
    if (q == 0) {
      o[r].w = round_h(cd[0][0]);
      o[r + 8].w = round_h(cd[0][2]);
    }
    if (!want_rgb) continue;
    {
      uint32_t ad[1][4];
      to_a<16, false>(cd, ad);
#pragma unroll
      for (int i = 0; i < 4; ++i) a32[0][i] = ad[0][i];
      load_a<16, kSsh>(sh, lane, 0, a32[1], mt);
    }
    {
      float c[8][4];
      layer<32, 64, kS32>(a32, wts + kOffWc1, lane, c);
      to_a<64, true>(c, a64);
    }
    {
      float c[8][4];
      layer<64, 64, kS64>(a64, wts + kOffWc2, lane, c);
      to_a<64, true>(c, a64);
    }
    float c3[1][4];
    layer<64, 8, kS64>(a64, wts + kOffWc3, lane, c3);
    if (q == 0) {
      o[r].x = round_h(c3[0][0]); o[r].y = round_h(c3[0][1]);
      o[r + 8].x = round_h(c3[0][2]); o[r + 8].y = round_h(c3[0][3]);
    } else if (q == 1) {
      o[r].z = round_h(c3[0][0]);
      o[r + 8].z = round_h(c3[0][2]);
    }
  }
}

// ---- the render kernel ---------------------------------------------------------------------------
struct Ray {
  int pix;          // -1: no ray
  int s;            // sample-per-pixel index of this ray
  int steps;
  bool alive;
  V3 d, id;
  float t, tu0, tu1;
  float r, g, b, a; // compositing state
  float maxw, dep;
};

// Ray of a pixel (pixel_to_ray with snap_to_pixel_centers: offset 0.5, screen centre 0.5, no parallax), its entry
// into the render box (BoundingBox::ray_intersect) and the part [tu0, tu1] inside the occupied bounds of every
// cascade it can probe (tu0 > tu1: nothing occupied along the ray).
__device__ __forceinline__ void setup_ray(const NerfParams& P, const V3& o, int pix, V3& dout, V3& idout, float& tentry,
                                          float& tu0, float& tu1) {
  const int px = pix % P.W, py = pix / P.W;
  // pixel_to_ray with snap_to_pixel_centers (offset 0.5), screen centre 0.5, no parallax
  const float u = ((float)px + 0.5f) / (float)P.W, v = ((float)py + 0.5f) / (float)P.H;
  const float cx = (u - 0.5f) * (float)P.W / P.focal, cy = (v - 0.5f) * (float)P.H / P.focal;
  V3 d;
  d.x = (cx * P.cam[0] + cy * P.cam[1]) + P.cam[2];
  d.y = (cx * P.cam[4] + cy * P.cam[5]) + P.cam[6];
  d.z = (cx * P.cam[8] + cy * P.cam[9]) + P.cam[10];
  const float nrm = sqrtf((d.x * d.x + d.y * d.y) + d.z * d.z);
  d.x /= nrm; d.y /= nrm; d.z /= nrm;
  dout = d;
  idout.x = 1.f / d.x; idout.y = 1.f / d.y; idout.z = 1.f / d.z;
  // BoundingBox::ray_intersect
  float t0 = (P.rmin[0] - o.x) / d.x, t1 = (P.rmax[0] - o.x) / d.x;
  float tmin = fminf(t0, t1), tmax = fmaxf(t0, t1);
  bool miss = false;
  t0 = (P.rmin[1] - o.y) / d.y; t1 = (P.rmax[1] - o.y) / d.y;
  float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
  miss = miss || (tmin > hi) || (lo > tmax);
  tmin = lo > tmin ? lo : tmin; tmax = hi < tmax ? hi : tmax;
  t0 = (P.rmin[2] - o.z) / d.z; t1 = (P.rmax[2] - o.z) / d.z;
  lo = fminf(t0, t1); hi = fmaxf(t0, t1);
  miss = miss || (tmin > hi) || (lo > tmax);
  tmin = lo > tmin ? lo : tmin; tmax = hi < tmax ? hi : tmax;
  tentry = miss ? 3.402823466e+38f : tmin;
  // part of the ray inside the occupied bounds of every cascade it can probe
  tu0 = 3.402823466e+38f;
  tu1 = -3.402823466e+38f;
  if (!miss) {
    const float far = fmaxf(tmax, kNear) + 1.f;
    int mh = P.corner_mip;
    {
      const float dtf = calc_dt(far, P.cone) * (2 * kGrid);
      int e = 0;
      if (dtf >= 1.f) frexpf(dtf, &e);
      mh = min(kCascades - 1, max(mh, e));
    }
    const float* U = P.occ[mh];
    float a0 = (U[0] - o.x) / d.x, a1 = (U[3] - o.x) / d.x;
    float un = fminf(a0, a1), ux = fmaxf(a0, a1);
    a0 = (U[1] - o.y) / d.y; a1 = (U[4] - o.y) / d.y;
    un = fmaxf(un, fminf(a0, a1)); ux = fminf(ux, fmaxf(a0, a1));
    a0 = (U[2] - o.z) / d.z; a1 = (U[5] - o.z) / d.z;
    un = fmaxf(un, fminf(a0, a1)); ux = fminf(ux, fmaxf(a0, a1));
    if (un <= ux && U[0] <= U[3]) {   // NaN-free hit (fminf/fmaxf drop NaNs of axis-parallel rays)
      tu0 = un;
      tu1 = ux;
    }
  }
}

// advance_pos_nerf for one (pixel, sample) ray: jittered start inside the render box, then on to the first
// occupied cell.  Returns the ray parameter of that position, or -1 when the ray dies before reaching one.
__device__ __forceinline__ float first_sample(const NerfParams& P, const V3& o, const V3& d, const V3& id, float tentry,
                                              float tu0, float tu1, int pix, int s) {
  float t = fmaxf(tentry, kNear) + 1e-6f;
  V3 p;
  p.x = o.x + d.x * t;
  p.y = o.y + d.y * t;
  p.z = o.z + d.z * t;
  if (!inside(P.rmin, P.rmax, p)) return -1.f;
  t += ld_random_val((uint32_t)s, (uint32_t)pix * 786433u) * calc_dt(t, P.cone);
  float dt;
  return skip_empty(P, o, d, id, tu0, tu1, t, p, dt) ? t : -1.f;
}

// A ray is finished: compact_kernel_nerf keeps it only above alpha 0.001, shade_kernel_nerf converts the colour to
// linear; the result goes to the per-ray frame buffer that nerf_resolve_kernel averages in sample order.
__device__ __forceinline__ void finish_ray(const NerfParams& P, const Ray& ry) {
  float fr = 0.f, fg = 0.f, fb = 0.f, fa = 0.f, fd = 0.f;
  if (ry.a > 0.001f) {
    fr = ry.r; fg = ry.g; fb = ry.b; fa = ry.a;
    if (!P.depth_mode) {
      fr = srgb_to_linear(fr); fg = srgb_to_linear(fg); fb = srgb_to_linear(fb);
    }
    if (ry.a > 0.2f) fd = ry.dep;
  }
  const size_t npix = (size_t)P.W * P.H;
  P.frames[(size_t)ry.s * npix + ry.pix] = make_float4(fr, fg, fb, fa);
  if (ry.s == P.spp - 1) P.frame_depth[ry.pix] = fd;
}

// Phase 1: the first sample position of every (pixel, sample) ray.  One thread per pixel, so the warps walk
// neighbouring rays through the empty space in front of the object together; inside the render kernel the
// same search would run whenever a lane starts a new ray, i.e. with one or two lanes of the warp active
// (measured: 2.1 active lanes, 57 % of all issued instructions).
__global__ void __launch_bounds__(256) nerf_start_kernel(const __grid_constant__ NerfParams P) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int npix = P.W * P.H;
  if (pix >= npix) return;
  const V3 o = {P.cam[3], P.cam[7], P.cam[11]};
  V3 d, id;
  float tentry, tu0, tu1;
  setup_ray(P, o, pix, d, id, tentry, tu0, tu1);
  const bool any = tu0 <= tu1;
  for (int s = 0; s < P.spp; ++s)
    P.starts[(size_t)s * npix + pix] = any ? first_sample(P, o, d, id, tentry, tu0, tu1, pix, s) : -1.f;
}

// Phase 3: accumulate_kernel (running mean over the samples, in sample order) + tonemap_kernel (background weighted
// by 1 - alpha, linear in / linear out, identity curve) + the uint8 conversion of get_nerf_image, one thread per pixel.
__global__ void __launch_bounds__(256) nerf_resolve_kernel(const __grid_constant__ NerfParams P) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int npix = P.W * P.H;
  if (pix >= npix) return;
  float ar = 0.f, ag = 0.f, ab = 0.f, aa = 0.f;
  bool last_alive = false;
  for (int s = 0; s < P.spp; ++s) {
    const bool live = P.starts[(size_t)s * npix + pix] >= 0.f;
    const float4 f = live ? P.frames[(size_t)s * npix + pix] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float n = (float)s;
    ar = (ar * n + f.x) / (n + 1.f);
    ag = (ag * n + f.y) / (n + 1.f);
    ab = (ab * n + f.z) / (n + 1.f);
    aa = (aa * n + f.w) / (n + 1.f);
    last_alive = live;
  }
  const float w = (1.f - aa) * P.bg[3];
  const float r = ar + P.bg[0] * w, g = ag + P.bg[1] * w, b = ab + P.bg[2] * w, a = aa + w;
  if (P.out_rgba) P.out_rgba[pix] = make_float4(r, g, b, a);
  if (P.out_u8) {   // run_vis_on_poses.py:52-54: (rgb * 255).astype(uint8)
    uint8_t* o = P.out_u8 + (size_t)pix * 3;
    o[0] = (uint8_t)((int)(r * 255.f) & 0xFF);
    o[1] = (uint8_t)((int)(g * 255.f) & 0xFF);
    o[2] = (uint8_t)((int)(b * 255.f) & 0xFF);
  }
  if (P.out_depth) P.out_depth[pix] = last_alive ? P.frame_depth[pix] : 0.f;
}

// kCtas = resident CTAs per SM the register allocation is bounded for (2: 128 registers, 3: 80).
template <int kCtas>
__global__ void __launch_bounds__(kThreads, kCtas) nerf_render_kernel(const __grid_constant__ NerfParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  __half* wts = reinterpret_cast<__half*>(smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* wbase = smem + kWeightHalfs * 2 + warp * (kWarpHalfs * 2 + 32 * 16);
  __half* feat = reinterpret_cast<__half*>(wbase);
  __half* sh = feat + 32 * kS32;
  float4* outv = reinterpret_cast<float4*>(wbase + kWarpHalfs * 2);

  // weights -> padded shared memory
  for (int i = threadIdx.x; i < 64 * 32; i += kThreads) {
    wts[kOffWd1 + (i >> 5) * kS32 + (i & 31)] = P.w[0][i];
    wts[kOffWc1 + (i >> 5) * kS32 + (i & 31)] = P.w[2][i];
  }
  for (int i = threadIdx.x; i < 16 * 64; i += kThreads) wts[kOffWd2 + (i >> 6) * kS64 + (i & 63)] = P.w[1][i];
  for (int i = threadIdx.x; i < 64 * 64; i += kThreads) wts[kOffWc2 + (i >> 6) * kS64 + (i & 63)] = P.w[3][i];
  for (int i = threadIdx.x; i < 8 * 64; i += kThreads) wts[kOffWc3 + (i >> 6) * kS64 + (i & 63)] = P.w[4][i];
  __syncthreads();

  const unsigned full = 0xffffffffu;
  const int npix = P.W * P.H;
  const unsigned nrays = (unsigned)npix * (unsigned)P.spp;
  const V3 o = {P.cam[3], P.cam[7], P.cam[11]};
  const V3 fwd = {P.cam[2], P.cam[6], P.cam[10]};
  Ray ry;
  ry.pix = -1;
  ry.alive = false;
  bool exhausted = false;
  unsigned st_samples = 0, st_steps = 0, st_rays = 0;   // per-warp statistics (identical in every lane; lane 0 reports)

  while (true) {
    // ---- take new rays until every lane has a live one or the queue is empty.  The work unit is a (pixel, sample)
    //      ray, pixel-major (the spp rays of a pixel follow the same path and run side by side in a warp); rays
    //      that nerf_start_kernel found dead cost one load here.  Ray-sized units balance the lanes much better
    //      than whole pixels (an object covers ~1e5 pixels, the grid has ~7.5e4 lanes). -----------------------
#pragma unroll 1
    while (true) {
      const bool need = ry.pix < 0 && !exhausted;
      const unsigned m = __ballot_sync(full, need);
      if (m == 0) break;
      unsigned base = 0;
      if (lane == __ffs(m) - 1) base = atomicAdd(P.counter, (unsigned)__popc(m));
      base = __shfl_sync(full, base, __ffs(m) - 1);
      if (need) {
        const unsigned mine = base + __popc(m & ((1u << lane) - 1u));
        if (mine < nrays) {
          const int pix = (int)(mine / (unsigned)P.spp), sidx = (int)(mine - (unsigned)pix * (unsigned)P.spp);
          const float t = __ldg(P.starts + (size_t)sidx * (size_t)npix + pix);
          if (t >= 0.f) {
            ++st_rays;
            float tentry;
            setup_ray(P, o, pix, ry.d, ry.id, tentry, ry.tu0, ry.tu1);
            ry.pix = pix;
            ry.s = sidx;
            ry.t = t;
            ry.alive = true;
            ry.steps = 1;
            ry.r = ry.g = ry.b = ry.a = 0.f;
            ry.maxw = ry.dep = 0.f;
          }
        } else {
          exhausted = true;
        }
      }
    }
    if (__ballot_sync(full, ry.pix >= 0) == 0) break;

    // ---- next sample of every live ray ------------------------------------------------------------
    bool sample = false;
    V3 pos = {0.f, 0.f, 0.f};
    float dt = 0.f;
    if (ry.pix >= 0 && ry.alive) {
      sample = skip_empty(P, o, ry.d, ry.id, ry.tu0, ry.tu1, ry.t, pos, dt);
      if (!sample) ry.alive = false;
    }
    const unsigned sm = __ballot_sync(full, sample);
    if (sm != 0) {
      st_samples += (unsigned)__popc(sm);
      ++st_steps;
      // warp_position -> network input in the unit cube of the training box
      const float wx = (pos.x - P.tmin[0]) / (P.tmax[0] - P.tmin[0]);
      const float wy = (pos.y - P.tmin[1]) / (P.tmax[1] - P.tmin[1]);
      const float wz = (pos.z - P.tmin[2]) / (P.tmax[2] - P.tmin[2]);
      if (sample) {
        hash_encode(P, wx, wy, wz, feat + lane * kS32);
        if (!P.depth_mode) sh_encode(ry.d, sh + lane * kSsh);
      } else {
        uint4* z = reinterpret_cast<uint4*>(feat + lane * kS32);
        z[0] = z[1] = z[2] = z[3] = make_uint4(0, 0, 0, 0);
        uint4* zs = reinterpret_cast<uint4*>(sh + lane * kSsh);
        zs[0] = zs[1] = make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
      run_network(wts, feat, sh, outv, lane, P.depth_mode == 0);
      __syncwarp();
      if (sample) {   // composite_kernel_nerf
        const float4 raw = outv[lane];
        ry.t += dt;
        const float ux = P.tmin[0] + wx * (P.tmax[0] - P.tmin[0]);     // unwarp_position
        const float uy = P.tmin[1] + wy * (P.tmax[1] - P.tmin[1]);
        const float uz = P.tmin[2] + wz * (P.tmax[2] - P.tmin[2]);
        const float udt = ((dt - kMinStep) / kWarpDtSpan) * kWarpDtSpan + kMinStep;   // unwarp_dt(warp_dt(dt))
        const float T = 1.f - ry.a;
        const float alpha = 1.f - __expf(-__expf(raw.w) * udt);
        const float weight = alpha * T;
        float cr, cg, cb;
        if (P.depth_mode) {
          cr = cg = cb = ((fwd.x * (ux - o.x) + fwd.y * (uy - o.y)) + fwd.z * (uz - o.z)) * P.depth_scale;
        } else {
          cr = 1.f / (1.f + expf(-raw.x));
          cg = 1.f / (1.f + expf(-raw.y));
          cb = 1.f / (1.f + expf(-raw.z));
        }
        ry.r += cr * weight;
        ry.g += cg * weight;
        ry.b += cb * weight;
        ry.a += weight;
        if (weight > ry.maxw) {
          ry.maxw = weight;
          ry.dep = (fwd.x * (ux - o.x) + fwd.y * (uy - o.y)) + fwd.z * (uz - o.z);
        }
        if (ry.a > P.one_minus_min_T) {
          ry.r /= ry.a; ry.g /= ry.a; ry.b /= ry.a; ry.a /= ry.a;
          ry.alive = false;
        }
        if (++ry.steps >= kMaxSteps) ry.alive = false;
      }
      __syncwarp();
    }
    if (ry.pix >= 0 && !ry.alive) {
      finish_ray(P, ry);
      ry.pix = -1;
    }
  }
  {   // statistics of the render (ptk_nerf_stats): one set of atomics per warp at exit
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) st_rays += __shfl_xor_sync(full, st_rays, m);
    if (lane == 0) {
      atomicAdd(P.stats + 0, (unsigned long long)st_samples);
      atomicAdd(P.stats + 1, (unsigned long long)st_steps);
      atomicAdd(P.stats + 2, (unsigned long long)st_rays);
      atomicAdd(P.stats + 3, (unsigned long long)st_steps * 32ull - (unsigned long long)st_samples);
    }
  }
}

// Network alone (NerfNetwork::inference_mixed_precision_impl, nerf_network.h:101-136): one warp per 32 inputs, the same
// hash_encode / sh_encode / run_network the render kernel uses.  pos01 [n][3] = positions in the unit cube of the training
// box (warp_position applied), dir [n][3] = unit view directions; out [n][4] = raw r, g, b and raw density (fp16 values).
__global__ void __launch_bounds__(kThreads) nerf_eval_kernel(const __grid_constant__ NerfParams P, const float* __restrict__ pos01,
                                                             const float* __restrict__ dir, int n, float4* __restrict__ out,
                                                             float* __restrict__ feat_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  __half* wts = reinterpret_cast<__half*>(smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* wbase = smem + kWeightHalfs * 2 + warp * (kWarpHalfs * 2 + 32 * 16);
  __half* feat = reinterpret_cast<__half*>(wbase);
  __half* sh = feat + 32 * kS32;
  float4* outv = reinterpret_cast<float4*>(wbase + kWarpHalfs * 2);
  for (int i = threadIdx.x; i < 64 * 32; i += kThreads) {
    wts[kOffWd1 + (i >> 5) * kS32 + (i & 31)] = P.w[0][i];
    wts[kOffWc1 + (i >> 5) * kS32 + (i & 31)] = P.w[2][i];
  }
  for (int i = threadIdx.x; i < 16 * 64; i += kThreads) wts[kOffWd2 + (i >> 6) * kS64 + (i & 63)] = P.w[1][i];
  for (int i = threadIdx.x; i < 64 * 64; i += kThreads) wts[kOffWc2 + (i >> 6) * kS64 + (i & 63)] = P.w[3][i];
  for (int i = threadIdx.x; i < 8 * 64; i += kThreads) wts[kOffWc3 + (i >> 6) * kS64 + (i & 63)] = P.w[4][i];
  __syncthreads();
  for (int base = (blockIdx.x * kWarpsPerCta + warp) * 32; base < n; base += gridDim.x * kWarpsPerCta * 32) {
    const int i = base + lane;
    if (i < n) {
      hash_encode(P, pos01[3 * i], pos01[3 * i + 1], pos01[3 * i + 2], feat + lane * kS32);
      const V3 d = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
      sh_encode(d, sh + lane * kSsh);
    } else {
      uint4* z = reinterpret_cast<uint4*>(feat + lane * kS32);
      z[0] = z[1] = z[2] = z[3] = make_uint4(0, 0, 0, 0);
      uint4* zs = reinterpret_cast<uint4*>(sh + lane * kSsh);
      zs[0] = zs[1] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    if (feat_out != nullptr && i < n)
      for (int k = 0; k < 32; ++k) feat_out[(size_t)i * 32 + k] = __half2float(feat[lane * kS32 + k]);
    run_network(wts, feat, sh, outv, lane, true);
    __syncwarp();
    if (i < n) out[i] = outv[lane];
    __syncwarp();
  }
}

}  // namespace

struct PtkNerf {
  PtkContext* ctx;
  NerfParams base;
  unsigned* counter;             // [2 + 2 * 4]: ray queue counter, pad, statistics (4 x u64)
  float* starts;        // workspace of the largest render so far: per-ray frames, first-sample positions, depth
  size_t starts_cap;    // floats
  int aabb_scale;
};

extern "C" int ptk_nerf_create(PtkContext* ctx, const PtkNerfModel* m, PtkNerf** out) {
  PTK_REQUIRE(ctx && m && out, "null argument");
  PtkDeviceGuard guard(ctx->device);
  PTK_REQUIRE(m->grid && m->bitfield, "null grid / bitfield");
  for (int i = 0; i < 5; ++i) PTK_REQUIRE(m->weights[i] != nullptr, "null weight matrix");
  PTK_REQUIRE(m->aabb_scale >= 1 && m->aabb_scale <= 128 && (m->aabb_scale & (m->aabb_scale - 1)) == 0,
              "aabb_scale must be a power of two in [1, 128]");
  PtkNerf* n = (PtkNerf*)calloc(1, sizeof(PtkNerf));
  n->ctx = ctx;
  n->aabb_scale = m->aabb_scale;
  NerfParams& P = n->base;
  P.grid = (const __half2*)m->grid;
  P.bitfield = m->bitfield;
  for (int i = 0; i < 5; ++i) P.w[i] = (const __half*)m->weights[i];
  // hash-grid layout: testbed.cu:2233-2244 (per_level_scale), grid.h:898-930 (offsets)
  const float pls = expf(logf(2048.0f * (float)m->aabb_scale / 16.0f) / 15.0f);
  const float log2pls = log2f(pls);
  uint32_t off = 0;
  for (int l = 0; l < 16; ++l) {
    const float scale = exp2f((float)l * log2pls) * 16.0f - 1.0f;
    const uint32_t res = (uint32_t)ceilf(scale) + 1u;
    const double cube = (double)res * res * res;
    uint64_t cnt = cube > 2147483647.0 ? 2147483647ull : (uint64_t)cube;
    cnt = (cnt + 7) / 8 * 8;
    if (cnt > (1u << 19)) cnt = 1u << 19;
    P.lv[l].scale = scale;
    P.lv[l].res = res;
    P.lv[l].offset = off;
    P.lv[l].size = (uint32_t)cnt;
    P.lv[l].hashed = cube > (double)cnt ? 1u : 0u;
    off += (uint32_t)cnt;
  }
  if (m->n_grid_entries != (int64_t)off) {
    ptk_set_error("hash grid has %lld entries, aabb_scale %d needs %u", (long long)m->n_grid_entries, m->aabb_scale, off);
    free(n);
    return PTK_ERR_INVALID;
  }
  const float half = 0.5f * (float)(m->aabb_scale < 128 ? m->aabb_scale : 128);
  for (int i = 0; i < 3; ++i) {
    P.tmin[i] = 0.5f - half;
    P.tmax[i] = 0.5f + half;
  }
  P.cone = m->aabb_scale <= 1 ? 0.f : 1.f / 256.f;   // testbed_nerf.cu:2596
  {
    // Bounds of the occupied cells (load-time, host): B_m = box of the set cells of cascade m in world
    // coordinates; U_m = box of B_0..B_m padded by two cells of cascade m.  A probe at cascade <= m outside U_m
    // cannot be occupied, which lets the marcher skip the voxel walk there (skip_empty).
    const size_t bytes = (size_t)kCascades * kGrid * kGrid * kGrid / 8;
    uint8_t* host = (uint8_t*)malloc(bytes);
    cudaError_t ce = cudaMemcpy(host, m->bitfield, bytes, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) {
      ptk_set_error("reading the occupancy bitfield: %s", cudaGetErrorString(ce));
      free(host);
      free(n);
      return PTK_ERR_CUDA;
    }
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    const size_t per = (size_t)kGrid * kGrid * kGrid;
    for (int mip = 0; mip < kCascades; ++mip) {
      const float cs = ldexpf(1.f, mip) / kGrid;
      for (size_t i = 0; i < per; ++i) {
        if (((host[(mip * per + i) >> 3] >> (i & 7)) & 1) == 0) continue;
        uint32_t c[3];
        for (int a = 0; a < 3; ++a) {   // Morton decode
          uint32_t x = ((uint32_t)i >> a) & 0x49249249u;
          x = (x | (x >> 2)) & 0xc30c30c3u;
          x = (x | (x >> 4)) & 0x0f00f00fu;
          x = (x | (x >> 8)) & 0xff0000ffu;
          x = (x | (x >> 16)) & 0x0000ffffu;
          c[a] = x;
        }
        for (int a = 0; a < 3; ++a) {
          const float w0 = ((float)c[a] / kGrid - 0.5f) * ldexpf(1.f, mip) + 0.5f;
          if (w0 < lo[a]) lo[a] = w0;
          if (w0 + cs > hi[a]) hi[a] = w0 + cs;
        }
      }
      const float pad = 2.f * cs + 1e-4f;
      for (int a = 0; a < 3; ++a) {
        P.occ[mip][a] = lo[a] - pad;
        P.occ[mip][3 + a] = hi[a] + pad;
      }
    }
    free(host);
  }
  cudaError_t e = cudaMalloc(&n->counter, sizeof(unsigned) * 2 + sizeof(unsigned long long) * 4);
  if (e != cudaSuccess) {
    ptk_set_error("cudaMalloc: %s", cudaGetErrorString(e));
    free(n);
    return PTK_ERR_CUDA;
  }
  PTK_CUDA_CHECK(cudaFuncSetAttribute(nerf_render_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  PTK_CUDA_CHECK(cudaFuncSetAttribute(nerf_render_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  *out = n;
  return PTK_OK;
}

extern "C" void ptk_nerf_destroy(PtkNerf* n) {
  if (n == nullptr) return;
  if (n->counter) cudaFree(n->counter);
  if (n->starts) cudaFree(n->starts);
  free(n);
}

extern "C" int ptk_nerf_eval(PtkNerf* n, const float* pos01, const float* dir, int32_t count, float* out_rgbd,
                             float* out_features, void* stream) {
  PTK_REQUIRE(n && out_rgbd && (count == 0 || (pos01 && dir)) && count >= 0, "bad argument");
  if (count == 0) return PTK_OK;
  PtkDeviceGuard guard(n->ctx->device);
  static bool configured = false;
  if (!configured) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(nerf_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  long long ctas = ((long long)count + kThreads - 1) / kThreads;
  const long long cap = 2LL * n->ctx->num_sms;
  if (ctas > cap) ctas = cap;
  nerf_eval_kernel<<<(unsigned)ctas, kThreads, kSmemBytes, (cudaStream_t)stream>>>(n->base, pos01, dir, count,
                                                                                 reinterpret_cast<float4*>(out_rgbd), out_features);
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}

// SYNCHRONISING: statistics of the last render on this object (benchmarks / profiles).
extern "C" int ptk_nerf_stats(PtkNerf* n, uint64_t* out4) {
  PTK_REQUIRE(n && out4, "null argument");
  PtkDeviceGuard guard(n->ctx->device);
  PTK_CUDA_CHECK(cudaDeviceSynchronize());
  PTK_CUDA_CHECK(cudaMemcpy(out4, n->counter + 2, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost));
  return PTK_OK;
}

extern "C" int64_t ptk_nerf_grid_entries(int32_t aabb_scale) {
  const float pls = expf(logf(2048.0f * (float)aabb_scale / 16.0f) / 15.0f);
  const float log2pls = log2f(pls);
  uint64_t off = 0;
  for (int l = 0; l < 16; ++l) {
    const float scale = exp2f((float)l * log2pls) * 16.0f - 1.0f;
    const uint32_t res = (uint32_t)ceilf(scale) + 1u;
    const double cube = (double)res * res * res;
    uint64_t cnt = cube > 2147483647.0 ? 2147483647ull : (uint64_t)cube;
    cnt = (cnt + 7) / 8 * 8;
    if (cnt > (1u << 19)) cnt = 1u << 19;
    off += cnt;
  }
  return (int64_t)off;
}

extern "C" int ptk_nerf_render(PtkNerf* n, const PtkNerfView* v, float* out_rgba, uint8_t* out_u8, float* out_depth,
                               void* stream) {
  PTK_REQUIRE(n && v, "null argument");
  PTK_REQUIRE(v->width >= 1 && v->height >= 1 && v->spp >= 1, "width, height, spp must be >= 1");
  PTK_REQUIRE(v->focal > 0.f, "focal must be positive");
  PTK_REQUIRE(out_rgba || out_u8, "no output buffer");
  PtkDeviceGuard guard(n->ctx->device);
  NerfParams P = n->base;
  for (int i = 0; i < 12; ++i) P.cam[i] = v->camera[i];
  for (int i = 0; i < 3; ++i) {
    P.rmin[i] = v->render_aabb_min[i];
    P.rmax[i] = v->render_aabb_max[i];
  }
  {
    float mx = 0.f;
    for (int i = 0; i < 3; ++i) mx = fmaxf(mx, fmaxf(fabsf(P.rmin[i] - 0.5f), fabsf(P.rmax[i] - 0.5f)));
    int e = 0;
    frexpf(mx, &e);
    P.corner_mip = e + 1 < 0 ? 0 : (e + 1 > kCascades - 1 ? kCascades - 1 : e + 1);
  }
  P.focal = v->focal;
  P.depth_scale = v->depth_scale;
  P.one_minus_min_T = 1.0f - v->min_transmittance;
  for (int i = 0; i < 4; ++i) P.bg[i] = v->background[i];
  P.W = v->width;
  P.H = v->height;
  P.spp = v->spp;
  P.depth_mode = v->depth_mode;
  P.out_rgba = (float4*)out_rgba;
  P.out_u8 = out_u8;
  P.out_depth = out_depth;
  P.counter = n->counter;
  P.stats = reinterpret_cast<unsigned long long*>(n->counter + 2);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t npix = (size_t)v->width * v->height;
  const size_t need = npix * v->spp;
  const size_t floats = 4 * need + need + npix + 16;       // frames (float4), starts, last-sample depth
  if (floats > n->starts_cap) {   // grows only when a larger view is rendered (synchronising free)
    if (n->starts) PTK_CUDA_CHECK(cudaFree(n->starts));
    n->starts = nullptr;
    n->starts_cap = 0;
    PTK_CUDA_CHECK(cudaMalloc(&n->starts, floats * sizeof(float)));
    n->starts_cap = floats;
  }
  P.frames = reinterpret_cast<float4*>(n->starts);           // 16-byte aligned start of the workspace
  P.starts = n->starts + 4 * need;
  P.frame_depth = P.starts + need;
  PTK_CUDA_CHECK(cudaMemsetAsync(n->counter, 0, sizeof(unsigned) * 2 + sizeof(unsigned long long) * 4, s));
  nerf_start_kernel<<<(unsigned)(((size_t)v->width * v->height + 255) / 256), 256, 0, s>>>(P);
  static int want_ctas = -1;   // PTK_NERF_CTAS = 2 | 3 resident CTAs per SM (register bound of the instantiation used)
  if (want_ctas < 0) {
    const char* e = getenv("PTK_NERF_CTAS");
    want_ctas = (e != nullptr && atoi(e) == 2) ? 2 : 3;
  }
  int per_sm = 1;
  if (want_ctas == 2) PTK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nerf_render_kernel<2>, kThreads, kSmemBytes));
  else PTK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nerf_render_kernel<3>, kThreads, kSmemBytes));
  if (per_sm < 1) per_sm = 1;
  const long long warps_needed = ((long long)v->width * v->height + 31) / 32;
  long long ctas = (warps_needed + kWarpsPerCta - 1) / kWarpsPerCta;
  const long long cap = (long long)n->ctx->num_sms * per_sm;
  if (ctas > cap) ctas = cap;
  if (want_ctas == 2) nerf_render_kernel<2><<<(unsigned)ctas, kThreads, kSmemBytes, s>>>(P);
  else nerf_render_kernel<3><<<(unsigned)ctas, kThreads, kSmemBytes, s>>>(P);
  nerf_resolve_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(P);
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}
