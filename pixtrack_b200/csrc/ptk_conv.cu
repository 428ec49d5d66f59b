// 3x3 (and 1x1) convolution on channels-last fp16 activations as an implicit GEMM on the
// Blackwell tensor cores: tcgen05.mma (kind::f16, fp32 accumulate in TMEM) fed by TMA.
//
// Replaces the cuDNN convolutions under UNet._forward (reference
// pixloc/pixloc/pixlib/models/unet.py:68-99 encoder blocks, :15-44 decoder blocks).
//
// Five kernels, chosen per layer by conv_dispatch (bottom of the file; every choice has an environment switch):
//   conv_halo_kernel   persistent, 16 x 16 tiles, one halo load per 64-channel chunk, taps = shifted descriptors (the
//                      full-resolution 64 -> 64 layer; see its header comment)
//   conv_halo2_kernel  the same on CTA pairs (tcgen05 cta_group::2, M = 256 x N = 32 / 64 / 128 with two accumulator sets
//                      in TMEM, weights streamed or resident): most layers on the large maps (PTK_CONV_PAIR)
//   conv_row2_kernel   CTA pairs on ROW tiles (an MMA's 128 rows = 128 consecutive pixels of an image row), N = 64 / 128,
//                      R = 1..3 rows per CTA: maps about 128 pixels wide, where 16 x 16 tiles quantise badly (PTK_CONV_ROW)
//   conv_row64_kernel  CTA pairs on maps <= 64 pixels wide (three column-shifted halo copies) (PTK_CONV_ROW64)
//   conv_tc_kernel     one tap-shifted TMA box per k-step; small maps and 1x1 (described next); SPLIT = 2 shares the
//                      K loop of a tile between the two CTAs of a cluster (partial sums through DSMEM)
// All epilogues add the bias (conv bias or folded BatchNorm) from shared memory, apply ReLU, write fp16 as 32-byte
// pieces per lane (lane pairs swap halves: pair_store) and can also write the 2x2 max pool of the result.  All kernels
// are launched with programmatic dependent launch and prefetch their first weight tiles before griddepcontrol.wait.
//
// conv_tc_kernel, GEMM view:  D[128 pixels, BLOCK_N channels] += A[128, 64] * B[BLOCK_N, 64]^T over
//             K = taps x (C_in / 64) steps.
//  * A: one TMA 3-D box {64 ch, 16 px wide, 8 px high} of the NHWC input at the tap-shifted
//    position lands in shared memory as 128 rows x 128 B, 128B-swizzled -- exactly the K-major
//    UMMA operand layout.  Out-of-bounds box elements are zero-filled by TMA, which IS the
//    convolution's zero padding (and, with the tensor-map extents set to the cropped size, the
//    skip-connection crop of DecoderBlock.forward, unet.py:39-43).
//  * An optional second input tensor continues the K loop: torch.cat([upsampled, skip], 1)
//    (unet.py:44) is never materialised.
//  * B: weights pre-packed as [tap][C_out][C_in] fp16, box {64, BLOCK_N, 1}.
//  * Roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer
//    (warp-uniform loop, one elected lane issues), warps 2-5 = epilogue: tcgen05.ld 32 lanes x 32
//    columns, + bias, ReLU, fp16, 64 B per pixel per store.
//  * STAGES-deep mbarrier ring between TMA and MMA; tcgen05.commit releases a stage when the
//    MMAs that read it retire.  Several CTAs are resident per SM (smem permitting) so one CTA's
//    epilogue overlaps another's main loop.
#include <cuda.h>
#include <cuda_fp16.h>

#include "ptk_common.cuh"

namespace {

constexpr int kTileW = 16, kTileH = 8;  // 128 output pixels per CTA
constexpr int kKChunk = 64;             // fp16 channels per k-step = 128 B rows
constexpr int kABytes = 128 * 128;      // one A stage
constexpr int kConvThreads = 192;

struct ConvParams {
  int H, W;            // output (= input) spatial size
  int Cout;
  int cin0, cin1;      // channels taken from input 0 / input 1 (0 = absent); multiples of 64
  int taps;            // 9 (3x3, pad 1) or 1 (1x1)
  int relu;
  int tiles_w;
  int tw_log2;         // the CTA's 128 output pixels form a (1 << tw_log2) wide, (128 >> tw_log2) high tile
  const float* bias;   // [Cout] fp32
  __half* out;         // [H][W][Cout]
  __half* pool;        // optional [H/2][W/2][Cout]: 2x2 max pool of `out` (16 x 8 tiles only)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  unsigned spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 22)) __trap();   // never hang the GPU on a protocol bug
  }
}
// mbar_wait that adds the cycles spent waiting to `acc` (stall attribution, PTK_CONV_DBG)
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& acc, bool timed) {
  if (!timed) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// One lane of a CONVERGED warp (always the same one).  The MMA / TMA issue loops run warp-uniformly and
// predicate only the issuing instruction with this: descriptors then live in uniform registers, whereas
// an `if (lane == 0)` region makes the compiler wrap every UTCHMMA in an R2UR + elect waterfall loop
// (measured: ~100 cycles of issue per MMA, which bounds every layer with N <= 128).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row atoms of 1024 B):
// start>>4 | LBO(=1, ignored for swizzled K-major)<<16 | SBO(1024 B >> 4)<<32 | version 1<<46 | SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}


// 2x2 max pool inside a warp: the four pixels of a pooling window sit in lanes l, l^1 (x neighbour) and l^kYBit
// (y neighbour).  max commutes with the fp16 rounding, so pooling the packed outputs equals pooling `out`.
template <int kYBit>
__device__ __forceinline__ void pool_quad(uint32_t (&pw)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    __half2 v = *reinterpret_cast<__half2*>(&pw[j]);
    uint32_t o = __shfl_xor_sync(0xffffffffu, pw[j], 1);
    v = __hmax2(v, *reinterpret_cast<__half2*>(&o));
    uint32_t cur = *reinterpret_cast<uint32_t*>(&v);
    o = __shfl_xor_sync(0xffffffffu, cur, kYBit);
    v = __hmax2(v, *reinterpret_cast<__half2*>(&o));
    pw[j] = *reinterpret_cast<uint32_t*>(&v);
  }
}

// Biases of a tile's N output channels, one private copy per epilogue warp in shared memory.  (Read straight from global
// memory in the epilogue loop they miss L1 every time -- the output stores stream through it -- and each 32-channel step
// then waits an L2 round trip: measured ~700 of the 860 cycles a step took.)  Loaded before the warp waits for the
// accumulators, so the latency is hidden.
template <int N>
__device__ __forceinline__ void stage_bias(float* s_bias, const float* bias, int lane) {
  __syncwarp();
#pragma unroll
  for (int i = lane * 4; i < N; i += 128) *reinterpret_cast<float4*>(s_bias + i) = __ldg(reinterpret_cast<const float4*>(bias + i));
  __syncwarp();
}

// The extractor's level-0 head fused into the epilogue of the last decoder convolution (C_out = 32): the lane holds its
// pixel's 32 output channels (already rounded to fp16, which is what the separate head kernel would read back); the
// 33 x 32 weights are kernel parameters, i.e. constant-bank operands of the FMAs: 1 056 FMAs per pixel on epilogue warps
// that otherwise wait for the MMAs 80 % of the time.  Same arithmetic as head_mma_kernel up to the fp32 summation order.
// (unet.py:175-188: adaptation layer, uncertainty layer, sigmoid(-u), L2 normalisation.)
template <bool kHead>
struct HeadArg {};
template <>
struct HeadArg<true> {
  PtkHeadConst c;
};
__device__ __forceinline__ void fused_head32(const uint32_t (&pw)[16], const PtkHeadConst& hc, size_t pixel, bool inside) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pw[j]));
    x[2 * j] = f.x;
    x[2 * j + 1] = f.y;
  }
  // k outer, n inner: 33 independent accumulator chains; the empty asm keeps the compiler from hoisting the constant
  // loads of later k steps (which spilled ~1.8 KB per thread)
  float y[33];
#pragma unroll
  for (int n = 0; n < 33; ++n) y[n] = hc.b[n];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
#pragma unroll
    for (int n = 0; n < 33; ++n) y[n] = fmaf(x[k], hc.w[n * 32 + k], y[n]);
    asm volatile("" ::: "memory");
  }
  float ss = 0.f;
#pragma unroll
  for (int n = 0; n < 32; ++n) ss = fmaf(y[n], y[n], ss);
  const float inv = hc.normalize ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
  if (inside) {
    float* dst = hc.feat + pixel * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(y[8 * g + i] * inv);
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * g), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                   "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                   : "memory");
    }
    hc.conf[pixel] = 1.f / (1.f + expf(y[32]));
  }
}

// Output rows are written by lane PAIRS: a lane holds 32 output channels (64 B) of its own pixel, but a warp-wide 16-byte
// store of those scatters over 32 different lines -- 128 requests of 16 B per 32 channels, and the SM's store path, not
// the tensor pipe, then bounds every layer with a large output map (measured: 9.4 K cycles of epilogue per 256 x 128 tile
// against 4.7 K cycles of MMAs).  Lanes 2i and 2i + 1 swap halves instead: each store instruction writes 32 B per lane
// and the pair covers one pixel's 64 contiguous bytes -- two full sectors per request, 32 requests per 32 channels.
struct PairStore {
  __half* row[2];     // output rows of the even / odd pixel of this lane's pair (already offset to this lane's half)
  bool ok[2];
};
__device__ __forceinline__ PairStore pair_store_setup(__half* orow, bool inside, int lane) {
  PairStore s;
  const unsigned long long p = reinterpret_cast<unsigned long long>(orow);
  const int half_off = (lane & 1) * 16;
  s.row[0] = reinterpret_cast<__half*>(__shfl_sync(0xffffffffu, p, lane & ~1)) + half_off;
  s.row[1] = reinterpret_cast<__half*>(__shfl_sync(0xffffffffu, p, lane | 1)) + half_off;
  const unsigned m = __ballot_sync(0xffffffffu, inside);
  s.ok[0] = (m >> (lane & ~1)) & 1u;
  s.ok[1] = (m >> (lane | 1)) & 1u;
  return s;
}
__device__ __forceinline__ void st_global_256(__half* dst, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// pw: this lane's 32 packed fp16 outputs for channels [c, c + 32) of its own pixel.  Called by the whole warp.
__device__ __forceinline__ void pair_store(const PairStore& s, int c, const uint32_t (&pw)[16], int lane) {
  const bool odd = lane & 1;
  uint32_t a[8], b[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t send = odd ? pw[r] : pw[8 + r];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
    a[r] = odd ? recv : pw[r];         // even pixel: low half from the even lane, high half from the even lane via the odd one
    b[r] = odd ? pw[8 + r] : recv;     // odd pixel
  }
  if (s.ok[0]) st_global_256(s.row[0] + c, a);
  if (s.ok[1]) st_global_256(s.row[1] + c, b);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank0(uint32_t smem_addr) {   // the same offset in cluster rank 0's shared memory
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_addr));
  return remote;
}
// "This accumulator set has been read": arrives on rank 0's copy of `bar` WITHOUT release semantics.  The only thing the
// waiter (the MMA warp) does afterwards is overwrite TMEM, and the reads of it have completed (tcgen05.wait::ld) before
// the arrive is issued.  A releasing arrive at cluster scope compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. the epilogue warp
// waited for all of its OUTPUT STORES to become visible before it could free the accumulators: 21 % of all stall samples
// of the pair kernel on the 64 -> 64 full-resolution layer (ncu source page), once per tile and warp.
__device__ __forceinline__ void mbar_arrive_leader_relaxed(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on rank 0's copy of `bar`
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_rank0(smem_u32(bar))) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquires remote (DSMEM) writes
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  unsigned spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 22)) __trap();
  }
}

// SPLIT = 2: the K loop of one output tile is shared by a cluster of two CTAs (blockIdx.z = cluster rank).  Rank 1
// ships its fp32 partial accumulators into rank 0's shared memory (DSMEM stores, column-major so a warp writes 128
// contiguous bytes) and arrives on rank 0's `peer_bar`; rank 0 adds them in its epilogue.  Used for layers whose
// tiles would otherwise occupy at most half of the SMs (long K, few pixels, few output channels).
template <int BLOCK_N, int STAGES, int SPLIT>
__global__ void __launch_bounds__(kConvThreads) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                               const __grid_constant__ CUtensorMap tmA1,
                                                               const __grid_constant__ CUtensorMap tmW,
                                                               const ConvParams P) {
  constexpr int kBBytes = BLOCK_N * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;   // power of two >= 32
  // instruction descriptor: D=F32 (bit 4), A=B=F16, K-major both, N>>3 at bit 17, M>>4 at bit 24
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* peer_bar = tmem_full_bar + 1;                       // SPLIT: rank 1's partial sums have landed
  uint32_t* tmem_slot = (uint32_t*)(peer_bar + 1);
  float* xbuf = (float*)(smem + STAGES * kStageBytes + 256);    // SPLIT: [BLOCK_N / 4][128][4] fp32 partial sums of rank 1
  float* s_bias = (float*)(smem + STAGES * kStageBytes + 256 + (SPLIT > 1 ? BLOCK_N * 128 * 4 : 0));   // [BLOCK_N], shared by the 4 epilogue warps

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int th = tile / P.tiles_w, tw = tile - th * P.tiles_w;
  const int h0 = th * (128 >> P.tw_log2), w0 = tw << P.tw_log2;
  const int n0 = blockIdx.y * BLOCK_N;
  const int chunks0 = P.cin0 / kKChunk, chunks1 = P.cin1 / kKChunk;
  const int steps_per_tap = chunks0 + chunks1;
  const int krank = SPLIT > 1 ? (int)cluster_ctarank() : 0;
  const int ks_begin = P.taps * steps_per_tap * krank / SPLIT;
  const int num_steps = P.taps * steps_per_tap * (krank + 1) / SPLIT - ks_begin;   // this CTA's share of the K loop

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (chunks1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(peer_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is a warp-wide instruction; this warp also frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (SPLIT > 1) cluster_sync_all();   // rank 0's peer_bar must exist before rank 1 can arrive on it
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      // The weight tiles of the first ring of stages do not depend on the previous kernel: they are requested BEFORE the
      // programmatic-dependent-launch wait, the activation tiles after it.
      const int pre = num_steps < STAGES ? num_steps : STAGES;
      for (int ks = 0; ks < num_steps + pre; ++ks) {
        // passes 0 .. pre-1: weights of step ks; pass pre .. 2 pre - 1: activations of step ks - pre; then both per step
        const bool w_only = ks < pre, a_only = ks >= pre && ks < 2 * pre;
        const int step = w_only ? ks : ks - pre;
        if (ks == pre) {
          ptk_pdl_wait();
          ptk_pdl_trigger();
        }
        const int stage = step % STAGES;
        const uint32_t phase = (uint32_t)(step / STAGES) & 1u;
        const int kg = ks_begin + step;
        const int tap = kg / steps_per_tap, cc = kg - tap * steps_per_tap;
        const int dy = (P.taps == 9) ? tap / 3 - 1 : 0, dx = (P.taps == 9) ? tap % 3 - 1 : 0;
        uint8_t* sa = smem + stage * kStageBytes;
        if (!a_only) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_3d(sa + kABytes, &tmW, &full_bar[stage], cc * kKChunk, n0, tap);
        }
        if (!w_only) {
          if (cc < chunks0) tma_load_3d(sa, &tmA0, &full_bar[stage], cc * kKChunk, w0 + dx, h0 + dy);
          else tma_load_3d(sa, &tmA1, &full_bar[stage], (cc - chunks0) * kKChunk, w0 + dx, h0 + dy);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues ----------------
    for (int ks = 0; ks < num_steps; ++ks) {
      const int stage = ks % STAGES;
      const uint32_t phase = (uint32_t)(ks / STAGES) & 1u;
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * kStageBytes);
      const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + kABytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kKChunk / 16; ++k) {
          // +32 B per K=16 step inside the 128 B swizzle row = +2 in the (addr >> 4) field
          tc_mma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, kIdesc, (ks | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);   // stage is free once these MMAs have read it
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(tmem_full_bar);   // accumulator complete
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> registers -> bias/ReLU -> fp16 -> global ----------------
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;                  // row of the tile = pixel
    const int h = h0 + (m >> P.tw_log2), w = w0 + (m & ((1 << P.tw_log2) - 1));
    const bool inside = (h < P.H) && (w < P.W);
    // biases of this tile's channels -> shared memory (see stage_bias), before waiting for the accumulators
    for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) s_bias[i] = __ldg(P.bias + n0 + i);
    asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (SPLIT > 1 && krank != 0) {   // hand the partial sums to rank 0 and leave
      // exchange layout [column / 4][128 pixels][4 columns]: a lane ships 16 bytes per store, a warp 512 contiguous bytes
      const uint32_t remote = map_to_rank0(smem_u32(xbuf)) + (uint32_t)m * 16u;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote + (uint32_t)(c / 4 + j) * 2048u),
                       "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                       : "memory");
      }
      mbar_arrive_leader(peer_bar);
    } else {
    if (SPLIT > 1) mbar_wait_cluster(peer_bar, 0);
    __half* orow = P.out + ((size_t)h * P.W + w) * P.Cout + n0;
    const PairStore ps = pair_store_setup(orow, inside, lane);
    const bool pool_writer = P.pool != nullptr && ((lane & 17) == 0) && (h >> 1) < (P.H >> 1) && (w >> 1) < (P.W >> 1);
    __half* prow = P.pool != nullptr ? P.pool + ((size_t)(h >> 1) * (P.W >> 1) + (w >> 1)) * P.Cout + n0 : nullptr;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t v[32];
      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (SPLIT > 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x4 = *reinterpret_cast<const float4*>(xbuf + ((c / 4 + j) * 128 + m) * 4);
          v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + x4.x);
          v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + x4.y);
          v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + x4.z);
          v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + x4.w);
        }
      }
      uint32_t pw[16];
      const float4* b4 = reinterpret_cast<const float4*>(s_bias + c);   // 32 consecutive biases, broadcast 16-byte loads
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 bq = b4[j >> 1];
        float a = __uint_as_float(v[2 * j]) + ((j & 1) ? bq.z : bq.x);
        float b = __uint_as_float(v[2 * j + 1]) + ((j & 1) ? bq.w : bq.y);
        if (P.relu) {
          a = fmaxf(a, 0.f);
          b = fmaxf(b, 0.f);
        }
        const __half2 hv = __floats2half2_rn(a, b);
        pw[j] = *reinterpret_cast<const uint32_t*>(&hv);
      }
      pair_store(ps, c, pw, lane);
      if (P.pool != nullptr) {   // x neighbour = lane ^ 1, y neighbour = lane ^ 16
        pool_quad<16>(pw);
        if (pool_writer) {
          uint4* dst = reinterpret_cast<uint4*>(prow + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pw[4 * j], pw[4 * j + 1], pw[4 * j + 2], pw[4 * j + 3]);
        }
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// conv_halo_kernel: persistent 3x3 convolution with ONE halo load per 64-channel chunk.
//
// The tap-shifted operand of conv_tc_kernel re-fetches every input pixel nine times from L2
// (L2 -> SMEM bandwidth, not the tensor pipe, bounds it: 85 flop per staged byte at best).  Here a
// CTA owns a 16 x 16 pixel output tile and stages the (16+2) x (16+2) halo of a 64-channel chunk
// once: TMA box {64 ch, 18 px, 18 px} -> 324 rows of 128 B, 128B-swizzled, densely packed.  The
// operand of tap (dy, dx) for the 8-wide half tile `sx` is the SAME buffer seen through a descriptor
// whose start address is shifted by ((1+dy)*18 + 8*sx + 1+dx) rows and whose 8-row groups are 18 rows
// (2304 B) apart.  The tensor core applies the 128B swizzle to the absolute shared-memory address
// bits (measured: the descriptor's base-offset field must stay 0 for a start that is not
// 1024-B aligned), so the shifted view reads exactly what TMA wrote -- whatever the row pitch.  Two M=128 MMAs (the two half tiles) share each weight tile, the weights of a
// layer with <= 128 KB of them stay resident in shared memory, accumulators are double-buffered in
// TMEM so the epilogue of one tile overlaps the MMAs of the next, and CTAs are persistent.
// Staged bytes per flop drop 2.5-5x against conv_tc_kernel.
// ---------------------------------------------------------------------------------------------
// Halo rows are staged densely, 18 pixels = 2 304 B apart (round 1 padded them to 24 pixels = 3 x 1 024 B so that every
// 8-pixel run started a swizzle atom; the row-tile kernel showed that the tensor core, like TMA, applies the 128B swizzle
// to absolute address bits whatever the pitch, and the padding cost 25 % of the activation bytes of layers that sit at
// the per-SM TMA ingest rate).  A stage is 40.5 KB, rounded up to 41 KB so that both stages start 1 024-B aligned.
constexpr int kHaloW = 18, kHaloH = 18;
constexpr int kHaloBytes = kHaloW * kHaloH * 128;      // 41 472: bytes one halo load brings in
constexpr int kHaloStage = (kHaloBytes + 1023) / 1024 * 1024;   // 41 984
constexpr int kBBudget = 131072;                       // weight slots: 128 KB (8 x 16 KB at N = 128: the ring has to
                                                       // cover ~2 us of commit -> refill -> TMA round trip)
constexpr int kHaloThreads = 192;      // row-tile pair kernels: TMA warp, MMA warp, 4 epilogue warps
constexpr int kPair8Threads = 320;     // 16x16 pair kernel: TMA warp, MMA warp, 8 epilogue warps (4 per half tile)
constexpr int kHalo1Threads = 320;     // single-CTA kernel: TMA warp, MMA warp, 8 epilogue warps (4 per half tile)

struct HaloParams {
  int H, W, Cout;
  int chunks0, chunks1;
  int relu;
  int tiles_w, tiles_hw, total_tiles;
  int resident;          // all taps x chunks weight tiles fit the slots and Cout == N: load them once
  int rows_per_op;       // halo rows per TMA operation (the halo is fetched as 18 / rows_per_op boxes in flight)
  long long* dbg;        // optional (PTK_CONV_DBG=1): cycles CTA 0 spent waiting on each barrier kind
  const float* bias;
  __half* out;
  __half* pool;          // optional [H/2][W/2][Cout]: 2x2 max pool of `out`, written by the same epilogue
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {   // (relaxed: see mbar_arrive_leader_relaxed)
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// SPLIT = 2: the 64-channel chunks (K) of ONE output tile are shared by a cluster of two CTAs (small maps: too few
// 16 x 16 tiles to fill the machine, long K).  Each CTA runs the main loop over its half of the chunks; rank 1 then
// ships its fp32 partial accumulators into rank 0's shared memory -- into the operand staging area, which is idle
// once rank 0's MMAs have retired (rank 0 tells rank 1 so through `go_bar`) -- column-major, so a warp writes 128
// contiguous bytes per store, and arrives on rank 0's `peer_bar`; rank 0 adds them in its epilogue.  One tile per
// cluster (not persistent): grid = 2 x total_tiles.
template <int N, int SPLIT, bool kHead = false>
__global__ void __launch_bounds__(kHalo1Threads, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                    const __grid_constant__ CUtensorMap tmA1,
                                                                    const __grid_constant__ CUtensorMap tmW,
                                                                    const HaloParams P,
                                                                    const __grid_constant__ HeadArg<kHead> head) {
  static_assert(!kHead || (N == 32 && SPLIT == 1), "the fused head needs all 32 channels of a pixel in one lane");
  constexpr int kBBytes = N * 128;
  constexpr int kSlots = kBBudget / kBBytes;
  static_assert(SPLIT == 1 || N * 256 * 4 <= 2 * kHaloStage + kBBudget, "partial sums must fit the operand staging area");
  constexpr uint32_t kTmemCols = 4 * N < 32 ? 32 : 4 * N;      // 2 buffers x 2 half tiles
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                              // 2 halo stages
  uint8_t* sB = smem + 2 * kHaloStage;             // kSlots weight tiles
  uint64_t* fullA = (uint64_t*)(sB + kSlots * kBBytes);
  uint64_t* emptyA = fullA + 2;
  uint64_t* fullB = emptyA + 2;
  uint64_t* emptyB = fullB + kSlots;
  uint64_t* tmem_full = emptyB + kSlots;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* go_bar = tmem_empty + 2;       // SPLIT, on rank 1: rank 0's MMAs have retired, its staging area may be written
  uint64_t* peer_bar = go_bar + 1;         // SPLIT, on rank 0: rank 1's partial sums have landed
  uint32_t* tmem_slot = (uint32_t*)(peer_bar + 1);
  float* s_bias_all = reinterpret_cast<float*>(tmem_slot + 4);   // [8 epilogue warps][N]
  float* xbuf = reinterpret_cast<float*>(smem);   // SPLIT: [N][256] fp32 partial sums of rank 1 (aliases sA / sB)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = P.chunks0 + P.chunks1;
  const int krank = SPLIT > 1 ? (int)cluster_ctarank() : 0;
  const int c_begin = SPLIT > 1 ? (krank == 0 ? 0 : (chunks + 1) / 2) : 0;
  const int c_end = SPLIT > 1 ? (krank == 0 ? (chunks + 1) / 2 : chunks) : chunks;
  // work distribution: persistent CTAs stride over the tiles; a SPLIT cluster owns exactly one tile
  const int tile_first = SPLIT > 1 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = SPLIT > 1 ? P.total_tiles : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (P.chunks1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(&fullA[s], 1);
      mbar_init(&emptyA[s], 1);
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);
    }
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], 1);
    }
    mbar_init(go_bar, 1);
    mbar_init(peer_bar, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (SPLIT > 1) cluster_sync_all();   // the peer's barriers must exist before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      uint32_t a_it = 0, b_it = 0;
      bool first = true, waited = false;
      const bool timed = P.dbg != nullptr && blockIdx.x == 0;
      long long wA = 0, wB = 0;
      const long long tstart = clock64();
      for (int tile = tile_first; tile < P.total_tiles; tile += tile_step) {
        const int nb = tile / P.tiles_hw, sp = tile - nb * P.tiles_hw;
        const int th = sp / P.tiles_w, tw = sp - th * P.tiles_w;
        const int h0 = th * 16, w0 = tw * 16, n0 = nb * N;
        for (int c = c_begin; c < c_end; ++c) {
          auto load_weights = [&](int tap_begin, int tap_end) {
            if (P.resident && !first) return;
            for (int tap = tap_begin; tap < tap_end; ++tap) {
              int sb;
              if (P.resident) {
                sb = (c - c_begin) * 9 + tap;
              } else {
                sb = b_it % kSlots;
                mbar_wait_t(&emptyB[sb], ((b_it / kSlots) & 1u) ^ 1u, wB, timed);
              }
              mbar_expect_tx(&fullB[sb], kBBytes);
              tma_load_3d(sB + sb * kBBytes, &tmW, &fullB[sb], c * kKChunk, n0, tap);
              ++b_it;
            }
          };
          // The first weight tiles do not depend on the previous kernel: they are requested BEFORE the
          // programmatic-dependent-launch wait (as many as have a free slot without any MMA having run).
          int taps_done = 0;
          if (!waited) {
            taps_done = kSlots < 9 ? kSlots : 9;
            load_weights(0, taps_done);
            ptk_pdl_wait();
            ptk_pdl_trigger();
            waited = true;
          }
          const int sa = a_it & 1;
          mbar_wait_t(&emptyA[sa], ((a_it >> 1) & 1u) ^ 1u, wA, timed);
          mbar_expect_tx(&fullA[sa], kHaloBytes);
          for (int r = 0; r < kHaloH; r += P.rows_per_op) {
            uint8_t* dst = sA + sa * kHaloStage + r * (kHaloW * 128);
            if (c < P.chunks0) tma_load_3d(dst, &tmA0, &fullA[sa], c * kKChunk, w0 - 1, h0 - 1 + r);
            else tma_load_3d(dst, &tmA1, &fullA[sa], (c - P.chunks0) * kKChunk, w0 - 1, h0 - 1 + r);
          }
          ++a_it;
          load_weights(taps_done, 9);
        }
        first = false;
      }
      if (!waited) {   // no work for this CTA: still part of the chain
        ptk_pdl_wait();
        ptk_pdl_trigger();
      }
      if (timed) {
        P.dbg[0] = clock64() - tstart;
        P.dbg[1] = wA;
        P.dbg[2] = wB;
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues ----------------
    uint32_t a_it = 0, b_it = 0, t_it = 0;
    const bool timed = P.dbg != nullptr && blockIdx.x == 0;
    long long wT = 0, wA = 0, wB = 0;
    const long long tstart = clock64();
    for (int tile = tile_first; tile < P.total_tiles; tile += tile_step, ++t_it) {
      const uint32_t buf = t_it & 1u;
      mbar_wait_t(&tmem_empty[buf], ((t_it >> 1) & 1u) ^ 1u, wT, timed);
      tc_fence_after();
      for (int c = c_begin; c < c_end; ++c) {
        const int sa = a_it & 1;
        mbar_wait_t(&fullA[sa], (a_it >> 1) & 1u, wA, timed);
        ++a_it;
        const uint32_t abase = smem_u32(sA + sa * kHaloStage);
        // descriptor of tap (0,0) of half tile 0; the other taps / half tile are constant offsets of it
        const uint64_t adesc0 = (uint64_t)((abase & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) |
                                ((uint64_t)((kHaloW * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        // Taps are issued in groups of three: one elected region per group keeps the tensor pipe's queue fed
        // across tap boundaries (every elect / fence / barrier poll between MMAs is a bubble when the MMAs
        // are short: 41 cycles at N = 32).
#pragma unroll
        for (int tg = 0; tg < 3; ++tg) {
          int sb[3];
#pragma unroll
          for (int t3 = 0; t3 < 3; ++t3) {
            const int tap = tg * 3 + t3;
            if (P.resident) {
              sb[t3] = (c - c_begin) * 9 + tap;
              if (t_it == 0) mbar_wait_t(&fullB[sb[t3]], 0, wB, timed);
            } else {
              sb[t3] = b_it % kSlots;
              mbar_wait_t(&fullB[sb[t3]], (b_it / kSlots) & 1u, wB, timed);
              ++b_it;
            }
          }
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int t3 = 0; t3 < 3; ++t3) {
              const int tap = tg * 3 + t3;
              const uint64_t bdesc = make_sw128_desc(smem_u32(sB + sb[t3] * kBBytes));
              // k outer, half tile inner: consecutive MMAs write DIFFERENT accumulators, so an MMA never has to wait for
              // the accumulate of the one issued just before it
#pragma unroll
              for (int k = 0; k < kKChunk / 16; ++k) {
#pragma unroll
                for (int sx = 0; sx < 2; ++sx) {
                  // start row of this tap's view: ((1+dy)*24 + 8*sx + 1+dx), 128 B per row, >> 4 in the descriptor
                  const int row0 = (tap / 3) * kHaloW + 8 * sx + (tap % 3);
                  const uint64_t adesc = adesc0 + (uint64_t)(row0 * 8);
                  const uint32_t d = tmem_base + (buf * 2u + (uint32_t)sx) * (uint32_t)N;
                  tc_mma_f16(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, ((c - c_begin) | tap | k) != 0 ? 1u : 0u);
                }
              }
              if (!P.resident) tc_commit(&emptyB[sb[t3]]);
            }
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(&emptyA[sa]);
        __syncwarp();
      }
      if (elect_one()) tc_commit(&tmem_full[buf]);
      __syncwarp();
    }
    if (timed && lane == 0) {
      P.dbg[3] = clock64() - tstart;
      P.dbg[4] = wT;
      P.dbg[5] = wA;
      P.dbg[6] = wB;
      P.dbg[7] = t_it;
    }
  } else {
    // ---------------- epilogue: 8 warps; warp w reads TMEM lane quadrant q = w % 4 (rows 32q .. 32q+31) of half
    //                  tile sx = (w - 2) / 4, so the two half tiles drain in parallel ----------------
    const int q = warp & 3;
    const int sx = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int y = m >> 3, xx = m & 7;
    uint32_t t_it = 0;
    const bool timed = P.dbg != nullptr && blockIdx.x == 0 && warp == 2;   // warp 2: quadrant 2 of half tile 0
    long long wF = 0;
    const long long tstart = clock64();
    for (int tile = tile_first; tile < P.total_tiles; tile += tile_step, ++t_it) {
      const int nb = tile / P.tiles_hw, sp = tile - nb * P.tiles_hw;
      const int th = sp / P.tiles_w, tw = sp - th * P.tiles_w;
      const int h = th * 16 + y, n0 = nb * N;
      const uint32_t buf = t_it & 1u;
      float* s_bias = s_bias_all + (warp - 2) * N;
      if (N >= 128) stage_bias<N>(s_bias, P.bias + n0, lane);
      else stage_bias<128>(s_bias, P.bias + n0, lane < N / 4 ? lane : 0);
      mbar_wait_t(&tmem_full[buf], (t_it >> 1) & 1u, wF, timed);
      tc_fence_after();
      if (SPLIT > 1) {
        const int mg = sx * 128 + m;                 // row of the 256-pixel tile
        if (krank != 0) {
          // rank 0's staging area is free once its MMAs have retired: wait for its word, then ship the partial sums
          mbar_wait_cluster(go_bar, 0);
          // exchange layout [column / 4][256 pixels][4 columns]: 16 bytes per lane and store, 512 contiguous bytes per warp
          const uint32_t remote = map_to_rank0(smem_u32(xbuf)) + (uint32_t)mg * 16u;
#pragma unroll 1
          for (int c = 0; c < N; c += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)sx * (uint32_t)N + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote + (uint32_t)(c / 4 + j) * 4096u),
                           "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                           : "memory");
          }
          mbar_arrive_leader(peer_bar);
          continue;
        }
        // rank 0: its own accumulators are complete, i.e. every operand it staged has been consumed
        if (warp == 2 && lane == 0) {
          uint32_t remote_go;
          asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(remote_go) : "r"(smem_u32(go_bar)));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_go) : "memory");
        }
        mbar_wait_cluster(peer_bar, 0);
      }
      {
        const int w = tw * 16 + 8 * sx + xx;
        const bool inside = (h < P.H) && (w < P.W);
        __half* orow = P.out + ((size_t)h * P.W + w) * P.Cout + n0;
        const PairStore ps = pair_store_setup(orow, inside, lane);
        const bool pool_writer = P.pool != nullptr && ((lane & 9) == 0) && (h >> 1) < (P.H >> 1) && (w >> 1) < (P.W >> 1);
        __half* prow = P.pool != nullptr ? P.pool + ((size_t)(h >> 1) * (P.W >> 1) + (w >> 1)) * P.Cout + n0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < N; c += 32) {
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2u + (uint32_t)sx) * (uint32_t)N + (uint32_t)c, v);
          if (SPLIT > 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x4 = *reinterpret_cast<const float4*>(xbuf + ((c / 4 + j) * 256 + sx * 128 + m) * 4);
              v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + x4.x);
              v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + x4.y);
              v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + x4.z);
              v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + x4.w);
            }
          }
          uint32_t pw[16];
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c);   // 32 consecutive biases, broadcast 16-byte loads
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bq = b4[j >> 1];
            float a = __uint_as_float(v[2 * j]) + ((j & 1) ? bq.z : bq.x);
            float b = __uint_as_float(v[2 * j + 1]) + ((j & 1) ? bq.w : bq.y);
            if (P.relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            const __half2 hv = __floats2half2_rn(a, b);
            pw[j] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          pair_store(ps, c, pw, lane);
          if constexpr (kHead) fused_head32(pw, head.c, (size_t)h * P.W + w, inside);
          if (P.pool != nullptr) {   // x neighbour = lane ^ 1, y neighbour = lane ^ 8
            pool_quad<8>(pw);
            if (pool_writer) {
              uint4* dst = reinterpret_cast<uint4*>(prow + c);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pw[4 * j], pw[4 * j + 1], pw[4 * j + 2], pw[4 * j + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
    if (timed && lane == 0) {
      P.dbg[8] = clock64() - tstart;
      P.dbg[9] = wF;
    }
  }

  tc_fence_before();
  if (SPLIT > 1) cluster_sync_all();   // rank 0's shared memory stays mapped until rank 1 has finished writing to it
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// conv_halo2_kernel: the halo kernel on CTA pairs (tcgen05 cta_group::2) for C_out = k * 256.
//
// A single-CTA MMA in SS mode is bounded by its operand fetch from shared memory (measured: cycles per
// M=128 MMA ~ (A bytes + B bytes) / 64, 147 cycles at N=128 against a floor of 64).  A CTA pair issues
// M=256 x N=256 MMAs: each SM still supplies its own 128 A rows but only HALF of the B rows, so the
// operand bytes per SM and per flop halve and the MMA runs at its floor (128 cycles per K=16 step).
// The pair owns two horizontally adjacent 16x16 tiles (CTA r: tile 2*pair + r): every CTA stages the halo
// of its own tile and rows [128 r, 128 r + 128) of each 256-row weight tile; TMA completions of both
// CTAs land on the leader's (rank 0) mbarriers, the leader's MMA warp issues for the pair and releases
// stages / publishes accumulators with multicast commits, and both CTAs run the epilogue on their own
// TMEM lanes (2 half tiles x 256 columns = all 512 columns, so the epilogue is not overlapped).
// ---------------------------------------------------------------------------------------------
constexpr int kPairN = 256;                          // widest pair tile (C_out = k * 256); N = 128 is the double-buffered form
constexpr int kPairBudget = 16 * 64 * 128;           // weight ring per CTA: 128 KB (16 half tiles of 8 KB at N = 128)

// TMA load whose completion bytes are credited to the mbarrier at the same offset in the pair's leader CTA
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                                 int c2) {
  const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;   // clear the peer bit: rank 0's copy of the barrier
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// N = 256: one accumulator set (2 half tiles x 256 columns = all of TMEM), the epilogue is not overlapped.
// N = 128: two accumulator sets, so the epilogue of a tile runs under the MMAs of the next one -- the pair then keeps what
//          it gains on weight traffic (each CTA stages only half of every weight tile).
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPair8Threads, 1)
    conv_halo2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmW, const HaloParams P) {
  constexpr int kPairBBytes = (N / 2) * 128;           // this CTA's half of a weight tile
  constexpr int kPairSlots = kPairBudget / kPairBBytes;
  constexpr int kBufs = (4 * N <= 512) ? 2 : 1;        // accumulator sets in TMEM
  constexpr uint32_t kTmemCols = 512;
  // D=F32, A=B=F16 K-major, N >> 3 at bit 17, M (= 256 for the pair) >> 4 at bit 24
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * kHaloStage;
  uint64_t* fullA = (uint64_t*)(sB + kPairSlots * kPairBBytes);
  uint64_t* emptyA = fullA + 2;
  uint64_t* fullB = emptyA + 2;
  uint64_t* emptyB = fullB + kPairSlots;
  uint64_t* tmem_full = emptyB + kPairSlots;
  uint64_t* tmem_empty = tmem_full + kBufs;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + kBufs);
  float* s_bias_all = reinterpret_cast<float*>(tmem_slot + 4);   // [8 epilogue warps][N]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int chunks = P.chunks0 + P.chunks1;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (P.chunks1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(&fullA[s], 1);
      mbar_init(&emptyA[s], 1);
    }
    for (int s = 0; s < kPairSlots; ++s) {
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < kBufs; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 16);  // 8 epilogue warps of each CTA arrive on the leader's copy
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // Both CTAs must be running before the pair allocation: tcgen05.alloc.cta_group::2 expands to a handshake through the
  // PEER's shared memory (the leader stores the column address and arrives on a barrier there).  Two CTAs of a cluster are
  // co-scheduled but do not start at the same instant -- on a busy GPU the gap is wide enough for the leader's writes to
  // land before the peer's shared memory is set up, and the peer then waits for ever (seen as an intermittent hang with
  // two extractor plans replayed concurrently).  The same barrier publishes the mbarrier initialisation cluster-wide.
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                   // every warp of this CTA sees the column address its own warp 1 received
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs: own halo, own half of the weight rows) ----------------
      uint32_t a_it = 0, b_it = 0;
      bool waited = false;
      const bool timed = P.dbg != nullptr && blockIdx.x == 0;
      long long wA = 0, wB = 0;
      const long long tstart = clock64();
      for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs) {
        const int nb = tile / P.tiles_hw, sp = tile - nb * P.tiles_hw;
        const int th = sp / P.tiles_w, tw = sp - th * P.tiles_w;     // tiles_w counts PAIRS of 16-pixel columns
        const int h0 = th * 16, w0 = (tw * 2 + (int)rank) * 16, n0 = nb * N + (int)rank * (N / 2);
        for (int c = 0; c < chunks; ++c) {
          auto load_weights = [&](int tap_begin, int tap_end) {
            if (P.resident && tile != pair_id) return;      // resident weights: loaded with the first tile only
            for (int tap = tap_begin; tap < tap_end; ++tap) {
              int sb;
              if (P.resident) {
                sb = c * 9 + tap;
              } else {
                sb = b_it % kPairSlots;
                mbar_wait_t(&emptyB[sb], ((b_it / kPairSlots) & 1u) ^ 1u, wB, timed);
              }
              if (leader) mbar_expect_tx(&fullB[sb], 2 * kPairBBytes);
              tma_load_3d_pair(sB + sb * kPairBBytes, &tmW, &fullB[sb], c * kKChunk, n0, tap);
              ++b_it;
            }
          };
          int taps_done = 0;
          if (!waited) {   // weights first, then the programmatic-dependent-launch wait, then activations
            taps_done = kPairSlots < 9 ? kPairSlots : 9;
            load_weights(0, taps_done);
            ptk_pdl_wait();
            ptk_pdl_trigger();
            waited = true;
          }
          const int sa = a_it & 1;
          mbar_wait_t(&emptyA[sa], ((a_it >> 1) & 1u) ^ 1u, wA, timed);
          if (leader) mbar_expect_tx(&fullA[sa], 2 * kHaloBytes);    // both CTAs' halos are credited here
          for (int r = 0; r < kHaloH; r += P.rows_per_op) {
            uint8_t* dst = sA + sa * kHaloStage + r * (kHaloW * 128);
            if (c < P.chunks0) tma_load_3d_pair(dst, &tmA0, &fullA[sa], c * kKChunk, w0 - 1, h0 - 1 + r);
            else tma_load_3d_pair(dst, &tmA1, &fullA[sa], (c - P.chunks0) * kKChunk, w0 - 1, h0 - 1 + r);
          }
          ++a_it;
          load_weights(taps_done, 9);
        }
      }
      if (!waited) {
        ptk_pdl_wait();
        ptk_pdl_trigger();
      }
      if (timed) {
        P.dbg[0] = clock64() - tstart;
        P.dbg[1] = wA;
        P.dbg[2] = wB;
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer of the pair (warp-uniform loop, one elected lane issues) ----------------
      uint32_t a_it = 0, b_it = 0, t_it = 0;
      const bool timed = P.dbg != nullptr && blockIdx.x == 0;
      long long wT = 0, wA = 0, wB = 0;
      const long long tstart = clock64();
      for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs, ++t_it) {
        const uint32_t buf = t_it % kBufs;
        mbar_wait_t(&tmem_empty[buf], ((t_it / kBufs) & 1u) ^ 1u, wT, timed);
        tc_fence_after();
        for (int c = 0; c < chunks; ++c) {
          const int sa = a_it & 1;
          mbar_wait_t(&fullA[sa], (a_it >> 1) & 1u, wA, timed);
          ++a_it;
          const uint32_t abase = smem_u32(sA + sa * kHaloStage);
          const uint64_t adesc0 = (uint64_t)((abase & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) |
                                  ((uint64_t)((kHaloW * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            int sb;
            if (P.resident) {      // every weight tile of the layer has its own slot, filled once
              sb = c * 9 + tap;
              if (t_it == 0) mbar_wait_t(&fullB[sb], 0, wB, timed);
            } else {
              sb = b_it % kPairSlots;
              mbar_wait_t(&fullB[sb], (b_it / kPairSlots) & 1u, wB, timed);
              ++b_it;
            }
            tc_fence_after();
            const uint64_t bdesc = make_sw128_desc(smem_u32(sB + sb * kPairBBytes));
            if (elect_one()) {
#pragma unroll
              for (int sx = 0; sx < 2; ++sx) {
                const int row0 = (tap / 3) * kHaloW + 8 * sx + (tap % 3);
                const uint64_t adesc = adesc0 + (uint64_t)(row0 * 8);
                const uint32_t d = tmem_base + (buf * 2u + (uint32_t)sx) * (uint32_t)N;
#pragma unroll
                for (int k = 0; k < kKChunk / 16; ++k)
                  tc_mma_f16_pair(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, (c | tap | k) != 0 ? 1u : 0u);
              }
              if (!P.resident) tc_commit_pair(&emptyB[sb]);
            }
            __syncwarp();
          }
          if (elect_one()) tc_commit_pair(&emptyA[sa]);
          __syncwarp();
        }
        if (elect_one()) tc_commit_pair(&tmem_full[buf]);
        __syncwarp();
      }
      if (timed && lane == 0) {
        P.dbg[3] = clock64() - tstart;
        P.dbg[4] = wT;
        P.dbg[5] = wA;
        P.dbg[6] = wB;
        P.dbg[7] = t_it;
      }
    }
  } else {
    // ---------------- epilogue (both CTAs, own TMEM lanes): 8 warps; warp w reads TMEM lane quadrant q = w % 4 of half
    //                  tile sx = (w - 2) / 4, so the two half tiles drain in parallel (with four warps a tile's epilogue
    //                  outlasted its MMAs on the one-chunk layers) ----------------
    const int q = warp & 3;
    const int sx = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int y = m >> 3, xx = m & 7;
    uint32_t t_it = 0;
    const bool timed = P.dbg != nullptr && blockIdx.x == 0 && warp == 2;
    long long wF = 0, t_ld = 0, t_st = 0;
    const long long tstart = clock64();
    for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs, ++t_it) {
      const int nb = tile / P.tiles_hw, sp = tile - nb * P.tiles_hw;
      const int th = sp / P.tiles_w, tw = sp - th * P.tiles_w;
      const int h = th * 16 + y, n0 = nb * N;
      const uint32_t buf = t_it % kBufs;
      float* s_bias = s_bias_all + (warp - 2) * N;
      if (N >= 128) stage_bias<N>(s_bias, P.bias + n0, lane);
      else stage_bias<128>(s_bias, P.bias + n0, lane < N / 4 ? lane : 0);
      mbar_wait_t(&tmem_full[buf], (t_it / kBufs) & 1u, wF, timed);
      tc_fence_after();
      {
        const int w = (tw * 2 + (int)rank) * 16 + 8 * sx + xx;
        const bool inside = (h < P.H) && (w < P.W);
        __half* orow = P.out + ((size_t)h * P.W + w) * P.Cout + n0;
        const PairStore ps = pair_store_setup(orow, inside, lane);
        const bool pool_writer = P.pool != nullptr && ((lane & 9) == 0) && (h >> 1) < (P.H >> 1) && (w >> 1) < (P.W >> 1);
        __half* prow = P.pool != nullptr ? P.pool + ((size_t)(h >> 1) * (P.W >> 1) + (w >> 1)) * P.Cout + n0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < N; c += 32) {
          uint32_t v[32];
          const long long tl0 = timed ? clock64() : 0;
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2u + (uint32_t)sx) * (uint32_t)N + (uint32_t)c, v);
          if (timed) t_ld += clock64() - tl0;
          uint32_t pw[16];
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bq = b4[j >> 1];
            float a = __uint_as_float(v[2 * j]) + ((j & 1) ? bq.z : bq.x);
            float b = __uint_as_float(v[2 * j + 1]) + ((j & 1) ? bq.w : bq.y);
            if (P.relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            const __half2 hv = __floats2half2_rn(a, b);
            pw[j] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          const long long ts0 = timed ? clock64() : 0;
          pair_store(ps, c, pw, lane);
          if (timed) t_st += clock64() - ts0;
          if (P.pool != nullptr) {
            pool_quad<8>(pw);
            if (pool_writer) {
              uint4* dst = reinterpret_cast<uint4*>(prow + c);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pw[4 * j], pw[4 * j + 1], pw[4 * j + 2], pw[4 * j + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader_relaxed(&tmem_empty[buf]);
    }
    if (timed && lane == 0) {
      P.dbg[8] = clock64() - tstart;
      P.dbg[9] = wF;
      P.dbg[10] = t_ld;
      P.dbg[11] = t_st;
    }
  }

  tc_fence_before();
  cluster_sync_all();                // the peer may still read this CTA's shared memory / TMEM until here
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// conv_row2_kernel: CTA pairs on ROW tiles, for the 1/8-scale maps (72 x 128, 94 x 126) with 512 channels.
//
// 16 x 16 tiles quantise badly on a 72-row map (80 pair tiles on 74 SM pairs) and the per-tap kernel that served these
// layers re-reads every activation nine times (663 MB of L2 -> shared traffic per layer: it ran at 1 040 TF/s).  Here an
// MMA's 128 rows are 128 CONSECUTIVE PIXELS OF ONE IMAGE ROW (core matrices 1 024 B apart in a densely staged row), a
// CTA owns R rows x 128 columns and stages their (R + 2) x 130-pixel halo once per 64-channel chunk, and a pair
// (rank r: rows [2R th + R r, + R)) issues M = 256 x N = 128 MMAs with each CTA staging half of every weight tile:
// 72 x 128 with R = 2 is 72 pair tiles = ONE wave on 74 SM pairs, 94 x 126 with R = 3 is 64; staged bytes per layer
// drop 4x (159 MB).  The halo rows are denser in bytes than the 16 x 16 form (2.0x the tile at R = 2 against 1.27x), which
// is why the larger maps stay on conv_halo2_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kRowW = 130;                           // staged pixels per halo row: 128 + 2

template <int R, int N_>
struct RowGeom {
  static constexpr int kN = N_;                                        // 128, or 64 for the decoder's 64-channel layers
  static constexpr int kHaloRows = R + 2;
  static constexpr int kHaloBytesExact = kHaloRows * kRowW * 128;
  static constexpr int kStageA = (kHaloBytesExact + 1023) / 1024 * 1024;
  static constexpr int kBBytes = (kN / 2) * 128;                       // this CTA's half of a weight tile: 8 KB (4 KB at N = 64)
  // halo stages: a single-row tile at N = 64 consumes a chunk in 1 600 cycles, less than one halo load takes to arrive
  // (measured: 37.8 us with two stages on the first decoder convolution), so it gets a third stage
  static constexpr int kAStages = (R == 1) ? 3 : 2;
  static constexpr int kSlotsRaw = (232448 - 1024 - kAStages * kStageA - 768 - 4 * kN * 4) / kBBytes;
  static constexpr int kSlots = kSlotsRaw > 27 ? 27 : kSlotsRaw;
  static constexpr int kBufs = (2 * R * kN <= 512) ? 2 : 1;            // accumulator sets in TMEM
  static constexpr int kSmem =
      kAStages * kStageA + kSlots * kBBytes + (2 * kAStages + 2 * kSlots + 2 * kBufs) * 8 + 16 + 4 * kN * 4 + 1024;
  static_assert(R * kN <= 512, "accumulators of one tile must fit TMEM");
  static_assert(kSlots >= 6, "weight ring too short");
  static_assert(kSmem <= 232448, "shared memory budget");
};

template <int R, int NN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kHaloThreads, 1)
    conv_row2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmW, const HaloParams P) {
  using G = RowGeom<R, NN>;
  constexpr int N = G::kN;
  constexpr int kSlots = G::kSlots;
  constexpr int kBufs = G::kBufs;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + G::kAStages * G::kStageA;
  uint64_t* fullA = (uint64_t*)(sB + kSlots * G::kBBytes);
  uint64_t* emptyA = fullA + G::kAStages;
  uint64_t* fullB = emptyA + G::kAStages;
  uint64_t* emptyB = fullB + kSlots;
  uint64_t* tmem_full = emptyB + kSlots;
  uint64_t* tmem_empty = tmem_full + kBufs;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + kBufs);
  float* s_bias_all = reinterpret_cast<float*>(tmem_slot + 4);   // [4 epilogue warps][N]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int chunks = P.chunks0 + P.chunks1;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (P.chunks1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    for (int s = 0; s < G::kAStages; ++s) {
      mbar_init(&fullA[s], 1);
      mbar_init(&emptyA[s], 1);
    }
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < kBufs; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);  // 4 epilogue warps of each CTA arrive on the leader's copy
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  cluster_sync_all();                // the peer is running before the pair allocation touches its shared memory (see conv_halo2_kernel)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs: own halo, own half of the weight rows) ----------------
      uint32_t a_it = 0, b_it = 0;
      bool waited = false;
      const bool timed = P.dbg != nullptr && blockIdx.x == 0;
      long long wA = 0, wB = 0;
      const long long tstart = clock64();
      for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs) {
        const int nb = tile / P.tiles_hw, sp = tile - nb * P.tiles_hw;
        const int th = sp / P.tiles_w, tw = sp - th * P.tiles_w;
        const int h0 = (th * 2 + (int)rank) * R, w0 = tw * 128, n0 = nb * N + (int)rank * (N / 2);
        for (int c = 0; c < chunks; ++c) {
          auto load_weights = [&](int tap_begin, int tap_end) {
            for (int tap = tap_begin; tap < tap_end; ++tap) {
              const int sb = b_it % kSlots;
              mbar_wait_t(&emptyB[sb], ((b_it / kSlots) & 1u) ^ 1u, wB, timed);
              if (leader) mbar_expect_tx(&fullB[sb], 2 * G::kBBytes);
              tma_load_3d_pair(sB + sb * G::kBBytes, &tmW, &fullB[sb], c * kKChunk, n0, tap);
              ++b_it;
            }
          };
          int taps_done = 0;
          if (!waited) {   // weights first, then the programmatic-dependent-launch wait, then activations
            taps_done = kSlots < 9 ? kSlots : 9;
            load_weights(0, taps_done);
            ptk_pdl_wait();
            ptk_pdl_trigger();
            waited = true;
          }
          const int sa = a_it % G::kAStages;
          mbar_wait_t(&emptyA[sa], ((a_it / G::kAStages) & 1u) ^ 1u, wA, timed);
          if (leader) mbar_expect_tx(&fullA[sa], 2 * G::kHaloBytesExact);    // both CTAs' halos are credited here
          uint8_t* dst = sA + sa * G::kStageA;
          if (c < P.chunks0) tma_load_3d_pair(dst, &tmA0, &fullA[sa], c * kKChunk, w0 - 1, h0 - 1);
          else tma_load_3d_pair(dst, &tmA1, &fullA[sa], (c - P.chunks0) * kKChunk, w0 - 1, h0 - 1);
          ++a_it;
          load_weights(taps_done, 9);
        }
      }
      if (!waited) {
        ptk_pdl_wait();
        ptk_pdl_trigger();
      }
      if (timed) {
        P.dbg[0] = clock64() - tstart;
        P.dbg[1] = wA;
        P.dbg[2] = wB;
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer of the pair (warp-uniform loop, one elected lane issues) ----------------
      uint32_t a_it = 0, b_it = 0, t_it = 0;
      const bool timed = P.dbg != nullptr && blockIdx.x == 0;
      long long wT = 0, wA = 0, wB = 0;
      const long long tstart = clock64();
      for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs, ++t_it) {
        const uint32_t buf = t_it % kBufs;
        mbar_wait_t(&tmem_empty[buf], ((t_it / kBufs) & 1u) ^ 1u, wT, timed);
        tc_fence_after();
        for (int c = 0; c < chunks; ++c) {
          const int sa = a_it % G::kAStages;
          mbar_wait_t(&fullA[sa], (a_it / G::kAStages) & 1u, wA, timed);
          ++a_it;
          // 128 consecutive pixels of a staged row: core matrices (8 pixels x 128 B) 1 024 B apart
          const uint64_t adesc0 = make_sw128_desc(smem_u32(sA + sa * G::kStageA));
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int sb = b_it % kSlots;
            mbar_wait_t(&fullB[sb], (b_it / kSlots) & 1u, wB, timed);
            ++b_it;
            tc_fence_after();
            const uint64_t bdesc = make_sw128_desc(smem_u32(sB + sb * G::kBBytes));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kKChunk / 16; ++k) {
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {       // consecutive MMAs write different accumulators
                  const int px0 = (rr + tap / 3) * kRowW + (tap % 3);
                  const uint64_t adesc = adesc0 + (uint64_t)(px0 * 8);
                  const uint32_t d = tmem_base + (buf * (uint32_t)R + (uint32_t)rr) * (uint32_t)N;
                  tc_mma_f16_pair(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, (c | tap | k) != 0 ? 1u : 0u);
                }
              }
              tc_commit_pair(&emptyB[sb]);
            }
            __syncwarp();
          }
          if (elect_one()) tc_commit_pair(&emptyA[sa]);
          __syncwarp();
        }
        if (elect_one()) tc_commit_pair(&tmem_full[buf]);
        __syncwarp();
      }
      if (timed && lane == 0) {
        P.dbg[3] = clock64() - tstart;
        P.dbg[4] = wT;
        P.dbg[5] = wA;
        P.dbg[6] = wB;
        P.dbg[7] = t_it;
      }
    }
  } else {
    // ---------------- epilogue (both CTAs, own TMEM lanes): lane m of an accumulator = column w0 + m of one row -------
    const int q = warp & 3;
    const int m = q * 32 + lane;
    uint32_t t_it = 0;
    const bool timed = P.dbg != nullptr && blockIdx.x == 0 && warp == 2;
    long long wF = 0;
    const long long tstart = clock64();
    for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs, ++t_it) {
      const int nb = tile / P.tiles_hw, sp = tile - nb * P.tiles_hw;
      const int th = sp / P.tiles_w, tw = sp - th * P.tiles_w;
      const int h0 = (th * 2 + (int)rank) * R, w = tw * 128 + m, n0 = nb * N;
      const uint32_t buf = t_it % kBufs;
      float* s_bias = s_bias_all + (warp - 2) * N;
      if (N >= 128) stage_bias<N>(s_bias, P.bias + n0, lane);
      else stage_bias<128>(s_bias, P.bias + n0, lane < N / 4 ? lane : 0);
      mbar_wait_t(&tmem_full[buf], (t_it / kBufs) & 1u, wF, timed);
      tc_fence_after();
      PairStore ps[R];
#pragma unroll
      for (int rr = 0; rr < R; ++rr)
        ps[rr] = pair_store_setup(P.out + ((size_t)(h0 + rr) * P.W + w) * P.Cout + n0, (h0 + rr) < P.H && w < P.W, lane);
#pragma unroll 1
      for (int c = 0; c < N; c += 32) {
        uint32_t pooled[16];
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + c);
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const int h = h0 + rr;
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (buf * (uint32_t)R + (uint32_t)rr) * (uint32_t)N + (uint32_t)c, v);
          uint32_t pw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bq = b4[j >> 1];
            float a = __uint_as_float(v[2 * j]) + ((j & 1) ? bq.z : bq.x);
            float b = __uint_as_float(v[2 * j + 1]) + ((j & 1) ? bq.w : bq.y);
            if (P.relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            const __half2 hv = __floats2half2_rn(a, b);
            pw[j] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          pair_store(ps[rr], c, pw, lane);
          if ((R & 1) == 0 && P.pool != nullptr) {      // rows pair up inside a CTA only for even R (the host checks)
            if ((rr & 1) == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pooled[j] = pw[j];
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __half2 x = __hmax2(*reinterpret_cast<__half2*>(&pooled[j]), *reinterpret_cast<__half2*>(&pw[j]));
                uint32_t cur = *reinterpret_cast<uint32_t*>(&x);
                const uint32_t o = __shfl_xor_sync(0xffffffffu, cur, 1);
                x = __hmax2(x, *reinterpret_cast<const __half2*>(&o));
                pooled[j] = *reinterpret_cast<uint32_t*>(&x);
              }
              if ((lane & 1) == 0 && (h >> 1) < (P.H >> 1) && (w >> 1) < (P.W >> 1)) {
                uint4* dst = reinterpret_cast<uint4*>(P.pool + ((size_t)(h >> 1) * (P.W >> 1) + (w >> 1)) * P.Cout + n0 + c);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  dst[j] = make_uint4(pooled[4 * j], pooled[4 * j + 1], pooled[4 * j + 2], pooled[4 * j + 3]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader_relaxed(&tmem_empty[buf]);
    }
    if (timed && lane == 0) {
      P.dbg[8] = clock64() - tstart;
      P.dbg[9] = wF;
    }
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync_all();                // the peer may still read this CTA's shared memory / TMEM until here
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// conv_row64_kernel: CTA pairs on maps at most 64 pixels wide (the 1/16-scale block: 36 x 64, 47 x 63, 512 channels).
//
// These layers ran on the per-tap kernel with its K loop split over two CTAs: 144 / 192 CTAs that each stage 1.15 MB
// for 5 us of MMA work -- bound by L2 -> shared traffic (166 MB per layer) at ~500 TFLOP/s.  A 64-pixel row cannot use
// conv_row2_kernel's view (an MMA's 128 rows must be 128 consecutive staged pixels, and a 66-pixel halo row breaks the
// run at every image row).  Instead the halo of a 64-channel chunk is staged as THREE column-shifted copies, one per
// dx: box {64 ch, 64 px, 4 rows} at x = dx - 1, dense 64-pixel pitch.  In copy dx, rows (dy, dy + 1) are 128 consecutive
// pixels = the operand of tap (dy, dx) for the CTA's two output rows.  A pair (cta_group::2, M = 256 x N = 128, each CTA
// stages half of every weight tile) owns four rows; activations are staged 3x instead of 9x and weights once per pair:
// 97 MB per layer.  Taps run dx-major so that a copy is released after its three taps.
// ---------------------------------------------------------------------------------------------
constexpr int kR64ABytes = 4 * 64 * 128;            // one column-shifted halo copy: 32 KB
constexpr int kR64ASlots = 4;
constexpr int kR64BBytes = 64 * 128;                // this CTA's half of a 128-row weight tile: 8 KB
constexpr int kR64BSlots = 11;
constexpr int kR64Smem = kR64ASlots * kR64ABytes + kR64BSlots * kR64BBytes + (2 * kR64ASlots + 2 * kR64BSlots + 4) * 8 + 16 +
                         4 * 128 * 4 + 1024;
static_assert(kR64Smem <= 232448, "shared memory budget");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kHaloThreads, 1)
    conv_row64_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmW, const HaloParams P) {
  constexpr int N = 128;
  constexpr uint32_t kTmemCols = 256;                // two accumulator sets of 128 columns
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kR64ASlots * kR64ABytes;
  uint64_t* fullA = (uint64_t*)(sB + kR64BSlots * kR64BBytes);
  uint64_t* emptyA = fullA + kR64ASlots;
  uint64_t* fullB = emptyA + kR64ASlots;
  uint64_t* emptyB = fullB + kR64BSlots;
  uint64_t* tmem_full = emptyB + kR64BSlots;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* s_bias_all = reinterpret_cast<float*>(tmem_slot + 4);   // [4 epilogue warps][N]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int chunks = P.chunks0 + P.chunks1;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (P.chunks1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    for (int s = 0; s < kR64ASlots; ++s) {
      mbar_init(&fullA[s], 1);
      mbar_init(&emptyA[s], 1);
    }
    for (int s = 0; s < kR64BSlots; ++s) {
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);  // 4 epilogue warps of each CTA arrive on the leader's copy
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  cluster_sync_all();                // the peer is running before the pair allocation touches its shared memory (see conv_halo2_kernel)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs: own halo copies, own half of the weight rows) ----------------
      uint32_t a_it = 0, b_it = 0;
      int b_ahead = 0;               // weight tiles already requested ahead of the sequence (before the dependent-launch wait)
      bool waited = false;
      for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs) {
        const int nb = tile / P.tiles_hw, sp = tile - nb * P.tiles_hw;
        const int th = sp / P.tiles_w, tw = sp - th * P.tiles_w;
        const int h0 = (th * 2 + (int)rank) * 2, w0 = tw * 64, n0 = nb * N + (int)rank * (N / 2);
        for (int c = 0; c < chunks; ++c) {
          auto load_weight = [&](int dx, int dy) {
            const int sb = b_it % kR64BSlots;
            mbar_wait(&emptyB[sb], ((b_it / kR64BSlots) & 1u) ^ 1u);
            if (leader) mbar_expect_tx(&fullB[sb], 2 * kR64BBytes);
            tma_load_3d_pair(sB + sb * kR64BBytes, &tmW, &fullB[sb], c * kKChunk, n0, dy * 3 + dx);
            ++b_it;
          };
          if (!waited) {   // the first chunk's nine weight tiles, then the programmatic-dependent-launch wait, then activations
            for (int t = 0; t < 9; ++t) load_weight(t / 3, t % 3);
            b_ahead = 9;
            ptk_pdl_wait();
            ptk_pdl_trigger();
            waited = true;
          }
          for (int dx = 0; dx < 3; ++dx) {
            const int sa = a_it % kR64ASlots;
            mbar_wait(&emptyA[sa], ((a_it / kR64ASlots) & 1u) ^ 1u);
            if (leader) mbar_expect_tx(&fullA[sa], 2 * kR64ABytes);    // both CTAs' copies are credited here
            uint8_t* dst = sA + sa * kR64ABytes;
            if (c < P.chunks0) tma_load_3d_pair(dst, &tmA0, &fullA[sa], c * kKChunk, w0 + dx - 1, h0 - 1);
            else tma_load_3d_pair(dst, &tmA1, &fullA[sa], (c - P.chunks0) * kKChunk, w0 + dx - 1, h0 - 1);
            ++a_it;
            for (int dy = 0; dy < 3; ++dy) {
              if (b_ahead > 0) --b_ahead;
              else load_weight(dx, dy);
            }
          }
        }
      }
      if (!waited) {
        ptk_pdl_wait();
        ptk_pdl_trigger();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer of the pair (warp-uniform loop, one elected lane issues) ----------------
      uint32_t a_it = 0, b_it = 0, t_it = 0;
      for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs, ++t_it) {
        const uint32_t buf = t_it & 1u;
        mbar_wait(&tmem_empty[buf], ((t_it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * (uint32_t)N;
        for (int c = 0; c < chunks; ++c) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int sa = a_it % kR64ASlots;
            mbar_wait(&fullA[sa], (a_it / kR64ASlots) & 1u);
            ++a_it;
            int sb[3];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              sb[dy] = b_it % kR64BSlots;
              mbar_wait(&fullB[sb[dy]], (b_it / kR64BSlots) & 1u);
              ++b_it;
            }
            tc_fence_after();
            // rows (dy, dy + 1) of the copy = 128 consecutive pixels, core matrices 1 024 B apart
            const uint64_t adesc0 = make_sw128_desc(smem_u32(sA + sa * kR64ABytes));
            if (elect_one()) {
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                const uint64_t adesc = adesc0 + (uint64_t)(dy * 64 * 8);
                const uint64_t bdesc = make_sw128_desc(smem_u32(sB + sb[dy] * kR64BBytes));
#pragma unroll
                for (int k = 0; k < kKChunk / 16; ++k)
                  tc_mma_f16_pair(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, (c | dx | dy | k) != 0 ? 1u : 0u);
                tc_commit_pair(&emptyB[sb[dy]]);
              }
              tc_commit_pair(&emptyA[sa]);
            }
            __syncwarp();
          }
        }
        if (elect_one()) tc_commit_pair(&tmem_full[buf]);
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue (both CTAs, own TMEM lanes): lane m = pixel (row m / 64, column m % 64) ----------------
    const int q = warp & 3;
    const int m = q * 32 + lane;
    uint32_t t_it = 0;
    for (int tile = pair_id; tile < P.total_tiles; tile += n_pairs, ++t_it) {
      const int nb = tile / P.tiles_hw, sp = tile - nb * P.tiles_hw;
      const int th = sp / P.tiles_w, tw = sp - th * P.tiles_w;
      const int h = (th * 2 + (int)rank) * 2 + (m >> 6), w = tw * 64 + (m & 63), n0 = nb * N;
      const uint32_t buf = t_it & 1u;
      float* s_bias = s_bias_all + (warp - 2) * N;
      stage_bias<N>(s_bias, P.bias + n0, lane);
      const PairStore ps = pair_store_setup(P.out + ((size_t)h * P.W + w) * P.Cout + n0, h < P.H && w < P.W, lane);
      mbar_wait(&tmem_full[buf], (t_it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)N + (uint32_t)c, v);
        uint32_t pw[16];
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + c);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 bq = b4[j >> 1];
          float a = __uint_as_float(v[2 * j]) + ((j & 1) ? bq.z : bq.x);
          float b = __uint_as_float(v[2 * j + 1]) + ((j & 1) ? bq.w : bq.y);
          if (P.relu) {
            a = fmaxf(a, 0.f);
            b = fmaxf(b, 0.f);
          }
          const __half2 hv = __floats2half2_rn(a, b);
          pw[j] = *reinterpret_cast<const uint32_t*>(&hv);
        }
        pair_store(ps, c, pw, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader_relaxed(&tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync_all();                // the peer may still read this CTA's shared memory / TMEM until here
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// fp16 tensor [d2][d1][d0] (d0 innermost, contiguous), box {b0, b1, b2}, 128B swizzle, zero OOB fill
int make_map_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                uint64_t stride2_elems, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    ptk_set_error("cuTensorMapEncodeTiled entry point not available");
    return PTK_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
  const cuuint32_t box[3] = {b0, b1, b2};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ptk_set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu %llu %llu, box %u %u %u)", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1, b2);
    return PTK_ERR_CUDA;
  }
  return PTK_OK;
}

template <int BLOCK_N, int STAGES, int SPLIT = 1>
int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const ConvParams& P,
                cudaStream_t stream) {
  // stage ring | barriers + TMEM slot (256 B) | SPLIT: fp32 exchange buffer | alignment slack
  constexpr int smem = STAGES * (kABytes + BLOCK_N * 128) + 256 + (SPLIT > 1 ? BLOCK_N * 128 * 4 : 0) + BLOCK_N * 4 + 1024;
  static_assert(smem <= 232448, "shared memory budget");
  static_assert(2 * STAGES + 2 <= 31, "barriers + TMEM slot must fit the 256-byte block");
  static bool configured = false;
  if (!configured) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, STAGES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tile_h = 128 >> P.tw_log2;
  const int tiles_h = (P.H + tile_h - 1) / tile_h;
  dim3 grid(P.tiles_w * tiles_h, P.Cout / BLOCK_N, SPLIT);
  if (SPLIT == 1) {
    PTK_CUDA_CHECK(ptk_launch_pdl(conv_tc_kernel<BLOCK_N, STAGES, SPLIT>, grid, dim3(kConvThreads), smem, stream, dim3(1, 1, 1), a0, a1, w, P));
  } else {
    PTK_CUDA_CHECK(ptk_launch_pdl(conv_tc_kernel<BLOCK_N, STAGES, SPLIT>, grid, dim3(kConvThreads), smem, stream, dim3(1, 1, SPLIT), a0, a1, w, P));
  }
  return PTK_OK;
}


template <int N, int SPLIT = 1>
int launch_halo(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const HaloParams& P, int grid,
                cudaStream_t stream) {
  constexpr int kSlots = kBBudget / (N * 128);
  constexpr int smem = 2 * kHaloStage + kSlots * N * 128 + (4 + 2 * kSlots + 6) * 8 + 16 + 8 * N * 4 + 1024;
  static_assert(smem <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (SPLIT == 1) {
    PTK_CUDA_CHECK(ptk_launch_pdl(conv_halo_kernel<N, SPLIT>, dim3(grid), dim3(kHalo1Threads), smem, stream, dim3(1, 1, 1), a0, a1, w, P,
                                  HeadArg<false>()));
  } else {   // one cluster of SPLIT CTAs per tile
    PTK_CUDA_CHECK(ptk_launch_pdl(conv_halo_kernel<N, SPLIT>, dim3(SPLIT * P.total_tiles), dim3(kHalo1Threads), smem, stream,
                                  dim3(SPLIT, 1, 1), a0, a1, w, P, HeadArg<false>()));
  }
  PTK_CUDA_CHECK(cudaGetLastError());
  return PTK_OK;
}

// conv_halo_kernel<32> with the level-0 head in its epilogue
int launch_halo32_head(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const HaloParams& P, int grid,
                       cudaStream_t stream, const PtkHeadConst& hc) {
  constexpr int kSlots = kBBudget / (32 * 128);
  constexpr int smem = 2 * kHaloStage + kSlots * 32 * 128 + (4 + 2 * kSlots + 6) * 8 + 16 + 8 * 32 * 4 + 1024;
  static bool configured = false;
  if (!configured) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<32, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  HeadArg<true> ha;
  ha.c = hc;
  PTK_CUDA_CHECK(ptk_launch_pdl(conv_halo_kernel<32, 1, true>, dim3(grid), dim3(kHalo1Threads), smem, stream, dim3(1, 1, 1), a0, a1, w, P, ha));
  return PTK_OK;
}

template <int N>
int launch_halo2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const HaloParams& P, int n_pairs,
                 cudaStream_t stream) {
  constexpr int kSlots = kPairBudget / ((N / 2) * 128);
  constexpr int smem = 2 * kHaloStage + kPairBudget + (4 + 2 * kSlots + 4) * 8 + 16 + 8 * N * 4 + 1024;
  static_assert(smem <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(conv_halo2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  // the cluster shape is the kernel's compile-time __cluster_dims__(2, 1, 1)
  PTK_CUDA_CHECK(ptk_launch_pdl(conv_halo2_kernel<N>, dim3(2 * n_pairs), dim3(kPair8Threads), smem, stream, dim3(1, 1, 1), a0, a1, w, P));
  return PTK_OK;
}

int launch_row64(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const HaloParams& P, int n_pairs,
                 cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(conv_row64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kR64Smem));
    configured = true;
  }
  // the cluster shape is the kernel's compile-time __cluster_dims__(2, 1, 1)
  PTK_CUDA_CHECK(ptk_launch_pdl(conv_row64_kernel, dim3(2 * n_pairs), dim3(kHaloThreads), kR64Smem, stream, dim3(1, 1, 1), a0, a1, w, P));
  return PTK_OK;
}

template <int R, int N>
int launch_row2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const HaloParams& P, int n_pairs,
                cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    PTK_CUDA_CHECK(cudaFuncSetAttribute(conv_row2_kernel<R, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, RowGeom<R, N>::kSmem));
    configured = true;
  }
  // the cluster shape is the kernel's compile-time __cluster_dims__(2, 1, 1)
  PTK_CUDA_CHECK(ptk_launch_pdl(conv_row2_kernel<R, N>, dim3(2 * n_pairs), dim3(kHaloThreads), RowGeom<R, N>::kSmem, stream, dim3(1, 1, 1), a0,
                                a1, w, P));
  return PTK_OK;
}
int launch_row2_any(int r, int n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const HaloParams& P, int n_pairs,
                    cudaStream_t s) {
  if (n == 64) {
    if (r == 1) return launch_row2<1, 64>(a0, a1, w, P, n_pairs, s);
    if (r == 2) return launch_row2<2, 64>(a0, a1, w, P, n_pairs, s);
    return launch_row2<3, 64>(a0, a1, w, P, n_pairs, s);
  }
  if (r == 1) return launch_row2<1, 128>(a0, a1, w, P, n_pairs, s);
  if (r == 2) return launch_row2<2, 128>(a0, a1, w, P, n_pairs, s);
  return launch_row2<3, 128>(a0, a1, w, P, n_pairs, s);
}

}  // namespace

// Picks the N tile so that small maps still fill the machine (>= ~1 CTA per SM when possible).
static int pick_block_n(int Cout, int m_tiles, int num_sms) {
  int bn = Cout >= 256 ? 256 : Cout;
  while (bn > 64 && m_tiles * (Cout / bn) < num_sms) bn >>= 1;
  return bn;
}

// ptk_conv_f16 plus an optional fused 2x2 max pool of the result (the extractor plan's encoder blocks)
static int conv_dispatch(PtkContext* ctx, const void* in0, int32_t cin0, const void* in1, int32_t cin1, int32_t H, int32_t W,
                         int32_t in0_H, int32_t in0_W, int32_t in1_H, int32_t in1_W, const void* weights, const float* bias,
                         int32_t Cout, int32_t taps, int32_t relu, void* out, void* pool_out, void* stream,
                         const PtkHeadConst* head, int* head_fused) {
  if (head_fused != nullptr) *head_fused = 0;
  PTK_REQUIRE(ctx && in0 && weights && bias && out, "null argument");
  PtkDeviceGuard guard(ctx->device);
  PTK_REQUIRE(taps == 9 || taps == 1, "taps must be 9 (3x3, pad 1) or 1 (1x1)");
  PTK_REQUIRE(cin0 > 0 && cin0 % kKChunk == 0 && cin1 >= 0 && cin1 % kKChunk == 0, "C_in must be a multiple of 64");
  PTK_REQUIRE(Cout == 32 || Cout == 64 || Cout == 128 || Cout % 256 == 0, "C_out must be 32, 64, 128 or k*256");
  PTK_REQUIRE((cin1 == 0) == (in1 == nullptr), "in1 and cin1 must be given together");
  PTK_REQUIRE(H >= 1 && W >= 1 && in0_H >= H && in0_W >= W, "input 0 smaller than the output");
  const int ctot = cin0 + cin1;
  PTK_REQUIRE(cin1 == 0 || (in1_H >= H && in1_W >= W), "input 1 smaller than the output");
  CUtensorMap a0, a1, wm;
  int rc;
  cudaStream_t s = (cudaStream_t)stream;
  // ---- halo kernel (3x3 only): one halo load per chunk, persistent CTAs ----
  {
    // PTK_CONV_HALO: 0 = never, 1 = when it fills the machine (default), 2 = whenever legal.  PTK_CONV_HALO_ROWS: halo rows
    // per TMA operation (18 = the whole halo in one box, the default: pieces would start at multiples of 2 304 B)
    static int mode = -1, rows_per_op = 18;
    if (mode < 0) {
      const char* e = getenv("PTK_CONV_HALO");
      mode = e ? atoi(e) : 1;
      const char* r = getenv("PTK_CONV_HALO_ROWS");
      if (r && (atoi(r) == 1 || atoi(r) == 2 || atoi(r) == 3 || atoi(r) == 6 || atoi(r) == 9 || atoi(r) == 18)) rows_per_op = atoi(r);
    }
    const int tiles_w16 = (W + 15) / 16, tiles_h16 = (H + 15) / 16;
    const int n_halo = Cout >= 128 ? 128 : Cout;
    const int total = tiles_w16 * tiles_h16 * (Cout / n_halo);
    // CTA pairs (cta_group::2): C_out = k * 256 and enough pairs of 16x16 tiles for most of a wave of SM pairs
    {
      // PTK_CONV_PAIR: 0 = never, 1 = when the pairs fill a wave and the weights are streamed (default), 2 = whenever legal.
      // Measured on the 256-channel block (144x256, 43.5 GFLOP per layer), 10 launches back to back: single-CTA halo
      // kernel 44.7 us (972 TF/s: 17 % of the MMA warp's time waits for weight tiles, 84 cycles per MMA); pair with
      // N = 256 (one accumulator set, epilogue not overlapped) 38.9 us; pair with N = 128 (two accumulator sets)
      // **34.5-35.8 us = 1 210-1 260 TF/s**.  Also N = 64 (decoder blocks 64+128->64 at 288x512: 51.2 -> 38.5 us,
      // 64+256->64 at 144x256: 28.6 -> 22.0 us; their weights do not fit the single-CTA kernel's resident slots, and 20 %
      // of its MMA warp's time waits for weight tiles).  `profiles/r2/conv_layers_pair_forced.txt` has every layer as pairs.
      static int pair_mode = -1, pair_n_env = 128;   // PTK_CONV_PAIR_N = 128 (double-buffered accumulators, default) | 256
      if (pair_mode < 0) {
        const char* e = getenv("PTK_CONV_PAIR");
        pair_mode = e ? atoi(e) : 1;
        const char* n = getenv("PTK_CONV_PAIR_N");
        if (n && atoi(n) == 256) pair_n_env = 256;
      }
      const int pair_n = Cout == 32 ? 32 : (Cout == 64 ? 64 : ((pair_n_env == 256 && Cout % 256 == 0) ? 256 : 128));
      const int pairs_w = (tiles_w16 + 1) / 2;
      const int total_pairs = pairs_w * tiles_h16 * (Cout / pair_n);
      const int pair_slots = ctx->num_sms / 2;
      const int pwaves = (total_pairs + pair_slots - 1) / pair_slots;
      const bool pair_legal = taps == 9 && Cout % pair_n == 0 && mode != 0;
      // not for layers whose weights the single-CTA kernel keeps resident in shared memory (C_out <= 64 with few
      // chunks): there is no weight traffic to halve, and the pair's ring re-streams them per tile (64->64 at full
      // resolution: 62 us single, 84 us as pairs)
      const bool single_resident = Cout == n_halo && 9 * (ctot / kKChunk) <= kBBudget / (n_halo * 128);
      // ... unless the pair keeps them resident too (PTK_CONV_PAIR_RES, default 1): each SM then reads half of every
      // weight tile per MMA, which is what the thin full-resolution layers are bound by (operand fetch from shared
      // memory).  Measured, 10 launches back to back: 64+64->32 @576x1024 83.7 -> 59.8 us, 64+128->64 @288x512
      // 33.5 -> 30.3, 64->128 @288x512 24.8 -> 23.0.  The one-chunk 64->64 layer (+pool) at first LOST as a pair
      // (54.8 -> 68.9 us: its epilogue outlasts its MMAs, and every epilogue warp waited at a GPU-scope fence per tile); with
      // eight epilogue warps and the relaxed accumulator-release arrive it wins too: 54.8 -> ~48 us at 576x1024,
      // 68.3 -> 58.0 us at 756x1008.  (PTK_CONV_PAIR_RES=3 keeps that shape on the single-CTA kernel.)
      static int pair_res_mode = -1;
      if (pair_res_mode < 0) pair_res_mode = getenv("PTK_CONV_PAIR_RES") ? atoi(getenv("PTK_CONV_PAIR_RES")) : 1;
      const bool pair_resident = Cout == pair_n && 9 * (ctot / kKChunk) <= kPairBudget / ((pair_n / 2) * 128) &&
                                 !(ctot == kKChunk && Cout <= 64 && pair_res_mode == 3);
      const bool pair_wanted = pair_mode == 2 ||
                               (pair_mode == 1 && (!single_resident || (pair_res_mode != 0 && pair_resident)) &&
                                total_pairs * 10 >= pwaves * pair_slots * 8);
      if (pair_legal && pair_wanted && head == nullptr) {   // (a fused head only exists in the single-CTA kernel's epilogue)
        rc = make_map_3d(&a0, in0, cin0, W, H, cin0, (uint64_t)in0_W * cin0, kKChunk, kHaloW, rows_per_op);
        if (rc != PTK_OK) return rc;
        if (cin1 > 0) {
          rc = make_map_3d(&a1, in1, cin1, W, H, cin1, (uint64_t)in1_W * cin1, kKChunk, kHaloW, rows_per_op);
          if (rc != PTK_OK) return rc;
        } else {
          a1 = a0;
        }
        rc = make_map_3d(&wm, weights, ctot, Cout, taps, ctot, (uint64_t)Cout * ctot, kKChunk, pair_n / 2, 1);
        if (rc != PTK_OK) return rc;
        HaloParams Q;
        Q.H = H; Q.W = W; Q.Cout = Cout; Q.chunks0 = cin0 / kKChunk; Q.chunks1 = cin1 / kKChunk; Q.relu = relu;
        Q.tiles_w = pairs_w; Q.tiles_hw = pairs_w * tiles_h16; Q.total_tiles = total_pairs;
        Q.resident = (pair_res_mode != 0 && pair_resident) ? 1 : 0;
        Q.dbg = nullptr;
        Q.rows_per_op = rows_per_op;
        Q.bias = bias;
        Q.out = (__half*)out;
        Q.pool = (__half*)pool_out;
        const int n_launch = total_pairs < pair_slots ? total_pairs : pair_slots;
        static int pair_dbg = -1;
        if (pair_dbg < 0) pair_dbg = getenv("PTK_CONV_DBG") ? atoi(getenv("PTK_CONV_DBG")) : 0;
        if (pair_dbg) {   // stall attribution (debug only: synchronises and prints)
          static long long* dbuf = nullptr;
          if (!dbuf) cudaMalloc(&dbuf, 16 * sizeof(long long));
          cudaMemsetAsync(dbuf, 0, 16 * sizeof(long long), s);
          Q.dbg = dbuf;
          const int lrc = pair_n == 32 ? launch_halo2<32>(a0, a1, wm, Q, n_launch, s)
                          : pair_n == 64 ? launch_halo2<64>(a0, a1, wm, Q, n_launch, s)
                                         : (pair_n == 256 ? launch_halo2<256>(a0, a1, wm, Q, n_launch, s) : launch_halo2<128>(a0, a1, wm, Q, n_launch, s));
          long long h[16];
          cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
          fprintf(stderr, "[halo2 N=%d %dx%d cin=%d cout=%d tiles/pair=%lld] producer: total %lld waitEmptyA %lld waitEmptyB %lld | "
                  "mma: total %lld waitTmemEmpty %lld waitFullA %lld waitFullB %lld | epilogue: total %lld waitTmemFull %lld tmemLoad %lld store %lld\n",
                  pair_n, H, W, cin0 + cin1, Cout, h[7], h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[8], h[9], h[10], h[11]);
          return lrc;
        }
        if (pair_n == 32) return launch_halo2<32>(a0, a1, wm, Q, n_launch, s);
        if (pair_n == 64) return launch_halo2<64>(a0, a1, wm, Q, n_launch, s);
        return pair_n == 256 ? launch_halo2<256>(a0, a1, wm, Q, n_launch, s) : launch_halo2<128>(a0, a1, wm, Q, n_launch, s);
      }
      // Maps at most 64 pixels wide with a long K and many channels (the 1/16-scale block): conv_row64_kernel.
      // PTK_CONV_ROW64: 0 = never, 1 = maps 48..64 pixels wide with >= 44 pair tiles (default), 2 = whenever legal.
      // Measured, 512->512, 10 launches back to back: 47 x 63 (48 pair tiles) 29.4 -> 27.6 us; 36 x 64 (36 pair tiles =
      // half of the SMs, each CTA pulling 72 B/clk at the MMA rate) 21.9 -> 26.1 us, so that one stays on the per-tap kernel.
      static int row64_mode = -1;
      if (row64_mode < 0) row64_mode = getenv("PTK_CONV_ROW64") ? atoi(getenv("PTK_CONV_ROW64")) : 1;
      if (row64_mode != 0 && mode != 0 && taps == 9 && Cout % 128 == 0 && pool_out == nullptr) {
        const int col_tiles = (W + 63) / 64, row_tiles = (H + 3) / 4;
        const int tiles = col_tiles * row_tiles * (Cout / 128);
        const bool wanted64 = row64_mode == 2 || (ctot >= 256 && W <= 64 && W >= 48 && tiles >= 44 && tiles <= 2 * pair_slots);
        if (wanted64) {
          rc = make_map_3d(&a0, in0, cin0, W, H, cin0, (uint64_t)in0_W * cin0, kKChunk, 64, 4);
          if (rc != PTK_OK) return rc;
          if (cin1 > 0) {
            rc = make_map_3d(&a1, in1, cin1, W, H, cin1, (uint64_t)in1_W * cin1, kKChunk, 64, 4);
            if (rc != PTK_OK) return rc;
          } else {
            a1 = a0;
          }
          rc = make_map_3d(&wm, weights, ctot, Cout, taps, ctot, (uint64_t)Cout * ctot, kKChunk, 64, 1);
          if (rc != PTK_OK) return rc;
          HaloParams Q;
          Q.H = H; Q.W = W; Q.Cout = Cout; Q.chunks0 = cin0 / kKChunk; Q.chunks1 = cin1 / kKChunk; Q.relu = relu;
          Q.tiles_w = col_tiles; Q.tiles_hw = col_tiles * row_tiles; Q.total_tiles = tiles;
          Q.resident = 0;
          Q.dbg = nullptr;
          Q.rows_per_op = 4;
          Q.bias = bias;
          Q.out = (__half*)out;
          Q.pool = nullptr;
          return launch_row64(a0, a1, wm, Q, tiles < pair_slots ? tiles : pair_slots, s);
        }
      }
      // Row tiles on CTA pairs (conv_row2_kernel): maps about 128 pixels wide with a long K, where 16x16 tiles quantise badly
      // and the per-tap kernel is bound by its L2 -> shared traffic.  PTK_CONV_ROW: 0 = never, 1 = when one of R = 2 / 3
      // fills >= 60 % of its waves of SM pairs (default; the pooled 94 x 126 layer needs R = 2 = 96 tiles on 74 pairs and
      // still beats the per-tap kernel, 61 us), 2 = whenever legal.
      static int row_mode = -1;
      if (row_mode < 0) row_mode = getenv("PTK_CONV_ROW") ? atoi(getenv("PTK_CONV_ROW")) : 1;
      const int row_n = Cout == 64 ? 64 : 128;      // N = 64: the decoder's 64-channel layers with a long K
      if (row_mode != 0 && mode != 0 && taps == 9 && Cout % row_n == 0 && ctot >= 256) {
        const int col_tiles = (W + 127) / 128;
        int best_r = 0;
        double best_eff = 0.0;
        for (int r = 3; r >= 1; --r) {                        // (ties go to the larger R: less halo per output row)
          if (pool_out != nullptr && (r & 1)) continue;       // pooled rows must pair up inside one CTA
          const long tiles = (long)col_tiles * ((H + 2 * r - 1) / (2 * r)) * (Cout / row_n);
          const long wv = (tiles + pair_slots - 1) / pair_slots;
          const double eff = (double)H * W * (Cout / row_n) / ((double)wv * pair_slots * 2 * r * 128);
          if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best_r = r;
          }
        }
        static double need = -1.0;      // PTK_CONV_ROW_EFF: smallest fill of the SM pairs this kernel is chosen for
        if (need < 0) need = getenv("PTK_CONV_ROW_EFF") ? atof(getenv("PTK_CONV_ROW_EFF")) : 0.6;
        if (best_r != 0 && (row_mode == 2 || best_eff >= need)) {
          rc = make_map_3d(&a0, in0, cin0, W, H, cin0, (uint64_t)in0_W * cin0, kKChunk, kRowW, best_r + 2);
          if (rc != PTK_OK) return rc;
          if (cin1 > 0) {
            rc = make_map_3d(&a1, in1, cin1, W, H, cin1, (uint64_t)in1_W * cin1, kKChunk, kRowW, best_r + 2);
            if (rc != PTK_OK) return rc;
          } else {
            a1 = a0;
          }
          rc = make_map_3d(&wm, weights, ctot, Cout, taps, ctot, (uint64_t)Cout * ctot, kKChunk, row_n / 2, 1);
          if (rc != PTK_OK) return rc;
          HaloParams Q;
          Q.H = H; Q.W = W; Q.Cout = Cout; Q.chunks0 = cin0 / kKChunk; Q.chunks1 = cin1 / kKChunk; Q.relu = relu;
          Q.tiles_w = col_tiles; Q.tiles_hw = col_tiles * ((H + 2 * best_r - 1) / (2 * best_r));
          Q.total_tiles = Q.tiles_hw * (Cout / row_n);
          Q.resident = 0;
          Q.dbg = nullptr;
          Q.rows_per_op = best_r + 2;
          Q.bias = bias;
          Q.out = (__half*)out;
          Q.pool = (__half*)pool_out;
          const int n_launch = Q.total_tiles < pair_slots ? Q.total_tiles : pair_slots;
          static int row_dbg = -1;
          if (row_dbg < 0) row_dbg = getenv("PTK_CONV_DBG") ? atoi(getenv("PTK_CONV_DBG")) : 0;
          if (row_dbg) {   // stall attribution (debug only: synchronises and prints)
            static long long* dbuf = nullptr;
            if (!dbuf) cudaMalloc(&dbuf, 16 * sizeof(long long));
            cudaMemsetAsync(dbuf, 0, 16 * sizeof(long long), s);
            Q.dbg = dbuf;
            const int lrc = launch_row2_any(best_r, row_n, a0, a1, wm, Q, n_launch, s);
            long long h[16];
            cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[row2 R=%d N=%d %dx%d cin=%d cout=%d tiles/pair=%lld] producer: total %lld waitEmptyA %lld waitEmptyB %lld | "
                    "mma: total %lld waitTmemEmpty %lld waitFullA %lld waitFullB %lld | epilogue: total %lld waitTmemFull %lld\n",
                    best_r, row_n, H, W, cin0 + cin1, Cout, h[7], h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[8], h[9]);
            return lrc;
          }
          return launch_row2_any(best_r, row_n, a0, a1, wm, Q, n_launch, s);
        }
      }
    }
    const bool legal = taps == 9 && (Cout == 32 || Cout == 64 || Cout % 128 == 0);
    // the 16x16 tiles are coarse: use them only when their last wave is reasonably full, otherwise the
    // finer-grained kernel fills the machine better
    const int waves = (total + ctx->num_sms - 1) / ctx->num_sms;
    const bool wanted = mode == 2 || (mode == 1 && total * 10 >= waves * ctx->num_sms * 8);
    // Small maps with a long K (the 1/16-scale block, the first decoder convolution): too few tiles for one CTA each,
    // but two CTAs per tile, each running half of the 64-channel chunks, fill half to all of the machine -- and the
    // halo staging moves 2-3x fewer L2 -> shared-memory bytes than the per-tap kernel, which is what bounds these
    // layers (PTK_CONV_HALO_SPLIT=0 switches it off).
    static int hsplit = -1;
    if (hsplit < 0) hsplit = getenv("PTK_CONV_HALO_SPLIT") ? atoi(getenv("PTK_CONV_HALO_SPLIT")) : 0;
    const int chunks_all = ctot / kKChunk;
    const bool split_wanted = hsplit != 0 && mode != 0 && legal && !wanted && chunks_all >= 4 &&
                              2 * total <= ctx->num_sms && (hsplit == 2 || 4 * total >= ctx->num_sms);
    if (split_wanted) {
      rc = make_map_3d(&a0, in0, cin0, W, H, cin0, (uint64_t)in0_W * cin0, kKChunk, kHaloW, rows_per_op);
      if (rc != PTK_OK) return rc;
      if (cin1 > 0) {
        rc = make_map_3d(&a1, in1, cin1, W, H, cin1, (uint64_t)in1_W * cin1, kKChunk, kHaloW, rows_per_op);
        if (rc != PTK_OK) return rc;
      } else {
        a1 = a0;
      }
      rc = make_map_3d(&wm, weights, ctot, Cout, taps, ctot, (uint64_t)Cout * ctot, kKChunk, n_halo, 1);
      if (rc != PTK_OK) return rc;
      HaloParams Q;
      Q.H = H; Q.W = W; Q.Cout = Cout; Q.chunks0 = cin0 / kKChunk; Q.chunks1 = cin1 / kKChunk; Q.relu = relu;
      Q.tiles_w = tiles_w16; Q.tiles_hw = tiles_w16 * tiles_h16; Q.total_tiles = total;
      Q.resident = 0;
      Q.rows_per_op = rows_per_op;
      Q.bias = bias;
      Q.out = (__half*)out;
      Q.pool = (__half*)pool_out;
      Q.dbg = nullptr;
      switch (n_halo) {
        case 32: return launch_halo<32, 2>(a0, a1, wm, Q, 2 * total, s);
        case 64: return launch_halo<64, 2>(a0, a1, wm, Q, 2 * total, s);
        default: return launch_halo<128, 2>(a0, a1, wm, Q, 2 * total, s);
      }
    }
    if (legal && wanted) {
      rc = make_map_3d(&a0, in0, cin0, W, H, cin0, (uint64_t)in0_W * cin0, kKChunk, kHaloW, rows_per_op);
      if (rc != PTK_OK) return rc;
      if (cin1 > 0) {
        rc = make_map_3d(&a1, in1, cin1, W, H, cin1, (uint64_t)in1_W * cin1, kKChunk, kHaloW, rows_per_op);
        if (rc != PTK_OK) return rc;
      } else {
        a1 = a0;
      }
      rc = make_map_3d(&wm, weights, ctot, Cout, taps, ctot, (uint64_t)Cout * ctot, kKChunk, n_halo, 1);
      if (rc != PTK_OK) return rc;
      HaloParams Q;
      Q.H = H; Q.W = W; Q.Cout = Cout; Q.chunks0 = cin0 / kKChunk; Q.chunks1 = cin1 / kKChunk; Q.relu = relu;
      Q.tiles_w = tiles_w16; Q.tiles_hw = tiles_w16 * tiles_h16; Q.total_tiles = total;
      Q.resident = (Cout == n_halo && 9 * (Q.chunks0 + Q.chunks1) <= kBBudget / (n_halo * 128)) ? 1 : 0;
      Q.rows_per_op = rows_per_op;
      Q.bias = bias;
      Q.out = (__half*)out;
      Q.pool = (__half*)pool_out;
      Q.dbg = nullptr;
      const int grid = total < ctx->num_sms ? total : ctx->num_sms;
      if (head != nullptr && n_halo == 32 && Cout == 32) {   // the level-0 head rides in this kernel's epilogue
        if (head_fused != nullptr) *head_fused = 1;
        return launch_halo32_head(a0, a1, wm, Q, grid, s, *head);
      }
      static int dbg_mode = -1;
      if (dbg_mode < 0) dbg_mode = getenv("PTK_CONV_DBG") ? atoi(getenv("PTK_CONV_DBG")) : 0;
      if (dbg_mode) {   // stall attribution (debug only: synchronises and prints)
        static long long* dbuf = nullptr;
        if (!dbuf) cudaMalloc(&dbuf, 16 * sizeof(long long));
        cudaMemsetAsync(dbuf, 0, 16 * sizeof(long long), s);
        Q.dbg = dbuf;
        int lrc;
        switch (n_halo) {
          case 32: lrc = launch_halo<32>(a0, a1, wm, Q, grid, s); break;
          case 64: lrc = launch_halo<64>(a0, a1, wm, Q, grid, s); break;
          default: lrc = launch_halo<128>(a0, a1, wm, Q, grid, s); break;
        }
        long long h[16];
        cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[halo N=%d %dx%d cin=%d cout=%d res=%d tiles/cta=%lld] producer: total %lld waitEmptyA %lld waitEmptyB %lld | "
                "mma: total %lld waitTmemEmpty %lld waitFullA %lld waitFullB %lld | epilogue: total %lld waitTmemFull %lld\n",
                n_halo, H, W, cin0 + cin1, Cout, Q.resident, h[7], h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[8], h[9]);
        return lrc;
      }
      switch (n_halo) {
        case 32: return launch_halo<32>(a0, a1, wm, Q, grid, s);
        case 64: return launch_halo<64>(a0, a1, wm, Q, grid, s);
        default: return launch_halo<128>(a0, a1, wm, Q, grid, s);
      }
    }
  }
  ConvParams P;
  P.H = H; P.W = W; P.Cout = Cout; P.cin0 = cin0; P.cin1 = cin1; P.taps = taps; P.relu = relu;
  // tile shape: the 128 pixels of a CTA as 16x8 (default; the only shape the fused pool handles), or 32x4 / 64x2 /
  // 8x16 when that covers the map with fewer tiles (e.g. 64x36: 18 tiles of 32x4 instead of 20 of 16x8)
  P.tw_log2 = 4;
  {
    static int shapes = -1;   // PTK_CONV_SHAPES=0 pins 16x8 (A/B measurements)
    if (shapes < 0) shapes = getenv("PTK_CONV_SHAPES") ? atoi(getenv("PTK_CONV_SHAPES")) : 1;
    int best = ((W + 15) / 16) * ((H + 7) / 8);
    if (pool_out == nullptr && shapes) {
      const int cand[3] = {5, 6, 3};
      for (int lg : cand) {
        const int n = ((W + (1 << lg) - 1) >> lg) * ((H + (128 >> lg) - 1) / (128 >> lg));
        if (n < best) {
          best = n;
          P.tw_log2 = lg;
        }
      }
    }
  }
  const int tile_w = 1 << P.tw_log2, tile_h = 128 >> P.tw_log2;
  // extents = the OUTPUT size: anything beyond (incl. a larger skip tensor's extra rows/cols) reads as zero
  rc = make_map_3d(&a0, in0, cin0, W, H, cin0, (uint64_t)in0_W * cin0, kKChunk, tile_w, tile_h);
  if (rc != PTK_OK) return rc;
  if (cin1 > 0) {
    rc = make_map_3d(&a1, in1, cin1, W, H, cin1, (uint64_t)in1_W * cin1, kKChunk, tile_w, tile_h);
    if (rc != PTK_OK) return rc;
  } else {
    a1 = a0;
  }
  P.tiles_w = (W + tile_w - 1) / tile_w;
  P.bias = bias;
  P.out = (__half*)out;
  P.pool = (__half*)pool_out;
  const int m_tiles = P.tiles_w * ((H + tile_h - 1) / tile_h);
  const int bn = pick_block_n(Cout, m_tiles, ctx->num_sms);
  rc = make_map_3d(&wm, weights, ctot, Cout, taps, ctot, (uint64_t)Cout * ctot, kKChunk, bn, 1);
  if (rc != PTK_OK) return rc;
  static int split_mode = -1;   // PTK_CONV_SPLIT=0 disables the two-CTA K split
  if (split_mode < 0) split_mode = getenv("PTK_CONV_SPLIT") ? atoi(getenv("PTK_CONV_SPLIT")) : 1;
  const int ctas = m_tiles * (Cout / bn);
  const int k_steps = taps * (ctot / kKChunk);
  switch (bn) {
    case 32: return launch_conv<32, 4>(a0, a1, wm, P, s);
    case 64:
      // many output channels on a small map: the layer is bound by L2 -> shared-memory bytes, and a 128-channel
      // tile shared by two CTAs (half of K each) stages 1/3 fewer bytes than two 64-channel tiles
      if (split_mode && Cout % 128 == 0 && 2 * m_tiles * (Cout / 128) <= ctx->num_sms && k_steps >= 16 && pool_out == nullptr) {
        rc = make_map_3d(&wm, weights, ctot, Cout, taps, ctot, (uint64_t)Cout * ctot, kKChunk, 128, 1);
        if (rc != PTK_OK) return rc;
        return launch_conv<128, 5, 2>(a0, a1, wm, P, s);
      }
      // tiles for at most half of the SMs and a long K loop: two CTAs per tile, each runs half of K
      if (split_mode && 2 * ctas <= ctx->num_sms && k_steps >= 16 && pool_out == nullptr)
        return launch_conv<64, 8, 2>(a0, a1, wm, P, s);
      // at most one CTA per SM: nothing else hides the load latency, so run a deeper ring
      if (ctas <= ctx->num_sms) return launch_conv<64, 8>(a0, a1, wm, P, s);
      return launch_conv<64, 4>(a0, a1, wm, P, s);
    case 128: return launch_conv<128, 3>(a0, a1, wm, P, s);
    default: return launch_conv<256, 3>(a0, a1, wm, P, s);
  }
}

extern "C" int ptk_conv_f16_pool(PtkContext* ctx, const void* in0, int32_t cin0, const void* in1, int32_t cin1, int32_t H, int32_t W,
                      int32_t in0_H, int32_t in0_W, int32_t in1_H, int32_t in1_W, const void* weights, const float* bias,
                      int32_t Cout, int32_t taps, int32_t relu, void* out, void* pool_out, void* stream) {
  return conv_dispatch(ctx, in0, cin0, in1, cin1, H, W, in0_H, in0_W, in1_H, in1_W, weights, bias, Cout, taps, relu, out, pool_out,
                       stream, nullptr, nullptr);
}

// ptk_conv_f16 with the extractor's level-0 head fused into the epilogue when the layer lands on conv_halo_kernel<32>
// (*fused tells; otherwise the caller launches the head itself).  Library-internal.
int ptk_conv_f16_head(PtkContext* ctx, const void* in0, int32_t cin0, const void* in1, int32_t cin1, int32_t H, int32_t W,
                      int32_t in0_H, int32_t in0_W, int32_t in1_H, int32_t in1_W, const void* weights, const float* bias,
                      int32_t Cout, int32_t taps, int32_t relu, void* out, void* stream, const PtkHeadConst* head, int* fused) {
  return conv_dispatch(ctx, in0, cin0, in1, cin1, H, W, in0_H, in0_W, in1_H, in1_W, weights, bias, Cout, taps, relu, out, nullptr,
                       stream, head, fused);
}

extern "C" int ptk_conv_f16(PtkContext* ctx, const void* in0, int32_t cin0, const void* in1, int32_t cin1, int32_t H,
                            int32_t W, int32_t in0_H, int32_t in0_W, int32_t in1_H, int32_t in1_W, const void* weights,
                            const float* bias, int32_t Cout, int32_t taps, int32_t relu, void* out, void* stream) {
  return ptk_conv_f16_pool(ctx, in0, cin0, in1, cin1, H, W, in0_H, in0_W, in1_H, in1_W, weights, bias, Cout, taps, relu, out,
                           nullptr, stream);
}
