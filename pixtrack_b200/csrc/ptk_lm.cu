// Fused Levenberg-Marquardt pose refinement for sm_100a (B200).
//
// One persistent launch runs the WHOLE damped Gauss-Newton loop of
// PixTrackOptimizer.run (reference pixloc/pixloc/pixlib/models/
// learned_optimizer.py:48-95 + pixtrack/optimizers/pixtrack_optimizer.py:6-18)
// for B independent problems: projection, bilinear sampling of the query map
// and of its confidence, residual, robust weight, Jacobian chain, 6x6 normal
// equations, damped Cholesky solve, SE(3) update, stop test, per-iteration log.
// Nothing returns to the host between iterations.
//
// Mapping to the machine
//  * HBM/L2-bound gather kernel (about 3 flop/B): no tensor cores.  The query
//    map is channels-last [H][W][C], so one texel is C*4 contiguous bytes and
//    LPP = C/4 lanes fetch it with one 16-byte load each (a full warp reads one
//    512-byte texel at C=128; at C=32 a warp serves 4 points at once).  All 12
//    texels of the 5-sample cross footprint are requested before any is used.
//  * The N x C x 6 Jacobian is never formed.  With J_c = [gx_c gy_c] * A
//    (A = 2x6 projection chain of the point) the per-point sums collapse to
//      H_n = A^T [Sxx Sxy; Sxy Syy] A,   g_n = A^T [Sxr; Syr],
//    so only 6 channel sums (+1 confidence) cross lanes (xor shuffles).
//  * Scalar per-point work (projection, robust weight, Jacobian chain, the 29
//    accumulated scalars: 21 H, 6 g, cost, count) is done by ONE thread per point;
//    the warps that gather texels do nothing else (see lm_kernel).  thread -> warp
//    -> CTA -> problem reduction is fixed-order (bit-reproducible), no float atomics.  G CTAs share one problem; they meet once per iteration at a
//    counter barrier in global memory, then every CTA redundantly sums the G
//    partial vectors in the same order, solves the 6x6 system and updates its own
//    copy of the pose -- identical arithmetic, so no broadcast is needed.
//  * Grid <= number of SMs (one 512-thread CTA per SM), launched cooperatively
//    so the spin barrier is safe.
#include <stdlib.h>

#include "ptk_common.cuh"

namespace {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kEntries = 29;             // 21 H (upper) + 6 g + cost_sum + n_valid
constexpr unsigned kSpinLimit = 1u << 22;  // ~seconds; then give up instead of hanging the GPU
constexpr float kZEps = 1e-3f;           // Camera.eps, wrappers.py:225

struct LmParams {
  PtkLmProblem p;
  PtkLmResult r;
  int G;         // CTAs per problem
  int n_groups;  // problems in flight
  float a2;      // loss scale squared
  float* partials;
  unsigned int* counters;
  int* error;
};

// One channel of the 12-texel cross footprint: bilinear value and the two
// central differences (interpolation.py:61-83), folded into the 6 sums.
__device__ __forceinline__ void fold_channel(float m0, float m1,                      // row y0-1: x0, x0+1
                                             float a_, float a0, float a1, float a2,  // row y0  : x0-1 .. x0+2
                                             float b_, float b0, float b1, float b2,  // row y0+1
                                             float p0, float p1,                      // row y0+2: x0, x0+1
                                             float ref, float wa, float wb, float wc, float wd,
                                             float& sxx, float& sxy, float& syy, float& sxr, float& syr,
                                             float& srr) {
  const float f = a0 * wa + a1 * wb + b0 * wc + b1 * wd;
  const float fxp = a1 * wa + a2 * wb + b1 * wc + b2 * wd;
  const float fxm = a_ * wa + a0 * wb + b_ * wc + b0 * wd;
  const float fyp = b0 * wa + b1 * wb + p0 * wc + p1 * wd;
  const float fym = m0 * wa + m1 * wb + a0 * wc + a1 * wd;
  const float gx = (fxp - fxm) * 0.5f;
  const float gy = (fyp - fym) * 0.5f;
  const float r = f - ref;
  sxx = fmaf(gx, gx, sxx);
  sxy = fmaf(gx, gy, sxy);
  syy = fmaf(gy, gy, syy);
  sxr = fmaf(gx, r, sxr);
  syr = fmaf(gy, r, syr);
  srr = fmaf(r, r, srr);
}

struct SolveOut {
  int stop;
  int failed;
};

// Damped solve + pose update + stop test; executed by one lane per CTA.
// optimization.py:13-47 (damping, masking, Cholesky), :62-76 (so3exp_map),
// wrappers.py:171-175 (compose), :208-218 (magnitude), pixtrack_optimizer.py:6-18.
__device__ __noinline__ void solve_and_update(const float* tot, const float* lam, float* T, int failed_in, const LmParams& P,
                                 float* logrec, SolveOut& out) {
  float H[6][6], g[6];
  {
    int e = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) {
        H[r][c] = tot[e];
        H[c][r] = tot[e];
        ++e;
      }
#pragma unroll
    for (int k = 0; k < 6; ++k) g[k] = tot[21 + k];
  }
  const float cost_sum = tot[27];
  const float n_valid = tot[28];
  int failed = failed_in | (n_valid < (float)P.p.min_valid);   // learned_optimizer.py:65
  float gnorm = 0.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) gnorm = fmaf(g[k], g[k], gnorm);
  gnorm = sqrtf(gnorm);

  // damped, masked system
  float L[6][6], rhs[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
#pragma unroll
    for (int c = 0; c < 6; ++c) L[r][c] = failed ? (r == c ? 1.f : 0.f) : H[r][c];
    if (!failed) L[r][r] = H[r][r] + fmaxf(H[r][r] * lam[r], 1e-6f);
    rhs[r] = failed ? 0.f : g[r];
  }
  int notpd = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float s = L[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    if (!(s > 0.f)) notpd = 1;
    const float d = sqrtf(s);
    L[j][j] = d;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      float v = L[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v / d;
    }
  }
  float delta[6];
  {
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float v = rhs[i];
#pragma unroll
      for (int k = 0; k < i; ++k) v -= L[i][k] * y[k];
      y[i] = v / L[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      float v = y[i];
#pragma unroll
      for (int k = i + 1; k < 6; ++k) v -= L[k][i] * delta[k];
      delta[i] = v / L[i][i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) delta[i] = -delta[i];
  }
  if (notpd) {  // the reference raises here -> success=False upstream; keep the pose, flag failure
    failed = 1;
#pragma unroll
    for (int i = 0; i < 6; ++i) delta[i] = 0.f;
  }

  // T_delta = (exp(dw), dt)
  const float wx = delta[3], wy = delta[4], wz = delta[5];
  const float theta = sqrtf(wx * wx + wy * wy + wz * wz);
  const bool small = theta < 1e-7f;
  const float div = small ? 1.f : theta;
  const float kx = wx / div, ky = wy / div, kz = wz / div;
  const float Wm[3][3] = {{0.f, -kz, ky}, {kz, 0.f, -kx}, {-ky, kx, 0.f}};
  float Rd[3][3];
  const float s = sinf(theta), omc = 1.f - cosf(theta);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float w2 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) w2 = fmaf(Wm[r][k], Wm[k][c], w2);
      const float res = small ? Wm[r][c] : (Wm[r][c] * s + w2 * omc);
      Rd[r][c] = (r == c ? 1.f : 0.f) + res;
    }
  float Tn[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) Tn[r * 3 + c] = Rd[r][0] * T[c] + Rd[r][1] * T[3 + c] + Rd[r][2] * T[6 + c];
    Tn[9 + r] = delta[r] + (Rd[r][0] * T[9] + Rd[r][1] * T[10] + Rd[r][2] * T[11]);
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = Tn[i];

  const float trace = Rd[0][0] + Rd[1][1] + Rd[2][2];
  const float cs = fminf(fmaxf((trace - 1.f) * 0.5f, -1.f), 1.f);
  const float dR = fabsf(acosf(cs)) / 3.14159265358979323846f * 180.f;
  const float dt = sqrtf(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  const bool small_step = (dt < P.p.dt_stop) && (dR < P.p.dR_stop);
  const bool small_grad = gnorm < P.p.grad_stop;
  out.stop = (small_step || small_grad || notpd) ? 1 : 0;
  out.failed = failed;

  if (logrec != nullptr) {
    logrec[0] = cost_sum;
    logrec[1] = n_valid;
#pragma unroll
    for (int i = 0; i < 12; ++i) logrec[2 + i] = Tn[i];
    logrec[14] = dt;
    logrec[15] = dR;
    logrec[16] = gnorm;
#pragma unroll
    for (int i = 0; i < 6; ++i) logrec[17 + i] = g[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) logrec[23 + i] = delta[i];
    logrec[29] = (float)out.stop;
    logrec[30] = (float)failed;
    logrec[31] = (float)notpd;
#pragma unroll
    for (int i = 0; i < 21; ++i) logrec[32 + i] = tot[i];
  }
}

// Per-problem constants of the camera model, hoisted out of the point loops.
struct CamK {
  float cw, ch, fx, fy, cx, cy, k1, k2, p1, p2;
  float limit, xmax, ymax, padf;
  int n_cam;
  bool limited;
};

// Projection of one 3D point (wrappers.py:177-185,308-355; utils.py:36-69) and, when kJac, the 2x6
// chain A = diag(f) * J_undist * J_project * [I | -skew(p_cam)] (wrappers.py:195-203,316-362; utils.py:72-95).
template <bool kJac>
__device__ __forceinline__ bool point_geometry(const CamK& c, const float* __restrict__ T, float X, float Y, float Z,
                                               float& u, float& v, float (&A0)[6], float (&A1)[6]) {
  const float px = fmaf(T[2], Z, fmaf(T[1], Y, T[0] * X)) + T[9];
  const float py = fmaf(T[5], Z, fmaf(T[4], Y, T[3] * X)) + T[10];
  const float pz = fmaf(T[8], Z, fmaf(T[7], Y, T[6] * X)) + T[11];
  bool valid = pz > kZEps;                               // wrappers.py:311
  const float z = fmaxf(pz, kZEps);
  const float xn = px / z, yn = py / z;
  float xd = xn, yd = yn;
  float jxx = 1.f, jyy = 1.f, jxy = 0.f, jyx = 0.f;
  if (c.n_cam > 6) {
    const float r2 = xn * xn + yn * yn;
    const float radial = c.k1 * r2 + c.k2 * r2 * r2;
    xd = xn + xn * radial;
    yd = yn + yn * radial;
    valid = valid && (!c.limited || r2 < c.limit);
    const float uvn = xn * yn;
    if (kJac) {
      const float drad = 2.f * c.k1 + 4.f * c.k2 * r2;
      jxx += radial + xn * xn * drad;
      jyy += radial + yn * yn * drad;
      jxy += uvn * drad;
      jyx += uvn * drad;
    }
    if (c.n_cam > 8) {
      xd += 2.f * c.p1 * uvn + c.p2 * (r2 + 2.f * xn * xn);
      yd += 2.f * c.p2 * uvn + c.p1 * (r2 + 2.f * yn * yn);
      if (kJac) {
        jxx += 2.f * c.p1 * yn + 6.f * c.p2 * xn;
        jyy += 2.f * c.p2 * xn + 6.f * c.p1 * yn;
        jxy += 2.f * c.p1 * xn + 2.f * c.p2 * yn;
        jyx += 2.f * c.p2 * yn + 2.f * c.p1 * xn;
      }
    }
  }
  u = xd * c.fx + c.cx;
  v = yd * c.fy + c.cy;
  valid = valid && (u >= 0.f) && (v >= 0.f) && (u <= c.cw - 1.f) && (v <= c.ch - 1.f);     // Camera.in_image
  valid = valid && (u >= c.padf) && (v >= c.padf) && (u <= c.xmax) && (v <= c.ymax);       // mask_in_image
  if (kJac) {
    const float iz = 1.f / z;
    const float jp02 = -px / (z * z), jp12 = -py / (z * z);
    const float m00 = c.fx * jxx, m01 = c.fx * jxy, m10 = c.fy * jyx, m11 = c.fy * jyy;
    const float q00 = m00 * iz, q01 = m01 * iz, q02 = m00 * jp02 + m01 * jp12;
    const float q10 = m10 * iz, q11 = m11 * iz, q12 = m10 * jp02 + m11 * jp12;
    A0[0] = q00; A0[1] = q01; A0[2] = q02;
    A0[3] = q02 * py - q01 * pz;
    A0[4] = q00 * pz - q02 * px;
    A0[5] = q01 * px - q00 * py;
    A1[0] = q10; A1[1] = q11; A1[2] = q12;
    A1[3] = q12 * py - q11 * pz;
    A1[4] = q10 * pz - q12 * px;
    A1[5] = q11 * px - q10 * py;
  }
  return valid;
}

constexpr int kPass = 1024;                   // points a CTA stages per pass
constexpr int kRounds = kPass / kThreads;     // points per thread per pass

// Cross-lane sum of 8 per-lane values inside groups of LPP lanes by recursive halving: each xor
// step halves the number of values a lane is responsible for (lanes with the mask bit clear keep
// the lower half, the others the upper half), so 8 values cost 4+2+1(+1+1) shuffles instead of
// 8 x log2(LPP).  On return vals[0 .. n-1] (n = max(1, 8*2/LPP... see kKeep) hold the group totals
// of value indices first .. first+n-1.
template <int LPP>
struct GroupReduce {
  static constexpr int kSteps = (LPP == 32) ? 5 : (LPP == 16) ? 4 : (LPP == 8) ? 3 : 2;
  static constexpr int kHalvings = kSteps < 3 ? kSteps : 3;
  static constexpr int kKeep = 8 >> kHalvings;   // values left per lane
  __device__ static __forceinline__ int run(float (&vals)[8], int sub) {
    const unsigned full = 0xffffffffu;
    int first = 0;
    int n = 8;
    int m = LPP >> 1;
#pragma unroll
    for (int s = 0; s < kSteps; ++s) {
      if (s < kHalvings) {
        const bool up = (sub & m) != 0;
        n >>= 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < n) {
            const float lo = vals[i], hi = vals[i + n];
            const float send = up ? lo : hi;
            const float keep = up ? hi : lo;
            vals[i] = keep + __shfl_xor_sync(full, send, m);
          }
        }
        first += up ? n : 0;
      } else {
        vals[0] += __shfl_xor_sync(full, vals[0], m);
      }
      m >>= 1;
    }
    return first;
  }
};

// One LM launch.  Per iteration and per pass of <= kPass points a CTA runs three phases:
//  A  one THREAD per point: projection + validity, (u,v) to shared memory, deterministic compaction
//     of the valid points into a list;
//  B  WARPS walk the list: LPP lanes per point gather the 12-texel footprint (float4 per lane per
//     texel), fold it into 6 channel sums + the confidence sample, reduce them inside the lane
//     group and park the 7 totals in shared memory;
//  C  one THREAD per point again: robust weight, Jacobian chain, and the 29 contributions
//     (21 H, 6 g, cost, count), warp-reduced and added to the warp's shared-memory accumulator.
// Only phase B touches the maps, and it carries no per-point scalar math and no state of the other
// phases in registers, so its issue slots go to loads and the interpolation FMAs.
// kFast: C == 4*LPP (one float4 per lane per texel, compile-time texel stride) and pad >= 1 (all
// 12 texels in bounds except the two that carry an exactly-zero weight, which are clamped).
// kCtas: resident CTAs per SM the register allocation is bounded for (1: ~116 registers, 2: 64 -- twice the warps to
// hide the L2 latency of the gathers behind).
template <int LPP, bool kFast, int kCtas>
__global__ void __launch_bounds__(kThreads, kCtas) lm_kernel(const LmParams P) {
  constexpr int PPW = 32 / LPP;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane % LPP;
  const int grp = lane / LPP;
  const int group = blockIdx.x / P.G;
  const int rank = blockIdx.x - group * P.G;
  const unsigned full = 0xffffffffu;

  __shared__ float2 sUV[kPass];
  __shared__ __align__(16) float sSums[kPass][8];
  __shared__ unsigned short sList[kPass];
  __shared__ unsigned sMask[kRounds * kWarps];
  __shared__ int sOffs[kRounds * kWarps];
  __shared__ int sNValid;
  __shared__ float sWarp[kWarps][32];
  __shared__ float sTot[32];
  __shared__ float sT[12];
  __shared__ float sCam[12];
  __shared__ float sLam[6];
  __shared__ CamK sCK;
  __shared__ int sStop, sFailed, sAbort;
  static_assert(kRounds * kWarps <= 32, "prefix scan is done by one warp");

  const PtkLmProblem& p = P.p;
  const int N = p.N;
  const int per = (N + P.G - 1) / P.G;
  const int start = min(N, rank * per);
  const int end = min(N, start + per);
  unsigned bar_count = 0;  // barriers this group has completed (identical in all its CTAs)
  if (threadIdx.x == 0) sAbort = 0;

  for (int b = group; b < p.B && group < P.n_groups; b += P.n_groups) {
    const float* p3d = p.p3d + (size_t)b * p.p3d_bstride;
    const float* wref = p.w_ref ? p.w_ref + (size_t)b * p.w_ref_bstride : nullptr;
    const uint8_t* mask = p.mask ? p.mask + (size_t)b * p.mask_bstride : nullptr;

    __syncthreads();  // previous problem fully retired before smem is rewritten
    if (threadIdx.x < 12) sT[threadIdx.x] = p.T_init[(size_t)b * p.T_bstride + threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 32 + 12) {
      const int i = threadIdx.x - 32;
      sCam[i] = (i < p.n_cam) ? p.cam[(size_t)b * p.cam_bstride + i] : 0.f;
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 6) sLam[threadIdx.x - 64] = p.lambda[(size_t)b * p.lambda_bstride + threadIdx.x - 64];
    if (threadIdx.x == 96) {
      sStop = 0;
      sFailed = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      CamK ck;
      ck.cw = sCam[0]; ck.ch = sCam[1]; ck.fx = sCam[2]; ck.fy = sCam[3]; ck.cx = sCam[4]; ck.cy = sCam[5];
      ck.k1 = sCam[6]; ck.k2 = sCam[7]; ck.p1 = sCam[8]; ck.p2 = sCam[9];
      ck.n_cam = p.n_cam;
      ck.limited = false;
      ck.limit = 0.f;
      if (p.n_cam > 6) {   // validity limit of the radial model, utils.py:49-60
        const float disc = 9.f * ck.k1 * ck.k1 - 20.f * ck.k2;
        ck.limited = ((ck.k2 > 0.f) && (disc > 0.f)) || ((ck.k2 <= 0.f) && (ck.k1 > 0.f));
        ck.limit = fabsf(ck.k2 > 0.f ? (sqrtf(disc) - 3.f * ck.k1) / (10.f * ck.k2) : 1.f / (3.f * ck.k1));
      }
      ck.xmax = (float)(p.W - 1 - p.pad);
      ck.ymax = (float)(p.H - 1 - p.pad);
      ck.padf = (float)p.pad;
      sCK = ck;
    }
    __syncthreads();

    const bool skipped = p.skip != nullptr && p.skip[b] != 0;
    int it = 0;
    if (!skipped) {
      for (it = 0; it < p.num_iters; ++it) {
        sWarp[warp][lane] = 0.f;   // this warp's accumulator row (only this warp touches it until the CTA sum)

        for (int pass0 = start; pass0 < end; pass0 += kPass) {
          const int cnt = min(kPass, end - pass0);
          // ---- phase A: projection, validity, compaction ------------------------------------
          unsigned myvalid = 0;
#pragma unroll
          for (int r = 0; r < kRounds; ++r) {
            const int i = r * kThreads + threadIdx.x;
            bool ok = false;
            if (i < cnt) {
              const int pt = pass0 + i;
              float u, v, d0[6], d1[6];
              ok = point_geometry<false>(sCK, sT, p3d[3 * pt], p3d[3 * pt + 1], p3d[3 * pt + 2], u, v, d0, d1);
              if (mask != nullptr) ok = ok && (mask[pt] != 0);
              if (ok) sUV[i] = make_float2(u, v);
            }
            const unsigned m = __ballot_sync(full, ok);
            if (lane == 0) sMask[r * kWarps + warp] = m;
            myvalid |= (ok ? 1u : 0u) << r;
          }
          __syncthreads();
          if (warp == 0) {
            const int c = (lane < kRounds * kWarps) ? __popc(sMask[lane]) : 0;
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int o = __shfl_up_sync(full, incl, d);
              if (lane >= d) incl += o;
            }
            if (lane < kRounds * kWarps) sOffs[lane] = incl - c;
            if (lane == 31) sNValid = incl;
          }
          __syncthreads();
#pragma unroll
          for (int r = 0; r < kRounds; ++r) {
            if ((myvalid >> r) & 1u) {
              const int pos = sOffs[r * kWarps + warp] + __popc(sMask[r * kWarps + warp] & ((1u << lane) - 1u));
              sList[pos] = (unsigned short)(r * kThreads + threadIdx.x);
            }
          }
          __syncthreads();
          // ---- phase B: gather + fold + reduce, LPP lanes per point ---------------------------
          {
            const int nv = sNValid;
            const int W = p.W, H = p.H;
            const int C4 = kFast ? LPP : (p.C >> 2);
            const float4* fq4 = reinterpret_cast<const float4*>(p.fq + (size_t)b * p.fq_bstride);
            const float4* fref4 = reinterpret_cast<const float4*>(p.f_ref + (size_t)b * p.f_ref_bstride);
            const float* wq = p.wq ? p.wq + (size_t)b * p.wq_bstride : nullptr;
            for (int kb = warp * PPW; kb < nv; kb += kWarps * PPW) {
              const int k = kb + grp;
              const bool active = k < nv;
              float s8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) s8[i] = 0.f;
              int li = 0;
              if (active) {
                li = sList[k];
                const float2 uv = sUV[li];
                const int pt = pass0 + li;
                const float x0f = floorf(uv.x), y0f = floorf(uv.y);
                const int x0 = (int)x0f, y0 = (int)y0f;
                const float ax = uv.x - x0f, ay = uv.y - y0f;
                const float wa = (1.f - ax) * (1.f - ay), wb = ax * (1.f - ay), wc = (1.f - ax) * ay, wd = ax * ay;
                if (kFast) {
                  // pad >= 1: x0-1, x0+1, y0-1, y0+1 are in range; x0+2 / y0+2 leave the map only when
                  // ax / ay is exactly 0, i.e. with zero weight: clamp them onto a valid texel.
                  constexpr int dx = LPP;   // float4 units between x-neighbours
                  const int dy = W * LPP;
                  const float4* r0 = fq4 + ((size_t)y0 * W + x0) * LPP + sub;
                  const float4* rm = r0 - dy;
                  const float4* r1 = r0 + dy;
                  const float4* r2 = (y0 + 2 < H) ? r1 + dy : r1;
                  const int x2 = (x0 + 2 < W) ? 2 * dx : dx;
                  const float4 m0 = __ldg(rm), m1 = __ldg(rm + dx);
                  const float4 a_ = __ldg(r0 - dx), a0 = __ldg(r0), a1 = __ldg(r0 + dx), a2 = __ldg(r0 + x2);
                  const float4 b_ = __ldg(r1 - dx), b0 = __ldg(r1), b1 = __ldg(r1 + dx), b2 = __ldg(r1 + x2);
                  const float4 q0 = __ldg(r2), q1 = __ldg(r2 + dx);
                  const float4 rf = __ldg(fref4 + (size_t)pt * LPP + sub);
                  fold_channel(m0.x, m1.x, a_.x, a0.x, a1.x, a2.x, b_.x, b0.x, b1.x, b2.x, q0.x, q1.x, rf.x, wa, wb, wc,
                               wd, s8[0], s8[1], s8[2], s8[3], s8[4], s8[5]);
                  fold_channel(m0.y, m1.y, a_.y, a0.y, a1.y, a2.y, b_.y, b0.y, b1.y, b2.y, q0.y, q1.y, rf.y, wa, wb, wc,
                               wd, s8[0], s8[1], s8[2], s8[3], s8[4], s8[5]);
                  fold_channel(m0.z, m1.z, a_.z, a0.z, a1.z, a2.z, b_.z, b0.z, b1.z, b2.z, q0.z, q1.z, rf.z, wa, wb, wc,
                               wd, s8[0], s8[1], s8[2], s8[3], s8[4], s8[5]);
                  fold_channel(m0.w, m1.w, a_.w, a0.w, a1.w, a2.w, b_.w, b0.w, b1.w, b2.w, q0.w, q1.w, rf.w, wa, wb, wc,
                               wd, s8[0], s8[1], s8[2], s8[3], s8[4], s8[5]);
                  if (wq != nullptr && sub < 4) {   // confidence: 4 texels, one per lane (costs.py:27-32)
                    const float wgt = (sub == 0) ? wa : (sub == 1) ? wb : (sub == 2) ? wc : wd;
                    s8[6] = wgt * __ldg(wq + (size_t)(y0 + (sub >> 1)) * W + x0 + (sub & 1));
                  }
                } else {
                  const bool xin_ = x0 - 1 >= 0, xin1 = x0 + 1 < W, xin2 = x0 + 2 < W;
                  const bool yin_ = y0 - 1 >= 0, yin1 = y0 + 1 < H, yin2 = y0 + 2 < H;
                  const size_t row0 = (size_t)y0 * W;
                  for (int c4 = sub; c4 < C4; c4 += LPP) {
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4* base0 = fq4 + (row0 + x0) * C4 + c4;   // texel (y0, x0)
                    const ptrdiff_t dx = C4, dy = (ptrdiff_t)W * C4;
                    const float4 m0 = yin_ ? __ldg(base0 - dy) : z4;
                    const float4 m1 = (yin_ && xin1) ? __ldg(base0 - dy + dx) : z4;
                    const float4 a_ = xin_ ? __ldg(base0 - dx) : z4;
                    const float4 a0 = __ldg(base0);
                    const float4 a1 = xin1 ? __ldg(base0 + dx) : z4;
                    const float4 a2 = xin2 ? __ldg(base0 + 2 * dx) : z4;
                    const float4 b_ = (yin1 && xin_) ? __ldg(base0 + dy - dx) : z4;
                    const float4 b0 = yin1 ? __ldg(base0 + dy) : z4;
                    const float4 b1 = (yin1 && xin1) ? __ldg(base0 + dy + dx) : z4;
                    const float4 b2 = (yin1 && xin2) ? __ldg(base0 + dy + 2 * dx) : z4;
                    const float4 q0 = yin2 ? __ldg(base0 + 2 * dy) : z4;
                    const float4 q1 = (yin2 && xin1) ? __ldg(base0 + 2 * dy + dx) : z4;
                    const float4 rf = __ldg(fref4 + (size_t)pt * C4 + c4);
                    fold_channel(m0.x, m1.x, a_.x, a0.x, a1.x, a2.x, b_.x, b0.x, b1.x, b2.x, q0.x, q1.x, rf.x, wa, wb,
                                 wc, wd, s8[0], s8[1], s8[2], s8[3], s8[4], s8[5]);
                    fold_channel(m0.y, m1.y, a_.y, a0.y, a1.y, a2.y, b_.y, b0.y, b1.y, b2.y, q0.y, q1.y, rf.y, wa, wb,
                                 wc, wd, s8[0], s8[1], s8[2], s8[3], s8[4], s8[5]);
                    fold_channel(m0.z, m1.z, a_.z, a0.z, a1.z, a2.z, b_.z, b0.z, b1.z, b2.z, q0.z, q1.z, rf.z, wa, wb,
                                 wc, wd, s8[0], s8[1], s8[2], s8[3], s8[4], s8[5]);
                    fold_channel(m0.w, m1.w, a_.w, a0.w, a1.w, a2.w, b_.w, b0.w, b1.w, b2.w, q0.w, q1.w, rf.w, wa, wb,
                                 wc, wd, s8[0], s8[1], s8[2], s8[3], s8[4], s8[5]);
                  }
                  if (wq != nullptr && sub < 4) {
                    const int ox = sub & 1, oy = sub >> 1;
                    const bool in = (ox == 0 || xin1) && (oy == 0 || yin1);
                    const float wgt = (sub == 0) ? wa : (sub == 1) ? wb : (sub == 2) ? wc : wd;
                    s8[6] = in ? wgt * __ldg(wq + row0 + (size_t)oy * W + x0 + ox) : 0.f;
                  }
                }
              }
              const int first = GroupReduce<LPP>::run(s8, sub);
              constexpr int kKeep = GroupReduce<LPP>::kKeep;
              constexpr int kDup = (LPP * kKeep) / 8;            // lanes holding the same value index
              if (active && (sub % kDup) == 0) {
#pragma unroll
                for (int i = 0; i < kKeep; ++i) sSums[li][first + i] = s8[i];
              }
            }
          }
          __syncthreads();
          // ---- phase C: weight, Jacobian chain, 29 contributions, one thread per point ---------
          {
            float acc[kEntries];
#pragma unroll
            for (int e = 0; e < kEntries; ++e) acc[e] = 0.f;
#pragma unroll
            for (int r = 0; r < kRounds; ++r) {
              if ((myvalid >> r) & 1u) {
                const int i = r * kThreads + threadIdx.x;
                const int pt = pass0 + i;
                float u, v, A0[6], A1[6];
                point_geometry<true>(sCK, sT, p3d[3 * pt], p3d[3 * pt + 1], p3d[3 * pt + 2], u, v, A0, A1);
                const float4 s0 = *reinterpret_cast<const float4*>(&sSums[i][0]);
                const float4 s1 = *reinterpret_cast<const float4*>(&sSums[i][4]);
                const float sxx = s0.x, sxy = s0.y, syy = s0.z, sxr = s0.w, syr = s1.x, srr = s1.y, cq = s1.z;
                const float x = srr / P.a2;                              // losses.py:17-19
                const float wl = 2.f / (x + 2.f);                        // losses.py:73
                const float loss = 2.f * log1pf(fminf(0.5f * x, 33e37f)) * P.a2;
                float w = wl;
                if (wref != nullptr) w *= __ldg(wref + pt) * (p.wq != nullptr ? cq : 1.f);
                float Xv[6], Yv[6];
#pragma unroll
                for (int l = 0; l < 6; ++l) {
                  Xv[l] = sxx * A0[l] + sxy * A1[l];
                  Yv[l] = sxy * A0[l] + syy * A1[l];
                }
                int e = 0;
#pragma unroll
                for (int rr = 0; rr < 6; ++rr)
#pragma unroll
                  for (int cc = rr; cc < 6; ++cc) {
                    acc[e] += w * (A0[rr] * Xv[cc] + A1[rr] * Yv[cc]);
                    ++e;
                  }
#pragma unroll
                for (int l = 0; l < 6; ++l) acc[21 + l] += w * (A0[l] * sxr + A1[l] * syr);
                acc[27] += loss;
                acc[28] += 1.f;
              }
            }
            // thread -> warp (fixed order), added to this warp's row
            float mine = 0.f;
#pragma unroll
            for (int e = 0; e < kEntries; ++e) {
              float v = acc[e];
#pragma unroll
              for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(full, v, m);
              if (lane == e) mine = v;
            }
            sWarp[warp][lane] += mine;
          }
        }  // passes

        __syncthreads();
        if (warp == 0) {
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) tot += sWarp[w][lane];
          if (P.G > 1) {
            const int parity = bar_count & 1u;
            float* slot = P.partials + ((size_t)(group * P.G) * 2) * 32;
            slot[((size_t)rank * 2 + parity) * 32 + lane] = tot;
            __threadfence();
            __syncwarp();
            if (lane == 0) {
              atomicAdd(&P.counters[group], 1u);
              const unsigned target = (bar_count + 1u) * (unsigned)P.G;
              unsigned spins = 0;
              volatile unsigned int* ctr = &P.counters[group];
              while (*ctr < target) {
                if (++spins > kSpinLimit) {
                  atomicExch(P.error, 1);
                  sAbort = 1;
                  break;
                }
              }
              __threadfence();
            }
            __syncwarp();
            tot = 0.f;
            for (int r = 0; r < P.G; ++r) tot += __ldcg(slot + ((size_t)r * 2 + parity) * 32 + lane);
          }
          sTot[lane] = tot;
          __syncwarp();
          if (lane == 0) {
            SolveOut so;
            float* logrec = (P.r.log != nullptr && rank == 0)
                                ? P.r.log + ((size_t)b * p.num_iters + it) * PTK_LOG_STRIDE
                                : nullptr;
            solve_and_update(sTot, sLam, sT, sFailed, P, logrec, so);
            sStop = so.stop;
            sFailed = so.failed;
          }
        }
        ++bar_count;
        __syncthreads();
        if (sAbort) break;
        if (sStop) {
          ++it;
          break;
        }
      }  // iterations
    }
    if (rank == 0 && threadIdx.x < 12) P.r.T[(size_t)b * 12 + threadIdx.x] = sT[threadIdx.x];
    if (rank == 0 && threadIdx.x == 32) {
      // a timed-out barrier (sAbort) leaves inconsistent sums behind: the problem is reported as failed
      P.r.failed[b] = (uint8_t)((sFailed != 0) || skipped || (sAbort != 0));
      P.r.n_iters[b] = skipped ? 0 : min(it, p.num_iters);
    }
    if (sAbort) {   // ... and so is every problem this group will no longer visit
      if (rank == 0 && threadIdx.x == 32)
        for (int bb = b + P.n_groups; bb < p.B; bb += P.n_groups) {
          P.r.failed[bb] = 1;
          P.r.n_iters[bb] = 0;
        }
      break;
    }
  }

  // self-cleaning workspace: the last CTA to leave zeroes the counters
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int* exitc = &P.counters[PTK_MAX_SMS];
    const unsigned prev = atomicAdd(exitc, 1u);
    if (prev == gridDim.x - 1) {
      for (int gI = 0; gI < P.n_groups; ++gI) P.counters[gI] = 0u;
      *exitc = 0u;
      __threadfence();
    }
  }
}

int pick_lpp(int C) {
  const int c4 = C / 4;
  if (c4 <= 4) return 4;
  if (c4 <= 8) return 8;
  if (c4 <= 16) return 16;
  return 32;
}

int ctas_per_sm() {   // PTK_LM_CTAS = 1 | 2 (default 2)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PTK_LM_CTAS");
    v = (e != nullptr && atoi(e) == 1) ? 1 : 2;
  }
  return v;
}

void plan(const PtkContext* ctx, const PtkLmProblem* p, int* G, int* n_groups) {
  const int sms = ctx->num_sms * ctas_per_sm();   // co-resident CTA slots of the cooperative launch
  const int lpp = pick_lpp(p->C);
  const int pts_per_pass = kWarps * (32 / lpp);
  int ng = p->B < ctx->num_sms ? p->B : ctx->num_sms;
  if (ng < 1) ng = 1;
  int g = sms / ng;
  // do not spread a problem so thin that a CTA has under two passes of work
  int want = (p->N + 2 * pts_per_pass - 1) / (2 * pts_per_pass);
  if (want < 1) want = 1;
  if (g > want) g = want;
  if (g < 1) g = 1;
  const char* env = getenv("PTK_LM_CTAS_PER_PROBLEM");
  if (env != nullptr) {
    const int v = atoi(env);
    if (v >= 1 && v * ng <= sms) g = v;
  }
  *G = g;
  *n_groups = ng;
}

}  // namespace

extern "C" int ptk_lm_plan(const PtkContext* ctx, const PtkLmProblem* prob, int32_t* ctas_per_problem,
                           int32_t* n_groups) {
  PTK_REQUIRE(ctx && prob && ctas_per_problem && n_groups, "null argument");
  int G, ng;
  plan(ctx, prob, &G, &ng);
  *ctas_per_problem = G;
  *n_groups = ng;
  return PTK_OK;
}

extern "C" int64_t ptk_lm_workspace_bytes(void) { return (int64_t)PTK_LM_WS_BYTES; }

extern "C" int ptk_lm_run(PtkContext* ctx, const PtkLmProblem* prob, const PtkLmResult* res, void* stream) {
  PTK_REQUIRE(ctx && prob && res, "null context/problem/result");
  const PtkLmProblem& p = *prob;
  PTK_REQUIRE(p.B >= 1 && p.N >= 0, "B >= 1 and N >= 0 required");
  PTK_REQUIRE(p.C >= 4 && p.C % 4 == 0 && p.C <= 512, "C must be a multiple of 4 in [4, 512]");
  PTK_REQUIRE(p.H >= 2 && p.W >= 2, "map must be at least 2x2");
  PTK_REQUIRE(p.n_cam == 6 || p.n_cam == 8 || p.n_cam == 10, "n_cam must be 6, 8 or 10");
  PTK_REQUIRE(p.num_iters >= 0 && p.pad >= 0, "num_iters and pad must be >= 0");
  // N == 0 (every point dropped upstream) is legal: the per-point arrays may then be null; the launch fails the
  // problems at the first iteration like learned_optimizer.py:65 does with an empty p3D
  PTK_REQUIRE(p.fq && p.cam && p.T_init && p.lambda, "null input pointer");
  PTK_REQUIRE(p.N == 0 || (p.p3d && p.f_ref), "null point / descriptor pointer");
  PTK_REQUIRE(p.N == 0 || (p.w_ref == nullptr) == (p.wq == nullptr), "w_ref and wq must be given together");
  PTK_REQUIRE(res->T && res->failed && res->n_iters, "null output pointer");
  PTK_REQUIRE(((uintptr_t)p.fq % 16 == 0) && ((uintptr_t)p.f_ref % 16 == 0), "fq and f_ref must be 16-byte aligned");
  PTK_REQUIRE((p.fq_bstride % 4 == 0) && (p.f_ref_bstride % 4 == 0), "batch strides of fq/f_ref must be multiples of 4");
  PTK_REQUIRE(p.loss_scale > 0.f, "loss_scale must be positive");

  LmParams P;
  P.p = p;
  P.r = *res;
  plan(ctx, prob, &P.G, &P.n_groups);
  P.a2 = (float)((double)p.loss_scale * (double)p.loss_scale);
  // 0.1f squared in double rounds one ulp above float(0.01); the reference divides by float(0.1**2)
  if (p.loss_scale == 0.1f) P.a2 = 0.01f;
  if (p.workspace != nullptr) {   // per-launch workspace: concurrent launches on one context are independent
    PTK_REQUIRE(p.workspace_bytes >= (int64_t)PTK_LM_WS_BYTES, "workspace smaller than ptk_lm_workspace_bytes()");
    PTK_REQUIRE((uintptr_t)p.workspace % 16 == 0, "workspace must be 16-byte aligned");
    P.partials = (float*)p.workspace;
    P.counters = (unsigned int*)((char*)p.workspace + PTK_LM_WS_PARTIAL_BYTES);
  } else {
    P.partials = ctx->lm_partials;
    P.counters = ctx->lm_counters;
  }
  P.error = ctx->lm_error;
  PtkDeviceGuard guard(ctx->device);

  const dim3 grid(P.G * P.n_groups), block(kThreads);
  void* args[] = {&P};
  const void* fn = nullptr;
  const int lpp = pick_lpp(p.C);
  const bool fast = (p.C == 4 * lpp) && (p.pad >= 1);
  if (ctas_per_sm() == 2) {
    switch (lpp) {
      case 4: fn = fast ? (const void*)lm_kernel<4, true, 2> : (const void*)lm_kernel<4, false, 2>; break;
      case 8: fn = fast ? (const void*)lm_kernel<8, true, 2> : (const void*)lm_kernel<8, false, 2>; break;
      case 16: fn = fast ? (const void*)lm_kernel<16, true, 2> : (const void*)lm_kernel<16, false, 2>; break;
      default: fn = fast ? (const void*)lm_kernel<32, true, 2> : (const void*)lm_kernel<32, false, 2>; break;
    }
  } else {
    switch (lpp) {
      case 4: fn = fast ? (const void*)lm_kernel<4, true, 1> : (const void*)lm_kernel<4, false, 1>; break;
      case 8: fn = fast ? (const void*)lm_kernel<8, true, 1> : (const void*)lm_kernel<8, false, 1>; break;
      case 16: fn = fast ? (const void*)lm_kernel<16, true, 1> : (const void*)lm_kernel<16, false, 1>; break;
      default: fn = fast ? (const void*)lm_kernel<32, true, 1> : (const void*)lm_kernel<32, false, 1>; break;
    }
  }
  if (P.G > 1) {
    PTK_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, grid, block, args, 0, (cudaStream_t)stream));
  } else {
    PTK_CUDA_CHECK(cudaLaunchKernel(fn, grid, block, args, 0, (cudaStream_t)stream));
  }
  return PTK_OK;
}
