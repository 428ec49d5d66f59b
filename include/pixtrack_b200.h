/*
 * pixtrack_b200.h -- C ABI of libpixtrack_b200.so (sm_100a).
 *
 * The reference (GiantAI/pixtrack) has no FFI on this path: its boundary is
 * Python duck typing.  This header is the C-level
 * contract the Python adapters in pixtrack_b200/ bind with ctypes; each entry
 * point names the reference interface it replaces (paths relative to
 * /root/reference).
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer owned by the caller (e.g. a
 *    torch allocation) unless the name says `host_`;
 *  - all calls are stream-ordered on `stream` (a cudaStream_t passed as
 *    void*; NULL = legacy default stream) and never synchronise the device;
 *  - return value 0 = ok, negative = error; `ptk_last_error()` gives the text
 *    of the last error on the calling thread; nothing throws across the ABI;
 *  - a PtkContext holds the device index, the SM count and a small default
 *    LM workspace.  Only ptk_lm_run uses a workspace (the partial sums and
 *    barrier counters of CTAs that share a problem): launches that pass their
 *    own `PtkLmProblem.workspace` may be in flight concurrently on any number
 *    of streams; launches that leave it NULL share the context's and must
 *    then be stream-ordered with respect to each other.  No other entry point
 *    keeps per-context device state, so extractor plans, samplers, renders and
 *    masks of one context may overlap freely (each PtkExtractor / PtkNerf owns
 *    its buffers).  Host-side, calls are thread-compatible, not thread-safe --
 *    like the reference's single Python caller;
 *  - every entry point that allocates or launches makes the context's device
 *    current for the duration of the call and restores the caller's;
 *  - batch strides (`*_bstride`) are in ELEMENTS between consecutive
 *    problems; 0 means "shared by all problems of the batch".
 */
#ifndef PIXTRACK_B200_H_
#define PIXTRACK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTK_ABI_VERSION 2

/* error codes */
#define PTK_OK 0
#define PTK_ERR_INVALID -1   /* bad argument (NULL pointer, unsupported size)  */
#define PTK_ERR_CUDA -2      /* a CUDA runtime/driver call failed               */
#define PTK_ERR_UNSUPPORTED -3

typedef struct PtkContext PtkContext;

int ptk_abi_version(void);
const char* ptk_last_error(void);
/* Creates the per-device context (workspace, SM count).  Fails (PTK_ERR_CUDA)
 * when no sm_100 device is present: there is no CPU fallback. */
int ptk_create(int device, PtkContext** out);
void ptk_destroy(PtkContext* ctx);
int ptk_num_sms(const PtkContext* ctx);
/* SYNCHRONISING health check for tests: PTK_OK unless a device-side abort (a
 * barrier that timed out instead of hanging the GPU) was recorded. */
int ptk_device_status(PtkContext* ctx);

/* ------------------------------------------------------------------------
 * Fused Levenberg-Marquardt pose refinement on one pyramid level.
 *
 * Replaces  PixTrackOptimizer.run / LearnedOptimizer._run
 *   pixloc/pixloc/pixlib/models/learned_optimizer.py:48-95,
 *   pixtrack/optimizers/pixtrack_optimizer.py:6-18 (stop test every iteration)
 * and everything it calls per iteration:
 *   DirectAbsoluteCost.residual_jacobian  pixlib/geometry/costs.py:15-67
 *   Camera.world2image / J_world2image    pixlib/geometry/wrappers.py:308-362
 *   undistort_points / J_undistort_points pixlib/geometry/utils.py:36-95
 *   interpolate_tensor_bilinear           pixlib/geometry/interpolation.py:57-89
 *   scaled_barron(0, c)                   pixlib/geometry/losses.py:8-19,38-82
 *   BaseOptimizer.build_system            pixlib/models/base_optimizer.py:83-92
 *   optimizer_step (damped Cholesky)      pixlib/geometry/optimization.py:13-47
 *   Pose.from_aa / compose, so3exp_map    wrappers.py:127-175, optimization.py:62-76
 * The whole iteration loop, the 6x6 solve, the SE(3) update and the stop test
 * run on the device: no host synchronisation between iterations.
 *
 * B independent problems (reference views / frames) are solved by one launch.
 * ---------------------------------------------------------------------- */
#define PTK_LOG_STRIDE 64
/* per-iteration log record, PTK_LOG_STRIDE floats:
 *  [0] sum over valid points of the robust cost      [1] number of valid points
 *  [2..13] pose AFTER the update (R row-major, t)     [14] |dt|   [15] dR in degrees
 *  [16] |g| (unmasked gradient)   [17..22] g          [23..28] delta (dt, dw)
 *  [29] 1 if the stop test fired  [30] 1 if failed    [31] 1 if the damped H was not
 *  positive definite (the reference raises there)     [32..52] H, upper triangle,
 *  row-major (00 01 .. 05 11 12 .. 55), BEFORE damping.
 * This is what DebugTracker.log_optim_iter (pixtrack/localization/tracker.py:32-46)
 * needs: cost_sum/n_valid, T, |dt|. */

typedef struct PtkLmProblem {
  int32_t B;          /* number of independent problems                          */
  int32_t N;          /* 3D points per problem                                   */
  int32_t C;          /* descriptor channels, multiple of 4, <= 512              */
  int32_t H, W;       /* query map size                                          */
  int32_t n_cam;      /* camera vector length: 6, 8 (k1,k2) or 10 (+p1,p2)       */
  int32_t num_iters;  /* conf.num_iters                                          */
  int32_t pad;        /* conf.interpolation.pad (>= 0)                           */
  int32_t min_valid;  /* fail when fewer valid points (reference: 10)            */
  int32_t reserved0;
  const float* p3d;      int64_t p3d_bstride;    /* [N][3]                        */
  const float* f_ref;    int64_t f_ref_bstride;  /* [N][C]                        */
  const float* w_ref;    int64_t w_ref_bstride;  /* [N] or NULL (no confidences)  */
  const float* fq;       int64_t fq_bstride;     /* [H][W][C]  channels-last      */
  const float* wq;       int64_t wq_bstride;     /* [H][W] or NULL                */
  const uint8_t* mask;   int64_t mask_bstride;   /* [N] or NULL                   */
  const float* cam;      int64_t cam_bstride;    /* [n_cam] w h fx fy cx cy k1 k2 p1 p2 */
  const float* T_init;   int64_t T_bstride;      /* [12] R row-major then t       */
  const float* lambda;   int64_t lambda_bstride; /* [6] damping (DampingNet out)  */
  const uint8_t* skip;                           /* [B] or NULL: nonzero -> problem is
                                                    passed through untouched (used to chain
                                                    levels: skip = previous level's failed) */
  float loss_scale;   /* c of scaled_barron(0, c): 0.1                            */
  float grad_stop;    /* conf.grad_stop_criteria                                  */
  float dt_stop;      /* conf.dt_stop_criteria                                    */
  float dR_stop;      /* conf.dR_stop_criteria (degrees)                          */
  void* workspace;    /* ZERO-INITIALISED device scratch of >= ptk_lm_workspace_bytes()
                         owned by this (prepared) launch, or NULL = the context's shared
                         workspace.  The kernel leaves it zeroed again.              */
  int64_t workspace_bytes;
} PtkLmProblem;

typedef struct PtkLmResult {
  float* T;           /* [B][12]                                                  */
  uint8_t* failed;    /* [B]  (sticky OR with skip[b])                            */
  int32_t* n_iters;   /* [B]  iterations executed                                 */
  float* log;         /* [B][num_iters][PTK_LOG_STRIDE] or NULL                   */
} PtkLmResult;

int ptk_lm_run(PtkContext* ctx, const PtkLmProblem* prob, const PtkLmResult* res, void* stream);
/* bytes of PtkLmProblem.workspace (host only) */
int64_t ptk_lm_workspace_bytes(void);

/* Launch geometry the next ptk_lm_run with this problem would use (for
 * benchmarks / tests): CTAs per problem and number of problem groups. */
int ptk_lm_plan(const PtkContext* ctx, const PtkLmProblem* prob, int32_t* ctas_per_problem, int32_t* n_groups);

/* ------------------------------------------------------------------------
 * Map layout / normalisation helpers.
 *
 * ptk_chw_to_hwc: [C][H][W] -> [H][W][C]; with normalize != 0 each pixel's C
 * vector is L2-normalised on the way (F.normalize(dim=0), eps 1e-12), i.e. the
 * query-side normalisation of BaseRefiner.refine_pose_using_features
 * (pixloc/pixloc/localization/base_refiner.py:92-94).
 * ---------------------------------------------------------------------- */
/* stream-ordered device-to-device copy (tests snapshot plan-owned activations with it) */
int ptk_copy_d2d(void* dst, const void* src, int64_t bytes, void* stream);
int ptk_chw_to_hwc(PtkContext* ctx, const float* src, float* dst, int32_t C, int32_t H, int32_t W,
                   int32_t normalize, void* stream);

/* ------------------------------------------------------------------------
 * Sparse bilinear sampling of a dense map at N pixel positions.
 *
 * Replaces Interpolator.__call__ / interpolate_tensor(mode='linear')
 *   pixloc/pixloc/pixlib/geometry/interpolation.py:57-141
 * as called by PoseTrackerRefiner.interp_sparse_observations
 *   pixtrack/localization/pixloc_pose_refiners.py:349-351.
 * The map is addressed as map[c*stride_c + y*stride_y + x*stride_x] (element
 * strides), so both [C][H][W] and channels-last [H][W][C] storage work.
 * pts [N][2] = (x, y) pixels.  Outputs: vals [N][C]; mask [N] (1 when
 * pad <= p <= size-1-pad), may be NULL; grads [N][C][2] central differences,
 * may be NULL.
 * ---------------------------------------------------------------------- */
int ptk_sample_points(PtkContext* ctx, const float* map, int64_t stride_c, int64_t stride_y, int64_t stride_x,
                      int32_t C, int32_t H, int32_t W, const float* pts, int32_t N, int32_t pad, float* vals,
                      uint8_t* mask, float* grads, void* stream);

/* ------------------------------------------------------------------------
 * Reference-view observations of the model points, all pyramid levels in one
 * launch.
 *
 * Replaces PoseTrackerRefiner.interp_sparse_observations
 *   pixtrack/localization/pixloc_pose_refiners.py:327-368
 * (per level: camera.scale(sc).world2image(T * p3d) in float64, interpolator at
 * the float pixel position, mask & valid; validity = AND over levels) and the
 * reference-side half of BaseRefiner.refine_pose_using_features
 *   pixloc/pixloc/localization/base_refiner.py:74-84
 * (stack per level, split descriptor / confidence, F.normalize(F_ref, dim=1)).
 * Level l reads feat [H][W][C] channels-last (NOT normalised) and conf [H][W],
 * and writes f_out [N][C] (L2-normalised when normalize != 0) and w_out [N].
 * p3d: DEVICE float64 [N][3].  host_cam: HOST float64 [n_cam] camera of the
 * reference image; level l uses Camera.scale((sx, sy)).  host_T: HOST float64
 * [12] world-to-camera pose (R row-major, t).  valid [N]: 1 when the point
 * projects inside every level (the reference drops the others from its lists;
 * pass `valid` as PtkLmProblem.mask for the same sums).
 * ---------------------------------------------------------------------- */
#define PTK_MAX_LEVELS 4
typedef struct PtkRefLevel {
  const float* feat;
  const float* conf;     /* NULL together with w_out: descriptors only            */
  float* f_out;
  float* w_out;
  double sx, sy;
  int32_t C, H, W;
  int32_t normalize;
} PtkRefLevel;

int ptk_sample_reference(PtkContext* ctx, const PtkRefLevel* levels, int32_t n_levels, const double* p3d, int32_t N,
                         const double* host_cam, int32_t n_cam, const double* host_T, int32_t pad, uint8_t* valid,
                         void* stream);

/* ------------------------------------------------------------------------
 * Feature extractor (PixLoc UNet, VGG19 encoder) on the tcgen05 tensor cores.
 *
 * ptk_conv_f16: one convolution layer on channels-last fp16 activations
 *   (3x3 pad 1 when taps == 9, 1x1 when taps == 1), fp32 accumulation, + bias,
 *   optional ReLU, fp16 output [H][W][C_out].  An optional second input continues
 *   the reduction over channels (the decoder's torch.cat([upsampled, skip], 1)
 *   is never materialised); inputs larger than the output are cropped to it
 *   (DecoderBlock.forward, pixloc/pixloc/pixlib/models/unet.py:33-44).
 *   weights: fp16 [taps][C_out][cin0 + cin1]; bias: fp32 [C_out].
 *   Replaces the cuDNN convolutions of unet.py:22-31,68-99.
 *
 * PtkExtractor: the whole UNet._forward (unet.py:158-190) with the PixLoc
 *   configuration as a native plan for one input size, including the
 *   pre-processing of PixTrackFeatureExtractor.__call__
 *   (pixtrack/localization/feature_extractor.py:34-59): the image
 *   [img_h][img_w][3] fp32 0..255 is bilinearly resized to the plan's H x W
 *   (cv2.INTER_LINEAR semantics), scaled to 0..1 and normalised.
 *   Outputs per level l = 0,1,2 (fine -> coarse): feat[l] fp32 [H_l][W_l][C_l]
 *   channels-last, conf[l] fp32 [H_l][W_l] = sigmoid(-uncertainty); with
 *   normalize != 0 the descriptors are L2-normalised over C per pixel
 *   (base_refiner.py:92-94 fused into the head).
 * ---------------------------------------------------------------------- */
int ptk_conv_f16(PtkContext* ctx, const void* in0, int32_t cin0, const void* in1, int32_t cin1, int32_t H, int32_t W,
                 int32_t in0_H, int32_t in0_W, int32_t in1_H, int32_t in1_W, const void* weights, const float* bias,
                 int32_t Cout, int32_t taps, int32_t relu, void* out, void* stream);
/* Same, and when pool_out != NULL the epilogue also writes the 2x2 / stride-2 max pool (floor mode) of the result,
 * fp16 [H/2][W/2][C_out]: the nn.MaxPool2d that opens the next encoder block (unet.py:68-99), fused. */
int ptk_conv_f16_pool(PtkContext* ctx, const void* in0, int32_t cin0, const void* in1, int32_t cin1, int32_t H, int32_t W,
                      int32_t in0_H, int32_t in0_W, int32_t in1_H, int32_t in1_W, const void* weights,
                      const float* bias, int32_t Cout, int32_t taps, int32_t relu, void* out, void* pool_out,
                      void* stream);

typedef struct PtkUnetWeights {
  const void* conv_w[20];   /* [0]: fp32 [64][28] ((ky,kx,c) taps + 1 pad); [1..19]: fp16 [9][C_out][C_in];
                               16 encoder convs then 4 decoder convs (BatchNorm folded in)              */
  const float* conv_b[20];  /* fp32 [C_out]                                                             */
  const void* head_w[3];    /* fp16 [C_l + 1][C_in]: adaptation rows, last row = uncertainty head       */
  const float* head_b[3];   /* fp32 [C_l + 1]                                                           */
} PtkUnetWeights;

typedef struct PtkExtractor PtkExtractor;
int ptk_extractor_create(PtkContext* ctx, const PtkUnetWeights* w, int32_t H, int32_t W, PtkExtractor** out);
void ptk_extractor_destroy(PtkExtractor* e);
int ptk_extractor_level_shape(const PtkExtractor* e, int32_t level, int32_t* C, int32_t* H, int32_t* W);
/* image: [img_h][img_w][3] RGB, img_dtype 0 = fp32 in 0..255, 1 = uint8 (converted to float before the
 * bilinear resize; the reference's cv2 path for uint8 rounds the resized image back to uint8 first).
 * Stream-ordered.  A binding (image pointer + the six output pointers + sizes) is launched kernel by kernel the
 * first time it is seen, captured into a CUDA graph the second time and replayed afterwards (8 bindings per extractor, least
 * recently used replaced; PTK_PLAN_GRAPH=0 disables it; inside a caller's own stream capture the kernels are launched directly),
 * so the buffers of a binding must stay allocated for as long as it is used -- as for any prepared launch. */
int ptk_extractor_run(PtkExtractor* e, const void* image, int32_t img_dtype, int32_t img_h, int32_t img_w,
                      float* const* feat, float* const* conf, int32_t normalize, void* stream);
/* Benchmark helper (SYNCHRONISES): runs the plan once with a CUDA event after every launch.  ms[i] = device
 * time of launch i; kinds[i]: 0 prep, 1 first conv (mma.sync), 3 tensor-core conv (2x2 max pools fused), 4 upsample,
 * 5 head; flops[i] = 2 x multiply-adds of launch i. */
int ptk_extractor_profile(PtkExtractor* e, const void* image, int32_t img_dtype, int32_t img_h, int32_t img_w,
                          float* const* feat, float* const* conf, int32_t normalize, void* stream, int32_t max_n,
                          float* ms, int32_t* kinds, double* flops, int32_t* n_out);
/* Kernel launches of the plan's last run: 28, or 27 when the level-0 head ran inside the epilogue of the last decoder
 * convolution (PTK_FUSE_HEAD=1; off by default, it measured slower). */
int ptk_extractor_launch_count(const PtkExtractor* e, int32_t* n);
/* test access to intermediate fp16 NHWC activations: kind 0 = encoder block output, 1 = decoder block output */
int ptk_extractor_activation(const PtkExtractor* e, int32_t kind, int32_t index, const void** ptr, int32_t* C,
                             int32_t* H, int32_t* W);

/* ------------------------------------------------------------------------
 * NeRF reference-view render (instant-ngp hash grid + fused MLPs + occupancy
 * marching + compositing) as one persistent launch.
 *
 * Replaces pyngp Testbed.render(width, height, spp, linear=True)
 *   instant-ngp/src/python_api.cu:127-173,352 -> Testbed::render_frame
 *   (src/testbed.cu:2591-2749) -> render_nerf / NerfTracer
 *   (src/testbed_nerf.cu:606-955,1721-2146,2228-2330) and
 *   NerfNetwork::inference_mixed_precision_impl (nerf_network.h:101-136) with
 *   tiny-cuda-nn's kernel_grid / kernel_sh / kernel_mlp_fused,
 * as called by get_nerf_image (pixtrack/visualization/run_vis_on_poses.py:28-57)
 * with the testbed settings of pixtrack/utils/ingp_utils.py:22-44
 * (snap_to_pixel_centers, fov_axis 0, exposure 0, identity tone curve).
 *
 * PtkNerfModel: an unpacked snapshot for the configs/nerf/base.json network
 *   (16-level hash grid F=2 T=2^19 base 16, density MLP 32-64-16, rgb MLP
 *   32-64-64-3, ReLU, SH degree 4).  grid: fp16 [n_grid_entries][2]; weights:
 *   fp16 row-major [out][in]: density [64][32], [16][64]; rgb [64][32],
 *   [64][64], [16][64]; bitfield: uint8 [8 * 128^3 / 8] occupancy bits, Morton
 *   order per cascade (density_grid_bitfield).  The arrays must stay alive.
 * PtkNerfView: camera = 3x4 row-major camera-to-world in NGP convention
 *   (NerfDataset::nerf_matrix_to_ngp applied); focal in pixels (both axes);
 *   background already in linear colour; depth_mode 0 = Shade, 1 = Depth.
 * Outputs (any may be NULL except that one of rgba/u8 is required):
 *   rgba float [H][W][4] (what Testbed.render returns), u8 [H][W][3] =
 *   (rgb * 255).astype(uint8) (what get_nerf_image returns), depth [H][W].
 * ---------------------------------------------------------------------- */
typedef struct PtkNerfModel {
  const void* grid;
  int64_t n_grid_entries;
  const void* weights[5];
  const uint8_t* bitfield;
  int32_t aabb_scale;
  int32_t reserved0;
} PtkNerfModel;

typedef struct PtkNerfView {
  float camera[12];
  float render_aabb_min[3];
  float render_aabb_max[3];
  float focal;
  float depth_scale;        /* 1 / dataset scale */
  float min_transmittance;  /* testbed.nerf.rendering_min_transmittance */
  float background[4];
  int32_t width, height, spp, depth_mode;
} PtkNerfView;

typedef struct PtkNerf PtkNerf;
int ptk_nerf_create(PtkContext* ctx, const PtkNerfModel* model, PtkNerf** out);
void ptk_nerf_destroy(PtkNerf* n);
/* hash-grid entries (of 2 fp16 features) the base.json encoding has at this aabb_scale (host only) */
int64_t ptk_nerf_grid_entries(int32_t aabb_scale);
int ptk_nerf_render(PtkNerf* n, const PtkNerfView* view, float* out_rgba, uint8_t* out_u8, float* out_depth,
                    void* stream);
/* The network alone: replaces NerfNetwork::inference_mixed_precision_impl
 *   (include/neural-graphics-primitives/nerf_network.h:101-136: hash-grid encoding -> density MLP -> [16 | SH16] -> rgb
 *   MLP) for `count` inputs.  pos01 [count][3]: positions in the unit cube of the training box (warp_position applied);
 *   dir [count][3]: unit view directions.  out_rgbd [count][4] fp32: raw r, g, b, raw density (the fp16 values the
 *   reference's network emits, before the activations); out_features [count][32] fp32 or NULL: the hash-grid encoding. */
int ptk_nerf_eval(PtkNerf* n, const float* pos01, const float* dir, int32_t count, float* out_rgbd, float* out_features,
                  void* stream);
/* SYNCHRONISING statistics of the last render of `n` (benchmarks): out4 = network samples evaluated, warp steps,
 * rays marched (those that reach an occupied cell), lane slots that sat out a warp step. */
int ptk_nerf_stats(PtkNerf* n, uint64_t* out4);

/* ------------------------------------------------------------------------
 * Query-frame object mask.
 *
 * Replaces PixLocPoseTrackerR9.get_mask and the multiply in refine()
 *   pixtrack/pose_trackers/pixloc_tracker_r9.py:207-214,224-225:
 *   mask = dilate_5x5^5(erode_5x5((depth != 0))) with OpenCV's default
 *   morphology border (outside pixels ignored), query = query * mask.
 * depth_u8: [H][W][3] uint8, the depth-mode render (ptk_nerf_render out_u8).
 * image / out_image: [H][W][3], img_dtype 0 = fp32, 1 = uint8 (may alias);
 *   both NULL to get only the mask.  out_mask: [H][W][3] uint8 0/1 or NULL.
 * workspace: device scratch of at least 2 * H * W * 3 bytes.
 * ---------------------------------------------------------------------- */
int ptk_query_mask(PtkContext* ctx, const uint8_t* depth_u8, int32_t H, int32_t W, const void* image,
                   int32_t img_dtype, void* out_image, uint8_t* out_mask, uint8_t* workspace, void* stream);

/* ------------------------------------------------------------------------
 * Result overlay (visualisation of a tracked frame).
 *
 * Replaces, per frame of pixtrack/visualization/run_vis_on_poses.py:289-371:
 *   blend_images (:215-219): out = uint8(query * alpha + swap_rb(nerf) * (1 - alpha)) in float64;
 *   draw_axes (:74-79): three thick lines between the end points add_pose_axes (:82-112) projects on the host.
 * query / out: [H][W][3] uint8 in the camera frame's channel order (BGR from cv2.imread); nerf: [H][W][3] uint8 RGB
 * render (ptk_nerf_render out_u8) or NULL = white; host_axes_px: HOST int16 [6][2] end points (x0 y0 x1 y1 per axis)
 * or NULL = no axes; thickness in pixels.
 * ---------------------------------------------------------------------- */
int ptk_overlay(PtkContext* ctx, const uint8_t* query, const uint8_t* nerf, int32_t H, int32_t W, double alpha,
                const int16_t* host_axes_px, int32_t thickness, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIXTRACK_B200_H_ */
