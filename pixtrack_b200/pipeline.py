"""FrameTracker: the per-frame hot path on one GPU over static device buffers.

This is the B200 form of one `PixLocPoseTrackerR9.refine` call restricted to the hot path
(reference pixtrack/pose_trackers/pixloc_tracker_r9.py:216-266):

  refresh_reference(view, image, camera, pose)
      = create_dynamic_reference_image's feature half (r9.py:154-160 ->
        PoseTrackerRefiner.extract_reference_features, pixtrack/localization/pixloc_pose_refiners.py:273-325):
        dense extraction of the rendered reference view + interp_sparse_observations (:327-368).
        Here: the native UNet plan, then ONE ptk_sample_reference launch that writes the
        normalised descriptors / confidences / validity of all levels straight into the view's slot
        of the observation cache the LM launches read.
  track(image, T_init)
      = PoseTrackerRefiner.refine_query_pose (:200-271): dense_feature_extraction of the query +
        refine_pose_using_features (pixloc/pixloc/localization/base_refiner.py:64-137) against the B
        cached views.  Here: the UNet plan writes the L2-normalised query pyramid into the buffers
        the prepared LM launches read; the 3-level coarse-to-fine chain is one CUDA-graph launch.

A frame is: image upload(s), ~35 extractor launches per image, one sampling launch, one graph
launch, one 13-float-per-view read-back; nothing synchronises inside.
"""
from typing import Optional, Sequence

import torch

from .extractor import B200FeatureExtractor
from .geometry import Camera
from .refiner import FramePlan
from .sampling import sample_reference

Tensor = torch.Tensor


def morton_order(xyz) -> torch.Tensor:
    """Permutation that sorts points [N,3] along a 3-D Morton (Z-order) curve over their bounding box (10 bits per axis).
    Points that are close in space become close in memory, hence close in the image under any pose: the LM kernel hands
    contiguous index ranges to its CTAs and walks them in passes of 1024, so the 12-texel footprints of a pass overlap and
    are served by the SM's L1 instead of L2 (measured effect: see FrameTracker's `sort_points`).  Sums are
    order-independent up to rounding; the reference keeps COLMAP's dictionary order, which carries no meaning."""
    x = torch.as_tensor(xyz, dtype=torch.float64).cpu()
    if x.shape[0] == 0:
        return torch.zeros(0, dtype=torch.long)
    lo, hi = x.min(0).values, x.max(0).values
    q = ((x - lo) / (hi - lo).clamp(min=1e-300) * 1023.0).round().clamp(0, 1023).to(torch.int64)

    def spread(v):                       # 10 bits -> every third bit
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    return torch.argsort(code, stable=True)


class FrameTracker:
    def __init__(self, extractor: B200FeatureExtractor, image_hw, camera: Tensor, p3d: Tensor, lams: Sequence[Tensor],
                 n_views: int, scale_image: int = 1, use_graph: bool = True, pad: int = 1,
                 overlap_reference: bool = True, sort_points: bool = False, **lm_conf):
        """camera: [n_cam] query camera at the IMAGE resolution; p3d [N,3] model points (float64 kept for
        the reference-side projection, float32 copy for the LM); lams[l] [6] damping per level;
        lm_conf -> LmLaunch (num_iters, stop criteria).  sort_points: keep the points in Morton order on the device
        (`self.order[i]` = caller's index of stored row i; poses do not depend on it beyond rounding).  Off by default:
        measured on B200 (profiles/r2/lm_order.json + lm_*_ncu.json) it raises the L1 hit rate of the LM gathers from 3 %
        to 33 % and cuts L2 -> SM traffic by 31 %, but the launch gets 4 % SLOWER (181 -> 188 us per iteration at
        C = 128, N = 20000, B = 16): the kernel is bound by issue + latency at 16 warps per SM, not by that traffic."""
        dev = extractor.device
        self.extractor, self.scale_image, self.B, self.pad = extractor, scale_image, n_views, pad
        ih, iw = image_hw
        H, W, sr = extractor.network_size(ih, iw, scale_image)
        shapes = extractor.plan(H, W).shapes
        N = p3d.shape[0]
        self.sort_points = sort_points
        self.order = morton_order(p3d) if sort_points else torch.arange(N)
        self.p3d64 = torch.as_tensor(p3d, dtype=torch.float64).cpu()[self.order].to(dev).contiguous()
        self.p3d32 = self.p3d64.float()
        # query pyramid (channels-last, L2-normalised by the fused head) + confidences
        self.feats = [torch.zeros((h, w, c), dtype=torch.float32, device=dev) for c, h, w in shapes]
        self.confs = [torch.zeros((h, w), dtype=torch.float32, device=dev) for c, h, w in shapes]
        # reference observation cache: F_ref[l] [B,N,C_l], W_ref[l] [B,N], valid [B,N]
        self.F_ref = [torch.zeros((n_views, N, c), dtype=torch.float32, device=dev) for c, _, _ in shapes]
        self.W_ref = [torch.zeros((n_views, N), dtype=torch.float32, device=dev) for _ in shapes]
        self.valid = torch.zeros((n_views, N), dtype=torch.uint8, device=dev)
        cam = Camera(camera.detach().cpu().double())
        self.cams = [cam.scale((sr[0] / s, sr[1] / s))._data.float().to(dev) for s in extractor.model.scales]
        self.T_init = torch.zeros((n_views, 12), dtype=torch.float32, device=dev)
        self.plan = FramePlan(self.feats, self.confs, self.cams, self.F_ref, self.W_ref, self.p3d32, self.T_init,
                              [l.to(dev) for l in lams], mask=self.valid, pad=pad, **lm_conf)
        self._use_graph, self._captured = use_graph, False
        self._ref_bufs = {}
        self._masked = None
        # the reference-view refresh runs on a side stream next to the query extraction (its low-occupancy layers
        # overlap the other plan's); the LM launches wait for it through an event
        self._side = torch.cuda.Stream(dev) if overlap_reference else None
        self._ref_done = None
        self._pre_extract = None
        self._query_ready = False
        self.n_active = N

    def set_points(self, xyz):
        """Replace the model points by `xyz` [n <= N, 3] (host or device): the point set follows the reference image
        (Model3D.get_p3did_to_dbids, pixloc/pixloc/localization/model3d.py:49-87).  Buffers keep their size and
        addresses (prepared launches / graphs stay valid); rows n.. are switched off through the validity flags at
        the next refresh_reference.  Stream-ordered."""
        n = int(xyz.shape[0])
        if n > self.p3d64.shape[0]:
            raise ValueError(f'{n} points exceed the capacity {self.p3d64.shape[0]} this tracker was built for')
        src = torch.as_tensor(xyz, dtype=torch.float64)
        if self.sort_points:
            self.order = morton_order(src)
            src = src.cpu()[self.order]
        else:
            self.order = torch.arange(n)
        src = src.to(self.p3d64.device, non_blocking=True)
        self.p3d64[:n].copy_(src)
        self.p3d32[:n].copy_(src)
        self.n_active = n

    def last_costs(self):
        """Per view, the final mean cost of every level the LM visited, coarse to fine -- what
        DebugTracker.costs[level][-1] holds in the reference (pixtrack/localization/tracker.py:37-46, read at
        pixloc_tracker_r9.py:251).  Synchronises (reads the iteration counts and one log record per level)."""
        out = [[] for _ in range(self.B)]
        for L in self.plan.launches:
            n_it = L.n_iters.cpu()
            for b in range(self.B):
                n = int(n_it[b])
                if n > 0 and L.log is not None:
                    rec = L.log[b, n - 1, :2].cpu()
                    out[b].append(float(rec[0] / rec[1]))
        return out

    def refresh_reference(self, view: int, image: Tensor, camera, T_w2cam, scale_image: int = 1):
        """image: CUDA [H,W,3] uint8/fp32 render of reference view `view`; camera / T_w2cam: its camera at the
        image resolution and its world-to-camera pose (host, float64).  Stream-ordered."""
        ih, iw = image.shape[:2]
        key = (ih, iw, scale_image)
        if key not in self._ref_bufs:
            shapes = self.extractor.level_shapes(ih, iw, scale_image)
            dev = self.extractor.device
            self._ref_bufs[key] = ([torch.empty((h, w, c), dtype=torch.float32, device=dev) for c, h, w in shapes],
                                   [torch.empty((h, w), dtype=torch.float32, device=dev) for c, h, w in shapes])
        bufs = self._ref_bufs[key]
        dev = self.extractor.device
        if self._side is None:
            feats, confs, scales = self.extractor.extract_device(image, scale_image, normalize=False, out=bufs)
            sample_reference(feats, confs, scales, camera, T_w2cam, self.p3d64, pad=self.pad, normalize=True,
                             out=([f[view] for f in self.F_ref], [w[view] for w in self.W_ref], self.valid[view]))
            if self.n_active < self.valid.shape[1]:
                self.valid[view, self.n_active:].zero_()
            return
        cur = torch.cuda.current_stream(dev)
        if self._query_ready and self._pre_extract is not None:
            # this frame's query extraction was enqueued ahead (extract_query): wait for what preceded it on the caller's
            # stream -- the previous frame's LM, the uploads the stream had waited for -- not for the extraction itself,
            # so that the two plans still run side by side
            self._side.wait_event(self._pre_extract)
        else:
            self._side.wait_stream(cur)      # the previous frame's LM has finished reading the observation cache
        with torch.cuda.stream(self._side):
            feats, confs, scales = self.extractor.extract_device(image, scale_image, normalize=False, out=bufs, slot=1)
            sample_reference(feats, confs, scales, camera, T_w2cam, self.p3d64, pad=self.pad, normalize=True,
                             out=([f[view] for f in self.F_ref], [w[view] for w in self.W_ref], self.valid[view]))
            if self.n_active < self.valid.shape[1]:
                self.valid[view, self.n_active:].zero_()
            self._ref_done = torch.cuda.Event()
            self._ref_done.record(self._side)
        image.record_stream(self._side)

    def extract_query(self, image: Tensor, mask_depth: Optional[Tensor] = None):
        """The pose-independent half of `track`: (mask and) extract the query frame's features into the tracker's maps.
        A caller that has to read frame i's poses on the host before it can start frame i+1's refinement can enqueue this
        for frame i+1 first -- the extraction then runs while the host waits for frame i's result -- and call
        `track(None, T_init)` afterwards.  Stream-ordered after the previous `track` (whose LM reads the same maps)."""
        if mask_depth is not None:
            from .mask import query_mask
            if self._masked is None or self._masked.shape != image.shape or self._masked.dtype != image.dtype:
                self._masked = torch.empty_like(image)
            image, _ = query_mask(mask_depth, image, out=self._masked)
        if self._side is not None:
            self._pre_extract = torch.cuda.Event()
            self._pre_extract.record(torch.cuda.current_stream(self.extractor.device))
        self.extractor.extract_device(image, self.scale_image, normalize=True, out=(self.feats, self.confs))
        self._query_ready = True

    def track(self, image: Optional[Tensor], T_init: Optional[Tensor] = None, mask_depth: Optional[Tensor] = None):
        """image: CUDA [H,W,3] uint8/fp32 query frame, or None when `extract_query` already ran for this frame.
        mask_depth: optional CUDA uint8 [H,W,3] depth-mode NeRF render at the query resolution; the frame is then
        multiplied by the eroded/dilated object mask first (r9.py:207-214,224-225), on the device.  Returns
        (T [B,12], failed [B]) device tensors; stream-ordered, no synchronisation."""
        if self._use_graph and not self._captured:
            self.plan.capture()
            self._captured = True
        if T_init is not None:
            self.T_init.copy_(T_init, non_blocking=True)
        if image is not None:
            self.extract_query(image, mask_depth)
        elif not self._query_ready:
            raise RuntimeError('track(None, ...) needs a preceding extract_query() for this frame')
        self._query_ready = False
        if self._ref_done is not None:       # observations refreshed on the side stream must be complete
            torch.cuda.current_stream(self.extractor.device).wait_event(self._ref_done)
            self._ref_done = None
        self.plan.run()
        return self.plan.T, self.plan.failed
