"""FrameTracker: the per-frame hot path on one GPU -- query feature extraction followed by the
coarse-to-fine LM refinement against B cached reference views -- over static device buffers.

This is the B200 form of one `PixLocPoseTrackerR9.refine` call restricted to the hot path
(reference pixtrack/pose_trackers/pixloc_tracker_r9.py:216-266 ->
PoseTrackerRefiner.refine_query_pose, pixtrack/localization/pixloc_pose_refiners.py:200-271 ->
dense_feature_extraction + refine_pose_using_features): the query pyramid is written by the native
UNet plan straight into the buffers the prepared LM launches read (descriptors already
L2-normalised by the fused head), so a frame is: one image upload, ~35 extractor launches, one
CUDA-graph launch of the 3-level LM chain, one 13-float-per-view read-back.
"""
from typing import Dict, List, Optional, Sequence

import torch

from .extractor import B200FeatureExtractor
from .geometry import Camera
from .refiner import FramePlan

Tensor = torch.Tensor


class FrameTracker:
    def __init__(self, extractor: B200FeatureExtractor, image_hw, camera: Tensor, p3d: Tensor,
                 F_ref: Sequence[Tensor], W_ref: Sequence[Tensor], lams: Sequence[Tensor], n_views: int,
                 scale_image: int = 1, use_graph: bool = True, **lm_conf):
        """camera: [n_cam] at the IMAGE resolution; p3d [N,3]; F_ref[l] [B,N,C_l] (normalised),
        W_ref[l] [B,N]; lams[l] [6]; lm_conf -> LmLaunch (num_iters, stop criteria, pad)."""
        dev = extractor.device
        self.extractor, self.scale_image, self.B = extractor, scale_image, n_views
        ih, iw = image_hw
        H, W, sr = extractor.network_size(ih, iw, scale_image)
        shapes = extractor.plan(H, W).shapes
        self.feats = [torch.zeros((h, w, c), dtype=torch.float32, device=dev) for c, h, w in shapes]
        self.confs = [torch.zeros((h, w), dtype=torch.float32, device=dev) for c, h, w in shapes]
        cam = Camera(camera.detach().cpu().double())
        self.cams = [cam.scale((sr[0] / s, sr[1] / s))._data.float().to(dev) for s in extractor.model.scales]
        self.T_init = torch.zeros((n_views, 12), dtype=torch.float32, device=dev)
        self.plan = FramePlan(self.feats, self.confs, self.cams, [f.to(dev) for f in F_ref],
                              [w.to(dev) for w in W_ref], p3d.to(dev), self.T_init, [l.to(dev) for l in lams], **lm_conf)
        if use_graph:
            self.plan.capture()

    def track(self, image: Tensor, T_init: Optional[Tensor] = None):
        """image: CUDA [H,W,3] uint8/fp32.  Returns (T [B,12], failed [B]) device tensors; stream-ordered,
        no synchronisation."""
        if T_init is not None:
            self.T_init.copy_(T_init, non_blocking=True)
        self.extractor.extract_device(image, self.scale_image, normalize=True, out=(self.feats, self.confs))
        self.plan.run()
        return self.plan.T, self.plan.failed
