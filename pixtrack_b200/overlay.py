"""Result overlay: the NeRF render at the tracked pose blended over the camera frame, with the pose axes.

Same image as the per-frame body of reference pixtrack/visualization/run_vis_on_poses.py:289-371 builds with OpenCV on the
host (`blend_images` :215-219, `add_pose_axes` :82-112 -> `draw_axes` :74-79), composed on the device from the tracker's own
render (`get_nerf_image(..., device_output=True)`), so a tracked sequence can be visualised without a second renderer or a
round trip per frame.  The 3-D -> 2-D end points of the axes are host arithmetic (six points); blending and line drawing
are one launch (csrc/ptk_overlay.cu).  No CPU fallback.
"""
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

Tensor = torch.Tensor


def project_3d_to_2d(pts_3d: np.ndarray, K: np.ndarray) -> np.ndarray:
    """run_vis_on_poses.py:66-70."""
    p = K @ np.asarray(pts_3d, np.float64).T
    p = p / p[2, :]
    return p[:2, :].T


def pose_axes_points(camera, pose: np.ndarray, axes_center: Sequence[float] = (0.1179, 1.1538, 1.3870, 0.0)) -> np.ndarray:
    """End points of the three axis segments in pixels, int16 [6, 2] (x-axis: rows 0-1, y: 2-3, z: 4-5), as
    add_pose_axes computes them (run_vis_on_poses.py:82-112): K from the camera's focal length with the principal point
    at the image centre, axes of length 0.025 along +x, -y, -z from `axes_center` (homogeneous, w = 0 is ADDED to the
    w = 1 of the axes), `pose` = 4x4 camera-in-world matrix (get_camera_in_world_from_pixpose)."""
    width, height = (float(v) for v in camera.size)
    focal = float(camera.f[0])
    K = np.array([[focal, 0.0, width / 2], [0.0, focal, height / 2], [0.0, 0.0, 1.0]])
    s = 0.25 * 0.1
    axes = np.array([[0, 0, 0], [s, 0, 0], [0, 0, 0], [0, -s, 0], [0, 0, 0], [0, 0, -s]], np.float64)
    axes = np.hstack((axes, np.ones((6, 1))))
    axes = axes + np.array(axes_center, np.float64)
    pts_3d = axes @ np.linalg.inv(np.asarray(pose, np.float64)).T[:, :3]
    return project_3d_to_2d(pts_3d, K).astype(np.int16)


def overlay(query_bgr: Tensor, nerf_rgb: Optional[Tensor], alpha: float = 0.3, axes_px: Optional[np.ndarray] = None,
            thickness: int = 2, out: Optional[Tensor] = None) -> Tensor:
    """query_bgr: CUDA uint8 [H,W,3] camera frame as cv2.imread returns it; nerf_rgb: CUDA uint8 [H,W,3] render
    (get_nerf_image output) or None (the reference substitutes a white image when a frame has no pose); axes_px: int16
    [6,2] from pose_axes_points or None.  Returns the result image (uint8 [H,W,3], the frame's channel order)."""
    if not query_bgr.is_cuda:
        raise _lib.PtkError('overlay needs CUDA tensors (no CPU fallback)')
    assert query_bgr.dtype == torch.uint8 and query_bgr.dim() == 3 and query_bgr.shape[2] == 3 and query_bgr.is_contiguous()
    H, W = query_bgr.shape[:2]
    if nerf_rgb is not None:
        assert nerf_rgb.shape == query_bgr.shape and nerf_rgb.dtype == torch.uint8 and nerf_rgb.is_contiguous()
    if out is None:
        out = torch.empty_like(query_bgr)
    dev = query_bgr.device
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    ax = None
    if axes_px is not None:
        a = np.ascontiguousarray(np.asarray(axes_px, np.int16).reshape(12))
        ax = a.ctypes.data_as(_lib.C.POINTER(_lib.C.c_int16))
    _lib.check(_lib.load().ptk_overlay(_lib.context(di), query_bgr.data_ptr(), None if nerf_rgb is None else nerf_rgb.data_ptr(),
                                       H, W, float(alpha), ax, int(thickness), out.data_ptr(), _lib.current_stream_ptr(dev)))
    return out
