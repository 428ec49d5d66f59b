"""Sparse map sampling (the `opt.interpolator` contract) on the B200 path.

`Interpolator(feats[C,H,W], p2d[N,2], return_gradients=False) ->
(vals[N,C], mask[N] bool, grads[N,C,2])` as reference
pixloc/pixloc/pixlib/geometry/interpolation.py:131-141; called by
`PoseTrackerRefiner.interp_sparse_observations`
(pixtrack/localization/pixloc_pose_refiners.py:349-351).
"""
import ctypes as C
from typing import Sequence, Tuple

import torch

from . import _lib

Tensor = torch.Tensor


def sample_points(feats: Tensor, p2d: Tensor, pad: int = 1, return_gradients: bool = False):
    if not feats.is_cuda:
        raise _lib.PtkError('sample_points needs CUDA tensors (no CPU fallback)')
    assert feats.dim() == 3 and feats.dtype == torch.float32
    Cc, H, W = feats.shape
    pts = p2d.to(feats.device, torch.float32).contiguous()
    N = pts.shape[0]
    vals = torch.empty((N, Cc), dtype=torch.float32, device=feats.device)
    mask = torch.empty((N,), dtype=torch.uint8, device=feats.device)
    grads = torch.empty((N, Cc, 2), dtype=torch.float32, device=feats.device) if return_gradients else None
    sc, sy, sx = feats.stride()
    dev = feats.device.index if feats.device.index is not None else torch.cuda.current_device()
    _lib.check(_lib.load().ptk_sample_points(
        _lib.context(dev), feats.data_ptr(), sc, sy, sx, Cc, H, W, pts.data_ptr(), N, int(pad), vals.data_ptr(),
        mask.data_ptr(), None if grads is None else grads.data_ptr(), _lib.current_stream_ptr(feats.device)))
    if grads is None:
        grads = torch.zeros((N, Cc, 2), dtype=torch.float32, device=feats.device)
    return vals, mask.bool(), grads


class Interpolator:
    def __init__(self, mode: str = 'linear', pad: int = 1):
        if mode != 'linear':
            raise NotImplementedError('the PixTrack configuration uses linear interpolation')
        self.mode, self.pad = mode, pad

    def __call__(self, tensor: Tensor, pts: Tensor, return_gradients: bool = False):
        return sample_points(tensor, pts, self.pad, return_gradients)


def sample_reference(feats_hwc: Sequence[Tensor], confs: Sequence[Tensor], scales: Sequence[Tuple[float, float]],
                     camera, T_w2cam, p3d: Tensor, pad: int = 1, normalize: bool = True, out=None):
    """All-level reference observations in one launch (csrc/ptk_sample_ref.cu): the B200 form of
    `PoseTrackerRefiner.interp_sparse_observations` (reference
    pixtrack/localization/pixloc_pose_refiners.py:327-368) + the reference-side normalisation of
    `refine_pose_using_features` (pixloc/pixloc/localization/base_refiner.py:74-84).

      feats_hwc[l] [H_l,W_l,C_l] fp32 channels-last, NOT normalised; confs[l] [H_l,W_l];
      scales[l] = (sx, sy) of level l w.r.t. the reference image; camera: Camera-like (`._data`) or a
      vector [w,h,fx,fy,cx,cy,dist...] at the reference-image resolution; T_w2cam: Pose-like or
      [12] (R row-major, t); p3d: CUDA float64 [N,3] (the reference projects in float64).
    Returns (F_ref list [N,C_l], W_ref list [N], valid [N] uint8).  `out=(F_ref, W_ref, valid)`
    writes into caller-owned buffers (static addresses for prepared LM launches).  Stream-ordered."""
    dev = feats_hwc[0].device
    if dev.type != 'cuda':
        raise _lib.PtkError('sample_reference needs CUDA tensors (no CPU fallback)')
    assert p3d.is_cuda and p3d.dtype == torch.float64 and p3d.is_contiguous()
    N = p3d.shape[0]
    cam = getattr(camera, '_data', camera).detach().double().cpu().reshape(-1)
    T = getattr(T_w2cam, '_data', T_w2cam).detach().double().cpu().reshape(-1)
    assert T.numel() == 12 and cam.numel() in (6, 8, 10)
    L = len(feats_hwc)
    if out is None:
        F_ref = [torch.empty((N, f.shape[2]), dtype=torch.float32, device=dev) for f in feats_hwc]
        W_ref = [torch.empty((N,), dtype=torch.float32, device=dev) for _ in feats_hwc]
        valid = torch.empty((N,), dtype=torch.uint8, device=dev)
    else:
        F_ref, W_ref, valid = out
    if N == 0:
        return F_ref, W_ref, valid
    lv = (_lib.RefLevel * L)()
    for l in range(L):
        f, c = feats_hwc[l], confs[l]
        assert f.dtype == torch.float32 and f.is_contiguous() and c.is_contiguous()
        H, W, Cc = f.shape
        lv[l].feat, lv[l].conf = f.data_ptr(), c.data_ptr()
        lv[l].f_out, lv[l].w_out = F_ref[l].data_ptr(), W_ref[l].data_ptr()
        lv[l].sx, lv[l].sy = float(scales[l][0]), float(scales[l][1])
        lv[l].C, lv[l].H, lv[l].W, lv[l].normalize = Cc, H, W, 1 if normalize else 0
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    cam_c = (C.c_double * cam.numel())(*cam.tolist())
    T_c = (C.c_double * 12)(*T.tolist())
    _lib.check(_lib.load().ptk_sample_reference(_lib.context(di), lv, L, p3d.data_ptr(), N, cam_c, cam.numel(), T_c,
                                                int(pad), valid.data_ptr(), _lib.current_stream_ptr(dev)))
    return F_ref, W_ref, valid
