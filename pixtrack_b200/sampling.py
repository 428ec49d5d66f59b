"""Sparse map sampling (the `opt.interpolator` contract) on the B200 path.

`Interpolator(feats[C,H,W], p2d[N,2], return_gradients=False) ->
(vals[N,C], mask[N] bool, grads[N,C,2])` as reference
pixloc/pixloc/pixlib/geometry/interpolation.py:131-141; called by
`PoseTrackerRefiner.interp_sparse_observations`
(pixtrack/localization/pixloc_pose_refiners.py:349-351).
"""
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
