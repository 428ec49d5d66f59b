"""Device-side query mask: drop-in for `PixLocPoseTrackerR9.get_mask` + `query_image * mask`.

Reference pixtrack/pose_trackers/pixloc_tracker_r9.py:207-214,224-225: the NeRF depth render at the query
resolution is thresholded (`depth != 0`), eroded once and dilated five times with a 5x5 kernel (OpenCV on the
host), and multiplied into the query frame.  Here the depth render never leaves the device
(`get_nerf_image(..., depth=True, device_output=True)`) and the morphology + multiply is four small launches
(csrc/ptk_mask.cu).  No CPU fallback.
"""
from typing import Optional, Tuple

import torch

from . import _lib

Tensor = torch.Tensor


def query_mask(depth_u8: Tensor, image: Optional[Tensor] = None, want_mask: bool = False,
               out: Optional[Tensor] = None) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """depth_u8: CUDA uint8 [H,W,3] depth-mode render; image: CUDA [H,W,3] uint8 or fp32 query frame.
    Returns (masked image or None, mask uint8 [H,W,3] of 0/1 or None).  Stream-ordered, no sync."""
    if not depth_u8.is_cuda:
        raise _lib.PtkError('query_mask needs CUDA tensors (no CPU fallback)')
    assert depth_u8.dtype == torch.uint8 and depth_u8.dim() == 3 and depth_u8.shape[2] == 3 and depth_u8.is_contiguous()
    H, W = depth_u8.shape[:2]
    dev = depth_u8.device
    if image is not None:
        assert image.shape == depth_u8.shape and image.is_contiguous() and image.dtype in (torch.uint8, torch.float32)
        if out is None:
            out = torch.empty_like(image)
    mask = torch.empty_like(depth_u8) if want_mask else None
    ws = torch.empty(2 * H * W * 3, dtype=torch.uint8, device=dev)
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    _lib.check(_lib.load().ptk_query_mask(
        _lib.context(di), depth_u8.data_ptr(), H, W, None if image is None else image.data_ptr(),
        0 if (image is not None and image.dtype == torch.float32) else 1, None if image is None else out.data_ptr(),
        None if mask is None else mask.data_ptr(), ws.data_ptr(), _lib.current_stream_ptr(dev)))
    return (out if image is not None else None), mask
