"""CPU oracle for the query-frame mask (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates reference pixtrack/pose_trackers/pixloc_tracker_r9.py:207-214 with numpy: `(depth != 0)`, one
5x5 erosion, five 5x5 dilations, OpenCV's default morphology border (pixels outside the image do not
take part: erosion pads with the maximum, dilation with the minimum; anchor at the kernel centre).
Pinned against cv2 itself in tests/test_mask.py (cv2 is the third-party routine the reference calls).
"""
import numpy as np


def _box(m: np.ndarray, r: int, use_max: bool) -> np.ndarray:
    pad_val = 0 if use_max else 1
    H, W = m.shape[:2]
    p = np.pad(m, ((r, r), (r, r)) + ((0, 0),) * (m.ndim - 2), constant_values=pad_val)
    out = m.copy()
    for dy in range(2 * r + 1):
        for dx in range(2 * r + 1):
            w = p[dy:dy + H, dx:dx + W]
            out = np.maximum(out, w) if use_max else np.minimum(out, w)
    return out


def query_mask(depth_u8: np.ndarray) -> np.ndarray:
    m = (depth_u8 != 0).astype(np.uint8)
    m = _box(m, 2, use_max=False)
    for _ in range(5):
        m = _box(m, 2, use_max=True)
    return m
