"""Result overlay (pixtrack_b200/overlay.py, csrc/ptk_overlay.cu) against what the unmodified reference code produced
(tests/golden/overlay.npz from tests/golden/gen/make_overlay_goldens.py: blend_images, add_pose_axes / draw_axes of
pixtrack/visualization/run_vis_on_poses.py with OpenCV doing the drawing)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import cases


def test_axis_end_points_match_the_reference():
    from pixtrack_b200.overlay import pose_axes_points
    g = cases.gold('overlay')
    cam = SimpleNamespace(size=g['cam_size'], f=g['cam_f'])
    pts = pose_axes_points(cam, g['pose'], g['centre'].tolist() + [0])
    assert pts.dtype == np.int16 and np.array_equal(pts, g['axes_px'])
    assert [tuple(c) for c in g['colors']] == [(255, 0, 0), (0, 255, 0), (0, 0, 255)] and set(g['thickness']) == {2}


@pytest.mark.gpu
def test_blend_is_exact_and_axes_cover_what_opencv_draws():
    from pixtrack_b200.overlay import overlay, pose_axes_points
    g = cases.gold('overlay')
    d = 'cuda:0'
    q, n = torch.from_numpy(g['query']).to(d), torch.from_numpy(g['nerf']).to(d)
    blend = overlay(q, n, alpha=0.3)
    torch.cuda.synchronize()
    assert np.array_equal(blend.cpu().numpy(), g['blend'])                     # float64 blend + truncation, bit for bit
    white = overlay(q, None, alpha=0.3)                                        # frames without a pose: white render
    assert np.array_equal(white.cpu().numpy(), (g['query'] * 0.3 + 255.0 * 0.7).astype(np.uint8))
    cam = SimpleNamespace(size=g['cam_size'], f=g['cam_f'])
    pts = pose_axes_points(cam, g['pose'], g['centre'].tolist() + [0])
    got = overlay(q, n, alpha=0.3, axes_px=pts, thickness=2).cpu().numpy()
    want = g['with_axes']
    drawn_ref = (want != g['blend']).any(-1)
    drawn = (got != g['blend']).any(-1)
    # outside the lines nothing changes; on the lines the colours are the reference's; the rasterised regions agree up
    # to the one-pixel boundary of OpenCV's polygon scan conversion
    assert np.array_equal(got[~drawn], g['blend'][~drawn])
    both = drawn & drawn_ref
    assert both.sum() >= 0.85 * drawn_ref.sum() and both.sum() >= 0.7 * drawn.sum(), (both.sum(), drawn_ref.sum(), drawn.sum())
    assert (got[both] == want[both]).all(-1).mean() > 0.9                       # overlapping axes near the origin may differ in order
    assert (drawn != drawn_ref).mean() < 0.01
