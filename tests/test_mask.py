"""Query mask: oracle pinned on OpenCV (the reference's own calls), CUDA path against both."""
import numpy as np
import pytest
import torch

from oracle import mask as omask


def _depth(seed, H=90, W=140):
    rng = np.random.default_rng(seed)
    d = np.zeros((H, W, 3), np.uint8)
    yy, xx = np.mgrid[:H, :W]
    for _ in range(4):                                   # blobs, one touching the border, plus speckle
        cy, cx, r = rng.integers(0, H), rng.integers(0, W), rng.integers(4, 25)
        d[((yy - cy) ** 2 + (xx - cx) ** 2) < r * r] = rng.integers(1, 255)
    sp = rng.random((H, W)) < 0.01
    d[sp] = 200
    d[:3, :40] = 9
    return d


def _cv2_mask(depth):
    import cv2
    kernel = np.ones((5, 5), np.uint8)                   # r9.py:211-213, verbatim
    img_erosion = cv2.erode((depth != 0).astype(np.uint8), kernel, iterations=1)
    return cv2.dilate(img_erosion, kernel, iterations=5)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_oracle_matches_opencv(seed):
    pytest.importorskip('cv2')
    d = _depth(seed)
    assert np.array_equal(omask.query_mask(d), _cv2_mask(d))


def test_oracle_edge_cases():
    assert not omask.query_mask(np.zeros((20, 30, 3), np.uint8)).any()
    assert omask.query_mask(np.full((20, 30, 3), 7, np.uint8)).all()     # the border does not erode
    d = np.zeros((40, 40, 3), np.uint8)
    d[20, 20] = 1                                                        # a speck smaller than the kernel vanishes
    assert not omask.query_mask(d).any()
    d[10:15, 10:15] = 1                                                  # a 5x5 block survives as one pixel -> 21x21
    m = omask.query_mask(d)
    assert m[..., 0].sum() == 21 * 21 and m[12, 12, 0] == 1


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.uint8, torch.float32])
def test_cuda_mask_matches_oracle_and_multiplies_the_query(dtype):
    from pixtrack_b200.mask import query_mask
    for seed, (H, W) in enumerate([(90, 140), (33, 47), (270, 480)]):
        d = _depth(seed, H, W)
        ref = omask.query_mask(d)
        img = torch.from_numpy(np.random.default_rng(seed).integers(0, 256, (H, W, 3)).astype(np.uint8)).to(dtype)
        out, m = query_mask(torch.from_numpy(d).cuda(), img.cuda(), want_mask=True)
        torch.cuda.synchronize()
        assert np.array_equal(m.cpu().numpy(), ref)
        assert torch.equal(out.cpu(), img * torch.from_numpy(ref).to(dtype))     # r9.py:225
    only, m2 = query_mask(torch.from_numpy(d).cuda(), None, want_mask=True)
    assert only is None and np.array_equal(m2.cpu().numpy(), ref)


@pytest.mark.gpu
def test_cuda_mask_full_hd_properties():
    from pixtrack_b200.mask import query_mask
    d = torch.zeros((1080, 1920, 3), dtype=torch.uint8, device='cuda')
    d[400:700, 800:1200] = 50
    img = torch.full((1080, 1920, 3), 255, dtype=torch.uint8, device='cuda')
    out, m = query_mask(d, img, want_mask=True)
    # erode by 2 then dilate by 10: the rectangle grows by 8 pixels on every side
    assert int(m[..., 0].sum()) == (300 + 16) * (400 + 16)
    assert bool((out[392:708, 792:1208] == 255).all()) and int(out.sum()) == 255 * 3 * 316 * 416
    out2, _ = query_mask(m * 255, out)                   # idempotent on its own support: the mask only grows
    assert bool((out2 == out).all())
