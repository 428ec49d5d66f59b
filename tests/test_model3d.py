"""CPU: array-table point selection / reference choice against the dict-and-loop oracle on a synthetic SfM model."""
import numpy as np
import pytest

from oracle import model3d as om
from pixtrack_b200.model3d import PointTables


def _rot(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    return q * np.sign(np.linalg.det(q))


def _model(seed, n_img=14, n_pts=400):
    rng = np.random.default_rng(seed)
    image_ids = list(rng.permutation(np.arange(3, 3 + 2 * n_img, 2))[:n_img])      # non-contiguous ids
    point_ids = list(rng.permutation(np.arange(10, 10 + 3 * n_pts, 3))[:n_pts])
    tracks = {}
    for p in point_ids:
        k = int(rng.integers(1, 7))
        tracks[int(p)] = [int(i) for i in rng.choice(image_ids, size=k, replace=False)]
    obs = {int(i): [] for i in image_ids}
    for p, tr in tracks.items():
        for i in tr:
            obs[i].append(p)
    img_pts = {}
    for i, ps in obs.items():
        arr = np.array(ps + [-1] * int(rng.integers(0, 30)), dtype=np.int64)
        img_pts[i] = rng.permutation(arr)
    R = {int(i): _rot(rng) for i in image_ids}
    xyz = {int(p): rng.normal(size=3) for p in point_ids}
    return img_pts, R, xyz, tracks, rng


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_tables_match_the_reference_loops(seed):
    img_pts, R, xyz, tracks, rng = _model(seed)
    tab = PointTables(img_pts, R, xyz, tracks)
    covis = om.extract_covisibility(img_pts, tracks)
    for i in img_pts:
        ref = om.p3did_to_dbids(img_pts, tracks, [i])
        ids, pts = tab.points_of_image(i)
        assert list(ids) == list(ref.keys())                              # same points, same (insertion) order
        assert np.array_equal(pts, np.stack([xyz[int(p)] for p in ids]) if len(ids) else np.zeros((0, 3)))
        assert tab.covisible(i) == covis.get(i, {})
    two = list(img_pts)[:2]
    ids2, _ = tab.points_of_images(two)
    assert list(ids2) == list(om.p3did_to_dbids(img_pts, tracks, two).keys())
    for _ in range(20):
        Rq = _rot(rng)
        cur = int(rng.choice(list(img_pts)))
        for N in (0, 3, 50):
            assert tab.nearest_reference(Rq, cur, min_covis=N) == om.update_reference_ids(covis, R, Rq, cur, N=N)
        assert tab.nearest_reference(Rq, cur, min_covis=1, K=3) == om.update_reference_ids(covis, R, Rq, cur, N=1, K=3)


def test_edge_cases():
    img_pts = {5: np.array([-1, -1]), 7: np.array([1, 2, 2, -1]), 9: np.array([2])}
    tracks = {1: [7], 2: [7, 7, 9]}                    # point 2 seen twice by image 7
    R = {5: np.eye(3), 7: np.eye(3), 9: np.eye(3)}
    xyz = {1: np.zeros(3), 2: np.ones(3)}
    tab = PointTables(img_pts, R, xyz, tracks, min_track_length=3)
    assert list(tab.points_of_image(5)[0]) == [] and tab.covisible(5) == {}
    assert list(tab.points_of_image(7)[0]) == [2]      # point 1 has a track of length 1
    covis = om.extract_covisibility(img_pts, tracks)
    assert tab.covisible(7) == covis[7] and tab.covisible(9) == covis[9]
    assert tab.nearest_reference(np.eye(3), 5) == [5]


def test_tracker_built_from_tables_chooses_references_like_the_tables():
    """B200PoseTracker.from_tables: its dict-based update_reference_ids (the pinned restatement of r9.py:120-143)
    and PointTables.nearest_reference agree on a synthetic SfM model."""
    from pixtrack_b200.tracker import B200PoseTracker, PoseRt
    img_pts, R, xyz, tracks, rng = _model(4)
    tab = PointTables(img_pts, R, xyz, tracks)
    assert tab.max_points == max(len(tab.points_of_image(i)[0]) for i in img_pts)
    trk = B200PoseTracker.from_tables(None, tab, {i: np.zeros(3) for i in img_pts}, int(list(img_pts)[0]), min_covis=3)
    for _ in range(20):
        cur = int(rng.choice(list(img_pts)))
        trk.reference_ids, trk.cache_hit = [cur], False
        trk.pose = PoseRt(_rot(rng), np.zeros(3))
        assert trk.update_reference_ids() == tab.nearest_reference(trk.pose.R, cur, min_covis=3)
