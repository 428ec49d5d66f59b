"""CPU: the frame-loop policy (pixtrack_b200/tracker.py) against the unmodified reference tracker.

tests/golden/tracker_policy.json was recorded by driving `PixLocPoseTrackerR9.run_single_frame`
(reference pixtrack/pose_trackers/pixloc_tracker_r9.py) over a scripted sequence of refinement outcomes
(tests/golden/gen/make_tracker_goldens.py); the same script is replayed here through B200PoseTracker with a
scripted engine and every recorded decision must match: masking, multiscale schedule, reference ids (used
and stored), the point set the dynamic reference was built from, cost threshold, success, carried pose,
relocalisation / miss counters and the record schema.  The pose conversions and the wrapper half of
get_nerf_image are pinned by the same fixture.
"""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from pixtrack_b200 import tracker as trk_mod
from pixtrack_b200.tracker import B200PoseTracker, PoseRt

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = json.load(open(os.path.join(HERE, 'golden', 'tracker_policy.json')))


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def pose_of(deg, t):
    return PoseRt(rot_y(np.deg2rad(deg)) @ rot_x(np.deg2rad(10.0)), t)


class ScriptedEngine:
    def __init__(self, frames):
        self.frames, self.i, self.events = frames, 0, []

    def query_camera(self, query_path):
        return 'query-camera'

    def create_reference(self, pose, ref_ids):
        return {'ref_ids': list(ref_ids), 'render_R': pose.R.copy()}

    def mask_query(self, image, pose):
        self.events[-1]['masked'] = True
        return image * 0.5

    def refine(self, query_path, image, camera, pose_init, ref_id, multiscale, features):
        f = self.frames[self.i]
        self.events[-1].update(multiscale=list(multiscale), run_ref_ids=[ref_id], pose_init_R=pose_init.R.tolist(),
                               pose_init_t=pose_init.t.tolist(), query_sum=float(np.asarray(image).sum()),
                               feature_ref_ids=features['ref_ids'], render_R=features['render_R'].tolist())
        costs = [np.float32(c[-1]) for c in f['costs']]
        if not f['ok']:
            return {'success': False, 'costs': costs}
        return {'success': True, 'T_refined': pose_of(f['deg'], f['t']), 'diff_R': 0.0, 'diff_t': 0.0, 'costs': costs}


def _run(tmp_path=None):
    sc = FIX['scenario']
    db = {int(k): v for k, v in sc['db'].items()}
    image_R = {k: rot_y(np.deg2rad(v['deg'])) @ rot_x(np.deg2rad(10.0)) for k, v in db.items()}
    image_t = {k: np.array(v['t'], float) for k, v in db.items()}
    covis = {int(k): {int(j): n for j, n in v.items()} for k, v in sc['covis'].items()}
    eng = ScriptedEngine(sc['frames'])
    tr = B200PoseTracker(eng, image_R, image_t, covis, sc['upright_ref'], eval_path=tmp_path)
    out = []
    for i, f in enumerate(sc['frames']):
        eng.i = i
        eng.events.append({'masked': False})
        n_before = len(tr.dynamic)
        tr.run_single_frame((f'/data/query/{f["name"]}', np.ones((4, 4, 3))))
        ret = tr.pose_history[f['name']]
        eng.events[-1].update(success=tr.success, pose_R=tr.pose.R.tolist(), pose_t=tr.pose.t.tolist(),
                              cost_threshold=tr.cost_threshold, reference_ids=list(tr.reference_ids),
                              relocalization_count=tr.relocalization_count, hits=tr.hits, misses=tr.misses,
                              new_dynamic=len(tr.dynamic) - n_before, ret_keys=sorted(ret.keys()), ret_success=ret['success'],
                              ret_reference_ids=list(ret['reference_ids']), ret_query_path=ret['query_path'],
                              ret_camera=ret['camera'],
                              tracker_costs_last=[c[-1] for c in tr.pose_tracker_history[f['name']].costs])
        out.append(eng.events[-1])
    return tr, out


def test_policy_decisions_match_the_reference_tracker_frame_by_frame():
    _, got = _run()
    assert len(got) == len(FIX['events']) == 9
    exact = ('masked', 'multiscale', 'run_ref_ids', 'reference_ids', 'feature_ref_ids', 'success', 'relocalization_count',
             'hits', 'misses', 'new_dynamic', 'ret_keys', 'ret_success', 'ret_reference_ids', 'ret_query_path', 'ret_camera')
    for i, (g, r) in enumerate(zip(got, FIX['events'])):
        for k in exact:
            assert g[k] == r[k], (i, k, g[k], r[k])
        for k in ('pose_R', 'pose_t', 'pose_init_R', 'pose_init_t', 'render_R', 'tracker_costs_last'):
            np.testing.assert_allclose(np.array(g[k], float), np.array(r[k], float), rtol=0, atol=1e-6, err_msg=f'{i} {k}')
        assert abs(g['cost_threshold'] - r['cost_threshold']) < 1e-6
        # the mask halves the frame in both scripts: 4*4*3 ones -> 24 when masked
        assert g['query_sum'] == r['query_sum'] == (24.0 if r['masked'] else 48.0)
    # the scenario covers: cold start, cost-threshold failure, LM failure, unmasked retry, reference hand-over 3 -> 4 -> 5
    assert [e['success'] for e in got] == [True, True, True, False, True, False, True, True, True]
    assert [e['run_ref_ids'][0] for e in got] == [3, 3, 3, 4, 4, 4, 4, 5, 5]


def test_dynamic_reference_store_is_bounded_and_outputs_are_written(tmp_path):
    tr, _ = _run(str(tmp_path))
    assert len(tr.dynamic) <= tr.max_dynamic and tr.dynamic_id in tr.dynamic
    tr.max_dynamic = 2
    tr.pose = pose_of(3.0, [0, 0, 2.0])
    for a in (4.0, 5.0, 6.0):
        tr.pose = pose_of(a, [0, 0, 2.0])
        tr.get_dynamic_id(tr.pose)
    assert len(tr.dynamic) == 2 and tr.dynamic_id in tr.dynamic
    tr.save_poses()
    poses = pickle.load(open(tmp_path / 'poses.pkl', 'rb'))
    logs = pickle.load(open(tmp_path / 'trackers.pkl', 'rb'))
    assert sorted(poses) == [f'f{i:03d}.jpg' for i in range(9)] == sorted(logs)
    ok = poses['f008.jpg']
    assert set(ok) == {'T_init', 'T_refined', 'camera', 'dbids', 'diff_R', 'diff_t', 'query_path', 'reference_ids', 'success'}
    R, t = ok['T_refined'].cpu().numpy()                        # how run_vis_on_poses.py / pose_utils.py:16-22 read it
    assert R.shape == (3, 3) and t.shape == (3,)
    assert 'T_refined' not in poses['f005.jpg'] and poses['f005.jpg']['success'] is False


def test_cache_hits_with_a_positive_threshold():
    """THRESH > 0 (the reference ships 0, r9.py:172): a pose within THRESH of a stored dynamic reference reuses it."""
    sc = FIX['scenario']
    db = {int(k): v for k, v in sc['db'].items()}
    image_R = {k: rot_y(np.deg2rad(v['deg'])) @ rot_x(np.deg2rad(10.0)) for k, v in db.items()}
    image_t = {k: np.array(v['t'], float) for k, v in db.items()}
    covis = {int(k): {int(j): n for j, n in v.items()} for k, v in sc['covis'].items()}
    eng = ScriptedEngine(sc['frames'])
    eng.events.append({})
    tr = B200PoseTracker(eng, image_R, image_t, covis, 3, thresh=np.deg2rad(5.0))
    tr.pose = pose_of(0.0, [0, 0, 2.0])
    a = tr.get_dynamic_id(tr.pose)
    tr.pose = pose_of(20.0, [0, 0, 2.0])
    b = tr.get_dynamic_id(tr.pose)                              # miss: new reference, ids move to image 4
    assert b != a and tr.misses == 1 and tr.reference_ids == [4]
    tr.pose = pose_of(22.0, [0, 0, 2.0])
    assert tr.get_dynamic_id(tr.pose) == b and tr.hits == 1     # 2 degrees from b
    tr.pose = pose_of(1.0, [0, 0, 2.0])
    assert tr.get_dynamic_id(tr.pose) == a and tr.hits == 2     # back near the first one: its stored ids apply


def test_pose_conversions_match_the_reference():
    for g in FIX['poses']:
        pose = PoseRt(g['R'], g['t'])
        cIw = trk_mod.camera_in_world_from_pose(pose)
        np.testing.assert_allclose(cIw, np.array(g['cIw']), atol=1e-12)
        n2s = dict(centroid=g['centroid'], avglen=g['avglen'], R=g['n2s_R'], totp=g['totp'])
        nerf = trk_mod.sfm_to_nerf_pose(n2s, cIw)
        np.testing.assert_allclose(nerf, np.array(g['nerf_pose']), atol=1e-12)
        # points transform like the camera centre
        c_sfm = cIw[:3, 3][None]
        np.testing.assert_allclose(trk_mod.sfm_to_nerf_points(n2s, c_sfm)[0], nerf[:3, 3], atol=1e-12)
    assert abs(trk_mod.geodesic_distance_for_rotations(rot_y(0.3), rot_y(0.1)) - 0.2) < 1e-12
    dR, dt = (pose_of(10, [0, 0, 1]).inv() @ pose_of(25, [0, 0, 1])).magnitude()
    assert 14.0 < dR < 16.0 and dt < 1.0


@pytest.mark.parametrize('case', FIX['nerf_image'], ids=['shade', 'depth', 'alpha_thresh'])
def test_get_nerf_image_wrapper_matches_the_reference_calls(case):
    """fov, camera matrix, render size / spp, render-mode toggling and the uint8 image, with the renderer replaced
    by a recorder that returns the fixture's float image (reference run_vis_on_poses.py:28-57)."""
    from types import SimpleNamespace
    from pixtrack_b200.nerf import RenderMode, get_nerf_image
    rgba = np.array(case['rgba'], np.float32)
    calls = []

    class TB:
        render_mode = RenderMode('Shade')
        fov = None

        def set_nerf_camera_matrix(self, m):
            calls.append(['cam', np.asarray(m).tolist()])

        def render_device(self, w, h, spp, want_rgba=True, want_u8=False, want_depth=False):
            calls.append(['render', w, h, spp, True, str(self.render_mode.name)])
            u8 = torch.from_numpy((rgba[:, :, :3] * 255.0).astype(np.uint8))     # the conversion the reference applies
            return torch.from_numpy(rgba), u8, None
    tb = TB()
    pose = np.eye(4)
    pose[:3, 3] = [0.1, 0.2, 0.3]
    cam = SimpleNamespace(size=np.array([case['W'], case['H']], np.float32), f=np.array([case['fl'], case['fl'] * 1.1], np.float32))
    img = get_nerf_image(tb, pose, cam, depth=case['depth'], alpha_thresh=case['alpha_thresh'])
    assert abs(tb.fov - case['fov']) < 1e-9
    assert calls == case['calls']
    assert tb.render_mode == case['mode_after']
    assert np.array_equal(img, np.array(case['image'], np.uint8))
