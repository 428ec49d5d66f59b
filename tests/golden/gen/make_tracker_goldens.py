"""Generate tests/golden/tracker_policy.json by driving the UNMODIFIED reference frame loop.

Run in the authoring container only (needs /root/reference):

    python tests/golden/gen/make_tracker_goldens.py

What is pinned (the host-side policy around the hot path):
  * `PixLocPoseTrackerR9.refine / relocalize / get_dynamic_id / update_reference_ids` and
    `PoseTracker.run_single_frame` (pixtrack/pose_trackers/pixloc_tracker_r9.py:95-266,
    base_pose_tracker.py:21-31): which frames are masked, the `multiscale` schedule, the cost
    threshold, success / pose carry-over, relocalisation count, reference-id choice, dynamic-reference
    bookkeeping and the `pose_history` record schema;
  * `sfm_to_nerf_pose` (pixtrack/utils/ingp_utils.py:47-63) and `get_camera_in_world_from_pixpose`
    (pixtrack/utils/pose_utils.py:16-27);
  * the wrapper half of `get_nerf_image` (pixtrack/visualization/run_vis_on_poses.py:28-57): fov, camera
    matrix, render arguments, render-mode toggling and the float -> uint8 conversion, against a recording
    fake testbed.

The tracker object is created without running its constructor (which needs COLMAP / HLoc / pyngp data on
disk); its attributes are set to what the constructor sets (r9.py:69-93), the localizer is a scripted fake
whose `run_query` replays the per-frame outcomes below through the real `DebugTracker`, and the renderer
calls (`get_reference_image`, `get_mask`, `get_query_camera`) are replaced by recorders.  Modules that are
not installed here (pycolmap, hloc, pyngp, ...) are satisfied by empty stand-ins; none of the pinned
functions touches them.
"""
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.abspath(os.path.join(HERE, '..'))
sys.path[:0] = [HERE, '/root/reference/pixloc', '/root/reference']
os.environ.setdefault('PROJECT_ROOT', '/root/reference')

MISSING = ('pycolmap', 'hloc', 'h5py', 'commentjson', 'pyngp', 'matplotlib', 'ycbvideo', 'pytorch3d', 'common',
           'scenes', 'pixsfm', 'plotly', 'open3d', 'trimesh')


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        return _Dummy()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, n):
        if n.startswith('__'):
            raise AttributeError(n)
        return _Dummy


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in MISSING:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _Finder())
# ingp_utils.py gets `np` through `from common import *` (instant-ngp/scripts/common.py imports numpy as np and
# needs imageio, absent here): the stand-in re-exports exactly that name
_common = types.ModuleType('common')
_common.np = np
sys.modules['common'] = _common
import torch  # noqa: E402

_six = types.ModuleType('torch._six')
_six.string_classes = (str, bytes)
sys.modules['torch._six'] = _six

from pixloc.pixlib.geometry import Pose  # noqa: E402
from pixtrack.pose_trackers.pixloc_tracker_r9 import PixLocPoseTrackerR9  # noqa: E402
from pixtrack.utils.ingp_utils import sfm_to_nerf_pose  # noqa: E402
from pixtrack.utils.pose_utils import get_camera_in_world_from_pixpose  # noqa: E402
from pixtrack.visualization.run_vis_on_poses import get_nerf_image  # noqa: E402


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


# ---------------------------------------------------------------------------------------------------------
# scenario (stored in the fixture so the test replays exactly these inputs)
# ---------------------------------------------------------------------------------------------------------
SCENARIO = dict(
    upright_ref=3,
    # database images: rotation = rot_y(deg) @ rot_x(10 deg), tvec
    db={1: dict(deg=-40.0, t=[0.1, 0.0, 2.0]), 2: dict(deg=-20.0, t=[0.0, 0.1, 2.1]), 3: dict(deg=0.0, t=[0.0, 0.0, 2.0]),
        4: dict(deg=18.0, t=[-0.1, 0.0, 2.2]), 5: dict(deg=35.0, t=[0.0, -0.1, 1.9]), 6: dict(deg=60.0, t=[0.2, 0.0, 2.0])},
    # covisibility counts (only neighbours with > 50 shared points are candidates, r9.py:131-132)
    covis={1: {2: 400, 3: 60}, 2: {1: 400, 3: 300, 4: 51}, 3: {2: 300, 4: 280, 5: 50, 1: 60}, 4: {3: 280, 5: 200, 2: 51},
           5: {4: 200, 6: 90, 3: 50}, 6: {5: 90}},
    # per frame: what the refinement returns.  `costs` = per pyramid level (coarse first) the per-iteration
    # mean costs the optimizer logs; deg / t = the refined pose
    frames=[
        dict(name='f000.jpg', ok=True, deg=2.0, t=[0.0, 0.0, 2.0], costs=[[5.0, 3.0, 2.0], [1.5, 1.2], [1.0, 0.9]]),
        dict(name='f001.jpg', ok=True, deg=9.0, t=[0.01, 0.0, 2.0], costs=[[2.1, 1.9], [1.2], [0.85]]),
        dict(name='f002.jpg', ok=True, deg=15.0, t=[0.02, 0.0, 2.0], costs=[[2.0], [1.3], [0.95]]),
        dict(name='f003.jpg', ok=True, deg=50.0, t=[0.5, 0.0, 2.0], costs=[[4.0, 3.9], [2.5], [2.0]]),   # cost too high
        dict(name='f004.jpg', ok=True, deg=22.0, t=[0.03, 0.0, 2.0], costs=[[2.2, 1.8], [1.1], [0.9]]),
        dict(name='f005.jpg', ok=False, deg=0.0, t=[0.0, 0.0, 0.0], costs=[[3.0, 2.9]]),                  # LM failed at level 0
        dict(name='f006.jpg', ok=True, deg=30.0, t=[0.04, 0.0, 2.0], costs=[[1.9], [1.25], [0.99]]),
        dict(name='f007.jpg', ok=True, deg=41.0, t=[0.05, 0.0, 2.0], costs=[[1.9], [1.25], [1.3]]),
        dict(name='f008.jpg', ok=True, deg=47.0, t=[0.05, 0.0, 2.0], costs=[[1.0], [1.0], [1.0]]),
    ],
)


def db_rotation(entry):
    return rot_y(np.deg2rad(entry['deg'])) @ rot_x(np.deg2rad(10.0))


def run_policy():
    sc = SCENARIO
    events = []

    class DbImage:
        def __init__(self, e):
            self.R, self.tvec = db_rotation(e), np.array(e['t'], float)

        def qvec2rotmat(self):
            return self.R

    class Opt:
        logging_fn = None

    refiner = SimpleNamespace(conf=SimpleNamespace(multiscale=[1]), features_dicts={}, optimizer=[Opt(), Opt(), Opt()],
                              reference_scale=0.5)
    refiner.extract_reference_features = lambda ref_ids, pose, img: {'ref_ids': list(ref_ids), 'img': img}
    cursor = dict(i=0)

    def run_query(query_path, camera, pose_init, ref_ids, image_query=None, pose=None, reference_images_raw=None,
                  dynamic_id=None):
        f = sc['frames'][cursor['i']]
        events[-1].update(multiscale=list(refiner.conf.multiscale), run_ref_ids=[int(r) for r in ref_ids],
                          pose_init_R=pose_init.numpy()[0].tolist(), pose_init_t=pose_init.numpy()[1].tolist(),
                          query_sum=float(np.asarray(image_query, float).sum()),
                          dynamic_is_current=bool(dynamic_id == tr.dynamic_id))
        T_ref = Pose.from_Rt(torch.tensor(rot_y(np.deg2rad(f['deg'])) @ rot_x(np.deg2rad(10.0))), torch.tensor(f['t'], dtype=torch.float64))
        # replay the optimizer log through whatever tracker object is attached (DebugTracker, r9.py:238)
        for lv, cs in enumerate(f['costs']):
            for i, c in enumerate(cs):
                refiner.optimizer[lv].logging_fn(i=i, T_init=pose_init, T=T_ref, T_delta=Pose.from_Rt(torch.eye(3), torch.zeros(3)),
                                                 cost=torch.tensor([c, c * 3.0]), valid=torch.tensor([True, False]))
        if not f['ok']:
            return {'success': False, 'T_init': pose_init, 'dbids': list(ref_ids)}
        return {'success': True, 'T_init': pose_init, 'T_refined': T_ref, 'diff_R': 0.0, 'diff_t': 0.0, 'dbids': list(ref_ids)}

    localizer = SimpleNamespace(refiner=refiner, run_query=run_query,
                                model3d=SimpleNamespace(dbs={k: DbImage(v) for k, v in sc['db'].items()}))
    tr = object.__new__(PixLocPoseTrackerR9)
    # what the constructor sets (r9.py:57-93)
    tr.debug = 1
    tr.localizer = localizer
    tr.eval_path = '/tmp'
    tr.covis = sc['covis']
    tr.pose_history, tr.pose_tracker_history = {}, {}
    tr.cold_start, tr.pose = True, None
    tr.reference_ids = [sc['upright_ref']]
    tr.reference_scale = 0.5
    tr.dynamic_id = None
    tr.hits = tr.misses = 0
    tr.cache_hit = False
    tr.cost_threshold = None
    tr.relocalization_count = 0
    tr.success = True
    tr.pbar = SimpleNamespace(set_description=lambda m: None)
    # renderer-facing calls -> recorders
    tr.get_query_camera = lambda q: 'query-camera'
    tr.get_reference_image = lambda pose: ('render', pose.numpy()[0].copy())

    def get_mask(pose):
        events[-1]['masked'] = True
        return np.full((4, 4, 1), 0.5)
    tr.get_mask = get_mask

    for i, f in enumerate(sc['frames']):
        cursor['i'] = i
        events.append(dict(masked=False))
        n_dyn_before = len(refiner.features_dicts)
        tr.run_single_frame((f'/data/query/{f["name"]}', np.ones((4, 4, 3))))
        R, t = tr.pose.numpy()
        ret = tr.pose_history[f['name']]
        events[-1].update(success=bool(tr.success), pose_R=np.asarray(R).tolist(), pose_t=np.asarray(t).tolist(),
                          cost_threshold=float(tr.cost_threshold), reference_ids=[int(r) for r in tr.reference_ids],
                          relocalization_count=int(tr.relocalization_count), hits=int(tr.hits), misses=int(tr.misses),
                          new_dynamic=len(refiner.features_dicts) - n_dyn_before, ret_keys=sorted(ret.keys()),
                          feature_ref_ids=[int(r) for r in refiner.features_dicts[tr.dynamic_id]['features']['ref_ids']],
                          render_R=np.asarray(refiner.features_dicts[tr.dynamic_id]['features']['img'][1]).tolist(),
                          ret_success=bool(ret['success']), ret_reference_ids=[int(r) for r in ret['reference_ids']],
                          ret_query_path=ret['query_path'], ret_camera=ret['camera'],
                          tracker_costs_last=[float(c[-1]) for c in tr.pose_tracker_history[f['name']].costs])
    return events


def pose_goldens():
    rng = np.random.default_rng(5)
    out = []
    for _ in range(4):
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] *= -1
        n2s = dict(centroid=rng.normal(size=3), avglen=float(rng.uniform(1.0, 4.0)), R=np.eye(4), totp=rng.normal(size=3))
        rr, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(rr) < 0:
            rr[:, 0] *= -1
        n2s['R'][:3, :3] = rr
        R = q
        t = rng.normal(size=3)
        pose = Pose.from_Rt(torch.tensor(R), torch.tensor(t))
        cIw = get_camera_in_world_from_pixpose(pose)
        nerf = sfm_to_nerf_pose(n2s, cIw.copy())
        out.append(dict(R=R.tolist(), t=t.tolist(), centroid=n2s['centroid'].tolist(), avglen=n2s['avglen'], n2s_R=n2s['R'].tolist(),
                        totp=n2s['totp'].tolist(), cIw=np.asarray(cIw).tolist(), nerf_pose=np.asarray(nerf).tolist()))
    return out


def nerf_image_goldens():
    """get_nerf_image against a recording testbed whose render returns a fixed float image."""
    rng = np.random.default_rng(9)
    out = []
    # rgba in [0, 1] like a real render (alpha is never negative, so the `alpha < 0.0` clear of the default
    # alpha_thresh cannot fire); the third case pins the thresholding branch
    for depth, thresh in ((False, 0.0), (True, 0.0), (False, 0.5)):
        W, H, fl = 7, 5, 11.5
        rgba = rng.uniform(0.0, 1.0, size=(H, W, 4)).astype(np.float32)
        calls = []

        class Mode:                      # like a pybind enum: members are reachable from a member
            def __init__(self, name):
                self.name = name

            def __str__(self):
                return self.name
        Mode.Depth, Mode.Shade = Mode('Depth'), Mode('Shade')

        class TB:
            render_mode = Mode.Shade
            fov = None

            def set_nerf_camera_matrix(self, m):
                calls.append(('cam', np.asarray(m).tolist()))

            def render(self, w, h, spp, linear):
                calls.append(('render', w, h, spp, bool(linear), str(self.render_mode)))
                return rgba.copy()
        tb = TB()
        pose = np.eye(4)
        pose[:3, 3] = [0.1, 0.2, 0.3]
        cam = SimpleNamespace(size=np.array([W, H], np.float32), f=np.array([fl, fl * 1.1], np.float32))
        with np.errstate(invalid='ignore'):
            img = get_nerf_image(tb, pose, cam, depth=depth, alpha_thresh=thresh)
        out.append(dict(depth=depth, alpha_thresh=thresh, W=W, H=H, fl=fl, rgba=rgba.tolist(), fov=float(tb.fov), calls=calls,
                        mode_after=str(tb.render_mode), image=img.tolist()))
    return out


if __name__ == '__main__':
    fixture = dict(scenario=SCENARIO, events=run_policy(), poses=pose_goldens(), nerf_image=nerf_image_goldens())
    # JSON keys must be strings
    fixture['scenario'] = json.loads(json.dumps(SCENARIO, default=str))
    path = os.path.join(OUT, 'tracker_policy.json')
    with open(path, 'w') as f:
        json.dump(fixture, f, indent=1)
    print('wrote', path, os.path.getsize(path), 'bytes')
    for e in fixture['events']:
        print({k: e[k] for k in ('success', 'masked', 'multiscale', 'run_ref_ids', 'reference_ids', 'cost_threshold',
                                 'relocalization_count', 'hits', 'misses', 'new_dynamic')})
