"""Generate tests/golden/model3d.json from the UNMODIFIED reference point-selection code.

Run in the authoring container only (needs /root/reference):

    python tests/golden/gen/make_model3d_goldens.py

Pinned: `Model3D.get_p3did_to_dbids` (point_selection='all', min_track_length 3 and 2) and `get_dbid_to_p3dids`
(pixloc/pixloc/localization/model3d.py:41-87), `extract_covisibility` (pixtrack/utils/hloc_utils.py:28-47) on synthetic
COLMAP-shaped models (namedtuples with the fields the reference reads: Image.point3D_ids, Point3D.image_ids / xyz).
The Model3D constructor (reads a COLMAP model from disk) is bypassed; `hloc.read_model`, absent here, is replaced by a
function returning the same synthetic model.  The fixture stores the models and what the reference returned,
INCLUDING the key order of the returned dicts (the point order downstream code stacks xyz in).
"""
import collections
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.abspath(os.path.join(HERE, '..'))
sys.path[:0] = [HERE, '/root/reference/pixloc', '/root/reference']
MISSING = ('pycolmap', 'hloc', 'h5py', 'commentjson', 'pyngp', 'matplotlib', 'ycbvideo', 'pytorch3d', 'common', 'scenes',
           'pixsfm', 'plotly', 'open3d', 'trimesh', 'tqdm_missing')


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        return _Dummy()


class _Stub(types.ModuleType):
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
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _Finder())
import torch  # noqa: E402,F401
_six = types.ModuleType('torch._six')
_six.string_classes = (str, bytes)
sys.modules['torch._six'] = _six

from pixloc.localization.model3d import Model3D  # noqa: E402
from pixtrack.utils import hloc_utils  # noqa: E402

Image = collections.namedtuple('Image', ['id', 'qvec', 'tvec', 'camera_id', 'name', 'xys', 'point3D_ids'])
Point3D = collections.namedtuple('Point3D', ['id', 'xyz', 'rgb', 'error', 'image_ids', 'point2D_idxs'])


def synthetic_model(seed, n_img=14, n_pts=400):
    """Same generator as tests/test_model3d.py::_model (non-contiguous ids, tracks of 1..6 images, unmatched keypoints,
    shuffled keypoint order) plus duplicate observations of a point in one image."""
    rng = np.random.default_rng(seed)
    image_ids = [int(i) for i in rng.permutation(np.arange(3, 3 + 2 * n_img, 2))[:n_img]]
    point_ids = [int(p) for p in rng.permutation(np.arange(10, 10 + 3 * n_pts, 3))[:n_pts]]
    tracks, obs = {}, {i: [] for i in image_ids}
    for p in point_ids:
        k = int(rng.integers(1, 7))
        tr = [int(i) for i in rng.choice(image_ids, size=k, replace=False)]
        if rng.random() < 0.1:
            tr.append(tr[0])                       # the same image observes the point twice
        tracks[p] = tr
        for i in tr:
            obs[i].append(p)
    img_pts = {i: [int(x) for x in rng.permutation(np.array(ps + [-1] * int(rng.integers(0, 30)), dtype=np.int64))]
               for i, ps in obs.items()}
    return dict(seed=seed, image_point3D_ids=img_pts, tracks=tracks)


def run_reference(model):
    images = {i: Image(i, None, None, 1, f'{i}.jpg', None, np.array(ids, dtype=np.int64))
              for i, ids in model['image_point3D_ids'].items()}
    points = {p: Point3D(p, np.zeros(3), None, 0.0, np.array(tr, dtype=np.int64), None) for p, tr in model['tracks'].items()}
    m3d = Model3D.__new__(Model3D)
    m3d.cameras, m3d.dbs, m3d.points3D = {}, images, points
    out = {'single': {}, 'pairs': [], 'dbid_to_p3dids': {}}
    ids = list(images)
    for i in ids:
        for mtl in (3, 2):
            sel = m3d.get_p3did_to_dbids([i], None, None, 'all', mtl)
            out['single'][f'{i}:{mtl}'] = [int(k) for k in sel.keys()]
    for a, b in ((ids[0], ids[1]), (ids[2], ids[5]), (ids[3], ids[3])):
        sel = m3d.get_p3did_to_dbids([a, b], None, None, 'all', 3)
        out['pairs'].append(dict(dbids=[a, b], p3dids=[int(k) for k in sel.keys()],
                                 dbids_of_point={str(int(k)): sorted(int(x) for x in v) for k, v in sel.items()}))
        d2p = m3d.get_dbid_to_p3dids(sel)
        out['dbid_to_p3dids'][f'{a},{b}'] = {str(int(k)): [int(x) for x in v] for k, v in d2p.items()}
    hloc_utils.read_model = lambda path: ({}, images, points)
    hloc_utils.tqdm = types.SimpleNamespace(tqdm=lambda x: x)
    covis = hloc_utils.extract_covisibility('unused')
    out['covis'] = {str(int(i)): {str(int(j)): int(n) for j, n in c.items()} for i, c in covis.items()}
    return out


if __name__ == '__main__':
    cases = []
    for seed in (0, 1, 2, 3):
        model = synthetic_model(seed)
        ref = run_reference(model)
        cases.append(dict(model={'seed': seed,
                                 'image_point3D_ids': {str(k): v for k, v in model['image_point3D_ids'].items()},
                                 'tracks': {str(k): v for k, v in model['tracks'].items()}}, reference=ref))
    path = os.path.join(OUT, 'model3d.json')
    with open(path, 'w') as f:
        json.dump(dict(generator='tests/golden/gen/make_model3d_goldens.py', cases=cases), f)
    print('wrote', path, os.path.getsize(path), 'bytes')
