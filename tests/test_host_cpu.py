"""CPU: host-side logic and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'pixtrack_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(ptk_[a-z0-9_]+)\s*\(', hdr)))


def test_library_loads_and_exports_every_declared_symbol():
    from pixtrack_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/pixtrack_b200.h but not exported'
        assert n in _lib.SYMBOLS, f'{n} has no ctypes prototype in _lib.py'
    assert lib.ptk_abi_version() == _lib.ABI_VERSION == 2


def test_struct_layout_matches_header():
    from pixtrack_b200 import _lib
    # 10 int32, 9 (pointer, int64) pairs, 1 pointer, 4 floats, workspace pointer + size
    assert ctypes.sizeof(_lib.LmProblem) == 40 + 9 * 16 + 8 + 16 + 16
    assert _lib.LmProblem.workspace.offset == 40 + 9 * 16 + 8 + 16
    assert _lib.LmProblem.p3d.offset == 40 and _lib.LmProblem.skip.offset == 40 + 9 * 16
    assert ctypes.sizeof(_lib.LmResult) == 32


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_path_fails_loudly_without_cuda():
    from pixtrack_b200 import _lib
    from pixtrack_b200.geometry import Camera, Pose
    from pixtrack_b200.optimizer import B200Optimizer
    from pixtrack_b200.sampling import sample_points
    with pytest.raises(_lib.PtkError):
        _lib.context(0)
    opt = B200Optimizer(dict(num_iters=3, pad=1, loss_fn='scaled_barron(0, 0.1)'))
    with pytest.raises(_lib.PtkError):
        opt.run(np.zeros((20, 3)), torch.zeros(20, 16), torch.zeros(16, 8, 8), Pose(torch.zeros(12)),
                Camera(torch.ones(8)))
    with pytest.raises(_lib.PtkError):
        sample_points(torch.zeros(4, 8, 8), torch.zeros(3, 2))


def test_context_creation_failure_raises_instead_of_deadlocking(monkeypatch):
    """ptk_create failing (wrong architecture, bad index, allocation failure) must surface as PtkError: check() re-enters
    load(), which takes the same non-reentrant lock context() holds while creating."""
    import threading
    from pixtrack_b200 import _lib
    lib = _lib.load()

    class FailingLib:
        def __getattr__(self, name):
            return getattr(lib, name)

        @staticmethod
        def ptk_create(device, out):
            return -3

    monkeypatch.setattr(_lib, '_lib', FailingLib())
    monkeypatch.setattr(_lib, '_contexts', {})
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    result = []

    def attempt():
        try:
            _lib.context(0)
            result.append('returned')
        except _lib.PtkError as e:
            result.append(e)
    th = threading.Thread(target=attempt, daemon=True)
    th.start()
    th.join(20)
    assert not th.is_alive(), 'context() deadlocked on its own lock'
    assert len(result) == 1 and isinstance(result[0], _lib.PtkError)
    assert 0 not in _lib._contexts


def test_camera_world2image_matches_the_reference_fixture():
    """Camera.world2image of the product (host side) on the points / cameras the unmodified reference was run on
    (tests/golden/geometry.npz: pinhole, radial +/-, tangential)."""
    import cases
    from pixtrack_b200.geometry import Camera
    g = cases.gold('geometry')
    pc = torch.from_numpy(g['pc'])
    for tag, dist in (('d0', []), ('d2', [0.1, 0.01]), ('d2n', [-0.2, 0.05]), ('d4', [0.15, -0.3, 0.002, -0.001])):
        cam = Camera(torch.tensor([640., 480., 300., 350., 320., 240.] + dist))
        uv, valid = cam.world2image(pc)
        assert np.array_equal(valid.numpy(), g[f'{tag}_valid'])
        np.testing.assert_allclose(uv.numpy()[g[f'{tag}_valid']], g[f'{tag}_uv'][g[f'{tag}_valid']], rtol=2e-6, atol=2e-4)
        uv64, valid64 = Camera(cam._data.double()).world2image(pc.double().numpy())      # numpy in, float64 chain
        assert uv64.dtype == torch.float64 and np.array_equal(valid64.numpy(), g[f'{tag}_valid'])


def test_morton_order_is_a_locality_preserving_permutation():
    from pixtrack_b200.pipeline import morton_order
    g = torch.Generator().manual_seed(0)
    x = torch.rand(4096, 3, generator=g, dtype=torch.float64)
    o = morton_order(x)
    assert sorted(o.tolist()) == list(range(4096))
    y = x[o]
    # neighbours along the curve are close in space: mean step far below the mean distance of random pairs (0.66)
    assert float((y[1:] - y[:-1]).norm(dim=1).mean()) < 0.15
    assert float((x[1:] - x[:-1]).norm(dim=1).mean()) > 0.5
    assert morton_order(torch.zeros(0, 3)).numel() == 0 and morton_order(torch.ones(5, 3)).tolist() == [0, 1, 2, 3, 4]


def test_product_package_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'pixtrack_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_conf_merge_and_damping():
    from pixtrack_b200.optimizer import B200Optimizer, parse_loss
    opt = B200Optimizer(dict(num_iters=150, pad=1, loss_fn='scaled_barron(0, 0.1)'))   # r9.py:46-49 style
    assert opt.conf.num_iters == 150 and opt.conf.interpolation.pad == 1 and opt.interpolator.pad == 1
    assert opt.conf.grad_stop_criteria == 1e-4 and opt.conf.dt_stop_criteria == 5e-3
    assert parse_loss('scaled_barron(0, 0.1)') == 0.1
    with pytest.raises(NotImplementedError):
        parse_loss('squared_loss')
    lam = opt.dampingnet()
    np.testing.assert_allclose(lam.detach().numpy(), np.full(6, 10 ** -0.5), rtol=1e-6)
    assert 'dampingnet.const' in opt.state_dict()       # checkpoint key optimizer.{l}.dampingnet.const


def test_geometry_wrappers():
    from pixtrack_b200.geometry import Camera, Pose
    cam = Camera.from_colmap(dict(model='SIMPLE_RADIAL', width=1920, height=1080,
                                  params=np.array([2304., 960., 540., 0.01])))
    np.testing.assert_allclose(cam._data.numpy(), [1920, 1080, 2304, 2304, 959.5, 539.5, 0.01, 0.0])
    s = cam.scale((0.5, 0.25))
    np.testing.assert_allclose(s._data.numpy(), [960, 270, 1152, 576, 479.5, 134.5, 0.01, 0.0])
    T = Pose.from_aa(torch.tensor([0.1, -0.2, 0.3]), torch.tensor([1., 2., 3.]))
    I = (T.inv() @ T)._data
    np.testing.assert_allclose(I.numpy(), np.r_[np.eye(3).ravel(), 0, 0, 0], atol=1e-6)
    dr, dt = T.magnitude()
    np.testing.assert_allclose(float(dr), np.degrees(np.linalg.norm([0.1, -0.2, 0.3])), rtol=1e-5)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_install_and_adapters_fail_loudly_without_cuda():
    """install() on a reference-shaped localizer must raise instead of leaving the PyTorch path in place."""
    from types import SimpleNamespace
    import synthetic as syn
    from pixtrack_b200 import _lib
    from pixtrack_b200.install import install
    from pixtrack_b200.nerf import NerfTestbed
    from pixtrack_b200.mask import query_mask

    class RefOpt(torch.nn.Module):                       # what BaseOptimizer exposes (base_optimizer.py:23-60)
        def __init__(self):
            super().__init__()
            self.conf = dict(num_iters=150, loss_fn='scaled_barron(0, 0.1)', jacobi_scaling=False, normalize_features=False,
                             grad_stop_criteria=1e-4, dt_stop_criteria=5e-3, dR_stop_criteria=5e-2,
                             interpolation=dict(mode='linear', pad=1), damping=dict(type='constant', log_range=[-6, 5]))
            self.dampingnet = SimpleNamespace(const=torch.nn.Parameter(torch.zeros(6)))
            self.logging_fn = None
    unet = torch.nn.Module()
    unet.state_dict = lambda: syn.unet_weights(0)
    ref_ext = SimpleNamespace(conf=SimpleNamespace(resize=1024, resize_by='max'), model=unet, device=torch.device('cpu'))
    loc = SimpleNamespace(optimizer=[RefOpt(), RefOpt(), RefOpt()], extractor=ref_ext,
                          refiner=SimpleNamespace(optimizer=None, feature_extractor=None, tracker=None))
    with pytest.raises(_lib.PtkError):
        install(loc)
    assert isinstance(loc.optimizer[0], RefOpt)          # nothing was half-swapped
    with pytest.raises(_lib.PtkError):
        NerfTestbed(np.zeros((8, 2), np.float16), [np.zeros((64, 32)), np.zeros((16, 64))],
                    [np.zeros((64, 32)), np.zeros((64, 64)), np.zeros((16, 64))], np.zeros(8 * 128 ** 3 // 8, np.uint8), 1, 'cpu')
    with pytest.raises(_lib.PtkError):
        query_mask(torch.zeros(4, 4, 3, dtype=torch.uint8))


def test_shared_structs_match_the_header_layout():
    from pixtrack_b200 import _lib
    assert ctypes.sizeof(_lib.NerfModelStruct) == 8 + 8 + 5 * 8 + 8 + 4 + 4
    assert ctypes.sizeof(_lib.NerfView) == (12 + 3 + 3 + 3 + 4 + 4) * 4
    assert ctypes.sizeof(_lib.RefLevel) == 4 * 8 + 2 * 8 + 4 * 4
    assert ctypes.sizeof(_lib.UnetWeights) == (20 + 20 + 3 + 3) * 8
