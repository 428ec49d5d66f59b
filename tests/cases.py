"""Seeded inputs shared by the oracle tests and the GPU parity tests.  They
rebuild exactly what tests/golden/gen/make_goldens.py fed to the reference."""
import os

import numpy as np
import torch

import synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

LEVEL_CASES = dict(
    l1_s0=dict(seed=0, N=500, C=128, H=144, W=256, level_scale=(1024 / 1920) / 4),
    l1_s1=dict(seed=1, N=500, C=128, H=144, W=256, level_scale=(1024 / 1920) / 4, fixed_iters=15),
    l2_s2=dict(seed=2, N=300, C=128, H=36, W=64, level_scale=(1024 / 1920) / 16, sigma=1.0),
    l0_s3=dict(seed=3, N=400, C=32, H=576, W=1024, level_scale=(1024 / 1920), sigma=4.0, rot_deg=0.3, trans=0.004),
    l1_k1=dict(seed=4, N=500, C=64, H=144, W=256, level_scale=(1024 / 1920) / 4, k1=-0.12, fixed_iters=12),
)
NO_STOP = dict(grad_stop=0.0, dt_stop=0.0, dR_stop=0.0)


def gold(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def checksum(*tensors):
    return np.array([float(t.double().abs().sum()) for t in tensors])


def damping(const):
    const = torch.as_tensor(const, dtype=torch.float32)
    return 10.0 ** (-6.0 + torch.sigmoid(const) * 11.0)


def level_case(name):
    """-> (problem dict, lm kwargs, golden npz)"""
    kw = dict(LEVEL_CASES[name])
    fixed = kw.pop('fixed_iters', 0)
    p = syn.level_problem(**kw)
    g = gold('lm_' + name)
    np.testing.assert_allclose(checksum(p['p3d'], p['F_ref'], p['F_q'], p['W_q']), g['chk'], rtol=1e-9)
    lm_kw = dict(lam=damping(g['const']), num_iters=fixed or 150, **(NO_STOP if fixed else {}))
    return p, lm_kw, g


def edge_few():
    p = syn.level_problem(seed=5, N=40, C=16, H=48, W=64, level_scale=0.05)
    p['p3d'] = p['p3d'].clone()
    p['p3d'][8:, 2] = -1.0
    return p


def edge_tangential():
    q = syn.level_problem(seed=6, N=600, C=16, H=120, W=160, level_scale=160 / 1920)
    cam10 = torch.cat([q['cam'][:6], torch.tensor([0.15, -0.3, 0.002, -0.001])])
    cam10[2:4] = cam10[2:4] * 0.35
    q['cam'] = cam10
    return q


def pyramid_scene(seed, N=600):
    dims = ((32, 144, 256), (128, 36, 64), (128, 9, 16))
    cam_q = syn.pixtrack_camera(1920, 1080)
    sr = 256 / 1920
    scales = [(sr / s, sr / s) for s in (1, 4, 16)]
    maps = []
    for lv, (C, H, W) in enumerate(dims):
        f = syn.smooth_feature_map(C, H, W, seed * 31 + lv, sigma=(6.0, 1.5, 0.6)[lv], normalize=False)
        c = syn.smooth_confidence(H, W, seed * 31 + 10 + lv)
        maps.append(torch.cat([f, c], 0))
    p3d = syn.object_points(N, seed * 31 + 20, depth=1.2, spread=0.25)
    R_gt = syn.axis_angle_to_R(torch.randn(3, generator=torch.Generator().manual_seed(seed * 31 + 21)) * 0.05)
    t_gt = torch.tensor([0.01, -0.02, 0.03])
    return cam_q, scales, maps, p3d, R_gt, t_gt


def interp_fixture():
    torch.random.manual_seed(0)
    w, h = 480, 240
    torch.rand(1000, 2)
    return torch.rand(16, h, w) * 100
