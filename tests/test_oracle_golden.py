"""CPU: the oracle (oracle/*.py) against fixtures produced by the unmodified
reference (tests/golden/gen/make_goldens.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

import cases
from oracle import lm, unet
import synthetic as syn

torch.set_grad_enabled(False)


def _check_run(out, g, prefix='', rtol=2e-4, atol_T=2e-5):
    n = int(g[prefix + 'n_iters'])
    assert out['n_iters'] == n
    assert bool(out['failed']) == bool(g[prefix + 'failed'])
    for i, e in enumerate(out['log']):
        H = g[prefix + 'H'][i]
        np.testing.assert_allclose(e['H'].numpy(), H, rtol=0, atol=rtol * np.abs(H).max())
        assert e['n_valid'] == g[prefix + 'n_valid'][i]
        np.testing.assert_allclose(e['cost_sum'], g[prefix + 'cost_sum'][i], rtol=1e-4)
        T = torch.cat([e['R'].reshape(-1), e['t']]).numpy()
        np.testing.assert_allclose(T, g[prefix + 'T'][i], atol=atol_T)
    T = torch.cat([out['R'].reshape(-1), out['t']]).numpy()
    np.testing.assert_allclose(T, g[prefix + 'T_final'], atol=atol_T)


def test_toy_fixture_fixed_iters():
    p, g = syn.toy_problem(0, 500), cases.gold('lm_toy')
    np.testing.assert_allclose(cases.checksum(p['p3d'], p['F_ref'], p['F_q'], p['W_q']), g['chk'], rtol=1e-9)
    out = lm.lm_run(p['p3d'], p['F_ref'], p['F_q'], p['R0'], p['t0'], p['cam'], p['W_ref'], p['W_q'],
                    num_iters=5, **cases.NO_STOP)
    _check_run(out, g, 'fixed_')


def test_toy_fixture_no_confidence_early_stop():
    p, g = syn.toy_problem(0, 500), cases.gold('lm_toy')
    out = lm.lm_run(p['p3d'], p['F_ref'], p['F_q'], p['R0'], p['t0'], p['cam'])
    _check_run(out, g, 'nowt_')


@pytest.mark.parametrize('name', list(cases.LEVEL_CASES))
def test_pixtrack_levels(name):
    p, kw, g = cases.level_case(name)
    out = lm.lm_run(p['p3d'], p['F_ref'][0], p['F_q'], p['R0'][0], p['t0'][0], p['cam'],
                    p['W_ref'][0], p['W_q'], **kw)
    _check_run(out, g)


def test_too_few_points_fails_and_keeps_pose():
    p, g = cases.edge_few(), cases.gold('lm_edge')
    out = lm.lm_run(p['p3d'], p['F_ref'][0], p['F_q'], p['R0'][0], p['t0'][0], p['cam'], p['W_ref'][0], p['W_q'],
                    num_iters=20)
    _check_run(out, g, 'few_')
    assert out['failed'] and out['n_iters'] == 1
    assert torch.equal(out['R'], p['R0'][0]) and torch.equal(out['t'], p['t0'][0])


def test_tangential_camera_with_radial_limit():
    q, g = cases.edge_tangential(), cases.gold('lm_edge')
    np.testing.assert_allclose(q['cam'].numpy(), g['cam10'])
    out = lm.lm_run(q['p3d'], q['F_ref'][0], q['F_q'], q['R0'][0], q['t0'][0], q['cam'], q['W_ref'][0], q['W_q'],
                    num_iters=8, **cases.NO_STOP)
    _check_run(out, g, 'tang_')


def test_projection_and_jacobian():
    g = cases.gold('geometry')
    pc = torch.from_numpy(g['pc'])
    for tag, dist in (('d0', []), ('d2', [0.1, 0.01]), ('d2n', [-0.2, 0.05]), ('d4', [0.15, -0.3, 0.002, -0.001])):
        cam = torch.tensor([640., 480., 300., 350., 320., 240.] + dist)
        uv, valid = lm.world_to_image(cam, pc)
        J = lm.world_to_image_jacobian(cam, pc)
        assert np.array_equal(valid.numpy(), g[f'{tag}_valid'])
        np.testing.assert_allclose(uv.numpy(), g[f'{tag}_uv'], rtol=1e-6, atol=1e-4)
        np.testing.assert_allclose(J.numpy(), g[f'{tag}_J'], rtol=1e-5, atol=1e-3)
    # the distortion limit really bites: some points land inside the image yet are rejected
    cam = torch.tensor([640., 480., 300., 350., 320., 240., 0.15, -0.3, 0.002, -0.001])
    uv, valid = lm.world_to_image(cam, pc)
    inside = ((uv >= 0) & (uv <= cam[:2] - 1)).all(-1) & (pc[:, 2] > 1e-3)
    assert (inside & ~valid).sum() > 0


def test_interpolation_fixture():
    g = cases.gold('geometry')
    tensor = cases.interp_fixture()
    np.testing.assert_allclose(cases.checksum(tensor), g['chk'], rtol=1e-9)
    pts = torch.from_numpy(g['pts'])
    val, mask, grad = lm.sample_map(tensor, pts, pad=1, grads=True)
    np.testing.assert_allclose(val.numpy(), g['interp_val'], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(grad.numpy(), g['interp_grad'], rtol=1e-6, atol=1e-5)
    assert np.array_equal(mask.numpy(), g['interp_mask'])
    assert np.array_equal(lm.sample_map(tensor, pts, pad=0)[1].numpy(), g['interp_mask_pad0'])


def test_reference_sampling_and_refine_levels():
    g = cases.gold('refine')
    cam_q, scales, maps, p3d, R_gt, t_gt = cases.pyramid_scene(1)
    np.testing.assert_allclose(cases.checksum(*maps, p3d), g['chk'], rtol=1e-9)
    obs, keep = lm.sample_reference(maps, scales, cam_q.double(), R_gt.double(), t_gt.double(), p3d.double())
    kept = torch.nonzero(keep)[:, 0]
    assert np.array_equal(kept.numpy(), g['kept'])
    for lv in range(3):
        np.testing.assert_allclose(obs[lv][kept].numpy(), g[f'obs{lv}'], rtol=1e-5, atol=1e-6)
    R0, t0 = syn.perturb_pose(R_gt, t_gt, 99, 1.5, 0.015)
    lams = [cases.damping(c) for c in g['consts']]
    out = lm.refine_levels(maps, scales, cam_q, R0, t0, [o[kept] for o in obs], p3d[kept], lams)
    assert out['success'] == bool(g['success'])
    # tracker.costs / num_iters are appended coarse -> fine (base_refiner.py:101)
    assert [r['n_iters'] for r in out['runs']] == list(g['num_iters'])
    last = [r['log'][-1]['cost_sum'] / r['log'][-1]['n_valid'] for r in out['runs']]
    np.testing.assert_allclose(last, g['last_costs'], rtol=1e-4)
    T = torch.cat([out['R'].reshape(-1), out['t']]).numpy()
    np.testing.assert_allclose(T, g['T_refined'], atol=2e-5)


def test_unet_forward_and_public_call():
    g = cases.gold('unet')
    sd = syn.unet_weights(0)
    np.testing.assert_allclose(cases.checksum(*[sd[k].float() for k in sorted(sd)]), g['chk'], rtol=1e-9)
    for tag, (h, w) in (('a', (64, 96)), ('b', (80, 112))):
        img = syn.textured_image(h, w, seed=3)
        feats, confs = unet.unet_forward(sd, (img.permute(2, 0, 1) / 255.)[None])
        for lv in range(3):
            np.testing.assert_allclose(feats[lv][0].numpy(), g[f'{tag}_f{lv}'], rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(confs[lv][0].numpy(), g[f'{tag}_c{lv}'], rtol=1e-4, atol=1e-5)
    feats, scales, confs = unet.extract(sd, syn.textured_image(150, 200, seed=4).numpy(), 1, resize=128)
    np.testing.assert_allclose(np.array(scales), g['x_scales'])
    for lv in range(3):
        np.testing.assert_allclose(feats[lv].numpy(), g[f'x_f{lv}'], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(confs[lv].numpy(), g[f'x_c{lv}'], rtol=1e-4, atol=1e-5)
