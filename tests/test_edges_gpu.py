"""GPU: edge cases across the hot path -- empty / maximum / ragged inputs."""
import numpy as np
import pytest
import torch

from oracle import lm as olm, nerf as onerf, unet as ounet
import synthetic as syn
import cases

pytestmark = pytest.mark.gpu
D = 'cuda:0'


def _args(p, B):
    from pixtrack_b200.optimizer import query_map_to_hwc
    T0 = torch.cat([p['R0'].reshape(B, 9), p['t0']], 1).to(D)
    return (p['p3d'].to(D), p['F_ref'].to(D), query_map_to_hwc(p['F_q'].to(D)), T0, p['cam'].to(D),
            cases.damping(torch.zeros(6)).to(D), p['W_ref'].reshape(B, -1).to(D), p['W_q'].to(D)), T0


def test_lm_no_points_and_zero_iterations():
    from pixtrack_b200.optimizer import lm_run_batched, query_map_to_hwc
    p = syn.level_problem(seed=3, N=64, C=32, H=40, W=60, level_scale=60 / 1920)
    args, T0 = _args(p, 1)
    # num_iters = 0: nothing runs, pose passes through, not failed
    T, failed, n, _ = lm_run_batched(*args, num_iters=0)
    torch.cuda.synchronize()
    assert torch.equal(T, T0) and int(n[0]) == 0 and not bool(failed[0])
    # N = 0 (every point dropped by the reference sampler): fails at the first iteration and keeps the pose,
    # like learned_optimizer.py:65 with an empty p3D
    empty = (args[0][:0], args[1][:, :0], args[2], args[3], args[4], args[5], args[6][:, :0], args[7])
    T, failed, n, log = lm_run_batched(*empty, num_iters=5)
    torch.cuda.synchronize()
    assert torch.equal(T, T0) and bool(failed[0]) and float(log[0, 0, 1]) == 0.0
    # all points masked out behaves the same
    T, failed, n, _ = lm_run_batched(*args, torch.zeros((1, 64), dtype=torch.uint8, device=D), num_iters=5)
    torch.cuda.synchronize()
    assert torch.equal(T, T0) and bool(failed[0])


def test_lm_more_problems_than_sms_and_widest_descriptor():
    """B = 300 independent problems (> 148 SMs: the problem groups loop) and C = 512 (the ABI maximum)."""
    from pixtrack_b200.optimizer import lm_run_batched
    B = 300
    p = syn.level_problem(seed=5, N=200, C=16, H=40, W=60, level_scale=60 / 1920, B=B)
    args, _ = _args(p, B)
    T, failed, n, _ = lm_run_batched(*args, num_iters=4, grad_stop=0, dt_stop=0, dR_stop=0)
    torch.cuda.synchronize()
    assert not bool(failed.any()) and int(n.min()) == 4
    for v in (0, 147, 148, 299):
        ref = olm.lm_run(p['p3d'], p['F_ref'][v], p['F_q'], p['R0'][v], p['t0'][v], p['cam'], p['W_ref'][v], p['W_q'],
                         num_iters=4, **cases.NO_STOP)
        np.testing.assert_allclose(T[v].cpu().numpy(), torch.cat([ref['R'].reshape(-1), ref['t']]).numpy(), atol=3e-5)
    q = syn.level_problem(seed=6, N=120, C=512, H=24, W=32, level_scale=32 / 1920)
    args, _ = _args(q, 1)
    T, failed, n, _ = lm_run_batched(*args, num_iters=3, grad_stop=0, dt_stop=0, dR_stop=0)
    ref = olm.lm_run(q['p3d'], q['F_ref'][0], q['F_q'], q['R0'][0], q['t0'][0], q['cam'], q['W_ref'][0], q['W_q'],
                     num_iters=3, **cases.NO_STOP)
    np.testing.assert_allclose(T[0].cpu().numpy(), torch.cat([ref['R'].reshape(-1), ref['t']]).numpy(), atol=3e-5)
    from pixtrack_b200 import _lib
    _lib.device_status(0)


@pytest.mark.parametrize('hw', [(16, 16), (37, 53), (271, 481)])
def test_extractor_ragged_sizes_against_oracle(hw):
    """Sizes that are not multiples of 16: floor pooling, cropped skips (unet.py:39-43), partial tiles in every kernel."""
    from pixtrack_b200.extractor import B200FeatureExtractor
    sd = syn.unet_weights(0)
    ext = B200FeatureExtractor(sd, D)
    img = syn.textured_image(hw[0], hw[1], seed=hw[0]).numpy().astype(np.float32)
    feats, scales, confs = ext(img)
    rf, rs, rc = ounet.extract(sd, img)
    assert [tuple(f.shape) for f in feats] == [tuple(f.shape) for f in rf] and scales == rs
    for f, c, of, oc in zip(feats, confs, rf, rc):
        rel = float((f.cpu() - of).norm() / of.norm().clamp_min(1e-6))
        assert rel < 2e-2, rel                           # fp16 operands vs the fp32 oracle, whole network
        assert float((c.cpu() - oc).abs().max()) < 3e-2


def test_nerf_camera_inside_box_cropped_render_box_and_single_sample():
    from pixtrack_b200.nerf import NerfTestbed, occupancy_bitfield
    sc = syn.nerf_scene(9, 2)
    bits = occupancy_bitfield(sc['density_grid'], sc['max_cascade'])
    box = np.array([[0.3, 0.3, 0.3], [0.9, 0.7, 0.8]], np.float32)     # cuts through the ball, not centred
    tb = NerfTestbed(sc['grid'], sc['w_density'], sc['w_rgb'], bits, 2, D, render_aabb=box)
    tb.nerf.rendering_min_transmittance = 1e-7
    m = onerf.NerfModel(2, sc['grid'], sc['w_density'], sc['w_rgb'], bits, render_aabb=box)
    for eye, fov, (W, H), spp in [((0.55, 0.45, 0.6), 70.0, (33, 21), 1),      # camera INSIDE the render box
                                  ((0.5, -1.2, 0.5), 30.0, (18, 40), 3)]:       # portrait, odd spp
        cam = syn.nerf_look_at(eye, target=(0.6, 0.9, 0.4) if eye[1] > 0 else (0.5, 0.5, 0.5))
        tb.fov = fov
        tb.set_ngp_camera_matrix(cam)
        got = tb.render(W, H, spp)
        ref = onerf.render(m, cam, W, H, fov, spp=spp)['rgba']
        assert got.shape == (H, W, 4) and ref[..., 3].max() > 0.5
        assert (np.abs(got - ref).max(-1) > 4e-3).mean() < 0.03
