"""GPU: reference-view sparse sampling (ptk_sample_reference) and the chained coarse-to-fine
refinement against the golden produced by the reference's
`PoseTrackerRefiner.interp_sparse_observations` + `BaseRefiner.refine_pose_using_features`
(tests/golden/refine.npz), and against the oracle on other seeds."""
import numpy as np
import pytest
import torch
import torch.nn.functional as tF

import cases
import synthetic as syn

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
D = 'cuda:0'


def _scene_on_device(seed, N=600):
    cam_q, scales, maps, p3d, R_gt, t_gt = cases.pyramid_scene(seed, N)
    feats = [m[:-1].permute(1, 2, 0).contiguous().to(D) for m in maps]
    confs = [m[-1].contiguous().to(D) for m in maps]
    return cam_q, scales, maps, p3d, R_gt, t_gt, feats, confs


def test_reference_sampling_matches_reference_golden():
    from pixtrack_b200.sampling import sample_reference
    g = cases.gold('refine')
    cam_q, scales, maps, p3d, R_gt, t_gt, feats, confs = _scene_on_device(1)
    T = torch.cat([R_gt.double().reshape(-1), t_gt.double()])
    F, Wr, valid = sample_reference(feats, confs, scales, cam_q.double(), T, p3d.double().to(D), pad=1, normalize=False)
    torch.cuda.synchronize()
    kept = torch.nonzero(valid.cpu())[:, 0].numpy()
    assert np.array_equal(kept, g['kept'])
    for lv in range(3):
        obs = g[f'obs{lv}']
        np.testing.assert_allclose(F[lv].cpu().numpy()[kept], obs[:, :-1], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(Wr[lv].cpu().numpy()[kept], obs[:, -1], rtol=1e-5, atol=1e-6)
    Fn, _, _ = sample_reference(feats, confs, scales, cam_q.double(), T, p3d.double().to(D), pad=1, normalize=True)
    for lv in range(3):
        ref = tF.normalize(torch.from_numpy(g[f'obs{lv}'][:, :-1]), dim=1).numpy()
        np.testing.assert_allclose(Fn[lv].cpu().numpy()[kept], ref, rtol=1e-5, atol=1e-6)


def test_sampling_then_chained_levels_match_reference_golden():
    """interp_sparse_observations -> refine_pose_using_features of the reference == one
    ptk_sample_reference launch + three chained ptk_lm_run launches (invalid points masked
    instead of dropped)."""
    from pixtrack_b200.optimizer import query_map_to_hwc
    from pixtrack_b200.refiner import refine_levels_batched
    from pixtrack_b200.sampling import sample_reference
    g = cases.gold('refine')
    cam_q, scales, maps, p3d, R_gt, t_gt, feats, confs = _scene_on_device(1)
    T = torch.cat([R_gt.double().reshape(-1), t_gt.double()])
    F, Wr, valid = sample_reference(feats, confs, scales, cam_q.double(), T, p3d.double().to(D), pad=1)
    R0, t0 = syn.perturb_pose(R_gt, t_gt, 99, 1.5, 0.015)
    T0 = torch.cat([R0.reshape(-1), t0])[None].to(D)
    fq = [query_map_to_hwc(m[:-1].to(D), normalize=True)[None] for m in maps]       # base_refiner.py:92-94
    wq = [c[None] for c in confs]
    cams = [syn.scale_cam(cam_q, s).to(D)[None] for s in scales]
    lams = [cases.damping(c).to(D) for c in g['consts']]
    out = refine_levels_batched(fq, wq, cams, [f[None] for f in F], [w[None] for w in Wr], p3d.to(D)[None], T0, lams,
                                mask=valid[None])
    torch.cuda.synchronize()
    assert bool(out['failed'][0]) != bool(g['success'])
    assert [int(n[0]) for n in out['n_iters']] == list(g['num_iters'])
    last = []
    for n, lg in zip(out['n_iters'], out['logs']):
        r = lg[0, int(n[0]) - 1].cpu().numpy()
        last.append(r[0] / r[1])
    np.testing.assert_allclose(last, g['last_costs'], rtol=1e-4)
    np.testing.assert_allclose(out['T'][0].cpu().numpy(), g['T_refined'], atol=2e-5)


@pytest.mark.parametrize('seed,dist', [(2, ()), (3, (-0.1, 0.02)), (4, (0.05, -0.01, 0.001, -0.002))])
def test_reference_sampling_against_oracle(seed, dist):
    """other seeds, radial and tangential cameras, a pose that pushes part of the cloud out of view,
    non-square level scales."""
    from oracle import lm
    from pixtrack_b200.sampling import sample_reference
    cam_q, scales, maps, p3d, R_gt, t_gt, feats, confs = _scene_on_device(seed, N=777)
    cam = torch.cat([cam_q[:6], torch.tensor(dist)]).double() if dist else cam_q[:6].double()
    scales = [(s[0] * 1.01, s[1] * 0.99) for s in scales]
    t = t_gt.double() + torch.tensor([0.35, 0.0, 0.0], dtype=torch.float64)
    obs, keep = lm.sample_reference(maps, scales, cam, R_gt.double(), t, p3d.double())
    T = torch.cat([R_gt.double().reshape(-1), t])
    F, Wr, valid = sample_reference(feats, confs, scales, cam, T, p3d.double().to(D), pad=1, normalize=False)
    torch.cuda.synchronize()
    assert 0 < int(keep.sum()) < 777
    assert torch.equal(valid.cpu().bool(), keep)
    for lv in range(3):
        np.testing.assert_allclose(F[lv].cpu().numpy()[keep], obs[lv][keep][:, :-1].numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(Wr[lv].cpu().numpy()[keep], obs[lv][keep][:, -1].numpy(), rtol=1e-5, atol=1e-6)


def test_reference_sampling_empty_and_preallocated_outputs():
    from pixtrack_b200.sampling import sample_reference
    cam_q, scales, maps, p3d, R_gt, t_gt, feats, confs = _scene_on_device(5, N=64)
    T = torch.cat([R_gt.double().reshape(-1), t_gt.double()])
    F, Wr, valid = sample_reference(feats, confs, scales, cam_q.double(), T, p3d.double().to(D)[:0].contiguous())
    assert valid.numel() == 0 and F[0].shape == (0, 32)
    # caller-owned slices of a [B, N, C] cache
    cache = [torch.zeros((2, 64, f.shape[2]), device=D) for f in feats]
    wc = [torch.zeros((2, 64), device=D) for _ in feats]
    vc = torch.zeros((2, 64), dtype=torch.uint8, device=D)
    sample_reference(feats, confs, scales, cam_q.double(), T, p3d.double().to(D),
                     out=([c[1] for c in cache], [w[1] for w in wc], vc[1]))
    F, Wr, valid = sample_reference(feats, confs, scales, cam_q.double(), T, p3d.double().to(D))
    torch.cuda.synchronize()
    for lv in range(3):
        assert torch.equal(cache[lv][1], F[lv]) and torch.equal(wc[lv][1], Wr[lv]) and not bool(cache[lv][0].any())
    assert torch.equal(vc[1], valid)
