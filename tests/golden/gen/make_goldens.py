"""Generate tests/golden/*.npz by executing the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/gen/make_goldens.py

The reference has no golden vectors for this path, so
these fixtures pin the oracle (and through it the CUDA path) to what the
reference's own classes compute: `PixTrackOptimizer.run`, `Camera.world2image`
/ `J_world2image`, `interpolate_tensor`, `PoseTrackerRefiner.
interp_sparse_observations`, `BaseRefiner.refine_pose_using_features`,
`UNet._forward`, `PixTrackFeatureExtractor.__call__`.

Inputs are NOT stored (a 640x480x16 map is 20 MB): tests regenerate them from
`tests/synthetic.py` with the seeds recorded here; each fixture carries
an input checksum so RNG drift is detected instead of mis-reported as a
parity failure.  Stand-ins used to import the reference: `omegaconf` (this
directory), empty `h5py`, `torch._six.string_classes`; torchvision's vgg19 is
wrapped to skip the ImageNet download (weights are then overwritten).
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..', '..'))
OUT = os.path.abspath(os.path.join(HERE, '..'))
sys.path[:0] = [HERE, '/root/reference/pixloc', '/root/reference', ROOT]
sys.modules['h5py'] = types.ModuleType('h5py')

import torch  # noqa: E402

_six = types.ModuleType('torch._six')
_six.string_classes = (str, bytes)
sys.modules['torch._six'] = _six
import torchvision  # noqa: E402

_vgg19 = torchvision.models.vgg19
torchvision.models.vgg19 = lambda pretrained=False, **kw: _vgg19(weights=None)

torch.set_grad_enabled(False)      # pixloc/pixloc/localization/localizer.py:21
torch.set_num_threads(1)            # fixed summation order in the fixtures

from pixloc.pixlib.geometry import Camera, Pose  # noqa: E402
from pixloc.pixlib.geometry.interpolation import interpolate_tensor  # noqa: E402
from pixloc.pixlib.models.learned_optimizer import LearnedOptimizer  # noqa: E402
from pixloc.pixlib.models.unet import UNet  # noqa: E402
from pixtrack.optimizers.pixtrack_optimizer import PixTrackOptimizer  # noqa: E402
from pixtrack.localization.pixloc_pose_refiners import PoseTrackerRefiner  # noqa: E402
from pixtrack.localization.feature_extractor import PixTrackFeatureExtractor  # noqa: E402
from pixtrack.localization.tracker import DebugTracker  # noqa: E402

import os as _os, sys as _sys  # noqa: E401,E402
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))), 'tests'))  # scene generators live with the tests
import synthetic as syn  # noqa: E402


def checksum(*tensors):
    return np.array([float(t.double().abs().sum()) for t in tensors])


def make_optimizer(num_iters=150, const=None, **over):
    conf = dict(num_iters=num_iters, pad=1, loss_fn='scaled_barron(0, 0.1)')   # r9.py:46-49 + checkpoint conf
    conf.update(over)
    opt = LearnedOptimizer(conf)
    opt.__class__ = PixTrackOptimizer                  # pixloc_pose_refiners.py:71-72
    opt.eval()
    if const is not None:
        opt.dampingnet.const.data.copy_(torch.as_tensor(const, dtype=torch.float32))
    return opt


def run_logged(opt, p3d, F_ref, F_q, R0, t0, cam, W_ref, W_q):
    rec = dict(g=[], H=[], T=[], n_valid=[], cost_sum=[], dt=[])

    def fn(**kw):
        v = kw['valid'].float()
        rec['H'].append(kw['H'].numpy().copy())
        rec['T'].append(kw['T']._data.numpy().copy())
        rec['n_valid'].append(float(v.sum()))
        rec['cost_sum'].append(float((v * kw['cost']).sum()))
        rec['dt'].append(float(kw['T_delta'].magnitude()[1]))
    opt.logging_fn = fn
    W = None if W_ref is None else (W_ref, W_q)
    T, failed = opt.run(p3d, F_ref, F_q, Pose.from_Rt(R0, t0), Camera(cam), W_ref_query=W)
    opt.logging_fn = None
    out = {k: np.array(v) for k, v in rec.items() if len(v)}
    out['T_final'] = T._data.numpy()
    out['failed'] = np.array(bool(failed))
    out['n_iters'] = np.array(len(rec['T']))
    return out


def save(name, **arrs):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrs)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


# --- A. the reference's own toy fixture (config C1a) -----------------------
def gold_lm_toy():
    p = syn.toy_problem(0, 500)
    fixed = make_optimizer(5, grad_stop_criteria=0, dt_stop_criteria=0, dR_stop_criteria=0)
    a = run_logged(fixed, p['p3d'], p['F_ref'], p['F_q'], p['R0'], p['t0'], p['cam'], p['W_ref'], p['W_q'])
    b = run_logged(make_optimizer(150), p['p3d'], p['F_ref'], p['F_q'], p['R0'], p['t0'], p['cam'], None, None)
    save('lm_toy', chk=checksum(p['p3d'], p['F_ref'], p['F_q'], p['W_q']),
         **{'fixed_' + k: v for k, v in a.items()}, **{'nowt_' + k: v for k, v in b.items()})


# --- B/C. PixTrack-shaped single levels (config C1b) -----------------------
def gold_lm_levels():
    cases = dict(
        l1_s0=dict(seed=0, N=500, C=128, H=144, W=256, level_scale=(1024 / 1920) / 4),
        l1_s1=dict(seed=1, N=500, C=128, H=144, W=256, level_scale=(1024 / 1920) / 4, fixed_iters=15),
        l2_s2=dict(seed=2, N=300, C=128, H=36, W=64, level_scale=(1024 / 1920) / 16, sigma=1.0),
        l0_s3=dict(seed=3, N=400, C=32, H=576, W=1024, level_scale=(1024 / 1920), sigma=4.0, rot_deg=0.3, trans=0.004),
        l1_k1=dict(seed=4, N=500, C=64, H=144, W=256, level_scale=(1024 / 1920) / 4, k1=-0.12, fixed_iters=12),
    )
    for name, kw in cases.items():
        fixed = kw.pop('fixed_iters', 0)
        p = syn.level_problem(**kw)
        const = torch.linspace(-1.0, 1.0, 6) * (kw['seed'] % 3)
        opt = make_optimizer(150, const) if not fixed else make_optimizer(
            fixed, const, grad_stop_criteria=0, dt_stop_criteria=0, dR_stop_criteria=0)
        out = run_logged(opt, p['p3d'], p['F_ref'][0], p['F_q'], p['R0'][0], p['t0'][0],
                         p['cam'], p['W_ref'][0], p['W_q'])
        save('lm_' + name, chk=checksum(p['p3d'], p['F_ref'], p['F_q'], p['W_q']), const=const.numpy(), **out)


# --- D. edge cases ----------------------------------------------------------
def gold_lm_edge():
    # (i) fewer than 10 valid points -> failed, pose untouched, loop stops at once
    p = syn.level_problem(seed=5, N=40, C=16, H=48, W=64, level_scale=0.05)
    p3d = p['p3d'].clone()
    p3d[8:, 2] = -1.0                      # behind the camera
    few = run_logged(make_optimizer(20), p3d, p['F_ref'][0], p['F_q'], p['R0'][0], p['t0'][0], p['cam'],
                     p['W_ref'][0], p['W_q'])
    # (ii) OPENCV-style camera with tangential terms + a radial limit that bites
    q = syn.level_problem(seed=6, N=600, C=16, H=120, W=160, level_scale=160 / 1920)
    cam10 = torch.cat([q['cam'][:6], torch.tensor([0.15, -0.3, 0.002, -0.001])])
    cam10[2:4] = cam10[2:4] * 0.35        # wide field of view so the limit is reached
    tang = run_logged(make_optimizer(8, grad_stop_criteria=0, dt_stop_criteria=0, dR_stop_criteria=0),
                      q['p3d'], q['F_ref'][0], q['F_q'], q['R0'][0], q['t0'][0], cam10, q['W_ref'][0], q['W_q'])
    save('lm_edge', chk=checksum(p['F_q'], q['F_q']), cam10=cam10.numpy(),
         **{'few_' + k: v for k, v in few.items()}, **{'tang_' + k: v for k, v in tang.items()})


# --- E. geometry primitives -------------------------------------------------
def gold_geometry():
    g = torch.Generator().manual_seed(11)
    pc = torch.randn(400, 3, generator=g) * torch.tensor([1.5, 1.0, 1.5]) + torch.tensor([0, 0, 1.5])
    out = {}
    for tag, dist in (('d0', []), ('d2', [0.1, 0.01]), ('d2n', [-0.2, 0.05]), ('d4', [0.15, -0.3, 0.002, -0.001])):
        cam = Camera(torch.tensor([640., 480., 300., 350., 320., 240.] + dist))
        uv, valid = cam.world2image(pc)
        J, _ = cam.J_world2image(pc)
        out.update({f'{tag}_uv': uv.numpy(), f'{tag}_valid': valid.numpy(), f'{tag}_J': J.numpy()})
    # interpolation fixture of the reference's test_run_all (interpolation.py:195-203), first 200 points
    torch.random.manual_seed(0)
    w, h = 480, 240
    pts = torch.rand(1000, 2) * torch.tensor([w - 1, h - 1])
    tensor = torch.rand(16, h, w) * 100
    pts = torch.cat([pts[:200], torch.tensor([[0.5, 10.0], [1.0, 1.0], [w - 2.0, h - 2.0], [w - 1.5, 5.0], [-3.0, 4.0]])])
    val, mask, grad = interpolate_tensor(tensor, pts, 'linear', 1, True)
    val0, mask0, _ = interpolate_tensor(tensor, pts, 'linear', 0, False)
    save('geometry', pc=pc.numpy(), pts=pts.numpy(), interp_val=val.numpy(), interp_mask=mask.numpy(),
         interp_grad=grad.numpy(), interp_mask_pad0=mask0.numpy(), chk=checksum(tensor), **out)


# --- F/G. reference sparse sampling + coarse-to-fine refine -----------------
def pyramid_scene(seed, N=600):
    """3-level (C+1)-channel query/reference pyramids sharing one scene."""
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


def fake_refiner(optimizers, p3d, cam_q, R, t):
    colcam = SimpleNamespace(model='SIMPLE_RADIAL', width=int(cam_q[0]), height=int(cam_q[1]),
                             params=np.array([float(cam_q[2]), float(cam_q[4]) + 0.5, float(cam_q[5]) + 0.5, float(cam_q[6])]))
    colcam._asdict = lambda: dict(model=colcam.model, width=colcam.width, height=colcam.height, params=colcam.params)
    img = SimpleNamespace(camera_id=1, qvec2rotmat=lambda: R.double().numpy(), tvec=t.double().numpy(), name='ref')
    pts = {i: SimpleNamespace(xyz=p3d[i].double().numpy()) for i in range(p3d.shape[0])}
    model3d = SimpleNamespace(dbs={7: img}, cameras={1: colcam._asdict()}, points3D=pts)
    conf = dict(num_dbs=1, multiscale=[1], point_selection='all', normalize_descriptors=True,
                average_observations=False, do_pose_approximation=False)                # r9.py:50-57
    return PoseTrackerRefiner(torch.device('cpu'), optimizers, model3d, None, None, conf)


def gold_refine():
    cam_q, scales, maps, p3d, R_gt, t_gt = pyramid_scene(1)
    consts = [torch.full((6,), v) for v in (0.3, -0.2, 0.1)]
    opts = [make_optimizer(150, c) for c in consts]
    refiner = fake_refiner(opts, p3d, cam_q, R_gt, t_gt)
    refiner.reference_scale = 1.0
    ids = list(range(p3d.shape[0]))
    # F: interp_sparse_observations at the render pose (pixloc_pose_refiners.py:327-368)
    fd = refiner.interp_sparse_observations(maps, scales, 7, ids, Pose.from_Rt(R_gt, t_gt).double())  # r9 poses are cpu float64 (base_refiner.py:128)
    kept = np.array(sorted(fd.keys()))
    obs = [torch.stack([fd[i][lv] for i in kept]).numpy() for lv in range(3)]
    # G: refine_pose_using_features from a perturbed pose, DebugTracker attached (r9.py:239,251)
    tracker = DebugTracker(refiner, debug=1)
    R0, t0 = syn.perturb_pose(R_gt, t_gt, 99, 1.5, 0.015)
    feats = [tuple(fd[i]) for i in kept]
    ret = refiner.refine_pose_using_features(maps, scales, Camera(cam_q), Pose.from_Rt(R0, t0).double(), feats, list(kept))
    save('refine', kept=kept, obs0=obs[0], obs1=obs[1], obs2=obs[2], consts=torch.stack(consts).numpy(),
         chk=checksum(*maps, p3d), success=np.array(ret['success']),
         T_refined=ret['T_refined']._data.numpy(), diff_R=np.array(ret['diff_R']), diff_t=np.array(ret['diff_t']),
         last_costs=np.array([float(c[-1]) for c in tracker.costs]), num_iters=np.array(tracker.num_iters))


# --- H/I. extractor ---------------------------------------------------------
def gold_unet():
    sd = syn.unet_weights(0)
    net = UNet(dict(encoder='vgg19', decoder=[64, 64, 64, 32], output_scales=[0, 2, 4],
                    output_dim=[32, 128, 128], compute_uncertainty=True)).eval()
    net.load_state_dict(sd)
    out = {}
    for tag, (h, w) in (('a', (64, 96)), ('b', (80, 112))):
        img = syn.textured_image(h, w, seed=3)
        x = (img.permute(2, 0, 1) / 255.)[None]
        pred = net({'image': x})
        for lv in range(3):
            out[f'{tag}_f{lv}'] = pred['feature_maps'][lv][0].numpy()
            out[f'{tag}_c{lv}'] = pred['confidences'][lv][0].numpy()
    # the public call, with the resize branch (feature_extractor.py:40-44): 150x200 -> max edge 128
    ext = PixTrackFeatureExtractor(net, torch.device('cpu'), dict(resize=128))
    img = syn.textured_image(150, 200, seed=4).numpy()
    feats, scales, confs = ext(img, 1)
    for lv in range(3):
        out[f'x_f{lv}'] = feats[lv].numpy()
        out[f'x_c{lv}'] = confs[lv].numpy()
    out['x_scales'] = np.array(scales)
    save('unet', chk=checksum(*[sd[k].float() for k in sorted(sd)]), **out)


if __name__ == '__main__':
    gold_lm_toy()
    gold_lm_levels()
    gold_lm_edge()
    gold_geometry()
    gold_refine()
    gold_unet()
