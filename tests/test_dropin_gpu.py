"""GPU: the drop-in proof.  A reference-shaped `PoseTrackerLocalizer` (same attributes the reference builds in
pixtrack/localization/pixloc_pose_refiners.py:29-93) gets `pixtrack_b200.install()`, then the REFERENCE'S OWN CONTROL
FLOW -- restated here line by line, because /root/reference does not exist on the GPU box -- is driven through the
swapped objects:

  * `PoseTrackerRefiner.interp_sparse_observations`   (pixloc_pose_refiners.py:327-368): per level
    `camera.scale(sc).world2image(p3d_cam)`, `opt.interpolator(feats, p2d)`, AND of masks, per-point lists of tensors;
  * `BaseRefiner.refine_pose_using_features`          (pixloc/pixloc/localization/base_refiner.py:64-137):
    `torch.stack` of the per-point tuples, `[:, :-1]` / `[:, -1:]`, `F.normalize(dim=1)`, query `[:-1]` / `[-1:]`,
    `F.normalize(dim=0)`, coarse-to-fine `opt.run(p3d, F_ref, F_q, T_i.to(F_q), qcamera_feat.to(F_q), W_ref_query=...)`;
  * `DebugTracker.log_optim_iter` as `logging_fn`     (pixtrack/localization/tracker.py:32-46, attached by
    `BaseTracker.__init__`, pixloc/pixloc/localization/tracker.py:5-13);
  * `BaseRefiner.dense_feature_extraction`            (base_refiner.py:189-207): `extractor(image, scale)` then
    `torch.cat([f, w], 0)` per level.

It has to land on what the unmodified reference produced on the same inputs: tests/golden/refine.npz (`T_refined`,
`num_iters`, `last_costs`, kept point ids, per-point observations) and tests/golden/unet.npz (the public extractor call).
"""
import copy
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import cases
import synthetic as syn

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
D = torch.device('cuda:0')


class AttrDict(dict):
    """Stands in for the OmegaConf DictConfig the reference modules carry (attribute and item access)."""
    __getattr__ = dict.__getitem__


def reference_shaped_optimizer(const):
    """What `PoseTrackerLocalizer.__init__` holds per level: an nn.Module with `.conf`, `.dampingnet.const`, `.logging_fn`
    (learned_optimizer.py:30-46, base_optimizer.py:23-60; conf values of the PixTrack experiment)."""
    m = torch.nn.Module()
    m.conf = AttrDict(num_iters=150, loss_fn='scaled_barron(0, 0.1)', jacobi_scaling=False, normalize_features=False,
                      lambda_=0.0, interpolation=AttrDict(mode='linear', pad=1), grad_stop_criteria=1e-4,
                      dt_stop_criteria=5e-3, dR_stop_criteria=5e-2, damping=AttrDict(type='constant', log_range=[-6, 5]),
                      learned_damping=True)
    m.dampingnet = torch.nn.Module()
    m.dampingnet.const = torch.nn.Parameter(torch.as_tensor(const, dtype=torch.float32).clone())
    m.logging_fn = None
    return m


class ReferenceShapedExtractor:
    """`PixTrackFeatureExtractor` as far as install() reads it (feature_extractor.py:15-32)."""

    def __init__(self, sd, device, resize):
        self.conf = AttrDict(resize=resize, resize_by='max')
        self.device = device
        self.model = SimpleNamespace(state_dict=lambda: sd, scales=[1, 4, 16])


class DebugTrackerRestated:
    """pixtrack/localization/tracker.py:5-46 with debug=1 + BaseTracker.__init__."""

    def __init__(self, refiner):
        refiner.tracker = self
        opts = refiner.optimizer
        opts = opts if isinstance(opts, (tuple, list)) else [opts]
        for opt in opts:
            opt.logging_fn = self.log_optim_iter
        self.costs, self.T, self.dt, self.num_iters, self.done = [], [], [], [], []

    def log_optim_done(self, **args):
        self.done.append(args['level'])

    def log_optim_iter(self, **args):
        if args['i'] == 0:
            self.costs.append([])
            self.T.append(args['T_init'].cpu())
            self.num_iters.append(None)
        valid = args['valid'].float()
        cost = (valid * args['cost']).sum(-1) / valid.sum(-1)
        self.costs[-1].append(cost.cpu().numpy())
        self.dt.append(args['T_delta'].magnitude()[1].cpu().numpy())
        self.num_iters[-1] = args['i'] + 1
        self.T.append(args['T'].cpu())


def build_localizer(consts, resize=128):
    opts = [reference_shaped_optimizer(c) for c in consts]
    ext = ReferenceShapedExtractor(syn.unet_weights(0), D, resize)
    refiner = SimpleNamespace(optimizer=opts, feature_extractor=ext, device=D,
                              conf=AttrDict(compute_uncertainty=True, normalize_descriptors=True, layer_indices=None))
    loc = SimpleNamespace(optimizer=opts, extractor=ext, refiner=refiner)
    tracker = DebugTrackerRestated(refiner)         # attaches itself to the OLD optimizers, like the reference
    return loc, tracker


# ---- the reference's control flow, restated --------------------------------------------------------------------
def interp_sparse_observations(refiner, feature_maps, feature_scales, camera, p3d, pose):
    """pixloc_pose_refiners.py:327-368 (camera / pose handed in instead of read from model3d)."""
    T_w2cam = copy.deepcopy(pose)
    p3d_cam = T_w2cam * p3d
    feature_obs, masks = [], []
    for i, (feats, sc) in enumerate(zip(feature_maps, feature_scales)):
        p2d_feat, valid = camera.scale(sc).world2image(p3d_cam)
        opt = refiner.optimizer
        opt = opt[len(opt) - i - 1] if isinstance(opt, (tuple, list)) else opt
        obs, mask, _ = opt.interpolator(feats, p2d_feat.to(feats))
        assert not obs.requires_grad
        feature_obs.append(obs)
        masks.append(mask & valid.to(mask))
    mask = torch.all(torch.stack(masks, dim=0), dim=0)
    n = p3d.shape[0]
    feature_obs = [[feature_obs[i][j] for i in range(len(feature_maps))] for j in range(n)]
    return {j: feature_obs[j] for j in range(n) if mask[j]}


def refine_pose_using_features(refiner, features_query, scales_query, qcamera, T_init, features_p3d, p3d):
    """base_refiner.py:64-137."""
    import torch.nn.functional as tF
    weights_ref, features_ref = [], []
    for level in range(len(features_p3d[0])):
        feats = torch.stack([feat[level] for feat in features_p3d], dim=0)
        feats = feats.to(refiner.device)
        if refiner.conf.compute_uncertainty:
            feats, weight = feats[:, :-1], feats[:, -1:]
            weights_ref.append(weight)
        if refiner.conf.normalize_descriptors:
            feats = tF.normalize(feats, dim=1)
        assert not feats.requires_grad
        features_ref.append(feats)
    features_query = [feat.to(refiner.device) for feat in features_query]
    if refiner.conf.compute_uncertainty:
        weights_query = [feat[-1:] for feat in features_query]
        features_query = [feat[:-1] for feat in features_query]
    if refiner.conf.normalize_descriptors:
        features_query = [tF.normalize(feat, dim=0) for feat in features_query]
    T_i = T_init
    ret = {'T_init': T_init}
    for idx, level in enumerate(reversed(range(len(features_query)))):
        F_q, F_ref = features_query[level], features_ref[level]
        qcamera_feat = qcamera.scale(scales_query[level])
        W_ref_query = (weights_ref[level], weights_query[level]) if refiner.conf.compute_uncertainty else None
        opt = refiner.optimizer
        if isinstance(opt, (tuple, list)):
            opt = opt[refiner.conf.layer_indices[level]] if refiner.conf.layer_indices else opt[level]
        T_opt, fail = opt.run(p3d, F_ref, F_q, T_i.to(F_q), qcamera_feat.to(F_q), W_ref_query=W_ref_query)
        refiner.tracker.log_optim_done(i=idx, T_opt=T_opt, fail=fail, level=level, p3d=p3d, T_init=T_init,
                                       camera=qcamera_feat)
        if fail:
            return {**ret, 'success': False}
        T_i = T_opt
    T_opt = T_opt.cpu().double()
    dR, dt = (T_init.inv() @ T_opt).magnitude()
    return {**ret, 'success': True, 'T_refined': T_opt, 'diff_R': dR.item(), 'diff_t': dt.item()}


# ---- tests -------------------------------------------------------------------------------------------------------
def test_install_swaps_every_reachable_object():
    from pixtrack_b200 import install as inst
    from pixtrack_b200.extractor import B200FeatureExtractor
    from pixtrack_b200.optimizer import B200Optimizer
    g = cases.gold('refine')
    loc, tracker = build_localizer(g['consts'])
    old = list(loc.optimizer)
    new_opts, new_ext = inst.install(loc)
    assert loc.optimizer is new_opts and loc.refiner.optimizer is new_opts
    assert loc.extractor is new_ext and loc.refiner.feature_extractor is new_ext
    assert isinstance(new_ext, B200FeatureExtractor) and new_ext.conf.resize == 128 and new_ext.model.scales == [1, 4, 16]
    for o, n, c in zip(old, new_opts, g['consts']):
        assert isinstance(n, B200Optimizer) and isinstance(n, torch.nn.Module)
        np.testing.assert_array_equal(n.dampingnet.const.detach().cpu().numpy(), c)
        assert n.conf.num_iters == 150 and n.conf.interpolation.pad == 1 and n.conf.dt_stop_criteria == 5e-3
        assert n.logging_fn == tracker.log_optim_iter          # the tracker's callback moved to the new optimizers
        assert callable(n.interpolator)


def test_reference_control_flow_through_swapped_objects_lands_on_the_reference_golden():
    from pixtrack_b200 import install as inst
    from pixtrack_b200.geometry import Camera, Pose
    g = cases.gold('refine')
    cam_q, scales, maps, p3d, R_gt, t_gt = cases.pyramid_scene(1)
    np.testing.assert_allclose(cases.checksum(*maps, p3d), g['chk'], rtol=1e-9)
    loc, tracker = build_localizer(g['consts'])
    inst.install(loc)
    refiner = loc.refiner
    maps_dev = [m.to(D) for m in maps]                      # [(C+1), H, W] CHW tensors, like dense_feature_extraction returns
    camera = Camera(cam_q.double())
    # reference side (pixloc_pose_refiners.py:327-368) at the render pose, float64 like the COLMAP model
    fd = interp_sparse_observations(refiner, maps_dev, scales, camera, p3d.double().numpy(),
                                    Pose.from_Rt(R_gt, t_gt).double())
    kept = np.array(sorted(fd.keys()))
    assert np.array_equal(kept, g['kept'])
    for lv in range(3):
        obs = torch.stack([fd[i][lv] for i in kept]).cpu().numpy()
        np.testing.assert_allclose(obs, g[f'obs{lv}'], rtol=1e-5, atol=1e-6)
    # query side (base_refiner.py:64-137) from the perturbed pose the golden used
    R0, t0 = syn.perturb_pose(R_gt, t_gt, 99, 1.5, 0.015)
    feats = [tuple(fd[i]) for i in kept]
    ret = refine_pose_using_features(refiner, maps_dev, scales, Camera(cam_q), Pose.from_Rt(R0, t0).double(), feats,
                                     p3d.double().numpy()[kept])
    assert ret['success'] == bool(g['success'])
    assert tracker.num_iters == list(g['num_iters'])
    assert tracker.done == [2, 1, 0]
    np.testing.assert_allclose([float(c[-1]) for c in tracker.costs], g['last_costs'], rtol=1e-4)
    np.testing.assert_allclose(ret['T_refined']._data.numpy(), g['T_refined'], atol=2e-5)
    assert ret['T_refined']._data.dtype == torch.float64 and ret['T_refined']._data.device.type == 'cpu'
    np.testing.assert_allclose(ret['diff_t'], float(g['diff_t']), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(ret['diff_R'], float(g['diff_R']), rtol=1e-3, atol=3e-2)   # degrees, acos-limited (float32)
    # every logged pose of the last level is a Pose on the host, T chain starts at the level's T_init
    assert len(tracker.T) == sum(tracker.num_iters) + 3
    from pixtrack_b200 import _lib
    _lib.device_status(0)


def test_dense_feature_extraction_flow_through_the_swapped_extractor():
    """base_refiner.py:189-207: `features, scales, weight = self.feature_extractor(image, image_scale)` then
    `torch.cat([f, w], 0)`; downstream slicing / normalisation on the CHW views of channels-last storage."""
    import torch.nn.functional as tF
    from pixtrack_b200 import install as inst
    g = cases.gold('unet')
    loc, _ = build_localizer(cases.gold('refine')['consts'], resize=128)
    inst.install(loc)
    image = syn.textured_image(150, 200, seed=4).numpy()
    features, scales, weight = loc.refiner.feature_extractor(image, 1)
    np.testing.assert_allclose(np.array(scales), g['x_scales'])
    assert weight is not None
    features = [torch.cat([f, w], 0) for f, w in zip(features, weight)]
    for lv in range(3):
        ref = torch.cat([torch.from_numpy(g[f'x_f{lv}']), torch.from_numpy(g[f'x_c{lv}'])], 0)
        assert features[lv].shape == ref.shape
        f, w = features[lv][:-1], features[lv][-1:]
        rel = float((f.cpu() - ref[:-1]).abs().max() / ref[:-1].abs().max())
        assert rel < 2e-2, (lv, rel)
        assert float((w.cpu() - ref[-1:]).abs().max()) < 1e-2
        n = tF.normalize(f, dim=0)
        assert abs(float(n.pow(2).sum(0).mean()) - 1.0) < 1e-4
