"""Importers: PixLoc checkpoint layout (pinned on the reference model's own state dict, last test) and instant-ngp
msgpack snapshot layout (synthetic files written in the reference's format; the real assets are not in the
container)."""
import numpy as np
import pytest
import torch

import synthetic as syn
from pixtrack_b200 import importers


def _checkpoint():
    sd = {f'extractor.{k}': v for k, v in syn.unet_weights(0).items()}
    for lv in range(3):
        sd[f'optimizer.{lv}.dampingnet.const'] = torch.full((6,), 0.1 * lv)
    conf = {'model': {'name': 'two_view_refiner', 'duplicate_optimizer_per_scale': True,
                      'extractor': {'name': 'unet', 'encoder': 'vgg19', 'decoder': [64, 64, 64, 32],
                                    'output_scales': [0, 2, 4], 'output_dim': [32, 128, 128], 'compute_uncertainty': True},
                      'optimizer': {'name': 'learned_optimizer', 'num_iters': 15, 'pad': 2, 'lambda_': 0.01,
                                    'loss_fn': 'scaled_barron(0, 0.1)', 'jacobi_scaling': False, 'learned_damping': True,
                                    'damping': {'type': 'constant'}}}}
    return {'conf': conf, 'model': sd, 'epoch': 3}


def test_split_pixloc_checkpoint(tmp_path):
    ck = _checkpoint()
    torch.save(ck, tmp_path / 'checkpoint_best.tar')
    ck2 = torch.load(tmp_path / 'checkpoint_best.tar', map_location='cpu', weights_only=False)
    ext, oconf, consts = importers.split_pixloc_checkpoint(ck2, importers.R9_OPTIMIZER_OVERRIDES)
    assert 'encoder.0.0.weight' in ext and 'adaptation.2.0.bias' in ext and not any(k.startswith('extractor.') for k in ext)
    assert oconf['num_iters'] == 150 and oconf['pad'] == 1 and oconf['loss_fn'] == 'scaled_barron(0, 0.1)'
    assert len(consts) == 3 and float(consts[2][0]) == pytest.approx(0.2)
    from pixtrack_b200.optimizer import B200Optimizer
    opt = B200Optimizer(oconf)          # the merged conf is accepted as is (legacy `pad` key, base_model.py:69-72)
    assert opt.conf.interpolation.pad == 1 and opt.conf.num_iters == 150 and opt.loss_scale == pytest.approx(0.1)
    ck['conf']['model']['extractor']['decoder'] = [64, 64]
    with pytest.raises(NotImplementedError):
        importers.split_pixloc_checkpoint(ck)
    with pytest.raises(ValueError):
        importers.split_pixloc_checkpoint({'conf': {'model': {}}, 'model': {'foo': torch.zeros(1)}})


def _snapshot_bytes(sc, as_json_binary=False):
    import msgpack
    params = np.concatenate([w.ravel() for w in (*sc['w_density'], *sc['w_rgb'])] + [sc['grid'].ravel()]).astype(np.float16)
    dens = sc['density_grid'].astype(np.float16)

    def blob(a):
        return {'bytes': list(a.tobytes()), 'subtype': None} if as_json_binary else a.tobytes()
    cfg = {'encoding': {'otype': 'HashGrid', 'n_levels': 16, 'n_features_per_level': 2, 'log2_hashmap_size': 19,
                        'base_resolution': 16},
           'snapshot': {'version': 1, 'n_params': int(params.size), 'params_type': '__half', 'params_binary': blob(params),
                        'density_grid_size': 128, 'density_grid_binary': blob(dens), 'training_step': 35000, 'loss': 0.001,
                        'aabb': {'min': [0.0, 0.0, 0.0], 'max': [1.0, 1.0, 1.0]}, 'bounding_radius': 1.0,
                        'nerf': {'aabb_scale': sc['aabb_scale'],
                                 'dataset': {'scale': 0.25, 'offset': [0.5, 0.4, 0.5], 'aabb_scale': sc['aabb_scale'],
                                             'render_aabb': {'min': [0.2, 0.2, 0.2], 'max': [0.8, 0.8, 0.8]}}}}}
    return msgpack.packb(cfg, use_bin_type=True), params, dens


def test_read_ingp_snapshot_round_trip(tmp_path):
    sc = syn.nerf_scene(4, 1, zero_network=True)      # small to serialise: zero grid, real occupancy
    raw, params, dens = _snapshot_bytes(sc)
    (tmp_path / 'weights.msgpack').write_bytes(raw)
    s = importers.read_ingp_snapshot(tmp_path / 'weights.msgpack')
    assert np.array_equal(s['params'], params) and np.array_equal(s['density_grid'], dens)
    assert s['aabb_scale'] == 1 and s['scale'] == 0.25 and s['offset'] == (0.5, 0.4, 0.5)
    assert np.allclose(s['render_aabb'], [[0.2] * 3, [0.8] * 3])
    from pixtrack_b200.nerf import split_params
    wd, wc, grid = split_params(s['params'], s['aabb_scale'])
    assert np.array_equal(wd[1], sc['w_density'][1]) and np.array_equal(grid, sc['grid'])


def test_snapshot_errors():
    import msgpack
    with pytest.raises(ValueError):
        importers.read_ingp_snapshot(msgpack.packb({'encoding': {}}))
    sc = syn.nerf_scene(4, 1, zero_network=True)
    raw, _, _ = _snapshot_bytes(sc)
    cfg = msgpack.unpackb(raw, raw=False)
    cfg['snapshot']['nerf']['aabb_scale'] = 4          # 1 cascade in the file, 3 needed
    with pytest.raises(ValueError):
        importers.read_ingp_snapshot(msgpack.packb(cfg, use_bin_type=True))
    cfg['snapshot']['nerf']['aabb_scale'] = 1
    cfg['encoding']['n_levels'] = 8
    with pytest.raises(NotImplementedError):
        importers.read_ingp_snapshot(msgpack.packb(cfg, use_bin_type=True))


@pytest.mark.gpu
def test_loaders_build_working_adapters(tmp_path):
    from oracle import nerf as onerf
    ext, opts = importers.load_pixloc_checkpoint(_checkpoint(), 'cuda:0')
    assert len(opts) == 3 and opts[1].dampingnet.const.is_cuda and float(opts[1].dampingnet.const.detach()[0]) == pytest.approx(0.1)
    img = syn.textured_image(120, 160, seed=1).numpy().astype(np.float32)
    feats, scales, confs = ext(img)
    # 120 -> 60 -> 30 -> 15 -> 7 down, x2 up with cropped skips (unet.py:39-43): 14, 28, 56, 112
    assert [tuple(f.shape) for f in feats] == [(32, 112, 160), (128, 28, 40), (128, 7, 10)]
    sc = syn.nerf_scene(4, 1)
    raw, _, _ = _snapshot_bytes(sc)
    (tmp_path / 'weights.msgpack').write_bytes(raw)
    tb = importers.load_ingp_snapshot(tmp_path / 'weights.msgpack', 'cuda:0')
    assert tb.scale == 0.25 and np.allclose(tb.render_aabb.min, 0.2) and tb.nerf.rendering_min_transmittance == 1e-7
    cam = syn.nerf_look_at((0.5, -0.9, 0.6))
    tb.fov = 50.0
    tb.set_ngp_camera_matrix(cam)
    got = tb.render(32, 24, 2)
    bits = onerf.bitfield_from_density_grid(sc['density_grid'].astype(np.float16).astype(np.float32), 0)
    m = onerf.NerfModel(1, sc['grid'], sc['w_density'], sc['w_rgb'], bits, scale=0.25, offset=(0.5, 0.4, 0.5),
                        render_aabb=np.array([[0.2] * 3, [0.8] * 3], np.float32))
    ref = onerf.render(m, cam, 32, 24, 50.0, spp=2)['rgba']
    assert ref[..., 3].max() > 0.9
    assert (np.abs(got - ref).max(-1) > 4e-3).mean() < 0.02


def test_importer_consumes_the_reference_models_state_dict_layout():
    """tests/golden/pixloc_checkpoint_layout.json lists every key / shape / dtype of
    `TwoViewRefiner(pixloc_megadepth conf).state_dict()` from the unmodified reference
    (tests/golden/gen/make_checkpoint_layout.py).  A checkpoint with exactly those entries goes through
    split_pixloc_checkpoint and the weight packer, every floating-point extractor tensor is consumed, and the synthetic
    weights the other tests use have the same layout."""
    import json
    import os
    from pixtrack_b200.extractor import pack_weights
    fix = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'pixloc_checkpoint_layout.json')))
    g = torch.Generator().manual_seed(0)
    sd = {}
    for key, shape, dtype in fix['state_dict']:
        if dtype == 'int64':
            sd[key] = torch.zeros(shape, dtype=torch.int64)
        elif key.endswith('running_var'):
            sd[key] = torch.rand(shape, generator=g) + 0.5
        else:
            sd[key] = torch.randn(shape, generator=g) * 0.05
    ext, oconf, consts = importers.split_pixloc_checkpoint({'conf': {'model': fix['conf_model']}, 'model': sd},
                                                           importers.R9_OPTIMIZER_OVERRIDES)
    assert len(consts) == 3 and all(tuple(c.shape) == (6,) for c in consts)
    assert oconf['num_iters'] == 150 and oconf['pad'] == 1 and oconf['lambda_'] == 0.01 and oconf['damping'] == {'type': 'constant'}
    packed = pack_weights(ext, 'cpu')
    assert len(packed['conv_w']) == len(packed['conv_b']) == 20 and len(packed['head_w']) == 3
    assert [tuple(w.shape) for w in packed['head_w']] == [(33, 32), (129, 64), (129, 512)]
    assert tuple(packed['conv_w'][0].shape) == (64, 28) and tuple(packed['conv_w'][-1].shape) == (9, 32, 128)
    # every floating-point tensor of the extractor is used by the packer (decoder convs carry no bias: BatchNorm follows)
    used = set()
    for b, idxs in enumerate(([0, 2], [1, 3], [1, 3, 5, 7], [1, 3, 5, 7], [1, 3, 5, 7])):
        for i in idxs:
            used |= {f'encoder.{b}.{i}.weight', f'encoder.{b}.{i}.bias'}
    for i in range(4):
        used |= {f'decoder.{i}.layers.0.weight'} | {f'decoder.{i}.layers.1.{n}' for n in ('weight', 'bias', 'running_mean', 'running_var')}
    for l in range(3):
        used |= {f'{h}.{l}.0.{n}' for h in ('adaptation', 'uncertainty') for n in ('weight', 'bias')}
    assert used == {k for k, v in ext.items() if v.is_floating_point()}
    # the synthetic generator mirrors the layout exactly
    syn_sd = syn.unet_weights(0)
    assert {k: tuple(v.shape) for k, v in syn_sd.items()} == {k: tuple(v.shape) for k, v in ext.items()}
