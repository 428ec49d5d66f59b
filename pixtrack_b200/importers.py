"""Weight / snapshot importers: the files the reference loads at start-up -> the packed device tensors of
the B200 path.

* PixLoc checkpoint `outputs/training/pixloc_megadepth/checkpoint_best.tar`
  (reference pixloc/pixloc/pixlib/utils/experiments.py:58-80 `load_experiment`: `torch.load` ->
  `{'conf': {... 'model': {...}}, 'model': state_dict}`; module tree of TwoViewRefiner,
  pixloc/pixloc/pixlib/models/two_view_refiner.py:42-61: `extractor.*` and `optimizer.{level}.*`).
* instant-ngp snapshot `weights.msgpack` (instant-ngp/src/testbed.cu:2905-3001 `save_snapshot` /
  `load_snapshot`; trainer part tiny-cuda-nn/include/tiny-cuda-nn/trainer.h:255-262): msgpack of the network
  config with `snapshot.params_binary` (fp16), `snapshot.density_grid_binary` (fp16),
  `snapshot.nerf.aabb_scale`, `snapshot.nerf.dataset.{scale, offset, render_aabb}`.

Parsing runs on the host (numpy); the returned adapters need a CUDA device.
"""
from typing import Dict, List, Mapping, Optional, Tuple

import numpy as np
import torch

PIXLOC_EXTRACTOR = dict(encoder='vgg19', decoder=[64, 64, 64, 32], output_scales=[0, 2, 4], output_dim=[32, 128, 128])
# pixtrack/pose_trackers/pixloc_tracker_r9.py:43-58: what the tracker overrides on top of the checkpoint conf
R9_OPTIMIZER_OVERRIDES = dict(num_iters=150, pad=1)


def _get(conf, key, default=None):
    try:
        return conf[key]
    except (KeyError, TypeError, AttributeError):
        return getattr(conf, key, default)


def split_pixloc_checkpoint(ckpt: Mapping, optimizer_overrides: Optional[dict] = None):
    """-> (extractor state dict without the `extractor.` prefix, optimizer conf dict, [damping const per level])."""
    model_conf = _get(_get(ckpt, 'conf'), 'model')
    econf = _get(model_conf, 'extractor', {})
    for k, want in PIXLOC_EXTRACTOR.items():
        got = _get(econf, k)
        if got is not None and (list(got) if isinstance(want, list) else got) != want:
            raise NotImplementedError(f'extractor.{k} = {got}: the B200 plan implements the PixLoc UNet ({want})')
    oc = _get(model_conf, 'optimizer', {})
    oconf = {k: _get(oc, k) for k in ('num_iters', 'loss_fn', 'jacobi_scaling', 'normalize_features', 'lambda_',
                                      'grad_stop_criteria', 'dt_stop_criteria', 'dR_stop_criteria', 'pad', 'learned_damping')
             if _get(oc, k) is not None}
    damping = _get(oc, 'damping')
    if damping is not None:
        oconf['damping'] = {k: (list(v) if k == 'log_range' else v) for k, v in dict(damping).items()}
    interp = _get(oc, 'interpolation')
    if interp is not None:
        oconf['interpolation'] = dict(interp)
    oconf.update(optimizer_overrides or {})
    sd = _get(ckpt, 'model')
    ext = {k[len('extractor.'):]: v for k, v in sd.items() if k.startswith('extractor.')}
    consts = []
    lv = 0
    while f'optimizer.{lv}.dampingnet.const' in sd:
        consts.append(sd[f'optimizer.{lv}.dampingnet.const'])
        lv += 1
    if not consts and 'optimizer.dampingnet.const' in sd:          # duplicate_optimizer_per_scale = false
        consts = [sd['optimizer.dampingnet.const']] * 3
    if not ext or not consts:
        raise ValueError('not a PixLoc TwoViewRefiner checkpoint (no extractor.* / optimizer.*.dampingnet.const keys)')
    return ext, oconf, consts


def load_pixloc_checkpoint(path_or_ckpt, device, optimizer_overrides: Optional[dict] = R9_OPTIMIZER_OVERRIDES,
                           preprocessing: Optional[dict] = None):
    """-> (B200FeatureExtractor, [B200Optimizer per level]) ready to be handed to a PoseTrackerRefiner."""
    from .extractor import B200FeatureExtractor
    from .optimizer import B200Optimizer
    ckpt = path_or_ckpt
    if not isinstance(ckpt, Mapping):
        ckpt = torch.load(str(path_or_ckpt), map_location='cpu', weights_only=False)
    ext_sd, oconf, consts = split_pixloc_checkpoint(ckpt, optimizer_overrides)
    extractor = B200FeatureExtractor(ext_sd, device, preprocessing)
    opts: List = []
    for c in consts:
        o = B200Optimizer(oconf)
        o.dampingnet.const.data.copy_(torch.as_tensor(c, dtype=torch.float32))
        opts.append(o.to(device))
    return extractor, opts


# ------------------------------------------------------------------------------------------------
def _binary(x) -> bytes:
    if isinstance(x, (bytes, bytearray, memoryview)):
        return bytes(x)
    if isinstance(x, Mapping) and 'bytes' in x:            # nlohmann's JSON form of a binary value
        return bytes(bytearray(x['bytes']))
    if hasattr(x, 'data'):                                  # msgpack.ExtType
        return bytes(x.data)
    raise ValueError(f'cannot read a binary blob from {type(x)}')


def _box(b) -> Optional[np.ndarray]:
    if b is None:
        return None
    if isinstance(b, Mapping):
        return np.array([b['min'], b['max']], np.float32)
    a = np.asarray(b, np.float32)
    return a.reshape(2, 3)


def read_ingp_snapshot(path_or_bytes) -> Dict[str, object]:
    """-> dict(params fp16 [n], density_grid fp16 [(max_cascade+1)*128^3], aabb_scale, scale, offset, render_aabb)."""
    import msgpack
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(str(path_or_bytes), 'rb').read()
    cfg = msgpack.unpackb(bytes(raw), raw=False, strict_map_key=False)
    if 'snapshot' not in cfg:
        raise ValueError('file does not contain a snapshot')                      # testbed.cu:2943-2945
    snap = cfg['snapshot']
    if snap.get('density_grid_size', 128) != 128:
        raise ValueError('Incompatible grid size.')                               # testbed.cu:2959-2961
    if snap.get('params_type', '__half') != '__half':
        raise NotImplementedError('only fp16 snapshots (the instant-ngp default) are supported')
    params = np.frombuffer(_binary(snap['params_binary']), np.float16).copy()
    if 'n_params' in snap and int(snap['n_params']) != params.size:
        raise ValueError('params_binary does not hold n_params values')
    density = np.frombuffer(_binary(snap['density_grid_binary']), np.float16).copy()
    nerf = snap.get('nerf', {})
    ds = nerf.get('dataset', {}) or {}
    aabb_scale = int(nerf.get('aabb_scale', ds.get('aabb_scale', 1)))
    n = 128 ** 3
    if density.size % n != 0 or density.size // n != aabb_scale.bit_length():
        # testbed.cu:2984-2986: the grid must have max_cascade + 1 = log2(aabb_scale) + 1 cascades
        raise ValueError('Incompatible number of grid cascades.')
    enc = cfg.get('encoding', {})
    for k, want in (('n_levels', 16), ('n_features_per_level', 2), ('log2_hashmap_size', 19), ('base_resolution', 16)):
        if k in enc and int(enc[k]) != want:
            raise NotImplementedError(f'encoding.{k} = {enc[k]}: the render kernel implements configs/nerf/base.json')
    return dict(params=params, density_grid=density, aabb_scale=aabb_scale, scale=float(ds.get('scale', 0.33)),
                offset=tuple(float(v) for v in ds.get('offset', (0.5, 0.5, 0.5))), render_aabb=_box(ds.get('render_aabb')),
                aabb=_box(snap.get('aabb')))


def load_ingp_snapshot(path_or_bytes, device, render_aabb=None):
    """-> NerfTestbed, the counterpart of `initialize_ingp(snapshot_path, aabb)` (pixtrack/utils/ingp_utils.py:22-44,
    which also sets render_aabb to the object box and rendering_min_transmittance = 1e-7)."""
    from .nerf import NerfTestbed
    s = read_ingp_snapshot(path_or_bytes)
    tb = NerfTestbed.from_snapshot_arrays(s['params'], s['density_grid'], s['aabb_scale'], device, scale=s['scale'],
                                          offset=s['offset'])
    box = render_aabb if render_aabb is not None else s['render_aabb']
    if box is not None and np.all(np.asarray(box)[1] > np.asarray(box)[0]):
        tb.render_aabb.min, tb.render_aabb.max = np.asarray(box[0], np.float32), np.asarray(box[1], np.float32)
    tb.nerf.rendering_min_transmittance = 1e-7
    return tb
