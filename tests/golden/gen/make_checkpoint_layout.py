"""Generate tests/golden/pixloc_checkpoint_layout.json from the UNMODIFIED reference model classes.

Run in the authoring container only (needs /root/reference):

    python tests/golden/gen/make_checkpoint_layout.py

The trained `checkpoint_best.tar` is not available offline, but its layout is whatever
`TwoViewRefiner(conf).state_dict()` produces for the pixloc_megadepth configuration
(pixloc/pixloc/pixlib/models/two_view_refiner.py:25-61, unet.py, learned_optimizer.py) and
`pixlib/utils/experiments.py:58-80` saves: {'model': state_dict, 'conf': ...}.  This script instantiates the
reference model with that configuration (random weights) and records every state-dict key with its shape and
dtype; tests/test_importers.py builds a checkpoint with exactly these entries and runs it through the importer.
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, '/root/reference/pixloc', '/root/reference']
sys.modules['h5py'] = types.ModuleType('h5py')
import torch  # noqa: E402

_six = types.ModuleType('torch._six')
_six.string_classes = (str, bytes)
sys.modules['torch._six'] = _six
import torchvision  # noqa: E402

_vgg19 = torchvision.models.vgg19
torchvision.models.vgg19 = lambda pretrained=False, **kw: _vgg19(weights=None)

from pixloc.pixlib.models.two_view_refiner import TwoViewRefiner  # noqa: E402

# the pixloc_megadepth experiment (pixloc/pixloc/pixlib/configs/train_pixloc_megadepth.yaml: model section)
CONF = dict(
    extractor=dict(name='unet', encoder='vgg19', decoder=[64, 64, 64, 32], output_scales=[0, 2, 4], output_dim=[32, 128, 128],
                   freeze_batch_normalization=False, do_average_pooling=False, compute_uncertainty=True, checkpointed=True),
    optimizer=dict(name='learned_optimizer', num_iters=15, pad=2, lambda_=0.01, verbose=False, loss_fn='scaled_barron(0, 0.1)',
                   jacobi_scaling=False, learned_damping=True, damping=dict(type='constant')),
    duplicate_optimizer_per_scale=True, success_thresh=3, clamp_error=7, normalize_features=True, normalize_dt=False,
)

if __name__ == '__main__':
    model = TwoViewRefiner(CONF)
    sd = model.state_dict()
    layout = [[k, list(v.shape), str(v.dtype).replace('torch.', '')] for k, v in sd.items()]
    path = os.path.join(HERE, '..', 'pixloc_checkpoint_layout.json')
    with open(path, 'w') as f:
        json.dump({'conf_model': CONF, 'state_dict': layout}, f, indent=0)
    print('wrote', os.path.abspath(path), len(layout), 'entries;', sum(1 for k, _, _ in layout if k.startswith('extractor.')), 'extractor,',
          [k for k, _, _ in layout if k.startswith('optimizer')])
