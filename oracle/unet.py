"""CPU oracle for the PixLoc UNet feature extractor (VGG19 encoder).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional PyTorch fp32
restatement driven by a plain state dict, following
/root/reference/pixloc/pixloc/pixlib/models/unet.py:15-190 with the PixLoc
configuration (pixlib/configs/train_pixloc_megadepth.yaml:22-31): vgg19
encoder split into 5 blocks at the max-pools, 4 decoder blocks
[64,64,64,32], 1x1 adaptation heads at scales 0/2/4 with dims 32/128/128 and
1x1 uncertainty heads; and the pre-processing of
/root/reference/pixtrack/localization/feature_extractor.py:34-59.
"""
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as tF

Tensor = torch.Tensor
IMAGENET_MEAN = (0.485, 0.456, 0.406)      # unet.py:64-65
IMAGENET_STD = (0.229, 0.224, 0.225)
N_BLOCKS = 5
OUTPUT_SCALES = (0, 2, 4)


def _block_convs(sd: Dict[str, Tensor], b: int) -> List[str]:
    keys = sorted((int(k.split('.')[2]) for k in sd if k.startswith(f'encoder.{b}.') and k.endswith('.weight')))
    return [f'encoder.{b}.{i}' for i in keys]


def encoder(sd: Dict[str, Tensor], x: Tensor) -> List[Tensor]:
    """unet.py:68-99 (block packing) and :163-167 (forward): each block after
    the first opens with a 2x2/stride-2 max-pool (floor mode), then
    conv3x3(pad 1)+ReLU repeated."""
    skips = []
    for b in range(N_BLOCKS):
        if b > 0:
            x = tF.max_pool2d(x, 2, 2)
        for name in _block_convs(sd, b):
            x = tF.relu(tF.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], padding=1))
        skips.append(x)
    return skips


def decoder_block(sd: Dict[str, Tensor], i: int, prev: Tensor, skip: Tensor) -> Tensor:
    """unet.py:15-44: x2 bilinear upsample (align_corners=False), crop the skip
    to the upsampled size, concat [upsampled, skip], conv3x3 (no bias),
    BatchNorm in eval mode, ReLU."""
    up = tF.interpolate(prev, scale_factor=2, mode='bilinear', align_corners=False)
    skip = skip[:, :, :up.shape[2], :up.shape[3]]
    p = f'decoder.{i}.layers.'
    y = tF.conv2d(torch.cat([up, skip], 1), sd[p + '0.weight'], None, padding=1)
    y = tF.batch_norm(y, sd[p + '1.running_mean'], sd[p + '1.running_var'],
                      sd[p + '1.weight'], sd[p + '1.bias'], training=False, eps=1e-5)
    return tF.relu(y)


def unet_forward(sd: Dict[str, Tensor], image01: Tensor) -> Tuple[List[Tensor], List[Tensor]]:
    """image01: [1,3,H,W] in 0..1.  Returns (feature maps, confidences), fine
    to coarse, as unet.py:158-190; confidence = sigmoid(-uncertainty)."""
    mean = image01.new_tensor(IMAGENET_MEAN)[:, None, None]
    std = image01.new_tensor(IMAGENET_STD)[:, None, None]
    skips = encoder(sd, (image01 - mean) / std)
    pre = [skips[-1]]
    for i, skip in enumerate(skips[:-1][::-1]):
        pre.append(decoder_block(sd, i, pre[-1], skip))
    pre = pre[::-1]
    feats, confs = [], []
    for idx, s in enumerate(OUTPUT_SCALES):
        feats.append(tF.conv2d(pre[s], sd[f'adaptation.{idx}.0.weight'], sd[f'adaptation.{idx}.0.bias']))
        unc = tF.conv2d(pre[s], sd[f'uncertainty.{idx}.0.weight'], sd[f'uncertainty.{idx}.0.bias'])
        confs.append(torch.sigmoid(-unc))
    return feats, confs


def resize_max_edge(image: np.ndarray, target: int) -> Tuple[np.ndarray, Tuple[float, float]]:
    """pixlib/datasets/view.py:31-48 with fn=max, interp='linear'."""
    import cv2
    h, w = image.shape[:2]
    s = target / max(h, w)
    h2, w2 = int(round(h * s)), int(round(w * s))
    return cv2.resize(image, (w2, h2), interpolation=cv2.INTER_LINEAR), (s, s)


def extract(sd: Dict[str, Tensor], image: np.ndarray, scale_image: int = 1, resize: int = 1024, device=None):
    """PixTrackFeatureExtractor.__call__ (feature_extractor.py:34-59): resize
    only if the longer edge exceeds resize//scale_image; /255; HWC->CHW;
    returns (features [C,H,W] list, scales list, confidences [1,H,W] list).
    `device`: where the network runs (`image_tensor.to(self.device)`, :47); `sd` must live there."""
    sr = (1.0, 1.0)
    target = resize // scale_image if resize is not None else None
    if resize is not None and max(image.shape[:2]) > target:
        image, sr = resize_max_edge(image, target)
    x = torch.from_numpy(np.ascontiguousarray(image.transpose(2, 0, 1)) / 255.).float()[None]
    if device is not None:
        x = x.to(device)
    feats, confs = unet_forward(sd, x)
    scales = [(sr[0] / 2 ** s, sr[1] / 2 ** s) for s in OUTPUT_SCALES]
    return [f[0] for f in feats], scales, [c[0] for c in confs]
