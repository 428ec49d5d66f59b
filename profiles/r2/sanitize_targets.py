"""Targets for `compute-sanitizer --tool memcheck` beyond smoke(): the kernels added or changed in round 2 -- every
convolution kernel variant the dispatch reaches by default (per-tap K split, single-CTA halo, 16x16 pairs, row-tile pairs,
narrow-map pairs; ragged and odd sizes, fused pool), the cp.async head / pipelined conv1 kernels inside a small extractor plan, the
LM launch on its own workspace with two CTAs per SM, the NeRF network entry point and the overlay."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, 'tests'))
import synthetic as syn  # noqa: E402
from pixtrack_b200.extractor import B200FeatureExtractor, conv_f16, pack_conv3x3  # noqa: E402
from pixtrack_b200.nerf import NerfTestbed, occupancy_bitfield  # noqa: E402
from pixtrack_b200.optimizer import lm_run_batched, query_map_to_hwc  # noqa: E402
from pixtrack_b200.overlay import overlay  # noqa: E402

torch.set_grad_enabled(False)
D = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
# (cin, cout, H, W, pool): per-tap K split; single-CTA halo; 16x16 pairs streamed / resident (N = 128, 64, 32) with the fused
# pool and ragged odd sizes; row-tile pairs R = 2 (pooled) / R = 3 and N = 64; the narrow-map pair kernel
for cin, cout, H, W, pool in ((512, 512, 36, 64, False), (1024, 64, 72, 128, False), (64, 64, 170, 200, False),
                              (256, 256, 145, 255, True), (64, 128, 289, 511, False), (128, 32, 577, 1023, False),
                              (192, 64, 287, 513, False), (512, 512, 72, 128, True), (512, 512, 94, 126, False),
                              (320, 64, 188, 252, False), (512, 512, 47, 63, False)):
    x = torch.randn(H, W, cin, generator=g).half().to(D)
    w = pack_conv3x3((torch.randn(cout, cin, 3, 3, generator=g) / 50).half().to(D))
    b = torch.randn(cout, generator=g).to(D)
    y = conv_f16(x, w, b, pool=pool)
    torch.cuda.synchronize()
    assert bool(torch.isfinite((y[0] if pool else y).float()).all())
ext = B200FeatureExtractor(syn.unet_weights(0), D, dict(resize=None))
f, c, _ = ext.extract_device(syn.textured_image(96, 160, seed=1).to(D), normalize=True)
torch.cuda.synchronize()
p = syn.level_problem(seed=7, N=3000, C=128, H=144, W=256, B=5)
T0 = torch.cat([p['R0'].reshape(5, 9), p['t0']], 1).to(D)
lam = (10.0 ** (-6.0 + torch.sigmoid(torch.zeros(6)) * 11.0)).to(D)
out = lm_run_batched(p['p3d'].to(D), p['F_ref'].to(D), query_map_to_hwc(p['F_q'].to(D)), T0, p['cam'].to(D), lam,
                     p['W_ref'].reshape(5, -1).to(D), p['W_q'].to(D), num_iters=10)
torch.cuda.synchronize()
sc = syn.nerf_scene(2, 1)
tb = NerfTestbed(sc['grid'], sc['w_density'], sc['w_rgb'], occupancy_bitfield(sc['density_grid'], sc['max_cascade']), 1, D)
o = tb.network(torch.rand(1000, 3), torch.nn.functional.normalize(torch.randn(1000, 3), dim=1))
q = torch.randint(0, 256, (90, 130, 3), dtype=torch.uint8, device=D)
r = overlay(q, q.flip(0).contiguous(), axes_px=np.array([[10, 10], [80, 40], [10, 10], [10, 70], [10, 10], [120, 85]], np.int16))
torch.cuda.synchronize()
print('sanitize targets ok', float(out[0].abs().sum()), float(o.abs().sum()), int(r.sum()))
