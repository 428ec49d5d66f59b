"""LM kernel at C4 level-1 scale (C=128, 144x256, N=20000, B=16, 30 fixed iterations) with the model points in random
order and in Morton order (pixtrack_b200.pipeline.morton_order): the effect of L1 locality on the L2-bound gather.
   python profiles/r2/lm_order.py [reps] [mode: both|random|morton]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, 'tests'))
import synthetic as syn  # noqa: E402
from pixtrack_b200.optimizer import LmLaunch, query_map_to_hwc  # noqa: E402
from pixtrack_b200.pipeline import morton_order  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
mode = sys.argv[2] if len(sys.argv) > 2 else 'both'
lam = (10.0 ** (-6.0 + torch.sigmoid(torch.zeros(6)) * 11.0)).to(dev)
out = {}
for (C, H, W, tag) in ((128, 144, 256, 'L1'), (32, 576, 1024, 'L0'), (128, 36, 64, 'L2')):
    B = 16
    p = syn.level_problem(seed=9, N=20000, C=C, H=H, W=W, B=B, noise=0.02, rot_deg=0.5, trans=0.005,
                          level_scale=(1024 / 1920) / {144: 4, 576: 1, 36: 16}[H], sigma={144: 2.0, 576: 4.0, 36: 1.0}[H])
    T0 = torch.cat([p['R0'].reshape(B, 9), p['t0']], 1).to(dev)
    for order_name in ('random', 'morton'):
        if mode not in ('both', order_name):
            continue
        o = morton_order(p['p3d']) if order_name == 'morton' else torch.arange(20000)
        L = LmLaunch(p['p3d'][o].to(dev), p['F_ref'][:, o].contiguous().to(dev), query_map_to_hwc(p['F_q'].to(dev)), T0,
                     p['cam'].to(dev), lam, p['W_ref'].reshape(B, -1)[:, o].contiguous().to(dev), p['W_q'].to(dev),
                     num_iters=30, grad_stop=0.0, dt_stop=0.0, dR_stop=0.0)
        for _ in range(3):
            L.launch()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            L.launch()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        nv = float(L.log[:, :, 1].sum())
        out[f'{tag}_{order_name}'] = dict(ms=ms, us_per_iter=1e3 * ms / 30, algorithmic_tbs=nv * (52 * C + 32) / ms / 1e9,
                                          T_checksum=float(L.T.double().abs().sum()), plan=L.plan())
print(json.dumps(out))
