"""Short driver for ncu: the fused LM kernel alone at config-4 scale (C=128 144x256, N=20000, B=16,
30 fixed iterations) and at the C2 level-2 shape.  Used by the ncu commands recorded in profiles/README.md."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import os as _os, sys as _sys  # noqa: E401,E402
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'tests'))  # scene generators live with the tests
import synthetic as syn  # noqa: E402
from pixtrack_b200.optimizer import LmLaunch, query_map_to_hwc  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
lam = (10.0 ** (-6.0 + torch.sigmoid(torch.zeros(6)) * 11.0)).to(dev)
B = 16
p = syn.level_problem(seed=9, N=20000, C=128, H=144, W=256, B=B, noise=0.02, rot_deg=0.5, trans=0.005)
T0 = torch.cat([p['R0'].reshape(B, 9), p['t0']], 1).to(dev)
L = LmLaunch(p['p3d'].to(dev), p['F_ref'].to(dev), query_map_to_hwc(p['F_q'].to(dev)), T0, p['cam'].to(dev), lam,
             p['W_ref'].reshape(B, -1).to(dev), p['W_q'].to(dev), num_iters=30, grad_stop=0.0, dt_stop=0.0, dR_stop=0.0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    L.launch()
torch.cuda.synchronize()
print('ok', L.plan(), float(L.log[:, :, 1].sum()))
