"""Short driver for `ncu --metrics gpu__time_duration.sum`: the C2 frame loop of bench.py (reference refresh + query
extraction + LM chain, images resident in HBM) without the CPU / library legs, so that the launch list holds only the
frame's own kernels.   python profiles/r2/frame_launches.py [frames] [workload c2|c5]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, 'tests'))
os.environ.setdefault('PTK_PLAN_GRAPH', '0')      # kernel-by-kernel launches: ncu serialises them anyway, and names stay visible
import bench  # noqa: E402
import synthetic as syn  # noqa: E402
from pixtrack_b200.extractor import B200FeatureExtractor  # noqa: E402
from pixtrack_b200.pipeline import FrameTracker  # noqa: E402

torch.set_grad_enabled(False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wl = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else 'c2']
dev = torch.device('cuda:0')
seq = bench.make_sequence(wl, 100)
ext = B200FeatureExtractor(syn.unet_weights(0), dev)
lam = bench.lam0().to(dev)
fr0 = seq['frames'][0]
trk = FrameTracker(ext, fr0['img_q'].shape[:2], seq['cam_q'], seq['p3d'], [lam] * 3, wl['n_views'], use_graph=False, **wl['stop'])
imgs = [dict(q=f['img_q'].to(dev), r=f['img_r'].to(dev)) for f in seq['frames']]
T_ref = [torch.cat([f['R_r'].reshape(-1), f['t_r']]) for f in seq['frames']]
T0 = [f['T_init'].to(dev) for f in seq['frames']]
tbs = bench.make_nerf_objects(dev, 1, 11) if wl['nerf'] else None
for v in range(wl['n_views']):
    trk.refresh_reference(v, imgs[0]['r'], seq['cam_r'], T_ref[0])
torch.cuda.synchronize()
for i in range(n):
    k = i % bench.RING
    if tbs is None:
        trk.refresh_reference(i % wl['n_views'], imgs[k]['r'], seq['cam_r'], T_ref[k])
        trk.track(imgs[k]['q'], T0[k])
    else:
        depth, ref = bench.nerf_renders(tbs[0], i)
        trk.refresh_reference(i % wl['n_views'], ref, seq['cam_r'], T_ref[k])
        trk.track(imgs[k]['q'], T0[k], mask_depth=depth)
torch.cuda.synchronize()
print('ok', n, 'frames')
