import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import tracker_demo as td
from pixtrack_b200.tracker import PoseRt
tb, cam_q, trk = td.build()
eng = trk.engine
gt = td.orbit_pose(1.0)
img = td.query_frame(tb, cam_q, gt)
print('query image mean', float(img.float().mean()), 'nonzero frac', float((img != 0).any(-1).float().mean()))
feat = eng.create_reference(gt, [3])
print('ref image', tuple(feat['image'].shape), float(feat['image'].float().mean()), 'std', float(feat['image'].float().std()))
for init_yaw in (1.0, 0.0, 3.0):
    init = td.orbit_pose(init_yaw)
    out = eng.refine('q', img, cam_q, init, 3, [1], feat)
    t = eng._tracker(1)
    print('init yaw', init_yaw, 'success', out['success'], 'costs', out['costs'], 'n_iters', [int(x[0]) for x in t.plan.n_iters],
          'valid', int(t.valid.sum()), 'of', t.n_active)
    if out['success']:
        print('   err vs gt', (gt.inv() @ out['T_refined']).magnitude(), ' moved', out['diff_R'], out['diff_t'])
