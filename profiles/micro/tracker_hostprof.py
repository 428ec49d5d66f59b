import cProfile, os, pstats, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import tracker_demo as td
tb, cam_q, trk = td.build(n_points=2000)
gt = td.orbit_pose(0.0)
img = td.query_frame(tb, cam_q, gt)
for f in range(3):
    trk.run_single_frame((f'w{f}.png', img))
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for f in range(10):
    trk.run_single_frame((f'p{f}.png', img))
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
