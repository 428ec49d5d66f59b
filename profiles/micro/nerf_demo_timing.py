import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import tracker_demo as td
from pixtrack_b200.nerf import get_nerf_image
from pixtrack_b200.tracker import sfm_to_nerf_pose, camera_in_world_from_pose
tb, cam_q, trk = td.build(n_points=2000)
eng = trk.engine
pose = td.orbit_pose(0.0)
npose = sfm_to_nerf_pose(td.N2S, camera_in_world_from_pose(pose))
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, cam in (('ref 1008x756', eng.camera_r), ('query 1920x1080', eng.camera_q)):
    for depth in (False, True):
        print(name, 'depth' if depth else 'shade', f'{timed(lambda: get_nerf_image(tb, npose, cam, depth=depth, device_output=True)):.2f} ms', flush=True)
img = td.query_frame(tb, cam_q, pose)
feat = eng.create_reference(pose, [3])
print('refine', f'{timed(lambda: eng.refine("q", img, cam_q, pose, 3, [1], feat)):.2f} ms')
print('mask_query', f'{timed(lambda: eng.mask_query(img, pose)):.2f} ms')
