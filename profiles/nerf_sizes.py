import sys; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
import os as _os, sys as _sys  # noqa: E401,E402
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'tests'))  # scene generators live with the tests
import synthetic as syn  # noqa: E402
from pixtrack_b200.nerf import NerfTestbed, occupancy_bitfield
sc = syn.nerf_scene(11, 2)
bits = occupancy_bitfield(sc['density_grid'], sc['max_cascade'])
tb = NerfTestbed(sc['grid'], sc['w_density'], sc['w_rgb'], bits, 2, 'cuda:0')
tb.nerf.rendering_min_transmittance = 1e-7
tb.fov = 40.0
tb.set_ngp_camera_matrix(syn.nerf_look_at((0.4, -1.3, 0.8)))
for (w, h, spp) in [(1008,756,1),(1008,756,8),(1008,756,8),(1920,1080,8)]:
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a, _, _ = tb.render_device(w, h, spp)
    e1.record()
    torch.cuda.synchronize()
    al = a[..., 3]
    print(w, h, spp, 'ms', e0.elapsed_time(e1), 'amax', float(al.max()), 'cover', float((al > 0.5).float().mean()),
          'zero', float((al == 0).float().mean()), 'center', a[h // 2, w // 2].tolist())
