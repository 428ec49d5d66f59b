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
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(3):
    a, _, _ = tb.render_device(1008, 756, spp)
torch.cuda.synchronize()
print('ok', float(a[..., 3].mean()))
