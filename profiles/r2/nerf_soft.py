"""NeRF renders of one r9 frame on the bench's soft-shell object (bench.py make_nerf_objects): reference view 1008x756
spp 8 (Shade) and mask source 1920x1080 spp 8 (Depth).  Prints CUDA-event times and the render statistics
(samples per ray, lane utilisation).   python profiles/r2/nerf_soft.py [reps] [density_gain]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, 'tests'))
import synthetic as syn  # noqa: E402
from pixtrack_b200.nerf import NerfTestbed, occupancy_bitfield  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
gain = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
dev = torch.device('cuda:0')
sc = syn.nerf_scene(11, 2, density_gain=gain)
tb = NerfTestbed(sc['grid'], sc['w_density'], sc['w_rgb'], occupancy_bitfield(sc['density_grid'], sc['max_cascade']), 2, dev)
tb.nerf.rendering_min_transmittance = 0.01
tb.set_ngp_camera_matrix(syn.nerf_look_at((0.4, -1.3, 0.8)))
out = {}
for name, (w, h, f, depth) in dict(reference=(1008, 756, 1209.6, False), depth=(1920, 1080, 2304.0, True)).items():
    tb.fov = 2 * np.degrees(np.arctan(w / (2 * f)))
    tb.render_mode = tb.render_mode.Depth if depth else tb.render_mode.Shade
    buf = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
    for _ in range(2):
        tb.render_device(w, h, 8, want_rgba=False, out_u8=buf)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        tb.render_device(w, h, 8, want_rgba=False, out_u8=buf)
    b.record()
    torch.cuda.synchronize()
    st = tb.last_stats()
    ms = a.elapsed_time(b) / reps
    out[name] = dict(ms=ms, rays_total=w * h * 8, **st, samples_per_ray=st['samples'] / max(1, st['rays']),
                     lane_utilisation=st['samples'] / max(1, 32 * st['warp_steps']),
                     gsamples_per_s=st['samples'] / ms / 1e6, covered=float((buf != 0).any(-1).float().mean()))
    tb.render_mode = tb.render_mode.Shade
print(json.dumps(out))
