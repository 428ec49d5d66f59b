"""Per-launch device times of the native UNet plan on a 1080p frame (resized to 576x1024 on the device).
Prints one line per launch and the totals; used for the numbers in profiles/README.md and as the short
driver for ncu captures of conv_tc_kernel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import os as _os, sys as _sys  # noqa: E401,E402
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'tests'))  # scene generators live with the tests
import synthetic as syn  # noqa: E402
from pixtrack_b200.extractor import B200FeatureExtractor  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ext = B200FeatureExtractor(syn.unet_weights(0), dev)
img = (syn.textured_image(1080, 1920, seed=6)).to(torch.uint8).to(dev)
for _ in range(2):
    ext.extract_device(img, normalize=True)
torch.cuda.synchronize()
best = None
for _ in range(reps):
    rows = ext.profile(img)
    if best is None:
        best = [list(r) for r in rows]
    else:
        for b, r in zip(best, rows):
            b[1] = min(b[1], r[1])
tot_ms = sum(r[1] for r in best)
tot_fl = sum(r[2] for r in best)
tc_ms = sum(r[1] for r in best if r[0] == 'conv_tc')
tc_fl = sum(r[2] for r in best if r[0] == 'conv_tc')
for i, (k, ms, fl) in enumerate(best):
    print(f'{i:2d} {k:13s} {ms * 1e3:9.1f} us  {fl / 1e9:8.2f} GFLOP  {fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:7.1f} TF/s')
# whole plan back to back (no events in between)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ext.extract_device(img, normalize=True)
b.record()
torch.cuda.synchronize()
e2e = a.elapsed_time(b) / 10
print(json.dumps(dict(sum_of_launches_ms=tot_ms, back_to_back_ms=e2e, gflop=tot_fl / 1e9,
                      tflops_back_to_back=tot_fl / (e2e * 1e-3) / 1e12, conv_tc_ms=tc_ms,
                      conv_tc_tflops=tc_fl / (tc_ms * 1e-3) / 1e12)))
