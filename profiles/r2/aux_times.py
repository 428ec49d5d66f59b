"""Per-launch times of the non-tensor-core kernels of the 1024x576 plan (prep, conv1, upsample x4, heads x3): best of N
event-bracketed plan runs.   python profiles/r2/aux_times.py [runs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, 'tests'))
import synthetic as syn  # noqa: E402
from pixtrack_b200.extractor import B200FeatureExtractor  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
ext = B200FeatureExtractor(syn.unet_weights(0), dev)
img = syn.textured_image(1080, 1920, seed=6).to(torch.uint8).to(dev)
for _ in range(3):
    ext.extract_device(img, normalize=True)
torch.cuda.synchronize()
best = None
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    rows = [list(r) for r in ext.profile(img)]
    best = rows if best is None else [[b[0], min(b[1], r[1])] + list(b[2:]) for b, r in zip(best, rows)]
aux = [(r[0], round(r[1] * 1e3, 1)) for r in best if r[0] != 'conv_tc']
print(aux, 'sum_us', round(sum(a[1] for a in aux), 1), 'conv_us', round(sum(r[1] for r in best if r[0] == 'conv_tc') * 1e3, 1))
