"""Every tensor-core convolution of the 1024x576 (and 1008x756) extractor plans in isolation: 10 back-to-back launches
per sample, median of 15 samples.  Environment switches of ptk_conv.cu (PTK_CONV_HALO, PTK_CONV_HALO_SPLIT,
PTK_CONV_SPLIT, PTK_CONV_PAIR ...) select kernel variants for A/B runs.
    python profiles/r2/conv_layers.py [576x1024|756x1008|small]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pixtrack_b200.extractor import conv_f16, pack_conv3x3  # noqa: E402

D = 'cuda:0'
which = sys.argv[1] if len(sys.argv) > 1 else '576x1024'
H0, W0 = (576, 1024) if which != '756x1008' else (756, 1008)
ENC = ((64, 64), (128, 128), (256, 256, 256, 256), (512, 512, 512, 512), (512, 512, 512, 512))
layers = []
cin = 64
h, w = H0, W0
for b, chans in enumerate(ENC):
    if b > 0:
        h, w = h // 2, w // 2
    for i, c in enumerate(chans):
        if not (b == 0 and i == 0):
            layers.append((f'enc{b}.{i}', cin, 0, c, h, w, i == len(chans) - 1 and b < 4))
        cin = c
sizes = [(H0 >> k, W0 >> k) for k in range(5)]
prev, ph, pw = 512, sizes[4][0], sizes[4][1]
for i, (out, skip) in enumerate(zip((64, 64, 64, 32), (512, 256, 128, 64))):
    ph, pw = 2 * ph, 2 * pw
    layers.append((f'dec{i}', prev, skip, out, ph, pw, False))
    prev = out
if which == 'small':
    layers = [l for l in layers if l[4] * l[5] <= 144 * 256]
rows = []
for name, c0, c1, cout, h, w, pool in layers:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(h, w, c0, generator=g).half().to(D)
    x1 = torch.randn(h, w, c1, generator=g).half().to(D) if c1 else None
    wt = pack_conv3x3((torch.randn(cout, c0 + c1, 3, 3, generator=g) / 50).half().to(D))
    bias = torch.randn(cout, generator=g).to(D)
    for _ in range(3):
        conv_f16(x, wt, bias, relu=True, x1=x1, pool=pool)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(15):
        e0.record()
        for _ in range(10):
            conv_f16(x, wt, bias, relu=True, x1=x1, pool=pool)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e2)
    ts.sort()
    us = ts[len(ts) // 2]
    fl = 2.0 * h * w * 9 * (c0 + c1) * cout
    rows.append(dict(layer=name, shape=f'{c0}+{c1}->{cout} @{h}x{w}' + (' +pool' if pool else ''), us=round(us, 2),
                     tflops=round(fl / us / 1e6, 1), gflop=round(fl / 1e9, 2)))
    print(f'{name:8s} {rows[-1]["shape"]:32s} {us:8.1f} us {rows[-1]["tflops"]:8.1f} TF/s', flush=True)
tot_us = sum(r['us'] for r in rows)
tot_fl = sum(r['gflop'] for r in rows)
print(json.dumps(dict(env={k: v for k, v in os.environ.items() if k.startswith('PTK_')}, total_us=tot_us, total_gflop=tot_fl,
                      tflops=tot_fl / tot_us * 1e3 / 1e3, layers=rows)))
