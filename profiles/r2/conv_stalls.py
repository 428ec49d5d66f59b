"""Stall attribution of the halo conv kernel per layer shape (run with PTK_CONV_DBG=1: the library then synchronises after
every halo launch and prints, for CTA 0, the cycles its TMA / MMA / epilogue warps spent in total and waiting on each
barrier kind).   PTK_CONV_DBG=1 python profiles/r2/conv_stalls.py 2> stalls.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pixtrack_b200.extractor import conv_f16, pack_conv3x3  # noqa: E402

D = 'cuda:0'
LAYERS = [('enc0.1', 64, 0, 64, 576, 1024, True), ('enc1.0', 64, 0, 128, 288, 512, False), ('enc1.1', 128, 0, 128, 288, 512, True),
          ('enc2.0', 128, 0, 256, 144, 256, False), ('enc2.1', 256, 0, 256, 144, 256, False), ('dec1', 64, 256, 64, 144, 256, False),
          ('dec2', 64, 128, 64, 288, 512, False), ('dec3', 64, 64, 32, 576, 1024, False),
          # the 1/8-scale block on conv_row2_kernel (row tiles, CTA pairs)
          ('enc3.0', 256, 0, 512, 72, 128, False), ('enc3.1', 512, 0, 512, 72, 128, False), ('enc3.3', 512, 0, 512, 72, 128, True),
          ('enc3.1r', 512, 0, 512, 94, 126, False),
          # first / second decoder convolutions (N = 64 row tiles where the dispatch picks them)
          ('dec0', 512, 512, 64, 72, 128, False), ('dec0r', 512, 512, 64, 94, 126, False), ('dec1r', 64, 256, 64, 188, 252, False)]
if len(sys.argv) > 1:
    LAYERS = [l for l in LAYERS if l[0] in sys.argv[1:]]
for name, c0, c1, cout, h, w, pool in LAYERS:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(h, w, c0, generator=g).half().to(D)
    x1 = torch.randn(h, w, c1, generator=g).half().to(D) if c1 else None
    wt = pack_conv3x3((torch.randn(cout, c0 + c1, 3, 3, generator=g) / 50).half().to(D))
    bias = torch.randn(cout, generator=g).to(D)
    for i in range(3):
        sys.stderr.write(f'## {name} run {i}\n')
        sys.stderr.flush()
        conv_f16(x, wt, bias, relu=True, x1=x1, pool=pool)
        torch.cuda.synchronize()
