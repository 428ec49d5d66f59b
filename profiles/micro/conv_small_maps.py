'''Small-map / few-channel conv layers in isolation: 10 back-to-back launches per sample (median of 20).
PTK_CONV_SHAPES / PTK_CONV_SPLIT = 0 switch the tile-shape choice / the two-CTA K split off.'''
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from pixtrack_b200.extractor import conv_f16, pack_conv3x3
D='cuda:0'
def run(cin, cout, H, W, x1c=0):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(H, W, cin, generator=g).half().to(D)
    x1 = torch.randn(H, W, x1c, generator=g).half().to(D) if x1c else None
    w = (torch.randn(cout, cin + x1c, 3, 3, generator=g) / 50).half().to(D)
    b = torch.randn(cout, generator=g).to(D)
    pw = pack_conv3x3(w)
    for _ in range(5): y = conv_f16(x, pw, b, relu=True, x1=x1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(20):
        e0.record()
        for _ in range(10): y = conv_f16(x, pw, b, relu=True, x1=x1)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e2)
    ts.sort()
    fl = 2 * H * W * 9 * (cin + x1c) * cout
    print(f'{os.environ.get("PTK_CONV_SHAPES","1")}/{os.environ.get("PTK_CONV_SPLIT","1")} {cin}+{x1c}->{cout} {H}x{W}: {ts[len(ts)//2]:.1f} us  {fl / ts[len(ts)//2] / 1e6:.0f} TF/s', flush=True)
run(512, 512, 36, 64)
run(512, 64, 72, 128, 512)
run(512, 512, 47, 63)
run(512, 64, 94, 126, 512)
run(64, 64, 144, 256, 256)
