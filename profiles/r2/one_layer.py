import os, sys, torch
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
from pixtrack_b200.extractor import conv_f16, pack_conv3x3
D='cuda:0'
g = torch.Generator().manual_seed(1)
c0, cout, h, w, pool = 64, 64, 576, 1024, True
x = torch.randn(h, w, c0, generator=g).half().to(D)
wt = pack_conv3x3((torch.randn(cout, c0, 3, 3, generator=g) / 50).half().to(D))
bias = torch.randn(cout, generator=g).to(D)
for _ in range(4):
    conv_f16(x, wt, bias, relu=True, pool=pool)
torch.cuda.synchronize()
print('ok')
