"""CPU study (no GPU): how far does operand precision in the UNet move the converged LM pose at the C2 sizes?

Emulates the extractor's arithmetic inside the fp32 oracle UNet: every conv multiplies operands rounded to
  f16   : one fp16 value                      (the tcgen05 kind::f16 path)
  split : hi + lo, two fp16 values (22 bits)  (3 MMAs: hi*hi + hi*lo + lo*hi)
with fp32 accumulation, and activations are stored in the same form.  Then runs the oracle LM on both pyramids
and prints the pose differences.   python profiles/r2/precision_study.py [n_frames]
"""
import os, sys, json, time
import numpy as np, torch, torch.nn.functional as tF
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lm, unet
import os as _os, sys as _sys  # noqa: E401,E402
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))), 'tests'))  # scene generators live with the tests
import synthetic as syn  # noqa: E402

torch.set_grad_enabled(False)
STOP = dict(num_iters=150, grad_stop=1e-4, dt_stop=5e-3, dR_stop=5e-2)


def q16(x):
    return x.half().float()


def q22(x):
    hi = x.half().float()
    return hi + (x - hi).half().float()


def forward(sd, image01, q, qhead=None, qfirst=None):
    qhead = qhead or q
    qfirst = qfirst or q
    mean = image01.new_tensor(unet.IMAGENET_MEAN)[:, None, None]
    std = image01.new_tensor(unet.IMAGENET_STD)[:, None, None]
    x = qfirst((image01 - mean) / std)
    skips = []
    first = True
    for b in range(5):
        if b > 0:
            x = tF.max_pool2d(x, 2, 2)
        for name in unet._block_convs(sd, b):
            qq = qfirst if first else q
            x = qq(tF.relu(tF.conv2d(qq(x), qq(sd[name + '.weight']), sd[name + '.bias'], padding=1)))
            first = False
        skips.append(x)
    pre = [skips[-1]]
    for i, skip in enumerate(skips[:-1][::-1]):
        up = q(tF.interpolate(pre[-1], scale_factor=2, mode='bilinear', align_corners=False))
        skip = skip[:, :, :up.shape[2], :up.shape[3]]
        p = f'decoder.{i}.layers.'
        scale = sd[p + '1.weight'] / torch.sqrt(sd[p + '1.running_var'] + 1e-5)
        w = q(sd[p + '0.weight'] * scale[:, None, None, None])
        bias = sd[p + '1.bias'] - sd[p + '1.running_mean'] * scale
        pre.append(q(tF.relu(tF.conv2d(torch.cat([up, skip], 1), w, bias, padding=1))))
    pre = pre[::-1]
    feats, confs = [], []
    for idx, s in enumerate(unet.OUTPUT_SCALES):
        xin = qhead(pre[s])
        feats.append(tF.conv2d(xin, qhead(sd[f'adaptation.{idx}.0.weight']), sd[f'adaptation.{idx}.0.bias'])[0])
        unc = tF.conv2d(xin, qhead(sd[f'uncertainty.{idx}.0.weight']), sd[f'uncertainty.{idx}.0.bias'])
        confs.append(torch.sigmoid(-unc)[0])
    return feats, confs


def extract(sd, image, q, **kw):
    sr = (1.0, 1.0)
    if max(image.shape[:2]) > 1024:
        image, sr = unet.resize_max_edge(image, 1024)
    x = torch.from_numpy(np.ascontiguousarray(image.transpose(2, 0, 1)) / 255.).float()[None]
    f, c = forward(sd, x, q, **kw)
    return f, [(sr[0] / 2 ** s, sr[1] / 2 ** s) for s in unet.OUTPUT_SCALES], c


def rot_angle(Ra, Rb):
    """Geodesic angle, well conditioned near 0: atan2(|vee(M - M^T)| / 2, (tr M - 1) / 2).  (acos of the trace alone has
    a noise floor of sqrt(2 * 1e-7) = 4e-4 rad on float32 rotation matrices.)"""
    M = Ra.double() @ Rb.double().t()
    s = 0.5 * torch.stack([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]]).norm()
    return float(torch.atan2(s, (M.diagonal().sum() - 1) / 2))


def poses(sd, seq, fr, q, n_views, **kw):
    fr_f, sc_r, cf_r = extract(sd, fr['img_r'].numpy().astype(np.float32), q, **kw)
    maps_r = [torch.cat([f, c], 0) for f, c in zip(fr_f, cf_r)]
    obs, keep = lm.sample_reference(maps_r, sc_r, seq['cam_r'], fr['R_r'], fr['t_r'], seq['p3d'])
    fq, sc_q, cf_q = extract(sd, fr['img_q'].numpy().astype(np.float32), q, **kw)
    maps_q = [torch.cat([f, c], 0) for f, c in zip(fq, cf_q)]
    lam = lm.damping_lambda(torch.zeros(6))
    res = []
    for v in range(n_views):
        T0 = fr['T_init'][v]
        out = lm.refine_levels(maps_q, sc_q, seq['cam_q'].float(), T0[:9].reshape(3, 3), T0[9:], [o[keep] for o in obs],
                               seq['p3d'][keep].float(), [lam] * 3, **STOP)
        res.append((out['R'].double(), out['t'].double(), [r['n_iters'] for r in out['runs']]))
    return res, (fq, cf_q), int(keep.sum())


if __name__ == '__main__':
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    torch.set_num_threads(os.cpu_count())
    seq = syn.tracked_sequence(100, n_frames=max(nf, 1), N=5000, n_views=8)
    sd = syn.unet_weights(0)
    ident = lambda x: x
    modes = {'f16': dict(q=q16), 'split': dict(q=q22), 'f16_body+split_heads': dict(q=q16, qhead=q22)}
    for k, fr in enumerate(seq['frames'][:nf]):
        t0 = time.time()
        base, (f0, c0), nk = poses(sd, seq, fr, ident, nv)
        print(f'frame {k}: fp32 oracle {time.time() - t0:.1f}s, kept {nk}, iters {[b[2] for b in base]}', flush=True)
        gt = [(rot_angle(b[0], fr['R_q']), float((b[1] - fr['t_q']).norm())) for b in base]
        print('  fp32 vs GT   dR max %.2e  dt max %.2e' % (max(g[0] for g in gt), max(g[1] for g in gt)))
        for name, kw in modes.items():
            got, (f1, c1), nk1 = poses(sd, seq, fr, nv=None, **kw) if False else poses(sd, seq, fr, kw['q'], nv, **{a: b for a, b in kw.items() if a != 'q'})
            dR = [rot_angle(a[0], b[0]) for a, b in zip(got, base)]
            dt = [float((a[1] - b[1]).norm()) for a, b in zip(got, base)]
            rel = [float((a - b).norm() / b.norm()) for a, b in zip(f1, f0)]
            its = [[x - y for x, y in zip(a[2], b[2])] for a, b in zip(got, base)]
            print(f'  {name:22s} dR max {max(dR):.2e} med {np.median(dR):.2e}  dt max {max(dt):.2e} med {np.median(dt):.2e} '
                  f'feat rel {[f"{r:.1e}" for r in rel]} kept {nk1} iter deltas {its}', flush=True)
