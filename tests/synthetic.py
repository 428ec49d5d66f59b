"""Deterministic synthetic scenes for tests and bench.

No real assets (PixLoc checkpoint, premier_protein, YCB, NeRF snapshots) exist
in the build container, so every configuration is restated on seeded
synthetic inputs.  All randomness comes from `torch.Generator(seed)` on the
CPU, so the same seed gives the same bytes here and on the GPU box.

Nothing here is on the measured path: it only manufactures inputs.
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as tF

Tensor = torch.Tensor


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device='cpu')
    g.manual_seed(int(seed))
    return g


def _blur(x: Tensor, sigma: float) -> Tensor:
    """Separable Gaussian low-pass of a [C,H,W] map (reflect borders)."""
    if sigma <= 0:
        return x
    r = max(1, int(math.ceil(3 * sigma)))
    k = torch.exp(-0.5 * (torch.arange(-r, r + 1, dtype=x.dtype) / sigma) ** 2)
    k = k / k.sum()
    C = x.shape[0]
    x = x[None]
    x = tF.conv2d(tF.pad(x, (r, r, 0, 0), mode='replicate'), k.view(1, 1, 1, -1).expand(C, 1, 1, -1), groups=C)
    x = tF.conv2d(tF.pad(x, (0, 0, r, r), mode='replicate'), k.view(1, 1, -1, 1).expand(C, 1, -1, 1), groups=C)
    return x[0]


def smooth_feature_map(C: int, H: int, W: int, seed: int, sigma: float = 2.0,
                       normalize: bool = True) -> Tensor:
    """Low-pass Gaussian noise [C,H,W], unit L2 norm over C at every pixel, so
    that the feature-metric cost has a basin of a few pixels."""
    x = _blur(torch.randn(C, H, W, generator=_gen(seed)), sigma)
    x = x / x.std()
    return tF.normalize(x, dim=0) if normalize else x


def smooth_confidence(H: int, W: int, seed: int, sigma: float = 4.0) -> Tensor:
    x = _blur(torch.randn(1, H, W, generator=_gen(seed)), sigma)
    return torch.sigmoid(2.0 * x / x.std())


def axis_angle_to_R(w: Tensor) -> Tensor:
    th = float(w.norm())
    if th < 1e-12:
        return torch.eye(3, dtype=w.dtype)
    k = w / th
    K = torch.tensor([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=w.dtype)
    return torch.eye(3, dtype=w.dtype) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)


def pixtrack_camera(width: int = 1920, height: int = 1080, k1: float = 0.0) -> Tensor:
    """The prior r9 gets from pycolmap.infer_camera_from_image: SIMPLE_RADIAL,
    f = 1.2*max(w,h), c = (w/2, h/2), then Camera.from_colmap's -0.5 shift
    (reference pixtrack/pose_trackers/pixloc_tracker_r9.py:108-118,
    pixloc/pixloc/pixlib/geometry/wrappers.py:229-255)."""
    f = 1.2 * max(width, height)
    return torch.tensor([width, height, f, f, width / 2 - 0.5, height / 2 - 0.5, k1, 0.0])


def scale_cam(cam: Tensor, s: Sequence[float]) -> Tensor:
    s = torch.tensor([float(s[0]), float(s[1])], dtype=cam.dtype)
    return torch.cat([cam[0:2] * s, cam[2:4] * s, (cam[4:6] + 0.5) * s - 0.5, cam[6:]])


def project_pinhole_radial(cam: Tensor, R: Tensor, t: Tensor, p3d: Tensor) -> Tensor:
    """Plain projection used only to place synthetic reference descriptors."""
    pc = p3d @ R.t() + t
    xy = pc[:, :2] / pc[:, 2:3]
    if cam.numel() > 6:
        r2 = (xy ** 2).sum(-1, keepdim=True)
        xy = xy * (1 + cam[6] * r2 + cam[7] * r2 ** 2)
    return xy * cam[2:4] + cam[4:6]


def bilinear(Fm: Tensor, uv: Tensor) -> Tensor:
    C, H, W = Fm.shape
    g = uv / torch.tensor([W - 1, H - 1], dtype=uv.dtype) * 2 - 1
    return tF.grid_sample(Fm[None], g[None, :, None], mode='bilinear', align_corners=True).reshape(C, -1).t()


def object_points(N: int, seed: int, depth: float = 1.2, spread: float = 0.15) -> Tensor:
    p = torch.randn(N, 3, generator=_gen(seed)) * spread
    p[:, 2] += depth
    return p


def perturb_pose(R: Tensor, t: Tensor, seed: int, rot_deg: float, trans: float):
    g = _gen(seed)
    ax = torch.randn(3, generator=g)
    ax = ax / ax.norm() * math.radians(rot_deg)
    dt = torch.randn(3, generator=g)
    dt = dt / dt.norm() * trans
    Rd = axis_angle_to_R(ax)
    return Rd @ R, Rd @ t + dt


def level_problem(seed: int = 0, N: int = 500, C: int = 128, H: int = 144, W: int = 256,
                  level_scale: float = (1024 / 1920) / 4, B: int = 1, noise: float = 0.05,
                  rot_deg: float = 2.0, trans: float = 0.02, sigma: float = 2.0,
                  k1: float = 0.0) -> Dict[str, Tensor]:
    """One pyramid level of a PixTrack-shaped problem (SURVEY config C1b): a
    1920x1080 SIMPLE_RADIAL camera scaled to the level, smooth unit-norm query
    map, reference descriptors = query map sampled at the ground-truth
    projection + noise (re-normalised), B perturbed initial poses."""
    cam = scale_cam(pixtrack_camera(k1=k1), (level_scale, level_scale))
    cam[0], cam[1] = float(W), float(H)
    Fq = smooth_feature_map(C, H, W, seed * 7 + 1, sigma)
    Wq = smooth_confidence(H, W, seed * 7 + 2)
    p3d = object_points(N, seed * 7 + 3)
    R_gt = axis_angle_to_R(torch.randn(3, generator=_gen(seed * 7 + 4)) * 0.05)
    t_gt = torch.randn(3, generator=_gen(seed * 7 + 5)) * 0.02
    uv = project_pinhole_radial(cam, R_gt, t_gt, p3d)
    g = _gen(seed * 7 + 6)
    F_ref, W_ref, R0, t0 = [], [], [], []
    for b in range(B):
        fr = bilinear(Fq, uv) + noise * torch.randn(N, C, generator=g) / math.sqrt(C)
        F_ref.append(tF.normalize(fr, dim=1))
        W_ref.append(0.5 + 0.5 * torch.rand(N, 1, generator=g))
        Rb, tb = perturb_pose(R_gt, t_gt, seed * 1000 + b, rot_deg, trans)
        R0.append(Rb)
        t0.append(tb)
    return dict(cam=cam, F_q=Fq, W_q=Wq, p3d=p3d, R_gt=R_gt, t_gt=t_gt,
                F_ref=torch.stack(F_ref), W_ref=torch.stack(W_ref),
                R0=torch.stack(R0), t0=torch.stack(t0))


def toy_problem(seed: int = 0, n_points: int = 500) -> Dict[str, Tensor]:
    """The reference's own fixture, regenerated with the same RNG call order
    (reference pixloc/pixloc/pixlib/geometry/check_jacobians.py:32-54):
    640x480, f=(300,350), c=(320,240), radial (0.1, 0.01), 16 channels.
    Adds the confidences SURVEY config C1a asks for (drawn AFTER the fixture's
    own draws so those stay identical to the reference's)."""
    torch.random.manual_seed(seed)
    aa = torch.randn(3) / 10
    t = torch.randn(3) / 5
    w, h = 640, 480
    cam = torch.tensor([w, h, 300., 350., w / 2, h / 2, 0.1, 0.01])
    p3d = torch.randn(n_points, 3)
    p3d[:, -1] += 2
    F_ref = torch.randn(n_points, 16)
    F_q = torch.randn(16, h, w)
    W_ref = torch.rand(n_points, 1)
    W_q = torch.rand(1, h, w)
    return dict(aa=aa, t0=t, R0=axis_angle_to_R(aa), cam=cam, p3d=p3d, F_ref=F_ref, F_q=F_q,
                W_ref=W_ref, W_q=W_q)


# ---------------------------------------------------------------------------
# procedural image + random extractor weights
# ---------------------------------------------------------------------------
def textured_image(H: int, W: int, seed: int) -> Tensor:
    """HxWx3 float32 in 0..255: multi-octave smooth noise (object-like texture)."""
    g = _gen(seed)
    img = torch.zeros(3, H, W)
    for octave, amp in ((4, 1.0), (16, 0.6), (64, 0.35)):
        h, w = max(2, H // octave), max(2, W // octave)
        n = torch.randn(1, 3, h, w, generator=g)
        img += amp * tF.interpolate(n, size=(H, W), mode='bilinear', align_corners=False)[0]
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255.0).permute(1, 2, 0).contiguous()


VGG19_BLOCKS = ((64, 64), (128, 128), (256, 256, 256, 256), (512, 512, 512, 512), (512, 512, 512, 512))
DECODER = (64, 64, 64, 32)
OUTPUT_SCALES = (0, 2, 4)
OUTPUT_DIM = (32, 128, 128)


def unet_weights(seed: int = 0) -> Dict[str, Tensor]:
    """Random (He-scaled) weights for the PixLoc UNet, keyed like the
    checkpoint's `extractor.*` state dict (reference
    pixloc/pixloc/pixlib/models/unet.py:68-156; config
    pixloc/pixloc/pixlib/configs/train_pixloc_megadepth.yaml: vgg19 encoder,
    decoder [64,64,64,32], output_dim [32,128,128], output_scales [0,2,4],
    uncertainty heads).  Encoder keys follow torchvision's vgg19.features
    indexing regrouped into 5 blocks (block b, position i inside the block)."""
    g = _gen(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin, k, bias=True, gain=2.0):
        fan = cin * k * k
        sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(gain / fan)
        if bias:
            sd[name + '.bias'] = 0.1 * torch.randn(cout, generator=g)

    cin = 3
    for b, chans in enumerate(VGG19_BLOCKS):
        pos = 0 if b == 0 else 1            # blocks 1.. start with the max-pool
        for c in chans:
            conv(f'encoder.{b}.{pos}', c, cin, 3)
            cin = c
            pos += 2                        # conv, relu
    skip = [c[-1] for c in VGG19_BLOCKS]
    prev = skip[-1]
    for i, (out, sk) in enumerate(zip(DECODER, skip[:-1][::-1])):
        conv(f'decoder.{i}.layers.0', out, prev + sk, 3, bias=False)
        sd[f'decoder.{i}.layers.1.weight'] = 1.0 + 0.1 * torch.randn(out, generator=g)
        sd[f'decoder.{i}.layers.1.bias'] = 0.1 * torch.randn(out, generator=g)
        sd[f'decoder.{i}.layers.1.running_mean'] = 0.1 * torch.randn(out, generator=g)
        sd[f'decoder.{i}.layers.1.running_var'] = 0.5 + torch.rand(out, generator=g)
        sd[f'decoder.{i}.layers.1.num_batches_tracked'] = torch.tensor(1)
        prev = out
    for idx, s in enumerate(OUTPUT_SCALES):
        cin = skip[s] if s == len(VGG19_BLOCKS) - 1 else DECODER[-1 - s]
        conv(f'adaptation.{idx}.0', OUTPUT_DIM[idx], cin, 1, gain=1.0)
        conv(f'uncertainty.{idx}.0', 1, cin, 1, gain=1.0)
    return sd


# ---------------------------------------------------------------------------
# full frames (SURVEY configs C2 / C4): 3-level pyramid, B views, N points
# ---------------------------------------------------------------------------
def frame_problem(seed: int, N: int = 5000, B: int = 8, size: Tuple[int, int] = (576, 1024),
                  dims: Sequence[int] = (32, 128, 128), strides: Sequence[int] = (1, 4, 16),
                  sigmas: Sequence[float] = (4.0, 2.0, 1.0), noise: float = 0.05, rot_deg: float = 1.0,
                  trans: float = 0.005) -> Dict:
    """One tracked frame: a 1920x1080 query resized to 1024x576 gives level
    maps 32x576x1024 / 128x144x256 / 128x36x64.  Each of the B
    reference views contributes N observations whose descriptors are the query
    maps sampled at the ground-truth projection + noise; each view starts from
    its own perturbed pose (frame-to-frame motion of SURVEY C2).
    Returns per-level lists (fine -> coarse) of CHW query maps `F_q`, `W_q`
    [1,H,W], cameras, `F_ref` [B,N,C], `W_ref` [B,N], plus p3d, T_init [B,12],
    R_gt, t_gt."""
    Hf, Wf = size
    cam0 = scale_cam(pixtrack_camera(), (Wf / 1920.0, Wf / 1920.0))
    p3d = object_points(N, seed * 13 + 1)
    R_gt = axis_angle_to_R(torch.randn(3, generator=_gen(seed * 13 + 2)) * 0.05)
    t_gt = torch.randn(3, generator=_gen(seed * 13 + 3)) * 0.02
    g = _gen(seed * 13 + 4)
    out = dict(F_q=[], W_q=[], cam=[], F_ref=[], W_ref=[], p3d=p3d, R_gt=R_gt, t_gt=t_gt)
    for lv, (C, s, sg) in enumerate(zip(dims, strides, sigmas)):
        H, W = Hf // s, Wf // s
        cam = scale_cam(cam0, (1.0 / s, 1.0 / s))
        Fq = smooth_feature_map(C, H, W, seed * 13 + 5 + lv, sg)
        Wq = smooth_confidence(H, W, seed * 13 + 8 + lv)
        uv = project_pinhole_radial(cam, R_gt, t_gt, p3d)
        base = bilinear(Fq, uv)
        fr = base[None] + noise * torch.randn(B, N, C, generator=g) / math.sqrt(C)
        out['F_q'].append(Fq)
        out['W_q'].append(Wq)
        out['cam'].append(cam)
        out['F_ref'].append(tF.normalize(fr, dim=2))
        out['W_ref'].append(0.5 + 0.5 * torch.rand(B, N, generator=g))
    T0 = []
    for b in range(B):
        Rb, tb = perturb_pose(R_gt, t_gt, seed * 1000 + b, rot_deg, trans)
        T0.append(torch.cat([Rb.reshape(-1), tb]))
    out['T_init'] = torch.stack(T0)
    return out


# ---------------------------------------------------------------------------
# geometrically consistent image pairs: a textured plane seen by two cameras
# ---------------------------------------------------------------------------
def plane_texture(seed: int, size: int = 1536) -> Tensor:
    """[3, size, size] in 0..255: multi-octave smooth noise, the 'object' texture on the plane z = 0."""
    g = _gen(seed)
    tex = torch.zeros(3, size, size)
    for cells, amp in ((6, 1.0), (24, 0.7), (96, 0.45), (384, 0.25)):
        n = torch.randn(1, 3, cells, cells, generator=g)
        tex += amp * tF.interpolate(n, size=(size, size), mode='bicubic', align_corners=False)[0]
    tex = (tex - tex.mean()) / tex.std()
    return (torch.sigmoid(1.2 * tex) * 255.0).contiguous()


def render_plane(tex: Tensor, cam: Tensor, R: Tensor, t: Tensor, half_extent: float = 0.6) -> Tensor:
    """Pinhole view (cam = [w,h,fx,fy,cx,cy,...], distortion ignored; world -> camera p_c = R p + t) of the
    plane z = 0 whose square [-half_extent, half_extent]^2 carries `tex`.  Returns uint8 [h, w, 3] (outside
    the square: mid grey), i.e. what the NeRF render / the camera frame would hand to the extractor."""
    w, h = int(cam[0]), int(cam[1])
    R, t = R.double(), t.double()
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing='ij')
    d = torch.stack([(xs - float(cam[4])) / float(cam[2]), (ys - float(cam[5])) / float(cam[3]), torch.ones_like(xs)], -1)
    dw = d @ R                      # R^T d for every pixel (row vectors)
    c = -(R.t() @ t)                # camera centre in world
    lam = -c[2] / dw[..., 2]
    P = c + lam[..., None] * dw     # intersection with z = 0
    g = (P[..., :2] / half_extent).float()
    img = tF.grid_sample(tex[None], g[None], mode='bilinear', padding_mode='zeros', align_corners=False)[0]
    inside = ((g.abs() <= 1).all(-1) & (lam > 0))[None]
    img = torch.where(inside, img, torch.full_like(img, 127.0))
    return img.permute(1, 2, 0).round().clamp(0, 255).to(torch.uint8).contiguous()


def tracked_sequence(seed: int, n_frames: int = 4, N: int = 5000, n_views: int = 8, query_wh=(1920, 1080),
                     ref_wh=(1008, 756), rot_deg: float = 1.0, trans: float = 0.005, cam_q: Optional[Tensor] = None,
                     cam_r: Optional[Tensor] = None) -> Dict:
    """A short tracked sequence with REAL pixels (BASELINE config 2 restated): one textured
    plane carrying N model points, and per frame a query image (1920x1080 camera frame) plus the
    reference-view render r9 makes at the previous pose estimate with the SfM camera x 0.5 (1008x756;
    reference pixtrack/pose_trackers/pixloc_tracker_r9.py:145-160), and n_views perturbed initial poses.
    Everything the per-frame hot path consumes: images (uint8 HWC), cameras at image resolution,
    world->camera poses (float64), points (float64)."""
    tex = plane_texture(seed * 3 + 1)
    g = _gen(seed * 3 + 2)
    p = torch.zeros(N, 3, dtype=torch.float64)
    p[:, :2] = (torch.rand(N, 2, generator=g, dtype=torch.float64) - 0.5) * 0.5       # central 0.5 m square
    cam_q = pixtrack_camera(*query_wh) if cam_q is None else cam_q.clone()      # cam_q / cam_r: explicit intrinsics
    cam_r = pixtrack_camera(*ref_wh) if cam_r is None else cam_r.clone()        # (e.g. the YCB-Video camera, config C3)
    cam_q[6:] = 0
    cam_r[6:] = 0
    base = torch.rand(6, generator=g, dtype=torch.float64) - 0.5
    frames = []
    for f in range(n_frames):
        d = (torch.rand(12, generator=g, dtype=torch.float64) - 0.5)
        aq = torch.stack([0.25 * base[0] + 0.03 * d[0], 0.25 * base[1] + 0.03 * d[1], 0.3 * base[2] + 0.03 * d[2]])
        R_q = axis_angle_to_R(aq)
        t_q = torch.stack([0.05 * base[3] + 0.01 * d[3], 0.05 * base[4] + 0.01 * d[4], 1.25 + 0.1 * base[5] + 0.01 * d[5]])
        R_r = axis_angle_to_R(aq + 0.02 * d[6:9])                # render pose = a nearby earlier estimate
        t_r = t_q + 0.01 * d[9:12]
        T0 = []
        for b in range(n_views):
            Rb, tb = perturb_pose(R_q.float(), t_q.float(), seed * 1000 + f * 37 + b, rot_deg, trans)
            T0.append(torch.cat([Rb.reshape(-1), tb]))
        frames.append(dict(img_q=render_plane(tex, cam_q, R_q, t_q), img_r=render_plane(tex, cam_r, R_r, t_r),
                           R_q=R_q, t_q=t_q, R_r=R_r, t_r=t_r, T_init=torch.stack(T0)))
    return dict(frames=frames, cam_q=cam_q.double(), cam_r=cam_r.double(), p3d=p)


# ---------------------------------------------------------------------------
# NeRF scenes (SURVEY config C5): random hash grid / MLPs, procedural occupancy
# ---------------------------------------------------------------------------
NERF_GRID = 128
NERF_CASCADES = 8


def nerf_grid_size(aabb_scale: int) -> int:
    """Number of hash-grid entries (x2 features) for the base.json encoding at this aabb_scale:
    16 levels, T = 2^19, base resolution 16, finest resolution 2048 * aabb_scale
    (instant-ngp/src/testbed.cu:2233-2244, tiny-cuda-nn encodings/grid.h:898-930)."""
    import numpy as np
    pls = np.float32(math.exp(math.log(2048.0 * aabb_scale / 16) / 15))
    log2 = np.float32(np.log2(pls))
    total = 0
    for lv in range(16):
        scale = np.float32(np.exp2(np.float32(lv) * log2) * np.float32(16) - np.float32(1))
        res = int(np.ceil(scale)) + 1
        total += min((min(res ** 3, 2 ** 31 - 1) + 7) // 8 * 8, 1 << 19)
    return total


def _morton_cell_centers(level: int):
    import numpy as np

    def compact(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xc30c30c3
        x = (x | (x >> 4)) & 0x0f00f00f
        x = (x | (x >> 8)) & 0xff0000ff
        x = (x | (x >> 16)) & 0x0000ffff
        return x
    i = np.arange(NERF_GRID ** 3, dtype=np.int64)
    xyz = np.stack([compact(i), compact(i >> 1), compact(i >> 2)], 1).astype(np.float64)
    return ((xyz + 0.5) / NERF_GRID - 0.5) * (2.0 ** level) + 0.5


def nerf_level_offsets(aabb_scale: int):
    """Entry offset of each of the 16 hash-grid levels (and the total as a 17th element); same layout as
    nerf_grid_size."""
    import numpy as np
    pls = np.float32(math.exp(math.log(2048.0 * aabb_scale / 16) / 15))
    log2 = np.float32(np.log2(pls))
    offs = [0]
    for lv in range(16):
        scale = np.float32(np.exp2(np.float32(lv) * log2) * np.float32(16) - np.float32(1))
        res = int(np.ceil(scale)) + 1
        offs.append(offs[-1] + min((min(res ** 3, 2 ** 31 - 1) + 7) // 8 * 8, 1 << 19))
    return offs


def nerf_scene(seed: int = 0, aabb_scale: int = 1, radius: float = 0.28, density_gain: float = 8.0,
               zero_network: bool = False, texture_levels: int = 16) -> Dict[str, object]:
    """A random-weight instant-ngp model with a procedurally filled occupancy grid (a ball of
    `radius` around the cube centre, in every cascade).  Returns numpy arrays shaped like an
    unpacked snapshot: grid fp16 [n,2], w_density ([64,32],[16,64]), w_rgb ([64,32],[64,64],[16,64]),
    density_grid float32 [(max_cascade+1) * 128^3] (Morton order), aabb_scale.  texture_levels < 16 zeroes
    the finer hash levels."""
    import numpy as np
    g = _gen(seed)
    n = nerf_grid_size(aabb_scale)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).numpy().astype(np.float16)
    if zero_network:
        grid = np.zeros((n, 2), np.float16)
    else:
        grid = ((torch.rand(n, 2, generator=g) * 2 - 1) * 0.6).numpy().astype(np.float16)
        if texture_levels < 16:      # only the coarse (dense) levels carry signal: a smooth, trackable texture
            grid[nerf_level_offsets(aabb_scale)[texture_levels]:] = 0
    w_density = (rnd(64, 32, scale=math.sqrt(2.0 / 32)), rnd(16, 64, scale=density_gain * math.sqrt(1.0 / 64)))
    w_density[1][0] = np.abs(w_density[1][0])      # density row: positive weights on ReLU outputs -> dense medium
    w_rgb = (rnd(64, 32, scale=math.sqrt(2.0 / 32)), rnd(64, 64, scale=math.sqrt(2.0 / 64)),
             rnd(16, 64, scale=2.0 * math.sqrt(1.0 / 64)))
    max_cascade = 0
    while (1 << max_cascade) < aabb_scale:
        max_cascade += 1
    dens = []
    for lv in range(max_cascade + 1):
        c = _morton_cell_centers(lv)
        inside = ((c - 0.5) ** 2).sum(1) < (radius + 0.87 * (2.0 ** lv) / NERF_GRID) ** 2
        dens.append(np.where(inside, 1.0, -1.0).astype(np.float32))
    return dict(aabb_scale=aabb_scale, grid=grid, w_density=w_density, w_rgb=w_rgb,
                density_grid=np.concatenate(dens), max_cascade=max_cascade)


def nerf_textured_scene(seed: int = 0, aabb_scale: int = 1, radius: float = 0.2, texture_levels: int = 5,
                        contrast: float = 4.0) -> Dict[str, object]:
    """Like nerf_scene, but with CONSTRUCTED networks so that the ball is opaque and carries a smooth,
    zero-mean, view-independent, high-contrast colour pattern (a trackable object): the density MLP passes
    the signed hash features through (relu(x) - relu(-x)) and sums their magnitudes into the density, the
    colour MLP ignores the view direction and mixes the features of the `texture_levels` coarsest levels
    with random +-contrast weights in front of the logistic.  Same array layout as nerf_scene."""
    import numpy as np
    sc = nerf_scene(seed, aabb_scale, radius=radius, texture_levels=texture_levels)
    g = _gen(seed + 7919)
    wd0 = np.zeros((64, 32), np.float16)
    for i in range(32):
        wd0[2 * i, i], wd0[2 * i + 1, i] = 1.0, -1.0
    wd1 = np.zeros((16, 64), np.float16)
    wd1[0, :] = 6.0                                         # density raw = 6 * sum |feature|: opaque within a step or two
    for k in range(1, 16):
        wd1[k, 2 * (k - 1)], wd1[k, 2 * (k - 1) + 1] = 1.0, -1.0
    wc0 = np.zeros((64, 32), np.float16)
    for k in range(1, 16):
        wc0[2 * k, k], wc0[2 * k + 1, k] = 1.0, -1.0
    wc1 = np.eye(64, dtype=np.float16)
    wc2 = np.zeros((16, 64), np.float16)
    n_feat = min(15, 2 * texture_levels)
    a = (torch.randint(0, 2, (3, n_feat), generator=g).float() * 2 - 1).numpy() * contrast
    for c in range(3):
        for k in range(1, n_feat + 1):
            wc2[c, 2 * k], wc2[c, 2 * k + 1] = a[c, k - 1], -a[c, k - 1]
    sc['w_density'], sc['w_rgb'] = (wd0, wd1), (wc0, wc1, wc2)
    return sc


def nerf_look_at(eye, target=(0.5, 0.5, 0.5), up=(0.0, 0.0, 1.0)):
    """3x4 camera-to-world matrix in the NGP convention used by the renderer (columns: right, down,
    forward, origin), looking from `eye` at `target`."""
    import numpy as np
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    return np.stack([r, d, f, eye], 1).astype(np.float32)
