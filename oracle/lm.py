"""CPU oracle for the feature-metric Levenberg-Marquardt pose refinement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PyTorch fp32, using the
same library primitives the reference uses (`grid_sample`, `einsum`,
`linalg.cholesky`) so that (i) results match the reference to rounding and
(ii) its wall-clock is a fair stand-in for the reference's own path.  Every
function follows the device of its inputs: on CPU tensors it is the checker
and the CPU baseline; on CUDA tensors it is what the reference executes with
`device='cuda'` (pixtrack/localization/pixloc_pose_refiners.py:35-39): library
kernels per op, the 6x6 factorisation on the host with a device round trip
per iteration -- bench.py's `library_baseline` leg times that.

Conventions (all tensors fp32 unless noted):
  pose   : R [3,3], t [3]          world -> camera, p_c = R p + t
  cam    : [w, h, fx, fy, cx, cy, (k1, k2, (p1, p2))]   6 / 8 / 10 floats
  F_q    : [C, H, W] dense query map,  F_ref : [N, C]
  delta  : [dt(3), dw(3)]          (translation first, then axis-angle)
Paths cited are relative to /root/reference/pixloc/pixloc/pixlib/.
"""
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as tF

Tensor = torch.Tensor
Z_EPS = 1e-3  # geometry/wrappers.py:225  (Camera.eps)


# ----------------------------------------------------------------------------
# SE(3) helpers                                  geometry/optimization.py:50-76
# ----------------------------------------------------------------------------
def hat(v: Tensor) -> Tensor:
    """3-vector -> 3x3 cross-product matrix (optimization.py:50-59)."""
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    o = torch.zeros_like(x)
    rows = [o, -z, y, z, o, -x, -y, x, o]
    return torch.stack(rows, -1).reshape(v.shape[:-1] + (3, 3))


def rodrigues(w: Tensor, eps: float = 1e-7) -> Tensor:
    """Axis-angle -> rotation, as optimization.py:62-76: the skew matrix is
    built from w/theta, R = I + sin(theta) W + (1-cos(theta)) W^2, and below
    `eps` the first-order form R = I + hat(w) is used."""
    theta = w.norm(dim=-1, keepdim=True)
    tiny = theta < eps
    unit = w / torch.where(tiny, torch.ones_like(theta), theta)
    W = hat(unit)
    th = theta[..., None]
    full = torch.sin(th) * W + (1.0 - torch.cos(th)) * (W @ W)
    return torch.eye(3, dtype=w.dtype, device=w.device) + torch.where(tiny[..., None], W, full)


def compose(Ra: Tensor, ta: Tensor, Rb: Tensor, tb: Tensor):
    """(Ra,ta) o (Rb,tb): geometry/wrappers.py:171-175."""
    return Ra @ Rb, ta + (Ra @ tb[..., None])[..., 0]


def pose_magnitude(R: Tensor, t: Tensor):
    """Rotation angle in DEGREES and translation norm, wrappers.py:208-218."""
    c = ((torch.diagonal(R, dim1=-1, dim2=-2).sum(-1) - 1.0) / 2.0).clamp(-1, 1)
    return torch.acos(c).abs() / math.pi * 180.0, t.norm(dim=-1)


# ----------------------------------------------------------------------------
# camera                 geometry/wrappers.py:282-365, geometry/utils.py:36-95
# ----------------------------------------------------------------------------
def scale_camera(cam: Tensor, s: Sequence[float]) -> Tensor:
    """Camera.scale (wrappers.py:276-286): size*s, f*s, (c+0.5)*s-0.5."""
    s = cam.new_tensor([float(s[0]), float(s[1])]) if not isinstance(s, (int, float)) \
        else cam.new_tensor([float(s), float(s)])
    return torch.cat([cam[0:2] * s, cam[2:4] * s, (cam[4:6] + 0.5) * s - 0.5, cam[6:]])


def distort(xy: Tensor, dist: Tensor) -> Tuple[Tensor, Tensor]:
    """Radial (+tangential) model and its validity limit, utils.py:36-69."""
    ok = torch.ones(xy.shape[:-1], dtype=torch.bool, device=xy.device)
    out = xy
    if dist.numel() > 0:
        k1, k2 = dist[0], dist[1]
        r2 = (xy ** 2).sum(-1, keepdim=True)
        out = out + xy * (k1 * r2 + k2 * r2 ** 2)
        disc = 9 * k1 ** 2 - 20 * k2
        has_limit = ((k2 > 0) & (disc > 0)) | ((k2 <= 0) & (k1 > 0))
        limit = torch.abs(torch.where(k2 > 0, (torch.sqrt(disc) - 3 * k1) / (10 * k2),
                                      1 / (3 * k1)))
        ok = ok & (~has_limit | (r2[..., 0] < limit))
        if dist.numel() > 2:
            p12 = dist[2:4]
            p21 = p12.flip(-1)
            uv = xy.prod(-1, keepdim=True)
            out = out + 2 * p12 * uv + p21 * (r2 + 2 * xy ** 2)
    return out, ok


def distort_jacobian(xy: Tensor, dist: Tensor) -> Tensor:
    """2x2 Jacobian of `distort`, utils.py:72-95."""
    diag = torch.ones_like(xy)
    cross = torch.zeros_like(xy)
    if dist.numel() > 0:
        k1, k2 = dist[0], dist[1]
        r2 = (xy ** 2).sum(-1, keepdim=True)
        uv = xy.prod(-1, keepdim=True)
        d_rad = 2 * k1 + 4 * k2 * r2
        diag = diag + (k1 * r2 + k2 * r2 ** 2) + xy ** 2 * d_rad
        cross = cross + uv * d_rad
        if dist.numel() > 2:
            p12 = dist[2:4]
            p21 = p12.flip(-1)
            diag = diag + 2 * p12 * xy.flip(-1) + 6 * p21 * xy
            cross = cross + 2 * p12 * xy + 2 * p21 * xy.flip(-1)
    return torch.diag_embed(diag) + torch.diag_embed(cross).flip(-1)


def world_to_image(cam: Tensor, p_cam: Tensor) -> Tuple[Tensor, Tensor]:
    """Camera.world2image (wrappers.py:349-355): project (z clamped to 1e-3,
    :301-307... :308-314), distort, x*f+c, and the in-image test on the FLOAT
    camera size 0 <= p <= size-1 (:299-306)."""
    z = p_cam[..., 2]
    front = z > Z_EPS
    xy = p_cam[..., :2] / z.clamp(min=Z_EPS)[..., None]
    xy, ok = distort(xy, cam[6:])
    uv = xy * cam[2:4] + cam[4:6]
    inside = ((uv >= 0) & (uv <= cam[0:2] - 1)).all(-1)
    return uv, front & ok & inside


def world_to_image_jacobian(cam: Tensor, p_cam: Tensor) -> Tensor:
    """J_world2image (wrappers.py:357-362) = diag(f) . J_dist . J_proj, N x 2 x 3."""
    x, y = p_cam[..., 0], p_cam[..., 1]
    z = p_cam[..., 2].clamp(min=Z_EPS)
    o = torch.zeros_like(z)
    J_proj = torch.stack([1 / z, o, -x / z ** 2, o, 1 / z, -y / z ** 2], -1)
    J_proj = J_proj.reshape(p_cam.shape[:-1] + (2, 3))          # wrappers.py:316-326
    xy = p_cam[..., :2] / z[..., None]
    return torch.diag_embed(cam[2:4]) @ distort_jacobian(xy, cam[6:]) @ J_proj


# ----------------------------------------------------------------------------
# dense-map sampling                           geometry/interpolation.py:57-141
# ----------------------------------------------------------------------------
def sample_map(F: Tensor, uv: Tensor, pad: int = 1, grads: bool = False):
    """Bilinear sample of F [C,H,W] at pixel coords uv [N,2] (x, y).

    interpolation.py:57-89: coords are normalised with (w-1, h-1), clamped to
    [-2, 2] and sampled with grid_sample(bilinear, align_corners=True, zero
    padding); the gradient is the central difference of two more bilinear
    samples one pixel away in x and in y, halved.  Mask (:92-95, :116): the
    point lies in [pad, W-1-pad] x [pad, H-1-pad] of the TENSOR grid."""
    C, H, W = F.shape
    span = uv.new_tensor([W - 1, H - 1])
    mask = ((uv >= pad) & (uv <= uv.new_tensor([W - pad - 1, H - pad - 1]))).all(-1)
    g = ((uv / span) * 2 - 1).clamp(-2, 2)

    def gs(pts):
        out = tF.grid_sample(F[None], pts[None, :, None], mode='bilinear', align_corners=True)
        return out.reshape(C, -1).t()

    val = gs(g)
    if not grads:
        return val, mask, None
    step = torch.eye(2, dtype=uv.dtype, device=uv.device) / span * 2
    fx0, fx1 = gs(g - step[0]), gs(g + step[0])
    fy0, fy1 = gs(g - step[1]), gs(g + step[1])
    dF = torch.stack([(fx1 - fx0) / 2, (fy1 - fy0) / 2], -1)        # N x C x 2
    return val, mask, dF


# ----------------------------------------------------------------------------
# robust loss                                     geometry/losses.py:8-19,38-82
# ----------------------------------------------------------------------------
def barron0_scaled(sq: Tensor, scale: float = 0.1) -> Tuple[Tensor, Tensor]:
    """scaled_barron(alpha=0, c=scale) applied to squared residual norms:
    loss = c^2 * 2 log1p(x / 2c^2),  weight = 2 / (x/c^2 + 2)."""
    a2 = scale ** 2
    x = sq / a2
    loss = 2 * torch.log1p(torch.clamp(0.5 * x, max=33e37))
    return loss * a2, 2 / (x + 2)


# ----------------------------------------------------------------------------
# one residual / Jacobian evaluation                  geometry/costs.py:15-67
# ----------------------------------------------------------------------------
def residual_jacobian(R: Tensor, t: Tensor, cam: Tensor, p3d: Tensor, F_ref: Tensor,
                      F_q: Tensor, W_ref: Optional[Tensor], W_q: Optional[Tensor],
                      pad: int = 1):
    p_cam = p3d @ R.t() + t                                    # wrappers.py:177-185
    uv, visible = world_to_image(cam, p_cam)
    Fp, inb, dF = sample_map(F_q, uv, pad, grads=True)
    valid = inb & visible
    w_unc = None
    if W_ref is not None:
        cq, _, _ = sample_map(W_q, uv, pad)
        w_unc = (W_ref * cq)[..., 0].masked_fill(~valid, 0.0)       # costs.py:27-32
    res = Fp - F_ref                                                # costs.py:41
    # d p_cam / d delta = [ I | -hat(p_cam) ]   wrappers.py:195-203
    J_pose = torch.cat([torch.eye(3, dtype=p_cam.dtype, device=p_cam.device).expand(p_cam.shape[0], 3, 3), -hat(p_cam)], -1)
    J_uv = world_to_image_jacobian(cam, p_cam) @ J_pose             # N x 2 x 6
    J = dF @ J_uv                                                   # N x C x 6
    return res, valid, w_unc, J, uv


def normal_equations(J: Tensor, res: Tensor, w: Tensor):
    """models/base_optimizer.py:83-92."""
    g = (w[:, None] * torch.einsum('ndi,nd->ni', J, res)).sum(0)
    H = (w[:, None, None] * torch.einsum('ndk,ndl->nkl', J, J)).sum(0)
    return g, H


def damped_solve(g: Tensor, H: Tensor, lam: Tensor, ok: bool, eps: float = 1e-6):
    """geometry/optimization.py:13-47: H += diag(max(diag(H)*lam, eps)); failed
    problems get (I, 0); Cholesky solve on the CPU; delta = -H^-1 g.
    A non-positive-definite H raises in modern torch (the reference's 'singular
    U' fallback string no longer matches and `torch.solve` is gone), which
    `refine_query_pose` turns into success=False; the oracle returns None."""
    H = H + torch.diag_embed((torch.diagonal(H) * lam).clamp(min=eps))
    if not ok:
        H, g = torch.eye(6).to(H), torch.zeros(6).to(g)
    H_, g_ = H.cpu(), g.cpu()             # optimization.py:33: the factorisation runs on the host, also for CUDA tensors
    try:
        L = torch.linalg.cholesky(H_)
    except RuntimeError:
        return None
    return (-torch.cholesky_solve(g_[:, None], L)[:, 0]).to(H.device)


def damping_lambda(const: Tensor, log_range=(-6.0, 5.0)) -> Tensor:
    """DampingNet.forward, models/learned_optimizer.py:25-27."""
    lo, hi = log_range
    return 10.0 ** (lo + torch.sigmoid(const) * (hi - lo))


# ----------------------------------------------------------------------------
# the LM loop          models/learned_optimizer.py:48-95 (as PixTrackOptimizer,
#                      /root/reference/pixtrack/optimizers/pixtrack_optimizer.py:6-18)
# ----------------------------------------------------------------------------
def lm_run(p3d: Tensor, F_ref: Tensor, F_q: Tensor, R0: Tensor, t0: Tensor, cam: Tensor,
           W_ref: Optional[Tensor] = None, W_q: Optional[Tensor] = None,
           mask: Optional[Tensor] = None, lam: Optional[Tensor] = None,
           num_iters: int = 150, pad: int = 1, loss_scale: float = 0.1,
           grad_stop: float = 1e-4, dt_stop: float = 5e-3, dR_stop: float = 5e-2,
           on_iter: Optional[Callable] = None) -> Dict:
    """Returns dict(R, t, failed, n_iters, raised, log=[per-iteration dicts]).

    Per iteration (learned_optimizer.py:62-92): residual+Jacobian; failed |=
    n_valid < 10; robust loss/weights; g, H; damped solve masked by ~failed;
    T <- exp(delta) o T; log; stop test EVERY iteration (pixtrack_optimizer.py:8)
    with the UNMASKED gradient norm."""
    if lam is None:
        lam = damping_lambda(torch.zeros(6))
    R, t = R0.clone(), t0.clone()
    failed, raised = False, False
    log: List[Dict] = []
    it = 0
    for it in range(num_iters):
        res, valid, w_unc, J, _ = residual_jacobian(R, t, cam, p3d, F_ref, F_q, W_ref, W_q, pad)
        if mask is not None:
            valid = valid & mask
        n_valid = int(valid.sum())
        failed = failed or n_valid < 10
        sq = (res ** 2).sum(-1)
        cost, w_loss = barron0_scaled(sq, loss_scale)
        w = w_loss * valid.float()
        if w_unc is not None:
            w = w * w_unc
        g, H = normal_equations(J, res, w)
        delta = damped_solve(g, H, lam, not failed)
        if delta is None:
            raised = True
            break
        Rd = rodrigues(delta[3:])
        R, t = compose(Rd, delta[:3], R, t)
        dR_deg, dt_norm = pose_magnitude(Rd, delta[:3])
        entry = dict(i=it, g=g.clone(), H=H.clone(), delta=delta.clone(), R=R.clone(), t=t.clone(),
                     n_valid=n_valid, cost_sum=float((valid.float() * cost).sum()),
                     dt=float(dt_norm), dR=float(dR_deg), gnorm=float(g.norm()))
        log.append(entry)
        if on_iter is not None:
            on_iter(entry, valid=valid, cost=cost, w_unc=w_unc, w_loss=w_loss, J=J)
        small_step = (dt_norm < dt_stop) and (dR_deg < dR_stop)
        if small_step or (g.norm() < grad_stop):
            break
    return dict(R=R, t=t, failed=failed, raised=raised, n_iters=len(log), log=log)


# ----------------------------------------------------------------------------
# reference-side sparse sampling
#   /root/reference/pixtrack/localization/pixloc_pose_refiners.py:327-368
# ----------------------------------------------------------------------------
def sample_reference(maps: Sequence[Tensor], scales: Sequence[Tuple[float, float]],
                     cam: Tensor, R: Tensor, t: Tensor, p3d: Tensor, pad: int = 1):
    """maps[l] is [(C_l+1), H_l, W_l] (descriptor channels + confidence last).
    Returns per-level observations [N, C_l+1] and the AND-over-levels keep mask
    (points are kept only if they project inside every level).
    The reference does the projection in the dtype of the COLMAP model, i.e.
    float64 (numpy xyz, qvec2rotmat, `Camera.from_colmap`), and casts the pixel
    coordinates to the map dtype only for the interpolation
    (`p2d_feat.to(feats)`, :349-351): pass float64 cam/R/t/p3d to get that."""
    p_cam = p3d @ R.t() + t
    obs, keep = [], torch.ones(p3d.shape[0], dtype=torch.bool, device=p3d.device)
    for Fm, sc in zip(maps, scales):
        uv, vis = world_to_image(scale_camera(cam, sc), p_cam)
        val, inb, _ = sample_map(Fm, uv.to(Fm.dtype), pad)
        obs.append(val)
        keep &= inb & vis
    return obs, keep


# ----------------------------------------------------------------------------
# coarse-to-fine level loop
#   /root/reference/pixloc/pixloc/localization/base_refiner.py:64-137
# ----------------------------------------------------------------------------
def refine_levels(maps_q: Sequence[Tensor], scales_q: Sequence[Tuple[float, float]],
                  cam: Tensor, R0: Tensor, t0: Tensor, obs_ref: Sequence[Tensor],
                  p3d: Tensor, lams: Sequence[Tensor], **lm_kw) -> Dict:
    """maps_q[l]: [(C_l+1),H_l,W_l]; obs_ref[l]: [N, C_l+1] (already filtered).
    Splits descriptor/confidence, L2-normalises the reference descriptors over
    C (base_refiner.py:80-84) and the query map over C at EVERY pixel
    (:92-94), then runs LM from the coarsest level to the finest, chaining T
    (:101-126).  Stops at the first failed level (:124-125)."""
    R, t = R0, t0
    runs = []
    for level in reversed(range(len(maps_q))):
        Fq, Wq = maps_q[level][:-1], maps_q[level][-1:]
        Fq = tF.normalize(Fq, dim=0)
        Fr, Wr = obs_ref[level][:, :-1], obs_ref[level][:, -1:]
        Fr = tF.normalize(Fr, dim=1)
        out = lm_run(p3d, Fr, Fq, R, t, scale_camera(cam, scales_q[level]),
                     W_ref=Wr, W_q=Wq, lam=lams[level], **lm_kw)
        runs.append(out)
        if out['failed'] or out['raised']:
            return dict(success=False, R=R, t=t, runs=runs)
        R, t = out['R'], out['t']
    return dict(success=True, R=R, t=t, runs=runs)
