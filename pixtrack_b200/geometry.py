"""Pose / Camera tensor wrappers with the reference's interface.

The drop-in optimizer accepts the reference's own `pixloc.pixlib.geometry.
{Pose, Camera}` objects (duck-typed through `._data`); these local classes
exist so the package, tests and bench work where /root/reference is absent.
Same data layout as reference pixloc/pixloc/pixlib/geometry/wrappers.py:
Pose._data = [R row-major (9), t (3)], Camera._data = [w, h, fx, fy, cx, cy,
dist...] with 0, 2 or 4 distortion terms.
"""
import math
from typing import Sequence, Tuple, Union

import numpy as np
import torch

Tensor = torch.Tensor


def _as_tensor(x, like: Tensor = None) -> Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if like is not None:
        x = x.to(device=like.device, dtype=like.dtype)
    return x


def rotation_angle(Ra: Tensor, Rb: Tensor) -> Tensor:
    """Geodesic angle in radians between rotations [...,3,3], accurate near zero: atan2(|vee(M - M^T)| / 2,
    (tr M - 1) / 2) with M = Ra Rb^T in float64.  (acos of the trace alone -- Pose.magnitude, wrappers.py:208-218 -- has
    a noise floor of sqrt(2 eps) = 3e-4..5e-4 rad on float32 matrices, far above the 1e-4 rad parity bound.)"""
    M = Ra.double() @ Rb.double().transpose(-1, -2)
    v = torch.stack([M[..., 2, 1] - M[..., 1, 2], M[..., 0, 2] - M[..., 2, 0], M[..., 1, 0] - M[..., 0, 1]], -1)
    return torch.atan2(0.5 * v.norm(dim=-1), 0.5 * (M.diagonal(dim1=-2, dim2=-1).sum(-1) - 1.0))


def pose_distance(Ta: Tensor, Tb: Tensor) -> Tuple[Tensor, Tensor]:
    """(rotation angle in rad, translation distance) between 12-vectors [R row-major, t]."""
    Ta, Tb = Ta.double(), Tb.double()
    return (rotation_angle(Ta[..., :9].reshape(Ta.shape[:-1] + (3, 3)), Tb[..., :9].reshape(Tb.shape[:-1] + (3, 3))),
            (Ta[..., 9:] - Tb[..., 9:]).norm(dim=-1))


class _Wrapper:
    def __init__(self, data: Tensor):
        self._data = _as_tensor(data)

    @property
    def shape(self):
        return self._data.shape[:-1]

    @property
    def device(self):
        return self._data.device

    @property
    def dtype(self):
        return self._data.dtype

    def to(self, *a, **k):
        if a and isinstance(a[0], _Wrapper):
            a = (a[0]._data,) + a[1:]
        return self.__class__(self._data.to(*a, **k))

    def cpu(self):
        return self.__class__(self._data.cpu())

    def cuda(self):
        return self.__class__(self._data.cuda())

    def float(self):
        return self.__class__(self._data.float())

    def double(self):
        return self.__class__(self._data.double())

    def __getitem__(self, i):
        return self.__class__(self._data[i])


class Pose(_Wrapper):
    def __init__(self, data):
        super().__init__(data)
        assert self._data.shape[-1] == 12

    @classmethod
    def from_Rt(cls, R, t):
        R, t = _as_tensor(R), _as_tensor(t)
        return cls(torch.cat([R.flatten(start_dim=-2), t.to(R)], -1))

    @classmethod
    def from_aa(cls, aa, t):
        aa = _as_tensor(aa)
        theta = aa.norm(dim=-1, keepdim=True)
        small = theta < 1e-7
        k = aa / torch.where(small, torch.ones_like(theta), theta)
        z = torch.zeros_like(k[..., 0])
        W = torch.stack([z, -k[..., 2], k[..., 1], k[..., 2], z, -k[..., 0], -k[..., 1], k[..., 0], z], -1)
        W = W.reshape(aa.shape[:-1] + (3, 3))
        th = theta[..., None]
        res = torch.where(small[..., None], W, W * torch.sin(th) + (W @ W) * (1 - torch.cos(th)))
        return cls.from_Rt(torch.eye(3).to(W) + res, t)

    @property
    def R(self) -> Tensor:
        return self._data[..., :9].reshape(self._data.shape[:-1] + (3, 3))

    @property
    def t(self) -> Tensor:
        return self._data[..., 9:]

    def inv(self) -> 'Pose':
        Rt = self.R.transpose(-1, -2)
        return Pose.from_Rt(Rt, -(Rt @ self.t[..., None])[..., 0])

    def compose(self, other: 'Pose') -> 'Pose':
        return Pose.from_Rt(self.R @ other.R, self.t + (self.R @ other.t[..., None])[..., 0])

    __matmul__ = compose

    def transform(self, p3d) -> Tensor:
        p3d = _as_tensor(p3d, self._data)
        return p3d @ self.R.transpose(-1, -2) + self.t[..., None, :]

    __mul__ = transform

    def magnitude(self) -> Tuple[Tensor, Tensor]:
        tr = torch.diagonal(self.R, dim1=-1, dim2=-2).sum(-1)
        cos = torch.clamp((tr - 1) / 2, -1, 1)
        return torch.acos(cos).abs() / math.pi * 180, torch.norm(self.t, dim=-1)

    def numpy(self):
        return self.R.numpy(), self.t.numpy()


class Camera(_Wrapper):
    def __init__(self, data):
        super().__init__(data)
        assert self._data.shape[-1] in (6, 8, 10)

    @classmethod
    def from_colmap(cls, camera) -> 'Camera':
        """COLMAP camera dict/tuple -> Camera; pixel-centre origin (c - 0.5)."""
        if hasattr(camera, '_asdict'):
            camera = camera._asdict()
        model, params = camera['model'], np.asarray(camera['params'], dtype=np.float64)
        if model in ('OPENCV', 'PINHOLE'):
            (fx, fy, cx, cy), rest = params[:4], params[4:]
        elif model in ('SIMPLE_PINHOLE', 'SIMPLE_RADIAL', 'RADIAL'):
            (f, cx, cy), rest = params[:3], params[3:]
            fx = fy = f
            if model == 'SIMPLE_RADIAL':
                rest = np.r_[rest, 0.]
        else:
            raise NotImplementedError(model)
        return cls(torch.from_numpy(np.r_[camera['width'], camera['height'], fx, fy, cx - 0.5, cy - 0.5, rest]))

    @property
    def size(self):
        return self._data[..., :2]

    @property
    def f(self):
        return self._data[..., 2:4]

    @property
    def c(self):
        return self._data[..., 4:6]

    @property
    def dist(self):
        return self._data[..., 6:]

    def scale(self, scales: Union[float, int, Sequence[float]]) -> 'Camera':
        if isinstance(scales, (int, float)):
            scales = (scales, scales)
        s = self._data.new_tensor(scales)
        return Camera(torch.cat([self.size * s, self.f * s, (self.c + 0.5) * s - 0.5, self.dist], -1))

    # ---- host-side projection chain (interface of wrappers.py:299-355; not on the per-iteration path: the LM kernel and
    #      the reference sampler carry their own copies in registers; this serves code that calls the Camera directly,
    #      e.g. the reference's interp_sparse_observations, pixloc_pose_refiners.py:347) -------------------------------
    eps = 1e-3                                                  # wrappers.py:225

    def in_image(self, p2d: Tensor) -> Tensor:
        """0 <= p <= size - 1 on the (float) camera size."""
        p2d = _as_tensor(p2d, self._data)
        hi = (self.size - 1).unsqueeze(-2)
        return ((p2d >= 0) & (p2d <= hi)).all(-1)

    def project(self, p3d: Tensor) -> Tuple[Tensor, Tensor]:
        p3d = _as_tensor(p3d, self._data)
        z = p3d[..., 2]
        return p3d[..., :2] / z.clamp(min=self.eps).unsqueeze(-1), z > self.eps

    def undistort(self, pts: Tensor) -> Tuple[Tensor, Tensor]:
        """Applies the radial (+ tangential) model to normalised coordinates and flags points beyond the model's inflection
        radius (utils.py:36-69; the reference calls this direction `undistort` too)."""
        pts = _as_tensor(pts, self._data)
        d = self.dist.unsqueeze(-2)
        ok = torch.ones(pts.shape[:-1], dtype=torch.bool, device=pts.device)
        if d.shape[-1] == 0:
            return pts, ok
        k1, k2 = d[..., 0:1], d[..., 1:2]
        r2 = pts.pow(2).sum(-1, keepdim=True)
        out = pts * (1 + k1 * r2 + k2 * r2 * r2)
        disc = 9 * k1 * k1 - 20 * k2
        bounded = ((k2 > 0) & (disc > 0)) | ((k2 <= 0) & (k1 > 0))
        radius2 = torch.where(k2 > 0, (disc.clamp(min=0).sqrt() - 3 * k1) / (10 * k2), 1 / (3 * k1)).abs()
        ok = ok & (~bounded | (r2 < radius2)).squeeze(-1)
        if d.shape[-1] > 2:
            p = d[..., 2:4]
            out = out + 2 * p * pts.prod(-1, keepdim=True) + p.flip(-1) * (r2 + 2 * pts * pts)
        return out, ok

    def denormalize(self, p2d: Tensor) -> Tensor:
        return _as_tensor(p2d, self._data) * self.f.unsqueeze(-2) + self.c.unsqueeze(-2)

    def world2image(self, p3d: Tensor) -> Tuple[Tensor, Tensor]:
        """Camera-frame 3-D points -> (pixel coordinates, valid = in front & inside the model's radius & inside the image)."""
        xy, front = self.project(p3d)
        xy, ok = self.undistort(xy)
        uv = self.denormalize(xy)
        return uv, front & ok & self.in_image(uv)
