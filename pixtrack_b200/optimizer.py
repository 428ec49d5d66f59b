"""B200Optimizer: drop-in for the reference's per-level LM optimizer.

Mirrors the interface of `PixTrackOptimizer` (reference
pixtrack/optimizers/pixtrack_optimizer.py:5, pixloc/pixloc/pixlib/models/
learned_optimizer.py:30-95, base_optimizer.py:23-105): a `torch.nn.Module`
with `.conf`, `.dampingnet.const`, `.interpolator`, `.logging_fn` and
`run(p3D, F_ref, F_query, T_init, camera, mask=None, W_ref_query=None)
-> (Pose, failed)`, so `BaseRefiner.refine_pose_using_features`
(pixloc/pixloc/localization/base_refiner.py:117-119) and
`PoseTrackerRefiner.interp_sparse_observations`
(pixtrack/localization/pixloc_pose_refiners.py:349-351) call it unchanged.
The whole iteration loop runs inside one CUDA launch (csrc/ptk_lm.cu) through
the C ABI; there is no PyTorch/CPU fallback.
"""
import ctypes as C
import re
from types import SimpleNamespace
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .geometry import Camera, Pose

Tensor = torch.Tensor

DEFAULT_CONF = dict(
    num_iters=100,                    # base_optimizer.py:24-36
    loss_fn='squared_loss',
    jacobi_scaling=False,
    normalize_features=False,
    lambda_=0.0,
    interpolation=dict(mode='linear', pad=4),
    grad_stop_criteria=1e-4,
    dt_stop_criteria=5e-3,
    dR_stop_criteria=5e-2,
    damping=dict(type='constant', log_range=[-6, 5]),   # learned_optimizer.py:31-35
    learned_damping=True,
)


def _ns(d):
    return SimpleNamespace(**{k: _ns(v) if isinstance(v, dict) else v for k, v in d.items()})


def _merge_conf(user) -> SimpleNamespace:
    conf = {k: (dict(v) if isinstance(v, dict) else v) for k, v in DEFAULT_CONF.items()}
    user = dict(user or {})
    if 'pad' in user:                 # legacy key, base_model.py:69-72
        user.setdefault('interpolation', {})
        user['interpolation'] = {**dict(user['interpolation']), 'pad': user.pop('pad')}
    for k, v in user.items():
        if k in ('interpolation', 'damping'):
            conf[k].update(dict(v))
        else:
            conf[k] = v
    return _ns(conf)


def _conf_from_reference(ref_conf) -> dict:
    keys = ('num_iters', 'loss_fn', 'jacobi_scaling', 'normalize_features', 'grad_stop_criteria',
            'dt_stop_criteria', 'dR_stop_criteria')
    out = {k: ref_conf[k] for k in keys if k in ref_conf}
    out['interpolation'] = dict(mode=ref_conf['interpolation']['mode'], pad=ref_conf['interpolation']['pad'])
    if 'damping' in ref_conf:
        out['damping'] = dict(type=ref_conf['damping']['type'], log_range=list(ref_conf['damping']['log_range']))
    return out


def parse_loss(loss_fn: str) -> float:
    """Only the PixLoc/PixTrack loss is on the hot path: scaled_barron(0, c)
    (checkpoint conf; geometry/losses.py:80-82).  Returns c."""
    m = re.fullmatch(r'\s*scaled_barron\(\s*0(?:\.0*)?\s*,\s*([0-9.eE+-]+)\s*\)\s*', loss_fn)
    if not m:
        raise NotImplementedError(f'loss_fn {loss_fn!r}: the B200 path implements scaled_barron(0, c) only')
    return float(m.group(1))


class DampingNet(torch.nn.Module):
    """learned_optimizer.py:14-27: lambda = 10^(min + sigmoid(const) * (max - min))."""

    def __init__(self, log_range=(-6.0, 5.0), num_params: int = 6):
        super().__init__()
        self.log_range = (float(log_range[0]), float(log_range[1]))
        self.const = torch.nn.Parameter(torch.zeros(num_params))

    def forward(self) -> Tensor:
        lo, hi = self.log_range
        return 10. ** (lo + self.const.sigmoid() * (hi - lo))


def _f32c(x: Tensor) -> Tensor:
    return x.to(torch.float32).contiguous()


def _ptr(x: Optional[Tensor]):
    return None if x is None else x.data_ptr()


def query_map_to_hwc(F_q: Tensor, normalize: bool = False) -> Tensor:
    """[C,H,W] (any strides) -> contiguous [H,W,C] fp32 on the same device.
    Zero-copy when F_q is already a channels-last view (what B200FeatureExtractor
    returns); otherwise one pass of csrc ptk_chw_to_hwc."""
    assert F_q.is_cuda and F_q.dim() == 3
    hwc = F_q.permute(1, 2, 0)
    if not normalize and F_q.dtype == torch.float32 and hwc.is_contiguous():
        return hwc
    src = _f32c(F_q)
    Cc, H, W = src.shape
    dst = torch.empty((H, W, Cc), dtype=torch.float32, device=src.device)
    dev = src.device.index if src.device.index is not None else torch.cuda.current_device()
    _lib.check(_lib.load().ptk_chw_to_hwc(_lib.context(dev), src.data_ptr(), dst.data_ptr(), Cc, H, W,
                                          1 if normalize else 0, _lib.current_stream_ptr(src.device)))
    return dst


class LmLaunch:
    """A prepared ptk_lm_run call: argument struct, output buffers and keep-alive
    references are built once; `launch()` is then a single ctypes call (a few
    microseconds of host time), so chains of launches stay GPU-bound and can be
    captured in a CUDA graph.

    Shapes (fp32 CUDA, contiguous; a leading batch dim of 1 / a missing batch
    dim means "shared by all B problems"):
      p3d [B|1,N,3]  F_ref [B,N,C]  fq_hwc [B|1,H,W,C]  T_init [B,12]
      cam [B|1,6|8|10]  lam [B|1,6]  W_ref [B,N]  wq [B|1,H,W]  mask [B,N] uint8/bool
      skip [B] uint8 (e.g. the `failed` output of the previous level's launch).
    Outputs: T [B,12], failed [B] uint8, n_iters [B] int32, log [B,num_iters,64] or None."""

    def __init__(self, p3d: Tensor, F_ref: Tensor, fq_hwc: Tensor, T_init: Tensor, cam: Tensor, lam: Tensor,
                 W_ref: Optional[Tensor] = None, wq: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                 skip: Optional[Tensor] = None, *, num_iters: int, pad: int = 1, loss_scale: float = 0.1,
                 grad_stop: float = 1e-4, dt_stop: float = 5e-3, dR_stop: float = 5e-2, min_valid: int = 10,
                 want_log: bool = True):
        dev = F_ref.device
        if dev.type != 'cuda':
            raise _lib.PtkError('the LM path needs CUDA tensors (no CPU fallback)')
        B, N, Cc = F_ref.shape

        def prep(x, nd):
            x = _f32c(x)
            return x if x.dim() == nd else x[None]

        p3d, fq_hwc, cam, lam = prep(p3d, 3), prep(fq_hwc, 4), prep(cam, 2), prep(lam, 2)
        F_ref, T_init = _f32c(F_ref), _f32c(T_init).reshape(B, 12)
        H, W = fq_hwc.shape[1:3]
        assert fq_hwc.shape[3] == Cc and p3d.shape[1] == N

        def bs(x, per):
            return 0 if x.shape[0] == 1 and B > 1 else per

        prob = _lib.LmProblem()
        prob.B, prob.N, prob.C, prob.H, prob.W = B, N, Cc, H, W
        prob.n_cam, prob.num_iters, prob.pad, prob.min_valid = cam.shape[-1], int(num_iters), int(pad), int(min_valid)
        prob.p3d, prob.p3d_bstride = p3d.data_ptr(), bs(p3d, N * 3)
        prob.f_ref, prob.f_ref_bstride = F_ref.data_ptr(), N * Cc
        keep = [p3d, F_ref, fq_hwc, cam, lam, T_init]
        if W_ref is not None:
            W_ref = _f32c(W_ref).reshape(B, N)
            wq = prep(wq.reshape(-1, H, W) if wq.dim() > 2 else wq, 3)
            prob.w_ref, prob.w_ref_bstride = W_ref.data_ptr(), N
            prob.wq, prob.wq_bstride = wq.data_ptr(), bs(wq, H * W)
            keep += [W_ref, wq]
        prob.fq, prob.fq_bstride = fq_hwc.data_ptr(), bs(fq_hwc, H * W * Cc)
        if mask is not None:
            mask = mask.to(torch.uint8).contiguous().reshape(B, N)
            prob.mask, prob.mask_bstride = mask.data_ptr(), N
            keep.append(mask)
        prob.cam, prob.cam_bstride = cam.data_ptr(), bs(cam, cam.shape[-1])
        prob.T_init, prob.T_bstride = T_init.data_ptr(), 12
        prob.lambda_, prob.lambda_bstride = lam.data_ptr(), bs(lam, 6)
        if skip is not None:
            assert skip.dtype == torch.uint8 and skip.is_contiguous()
            prob.skip = skip.data_ptr()
            keep.append(skip)
        prob.loss_scale, prob.grad_stop, prob.dt_stop, prob.dR_stop = loss_scale, grad_stop, dt_stop, dR_stop

        self.T = torch.empty((B, 12), dtype=torch.float32, device=dev)
        self.failed = torch.empty((B,), dtype=torch.uint8, device=dev)
        self.n_iters = torch.empty((B,), dtype=torch.int32, device=dev)
        self.log = (torch.zeros((B, max(1, num_iters), _lib.LOG_STRIDE), dtype=torch.float32, device=dev)
                    if want_log else None)
        # own barrier / partial-sum scratch: prepared launches on one device may run concurrently on different streams
        # (zero-initialised once; the kernel leaves it zeroed)
        ws_bytes = int(_lib.load().ptk_lm_workspace_bytes())
        self._workspace = torch.zeros((ws_bytes + 3) // 4, dtype=torch.int32, device=dev)
        prob.workspace, prob.workspace_bytes = self._workspace.data_ptr(), ws_bytes
        self._res = _lib.LmResult(self.T.data_ptr(), self.failed.data_ptr(), self.n_iters.data_ptr(), _ptr(self.log))
        self._prob, self._keep, self.device = prob, keep, dev
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        self._ctx, self._fn = _lib.context(di), _lib.load().ptk_lm_run
        self.shape = (B, N, Cc, H, W)

    def launch(self, stream_ptr: Optional[int] = None):
        """Stream-ordered, never synchronises.  Returns self for chaining."""
        if stream_ptr is None:
            stream_ptr = torch.cuda.current_stream(self.device).cuda_stream
        code = self._fn(self._ctx, C.byref(self._prob), C.byref(self._res), stream_ptr)
        if code != 0:
            _lib.check(code)
        return self

    def plan(self):
        """(CTAs per problem, problems in flight) the launch uses."""
        g, n = C.c_int32(), C.c_int32()
        _lib.check(_lib.load().ptk_lm_plan(self._ctx, C.byref(self._prob), C.byref(g), C.byref(n)))
        return g.value, n.value


def lm_run_batched(*args, **kw):
    """One-shot form: build an LmLaunch, launch it, return
    (T [B,12], failed [B] uint8, n_iters [B] int32, log or None); nothing is synchronised."""
    L = LmLaunch(*args, **kw).launch()
    return L.T, L.failed, L.n_iters, L.log


def unpack_H(rec: np.ndarray) -> np.ndarray:
    H = np.zeros((6, 6), dtype=rec.dtype)
    H[np.triu_indices(6)] = rec[32:53]
    return H + np.triu(H, 1).T


class B200Optimizer(torch.nn.Module):
    """One pyramid level's optimizer (the reference keeps a ModuleList of 3)."""
    logging_fn = None   # set by BaseTracker.__init__ (pixloc/pixloc/localization/tracker.py:5-13)

    def __init__(self, conf=None):
        super().__init__()
        self.conf = _merge_conf(conf)
        assert self.conf.interpolation.mode == 'linear', 'B200 path implements linear interpolation (PixTrack conf)'
        assert not self.conf.jacobi_scaling and not self.conf.normalize_features
        self.loss_scale = parse_loss(self.conf.loss_fn)
        self.dampingnet = DampingNet(self.conf.damping.log_range)
        from .sampling import Interpolator
        self.interpolator = Interpolator(mode='linear', pad=self.conf.interpolation.pad)

    @classmethod
    def from_reference(cls, ref_opt) -> 'B200Optimizer':
        """Build from a live reference LearnedOptimizer / PixTrackOptimizer."""
        new = cls(_conf_from_reference(ref_opt.conf))
        new.dampingnet.const.data.copy_(ref_opt.dampingnet.const.data)
        new.logging_fn = getattr(ref_opt, 'logging_fn', None)
        return new.to(ref_opt.dampingnet.const.device)

    def early_stop(self, **args):   # kept for API parity; the test runs on the device (ptk_lm.cu)
        raise RuntimeError('early_stop is evaluated inside the CUDA kernel')

    def forward(self, data):
        return self._run(data['p3D'], data['F_ref'], data['F_q'], data['T_init'], data['cam_q'], data['mask'],
                         data.get('W_ref_q'))

    def run(self, p3D, F_ref, F_query, T_init, camera, mask=None, W_ref_query=None):
        """numpy inputs are moved to the tensors' device/dtype, like @torchify
        (pixloc/pixloc/utils/tools.py:6-67)."""
        dev = F_query.device
        if isinstance(p3D, np.ndarray):
            p3D = torch.from_numpy(p3D).to(dev, torch.float32)
        if isinstance(F_ref, np.ndarray):
            F_ref = torch.from_numpy(F_ref).to(dev, torch.float32)
        return self._run(p3D, F_ref, F_query, T_init, camera, mask, W_ref_query)

    @torch.no_grad()
    def _run(self, p3D: Tensor, F_ref: Tensor, F_query: Tensor, T_init, camera, mask: Optional[Tensor] = None,
             W_ref_query: Optional[Tuple[Tensor, Tensor]] = None):
        dev = F_query.device
        if dev.type != 'cuda':
            raise _lib.PtkError('B200Optimizer needs CUDA tensors: the LM loop exists only as an sm_100a kernel')
        assert T_init._data.dim() == 1, 'one problem per call at this boundary (use lm_run_batched for batches)'
        W_ref = wq = None
        if W_ref_query is not None:
            W_ref, wq = W_ref_query
            W_ref, wq = W_ref.to(dev).reshape(1, -1), wq.to(dev)
        lam = self.dampingnet().to(dev)
        c = self.conf
        T, failed, n_iters, log = lm_run_batched(
            p3D.to(dev)[None], F_ref.to(dev)[None], query_map_to_hwc(F_query)[None], T_init._data.to(dev)[None],
            camera._data.to(dev)[None], lam[None], W_ref, wq, None if mask is None else mask.to(dev)[None],
            num_iters=c.num_iters, pad=c.interpolation.pad, loss_scale=self.loss_scale,
            grad_stop=c.grad_stop_criteria, dt_stop=c.dt_stop_criteria, dR_stop=c.dR_stop_criteria,
            want_log=self.logging_fn is not None)
        PoseCls = T_init.__class__
        T_out = PoseCls(T[0].to(T_init._data.dtype))
        if self.logging_fn is not None:
            self.replay_log(T_init, log[0], int(n_iters[0]))
        return T_out, failed[0].bool()

    def replay_log(self, T_init, log: Tensor, n: int):
        """Calls logging_fn once per executed iteration with the reference's
        kwarg names (base_optimizer.py:140).  Per-point arrays are not kept on
        the fast path: `valid`=[1], `cost`=[cost_sum/n_valid] reproduce exactly
        what DebugTracker/SimpleTracker compute from them
        ((valid*cost).sum(-1)/valid.sum(-1), pixtrack/localization/tracker.py:40-41)."""
        rec = log[:n].cpu()
        PoseCls = T_init.__class__
        for i in range(n):
            r = rec[i]
            T_i = PoseCls(r[2:14].clone())
            T_delta = PoseCls.from_aa(r[26:29].clone(), r[23:26].clone())
            cost = (r[0] / r[1]).reshape(1)
            self.logging_fn(i=i, T_init=T_init, T=T_i, T_delta=T_delta, cost=cost,
                            valid=torch.ones(1, dtype=torch.bool), w_unc=None, w_loss=None,
                            H=torch.from_numpy(unpack_H(r.numpy())), J=None)
