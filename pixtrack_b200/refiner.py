"""Coarse-to-fine refinement over the feature pyramid, chained on the device.

`refine_levels_batched` is the B200 form of the level loop in
`BaseRefiner.refine_pose_using_features` (reference
pixloc/pixloc/localization/base_refiner.py:96-126): coarsest level first, the
pose of one level initialises the next, and a failed level stops the chain
(:124-125).  Here the chain is three stream-ordered launches of the fused LM
kernel with the previous level's `failed` flags passed as `skip`, so there is
no host synchronisation inside a frame; B reference views / frames are solved
together.
"""
from typing import List, Optional, Sequence

import torch

from .optimizer import LmLaunch, lm_run_batched

Tensor = torch.Tensor


def refine_levels_batched(fq_hwc: Sequence[Tensor], wq: Sequence[Optional[Tensor]], cams: Sequence[Tensor],
                          F_ref: Sequence[Tensor], W_ref: Sequence[Optional[Tensor]], p3d: Tensor, T_init: Tensor,
                          lams: Sequence[Tensor], *, num_iters: int = 150, pad: int = 1, loss_scale: float = 0.1,
                          grad_stop: float = 1e-4, dt_stop: float = 5e-3, dR_stop: float = 5e-2,
                          want_log: bool = True, mask: Optional[Tensor] = None):
    """All sequences are indexed by pyramid level, fine (0) to coarse (L-1);
    the levels are visited coarse to fine.
      fq_hwc[l] [B|1,H_l,W_l,C_l] (already L2-normalised over C), wq[l] [B|1,H_l,W_l],
      cams[l] [B|1,n_cam] (camera already scaled to the level), F_ref[l] [B,N,C_l]
      (already normalised), W_ref[l] [B,N], p3d [B|1,N,3], T_init [B,12], lams[l] [6];
      mask [B,N] uint8: points to use (the `valid` flags of sample_reference).
    Returns dict(T [B,12], failed [B] uint8, n_iters [L][B], logs [L]) -- levels
    in visiting order (coarse first), like DebugTracker.costs."""
    T, skip = T_init, None
    n_all, logs = [], []
    for lv in reversed(range(len(fq_hwc))):
        T, failed, n_it, log = lm_run_batched(
            p3d, F_ref[lv], fq_hwc[lv], T, cams[lv], lams[lv], W_ref[lv], wq[lv], mask, skip,
            num_iters=num_iters, pad=pad, loss_scale=loss_scale, grad_stop=grad_stop, dt_stop=dt_stop,
            dR_stop=dR_stop, want_log=want_log)
        skip = failed
        n_all.append(n_it)
        logs.append(log)
    return dict(T=T, failed=skip, n_iters=n_all, logs=logs)


class FramePlan:
    """The coarse-to-fine chain of one frame batch as prepared launches.

    Built once over STATIC buffers (query pyramid, reference observations,
    initial poses); `run()` is then L bare ctypes launches (coarse -> fine),
    each reading the previous level's pose / failed outputs directly on the
    device.  `capture()` records the chain into a CUDA graph so a frame costs
    one graph launch."""

    def __init__(self, fq_hwc, wq, cams, F_ref, W_ref, p3d, T_init, lams, mask=None, **kw):
        self.launches = []
        T, skip = T_init, None
        for lv in reversed(range(len(fq_hwc))):
            L = LmLaunch(p3d, F_ref[lv], fq_hwc[lv], T, cams[lv], lams[lv], W_ref[lv], wq[lv], mask, skip, **kw)
            self.launches.append(L)
            T, skip = L.T, L.failed
        self.T, self.failed = T, skip
        self.graph = None

    def run(self):
        if self.graph is not None:
            self.graph.replay()
            return self
        sp = torch.cuda.current_stream(self.launches[0].device).cuda_stream
        for L in self.launches:
            L.launch(sp)
        return self

    def capture(self):
        """Capture the chain in a CUDA graph (falls back to direct launches if
        the driver refuses to capture a cooperative launch)."""
        dev = self.launches[0].device
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            self.run()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=s):
                for L in self.launches:
                    L.launch(s.cuda_stream)
            self.graph = g
        except Exception:       # noqa: BLE001 - capture is an optimisation, the direct chain is equivalent
            self.graph = None
            torch.cuda.synchronize(dev)
        return self

    @property
    def n_iters(self):
        return [L.n_iters for L in self.launches]

    @property
    def logs(self):
        return [L.log for L in self.launches]
