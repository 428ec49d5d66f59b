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

from .optimizer import lm_run_batched

Tensor = torch.Tensor


def refine_levels_batched(fq_hwc: Sequence[Tensor], wq: Sequence[Optional[Tensor]], cams: Sequence[Tensor],
                          F_ref: Sequence[Tensor], W_ref: Sequence[Optional[Tensor]], p3d: Tensor, T_init: Tensor,
                          lams: Sequence[Tensor], *, num_iters: int = 150, pad: int = 1, loss_scale: float = 0.1,
                          grad_stop: float = 1e-4, dt_stop: float = 5e-3, dR_stop: float = 5e-2,
                          want_log: bool = True):
    """All sequences are indexed by pyramid level, fine (0) to coarse (L-1);
    the levels are visited coarse to fine.
      fq_hwc[l] [B|1,H_l,W_l,C_l] (already L2-normalised over C), wq[l] [B|1,H_l,W_l],
      cams[l] [B|1,n_cam] (camera already scaled to the level), F_ref[l] [B,N,C_l]
      (already normalised), W_ref[l] [B,N], p3d [B|1,N,3], T_init [B,12], lams[l] [6].
    Returns dict(T [B,12], failed [B] uint8, n_iters [L][B], logs [L]) -- levels
    in visiting order (coarse first), like DebugTracker.costs."""
    T, skip = T_init, None
    n_all, logs = [], []
    for lv in reversed(range(len(fq_hwc))):
        T, failed, n_it, log = lm_run_batched(
            p3d, F_ref[lv], fq_hwc[lv], T, cams[lv], lams[lv], W_ref[lv], wq[lv], None, skip,
            num_iters=num_iters, pad=pad, loss_scale=loss_scale, grad_stop=grad_stop, dt_stop=dt_stop,
            dR_stop=dR_stop, want_log=want_log)
        skip = failed
        n_all.append(n_it)
        logs.append(log)
    return dict(T=T, failed=skip, n_iters=n_all, logs=logs)
