"""Swap the B200 path into a live reference `PoseTrackerLocalizer`.

The reference builds its pipeline in `PoseTrackerLocalizer.__init__`
(pixtrack/localization/pixloc_pose_refiners.py:29-93): the per-level optimizers come out of the
checkpoint and are re-classed in place (`opt.__class__ = PixTrackOptimizer`, :71-72), the UNet is
wrapped in `PixTrackFeatureExtractor` (:75-77) and both are handed to `PoseTrackerRefiner`
(:84-92, stored as `refiner.optimizer` / `refiner.feature_extractor`,
pixloc/pixloc/localization/base_refiner.py:41-44).  `install()` does the same swap one level
further: every object the reference code reaches through those attributes becomes the CUDA
adapter, nothing else in the tracker changes.
"""
from .extractor import B200FeatureExtractor
from .optimizer import B200Optimizer


def install(localizer):
    """localizer: a reference PoseTrackerLocalizer (duck-typed: `.optimizer` list of per-level
    LearnedOptimizer/PixTrackOptimizer modules, `.extractor` PixTrackFeatureExtractor,
    `.refiner` with `.optimizer`, `.feature_extractor` and an optional `.tracker`).
    Returns (optimizers, extractor).  Raises PtkError if the CUDA library or an sm_100 device is
    missing -- there is no fallback to the PyTorch path."""
    opts = localizer.optimizer
    opts = list(opts) if hasattr(opts, '__iter__') else [opts]     # list, tuple or ModuleList
    new_opts = [B200Optimizer.from_reference(o) for o in opts]
    new_ext = B200FeatureExtractor.from_reference(localizer.extractor)
    localizer.optimizer = new_opts
    localizer.extractor = new_ext
    refiner = getattr(localizer, 'refiner', None)
    if refiner is not None:
        refiner.optimizer = new_opts
        refiner.feature_extractor = new_ext
        tracker = getattr(refiner, 'tracker', None)
        if tracker is not None:      # BaseTracker.__init__ attached its callback to the OLD optimizers
            for o in new_opts:
                o.logging_fn = tracker.log_optim_iter
    return new_opts, new_ext
