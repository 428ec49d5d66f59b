"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (PyTorch fp32, no CUDA) restatement of the reference algorithms on the
PixTrack pose-refinement hot path.  Every function cites
the reference file:line it follows (paths relative to /root/reference).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs -- only as the checker or
as the timed CPU baseline.  Nothing under `pixtrack_b200/` imports it; the
product path fails loudly when its CUDA library is missing.

Parity pinning: the reference holds NO golden vectors for this path
(SURVEY.md section 4 / 8c).  The oracle is therefore pinned against outputs of
the unmodified reference executed in the authoring container:
`tests/golden/gen/make_goldens.py` imports `/root/reference/{pixloc,pixtrack}`
(with a throw-away omegaconf stand-in), runs `PixTrackOptimizer.run`,
`interpolate_tensor`, `Camera.world2image`, `UNet._forward`, ... on seeded
inputs and stores inputs + outputs in `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every oracle function against them.
The NeRF render has no runnable reference here (pyngp cannot be built or run):
its host-callable header code is compiled in place (oracle/build_ref.py ->
oracle/_ref/ngp_host) and pins the jitter, colour transfer, camera conversion,
ray generation, box test, step sizes, cascade / cell / table indices, occupancy
lookup, empty-space stepping, hash-grid and SH encodings, ray start, the
compositing loop, shade, accumulate, compaction and tonemap of oracle/nerf.py
(tests/golden/nerf_host.json); the fused MLPs and the order in which render()
combines the pieces stay "parity unpinned".
"""
