"""NerfTestbed: drop-in for the slice of pyngp's `Testbed` that PixTrack uses.

The reference renders reference views (and a depth mask) with instant-ngp through
`get_nerf_image(testbed, nerf_pose, camera, depth=False)` (reference
pixtrack/visualization/run_vis_on_poses.py:28-57), which touches exactly: `testbed.fov`,
`testbed.set_nerf_camera_matrix(3x4)`, `testbed.render_mode` (+ `.Depth` / `.Shade`),
`testbed.render(width, height, spp, linear)`; `initialize_ingp` (pixtrack/utils/ingp_utils.py:22-44)
additionally sets `background_color`, `snap_to_pixel_centers`, `nerf.rendering_min_transmittance`,
`nerf.render_with_camera_distortion`, `nerf.sharpen`, `fov_axis`, `shall_train`, `render_aabb.min/max`,
`exposure`.  This class keeps those attribute names; the render itself is one persistent CUDA launch
(csrc/ptk_nerf.cu) through the C ABI.  There is no CPU fallback.
"""
import ctypes as C
import math
from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

GRID, CASCADES = 128, 8


def _expand_bits(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def _compact_bits(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32) & np.uint32(0x49249249)
    x = (x | (x >> np.uint32(2))) & np.uint32(0xc30c30c3)
    x = (x | (x >> np.uint32(4))) & np.uint32(0x0f00f00f)
    x = (x | (x >> np.uint32(8))) & np.uint32(0xff0000ff)
    x = (x | (x >> np.uint32(16))) & np.uint32(0x0000ffff)
    return x


def occupancy_bitfield(density_grid: np.ndarray, max_cascade: int) -> np.ndarray:
    """Snapshot `density_grid` (float/half, Morton order, (max_cascade+1) x 128^3 cells) -> the marcher's
    bitfield over all 8 cascades: Testbed::update_density_grid_mean_and_bitfield
    (instant-ngp/src/testbed_nerf.cu:2709-2724) = grid_to_bitfield (:555-581, threshold
    min(0.01, mean of the first cascade)) + bitfield_max_pool (:583-604).  Load-time host code."""
    n = GRID ** 3
    d = np.asarray(density_grid, np.float32).ravel()
    assert d.size >= (max_cascade + 1) * n, 'density grid too small'
    mean = np.float32(np.maximum(d[:n], 0).sum(dtype=np.float64) / n)
    thresh = min(np.float32(0.01), mean)
    bits = np.zeros(CASCADES * n // 8, np.uint8)
    occ = d[:(max_cascade + 1) * n] > thresh
    bits[:occ.size // 8] = np.packbits(occ.reshape(-1, 8), axis=1, bitorder='little').ravel()
    i = np.arange(n // 64, dtype=np.uint32)
    dst = (_expand_bits(_compact_bits(i) + GRID // 8) | (_expand_bits(_compact_bits(i >> 1) + GRID // 8) << 1) |
           (_expand_bits(_compact_bits(i >> 2) + GRID // 8) << 2)).astype(np.int64)
    for lv in range(1, CASCADES):
        prev = bits[(lv - 1) * n // 8: lv * n // 8]
        pooled = np.packbits(prev.reshape(-1, 8) > 0, axis=1, bitorder='little').ravel()
        np.bitwise_or.at(bits[lv * n // 8:(lv + 1) * n // 8], dst, pooled)
    return bits


def split_params(params: np.ndarray, aabb_scale: int):
    """`snapshot['params_binary']` (fp16) -> (density weights, rgb weights, grid) in
    NerfNetwork::set_params order (instant-ngp/include/neural-graphics-primitives/nerf_network.h:361-395)."""
    p = np.asarray(params, np.float16).ravel()
    shapes = [(64, 32), (16, 64), (64, 32), (64, 64), (16, 64)]
    mats, o = [], 0
    for r, c in shapes:
        mats.append(p[o:o + r * c].reshape(r, c))
        o += r * c
    n = int(_lib.load().ptk_nerf_grid_entries(int(aabb_scale)))
    grid = p[o:o + 2 * n]
    if grid.size != 2 * n:
        raise ValueError(f'params_binary holds {p.size} values; aabb_scale {aabb_scale} needs {o + 2 * n}')
    return (mats[0], mats[1]), (mats[2], mats[3], mats[4]), grid.reshape(n, 2)


class RenderMode:
    """Enum-like stand-in for pyngp.RenderMode: instances reach `.Shade` / `.Depth` through the class, which
    is how the reference switches modes (`testbed.render_mode = testbed.render_mode.Depth`)."""

    def __init__(self, name: str):
        self.name = name

    def __eq__(self, other):
        return self.name == (other.name if isinstance(other, RenderMode) else other)

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return f'RenderMode.{self.name}'


RenderMode.Shade = RenderMode('Shade')
RenderMode.Depth = RenderMode('Depth')


class NerfTestbed:
    def __init__(self, grid, w_density: Sequence, w_rgb: Sequence, bitfield, aabb_scale: int, device,
                 scale: float = 0.33, offset=(0.5, 0.5, 0.5), render_aabb=None):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.PtkError('NerfTestbed needs a CUDA device (no CPU fallback)')
        self._lib = _lib.load()
        self._ctx = _lib.context(self.device.index if self.device.index is not None else torch.cuda.current_device())

        def dev(a, dt):
            return torch.as_tensor(np.ascontiguousarray(a)).to(self.device, dt).contiguous()
        self._grid = dev(np.asarray(grid, np.float16), torch.float16)
        self._w = [dev(np.asarray(w, np.float16), torch.float16) for w in (*w_density, *w_rgb)]
        self._bits = dev(np.asarray(bitfield, np.uint8), torch.uint8)
        assert [tuple(w.shape) for w in self._w] == [(64, 32), (16, 64), (64, 32), (64, 64), (16, 64)]
        assert self._bits.numel() == CASCADES * GRID ** 3 // 8
        m = _lib.NerfModelStruct()
        m.grid, m.n_grid_entries = self._grid.data_ptr(), self._grid.shape[0]
        for i, w in enumerate(self._w):
            m.weights[i] = w.data_ptr()
        m.bitfield, m.aabb_scale = self._bits.data_ptr(), int(aabb_scale)
        self._h = C.c_void_p()
        _lib.check(self._lib.ptk_nerf_create(self._ctx, C.byref(m), C.byref(self._h)))
        self.aabb_scale, self.scale, self.offset = int(aabb_scale), float(scale), np.asarray(offset, np.float32)
        half = 0.5 * min(1 << (CASCADES - 1), self.aabb_scale)
        box = np.array([[0.5 - half] * 3, [0.5 + half] * 3], np.float32) if render_aabb is None else \
            np.asarray(render_aabb, np.float32)
        # attributes the reference sets / reads (ingp_utils.py:31-43, run_vis_on_poses.py:38-56)
        self.render_aabb = SimpleNamespace(min=box[0].copy(), max=box[1].copy())
        self.background_color = [255, 255, 255, 0.0]
        self.snap_to_pixel_centers = True
        self.fov_axis = 0
        self.fov = 50.625
        self.exposure = 0.0
        self.shall_train = False
        self.nerf = SimpleNamespace(sharpen=0.0, render_with_camera_distortion=True, rendering_min_transmittance=0.01)
        self.render_mode = RenderMode.Shade
        self._camera = np.eye(4, dtype=np.float32)[:3]

    @classmethod
    def from_snapshot_arrays(cls, params_binary, density_grid, aabb_scale: int, device, **kw):
        """params_binary / density_grid: the fp16 arrays of a `weights.msgpack` snapshot
        (instant-ngp/src/testbed.cu:2905-3001)."""
        wd, wc, grid = split_params(params_binary, aabb_scale)
        mc = 0
        while (1 << mc) < aabb_scale:
            mc += 1
        return cls(grid, wd, wc, occupancy_bitfield(density_grid, mc), aabb_scale, device, **kw)

    def __del__(self):
        try:
            self._lib.ptk_nerf_destroy(self._h)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # ---- pyngp surface --------------------------------------------------------------------------
    def set_nerf_camera_matrix(self, cam):
        """Testbed::set_nerf_camera_matrix (instant-ngp/src/testbed.cu:216-218) =
        NerfDataset::nerf_matrix_to_ngp (include/neural-graphics-primitives/nerf_loader.h:113-131)."""
        r = np.array(cam, np.float32)[:3, :4].copy()
        r[:, 1] *= -1
        r[:, 2] *= -1
        r[:, 3] = r[:, 3] * np.float32(self.scale) + self.offset
        self._camera = r[[1, 2, 0], :]

    def set_ngp_camera_matrix(self, cam):
        self._camera = np.array(cam, np.float32)[:3, :4].copy()

    def _view(self, width, height, spp) -> _lib.NerfView:
        if not self.snap_to_pixel_centers:
            raise NotImplementedError('only snap_to_pixel_centers=True (the PixTrack setting) is implemented')
        if self.exposure != 0.0:
            raise NotImplementedError('only exposure 0 (the PixTrack setting) is implemented')
        v = _lib.NerfView()
        for i, x in enumerate(self._camera.ravel()):
            v.camera[i] = float(x)
        for i in range(3):
            v.render_aabb_min[i] = float(self.render_aabb.min[i])
            v.render_aabb_max[i] = float(self.render_aabb.max[i])
        res = (width, height)[self.fov_axis]
        f32 = np.float32   # fov_to_focal_length(1, fov) * resolution[fov_axis] (common_device.cuh:470-472, testbed.cu:2509-2511)
        rel = f32(0.5) * f32(1) / f32(np.tan(f32(0.5) * f32(self.fov) * f32(math.pi) / f32(180)))
        v.focal = float(f32(rel * f32(res)))
        v.depth_scale = float(f32(1.0 / self.scale))
        v.min_transmittance = float(self.nerf.rendering_min_transmittance)
        bg = np.asarray(self.background_color, np.float32)
        lin = np.where(bg[:3] <= 0.04045, bg[:3] / np.float32(12.92),
                       np.power((bg[:3] + np.float32(0.055)) / np.float32(1.055), np.float32(2.4)))
        for i in range(3):
            v.background[i] = float(lin[i])
        v.background[3] = float(bg[3])
        v.width, v.height, v.spp = int(width), int(height), int(spp)
        v.depth_mode = 1 if self.render_mode == 'Depth' else 0
        return v

    def render_device(self, width: int, height: int, spp: int = 8, want_rgba: bool = True, want_u8: bool = False,
                      want_depth: bool = False, out_u8: Optional[torch.Tensor] = None):
        """Stream-ordered render; returns (rgba float32 [H,W,4] | None, u8 [H,W,3] | None, depth | None) CUDA
        tensors.  The uint8 image feeds FrameTracker.refresh_reference without leaving the device; `out_u8` is a
        caller-owned [H,W,3] uint8 buffer to render into (a static address lets the extractor replay its plan graph)."""
        v = self._view(width, height, spp)
        rgba = torch.empty((height, width, 4), dtype=torch.float32, device=self.device) if want_rgba else None
        if out_u8 is not None:
            assert out_u8.is_cuda and out_u8.dtype == torch.uint8 and out_u8.is_contiguous()
            assert tuple(out_u8.shape) == (height, width, 3)
            u8 = out_u8
        else:
            u8 = torch.empty((height, width, 3), dtype=torch.uint8, device=self.device) if want_u8 else None
        dep = torch.empty((height, width), dtype=torch.float32, device=self.device) if want_depth else None
        _lib.check(self._lib.ptk_nerf_render(self._h, C.byref(v), None if rgba is None else rgba.data_ptr(),
                                             None if u8 is None else u8.data_ptr(),
                                             None if dep is None else dep.data_ptr(),
                                             _lib.current_stream_ptr(self.device)))
        return rgba, u8, dep

    def network(self, pos01: torch.Tensor, dirs: torch.Tensor, want_features: bool = False):
        """The network alone (NerfNetwork::inference, nerf_network.h:101-136) for [n,3] unit-cube positions and [n,3] unit
        directions (CUDA fp32): -> raw (r, g, b, density) [n,4] fp32 holding the fp16 network outputs, and optionally the
        [n,32] hash-grid encoding.  Stream-ordered."""
        pos01 = pos01.to(self.device, torch.float32).contiguous()
        dirs = dirs.to(self.device, torch.float32).contiguous()
        n = pos01.shape[0]
        out = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        feats = torch.empty((n, 32), dtype=torch.float32, device=self.device) if want_features else None
        _lib.check(self._lib.ptk_nerf_eval(self._h, pos01.data_ptr(), dirs.data_ptr(), n, out.data_ptr(),
                                           None if feats is None else feats.data_ptr(), _lib.current_stream_ptr(self.device)))
        return (out, feats) if want_features else out

    def last_stats(self) -> dict:
        """Statistics of the last render (synchronises): network samples, warp steps, rays that reached the object."""
        out = (C.c_uint64 * 4)()
        _lib.check(self._lib.ptk_nerf_stats(self._h, out))
        return dict(samples=int(out[0]), warp_steps=int(out[1]), rays=int(out[2]), idle_lane_steps=int(out[3]))

    def render(self, width: int, height: int, spp: int = 8, linear: bool = True) -> np.ndarray:
        """pyngp Testbed.render -> float32 [H,W,4] on the host (python_api.cu:127-173)."""
        if not linear:
            raise NotImplementedError('PixTrack renders with linear=True; sRGB output is not implemented')
        return self.render_device(width, height, spp)[0].cpu().numpy()


def get_nerf_image(testbed: NerfTestbed, nerf_pose, camera, depth: bool = False, alpha_thresh: float = 0.0,
                   device_output: bool = False):
    """Same call as reference pixtrack/visualization/run_vis_on_poses.py:28-57: `camera` needs `.size` and
    `.f` (pixloc Camera).  Returns uint8 [H,W,3] (numpy, or a CUDA tensor with device_output=True)."""
    spp = 8
    width, height = (int(x) for x in camera.size)
    fl_x = float(camera.f[0])
    angle_x = math.atan(width / (fl_x * 2)) * 2
    testbed.fov = angle_x * 180 / np.pi
    testbed.set_nerf_camera_matrix(np.asarray(nerf_pose)[:3, :])
    if depth:
        testbed.render_mode = testbed.render_mode.Depth
    try:
        rgba, u8, _ = testbed.render_device(width, height, spp, want_rgba=alpha_thresh > 0.0, want_u8=True)
    finally:
        if depth:
            testbed.render_mode = testbed.render_mode.Shade
    if alpha_thresh > 0.0:
        u8[rgba[:, :, 3] < alpha_thresh] = 0
    return u8 if device_output else u8.cpu().numpy()
