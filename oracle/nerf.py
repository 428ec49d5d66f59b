"""CPU oracle for the instant-ngp reference-view render used by PixTrack.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy float32/float16, vectorised over rays.

PARITY PARTLY PINNED.  The reference renderer is the native `pyngp` module (instant-ngp + tiny-cuda-nn,
CUDA only): it cannot be built or run in this container (cmake project with GLFW/GLEW/Vulkan, no GPU),
no snapshot ships with the reference, and the reference holds no test vectors for it.  What CAN run here
is the reference's host-callable header code, compiled in place by oracle/build_ref.py
(oracle/ngp_ref/ngp_host.cu -> oracle/_ref/ngp_host); its outputs are stored in
tests/golden/nerf_host.json and pin, in tests/test_nerf_oracle.py:
    ld_random_val (bit-exact), srgb_to_linear / linear_to_srgb, fov_to_focal, nerf_matrix_to_ngp (bit-exact),
    pixel_rays (= pixel_to_ray + normalisation), ray_box / _contains (= BoundingBox::ray_intersect / contains),
    and -- from definitions that build_ref.py lifts out of testbed_nerf.cu / tcnn's grid.h at build time, widening
    __device__ to __host__ __device__ without touching the bodies -- the step-size constants, calc_dt, mip_from_pos,
    mip_from_dt, cascaded_grid_idx (bit-exact), occupied_bits (bit-exact), distance_to_next_voxel,
    advance_to_next_voxel, the position / direction / dt warps, morton3d, fast_hash and grid_index (bit-exact),
    hash_encode (= tcnn kernel_grid run as a host loop: bit-exact given the same level scales; the scales themselves
    agree to 2e-7, two of sixteen differ in the last bit between math libraries), sh_encode (= kernel_sh) and
    composite_sample (= the sample loop of composite_kernel_nerf with its activations, run as a host function),
    first_advance (= init_rays_with_payload_kernel_nerf + advance_pos_nerf), next_sample (= the marching kernel
    generate_next_nerf_network_inputs), shade, accumulate, the compaction threshold and tonemap (= shade_kernel_nerf,
    accumulate_kernel, compact_kernel_nerf, tonemap_kernel).
UNPINNED: the fused MLPs (wmma fragments, no host form) and the ORDER in which render() strings the pinned pieces
together (restated from render_nerf / NerfTracer, testbed_nerf.cu:2035-2330); checked against analytic cases
(empty occupancy -> nothing rendered, zero network -> closed-form transmittance).
One deliberate numerical difference: tiny-cuda-nn's fully fused MLP accumulates in fp16 inside
wmma fragments (fully_fused_mlp.cu:67-69); here, and in csrc/ptk_nerf.cu, products of fp16
operands are accumulated in fp32 and rounded to fp16 once per layer.  FP16_ACCUMULATE = True emulates the
reference's running fp16 sum; on the test scenes the rendered RGBA moves by at most 8e-4 (mean 1e-5), inside the
4e-3 the GPU tests allow (tests/test_nerf_oracle.py, last test).

Paths cited: `ngp/` = instant-ngp/, `tcnn/` = instant-ngp/dependencies/tiny-cuda-nn/.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

f32 = np.float32
GRID = 128                      # NERF_GRIDSIZE, ngp/include/neural-graphics-primitives/common.h
CASCADES = 8                    # ngp/src/testbed_nerf.cu:48
NEAR = f32(0.05)                # :46
SQRT3 = f32(1.73205080757)      # :50
STEPSIZE = f32(SQRT3 / f32(1024))                      # :51  MIN_CONE_STEPSIZE
MAX_STEPSIZE = f32(STEPSIZE * f32(1 << (CASCADES - 1)) * f32(1024) / f32(GRID))   # :54
MARCH_ITER = 10000              # :63
N_LEVELS, N_FEAT, LOG2_T, BASE_RES = 16, 2, 19, 16     # ngp/configs/nerf/base.json:23-29


# ------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------
def grid_layout(aabb_scale: int, desired_resolution: float = 2048.0):
    """Per-level scale / resolution / parameter offsets of the hash grid.
    per_level_scale: ngp/src/testbed.cu:2233-2244; offsets: tcnn/include/tiny-cuda-nn/encodings/grid.h:898-930."""
    pls = f32(math.exp(math.log(desired_resolution * aabb_scale / BASE_RES) / (N_LEVELS - 1)))
    log2_pls = f32(np.log2(pls))
    scales, ress, offs = [], [], [0]
    for lv in range(N_LEVELS):
        scale = f32(np.exp2(f32(lv) * log2_pls) * f32(BASE_RES) - f32(1.0))
        res = int(np.ceil(scale)) + 1
        n = res ** 3 if float(res) ** 3 <= 2 ** 31 - 1 else 2 ** 31 - 1
        n = (n + 7) // 8 * 8
        n = min(n, 1 << LOG2_T)
        scales.append(scale)
        ress.append(res)
        offs.append(offs[-1] + n)
    return np.array(scales, f32), np.array(ress, np.int64), np.array(offs, np.int64)


@dataclass
class NerfModel:
    """What a snapshot holds (ngp/src/testbed.cu:2905-3001), unpacked.  Weights are [out][in] fp16
    (tcnn/src/fully_fused_mlp.cu:86-92 loads them col-major with ld = width)."""
    aabb_scale: int
    grid: np.ndarray                 # fp16 [offsets[-1]][2]
    w_density: Tuple[np.ndarray, np.ndarray]            # [64,32], [16,64]
    w_rgb: Tuple[np.ndarray, np.ndarray, np.ndarray]    # [64,32], [64,64], [16,64] (rows 0..2 used)
    bitfield: np.ndarray             # uint8 [CASCADES * GRID^3 / 8], Morton order per cascade
    scale: float = 0.33              # dataset scale / offset: ngp/include/.../nerf_loader.h:28,84-87
    offset: Tuple[float, float, float] = (0.5, 0.5, 0.5)
    render_aabb: Optional[np.ndarray] = None            # [2,3]; default = the training box
    layout: tuple = field(default=None)

    def __post_init__(self):
        self.layout = grid_layout(self.aabb_scale)
        s = f32(0.5) * f32(min(1 << (CASCADES - 1), self.aabb_scale))        # testbed_nerf.cu:2581-2582
        self.aabb = np.array([[0.5 - s] * 3, [0.5 + s] * 3], f32)
        if self.render_aabb is None:
            self.render_aabb = self.aabb.copy()
        self.render_aabb = np.asarray(self.render_aabb, f32)
        self.cone_angle = f32(0.0) if self.aabb_scale <= 1 else f32(1.0 / 256.0)   # :2596


def split_params(params: np.ndarray, aabb_scale: int):
    """`snapshot.params_binary` (fp16) -> (w_density, w_rgb, grid) in NerfNetwork::set_params order
    (ngp/include/neural-graphics-primitives/nerf_network.h:361-395): density MLP, rgb MLP, position
    encoding (the direction encoding has no parameters)."""
    p = np.asarray(params, np.float16).ravel()
    shapes = [(64, 32), (16, 64), (64, 32), (64, 64), (16, 64)]
    out, o = [], 0
    for r, c in shapes:
        out.append(p[o:o + r * c].reshape(r, c))
        o += r * c
    n_grid = int(grid_layout(aabb_scale)[2][-1]) * N_FEAT
    grid = p[o:o + n_grid].reshape(-1, N_FEAT)
    assert grid.shape[0] * N_FEAT == n_grid, 'params_binary too short for this aabb_scale'
    return (out[0], out[1]), (out[2], out[3], out[4]), grid


# ------------------------------------------------------------------------------------------------
# occupancy bitfield                                   ngp/src/testbed_nerf.cu:555-604,2709-2724
# ------------------------------------------------------------------------------------------------
def _expand_bits(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3d(x, y, z):
    """tcnn/include/tiny-cuda-nn/common_device.h:349-354."""
    return _expand_bits(x) | (_expand_bits(y) << np.uint32(1)) | (_expand_bits(z) << np.uint32(2))


def morton3d_invert(x):
    x = x.astype(np.uint32) & np.uint32(0x49249249)
    x = (x | (x >> np.uint32(2))) & np.uint32(0xc30c30c3)
    x = (x | (x >> np.uint32(4))) & np.uint32(0x0f00f00f)
    x = (x | (x >> np.uint32(8))) & np.uint32(0xff0000ff)
    x = (x | (x >> np.uint32(16))) & np.uint32(0x0000ffff)
    return x


def bitfield_from_density_grid(density: np.ndarray, max_cascade: int) -> np.ndarray:
    """density: float [(max_cascade+1) * GRID^3] in Morton order per cascade (the snapshot's
    `density_grid_binary`).  grid_to_bitfield + bitfield_max_pool."""
    n = GRID ** 3
    d = np.asarray(density, f32)
    mean = f32(np.maximum(d[:n], 0).astype(np.float64).sum() / n)
    thresh = min(f32(0.01), mean)                                      # NERF_MIN_OPTICAL_THICKNESS
    bits = np.zeros(CASCADES * n // 8, np.uint8)
    occ = (d[:(max_cascade + 1) * n] > thresh).reshape(-1, 8)
    bits[:occ.shape[0]] = (occ * (1 << np.arange(8))).sum(1).astype(np.uint8)
    for lv in range(1, CASCADES):
        prev = bits[(lv - 1) * n // 8: lv * n // 8]
        i = np.arange(n // 64, dtype=np.uint32)
        pooled = ((prev.reshape(-1, 8) > 0) * (1 << np.arange(8))).sum(1).astype(np.uint8)
        x = morton3d_invert(i >> np.uint32(0)) + np.uint32(GRID // 8)
        y = morton3d_invert(i >> np.uint32(1)) + np.uint32(GRID // 8)
        z = morton3d_invert(i >> np.uint32(2)) + np.uint32(GRID // 8)
        dst = morton3d(x, y, z).astype(np.int64)
        np.bitwise_or.at(bits[lv * n // 8:(lv + 1) * n // 8], dst, pooled)
    return bits


# ------------------------------------------------------------------------------------------------
# low-discrepancy jitter                 ngp/include/neural-graphics-primitives/random_val.cuh:204-289
# ------------------------------------------------------------------------------------------------
def _u32(x):
    return np.asarray(x).astype(np.uint64) & np.uint64(0xFFFFFFFF)


def reverse_bits(x):
    x = _u32(x)
    x = ((x & np.uint64(0xaaaaaaaa)) >> np.uint64(1)) | ((x & np.uint64(0x55555555)) << np.uint64(1))
    x = ((x & np.uint64(0xcccccccc)) >> np.uint64(2)) | ((x & np.uint64(0x33333333)) << np.uint64(2))
    x = ((x & np.uint64(0xf0f0f0f0)) >> np.uint64(4)) | ((x & np.uint64(0x0f0f0f0f)) << np.uint64(4))
    x = ((x & np.uint64(0xff00ff00)) >> np.uint64(8)) | ((x & np.uint64(0x00ff00ff)) << np.uint64(8))
    return _u32((x >> np.uint64(16)) | (x << np.uint64(16)))


def laine_karras(x, seed):
    x = _u32(_u32(x) + _u32(seed))
    for c in (0x6c50b47c, 0xb82f1e52, 0xc7afe638, 0x8d22f6e6):
        x = x ^ _u32(x * np.uint64(c))
    return _u32(x)


def nested_uniform_scramble(x, seed):
    return reverse_bits(laine_karras(reverse_bits(x), seed))


def hash_combine(seed, v):
    seed = _u32(seed)
    return _u32(seed ^ _u32(np.uint64(v) + _u32(seed << np.uint64(6)) + (seed >> np.uint64(2))))


def ld_random_val(index, seed):
    """Owen-scrambled Sobol point, dimension 0 (random_val.cuh:284-288).  The first Sobol dimension
    is the base-2 radical inverse, i.e. sobol(i, 0) == reverse_bits(i) (direction numbers
    0x80000000 >> bit, random_val.cuh:160-167)."""
    seed = _u32(seed)
    idx = nested_uniform_scramble(index, seed)
    x = nested_uniform_scramble(reverse_bits(idx), hash_combine(seed, 0))
    return x.astype(np.uint32).astype(f32) * f32(1.0 / (1 << 32))


# ------------------------------------------------------------------------------------------------
# network                                              nerf_network.h:101-136
# ------------------------------------------------------------------------------------------------
def fast_hash(c) -> np.ndarray:
    """tcnn/include/tiny-cuda-nn/encodings/grid.h:82-98 for 3 dimensions (uint32 wrap-around arithmetic)."""
    primes = (np.uint64(1), np.uint64(2654435761), np.uint64(805459861))
    return _u32(c[0] * primes[0]) ^ _u32(c[1] * primes[1]) ^ _u32(c[2] * primes[2])


def grid_index(c, res: int, size: int) -> np.ndarray:
    """grid.h:100-116 (GridType::Hash): entry of grid vertex c = (cx, cy, cz) (uint64 arrays holding uint32 values)
    inside a level with `size` entries: dense strides while stride <= size, the hash when the level does not fit."""
    stride, index, d = 1, np.zeros(c[0].shape[0], np.uint64), 0
    while d < 3 and stride <= size:
        index = _u32(index + _u32(c[d] * np.uint64(stride)))
        stride = (stride * res) & 0xFFFFFFFF
        d += 1
    if size < stride:
        index = fast_hash(c)
    return index % np.uint64(size)


def hash_encode(m: NerfModel, x: np.ndarray) -> np.ndarray:
    """x [n,3] fp32 in the unit cube of the training box -> fp16 [n,32].
    tcnn/include/tiny-cuda-nn/encodings/grid.h:81-116 (index / hash), :139-275 (kernel_grid, linear
    interpolation accumulated in the parameter type, fp16), common_device.h:424-431 (pos_fract)."""
    scales, ress, offs = m.layout
    n = x.shape[0]
    out = np.zeros((n, N_LEVELS * N_FEAT), np.float16)
    for lv in range(N_LEVELS):
        scale, res = scales[lv], int(ress[lv])
        size = int(offs[lv + 1] - offs[lv])
        pos = x * scale + f32(0.5)
        fl = np.floor(pos)
        g = fl.astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
        w = (pos - fl).astype(f32)
        acc = np.zeros((n, N_FEAT), np.float16)
        for corner in range(8):
            wt = np.ones(n, f32)
            c = []
            for d in range(3):
                if (corner >> d) & 1:
                    wt = wt * w[:, d]
                    c.append(_u32(g[:, d] + np.uint64(1)))
                else:
                    wt = wt * (f32(1) - w[:, d])
                    c.append(g[:, d])
            index = grid_index(c, res, size).astype(np.int64) + offs[lv]
            val = m.grid[index].astype(f32)
            acc = (acc + (wt[:, None] * val).astype(np.float16)).astype(np.float16)
        out[:, lv * N_FEAT:(lv + 1) * N_FEAT] = acc
    return out


def sh_encode(d01: np.ndarray) -> np.ndarray:
    """Degree-4 real spherical harmonics of the direction packed to [0,1] (warp_direction), fp16 [n,16].
    tcnn/include/tiny-cuda-nn/encodings/spherical_harmonics.h:62-93."""
    x = d01[:, 0] * f32(2) - f32(1)
    y = d01[:, 1] * f32(2) - f32(1)
    z = d01[:, 2] * f32(2) - f32(1)
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = np.empty((d01.shape[0], 16), f32)
    o[:, 0] = f32(0.28209479177387814)
    o[:, 1] = f32(-0.48860251190291987) * y
    o[:, 2] = f32(0.48860251190291987) * z
    o[:, 3] = f32(-0.48860251190291987) * x
    o[:, 4] = f32(1.0925484305920792) * xy
    o[:, 5] = f32(-1.0925484305920792) * yz
    o[:, 6] = f32(0.94617469575755997) * z2 - f32(0.31539156525251999)
    o[:, 7] = f32(-1.0925484305920792) * xz
    o[:, 8] = f32(0.54627421529603959) * x2 - f32(0.54627421529603959) * y2
    o[:, 9] = f32(0.59004358992664352) * y * (f32(-3.0) * x2 + y2)
    o[:, 10] = f32(2.8906114426405538) * xy * z
    o[:, 11] = f32(0.45704579946446572) * y * (f32(1.0) - f32(5.0) * z2)
    o[:, 12] = f32(0.3731763325901154) * z * (f32(5.0) * z2 - f32(3.0))
    o[:, 13] = f32(0.45704579946446572) * x * (f32(1.0) - f32(5.0) * z2)
    o[:, 14] = f32(1.4453057213202769) * z * (x2 - y2)
    o[:, 15] = f32(0.59004358992664352) * x * (-x2 + f32(3.0) * y2)
    return o.astype(np.float16)


FP16_ACCUMULATE = False     # True: emulate the reference's fp16 wmma accumulators (only to measure the difference)


def _layer(x16: np.ndarray, w16: np.ndarray, relu: bool) -> np.ndarray:
    if FP16_ACCUMULATE:
        # fully_fused_mlp.cu:67-69,125-140: 16x16x16 wmma steps with __half accumulator fragments -- every K block
        # of 16 adds its (exact) partial dot product to an fp16 running sum
        x, w = x16.astype(f32), w16.astype(f32)
        acc = np.zeros((x.shape[0], w.shape[0]), np.float16)
        for k0 in range(0, x.shape[1], 16):
            acc = (acc.astype(f32) + x[:, k0:k0 + 16] @ w[:, k0:k0 + 16].T).astype(np.float16)
        y = acc.astype(f32)
    else:
        y = x16.astype(f32) @ w16.astype(f32).T       # fp16 operands, fp32 accumulate (see header)
    if relu:
        y = np.maximum(y, 0)
    return y.astype(np.float16)


def network(m: NerfModel, pos01: np.ndarray, dir01: np.ndarray, want_rgb: bool = True) -> np.ndarray:
    """-> fp16 [n,4]: raw r, g, b and raw density (before the activations)."""
    enc = hash_encode(m, pos01)
    h = _layer(enc, m.w_density[0], True)
    dens = _layer(h, m.w_density[1], False)                      # [n,16]; column 0 = density
    out = np.zeros((pos01.shape[0], 4), np.float16)
    out[:, 3] = dens[:, 0]
    if want_rgb:
        x = np.concatenate([dens, sh_encode(dir01)], 1)          # [n,32]
        x = _layer(x, m.w_rgb[0], True)
        x = _layer(x, m.w_rgb[1], True)
        out[:, :3] = _layer(x, m.w_rgb[2], False)[:, :3]
    return out


# ------------------------------------------------------------------------------------------------
# marching                                             ngp/src/testbed_nerf.cu:88-92,185-207,312-336,451-464
# ------------------------------------------------------------------------------------------------
def calc_dt(t, cone):
    return np.clip(t * cone, STEPSIZE, MAX_STEPSIZE).astype(f32)


def mip_from_pos(pos):
    mx = np.abs(pos - f32(0.5)).max(1)
    _, e = np.frexp(mx)
    return np.minimum(CASCADES - 1, np.maximum(0, e + 1)).astype(np.int64)


def mip_from_dt(dt, pos):
    mip = mip_from_pos(pos)
    d = dt * f32(2 * GRID)
    _, e = np.frexp(d)
    return np.where(d < 1, mip, np.minimum(CASCADES - 1, np.maximum(e, mip))).astype(np.int64)


def cascaded_grid_idx(pos, mip):
    """cascaded_grid_idx_at (:312-331): Morton index of the occupancy cell of `pos` in cascade `mip`."""
    scale = np.ldexp(f32(1), -mip).astype(f32)
    p = (pos - f32(0.5)) * scale[:, None] + f32(0.5)
    i = np.clip((p * f32(GRID)).astype(np.int32), 0, GRID - 1).astype(np.uint32)
    return morton3d(i[:, 0], i[:, 1], i[:, 2]).astype(np.int64)


def occupied_bits(bitfield, pos, mip):
    """density_grid_occupied_at (:333-336)."""
    idx = cascaded_grid_idx(pos, mip)
    byte = bitfield[idx // 8 + mip * (GRID ** 3 // 8)]
    return (byte >> (idx % 8).astype(np.uint8)) & 1 > 0


def occupied(m: NerfModel, pos, mip):
    return occupied_bits(m.bitfield, pos, mip)


def distance_to_next_voxel(pos, d, idir, res):
    p = pos * res[:, None].astype(f32)
    sg = np.copysign(f32(1), d)
    tt = (np.floor(p + f32(0.5) + f32(0.5) * sg) - p) * idir
    return np.maximum(np.fmin(np.fmin(tt[:, 0], tt[:, 1]), tt[:, 2]) / res.astype(f32), f32(0))   # device min() drops NaN


def advance_to_next_voxel(t, cone, pos, d, idir, res):
    """:195-207: regular stepping past the current empty cell: do { t += calc_dt(t) } while (t < t_target)."""
    target = (t + distance_to_next_voxel(pos, d, idir, res)).astype(f32)
    tk = t.astype(f32).copy()
    adv = np.ones(tk.shape[0], bool)
    while adv.any():
        tk[adv] = (tk[adv] + calc_dt(tk[adv], cone)).astype(f32)
        adv = adv & (tk < target)
    return tk


def _dot3(a, b):
    return ((a[:, 0] * b[0] + a[:, 1] * b[1]) + a[:, 2] * b[2]).astype(f32)


def _contains(box, p):
    return ((p >= box[0]) & (p <= box[1])).all(1)


def _skip_empty(m: NerfModel, o, d, idir, t, alive):
    """Common loop of advance_pos_nerf (:606-657) and generate_next_nerf_network_inputs (:717-740):
    advance t to the next sample position inside an occupied cell; rays leaving the render box die.
    Returns (t, alive, pos, dt)."""
    n = t.shape[0]
    pos = np.zeros((n, 3), f32)
    dt = np.zeros(n, f32)
    todo = alive.copy()
    while todo.any():
        k = np.nonzero(todo)[0]
        p = (o[k] + d[k] * t[k, None]).astype(f32)
        inside = _contains(m.render_aabb, p)
        alive[k[~inside]] = False
        todo[k[~inside]] = False
        k, p = k[inside], p[inside]
        if k.size == 0:
            break
        dtk = calc_dt(t[k], m.cone_angle)
        mip = mip_from_dt(dtk, p)
        occ = occupied(m, p, mip)
        pos[k[occ]], dt[k[occ]] = p[occ], dtk[occ]
        todo[k[occ]] = False
        k, p, mip = k[~occ], p[~occ], mip[~occ]
        if k.size == 0:
            continue
        res = (GRID >> mip).astype(np.int64)
        t[k] = advance_to_next_voxel(t[k], m.cone_angle, p, d[k], idir[k], res)
    return t, alive, pos, dt


def srgb_to_linear(x):
    """ngp/include/neural-graphics-primitives/common_device.cuh:31-37."""
    x = x.astype(f32)
    return np.where(x <= f32(0.04045), x / f32(12.92), np.power((np.maximum(x, 0) + f32(0.055)) / f32(1.055), f32(2.4))).astype(f32)


def fov_to_focal(res: int, degrees: float) -> np.float32:
    """common_device.cuh:470-472."""
    return f32(f32(0.5) * f32(res) / f32(np.tan(f32(0.5) * f32(degrees) * f32(math.pi) / f32(180))))


def nerf_matrix_to_ngp(m: NerfModel, mat: np.ndarray) -> np.ndarray:
    """nerf_loader.h:113-131 (not from_mitsuba)."""
    r = np.array(mat, f32)[:3, :4].copy()
    r[:, 1] *= -1
    r[:, 2] *= -1
    r[:, 3] = r[:, 3] * f32(m.scale) + np.asarray(m.offset, f32)
    return r[[1, 2, 0], :]


def pixel_rays(cam: np.ndarray, width: int, height: int, fov_deg: float, fov_axis: int = 0):
    """Origin and unit direction of the ray through every pixel centre, row-major: pixel_to_ray with
    snap_to_pixel_centers (the offset is fract(0.5 - v + v) = 0.5), screen centre 0.5, no parallax / aperture /
    distortion (common_device.cuh:260-307), then the normalisation of init_rays_with_payload_kernel_nerf
    (testbed_nerf.cu:1842-1849).  focal = calc_focal_length at zoom 1 (testbed.cu:2460-2462)."""
    cam = np.asarray(cam, f32)
    res = (width, height)
    focal = f32(fov_to_focal(1, fov_deg) * f32(res[fov_axis]))
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing='ij')
    px, py = xs.ravel().astype(f32), ys.ravel().astype(f32)
    n = px.size
    u = (px + f32(0.5)) / f32(width)
    v = (py + f32(0.5)) / f32(height)
    dcam = np.stack([(u - f32(0.5)) * f32(width) / focal, (v - f32(0.5)) * f32(height) / focal, np.ones(n, f32)], 1)
    d = ((dcam[:, 0:1] * cam[None, :, 0] + dcam[:, 1:2] * cam[None, :, 1]) + dcam[:, 2:3] * cam[None, :, 2]).astype(f32)
    d = (d / np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])[:, None]).astype(f32)
    o = np.broadcast_to(cam[:, 3], (n, 3)).astype(f32)
    return o, d


def ray_box(box, o: np.ndarray, d: np.ndarray):
    """BoundingBox::ray_intersect (bounding_box.cuh:163-208): (tmin, tmax) per ray, both FLT_MAX on a miss."""
    with np.errstate(divide='ignore', invalid='ignore'):
        t0 = ((box[0] - o) / d).astype(f32)
        t1 = ((box[1] - o) / d).astype(f32)
    lo, hi = np.minimum(t0, t1), np.maximum(t0, t1)
    tmin, tmax = lo[:, 0].copy(), hi[:, 0].copy()
    miss = np.zeros(o.shape[0], bool)
    for a in (1, 2):
        miss |= (tmin > hi[:, a]) | (lo[:, a] > tmax)
        tmin = np.where(lo[:, a] > tmin, lo[:, a], tmin)
        tmax = np.where(hi[:, a] < tmax, hi[:, a], tmax)
    big = np.finfo(f32).max
    return np.where(miss, big, tmin).astype(f32), np.where(miss, big, tmax).astype(f32)


def linear_to_srgb(x):
    """common_device.cuh:55-61."""
    x = x.astype(f32)
    return np.where(x < f32(0.0031308), f32(12.92) * x, f32(1.055) * np.power(np.maximum(x, 0), f32(0.41666)) - f32(0.055)).astype(f32)


def composite_sample(rgba, maxw, dep, k, out, wpos, wdt, aabb, origin, cam, depth_scale, depth_mode: bool,
                     min_transmittance) -> np.ndarray:
    """One iteration of composite_kernel_nerf's sample loop (testbed_nerf.cu:797-954) for the rays `k`: `out` [n,4] raw
    network outputs (rgb, density), `wpos` / `wdt` the warped position / step the network saw.  Updates rgba, maxw
    (payload.max_weight) and dep in place and returns which of the rays terminated (alpha above 1 - min_transmittance;
    their colour is then normalised by alpha).  Activations: logistic rgb, exponential density (network_to_rgb /
    network_to_density, :209-259)."""
    diag = aabb[1] - aabb[0]
    cam_fwd = cam[:, 2]
    upos = (aabb[0] + wpos * diag).astype(f32)                                      # unwarp_position
    udt = (wdt * (STEPSIZE * f32(1 << (CASCADES - 1)) - STEPSIZE) + STEPSIZE).astype(f32)   # unwarp_dt
    T = f32(1) - rgba[k, 3]
    alpha = f32(1) - np.exp(-np.exp(out[:, 3]) * udt).astype(f32)
    weight = (alpha * T).astype(f32)
    if depth_mode:
        val = _dot3(upos - origin, cam_fwd) * depth_scale
        rgb = np.repeat(val[:, None], 3, 1)
    else:
        rgb = (f32(1) / (f32(1) + np.exp(-out[:, :3]))).astype(f32)                # Logistic
    rgba[k, :3] += rgb * weight[:, None]
    rgba[k, 3] += weight
    better = weight > maxw[k]
    maxw[k[better]] = weight[better]
    dep[k[better]] = _dot3(upos[better] - cam[:, 3], cam_fwd)
    done = rgba[k, 3] > f32(1.0 - min_transmittance)
    kd = k[done]
    rgba[kd] = rgba[kd] / rgba[kd, 3:4]
    return done


def first_advance(m: NerfModel, o, d, idir, tmin, s: int):
    """Start of every ray for sample pass `s`: init_rays_with_payload_kernel_nerf (:1864-1889: start at the box entry
    or the near distance, + 1e-6; dead when that point is outside the render box) followed by advance_pos_nerf
    (:606-657: jitter by a fraction of the step from the Sobol sequence seeded by the pixel, then skip empty space).
    Returns (t_init, alive_init, t, alive)."""
    n = o.shape[0]
    t0 = (np.maximum(tmin, NEAR) + f32(1e-6)).astype(f32)
    start = (o + d * t0[:, None]).astype(f32)
    alive0 = _contains(m.render_aabb, start)
    dt0 = calc_dt(t0, m.cone_angle)
    pix = np.arange(n, dtype=np.uint64)
    t = (t0 + ld_random_val(np.full(n, s, np.uint64), pix * np.uint64(786433)) * dt0).astype(f32)
    t, alive, _, _ = _skip_empty(m, o, d, idir, t, alive0.copy())
    return t0, alive0, t, alive


def next_sample(m: NerfModel, o, d, idir, t, alive):
    """One iteration of generate_next_nerf_network_inputs' sample loop (:693-752) for all live rays: skip empty space,
    emit the network input of the sample found (warped position / direction / step) and step past it.  Rays that
    leave the render box die.  Returns (t, alive, k, wpos, wdir, wdt) with k the indices of the rays that produced a
    sample."""
    t, alive, pos, dt = _skip_empty(m, o, d, idir, t, alive)
    k = np.nonzero(alive)[0]
    diag = m.aabb[1] - m.aabb[0]
    wpos = ((pos[k] - m.aabb[0]) / diag).astype(f32)                       # warp_position
    wdt = ((dt[k] - STEPSIZE) / (STEPSIZE * f32(1 << (CASCADES - 1)) - STEPSIZE)).astype(f32)   # warp_dt
    wdir = ((d[k] + f32(1)) * f32(0.5)).astype(f32)                        # warp_direction
    t[k] = (t[k] + dt[k]).astype(f32)
    return t, alive, k, wpos, wdir, wdt


def shade(rgba, dep, depth_mode: bool = False):
    """shade_kernel_nerf (:1721-1754) into a cleared frame buffer: colours to linear (Shade mode), depth kept where
    alpha > 0.2.  -> (frame [n,4], depth [n])."""
    frame = rgba.astype(f32).copy()
    if not depth_mode:
        frame[:, :3] = srgb_to_linear(rgba[:, :3])
    return frame, np.where(rgba[:, 3] > f32(0.2), dep, f32(0)).astype(f32)


def accumulate(accum, frame, s: int):
    """accumulate_kernel, linear colour space (render_buffer.cu:236-271): running mean over the sample passes."""
    return ((accum * f32(s) + frame) / f32(s + 1)).astype(f32)


COMPACTION_MIN_ALPHA = f32(0.001)      # compact_kernel_nerf (:1773): finished rays below this never reach the frame


def tonemap(accum, background=(1.0, 1.0, 1.0, 0.0)):
    """tonemap_kernel (render_buffer.cu:542-569) for linear in / linear out, exposure 0, identity curve, no clamp: the
    background (given in sRGB) is blended in behind, weighted by (1 - alpha) * background alpha."""
    bg = np.asarray(background, f32)
    w = (f32(1) - accum[:, 3]) * bg[3]
    outp = accum.astype(f32).copy()
    outp[:, :3] += srgb_to_linear(bg[:3])[None] * w[:, None]
    outp[:, 3] += w
    return outp


def render(m: NerfModel, camera_matrix: np.ndarray, width: int, height: int, fov_deg: float, spp: int = 8,
           depth_mode: bool = False, min_transmittance: float = 1e-7, fov_axis: int = 0,
           background=(1.0, 1.0, 1.0, 0.0), network_fn=None) -> Dict[str, np.ndarray]:
    """Testbed.render(width, height, spp, linear=True) with the settings of
    pixtrack/utils/ingp_utils.py:22-44 (snap_to_pixel_centers, min transmittance 1e-7, fov_axis 0,
    exposure 0, background alpha 0 by default) -> dict(rgba float32 [H,W,4], depth [H,W]).
    camera_matrix: 3x4 in NGP convention (nerf_matrix_to_ngp applied).
    python_api.cu:127-173 -> testbed.cu:2591-2749 (render_frame) -> testbed_nerf.cu:2228-2330
    (render_nerf) -> :1781-1890 (ray init), :606-657 (first advance), :2035-2146 (trace),
    :693-752 (samples), :754-955 (composite), :1721-1754 (shade); render_buffer.cu:236-275
    (accumulate), :542-570 (tonemap, linear, identity curve).
    network_fn(wpos, wdir) -> fp16 [n,4] replaces the hash grid + MLPs (the orchestration is pinned against the
    reference's kernels chained in the reference's order with an analytic network, tests/test_nerf_oracle.py)."""
    cam = np.asarray(camera_matrix, f32)
    o, d = pixel_rays(cam, width, height, fov_deg, fov_axis)
    n = o.shape[0]
    cam_fwd = cam[:, 2]
    depth_scale = f32(1.0 / m.scale)
    with np.errstate(divide='ignore', invalid='ignore'):
        idir = (f32(1) / d).astype(f32)
    tmin, _ = ray_box(m.render_aabb, o, d)
    accum = np.zeros((n, 4), f32)
    depth_out = np.zeros(n, f32)
    pix = np.arange(n, dtype=np.uint64)
    for s in range(spp):
        _, _, t, alive = first_advance(m, o, d, idir, tmin, s)
        rgba = np.zeros((n, 4), f32)
        dep = np.zeros(n, f32)
        maxw = np.zeros(n, f32)
        steps = 1
        while alive.any() and steps < MARCH_ITER:
            t, alive, k, wpos, wdir, wdt = next_sample(m, o, d, idir, t, alive)
            if k.size == 0:
                break
            out = (network(m, wpos, wdir, want_rgb=not depth_mode) if network_fn is None
                   else network_fn(wpos, wdir)).astype(f32)
            done = composite_sample(rgba, maxw, dep, k, out, wpos, wdt, m.aabb, o[k], cam, depth_scale, depth_mode,
                                    min_transmittance)
            alive[k[done]] = False
            steps += 1
        # compaction keeps finished rays only when alpha > 0.001 (:1771)
        hit = rgba[:, 3] > COMPACTION_MIN_ALPHA
        frame, dbuf = shade(np.where(hit[:, None], rgba, f32(0)), np.where(hit, dep, f32(0)), depth_mode)
        accum = accumulate(accum, frame, s)
        depth_out = dbuf
    outp = tonemap(accum, background)
    return dict(rgba=outp.reshape(height, width, 4), depth=depth_out.reshape(height, width))


def get_nerf_image(m: NerfModel, nerf_pose: np.ndarray, width: int, height: int, fl_x: float, depth: bool = False,
                   spp: int = 8) -> np.ndarray:
    """pixtrack/visualization/run_vis_on_poses.py:28-57 -> uint8 [H,W,3]."""
    angle_x = math.atan(width / (fl_x * 2)) * 2
    out = render(m, nerf_matrix_to_ngp(m, np.asarray(nerf_pose)[:3, :]), width, height, angle_x * 180 / np.pi, spp,
                 depth_mode=depth)
    img = out['rgba'][:, :, :3] * f32(255.0)
    return img.astype(np.uint8)
