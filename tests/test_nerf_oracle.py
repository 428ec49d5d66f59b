"""CPU: the NeRF-render oracle (oracle/nerf.py) against analytic cases and known answers.

The reference renderer (pyngp) is not runnable here and ships no test vectors.  Its host-callable header code
is, and so are the small marching / indexing functions once lifted out of their .cu / template headers: the jitter
sequence, colour transfer, focal length, camera-matrix conversion, ray generation, box intersection, step sizes,
cascade choice, occupancy lookup, empty-space stepping, Morton codes, hash and grid index, and the hash-grid / SH
encoding kernels, and the ray-init / first-advance / compositing / shade / accumulate / compaction / tonemap kernels
(run as host loops) of the oracle are pinned against them (bottom of this file, fixture
tests/golden/nerf_host.json made by tests/golden/gen/make_nerf_goldens.py).  The MLPs and the order in which render()
combines the pieces are checked against properties of the published algorithm instead ("parity partly pinned"):
hash-grid layout numbers, Morton codes, the (0,1)-sequence property of the Owen-scrambled Sobol jitter, occupancy
pooling, empty space, and a closed-form transmittance for a zero network.
"""
import numpy as np
import pytest

from oracle import nerf
import synthetic as syn

f32 = np.float32


def model(scene, **kw):
    bits = nerf.bitfield_from_density_grid(scene['density_grid'], scene['max_cascade'])
    return nerf.NerfModel(scene['aabb_scale'], scene['grid'], scene['w_density'], scene['w_rgb'], bits, **kw)


def test_hash_grid_layout_of_base_config():
    scales, ress, offs = nerf.grid_layout(1)
    # 16 levels from 16 to 2048 (b = exp(ln(2048/16)/15) = 1.3819), dense while res^3 <= 2^19
    assert ress[0] == 16 and ress[-1] == 2048 and len(ress) == 16
    assert np.all(np.diff(ress) > 0)
    assert offs[1] == 16 ** 3 and offs[2] - offs[1] == (23 ** 3 + 7) // 8 * 8
    assert np.all(np.diff(offs)[5:] == 1 << 19)                 # 81^3 > 2^19: hashed from level 5 on
    assert offs[-1] == syn.nerf_grid_size(1)
    assert nerf.grid_layout(4)[1][-1] == 8192                   # finest resolution scales with the box


def test_morton_and_bit_tricks():
    rng = np.random.default_rng(0)
    xyz = rng.integers(0, 128, (1000, 3)).astype(np.uint32)
    code = nerf.morton3d(xyz[:, 0], xyz[:, 1], xyz[:, 2])
    assert np.array_equal(nerf.morton3d_invert(code), xyz[:, 0])
    assert np.array_equal(nerf.morton3d_invert(code >> np.uint32(1)), xyz[:, 1])
    assert np.array_equal(nerf.morton3d_invert(code >> np.uint32(2)), xyz[:, 2])
    assert int(nerf.morton3d(np.uint32([1]), np.uint32([0]), np.uint32([0]))[0]) == 1
    assert int(nerf.morton3d(np.uint32([0]), np.uint32([1]), np.uint32([1]))[0]) == 6
    assert int(nerf.reverse_bits(np.uint64([1]))[0]) == 0x80000000


@pytest.mark.parametrize('seed', [0, 12345, 786433 * 77])
def test_jitter_is_a_scrambled_01_sequence(seed):
    """Owen scrambling preserves the net property of the Sobol/van der Corput sequence: the first 2^k
    points put exactly one point in each interval [j/2^k, (j+1)/2^k)."""
    for k in (3, 6):
        v = nerf.ld_random_val(np.arange(1 << k, dtype=np.uint64), np.full(1 << k, seed, np.uint64)).astype(np.float64)
        assert v.min() >= 0.0 and v.max() <= 1.0
        assert sorted(np.floor(v * (1 << k)).astype(int).tolist()) == list(range(1 << k))


def test_bitfield_threshold_and_pooling():
    from pixtrack_b200.nerf import occupancy_bitfield
    sc = syn.nerf_scene(1, aabb_scale=2)
    bits = nerf.bitfield_from_density_grid(sc['density_grid'], sc['max_cascade'])
    assert np.array_equal(bits, occupancy_bitfield(sc['density_grid'], sc['max_cascade']))   # host loader == oracle
    n = 128 ** 3
    lv = [np.unpackbits(bits[i * n // 8:(i + 1) * n // 8], bitorder='little') for i in range(8)]
    assert lv[0].sum() > 0 and lv[7].sum() > 0
    # a cell occupied at cascade l is covered by an occupied cell at cascade l+1
    idx = np.nonzero(lv[2])[0][:5000].astype(np.uint32)
    x, y, z = (nerf.morton3d_invert(idx >> np.uint32(s)) for s in (0, 1, 2))
    up = nerf.morton3d(x // 2 + 32, y // 2 + 32, z // 2 + 32)
    assert lv[3][up].all()
    # the ball of radius 0.28: centre occupied, corner free at cascade 0
    m = nerf.NerfModel(2, sc['grid'], sc['w_density'], sc['w_rgb'], bits)
    assert nerf.occupied(m, np.array([[0.5, 0.5, 0.5]], f32), np.array([0]))[0]
    assert not nerf.occupied(m, np.array([[0.05, 0.05, 0.05]], f32), np.array([0]))[0]


def test_split_params_round_trip():
    sc = syn.nerf_scene(3, 1)
    flat = np.concatenate([w.ravel() for w in (*sc['w_density'], *sc['w_rgb'])] + [sc['grid'].ravel()])
    wd, wc, grid = nerf.split_params(flat, 1)
    assert np.array_equal(wd[1], sc['w_density'][1]) and np.array_equal(wc[2], sc['w_rgb'][2])
    assert np.array_equal(grid, sc['grid'])
    from pixtrack_b200.nerf import split_params
    wd2, wc2, grid2 = split_params(flat, 1)
    assert np.array_equal(wd2[0], wd[0]) and np.array_equal(wc2[1], wc[1]) and np.array_equal(grid2, grid)


def test_hash_encode_interpolates_grid_values():
    """At a grid vertex of a dense level the encoding is that vertex's entry; a constant table gives the
    constant back (weights sum to one)."""
    sc = syn.nerf_scene(0, 1)
    m = model(sc)
    scales, ress, offs = m.layout
    v = np.array([[3, 5, 7]], np.int64)
    x = ((v.astype(f32) - f32(0.5)) / scales[0]).astype(f32) + f32(1e-6)      # pos = x*scale + 0.5 -> vertex (3,5,7)
    enc = nerf.hash_encode(m, x).astype(f32)
    idx = 3 + 5 * 16 + 7 * 256
    assert np.allclose(enc[0, :2], m.grid[idx].astype(f32), atol=2e-3)
    m.grid = np.full_like(m.grid, 0.25)
    enc = nerf.hash_encode(m, np.random.default_rng(0).random((64, 3)).astype(f32)).astype(f32)
    assert np.allclose(enc, 0.25, atol=2e-3)


def test_empty_occupancy_renders_nothing():
    sc = syn.nerf_scene(0, 1)
    sc['density_grid'] = np.full_like(sc['density_grid'], -1.0)
    m = model(sc)
    out = nerf.render(m, syn.nerf_look_at((0.5, -0.9, 0.6)), 24, 16, 50.0, spp=2)
    assert np.all(out['rgba'] == 0)
    img = nerf.get_nerf_image(m, np.eye(4)[:3], 24, 16, 30.0, spp=1)
    assert img.dtype == np.uint8 and img.shape == (16, 24, 3) and not img.any()


def test_zero_network_closed_form_transmittance():
    """All parameters zero -> sigma = exp(0) = 1 and colour = logistic(0) = 0.5 everywhere.  Along the
    optical axis through a fully occupied unit cube the n samples have equal dt, so
    alpha = 1 - exp(-n dt), rgb = srgb_to_linear(0.5 alpha); n follows from the jittered start."""
    sc = syn.nerf_scene(0, 1, zero_network=True)
    sc['density_grid'] = np.ones_like(sc['density_grid'])
    for w in (*sc['w_density'], *sc['w_rgb']):
        w[...] = 0
    m = model(sc)
    W, H, spp = 5, 5, 4
    cam = syn.nerf_look_at((0.5, -1.0, 0.5), target=(0.5, 0.5, 0.5))
    out = nerf.render(m, cam, W, H, 20.0, spp=spp)
    pix = (H // 2) * W + W // 2
    dt = float(nerf.STEPSIZE)
    exp_a, exp_c = [], []
    for s in range(spp):
        jit = float(nerf.ld_random_val(np.uint64([s]), np.uint64([pix * 786433]))[0])
        t0 = 1.0 + 1e-6 + jit * dt                      # box entry at distance 1 (>= near distance)
        n = int(np.floor((2.0 - t0) / dt)) + 1          # samples while inside [1, 2]
        a = 1.0 - np.exp(-n * dt)
        exp_a.append(a)
        exp_c.append(float(nerf.srgb_to_linear(np.array([0.5 * a], f32))[0]))
    got = out['rgba'][H // 2, W // 2]
    assert abs(got[3] - np.mean(exp_a)) < 2e-3, (got, np.mean(exp_a))
    assert np.allclose(got[:3], np.mean(exp_c), atol=2e-3)
    assert 0.6 < got[3] < 0.65                           # 1 - exp(-1) = 0.632


def test_opaque_scene_saturates_and_depth_mode_is_consistent():
    m = model(syn.nerf_scene(0, 1))
    cam = syn.nerf_look_at((0.5, -0.9, 0.6))
    out = nerf.render(m, cam, 20, 14, 50.0, spp=1)
    a = out['rgba'][..., 3]
    assert a.max() > 0.999 and a.min() == 0.0           # ball in the middle, empty corners
    assert np.all(out['rgba'][..., :3] <= 1.0) and np.all(out['rgba'] >= 0.0)
    dep = nerf.render(m, cam, 20, 14, 50.0, spp=1, depth_mode=True)['rgba']
    c = dep[7, 10]
    # depth mode composites (camera-forward distance / dataset scale) * weight: the camera is 1.40 from the
    # ball centre, the surface of the radius-0.28 ball about 1.12
    assert 1.05 / 0.33 < c[0] / c[3] < 1.25 / 0.33 and c[0] == c[1] == c[2]
    assert np.array_equal(dep[..., 3] > 0, a > 0)


# ------------------------------------------------------------------------------------------------------------
# Pinned pieces: tests/golden/nerf_host.json holds the outputs of the reference's own host-callable functions
# (instant-ngp headers compiled in place by oracle/build_ref.py, harness oracle/ngp_ref/ngp_host.cu).
# ------------------------------------------------------------------------------------------------------------
import json
import os

HOST = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nerf_host.json')))


def test_jitter_sequence_matches_the_reference_bit_for_bit():
    rec = np.array(HOST['ld_random_val'], dtype=np.float64)
    got = nerf.ld_random_val(rec[:, 0].astype(np.uint64), rec[:, 1].astype(np.uint64))
    assert np.array_equal(got.astype(f32), rec[:, 2].astype(f32))


def test_colour_transfer_and_focal_length_match_the_reference():
    rec = np.array(HOST['srgb'], dtype=np.float64).astype(f32)
    np.testing.assert_allclose(nerf.srgb_to_linear(rec[:, 0]), rec[:, 1], rtol=3e-7, atol=1e-9)       # powf: <= 2 ulp
    np.testing.assert_allclose(nerf.linear_to_srgb(rec[:, 0]), rec[:, 2], rtol=3e-7, atol=1e-9)
    for res, deg, focal in HOST['fov_to_focal']:
        assert abs(float(nerf.fov_to_focal(int(res), deg)) - focal) <= 2e-7 * focal


def test_camera_matrix_conversion_matches_the_reference():
    m = model(syn.nerf_scene(0, 1))
    got = nerf.nerf_matrix_to_ngp(m, np.array(HOST['nerf_matrix'], f32).reshape(3, 4))
    assert np.array_equal(got, np.array(HOST['ngp_matrix'], f32).reshape(3, 4))


def test_rays_and_box_intersection_match_the_reference():
    R = HOST['rays']
    cam = np.array(R['camera'], f32).reshape(3, 4)
    o, d = nerf.pixel_rays(cam, R['width'], R['height'], R['fov'])
    boxes = [np.array([[b[0]] * 3, [b[1]] * 3], f32) for b in R['boxes']]
    t1 = nerf.ray_box(boxes[0], o, d)
    t4 = nerf.ray_box(boxes[1], o, d)
    n_hit = 0
    for s in R['samples']:
        i = s['y'] * R['width'] + s['x']                       # snap_to_pixel_centers: the sample index does not matter
        assert np.array_equal(o[i], np.array(s['o'], f32))
        np.testing.assert_allclose(d[i], np.array(s['d'], f32), rtol=0, atol=2e-7)
        for got, key in ((t1, 't_box1'), (t4, 't_box4')):
            ref = np.array(s[key], f32)
            if ref[0] > 1e30:
                assert got[0][i] > 1e30 and got[1][i] > 1e30
            else:
                n_hit += key == 't_box1'
                np.testing.assert_allclose([got[0][i], got[1][i]], ref, rtol=2e-6, atol=2e-6)
        assert bool(nerf._contains(boxes[1], o[i:i + 1])[0]) == bool(s['in_box4'])
    assert n_hit >= 100


def test_marching_helpers_match_the_reference():
    """calc_dt, mip_from_pos / mip_from_dt, cascaded_grid_idx_at, density_grid_occupied_at, distance_to_next_voxel,
    advance_to_next_voxel, the warps, Morton codes and the constants: the reference's own definitions
    (testbed_nerf.cu, lifted at build time by oracle/build_ref.py) evaluated on the host."""
    c = HOST['constants']
    assert f32(c['near']) == nerf.NEAR and f32(c['stepsize']) == nerf.STEPSIZE and c['cascades'] == nerf.CASCADES
    assert f32(c['max_cone_stepsize']) == nerf.MAX_STEPSIZE
    rec = np.array(HOST['calc_dt'], f32)
    assert np.array_equal(nerf.calc_dt(rec[:, 0], rec[:, 1]), rec[:, 2])
    mo = np.array(HOST['morton'], dtype=np.int64)
    code = nerf.morton3d(mo[:, 0].astype(np.uint32), mo[:, 1].astype(np.uint32), mo[:, 2].astype(np.uint32))
    assert np.array_equal(code.astype(np.int64), mo[:, 3])
    assert np.array_equal(nerf.morton3d_invert(code >> np.uint32(1)).astype(np.int64), mo[:, 4])
    lg = np.array(HOST['logistic'], f32)
    np.testing.assert_allclose(f32(1) / (f32(1) + np.exp(-lg[:, 0])), lg[:, 1], rtol=3e-7)
    # warps: position (relative to the aabb), direction, dt
    box = np.array([[-1.5] * 3, [2.5] * 3], f32)
    span = nerf.STEPSIZE * f32(1 << (nerf.CASCADES - 1)) - nerf.STEPSIZE
    for w in HOST['warp']:
        p = np.array(w['p'], f32)
        np.testing.assert_allclose((p - box[0]) / (box[1] - box[0]), np.array(w['warped'], f32), rtol=0, atol=1e-7)
        np.testing.assert_allclose((p + f32(1)) * f32(0.5), np.array(w['warp_direction'], f32), rtol=0, atol=1e-7)
        wdt = (f32(w['dt']) - nerf.STEPSIZE) / span
        assert abs(wdt - w['warp_dt']) <= 1e-6 * max(1.0, abs(w['warp_dt']))
        assert abs((f32(w['warp_dt']) * span + nerf.STEPSIZE) - w['unwarp_dt']) <= 1e-7
    # the empty-space walk
    M = HOST['march']
    pos = np.array([m['pos'] for m in M], f32)
    d = np.array([m['dir'] for m in M], f32)
    t = np.array([m['t'] for m in M], f32)
    cone = np.array([m['cone'] for m in M], f32)
    with np.errstate(divide='ignore'):
        idir = (f32(1) / d).astype(f32)
    dt = nerf.calc_dt(t, cone)
    assert np.array_equal(dt, np.array([m['dt'] for m in M], f32))
    assert np.array_equal(nerf.mip_from_pos(pos), np.array([m['mip_from_pos'] for m in M]))
    mip = nerf.mip_from_dt(dt, pos)
    assert np.array_equal(mip, np.array([m['mip_from_dt'] for m in M]))
    assert len(set(mip.tolist())) >= 4
    assert np.array_equal(nerf.cascaded_grid_idx(pos, mip), np.array([m['grid_idx'] for m in M]))
    n_bytes = nerf.CASCADES * 128 ** 3 // 8
    bits = ((np.arange(n_bytes, dtype=np.uint64) * np.uint64(2654435761) & np.uint64(0xFFFFFFFF)) >> np.uint64(13)).astype(np.uint8)
    occ = nerf.occupied_bits(bits, pos, mip)
    assert np.array_equal(occ, np.array([m['occupied'] for m in M]) > 0) and 0.2 < occ.mean() < 0.8
    res = (128 >> mip).astype(np.int64)
    assert np.array_equal(res, np.array([m['res'] for m in M]))
    with np.errstate(invalid='ignore'):
        dist = nerf.distance_to_next_voxel(pos, d, idir, res)
    np.testing.assert_allclose(dist, np.array([m['dist'] for m in M], f32), rtol=1e-6, atol=1e-7)
    for c0 in (0.0, 1.0 / 256.0):                   # cone angle is one scalar per model in the oracle
        k = cone == f32(c0)
        with np.errstate(invalid='ignore'):
            adv = nerf.advance_to_next_voxel(t[k], f32(c0), pos[k], d[k], idir[k], res[k])
        np.testing.assert_allclose(adv, np.array([m['advance'] for m in M], f32)[k], rtol=2e-7, atol=0)


def test_hash_and_grid_index_match_tiny_cuda_nn():
    """fast_hash / grid_index of tiny-cuda-nn (encodings/grid.h:82-116, lifted at build time) on every level of the base
    configuration: dense levels, the first hashed level, and the finest one; corner vertex included."""
    rec = np.array(HOST['grid_index'], dtype=np.uint64)
    seen_dense = seen_hash = 0
    for res, size, x, y, z, idx2, h in rec:
        c = (np.array([x], np.uint64), np.array([y], np.uint64), np.array([z], np.uint64))
        assert int(nerf.fast_hash(c)[0]) == int(h)
        assert int(nerf.grid_index(c, int(res), int(size))[0]) * 2 == int(idx2)       # x N_FEATURES_PER_LEVEL, feature 0
        dense = int(res) ** 3 <= int(size)
        seen_dense += dense
        seen_hash += not dense
    assert seen_dense >= 12 and seen_hash >= 48
    # and the layout the oracle derives has these level sizes
    _, ress, offs = nerf.grid_layout(1)
    assert [int(r) for r in ress] == sorted({int(r[0]) for r in rec})
    assert [int(offs[i + 1] - offs[i]) for i in range(16)] == [int(rec[6 * i][1]) for i in range(16)]


def test_hash_grid_and_sh_encodings_match_the_tiny_cuda_nn_kernels():
    """kernel_grid (level scale, pos_fract, grid_index, trilinear interpolation accumulated in fp16) and kernel_sh of
    tiny-cuda-nn, lifted as host functions by oracle/build_ref.py and run on a 6.1 M-entry table with a reproducible
    pattern: the oracle's hash_encode / sh_encode must give the same fp16 values."""
    E = HOST['encoding']
    scales, ress, offs = nerf.grid_layout(1)
    lv = np.array(E['levels'], dtype=np.float64)
    np.testing.assert_allclose(scales, lv[:, 0].astype(f32), rtol=2e-7)
    assert np.array_equal(ress, lv[:, 1].astype(np.int64)) and np.array_equal(np.diff(offs), lv[:, 2].astype(np.int64))
    assert int(offs[-1]) == E['total_entries']
    i = np.arange(2 * E['total_entries'], dtype=np.uint64)
    pat = (((i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(16)) & np.uint64(0xFFFF)
    grid = (pat.astype(f32) / f32(65536) - f32(0.5)).astype(np.float16).reshape(-1, 2)
    sc = syn.nerf_scene(0, 1, zero_network=True)
    bits = nerf.bitfield_from_density_grid(sc['density_grid'], sc['max_cascade'])
    m = nerf.NerfModel(1, grid, sc['w_density'], sc['w_rgb'], bits)
    S = E['samples']
    pos = np.array([s['pos'] for s in S], f32)
    ref = np.array([s['enc'] for s in S], f32)
    # (a) with the oracle's own level scales: two of the 16 differ from the harness's in the last float bit (exp / log /
    # exp2 come from different math libraries -- the GPU's exp2f is a third one), which moves a few values by one fp16 ulp
    enc = nerf.hash_encode(m, pos).astype(f32)
    assert np.abs(enc - ref).max() <= 5e-4 and (enc == ref).mean() > 0.95
    # (b) with the harness's scales the interpolation must agree bit for bit: same vertices, same weights, same fp16
    # accumulation order
    m.layout = (lv[:, 0].astype(f32), m.layout[1], m.layout[2])
    enc = nerf.hash_encode(m, pos).astype(f32)
    assert np.array_equal(enc, ref)
    assert np.abs(ref).max() > 0.3 and ref[:, 20:].std() > 0.05            # hashed levels carry signal too
    d01 = np.array([s['dir01'] for s in S], f32)
    sh = nerf.sh_encode(d01).astype(f32)
    sh_ref = np.array([s['sh'] for s in S], f32)
    assert sh.shape == sh_ref.shape == (len(S), 16)
    assert np.abs(sh - sh_ref).max() <= 1e-3 and (sh == sh_ref).mean() > 0.97     # fp16 outputs, fma contraction differences


def test_compositing_matches_the_reference_kernel():
    """composite_kernel_nerf (lifted as a host function; __expf -> expf) over 24 rays x up to 6 samples, shade and
    depth mode, with partly composited and dead rays: final rgba, depth, max weight, termination flag and step
    count against oracle.composite_sample applied sample by sample."""
    for C in HOST['composite']:
        cam = np.array(C['camera'], f32).reshape(3, 4)
        aabb = np.array([[C['aabb'][0]] * 3, [C['aabb'][1]] * 3], f32)
        rays = C['rays']
        n = len(rays)
        rgba = np.array([r['rgba0'] for r in rays], f32)
        maxw, dep = np.zeros(n, f32), np.zeros(n, f32)
        alive = np.array([bool(r['alive']) for r in rays])
        steps_done = np.zeros(n, np.int64)
        terminated = np.zeros(n, bool)
        origin = np.broadcast_to(cam[:, 3], (n, 3)).astype(f32)
        for j in range(6):
            k = np.nonzero(alive & np.array([j < r['n_steps'] for r in rays]))[0]
            if k.size == 0:
                continue
            smp = np.array([rays[i]['samples'][j] for i in k], f32)
            done = nerf.composite_sample(rgba, maxw, dep, k, smp[:, 4:8], smp[:, 0:3], smp[:, 3], aabb, origin[k], cam,
                                         f32(C['depth_scale']), bool(C['depth_mode']), C['min_transmittance'])
            steps_done[k] = j + 1
            terminated[k[done]] = True
            alive[k[done]] = False
        ref = np.array(C['result'], dtype=np.float64)
        live = np.array([bool(r['alive']) for r in rays])
        np.testing.assert_allclose(rgba[live], ref[live, :4].astype(f32), rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(dep[live], ref[live, 4].astype(f32), rtol=2e-6, atol=2e-7)
        # alpha = 1 - exp(-sigma dt) cancels for thin samples: one ulp of exp (numpy vs libm) is 6e-8 absolute
        np.testing.assert_allclose(maxw[live], ref[live, 7].astype(f32), rtol=2e-6, atol=3e-7)
        # dead rays are untouched; a ray that broke out of the loop early is marked dead with j = index of the
        # terminating sample (payload.n_steps = j + current_step, :957-960)
        assert np.array_equal(rgba[~live], np.array([r['rgba0'] for r in rays], f32)[~live])
        early = terminated & (steps_done < 6 + 1)
        for i in np.nonzero(live)[0]:
            if terminated[i]:
                assert ref[i, 5] == 0 and int(ref[i, 6]) == steps_done[i] - 1
        assert terminated.sum() >= 4 and (live & ~terminated).sum() >= 4 and early.any()


def test_ray_start_matches_the_reference_kernels():
    """init_rays_with_payload_kernel_nerf + advance_pos_nerf (lifted as host functions) on a 16x10 view of the unit cube
    filled with the reproducible occupancy pattern, sample passes 0 and 1: start distance, liveness after the box test,
    and the jittered, empty-space-skipped first sample distance of oracle.first_advance."""
    R = HOST['ray_start']
    cam = np.array(R['camera'], f32).reshape(3, 4)
    o, d = nerf.pixel_rays(cam, R['width'], R['height'], R['fov'])
    with np.errstate(divide='ignore'):
        idir = (f32(1) / d).astype(f32)
    n_bytes = nerf.CASCADES * 128 ** 3 // 8
    bits = ((np.arange(n_bytes, dtype=np.uint64) * np.uint64(2654435761) & np.uint64(0xFFFFFFFF)) >> np.uint64(13)).astype(np.uint8)
    sc = syn.nerf_scene(0, 1, zero_network=True)
    m = nerf.NerfModel(1, sc['grid'], sc['w_density'], sc['w_rgb'], bits)
    assert m.cone_angle == 0 and np.array_equal(m.render_aabb, np.array([[0] * 3, [1] * 3], f32))
    tmin, _ = nerf.ray_box(m.render_aabb, o, d)
    for s, rec in enumerate(R['passes']):
        rec = np.array(rec, dtype=np.float64)
        t0, alive0, t, alive = nerf.first_advance(m, o, d, idir, tmin, s)
        assert np.array_equal(alive0, rec[:, 0] > 0) and 0.3 < alive0.mean() < 0.7
        np.testing.assert_allclose(t0[alive0], rec[alive0, 1].astype(f32), rtol=3e-7)
        assert np.array_equal(alive, rec[:, 2] > 0)
        np.testing.assert_allclose(t[alive], rec[alive, 3].astype(f32), rtol=5e-7)
        assert (t[alive] > t0[alive]).mean() > 0.9                  # the jitter / skipping moved them
    assert not np.array_equal(np.array(R['passes'][0])[:, 3], np.array(R['passes'][1])[:, 3])


def test_shade_and_accumulate_match_the_reference_kernels():
    """shade_kernel_nerf (scattered through payload.idx, sRGB -> linear, depth above alpha 0.2) and accumulate_kernel
    (running mean, linear colour space) over three sample passes."""
    accum = np.zeros((12, 4), f32)
    for s, P in enumerate(HOST['shade']['passes']):
        rec = np.array(P['rgba'], f32)
        frame, dbuf = nerf.shade(rec[:, :4], rec[:, 4])
        ref = np.array(P['frame'], f32)[::-1]                        # payload.idx = n - 1 - i in the harness
        np.testing.assert_allclose(frame, ref[:, :4], rtol=3e-7, atol=1e-9)
        assert np.array_equal(dbuf, ref[:, 4]) and (dbuf == 0).any() and (dbuf > 0).any()
        accum = nerf.accumulate(accum[::-1], frame, s)[::-1]          # the buffers are indexed by pixel = n - 1 - i
        np.testing.assert_allclose(accum, np.array(P['accumulated'], f32), rtol=5e-7, atol=1e-9)


def test_compaction_threshold_and_tonemap_match_the_reference_kernels():
    """compact_kernel_nerf (atomic counters replaced by host counters) and tonemap_kernel (surface write replaced by a
    host array), linear in / linear out, exposure 0, identity curve."""
    C = HOST['compaction']
    rays = np.array(C['rays'], dtype=np.float64)
    alive, alpha = rays[:, 0] > 0, rays[:, 1].astype(f32)
    assert sorted(C['still_alive']) == np.nonzero(alive)[0].tolist()
    assert sorted(C['final']) == np.nonzero(~alive & (alpha > nerf.COMPACTION_MIN_ALPHA))[0].tolist()
    assert (alpha[~alive] <= nerf.COMPACTION_MIN_ALPHA).sum() >= 3          # incl. one ray at exactly 0.001
    for T in HOST['tonemap']:
        acc = np.array(T['accumulated'], f32)
        out = nerf.tonemap(acc, T['background'])
        np.testing.assert_allclose(out, np.array(T['out'], f32), rtol=3e-7, atol=1e-9)
    assert np.array_equal(np.array(HOST['tonemap'][0]['out'], f32), np.array(HOST['tonemap'][0]['accumulated'], f32))   # alpha-0 background


def test_marching_kernel_matches_the_reference():
    """generate_next_nerf_network_inputs (lifted as a host function) continuing sample pass 1 of the ray-start fixture
    for up to 4 samples per ray: the network inputs (warped position, direction, step), the distance reached and which
    rays left the box, against oracle.next_sample applied four times."""
    R = HOST['ray_start']
    cam = np.array(R['camera'], f32).reshape(3, 4)
    o, d = nerf.pixel_rays(cam, R['width'], R['height'], R['fov'])
    with np.errstate(divide='ignore'):
        idir = (f32(1) / d).astype(f32)
    n_bytes = nerf.CASCADES * 128 ** 3 // 8
    bits = ((np.arange(n_bytes, dtype=np.uint64) * np.uint64(2654435761) & np.uint64(0xFFFFFFFF)) >> np.uint64(13)).astype(np.uint8)
    sc = syn.nerf_scene(0, 1, zero_network=True)
    m = nerf.NerfModel(1, sc['grid'], sc['w_density'], sc['w_rgb'], bits)
    tmin, _ = nerf.ray_box(m.render_aabb, o, d)
    _, _, t, alive = nerf.first_advance(m, o, d, idir, tmin, 1)
    M = R['march_from_pass_1']
    produced = np.zeros(len(M), np.int64)
    for j in range(4):
        t, alive, k, wpos, wdir, wdt = nerf.next_sample(m, o, d, idir, t, alive)
        for a, i in enumerate(k):
            ref = np.array(M[i]['samples'][j], f32)
            np.testing.assert_allclose(wpos[a], ref[0:3], rtol=0, atol=1e-6)       # one ulp of t (2.4e-7 at t = 2.5) times |d|
            assert abs(float(wdt[a]) - float(ref[3])) <= 1e-6
            np.testing.assert_allclose(wdir[a], ref[4:7], rtol=0, atol=1e-7)
        produced[k] += 1
    for i, r in enumerate(M):
        if r['alive']:
            # a ray that left the box mid-way keeps its payload alive with n_steps < 4 (the compositing kernel retires it)
            assert produced[i] == r['n_steps']
            if r['n_steps'] == 4:
                assert abs(float(t[i]) - r['t']) <= 5e-7 * r['t']
    assert (produced == 4).sum() >= 60 and ((produced > 0) & (produced < 4)).any()


def test_fp32_versus_fp16_mlp_accumulation_stays_inside_the_render_tolerance():
    """The one deliberate numerical difference from the reference: tiny-cuda-nn's fused MLP keeps fp16 accumulators in its
    wmma fragments, this package (oracle and kernel) accumulates in fp32 and rounds to fp16 once per layer.  With the fp16
    accumulation emulated (a running fp16 sum over the 16-wide K blocks) the rendered image moves by less than the
    tolerance the GPU tests use (4e-3 on float RGBA) and no uint8 value moves by more than one level."""
    sc = syn.nerf_scene(2, 1)
    m = model(sc)
    cam = syn.nerf_look_at((0.5, -0.9, 0.6))
    a = nerf.render(m, cam, 32, 22, 45.0, spp=2)['rgba']
    nerf.FP16_ACCUMULATE = True
    try:
        b = nerf.render(m, cam, 32, 22, 45.0, spp=2)['rgba']
    finally:
        nerf.FP16_ACCUMULATE = False
    d = np.abs(a - b)
    assert 0 < d.max() < 2e-3 and d.mean() < 1e-4
    ua, ub = (a[..., :3] * 255).astype(np.uint8).astype(int), (b[..., :3] * 255).astype(np.uint8).astype(int)
    assert np.abs(ua - ub).max() <= 1


def test_product_host_side_camera_code_matches_the_reference_too():
    """The renderer's HOST half lives in pixtrack_b200/nerf.py (camera-matrix conversion, focal length, background to
    linear): checked directly against the reference outputs of the fixture, without going through the oracle."""
    from types import SimpleNamespace
    from pixtrack_b200.nerf import NerfTestbed, RenderMode
    fake = SimpleNamespace(scale=0.33, offset=np.array([0.5, 0.5, 0.5], f32), _camera=None)
    NerfTestbed.set_nerf_camera_matrix(fake, np.array(HOST['nerf_matrix'], f32).reshape(3, 4))
    assert np.array_equal(fake._camera, np.array(HOST['ngp_matrix'], f32).reshape(3, 4))
    for res, deg, focal in HOST['fov_to_focal']:
        fake = SimpleNamespace(snap_to_pixel_centers=True, exposure=0.0, _camera=np.zeros((3, 4), f32),
                               render_aabb=SimpleNamespace(min=np.zeros(3), max=np.ones(3)), fov_axis=0, fov=deg, scale=0.33,
                               nerf=SimpleNamespace(rendering_min_transmittance=1e-7), background_color=[0.2, 0.5, 0.8, 1.0],
                               render_mode=RenderMode.Shade)
        v = NerfTestbed._view(fake, int(res), 7, 8)
        # the reference computes fov_to_focal_length(1, fov) * res; the fixture holds fov_to_focal_length(res, fov)
        assert abs(v.focal - focal) <= 3e-7 * focal
        assert abs(v.depth_scale - 1 / 0.33) < 1e-6 and v.depth_mode == 0 and (v.width, v.height, v.spp) == (int(res), 7, 8)
        srgb = {round(r[0], 6): r[1] for r in HOST['srgb']}
        lin = nerf.srgb_to_linear(np.array([0.2, 0.5, 0.8], f32))
        assert np.allclose([v.background[i] for i in range(3)], lin, rtol=1e-6) and v.background[3] == 1.0


# ------------------------------------------------------------------------------------------------------------
# The orchestration: the reference's kernels chained in the reference's order (render_to_cpu -> render_frame ->
# render_nerf -> NerfTracer::init_rays_from_camera / trace with its compaction rounds -> shade -> accumulate -> tonemap,
# oracle/ngp_ref/ngp_host.cu `render_chain`) with an analytic network, against oracle/nerf.py::render with the same one.
# ------------------------------------------------------------------------------------------------------------
def _analytic_network(wpos, wdir):
    """The stand-in network of the harness: plain float32 arithmetic (no transcendental functions, so C and numpy agree
    bit for bit), rounded to fp16 like the reference's network output."""
    x, y, z = wpos[:, 0].astype(f32), wpos[:, 1].astype(f32), wpos[:, 2].astype(f32)
    m = np.maximum(np.maximum(np.abs(x - f32(0.5)), np.abs(y - f32(0.5))), np.abs(z - f32(0.5))).astype(f32)
    raw = np.stack([f32(8) * x - f32(4), f32(8) * y - f32(4), f32(6) * wdir[:, 2].astype(f32) - f32(3),
                    f32(6) - f32(14) * m], 1).astype(f32)
    return raw.astype(np.float16)


def _pattern_model():
    """aabb_scale 1 model whose occupancy bitfield is the harness's reproducible pattern (byte i = (i * 2654435761) >> 13)."""
    i = np.arange(nerf.CASCADES * nerf.GRID ** 3 // 8, dtype=np.uint64)
    bits = (((i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(13)).astype(np.uint8)
    sc = syn.nerf_scene(0, 1, zero_network=True)
    return nerf.NerfModel(1, sc['grid'], sc['w_density'], sc['w_rgb'], bits)


@pytest.mark.parametrize('mode', [0, 1])
def test_render_orchestration_matches_the_reference_kernels_chained_in_the_reference_order(mode):
    R = HOST['render_chain'][mode]
    assert R['depth_mode'] == mode
    m = _pattern_model()
    cam = np.array(R['camera'], f32).reshape(3, 4)
    out = nerf.render(m, cam, R['width'], R['height'], R['fov'], spp=R['spp'], depth_mode=bool(mode),
                      min_transmittance=R['min_transmittance'], background=(255.0, 255.0, 255.0, 0.0),
                      network_fn=_analytic_network)
    want = np.array(R['rgba'], np.float64).reshape(R['height'], R['width'], 4)
    got = out['rgba'].astype(np.float64)
    assert (want[..., 3] > 0.5).sum() > 50 and (want[..., 3] == 0).sum() > 10        # object and background both present
    # expf / powf of the two math libraries differ by an ulp or two per sample, hundreds of samples per ray: 2e-5.  A ray
    # grazing the box can gain or lose ONE border sample when a position lands an ulp on the other side of a cell
    # boundary (weight ~ 7e-4 there): allowed on at most 2 pixels, and never more than 1e-3.
    close = np.isclose(got, want, rtol=2e-5, atol=2e-6)
    assert (~close).any(-1).sum() <= 2, np.argwhere(~close)
    np.testing.assert_allclose(got, want, rtol=3e-3, atol=1e-3)
    # the depth buffer of the last sample pass: the reference leaves MAX_DEPTH (1e10) where no ray was shaded
    wd = np.array(R['depth'], np.float64).reshape(R['height'], R['width'])
    shaded = wd < 1e9
    np.testing.assert_allclose(out['depth'].astype(np.float64)[shaded], wd[shaded], rtol=2e-5, atol=2e-6)
    assert not out['depth'][~shaded].any()


def test_snapshot_parameter_order_matches_nerf_network_set_params():
    """`params_binary` of a snapshot is the concatenation NerfNetwork::set_params walks (nerf_network.h:361-395, its body
    run in the harness over recording sub-modules): density MLP, rgb MLP, hash grid, (parameter-free) SH encoding."""
    rec = {r['module']: r for r in HOST['set_params']}
    assert [r['module'] for r in sorted(HOST['set_params'], key=lambda r: (r['offset'], r['n_params'] == 0))] == \
        ['density_network', 'rgb_network', 'pos_encoding', 'dir_encoding']
    n_grid = syn.nerf_grid_size(1)
    assert rec['pos_encoding']['n_params'] == 2 * n_grid
    total = rec['dir_encoding']['offset']
    params = (np.arange(total) % 2039).astype(np.float16)                  # every position identifiable (< 2048: exact in fp16)
    wd, wc, grid = nerf.split_params(params, 1)
    o = rec['density_network']['offset']
    assert np.array_equal(np.concatenate([w.ravel() for w in wd]), params[o:o + rec['density_network']['n_params']])
    o = rec['rgb_network']['offset']
    assert np.array_equal(np.concatenate([w.ravel() for w in wc]), params[o:o + rec['rgb_network']['n_params']])
    o = rec['pos_encoding']['offset']
    assert np.array_equal(grid.ravel(), params[o:o + 2 * n_grid])
    # the product's importer splits the same way (pixtrack_b200/nerf.py::split_params needs the built library for the
    # grid size only)
    from pixtrack_b200.nerf import split_params as product_split
    pd, pc, pg = product_split(params, 1)
    assert all(np.array_equal(a, b) for a, b in zip((*pd, *pc, pg), (*wd, *wc, grid)))
