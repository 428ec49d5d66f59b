"""GPU: every tcgen05 convolution kernel path the 1024x576 / 1008x756 extractor plans dispatch to, in isolation,
against an fp32 convolution of the SAME fp16-rounded operands (so the only differences are the fp32 summation order
and the fp16 rounding of the output).

Dispatch (pixtrack_b200/csrc/ptk_conv.cu, conv_dispatch), in this order: the 16x16 CTA-pair kernel conv_halo2_kernel<N>
(N = 32 / 64 / 128; weights streamed or resident) when the tile pairs fill >= 80 % of their last wave of 74 SM pairs; the
narrow-map pair kernel conv_row64_kernel (maps 48..64 pixels wide with >= 44 pair tiles); the row-tile pair kernel
conv_row2_kernel<R, N> (long K, >= 60 % fill); the single-CTA halo kernel conv_halo_kernel<N> when its 16x16 tiles fill
>= 80 % of their last wave of 148 SMs (with the default switches: layers whose weights it keeps resident and the pair
kernel does not take, e.g. 64 -> 64 at full resolution); everything else conv_tc_kernel (per-tap boxes, K split).  The
shapes below are chosen to land on each of them, with ragged right / bottom tiles, two inputs, the fused 2x2 max pool
on odd sizes (the 1008x756 plan pools 189 -> 94), and the switches that force the non-default choices run in child
processes (the library reads them once).
"""
import pytest
import torch
import torch.nn.functional as tF

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
D = 'cuda:0'


def _ref(x, w, b, relu, x1=None, hw=None):
    H, W = hw if hw else x.shape[:2]
    xs = [x.float().permute(2, 0, 1)[None][:, :, :H, :W]]
    if x1 is not None:
        xs.append(x1.float().permute(2, 0, 1)[None][:, :, :H, :W])
    y = tF.conv2d(torch.cat(xs, 1), w.float(), b, padding=1)
    return tF.relu(y) if relu else y


def _case(cin, cout, H, W, seed, cin1=0, grow=(0, 0)):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(H, W, cin, generator=g).half().to(D)
    x1 = torch.randn(H + grow[0], W + grow[1], cin1, generator=g).half().to(D) if cin1 else None
    w = (torch.randn(cout, cin + cin1, 3, 3, generator=g) / (3 * (cin + cin1) ** 0.5)).half().to(D)
    b = torch.randn(cout, generator=g).to(D)
    return x, x1, w, b


def _check(y, ref_nchw, what):
    ref = ref_nchw[0].permute(1, 2, 0)
    err = (y.float() - ref).abs().max().item()
    assert err < 4e-3 * max(1.0, ref.abs().max().item()), (what, err)


# (cin, cout, H, W): 11 x 13 = 143 tiles of 16x16 (x C_out / 128 groups) -> halo kernel; ragged edges on both axes
HALO_SHAPES = [
    (64, 32, 170, 200),      # conv_halo_kernel<32>
    (64, 64, 170, 200),      # conv_halo_kernel<64>, weights resident in shared memory
    (128, 128, 170, 200),    # conv_halo_kernel<128>, weights resident
    (128, 256, 170, 200),    # conv_halo_kernel<128>, two C_out groups (286 tiles = 2 waves), weights streamed
    (256, 256, 171, 203),    # conv_halo_kernel<128>, 4 K chunks, odd sizes
    (192, 64, 172, 198),     # 3 K chunks
]


@pytest.mark.parametrize('cin,cout,H,W', HALO_SHAPES)
@pytest.mark.parametrize('relu', [True, False])
def test_halo_kernels_single_input(cin, cout, H, W, relu):
    from pixtrack_b200.extractor import conv_f16, pack_conv3x3
    x, _, w, b = _case(cin, cout, H, W, seed=cin * 11 + cout)
    y = conv_f16(x, pack_conv3x3(w), b, relu=relu)
    torch.cuda.synchronize()
    _check(y, _ref(x, w, b, relu), (cin, cout, H, W))


@pytest.mark.parametrize('cin,cout,H,W', [(64, 64, 170, 200), (128, 128, 170, 200), (128, 256, 171, 203),
                                          (64, 64, 189, 252),      # the 1008x756 plan's odd 189 -> 94 pool
                                          (64, 128, 40, 56)])       # small map: conv_tc_kernel 16x8 tiles + pool
def test_fused_max_pool(cin, cout, H, W):
    from pixtrack_b200.extractor import conv_f16, pack_conv3x3
    x, _, w, b = _case(cin, cout, H, W, seed=cin * 13 + cout + H)
    y, p = conv_f16(x, pack_conv3x3(w), b, relu=True, pool=True)
    torch.cuda.synchronize()
    ref = _ref(x, w, b, True)
    _check(y, ref, 'conv')
    # the pool is taken of the fp16-rounded outputs (max commutes with rounding): exact against pooling `y`
    want = tF.max_pool2d(y.float().permute(2, 0, 1)[None], 2, 2)[0].permute(1, 2, 0)
    assert tuple(p.shape) == (H // 2, W // 2, cout)
    assert torch.equal(p.float(), want)


@pytest.mark.parametrize('cin0,cin1,cout,H,W,grow', [
    (64, 64, 32, 170, 200, (0, 0)),       # last decoder block: halo<32>, upsampled + skip
    (64, 128, 64, 170, 200, (1, 1)),      # decoder block 2: halo<64>, skip one row / column larger (cropped)
    (64, 256, 64, 171, 203, (1, 0)),      # decoder block 1 shape class on a big map
    (128, 128, 128, 170, 200, (0, 3)),    # halo<128> with a second input
])
def test_halo_kernels_two_inputs_with_crop(cin0, cin1, cout, H, W, grow):
    from pixtrack_b200.extractor import conv_f16, pack_conv3x3
    x, x1, w, b = _case(cin0, cout, H, W, seed=cin0 + cin1 * 3 + cout, cin1=cin1, grow=grow)
    y = conv_f16(x, pack_conv3x3(w), b, relu=True, x1=x1, out_hw=(H, W))
    torch.cuda.synchronize()
    _check(y, _ref(x, w, b, True, x1=x1, hw=(H, W)), (cin0, cin1, cout))


@pytest.mark.parametrize('cin0,cin1,cout,H,W,pool', [
    (256, 0, 256, 144, 256, False),      # the 256-channel block of the 1024x576 plan: 72 pairs x 2 C_out groups
    (256, 0, 256, 144, 256, True),       # its last layer writes the fused pool
    (256, 0, 256, 189, 252, True),       # the same block of the 1008x756 plan: odd height, pooled to 94 x 126
    (128, 128, 128, 150, 250, False),    # two inputs, ragged pair columns (16 tile columns -> 8 pairs, last one half empty)
    (64, 128, 64, 288, 512, False),      # decoder block 2 of the 1024x576 plan: conv_halo2_kernel<64>, 3 chunks over two inputs
    (64, 256, 64, 144, 256, False),      # decoder block 1: 5 chunks
])
def test_cta_pair_kernel_default_dispatch(cin0, cin1, cout, H, W, pool):
    """3x3 layers whose weights are streamed (not resident in shared memory) and whose tile pairs fill >= 80 % of their last
    wave of 74 SM pairs run on conv_halo2_kernel<N>: tcgen05 cta_group::2, M = 256 x N = 128 or 64, two (or more)
    accumulator sets in TMEM."""
    from pixtrack_b200.extractor import conv_f16, pack_conv3x3
    x, x1, w, b = _case(cin0, cout, H, W, seed=cin0 + cout + H + pool, cin1=cin1)
    out = conv_f16(x, pack_conv3x3(w), b, relu=True, x1=x1, out_hw=(H, W), pool=pool)
    torch.cuda.synchronize()
    y = out[0] if pool else out
    _check(y, _ref(x, w, b, True, x1=x1, hw=(H, W)), (cin0, cin1, cout, H, W))
    if pool:
        want = tF.max_pool2d(y.float().permute(2, 0, 1)[None], 2, 2)[0].permute(1, 2, 0)
        assert torch.equal(out[1].float(), want)
    again = conv_f16(x, pack_conv3x3(w), b, relu=True, x1=x1, out_hw=(H, W))
    torch.cuda.synchronize()
    assert torch.equal(again, y)


def _row_pair_checks(shapes):
    from pixtrack_b200.extractor import conv_f16, pack_conv3x3
    for cin0, cin1, cout, H, W, pool in shapes:
        x, x1, w, b = _case(cin0, cout, H, W, seed=cin0 + cin1 + cout + H + W, cin1=cin1)
        for relu in (True, False):
            out = conv_f16(x, pack_conv3x3(w), b, relu=relu, x1=x1, out_hw=(H, W), pool=pool)
            torch.cuda.synchronize()
            y = out[0] if pool else out
            _check(y, _ref(x, w, b, relu, x1=x1, hw=(H, W)), (cin0, cin1, cout, H, W, relu))
            if pool:
                want = tF.max_pool2d(y.float().permute(2, 0, 1)[None], 2, 2)[0].permute(1, 2, 0)
                assert tuple(out[1].shape) == (H // 2, W // 2, cout)
                assert torch.equal(out[1].float(), want)
        again = conv_f16(x, pack_conv3x3(w), b, relu=False, x1=x1, out_hw=(H, W))
        torch.cuda.synchronize()
        assert torch.equal(again, y)
    print('row pair ok')


ROW_DEFAULT_SHAPES = [
    (256, 0, 512, 72, 128, False),      # 1/8-scale block of the 1024x576 plan: conv_row2_kernel<2>, 72 pair tiles = one wave
    (512, 0, 512, 72, 128, True),       # its last layer, fused pool (rows pair up inside a CTA)
    (512, 0, 512, 94, 126, False),      # the 1008x756 plan: conv_row2_kernel<3>, 64 pair tiles, ragged right and bottom edges
    (256, 256, 512, 94, 126, False),    # two inputs
    (512, 512, 64, 72, 128, False),     # first decoder convolution (N = 64, R = 1: 36 single-row pair tiles, 16 chunks)
    (512, 512, 64, 94, 126, False),     # ... of the 1008x756 plan
    (64, 256, 64, 188, 252, False),     # second decoder convolution of the 1008x756 plan (N = 64, R = 3, two column tiles)
]
ROW_FORCED_SHAPES = [
    (256, 0, 64, 38, 200, True),        # N = 64 with the fused pool (R = 2), two column tiles
    (256, 0, 128, 37, 200, True),       # two column tiles (the second 72 pixels wide), odd height, pooled
    (256, 0, 128, 300, 256, False),     # 150 pair tiles on 74 SM pairs: the persistent loop and both accumulator sets
    (192, 64, 256, 41, 130, False),     # a 2-pixel-wide second column tile, two inputs, two C_out groups
]


def test_row_tile_pair_kernel_default_dispatch():
    """Long-K layers on maps about 128 pixels wide (the 512-channel 1/8-scale block) run on conv_row2_kernel<R>: an MMA's
    128 rows are 128 consecutive pixels of an image row, a CTA pair owns 2R rows, tcgen05 cta_group::2."""
    _row_pair_checks(ROW_DEFAULT_SHAPES)


def test_row_tile_pair_kernel_forced_shapes():
    """The same kernel on shapes it is not chosen for by default (PTK_CONV_ROW=2: whenever legal): several column tiles,
    more tiles than SM pairs, ragged edges."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PTK_CONV_ROW='2', PTK_CONV_PAIR='0')
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, '-c', 'import sys; sys.path[:0] = [%r, %r]; import test_conv_paths_gpu as t; '
                        't._row_pair_checks(t.ROW_FORCED_SHAPES)' % (here, os.path.dirname(here))], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'row pair ok' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


ROW64_DEFAULT_SHAPES = [
    (512, 0, 512, 47, 63, False),       # 1/16-scale block of the 1008x756 plan: 48 pair tiles, ragged last column and rows
    (256, 256, 512, 45, 60, False),     # two inputs
]
ROW64_FORCED_SHAPES = [
    (512, 0, 512, 36, 64, False),       # 1/16-scale block of the 1024x576 plan (by default on the per-tap kernel, which is faster there)
    (128, 0, 128, 23, 100, False),      # two column tiles (the second 36 pixels wide), 2 chunks
    (256, 0, 128, 610, 64, False),      # 153 pair tiles on 74 SM pairs: the persistent loop and both accumulator sets
    (64, 64, 256, 9, 17, False),        # tiny map, two inputs of one chunk each, two C_out groups
]


def test_narrow_map_pair_kernel_default_dispatch():
    """Long-K layers on maps 48..64 pixels wide (the 512-channel 1/16-scale block) run on conv_row64_kernel: the halo is
    staged as three column-shifted copies so that two image rows form one 128-row MMA operand; CTA pairs, cta_group::2."""
    _row_pair_checks(ROW64_DEFAULT_SHAPES)


def test_narrow_map_pair_kernel_forced_shapes():
    import os
    import subprocess
    import sys
    env = dict(os.environ, PTK_CONV_ROW64='2', PTK_CONV_PAIR='0', PTK_CONV_ROW='0')
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, '-c', 'import sys; sys.path[:0] = [%r, %r]; import test_conv_paths_gpu as t; '
                        't._row_pair_checks(t.ROW64_FORCED_SHAPES)' % (here, os.path.dirname(here))], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'row pair ok' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


SPLIT_SHAPES = [
    (512, 0, 512, 36, 64, False),       # 1/16-scale block of the 1024x576 plan: halo<128, SPLIT 2>, 48 tiles x 2 CTAs
    (512, 0, 512, 47, 63, False),       # ... of the 1008x756 plan (odd sizes, ragged tiles)
    (512, 512, 64, 72, 128, False),     # first decoder convolution: two inputs, 16 chunks, halo<64, SPLIT 2>
    (512, 512, 64, 94, 126, False),
    (320, 0, 64, 80, 128, True),        # odd number of chunks (5 -> 3 + 2) and a fused pool
]


def _split_k_halo_checks():
    """Body of test_split_k_halo_kernel; runs in a child process with PTK_CONV_HALO_SPLIT=2 (the dispatch reads its
    switches once per process)."""
    from pixtrack_b200.extractor import conv_f16, pack_conv3x3
    for cin0, cin1, cout, H, W, pool in SPLIT_SHAPES:
        x, x1, w, b = _case(cin0, cout, H, W, seed=cin0 + cin1 + cout + H, cin1=cin1)
        for relu in (True, False):
            out = conv_f16(x, pack_conv3x3(w), b, relu=relu, x1=x1, out_hw=(H, W), pool=pool)
            torch.cuda.synchronize()
            y = out[0] if pool else out
            _check(y, _ref(x, w, b, relu, x1=x1, hw=(H, W)), (cin0, cin1, cout, relu))
            if pool:
                want = tF.max_pool2d(y.float().permute(2, 0, 1)[None], 2, 2)[0].permute(1, 2, 0)
                assert torch.equal(out[1].float(), want)
        # run-to-run bit reproducibility (fixed summation order: rank 0's sum + rank 1's sum)
        again = conv_f16(x, pack_conv3x3(w), b, relu=False, x1=x1, out_hw=(H, W))
        torch.cuda.synchronize()
        assert torch.equal(again if not pool else again, y)
    print('split-k halo ok')


def test_split_k_halo_kernel():
    """Small maps with a long K can run the halo kernel with the 64-channel chunks of each tile shared by a cluster of two
    CTAs (partial sums through distributed shared memory; PTK_CONV_HALO_SPLIT, off by default because the per-tap
    kernel's own K split measured faster on B200): against an fp32 convolution of the same operands."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PTK_CONV_HALO_SPLIT='2')
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, '-c', 'import sys; sys.path[:0] = [%r, %r]; import test_conv_paths_gpu as t; '
                        't._split_k_halo_checks()' % (here, os.path.dirname(here))], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and 'split-k halo ok' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_benchmark_layer_shapes_at_full_resolution():
    """The two most expensive launches of the 1024x576 plan at their real size: 64->64 (pooled) and the two-input
    128->32 decoder conv at 576x1024, checked on a random subset of rows (the fp32 reference conv of the whole map is
    evaluated on the GPU by cuDNN in fp32 with TF32 disabled)."""
    from pixtrack_b200.extractor import conv_f16, pack_conv3x3
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        H, W = 576, 1024
        x, _, w, b = _case(64, 64, H, W, seed=101)
        y, p = conv_f16(x, pack_conv3x3(w), b, relu=True, pool=True)
        _check(y, _ref(x, w, b, True), 'enc0.1')
        assert torch.equal(p.float(), tF.max_pool2d(y.float().permute(2, 0, 1)[None], 2, 2)[0].permute(1, 2, 0))
        x, x1, w, b = _case(64, 32, H, W, seed=102, cin1=64)
        y = conv_f16(x, pack_conv3x3(w), b, relu=True, x1=x1, out_hw=(H, W))
        _check(y, _ref(x, w, b, True, x1=x1, hw=(H, W)), 'dec3')
    finally:
        torch.backends.cudnn.allow_tf32 = old
