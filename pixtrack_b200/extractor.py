"""B200FeatureExtractor: drop-in for `PixTrackFeatureExtractor`.

Same call contract as reference pixtrack/localization/feature_extractor.py:34-59:
`extractor(image: np.ndarray[H,W,3] float32 0..255 RGB, scale_image=1) ->
(features: list[Tensor[C_l,H_l,W_l]], scales: list[(sx, sy)], confidences:
list[Tensor[1,H_l,W_l]])`, `extractor.model.scales == [1, 4, 16]`.  The network
(UNet._forward, pixloc/pixloc/pixlib/models/unet.py:158-190) runs as a native plan:
tcgen05 implicit-GEMM convolutions on channels-last fp16 activations
(csrc/ptk_conv.cu, csrc/ptk_unet.cu).  Returned tensors are logical C x H x W views of
channels-last fp32 storage, which `B200Optimizer` consumes without a transpose.
There is no PyTorch fallback.
"""
import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib

Tensor = torch.Tensor
ENC_BLOCKS = ((64, 64), (128, 128), (256, 256, 256, 256), (512, 512, 512, 512), (512, 512, 512, 512))
DECODER = (64, 64, 64, 32)
OUTPUT_SCALES = (0, 2, 4)
BN_EPS = 1e-5


def pack_conv3x3(w: Tensor) -> Tensor:
    """[C_out, C_in, 3, 3] -> fp16 [9][C_out][C_in] (tap-major, C_in contiguous = K-major B operand)."""
    return w.permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1]).to(torch.float16).contiguous()


def pack_weights(sd: Dict[str, Tensor], device) -> Dict[str, List[Tensor]]:
    """State dict with the checkpoint's `extractor.*` keys (prefix stripped) -> packed device tensors.
    Encoder keys follow torchvision vgg19.features indices regrouped into blocks (unet.py:68-99);
    decoder BatchNorm (eval) is folded into the conv (unet.py:22-31)."""
    sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in sd.items() if v.is_floating_point()}
    conv_w, conv_b = [], []
    first = True
    for b in range(5):
        idxs = sorted(int(k.split('.')[2]) for k in sd if k.startswith(f'encoder.{b}.') and k.endswith('.weight'))
        assert len(idxs) == len(ENC_BLOCKS[b]), 'not a VGG19 encoder'
        for i in idxs:
            w, bias = sd[f'encoder.{b}.{i}.weight'], sd[f'encoder.{b}.{i}.bias']
            if first:   # 3 -> 64: fp32 [64][28], taps ordered (ky, kx, c), one pad column
                w0 = torch.zeros(64, 28, device=device)
                w0[:, :27] = w.permute(0, 2, 3, 1).reshape(64, 27)
                conv_w.append(w0.contiguous())
                first = False
            else:
                conv_w.append(pack_conv3x3(w))
            conv_b.append(bias.contiguous())
    for i in range(4):
        p = f'decoder.{i}.layers.'
        scale = sd[p + '1.weight'] / torch.sqrt(sd[p + '1.running_var'] + BN_EPS)
        conv_w.append(pack_conv3x3(sd[p + '0.weight'] * scale[:, None, None, None]))
        conv_b.append((sd[p + '1.bias'] - sd[p + '1.running_mean'] * scale).contiguous())
    head_w, head_b = [], []
    for l in range(3):
        wa, wu = sd[f'adaptation.{l}.0.weight'], sd[f'uncertainty.{l}.0.weight']
        head_w.append(torch.cat([wa.reshape(wa.shape[0], -1), wu.reshape(1, -1)], 0).to(torch.float16).contiguous())
        head_b.append(torch.cat([sd[f'adaptation.{l}.0.bias'], sd[f'uncertainty.{l}.0.bias']]).contiguous())
    return dict(conv_w=conv_w, conv_b=conv_b, head_w=head_w, head_b=head_b)


def conv_f16(x0: Tensor, weights: Tensor, bias: Tensor, relu: bool = True, x1: Optional[Tensor] = None,
             out_hw: Optional[Tuple[int, int]] = None, taps: int = 9, pool: bool = False):
    """One tcgen05 convolution layer (test / building-block entry).  x0 [H0,W0,C0] fp16 channels-last,
    optional x1 [H1,W1,C1]; weights fp16 [taps][C_out][C0+C1]; returns fp16 [H,W,C_out] with
    (H, W) = out_hw or x0's size (larger inputs are cropped to it).  pool=True also returns the fused 2x2 max pool
    [H//2, W//2, C_out] (the encoder's nn.MaxPool2d written by the epilogue): (out, pooled)."""
    assert x0.is_cuda and x0.dtype == torch.float16 and x0.is_contiguous()
    H, W = out_hw if out_hw is not None else x0.shape[:2]
    Cout = weights.shape[1]
    out = torch.empty((H, W, Cout), dtype=torch.float16, device=x0.device)
    pooled = torch.empty((H // 2, W // 2, Cout), dtype=torch.float16, device=x0.device) if pool else None
    dev = x0.device.index if x0.device.index is not None else torch.cuda.current_device()
    _lib.check(_lib.load().ptk_conv_f16_pool(
        _lib.context(dev), x0.data_ptr(), x0.shape[2], None if x1 is None else x1.data_ptr(),
        0 if x1 is None else x1.shape[2], H, W, x0.shape[0], x0.shape[1], 0 if x1 is None else x1.shape[0],
        0 if x1 is None else x1.shape[1], weights.data_ptr(), bias.data_ptr(), Cout, taps, 1 if relu else 0,
        out.data_ptr(), None if pooled is None else pooled.data_ptr(), _lib.current_stream_ptr(x0.device)))
    return (out, pooled) if pool else out


class _Plan:
    def __init__(self, lib, ctx, wts_struct, H, W):
        self.lib, self.h = lib, C.c_void_p()
        _lib.check(lib.ptk_extractor_create(ctx, C.byref(wts_struct), H, W, C.byref(self.h)))
        self.shapes = []
        for l in range(3):
            c, h, w = C.c_int32(), C.c_int32(), C.c_int32()
            _lib.check(lib.ptk_extractor_level_shape(self.h, l, C.byref(c), C.byref(h), C.byref(w)))
            self.shapes.append((c.value, h.value, w.value))

    def __del__(self):
        try:
            self.lib.ptk_extractor_destroy(self.h)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class B200FeatureExtractor(torch.nn.Module):
    default_conf = dict(resize=1024, resize_by='max')

    def __init__(self, state_dict: Dict[str, Tensor], device, conf: Optional[dict] = None):
        super().__init__()
        self.conf = SimpleNamespace(**{**self.default_conf, **dict(conf or {})})
        assert self.conf.resize_by in ('max', 'max_force')
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.PtkError('B200FeatureExtractor needs a CUDA device (no CPU fallback)')
        self.model = SimpleNamespace(scales=[2 ** s for s in OUTPUT_SCALES])
        self._packed = pack_weights(state_dict, self.device)
        w = _lib.UnetWeights()
        for i in range(20):
            w.conv_w[i] = self._packed['conv_w'][i].data_ptr()
            w.conv_b[i] = self._packed['conv_b'][i].data_ptr()
        for l in range(3):
            w.head_w[l] = self._packed['head_w'][l].data_ptr()
            w.head_b[l] = self._packed['head_b'][l].data_ptr()
        self._wts = w
        self._plans: Dict[Tuple[int, int], _Plan] = {}
        self._lib = _lib.load()
        self._ctx = _lib.context(self.device.index if self.device.index is not None else torch.cuda.current_device())

    @classmethod
    def from_reference(cls, ref_extractor) -> 'B200FeatureExtractor':
        """Build from a live reference PixTrackFeatureExtractor / FeatureExtractor (its UNet's weights)."""
        conf = dict(resize=ref_extractor.conf.resize, resize_by=ref_extractor.conf.resize_by)
        return cls(ref_extractor.model.state_dict(), ref_extractor.device, conf)

    def network_size(self, h: int, w: int, scale_image: int = 1) -> Tuple[int, int, Tuple[float, float]]:
        """Resize rule of feature_extractor.py:40-44 + view.py:31-37 (max edge -> resize // scale)."""
        if self.conf.resize is not None:
            target = self.conf.resize // scale_image
            if max(h, w) > target or self.conf.resize_by == 'max_force':
                s = target / max(h, w)
                return int(round(h * s)), int(round(w * s)), (s, s)
        return h, w, (1.0, 1.0)

    def plan(self, H: int, W: int, slot: int = 0) -> _Plan:
        """Plans own their activation arena: use a different `slot` for extractions that may run concurrently on
        different streams (e.g. the reference view next to the query frame)."""
        if (H, W, slot) not in self._plans:
            self._plans[(H, W, slot)] = _Plan(self._lib, self._ctx, self._wts, H, W)
        return self._plans[(H, W, slot)]

    def launch_counts(self):
        """Kernel launches of the last run of every plan: {(H, W, slot): 28 or 27} (27 = the level-0 head ran fused into
        the last decoder convolution)."""
        out = {}
        for key, plan in self._plans.items():
            n = C.c_int32()
            _lib.check(self._lib.ptk_extractor_launch_count(plan.h, C.byref(n)))
            out[key] = n.value
        return out

    def level_shapes(self, ih: int, iw: int, scale_image: int = 1):
        H, W, _ = self.network_size(ih, iw, scale_image)
        return self.plan(H, W).shapes

    def extract_device(self, image: Tensor, scale_image: int = 1, normalize: bool = False, out=None, slot: int = 0):
        """image: CUDA [H,W,3] RGB, fp32 in 0..255 or uint8.  Returns (feats_hwc list [H_l,W_l,C_l], confs list
        [H_l,W_l], scales); nothing is synchronised.  `out=(feats, confs)` writes into caller-owned buffers
        (static addresses for prepared LM launches / CUDA graphs)."""
        assert image.is_cuda and image.dtype in (torch.float32, torch.uint8) and image.is_contiguous()
        assert image.shape[2] == 3
        ih, iw = image.shape[:2]
        H, W, sr = self.network_size(ih, iw, scale_image)
        plan = self.plan(H, W, slot)
        if out is None:
            feats = [torch.empty((h, w, c), dtype=torch.float32, device=self.device) for c, h, w in plan.shapes]
            confs = [torch.empty((h, w), dtype=torch.float32, device=self.device) for c, h, w in plan.shapes]
        else:
            feats, confs = out
        fp = (C.c_void_p * 3)(*[t.data_ptr() for t in feats])
        cp = (C.c_void_p * 3)(*[t.data_ptr() for t in confs])
        _lib.check(self._lib.ptk_extractor_run(plan.h, image.data_ptr(), 0 if image.dtype == torch.float32 else 1, ih,
                                               iw, fp, cp, 1 if normalize else 0,
                                               _lib.current_stream_ptr(self.device)))
        scales = [(sr[0] / s, sr[1] / s) for s in self.model.scales]
        return feats, confs, scales

    def profile(self, image: Tensor, scale_image: int = 1, normalize: bool = True):
        """Per-launch device times of one extraction (synchronises): list of (kind, ms, flops)."""
        ih, iw = image.shape[:2]
        H, W, _ = self.network_size(ih, iw, scale_image)
        plan = self.plan(H, W)
        feats = [torch.empty((h, w, c), dtype=torch.float32, device=self.device) for c, h, w in plan.shapes]
        confs = [torch.empty((h, w), dtype=torch.float32, device=self.device) for c, h, w in plan.shapes]
        fp = (C.c_void_p * 3)(*[t.data_ptr() for t in feats])
        cp = (C.c_void_p * 3)(*[t.data_ptr() for t in confs])
        ms, kinds, flops, n = (C.c_float * 48)(), (C.c_int32 * 48)(), (C.c_double * 48)(), C.c_int32()
        _lib.check(self._lib.ptk_extractor_profile(plan.h, image.data_ptr(), 0 if image.dtype == torch.float32 else 1,
                                                   ih, iw, fp, cp, 1 if normalize else 0,
                                                   _lib.current_stream_ptr(self.device), 48, ms, kinds, flops,
                                                   C.byref(n)))
        names = {0: 'prep', 1: 'conv1_mma', 2: 'maxpool', 3: 'conv_tc', 4: 'upsample', 5: 'head'}
        return [(names[kinds[i]], ms[i], flops[i]) for i in range(n.value)]

    def activation(self, H: int, W: int, kind: int, index: int) -> Tensor:
        """Copy of an intermediate fp16 activation of the (H, W) plan (tests)."""
        p, c, h, w = C.c_void_p(), C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(self._lib.ptk_extractor_activation(self.plan(H, W).h, kind, index, C.byref(p), C.byref(c), C.byref(h),
                                                      C.byref(w)))
        n = c.value * h.value * w.value
        out = torch.empty((h.value, w.value, c.value), dtype=torch.float16, device=self.device)
        _lib.check(self._lib.ptk_copy_d2d(out.data_ptr(), p.value, n * 2, _lib.current_stream_ptr(self.device)))
        return out

    @torch.no_grad()
    def __call__(self, image: np.ndarray, scale_image: int = 1):
        image = np.ascontiguousarray(image)
        if image.dtype != np.uint8:
            image = image.astype(np.float32, copy=False)
        img = torch.from_numpy(image).to(self.device, non_blocking=True)
        feats, confs, scales = self.extract_device(img, scale_image, normalize=False)
        return [f.permute(2, 0, 1) for f in feats], scales, [c[None] for c in confs]
