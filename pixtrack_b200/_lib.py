"""ctypes binding of libpixtrack_b200.so (include/pixtrack_b200.h).

There is NO fallback: if the library is missing or no sm_100 device is
present, loading / context creation raises.  The oracle is never imported
from here.
"""
import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libpixtrack_b200.so')
LOG_STRIDE = 64
ABI_VERSION = 2

c_f32p = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_uint8)
c_i32p = C.POINTER(C.c_int32)


class PtkError(RuntimeError):
    pass


class LmProblem(C.Structure):
    _fields_ = [
        ('B', C.c_int32), ('N', C.c_int32), ('C', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
        ('n_cam', C.c_int32), ('num_iters', C.c_int32), ('pad', C.c_int32), ('min_valid', C.c_int32),
        ('reserved0', C.c_int32),
        ('p3d', C.c_void_p), ('p3d_bstride', C.c_int64),
        ('f_ref', C.c_void_p), ('f_ref_bstride', C.c_int64),
        ('w_ref', C.c_void_p), ('w_ref_bstride', C.c_int64),
        ('fq', C.c_void_p), ('fq_bstride', C.c_int64),
        ('wq', C.c_void_p), ('wq_bstride', C.c_int64),
        ('mask', C.c_void_p), ('mask_bstride', C.c_int64),
        ('cam', C.c_void_p), ('cam_bstride', C.c_int64),
        ('T_init', C.c_void_p), ('T_bstride', C.c_int64),
        ('lambda_', C.c_void_p), ('lambda_bstride', C.c_int64),
        ('skip', C.c_void_p),
        ('loss_scale', C.c_float), ('grad_stop', C.c_float), ('dt_stop', C.c_float), ('dR_stop', C.c_float),
        ('workspace', C.c_void_p), ('workspace_bytes', C.c_int64),
    ]


class UnetWeights(C.Structure):
    _fields_ = [('conv_w', C.c_void_p * 20), ('conv_b', C.c_void_p * 20), ('head_w', C.c_void_p * 3),
                ('head_b', C.c_void_p * 3)]


class RefLevel(C.Structure):
    _fields_ = [('feat', C.c_void_p), ('conf', C.c_void_p), ('f_out', C.c_void_p), ('w_out', C.c_void_p),
                ('sx', C.c_double), ('sy', C.c_double), ('C', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('normalize', C.c_int32)]


class NerfModelStruct(C.Structure):
    _fields_ = [('grid', C.c_void_p), ('n_grid_entries', C.c_int64), ('weights', C.c_void_p * 5),
                ('bitfield', C.c_void_p), ('aabb_scale', C.c_int32), ('reserved0', C.c_int32)]


class NerfView(C.Structure):
    _fields_ = [('camera', C.c_float * 12), ('render_aabb_min', C.c_float * 3), ('render_aabb_max', C.c_float * 3),
                ('focal', C.c_float), ('depth_scale', C.c_float), ('min_transmittance', C.c_float),
                ('background', C.c_float * 4), ('width', C.c_int32), ('height', C.c_int32), ('spp', C.c_int32),
                ('depth_mode', C.c_int32)]


class LmResult(C.Structure):
    _fields_ = [('T', C.c_void_p), ('failed', C.c_void_p), ('n_iters', C.c_void_p), ('log', C.c_void_p)]


# name -> (restype, argtypes); every symbol include/pixtrack_b200.h declares
SYMBOLS = {
    'ptk_abi_version': (C.c_int, []),
    'ptk_last_error': (C.c_char_p, []),
    'ptk_create': (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    'ptk_destroy': (None, [C.c_void_p]),
    'ptk_num_sms': (C.c_int, [C.c_void_p]),
    'ptk_device_status': (C.c_int, [C.c_void_p]),
    'ptk_lm_run': (C.c_int, [C.c_void_p, C.POINTER(LmProblem), C.POINTER(LmResult), C.c_void_p]),
    'ptk_lm_plan': (C.c_int, [C.c_void_p, C.POINTER(LmProblem), c_i32p, c_i32p]),
    'ptk_lm_workspace_bytes': (C.c_int64, []),
    'ptk_sample_points': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    'ptk_sample_reference': (C.c_int, [C.c_void_p, C.POINTER(RefLevel), C.c_int32, C.c_void_p, C.c_int32,
                                       C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double), C.c_int32, C.c_void_p,
                                       C.c_void_p]),
    'ptk_conv_f16': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                               C.c_void_p, C.c_void_p]),
    'ptk_conv_f16_pool': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ptk_extractor_create': (C.c_int, [C.c_void_p, C.POINTER(UnetWeights), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    'ptk_extractor_destroy': (None, [C.c_void_p]),
    'ptk_extractor_level_shape': (C.c_int, [C.c_void_p, C.c_int32, c_i32p, c_i32p, c_i32p]),
    'ptk_extractor_run': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    'ptk_extractor_profile': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_int32, c_f32p, c_i32p,
                                        C.POINTER(C.c_double), c_i32p]),
    'ptk_extractor_launch_count': (C.c_int, [C.c_void_p, c_i32p]),
    'ptk_extractor_activation': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), c_i32p, c_i32p,
                                           c_i32p]),
    'ptk_nerf_create': (C.c_int, [C.c_void_p, C.POINTER(NerfModelStruct), C.POINTER(C.c_void_p)]),
    'ptk_nerf_destroy': (None, [C.c_void_p]),
    'ptk_nerf_grid_entries': (C.c_int64, [C.c_int32]),
    'ptk_nerf_render': (C.c_int, [C.c_void_p, C.POINTER(NerfView), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ptk_nerf_eval': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ptk_nerf_stats': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    'ptk_query_mask': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    'ptk_overlay': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.POINTER(C.c_int16),
                              C.c_int32, C.c_void_p, C.c_void_p]),
    'ptk_copy_d2d': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'ptk_chw_to_hwc': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_void_p]),
}

_lib = None
_lock = threading.Lock()
_contexts = {}


def load():
    """dlopen the library and bind every declared symbol (no GPU needed)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PtkError(f'{LIB_PATH} not found: build it with `python -m pixtrack_b200.build` '
                               '(there is no CPU fallback)')
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)      # AttributeError if the symbol is not exported
                fn.restype = res
                fn.argtypes = args
            if lib.ptk_abi_version() != ABI_VERSION:
                raise PtkError('ABI version mismatch between _lib.py and libpixtrack_b200.so')
            _lib = lib
    return _lib


def check(code: int):
    if code != 0:
        raise PtkError(f'libpixtrack_b200 error {code}: {load().ptk_last_error().decode()}')


def context(device_index: int) -> int:
    """Per-device PtkContext handle (created on first use)."""
    lib = load()
    import torch
    if not torch.cuda.is_available():
        # never enter the CUDA runtime without a device (it can block for minutes on a GPU-less host)
        raise PtkError('no CUDA device: pixtrack_b200 runs only on sm_100a GPUs (no CPU fallback)')
    code = 0
    with _lock:
        h = _contexts.get(device_index)
        if h is None:
            h = C.c_void_p()
            code = lib.ptk_create(int(device_index), C.byref(h))
            if code == 0:
                _contexts[device_index] = h
    check(code)          # outside the lock: check() re-enters load(), which takes it
    return h


def current_stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def device_status(device_index: int):
    check(load().ptk_device_status(context(device_index)))
