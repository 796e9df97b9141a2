"""ctypes binding of the C ABI in include/keypoints_b200.h.

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C keypoints_b200/csrc`` into
``keypoints_b200/lib/libkeypoints_b200.so``.  There is no fallback: if the library is missing or a call
fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('KP_LIB') or os.path.join(_HERE, 'lib', 'libkeypoints_b200.so')   # KP_LIB: experiment builds

F32, BF16 = 0, 1
ACT_NONE, ACT_LEAKY, ACT_RELU = 0, 1, 2
POST_NONE, POST_POOL, POST_UP = 0, 1, 2
POST_CORESIDENT = 0x100        # hint for the forward BatchNorm passes: tensor-core convs run concurrently on another stream
ACTS = {None: ACT_NONE, 'none': ACT_NONE, 'leaky': ACT_LEAKY, 'relu': ACT_RELU}
COMBINE = {'max': 0, 'sum_and_clamp': 1, 'loop': 2}
POSTS = {None: POST_NONE, 'none': POST_NONE, 'pool': POST_POOL, 'up': POST_UP}


class KpError(RuntimeError):
    pass


class KpView(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('sn', C.c_int64), ('sy', C.c_int64), ('sx', C.c_int64), ('sc', C.c_int64),
                ('dtype', C.c_int32), ('_pad', C.c_int32)]


DP_MAX_WORLD = 8


class KpDpPeers(C.Structure):
    _fields_ = [('g', C.c_void_p * DP_MAX_WORLD), ('p', C.c_void_p * DP_MAX_WORLD), ('flag', C.c_void_p * DP_MAX_WORLD),
                ('mc_g', C.c_void_p), ('mc_p', C.c_void_p)]


_P = C.c_void_p
_VP = C.POINTER(KpView)
_I, _L, _F, _D = C.c_int, C.c_int64, C.c_float, C.c_double

# name -> argtypes (restype is always int); mirrors include/keypoints_b200.h
SIGNATURES = {
    'kp_device_info': [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    'kp_ctx_create': [C.POINTER(C.c_void_p)],
    'kp_ctx_destroy': [_P],
    'kp_ctx_set_current': [_P],
    'kp_ctx_set_sm_limit': [_P, _I],
    'kp_ctx_info': [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                    C.POINTER(C.c_int64)],
    'kp_conv_simt': [_P, _VP, _P, _P, _VP, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I],
    'kp_conv_wgrad_simt': [_P, _VP, _VP, _P, _I, _I, _I, _I, _I, _I],
    'kp_conv_tc': [_P, _P, _L, _I, _P, _I, C.POINTER(C.c_int32), _P, _P, _I, _P, _I, _I, _I, _I],
    'kp_conv_wgrad_tc': [_P, _P, _P, _L, _I, _I, _I, _I, C.POINTER(C.c_int32), _P, _P],
    'kp_conv_wgrad_tc_img': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    'kp_wgrad_finalize_multi': [_P, _P, _I],
    'kp_pack_weights': [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P],
    'kp_pack_weights_multi': [_P, _P, _I],
    'kp_bn_finalize': [_P, _P, _I, _D, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    'kp_bn_act_fwd': [_P, _VP, _VP, _P, _P, _I, _I, _I, _I, _I, _I, _I],
    'kp_bn_act_bwd_reduce': [_P, _VP, _VP, _VP, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I],
    'kp_bn_act_bwd_apply': [_P, _VP, _VP, _P, _P, _P, _P, _D, _I, _I, _I, _I, _P, _P],
    'kp_bn_act_bwd_apply_gather': [_P, _VP, _VP, _VP, _P, _P, _P, _P, _P, _D, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    'kp_bn_finalize_act_fwd': [_P, _P, _D, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _VP, _VP, _I, _I, _I, _I, _I, _I, _I],
    'kp_bn_grad_finalize': [_P, _P, _I, _P, _P],
    'kp_spatial_softmax_fwd': [_P, _P, _I, _I, _I, _P, _P, _P],
    'kp_spatial_softmax_bwd': [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    'kp_gaussian_fwd': [_P, _P, _I, _I, _I, _F, _F, _P],
    'kp_gaussian_bwd': [_P, _VP, _I, _P, _P, _I, _I, _I, _I, _F, _F, _P],
    'kp_transport_fwd': [_P, _VP, _VP, _P, _P, _VP, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F],
    'kp_transport_bwd': [_P, _VP, _I, _VP, _VP, _P, _P, _VP, _P, _I, _I, _I, _I],
    'kp_transport_mode_fwd': [_P, _I, _VP, _VP, _P, _P, _VP, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F],
    'kp_transport_mode_bwd': [_P, _I, _VP, _I, _VP, _VP, _P, _P, _P, _P, _P, _P, _VP, _P, _I, _I, _I, _I, _I, _F, _F],
    'kp_l2_loss': [_P, _P, _P, _P, _L, _F, _P, _P],
    'kp_tps_warp': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I],
    'kp_rotate_warp': [_P, _P, _P, _P, _I, _I, _I, _I],
    'kp_aug_draw': [_P, C.c_uint64, _P, _I, _I, _I, _F, _F, _P, _P, _P],
    'kp_zero': [_P, _P, _L],
    'kp_jpeg_create': [C.POINTER(C.c_void_p)],
    'kp_jpeg_destroy': [_P],
    'kp_jpeg_info': [_P, _P, _L, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    'kp_jpeg_decode': [_P, _P, _P, _L, _P, _I, _I],
    'kp_resize_to_f32': [_P, _P, _I, _I, _I, _P, _P, _I, _I],
    'kp_u8_to_f32': [_P, _P, _P, _I, _I, _I, _I, _F, _F],
    'kp_loss_ring_push': [_P, _P, _D, _P, _P, _I],
    'kp_dp_adam_step': [_P, C.POINTER(KpDpPeers), _I, _I, _L, _P, _P, _D, _D, _D, _D, _F, _P, _P, _I],
    'kp_adam_step': [_P, _P, _P, _P, _P, _L, _D, _D, _D, _D, _I, _F, _P],
}
EXPORTS = sorted(list(SIGNATURES) + ['kp_last_error', 'kp_version'])

_lib = None
launches = 0          # number of C-ABI calls that launch kernels (bench.py reports it)


def load():
    """Load the shared library (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KpError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                      f'or `make -C keypoints_b200/csrc`. There is no CPU / PyTorch fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.kp_last_error.restype = C.c_char_p
    lib.kp_last_error.argtypes = []
    lib.kp_version.restype = C.c_int
    _lib = lib
    return lib


timing = None         # when a list: (name, flops, start_event, end_event) per call (bench.py roofline leg)


def call(name, *args, flops=0.0, tag='', count=True):
    global launches
    lib = load()
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise KpError(f'{name} failed ({rc}): {lib.kp_last_error().decode()}')
    if timing is not None:
        e1.record()
        timing.append((name, flops, e0, e1, tag))
    if count:
        launches += 1


class Context:
    """kp_ctx: per-device SM budget + tensor-map cache (include/keypoints_b200.h).  `use()` binds it to the calling thread."""

    def __init__(self):
        self.handle = C.c_void_p()
        call('kp_ctx_create', C.byref(self.handle), count=False)

    def use(self):
        call('kp_ctx_set_current', self.handle, count=False)

    def set_sm_limit(self, sms: int):
        call('kp_ctx_set_sm_limit', self.handle, int(sms), count=False)

    def info(self):
        dev, sm, lim = C.c_int(), C.c_int(), C.c_int()
        hits, misses, cached = C.c_int64(), C.c_int64(), C.c_int64()
        call('kp_ctx_info', self.handle, C.byref(dev), C.byref(sm), C.byref(lim), C.byref(hits), C.byref(misses), C.byref(cached),
             count=False)
        return {'device': dev.value, 'sm_count': sm.value, 'sm_limit': lim.value, 'map_hits': hits.value,
                'map_misses': misses.value, 'maps_cached': cached.value}

    def close(self):
        if self.handle:
            load().kp_ctx_set_current(None)
            call('kp_ctx_destroy', self.handle, count=False)
            self.handle = C.c_void_p()


def zero(t):
    """Zero a contiguous tensor with a memset node on the current stream (no ATen fill kernel inside the step)."""
    call('kp_zero', stream(), ptr(t), t.numel() * t.element_size(), count=False)


def device_info():
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    call('kp_device_info', C.byref(sm), C.byref(ma), C.byref(mi))
    return sm.value, ma.value, mi.value


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise KpError(f'unsupported dtype {t.dtype}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise KpError('keypoints_b200 kernels need CUDA tensors (there is no CPU path)')
    return C.c_void_p(t.data_ptr())


def view(t):
    """kp_view of a 4-D tensor indexed [n, y, x, c] (use .permute(0,2,3,1) for NCHW tensors)."""
    if not t.is_cuda:
        raise KpError('keypoints_b200 kernels need CUDA tensors (there is no CPU path)')
    sn, sy, sx, sc = t.stride()
    return C.byref(KpView(t.data_ptr(), sn, sy, sx, sc, dt(t), 0))


def nchw(t):
    return view(t.permute(0, 2, 3, 1))


def shifts_array(vals):
    return (C.c_int32 * len(vals))(*vals)
