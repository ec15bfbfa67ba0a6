"""ctypes binding of ``libll_b200.so`` (the C ABI declared in ``include/ll_api.h``).

There is NO fallback: if the library is missing, or the device is not a B200-class
(sm_100) GPU, every op raises.  PyTorch is only used for device memory and streams.
"""
import ctypes
import os

import torch

from . import build as _build

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_float = ctypes.c_float
c_voidp = ctypes.c_void_p

LL_OK, LL_EINVAL, LL_EARCH, LL_ECUDA = 0, -1, -2, -3


class LLError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libll_b200 error {code}: {msg}")
        self.code = code


class ll_view3(ctypes.Structure):
    _fields_ = [("ptr", c_voidp), ("sb", c_i64), ("sy", c_i64), ("sx", c_i64)]


class ll_lift_job(ctypes.Structure):
    _fields_ = [("src", ll_view3), ("din", ll_view3), ("dout", ll_view3),
                ("nb", ctypes.c_int32), ("ny", ctypes.c_int32), ("nx", ctypes.c_int32)]


# name -> (restype, argtypes); mirrors include/ll_api.h declaration by declaration
_P = c_voidp  # device pointers are passed as integers (tensor.data_ptr())
SIGNATURES = {
    "ll_last_error": (ctypes.c_char_p, []),
    "ll_version": (c_int, []),
    "ll_check_device": (c_int, []),
    "ll_sm_count": (c_int, []),
    "ll_pack_lift_step": (c_int, [_P] * 10 + [_P]),
    "ll_lift_step": (c_int, [ctypes.POINTER(ll_lift_job), c_int, _P, c_float, c_float, c_int, c_int, _P]),
    "ll_lift_level_scratch_floats": (ctypes.c_size_t, [c_int, c_int, c_int]),
    "ll_lift_level_fwd": (c_int, [_P, c_i64, _P, c_i64, _P, c_i64, _P, c_int, c_int, c_int,
                                  ctypes.POINTER(c_voidp), c_float, c_int, c_int, _P, _P, c_int, _P]),
    "ll_lift_level_inv": (c_int, [_P, c_i64, _P, c_i64, _P, c_i64, _P, c_int, c_int, c_int,
                                  ctypes.POINTER(c_voidp), c_float, c_int, c_int, _P, _P, c_int, _P]),
    "ll_dwt97_fwd_level": (c_int, [_P, c_i64, _P, c_i64, _P, c_i64, c_int, c_int, c_int, _P]),
    "ll_dwt97_inv_level": (c_int, [_P, c_i64, _P, c_i64, _P, c_i64, c_int, c_int, c_int, _P]),
    "ll_dwt97_scratch_floats": (ctypes.c_size_t, [c_int, c_int, c_int, c_int]),
    "ll_dwt97_fwd": (c_int, [_P, _P, ctypes.POINTER(c_voidp), _P, c_int, c_int, c_int, c_int, _P]),
    "ll_dwt97_inv": (c_int, [_P, ctypes.POINTER(c_voidp), _P, _P, c_int, c_int, c_int, c_int, _P]),
    "ll_pack_ae1": (c_int, [_P] * 8 + [c_int, c_int, _P, _P]),
    "ll_ae1_apply": (c_int, [_P, _P, _P, _P, c_int, c_int, c_i64, _P]),
    "ll_conv2d": (c_int, [_P, c_i64, _P, _P, _P, c_i64] + [c_int] * 12 + [_P]),
    "ll_pw_mlp3": (c_int, [_P, c_i64, _P, _P, _P, _P, _P, _P, _P, c_i64, c_int, c_int, c_i64, _P]),
    "ll_ctx_conv_nhwc": (c_int, [_P, _P, _P, _P] + [c_int] * 15 + [_P]),
    "ll_ctx_im2col": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "ll_igemm_cgp_tail": (c_int, [_P, _P, _P] + [c_int] * 7 + [ctypes.POINTER(c_int), _P, _P, _P, _P, c_int, _P, c_i64, _P, _P, c_i64, _P, _P]),
    "ll_cgp_tail_rate": (c_int, [_P, c_i64, _P, _P, _P, _P, _P, c_i64, _P, _P, c_i64, _P, _P, c_int, c_int, c_int, c_int, c_i64, _P, _P]),
    "ll_pack_igemm_weight": (c_int, [_P, _P] + [c_int] * 5 + [_P]),
    "ll_nchw_to_nhwc_bf16": (c_int, [_P, c_i64, _P] + [c_int] * 6 + [_P]),
    "ll_igemm_conv": (c_int, [_P, _P, _P] + [c_int] * 9 + [ctypes.POINTER(c_int), c_int, _P, c_i64, c_int, c_int, c_int, _P,
                              c_int, c_int, c_int, _P]),
    "ll_pack_tf32_weight": (c_int, [_P, _P] + [c_int] * 6 + [_P]),
    "ll_igemm_tf32": (c_int, [_P, _P, _P] + [c_int] * 9 + [_P, _P, c_int, _P]),
    "ll_igemm_tf32_gdn": (c_int, [_P, _P, _P, _P, _P] + [c_int] * 7 + [_P, _P]),
    "ll_conv3_gdn_head": (c_int, [_P, _P, _P, _P, _P] + [c_int] * 6 + [_P, _P]),
    "ll_nchw_to_nhwc_split": (c_int, [_P, _P, _P] + [c_int] * 5 + [_P]),
    "ll_nhwc_split_to_nchw": (c_int, [_P, _P] + [c_int] * 4 + [_P]),
    "ll_nhwc_split_conv3": (c_int, [_P, _P, _P, _P] + [c_int] * 5 + [_P]),
    "ll_nhwc_lrelu_conv1": (c_int, [_P, _P, _P, _P, c_int, c_i64, c_int, c_int, c_int, c_int, _P]),
    "ll_rans_stream_cap": (c_i64, [c_i64, c_int]),
    "ll_rans_encode": (c_int, [c_int, _P, _P, c_int, c_int, c_i64, c_int, _P, _P, _P]),
    "ll_rans_pack": (c_int, [_P, _P, _P, c_i64, c_int, _P, _P]),
    "ll_rans_decode": (c_int, [c_int, _P, _P, _P, c_int, c_int, c_i64, c_int, _P, _P]),
    "ll_rgb_to_ycbcr_shift": (c_int, [_P, _P, c_int, c_i64, _P]),
    "ll_ycbcr_to_rgb_sse": (c_int, [_P, _P, _P, c_int, c_i64, _P, _P]),
    "ll_quantize": (c_int, [_P, _P, _P, c_i64, _P]),
    "ll_gauss_rate": (c_int, [_P, c_i64, _P, c_i64, _P, _P, c_i64, _P, c_int, c_int, c_i64, _P, _P]),
    "ll_pack_eb": (c_int, [ctypes.POINTER(c_voidp), c_int, _P, _P]),
    "ll_eb_rate": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_i64, _P, _P]),
}

# libll_probe.so (include/ll_probe.h): measurement / unit probes, test and bench tooling only
PROBE_SIGNATURES = {
    "ll_fma_peak_probe": (c_int, [_P, c_int, c_int, _P]),
    "ll_tc_tf32_probe": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "ll_tf32_peak_probe": (c_int, [_P, c_int, c_int, c_int, c_int, _P]),
    "ll_halo_probe": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "ll_probe_set_timeline": (c_int, [_P]),
    "ll_probe_set_nostore": (c_int, [c_int]),
    "ll_dbg_lift_switches": (c_int, [c_int]),
    "ll_dbg_lift_stamp_buffer": (c_int, [_P]),
    "ll_last_error": (ctypes.c_char_p, []),
}

_lib = None
_probe = None
_device_ok = {}


def load_probe():
    """The probe library (``libll_probe.so``): not part of the product ABI."""
    global _probe
    if _probe is not None:
        return _probe
    try:
        _build.build()
    except Exception:
        pass
    path = _build.PROBE_LIB_PATH
    if not os.path.isfile(path):
        raise RuntimeError(f"{path} is missing: run `python -m {__package__}.build`")
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROBE_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _probe = lib
    return lib


def check_probe(rc):
    if rc != LL_OK:
        raise LLError(rc, load_probe().ll_last_error().decode(errors="replace"))


def load(build_if_missing=True):
    """Load (building first if stale and nvcc is present) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if build_if_missing:
        try:
            path = _build.build()
        except Exception as e:  # no nvcc on this box: use the shipped .so if there is one
            if not os.path.isfile(path):
                raise RuntimeError(f"libll_b200.so is missing and could not be built: {e}") from e
    if not os.path.isfile(path):
        raise RuntimeError(f"{path} is missing: run `python -m "
                           f"{__package__}.build` (there is no CPU / PyTorch fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != LL_OK:
        raise LLError(rc, load().ll_last_error().decode(errors="replace"))


def require_device(t):
    """The tensor must live on a CUDA device this library can run on."""
    if not t.is_cuda:
        raise RuntimeError("ll_b200 ops need CUDA tensors: there is no CPU fallback")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _device_ok:
        with torch.cuda.device(idx):
            check(load().ll_check_device())
        _device_ok[idx] = True


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return t.data_ptr() if t is not None else None
