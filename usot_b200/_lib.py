"""ctypes binding of libusot_b200.so (the C ABI declared in include/usot_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("USOT_B200_LIB") or os.path.join(_HERE, "libusot_b200.so")   # (the override exists for same-job A/B runs of two builds)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "usot_b200.h")

PREC_FP32_SIMT, PREC_FP16X3_TC, PREC_FP16_TC = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32_SIMT, "fp16x3": PREC_FP16X3_TC, "fp16": PREC_FP16_TC}

_lib = None
_lock = threading.Lock()

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_I64 = ctypes.c_int64
_D = ctypes.c_double

# name -> (restype, argtypes); must list every USOT_API symbol of the header (tests/test_abi.py checks this)
SIGNATURES = {
    "usot_last_error": (ctypes.c_char_p, []),
    "usot_abi_version": (_I, []),
    "usot_set_tunable": (_I, [ctypes.c_char_p, _I]),
    "usot_profile_reset": (_I, [_I]),
    "usot_profile_family_count": (_I, []),
    "usot_profile_family_name": (ctypes.c_char_p, [_I]),
    "usot_profile_read": (_I, [_I, ctypes.POINTER(ctypes.c_double)]),
    "usot_profile_count": (_I, [_I, _I64]),
    "usot_prroi_pool_forward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P]),
    "usot_prroi_pool_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P]),
    "usot_prroi_pool_coor_backward": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "usot_xcorr_depthwise": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "usot_xcorr_depthwise_backward": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "usot_groupdw_xcorr": (_I, [_P] * 8 + [_I] * 5 + [_P]),
    "usot_conv2d_nhwc": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _I, _P]),
    "usot_conv2d_nhwc_scaled": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "usot_pred_conv": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, ctypes.c_float, _P, _P, _P, _P]),
    "usot_engine_track_frame": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _I, _P, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_double, _I, _P, _P, _P]),
    "usot_maxpool3x3s2p1_nhwc": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "usot_stem_conv": (_I, [_P, _I, _I, _P, _P, _P, _P, _I, _P]),
    "usot_stem_maxpool": (_I, [_P, _I, _I, _P, _P, _P, _P, _I, _P]),
    "usot_stem_conv_raw": (_I, [_P, _I, _I, _P, _P, _P]),
    "usot_conf_fusion": (_I, [_P, _P, _I, _I, _I64, _P, _P]),
    "usot_cycle_glue": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P]),
    "usot_weighted_bce": (_I, [_P, _P, _I, _P, _P]),
    "usot_iou_loss": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "usot_conv2d_wgrad_nhwc": (_I, [_P, _P] + [_I] * 12 + [_P, _I, _P]),
    "usot_stem_conv_wgrad": (_I, [_P, _P, _I, _I, _P, _P]),
    "usot_pow2_scale": (_I, [_P, _I64, _I, _P, _P, _P]),
    "usot_conv2d_dgrad_nhwc": (_I, [_P, _P] + [_I] * 12 + [_P, _P]),
    "usot_bn_stats": (_I, [_P, _P, _I64, _I, _P, _P, _P]),
    "usot_bn_apply": (_I, [_P, _P, _P, _P, _F, _P, _P, _P, _I, _I64, _I, _P, _P, _P]),
    "usot_bn_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I64, _I, _P, _P, _P, _P, _P]),
    "usot_channel_sum": (_I, [_P, _I64, _I, _P, _P]),
    "usot_maxpool3x3s2p1_backward_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "usot_conf_fusion_backward": (_I, [_P, _P, _P, _I, _I, _I64, _P, _P, _P]),
    "usot_weighted_sum3": (_I, [_P, _P, _P, _P, _I64, _P, _P]),
    "usot_weighted_sum3_backward": (_I, [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "usot_weighted_bce_backward": (_I, [_P, _P, _I, _P, _P, _P]),
    "usot_iou_loss_backward": (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "usot_crop_resize": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "usot_nchw_to_nhwc": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "usot_nhwc_to_nchw": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "usot_engine_create": (_I, [ctypes.POINTER(_P), _I, _I]),
    "usot_engine_destroy": (_I, [_P]),
    "usot_engine_load_tensor": (_I, [_P, ctypes.c_char_p, _P, _I64]),
    "usot_engine_finalize": (_I, [_P]),
    "usot_engine_device_bytes": (_I64, [_P]),
    "usot_engine_packed_size": (_I64, [_P]),
    "usot_engine_export_packed": (_I, [_P, _P, _I64]),
    "usot_engine_import_packed": (_I, [_P, _P, _I64]),
    "usot_engine_backbone_neck": (_I, [_P, _P, _I, _I, _P, _P]),
    "usot_feature_size": (_I, [_I]),
    "usot_engine_template": (_I, [_P, _P, _I, _I, _P, _P, _P, _P]),
    "usot_engine_track": (_I, [_P, _P, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P, _P]),
    "usot_engine_extract_memory_feature": (_I, [_P, _P, _I, _I, _P, _I, _P, _P, _P]),
    "usot_engine_forward_train": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P]),
    "usot_tracker_postprocess": (_I, [_P, _P, _P, _P, _I, _I, _D, _D, _D, _D, _D, _P, _P]),
}


def header_symbols():
    """Every function the public header declares with USOT_API."""
    with open(HEADER_PATH) as f:
        src = f.read()
    return re.findall(r"USOT_API\s+[\w\s\*]+?\b(usot_\w+)\s*\(", src)


def load():
    """Load (once) and return the ctypes library.  Raises RuntimeError if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"usot_b200: {LIB_PATH} is missing -- build it with `python -m usot_b200.build` "
                    "(there is no CPU / PyTorch fallback for this path)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            # USOT_B200_TUNABLES="name=value,name=value": performance knobs (usot_set_tunable) applied once at load
            for kv in filter(None, os.environ.get("USOT_B200_TUNABLES", "").split(",")):
                name, _, val = kv.partition("=")
                if lib.usot_set_tunable(name.strip().encode(), int(val)) != 0:
                    raise RuntimeError((lib.usot_last_error() or b"usot_b200: bad USOT_B200_TUNABLES entry").decode())
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError((load().usot_last_error() or b"usot_b200: unknown error").decode())


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def profile_reset(on=False):
    check(load().usot_profile_reset(1 if on else 0))


def profile_count(counts):
    """Add {family name: launches} to the library's counters (replays of a caller-captured CUDA graph)."""
    lib = load()
    names = {lib.usot_profile_family_name(i).decode(): i for i in range(lib.usot_profile_family_count())}
    for k, n in counts.items():
        if n:
            check(lib.usot_profile_count(names[k], int(n)))


def profile_read():
    """{family: dict(launches, ms, flops, bytes)} since the last profile_reset (synchronises the device)."""
    lib = load()
    out = {}
    buf = (ctypes.c_double * 4)()
    for i in range(lib.usot_profile_family_count()):
        check(lib.usot_profile_read(i, buf))
        out[lib.usot_profile_family_name(i).decode()] = dict(launches=int(buf[0]), ms=buf[1], flops=buf[2], bytes=buf[3])
    return out
