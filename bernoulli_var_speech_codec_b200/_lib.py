"""ctypes binding of libbvc.so (the C ABI declared in include/bvc.h).

There is no fallback: if the shared library is missing this module raises, and every
compute entry point of the package goes through it.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BVC_LIBRARY") or os.path.join(PKG, "libbvc.so")   # BVC_LIBRARY: another build of the same ABI (A/B timing)

BVC_OK = 0
STATUS_NAMES = {0: "BVC_OK", -1: "BVC_ERR_INVALID", -2: "BVC_ERR_SCHEMA", -3: "BVC_ERR_DEVICE",
                -4: "BVC_ERR_STATE", -5: "BVC_ERR_NOMEM"}


class BvcConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("x_dim", C.c_int32), ("h_dim", C.c_int32), ("z_dim", C.c_int32),
        ("var_bit", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32), ("pad_left", C.c_int32),
        ("voc_initial_channel", C.c_int32), ("voc_num_stages", C.c_int32),
        ("voc_up_rates", C.c_int32 * 4), ("voc_up_kernels", C.c_int32 * 4),
        ("voc_num_kernels", C.c_int32), ("voc_res_kernels", C.c_int32 * 3), ("voc_res_dilations", C.c_int32 * 3),
        ("voc_antialias", C.c_int32 * 4), ("voc_antialias_post", C.c_int32),
    ]


class BvcTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


# every symbol include/bvc.h declares: name -> (restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int32, C.c_float
SYMBOLS = {
    "bvc_abi_version": (C.c_int, []),
    "bvc_last_error": (C.c_char_p, []),
    "bvc_create": (C.c_int, [C.POINTER(_P), C.POINTER(BvcConfig)]),
    "bvc_destroy": (C.c_int, [_P]),
    "bvc_load_bvrnn": (C.c_int, [_P, C.POINTER(BvcTensor), _I]),
    "bvc_load_vocoder": (C.c_int, [_P, C.POINTER(BvcTensor), _I]),
    "bvc_set_frontend": (C.c_int, [_P, _P, _P]),
    "bvc_logmel": (C.c_int, [_P, _P, _I, _I, _F, _P, _P]),
    "bvc_encode": (C.c_int, [_P, _P, _P, _F, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "bvc_encode_mel": (C.c_int, [_P, _P, _P, _F, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "bvc_encode_ex": (C.c_int, [_P, _P, _P, _F, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "bvc_unpack_codes": (C.c_int, [_P, _P, _P, _F, _I, _I, _P, _P]),
    "bvc_decode_mel": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "bvc_decode_packed": (C.c_int, [_P, _P, _P, _F, _P, _I, _I, _P, _P, _P]),
    "bvc_bitstream_bytes": (C.c_size_t, [_P, _I, _I]),
    "bvc_pack_bitstream": (C.c_int, [_P, _P, _P, _F, _I, _I, _P, C.c_size_t, _P]),
    "bvc_unpack_bitstream": (C.c_int, [_P, _P, C.c_size_t, _I, _I, _P, _P, _P]),
    "bvc_stream_create": (C.c_int, [_P, _I, C.POINTER(_P)]),
    "bvc_stream_destroy": (C.c_int, [_P]),
    "bvc_stream_reset": (C.c_int, [_P, _P, _P]),
    "bvc_stream_encode_step": (C.c_int, [_P, _P, _P, _P, _F, _F, _P, _P, _P]),
    "bvc_stream_decode_step": (C.c_int, [_P, _P, _P, _P, _F, _F, _P, _P]),
    "bvc_vocode": (C.c_int, [_P, _P, _I, _I, _I, _F, _P, _P]),
    "bvc_vocoder_out_len": (C.c_int64, [_P, _I]),
    "bvc_encode_host": (C.c_int, [_P, _P, _I, _I, _F, _F, _P]),
    "bvc_decode_host": (C.c_int, [_P, _P, _I, _I, _I, _F, _P]),
    "bvc_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "bvc_host_free": (C.c_int, [_P]),
    "bvc_workspace_bytes": (C.c_size_t, [_P]),
    "bvc_kernel_launches": (C.c_int64, [_P]),
    "bvc_set_precision": (C.c_int, [_P, _I]),
    "bvc_last_recurrent_ms": (C.c_float, [_P]),
    "bvc_recurrent_ms": (C.c_float, [_P, _I, _I]),
    "bvc_check": (C.c_int, [_P]),
    "bvc_debug_read": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t]),
}

_lib = None


def load():
    """Loads libbvc.so and binds every symbol; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m bernoulli_var_speech_codec_b200.build` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != BVC_OK:
        msg = load().bvc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what or 'libbvc'}: {STATUS_NAMES.get(rc, rc)}: {msg}")
