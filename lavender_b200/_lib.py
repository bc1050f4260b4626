"""ctypes binding of liblavender_b200.so (the C ABI declared in include/lavender_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblavender_b200.so")

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


class LavError(RuntimeError):
    pass


class Dropout(ctypes.Structure):
    """LavDropout: device rng pointer ([seed, step] uint64), call-site id, drop probability."""
    _fields_ = [("rng", c_void_p), ("site", ctypes.c_uint32), ("p", c_float)]


class GemmEpilogue(ctypes.Structure):
    _fields_ = [
        ("out", c_void_p), ("ldo", c_int64), ("out_dtype", ctypes.c_int32), ("act", ctypes.c_int32),
        ("bias", c_void_p), ("aux", c_void_p), ("ldaux", c_int64), ("residual", c_void_p), ("ldres", c_int64),
        ("row_map", c_void_p), ("row_scale", c_void_p), ("rows_per_scale", ctypes.c_int32), ("alpha", c_float),
        ("accumulate", ctypes.c_int32), ("reserved", ctypes.c_int32), ("bias_grad", c_void_p), ("drop", Dropout),
    ]


MAJOR_K, MAJOR_MN = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_BWD = 0, 1, 2
OUT_F16, OUT_F32 = 0, 1
STORE, ACCUMULATE = 0, 1

_lib = None

# symbol -> (restype, argtypes); every function declared in include/lavender_b200.h must be listed here
# (tests/test_abi.py cross-checks the header against this table and the built library).
SIGNATURES = {
    "lav_abi_version": (c_int, []),
    "lav_last_error": (ctypes.c_char_p, []),
    "lav_launch_count": (c_int64, []),
    "lav_device_info": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "lav_debug_set_trace": (c_int, [c_void_p, c_int64]),
    "lav_frames_resize_crop_norm_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int, c_int, c_int, c_int, c_int,
                                               ctypes.POINTER(c_float), ctypes.POINTER(c_float), c_void_p, c_void_p]),
    "lav_gemm_f16": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                             ctypes.POINTER(GemmEpilogue), c_int, c_void_p]),
    "lav_layernorm_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p,
                                  c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "lav_layernorm_bwd": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_int, ctypes.POINTER(Dropout), c_void_p]),
    "lav_layernorm_bwd_ex": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                     c_void_p, c_int, c_void_p, c_int,
                                     c_void_p, c_void_p, c_void_p, c_int64, c_int, ctypes.POINTER(Dropout), c_void_p]),
    "lav_dropout_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, ctypes.POINTER(Dropout), c_void_p]),
    "lav_dropout_mask": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(Dropout), c_void_p]),
    "lav_scale_cast_f16": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int64, c_int,
                                   c_int, c_void_p]),
    "lav_cast_f32_to_f16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "lav_gelu_bwd_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "lav_colsum_f16": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_float, c_void_p]),
    "lav_attn_fwd_f16": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p,
                                 ctypes.POINTER(Dropout), c_void_p]),
    "lav_attn_bwd_f16": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p,
                                 c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int,
                                 ctypes.POINTER(Dropout), c_void_p]),
    "lav_attn_fwd_ex": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p,
                                c_int64, c_void_p, ctypes.POINTER(Dropout), c_void_p]),
    "lav_split3_f16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "lav_relpos_bias_expand": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_float,
                                       c_void_p]),
    "lav_relpos_bias_grad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "lav_bert_embed_ln_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "lav_bert_embed_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "lav_vid_embed_ln_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "lav_vid_embed_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lav_xent_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p]),
    "lav_grad_stats": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "lav_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_float,
                               c_float, c_float, c_float, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                               c_int, c_void_p]),
    "lav_xent_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int64, c_void_p, c_int64, c_void_p]),
}


def lib():
    """Loads the library on first use. Raises LavError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LavError(f"{LIB_PATH} not found - build it with `python -m lavender_b200.build` "
                           f"(there is no CPU fallback for the CUDA path)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().lav_last_error()
        raise LavError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count():
    return int(lib().lav_launch_count())
