"""ctypes binding of libhealnet_b200.so — the C ABI declared in include/healnet_b200.h.

The library is the product: there is no Python/PyTorch fallback. If it is missing (not built) or the
process has no CUDA device, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_long, c_size_t, c_ubyte, c_void_p

HN_MAX_MODALITIES = 16
HN_MAX_AXES = 4

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhealnet_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")


class hn_desc(ctypes.Structure):
    """Mirror of `struct hn_desc` (include/healnet_b200.h); field names follow HealNet.__init__."""

    _fields_ = [
        ("n_modalities", c_int),
        ("depth", c_int),
        ("l_c", c_int),
        ("l_d", c_int),
        ("x_heads", c_int),
        ("cross_dim_head", c_int),
        ("l_heads", c_int),
        ("latent_dim_head", c_int),
        ("num_freq_bands", c_int),
        ("out_dims", c_int),
        ("self_per_cross_attn", c_int),
        ("snn", c_int),
        ("final_classifier_head", c_int),
        ("fourier_encode_data", c_int),
        ("max_freq", c_float),
        ("channel_dims", c_int * HN_MAX_MODALITIES),
        ("num_spatial_axes", c_int * HN_MAX_MODALITIES),
    ]


# name -> (restype, argtypes); every symbol include/healnet_b200.h declares
SIGNATURES = {
    "hn_create": (c_int, [POINTER(hn_desc), POINTER(c_void_p)]),
    "hn_destroy": (c_int, [c_void_p]),
    "hn_last_error": (c_char_p, []),
    "hn_set_weights": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p), c_int]),
    "hn_pack_weights": (c_int, [c_void_p, c_void_p]),
    "hn_workspace_bytes": (c_size_t, [c_void_p, c_int, POINTER(c_int)]),
    "hn_forward": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), c_void_p, c_long,
                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hn_forward_ex": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int),
                              c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hn_exchange_bytes": (c_size_t, [c_void_p, c_int]),
    "hn_exchange_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "hn_exchange_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "hn_exchange_close": (c_int, [c_void_p]),
    "hn_exchange_free": (c_int, [c_void_p]),
    "hn_set_exchange": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p), c_size_t]),
    "hn_exchange_error": (c_int, [c_void_p, POINTER(c_int)]),
    "hn_exchange_error_async": (c_int, [c_void_p, c_void_p, c_void_p]),
    "hn_set_exchange_timeout": (c_int, [c_void_p, ctypes.c_double]),
    "hn_workspace_bytes_split": (c_size_t, [c_void_p, c_int, POINTER(c_int), POINTER(c_long)]),
    "hn_forward_split": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_long),
                                 POINTER(c_long), POINTER(c_int), c_void_p, c_long, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "hn_tape_bytes": (c_size_t, [c_void_p, c_int, POINTER(c_int)]),
    "hn_backward_scratch_bytes": (c_size_t, [c_void_p, c_int, POINTER(c_int)]),
    "hn_forward_train": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int),
                                 c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "hn_set_grads": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p), c_int]),
    "hn_set_backward_variant": (c_int, [c_void_p, c_int]),
    "hn_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t,
                            c_void_p]),
    "hn_set_io_dtype": (c_int, [c_void_p, c_int]),
    "hn_last_launch_count": (c_int, [c_void_p]),
    "hn_set_attention_export": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "hn_profile_enable": (c_int, [c_void_p, c_int]),
    "hn_profile_read": (c_int, [c_void_p, c_int, c_int, POINTER(c_float), POINTER(c_int), POINTER(ctypes.c_double),
                                POINTER(ctypes.c_double), POINTER(ctypes.c_double)]),
    "hn_attention_workspace_bytes": (c_size_t, [c_int, c_int, c_long, c_int, c_int, c_int, c_int]),
    "hn_attention_forward": (c_int, [c_int, c_int, c_long, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "hn_attention_forward_cached": (c_int, [c_int, c_int, c_long, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_size_t, c_int, c_void_p]),
    "hn_op_gemm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                           c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hn_op_layernorm_f16": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_long,
                                    c_int, c_void_p]),
    "hn_op_build_context": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int,
                                    c_float, c_int, c_void_p, c_void_p]),
    "hn_op_attention_nsplit": (c_int, [c_int, c_int, c_int, c_long, c_int]),
    "hn_op_attention": (c_int, [c_void_p, c_int, c_void_p, c_long, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_long, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hn_op_combine": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                              c_void_p, c_void_p, c_int, c_void_p]),
}

_lib = None


class HealNetLibraryError(RuntimeError):
    pass


def build_library(verbose: bool = False) -> str:
    """Compiles healnet_b200/csrc/*.cu for sm_100a into libhealnet_b200.so (in-tree). Needs nvcc, not a GPU."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise HealNetLibraryError("building libhealnet_b200.so failed (see output above)")
    return LIB_PATH


def load_library() -> ctypes.CDLL:
    """Loads the shared library and types every exported entry point. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HealNetLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C {CSRC_DIR}`. healnet_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library ever drift apart
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    msg = load_library().hn_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise HealNetLibraryError(f"{what} failed (code {rc}): {last_error()}")
