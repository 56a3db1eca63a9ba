"""ctypes binding of librn_b200.so (C ABI declared in include/rn_b200.h).

The library is built in-tree by ``build()`` (``make -C csrc``: nvcc, sm_100a) and lives next to this
file so it travels with the source tree.  There is no fallback: if the library is missing or a call
fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librn_b200.so")
CSRC = os.path.join(_HERE, "csrc")

RN_ABI_VERSION = 3
PRECISION = {"fp32": 0, "parity": 1, "fast": 2}
MAX_G_LAYERS = 8


class RelationCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "n", "k", "Q", "G", "L", "qinj", "precision", "training")] + [("flags", C.c_uint32)]


class FCfg(C.Structure):
    _fields_ = [("B", C.c_int32), ("G", C.c_int32), ("F1", C.c_int32), ("F2", C.c_int32), ("A", C.c_int32),
                ("keep_scale", C.c_float)]


class ConvCfg(C.Structure):
    _fields_ = [("B", C.c_int32), ("side", C.c_int32), ("training", C.c_int32), ("eps", C.c_float),
                ("momentum", C.c_float), ("img_u8", C.c_int32), ("flags", C.c_int32)]


class ConvLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w", "bias", "gamma", "beta", "running_mean", "running_var")]


class ConvGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dw", "dbias", "dgamma", "dbeta")]


class LstmCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "T", "V", "E", "H", "training")]


class AdamCfg(C.Structure):
    _fields_ = [("n", C.c_int64), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("clip_norm", C.c_float), ("grad_scale", C.c_float), ("step", C.c_int32)]


# symbol -> (restype, argtypes); every symbol include/rn_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    "rn_abi_version": (C.c_int, []),
    "rn_last_error": (C.c_char_p, []),
    "rn_launch_count": (C.c_ulonglong, []),
    "rn_abi_struct_sizes": (C.c_int, [C.POINTER(C.c_int32), C.c_int]),
    "rn_device_check": (C.c_int, [C.c_int]),
    "rn_relation_tc_supported": (C.c_int, [C.POINTER(RelationCfg)]),
    "rn_relation_workspace": (C.c_int, [C.POINTER(RelationCfg), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "rn_relation_fwd": (C.c_int, [C.POINTER(RelationCfg), _P, _P, C.POINTER(_P), C.POINTER(_P), _P, _P, _P, _P]),
    "rn_relation_bwd": (C.c_int, [C.POINTER(RelationCfg), _P, _P, _P, C.POINTER(_P), _P, _P, _P, C.POINTER(_P),
                                  C.POINTER(_P), _P, _P]),
    "rn_extract_stats": (C.c_int, [_P, C.c_int, C.c_longlong, C.c_int, C.c_int, _P, _P, _P, _P]),
    "rn_relation_activation": (C.c_int, [C.POINTER(RelationCfg), _P, C.c_int, C.POINTER(_P)]),
    "rn_f_fwd": (C.c_int, [C.POINTER(FCfg)] + [_P] * 11),
    "rn_f_bwd": (C.c_int, [C.POINTER(FCfg)] + [_P] * 17),
    "rn_conv_workspace": (C.c_int, [C.POINTER(ConvCfg), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "rn_conv_fwd": (C.c_int, [C.POINTER(ConvCfg), _P, C.POINTER(ConvLayer), _P, _P, _P, _P]),
    "rn_conv_bwd": (C.c_int, [C.POINTER(ConvCfg), _P, _P, C.POINTER(ConvLayer), _P, C.POINTER(ConvGrads), _P, _P]),
    "rn_lstm_supported": (C.c_int, [C.POINTER(LstmCfg)]),
    "rn_lstm_workspace": (C.c_int, [C.POINTER(LstmCfg), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "rn_lstm_fwd": (C.c_int, [C.POINTER(LstmCfg)] + [_P] * 9),
    "rn_lstm_bwd": (C.c_int, [C.POINTER(LstmCfg)] + [_P] * 13),
    "rn_clip_adam": (C.c_int, [C.POINTER(AdamCfg), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()


def build(verbose: bool = False, jobs: int = 8) -> str:
    """Compile librn_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building librn_b200.so failed (see compiler output above)")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library.  Raises if it has not been built -- there is no CPU or eager fallback."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C relationnetworks_clevr_b200/csrc`.  This package has no fallback path.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)      # AttributeError if the symbol is not exported
                    fn.restype = res
                    fn.argtypes = args
                if handle.rn_abi_version() != RN_ABI_VERSION:
                    raise RuntimeError("librn_b200.so ABI version mismatch; rebuild it")
                mirrors = (RelationCfg, FCfg, ConvCfg, ConvLayer, ConvGrads, LstmCfg, AdamCfg)
                sizes = (C.c_int32 * len(mirrors))()
                n = handle.rn_abi_struct_sizes(sizes, len(mirrors))
                bad = [m.__name__ for m, sz in zip(mirrors, sizes) if C.sizeof(m) != sz]
                if n != len(mirrors) or bad:
                    raise RuntimeError(f"ctypes mirrors out of date with include/rn_b200.h: {bad or n}")
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().rn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr_array(tensors) -> C.Array:
    arr = (_P * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
