"""ctypes binding of libwmd_b200.so (the C ABI declared in include/wmd_b200.h).

The library is the product; there is no Python or CPU implementation behind these calls.  If the
shared object is missing it is built in-tree with nvcc (``build.py``); if that fails, or the
library cannot be loaded, importing callers get a loud ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_handle = ctypes.c_void_p
IPC_HANDLE_BYTES = 64                                    # WMD_IPC_HANDLE_BYTES
MAX_FANOUT = 7                                           # WMD_MAX_FANOUT

# name -> (restype, argtypes); mirrors include/wmd_b200.h one to one
SIGNATURES = {
    "wmd_create": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                                  ctypes.c_int32, ctypes.POINTER(c_handle)]),
    "wmd_destroy": (ctypes.c_int, [c_handle]),
    "wmd_last_error": (ctypes.c_char_p, []),
    "wmd_version": (ctypes.c_char_p, []),
    "wmd_set_token_map": (ctypes.c_int, [c_handle, c_i32p, ctypes.c_int64]),
    "wmd_set_rank": (ctypes.c_int, [c_handle, c_i32p, ctypes.c_int64]),
    "wmd_get_table": (ctypes.c_int, [c_handle, c_f32p]),
    "wmd_pairs_host": (ctypes.c_int, [c_handle, c_i32p, c_i64p, c_i32p, c_i64p, ctypes.c_int64, ctypes.c_int32,
                                      c_f64p, c_i32p]),
    "wmd_pairs_host_in_dev_out": (ctypes.c_int, [c_handle, c_i32p, c_i64p, c_i32p, c_i64p, ctypes.c_int64, ctypes.c_int32,
                                                 ctypes.c_void_p, ctypes.c_void_p]),
    "wmd_pairs_submit": (ctypes.c_int, [c_handle, c_i32p, c_i64p, c_i32p, c_i64p, ctypes.c_int64, ctypes.c_int32]),
    "wmd_pairs_wait": (ctypes.c_int, [c_handle, c_f64p, c_i32p]),
    "wmd_workspace_bytes": (ctypes.c_int, [c_handle, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, c_i64p, c_i64p]),
    "wmd_nbow_dev": (ctypes.c_int, [c_handle, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "wmd_rwmd_pairs_dev": (ctypes.c_int, [c_handle, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "wmd_pairs_dev": (ctypes.c_int, [c_handle, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                     ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "wmd_pairs_padded_dev": (ctypes.c_int, [c_handle, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                            ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "wmd_nbow_host": (ctypes.c_int, [c_handle, c_i32p, c_i64p, ctypes.c_int64, c_i32p, c_i32p, c_f64p, c_i32p]),
    "wmd_rwmd_pairs_host": (ctypes.c_int, [c_handle, c_i32p, c_i64p, c_i32p, c_i64p, ctypes.c_int64,
                                           c_f64p, c_f64p, c_f64p, c_i32p, c_i32p, c_i32p]),
    "wmd_allpairs_topk_host": (ctypes.c_int, [c_handle, c_i32p, c_i64p, ctypes.c_int64, c_i32p, c_i64p, ctypes.c_int64,
                                              ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, c_i32p, c_f64p, c_i64p, c_f64p]),
    "wmd_allpairs_topk_dev": (ctypes.c_int, [c_handle, c_i32p, c_i64p, ctypes.c_int64, c_i32p, c_i64p, ctypes.c_int64,
                                             ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, c_i64p, c_f64p]),
    "wmd_emd_batch_host": (ctypes.c_int, [c_handle, c_f64p, c_f64p, c_f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_double, c_f64p]),
    "wmd_peer_alloc": (ctypes.c_int, [c_handle, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    "wmd_peer_open": (ctypes.c_int, [c_handle, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "wmd_peer_close": (ctypes.c_int, [c_handle, ctypes.c_void_p, ctypes.c_int32]),
    "wmd_set_fanout": (ctypes.c_int, [c_handle, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)]),
    "wmd_set_profiling": (ctypes.c_int, [c_handle, ctypes.c_int32]),
    "wmd_set_serial": (ctypes.c_int, [c_handle, ctypes.c_int32]),
    "wmd_set_distance_table": (ctypes.c_int, [c_handle, ctypes.c_int32]),
    "wmd_distance_table_info": (ctypes.c_int, [c_handle, c_i64p, c_f64p, c_i32p, c_i32p]),
    "wmd_get_profile": (ctypes.c_int, [c_handle, c_f64p, c_i64p, ctypes.c_int32]),
    "wmd_get_last_stats": (ctypes.c_int, [c_handle, c_i64p]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load libwmd_b200.so, declaring every prototype.  Raises RuntimeError loudly on any failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and _build.is_stale():
        try:
            _build.build()
        except Exception as exc:                       # stale-but-present library is still usable
            if not os.path.exists(path):
                raise RuntimeError(f"libwmd_b200.so is missing and could not be built: {exc}") from exc
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -m consistent__style_transfer_b200.build`; "
                           "there is no CPU fallback")
    try:
        L = ctypes.CDLL(path)
    except OSError as exc:
        raise RuntimeError(f"cannot load {path}: {exc}; there is no CPU fallback") from exc
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)                          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = load().wmd_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libwmd_b200 error {rc}: {msg}")
