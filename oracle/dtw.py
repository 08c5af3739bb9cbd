"""ORACLE / TEST INFRASTRUCTURE ONLY: ctypes front for dtw_oracle.c (built by
`make -C oracle`, or on demand here with gcc)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdtw_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dtw_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"] if force else ["make", "-C", _HERE, "-s"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.dtw_oracle_f32.restype = ctypes.c_long
        lib.dtw_oracle_f32.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_long, ctypes.c_long, i64p, i64p]
        lib.jump_frames_oracle.restype = ctypes.c_long
        lib.jump_frames_oracle.argtypes = [i64p, i64p, ctypes.c_long, i64p]
        _lib = lib
    return _lib


def dtw_path(x: np.ndarray):
    """x: (N, M) float32 cost.  Returns (text_indices, time_indices) int64, forward order."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, m = x.shape
    ti = np.empty(n + m, dtype=np.int64)
    tj = np.empty(n + m, dtype=np.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    length = _load().dtw_oracle_f32(
        x.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n, m, ti.ctypes.data_as(i64p), tj.ctypes.data_as(i64p)
    )
    if length < 0:
        raise MemoryError("dtw_oracle_f32")
    return ti[:length].copy(), tj[:length].copy()


def jump_frames(text_idx: np.ndarray, time_idx: np.ndarray) -> np.ndarray:
    text_idx = np.ascontiguousarray(text_idx, dtype=np.int64)
    time_idx = np.ascontiguousarray(time_idx, dtype=np.int64)
    out = np.empty(len(text_idx), dtype=np.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    k = _load().jump_frames_oracle(
        text_idx.ctypes.data_as(i64p), time_idx.ctypes.data_as(i64p), len(text_idx), out.ctypes.data_as(i64p)
    )
    return out[:k].copy()
