"""ctypes front-end of the from-spec MJPEG423 stream producer (csrc/synth_encoder.cpp).

Host-only helper for tests and bench: makes procedural frames and complete in-memory .mpg files.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.LIB_SYNTH
        if not os.path.exists(path):
            _build.build_synth()
        lib = C.CDLL(path)
        lib.mjpeg423_synth_frame.argtypes = [C.c_uint32] * 5 + [C.c_void_p]
        lib.mjpeg423_synth_frame.restype = None
        lib.mjpeg423_encode_mpg.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
        lib.mjpeg423_encode_mpg.restype = C.c_size_t
        lib.mjpeg423_synth_mpg.argtypes = [C.c_uint32] * 6 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        lib.mjpeg423_synth_mpg.restype = C.c_size_t
        _lib = lib
    return _lib


def _qptr(q):
    if q is None:
        return None, None
    a = np.ascontiguousarray(np.asarray(q, dtype=np.int16).reshape(64))
    return a, a.ctypes.data


def synth_frame(W: int, H: int, frame_index: int = 0, amp: int = 16, flat_rows: int = 0) -> np.ndarray:
    """Procedural BGRA picture, shape (H, W, 4) uint8."""
    out = np.empty((H, W, 4), dtype=np.uint8)
    _load().mjpeg423_synth_frame(W, H, frame_index, amp, flat_rows, out.ctypes.data)
    return out


def encode_mpg(frames: np.ndarray, yq=None, cq=None, gop: int = 1) -> np.ndarray:
    """Encode (n, H, W, 4) BGRA frames into a .mpg byte array. gop > 1 makes P frames."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, H, W, _ = frames.shape
    ya, yp = _qptr(yq)
    ca, cp = _qptr(cq)
    lib = _load()
    size = lib.mjpeg423_encode_mpg(W, H, n, frames.ctypes.data, yp, cp, gop, None, 0)
    if size == 0:
        raise ValueError("bad geometry (W, H must be non-zero multiples of 8)")
    out = np.zeros(size + 64, dtype=np.uint8)  # slack keeps the decoder look-ahead inside the buffer
    got = lib.mjpeg423_encode_mpg(W, H, n, frames.ctypes.data, yp, cp, gop, out.ctypes.data, size)
    assert got == size
    return out[:size]


def synth_mpg(W: int, H: int, n: int, n_unique: int = 0, amp: int = 16, flat_rows: int = 0, yq=None, cq=None,
              nthreads: int | None = None) -> np.ndarray:
    """Intra-only .mpg of n frames cycling n_unique procedural pictures (bench-scale generator)."""
    ya, yp = _qptr(yq)
    ca, cp = _qptr(cq)
    lib = _load()
    if nthreads is None:
        nthreads = max(1, min(os.cpu_count() or 1, 64))
    size = lib.mjpeg423_synth_mpg(W, H, n, n_unique, amp, flat_rows, yp, cp, None, 0, nthreads)
    if size == 0:
        raise ValueError("stream does not fit the container (32-bit offsets) or bad geometry")
    out = np.zeros(size + 64, dtype=np.uint8)
    got = lib.mjpeg423_synth_mpg(W, H, n, n_unique, amp, flat_rows, yp, cp, out.ctypes.data, size, nthreads)
    assert got == size
    return out[:size]
