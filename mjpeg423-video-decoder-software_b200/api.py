"""Python host mirror of the reference decoder interface, calling the C-ABI of libmjpeg423_b200.so.

The functions keep the reference's names and argument meaning
(LIB/decoder/mjpeg423_decoder.h:14-17, LIB = /root/reference/core0/software/common/libs/mjpeg423):

    lossless_decode(num_blocks, bitstream, DCACq, quant, P)
    idct(DCAC, block)                       (also whole planes: (n, 8, 8) arrays)
    ycbcr_to_rgb(h, w, w_size, Y, Cb, Cr, rgbblock)
    mjpeg423_decode(filename_in, filenamebase_out)

plus `Decoder`, the batched frame-range API (C group 3).  Everything goes through ctypes into the CUDA
library; there is no Python or CPU implementation of the arithmetic here.  If the library is missing,
cannot be loaded, or no CUDA device is present, calls raise `RuntimeError`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import build as _build

YQUANT = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
                   14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                   18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                   49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int16).reshape(8, 8)
CQUANT = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                   24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32, dtype=np.int16).reshape(8, 8)
ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                   41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63], dtype=np.int32)

OPT_PROFILE, OPT_STAGED, OPT_CHUNK_FRAMES, OPT_VALIDATE = 1, 2, 3, 4
E_ARG, E_FORMAT, E_CUDA, E_NOMEM, E_STREAM, E_PFRAME = -1, -2, -3, -4, -5, -6

# every symbol include/mjpeg423_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "lossless_decode", "idct", "ycbcr_to_rgb", "mjpeg423_decode", "Yquant", "Cquant", "zigzag_table",
    "mjpeg423_b200_set_read_limit", "mjpeg423_b200_get_read_limit", "mjpeg423_b200_idct_blocks",
    "mjpeg423_b200_ycbcr_to_rgb_frame", "mjpeg423_b200_lossless_decode",
    "init_idct_ycbcr_to_rgb_accel", "idct_accel_calculate_buffer_y", "idct_accel_calculate_buffer_cb",
    "idct_accel_calculate_buffer_cr", "ycbcr_to_rgb_accel_get_results", "wait_for_ycbcr_to_rgb_finsh",
    "wait_for_idct_y_finsh", "mjpeg423_b200_accel_set_geometry", "ycbcr_to_rgb_accel_calculate_buffer",
    "mjpeg423_b200_index", "mjpeg423_b200_seek_iframe", "mjpeg423_b200_fast_forward", "mjpeg423_b200_rewind",
    "mjpeg423_b200_decode_frames_multi",
    "mjpeg423_b200_create", "mjpeg423_b200_destroy", "mjpeg423_b200_set_option", "mjpeg423_b200_last_error",
    "mjpeg423_b200_probe", "mjpeg423_b200_set_quant", "mjpeg423_b200_decode_frames", "mjpeg423_b200_upload",
    "mjpeg423_b200_decode_resident", "mjpeg423_b200_get_stats", "mjpeg423_b200_resident_entropy",
    "mjpeg423_b200_resident_idct", "mjpeg423_b200_resident_colour", "mjpeg423_b200_resident_idct_colour",
    "mjpeg423_b200_host_alloc", "mjpeg423_b200_host_free", "mjpeg423_b200_device_alloc",
    "mjpeg423_b200_device_free", "mjpeg423_b200_memcpy_d2h", "mjpeg423_b200_memcpy_h2d", "mjpeg423_b200_sync",
    "mjpeg423_b200_device_count", "mjpeg423_b200_hash_frames",
    "mjpeg423_b200_encode_bound", "mjpeg423_b200_encode_frames",
    "mjpeg423_b200_display_init", "mjpeg423_b200_display_free", "mjpeg423_b200_display_register_written_buffer",
    "mjpeg423_b200_display_buffer_is_available", "mjpeg423_b200_display_switch_frames", "mjpeg423_b200_display_get_buffer",
    "mjpeg423_b200_display_get_displayed_buffer", "mjpeg423_b200_display_clear_screen", "mjpeg423_b200_display_num_buffers",
    "mjpeg423_b200_play", "mjpeg423_b200_write_bmp",
]


DISPLAY_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_void_p)     # on_display(user, frame_index, frame)


class _Info(C.Structure):
    _fields_ = [("num_frames", C.c_uint32), ("w_size", C.c_uint32), ("h_size", C.c_uint32),
                ("num_iframes", C.c_uint32), ("payload_size", C.c_uint32), ("num_pframes", C.c_uint32),
                ("frame_bytes", C.c_uint64), ("max_frame_payload", C.c_uint64)]


class _Stats(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("sync_ms", C.c_float), ("chain_ms", C.c_float), ("index_ms", C.c_float),
                ("decode_ms", C.c_float), ("idct_colour_ms", C.c_float), ("idct_ms", C.c_float),
                ("colour_ms", C.c_float), ("kernel_launches", C.c_uint64), ("payload_bytes", C.c_uint64),
                ("segments", C.c_uint64), ("fixups", C.c_uint64), ("frames", C.c_uint64), ("list_entries", C.c_uint64)]


@dataclass
class MpgInfo:
    num_frames: int
    w_size: int
    h_size: int
    num_iframes: int
    payload_size: int
    num_pframes: int
    frame_bytes: int
    max_frame_payload: int


_lib = None


def load_library(build_if_missing: bool = False) -> C.CDLL:
    """Load libmjpeg423_b200.so (in-tree). Raises RuntimeError if it is absent: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_CUDA
    if not os.path.exists(path):
        if build_if_missing:
            _build.build_cuda()
        else:
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the MJPEG423 decode path has no CPU fallback)")
    lib = C.CDLL(path)
    p, u32, i32, sz, u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t, C.c_uint64
    sig = {
        "lossless_decode": (None, [i32, p, p, p, i32]),
        "idct": (None, [p, p]),
        "ycbcr_to_rgb": (None, [i32, i32, u32, p, p, p, p]),
        "mjpeg423_decode": (None, [C.c_char_p, C.c_char_p]),
        "mjpeg423_b200_set_read_limit": (None, [sz]),
        "mjpeg423_b200_get_read_limit": (sz, []),
        "mjpeg423_b200_idct_blocks": (i32, [p, p, sz]),
        "mjpeg423_b200_ycbcr_to_rgb_frame": (i32, [p, p, p, u32, u32, p]),
        "mjpeg423_b200_lossless_decode": (i32, [i32, p, sz, p, p, i32]),
        "init_idct_ycbcr_to_rgb_accel": (i32, []),
        "idct_accel_calculate_buffer_y": (None, [p, u32]),
        "idct_accel_calculate_buffer_cb": (None, [p, u32]),
        "idct_accel_calculate_buffer_cr": (None, [p, u32]),
        "ycbcr_to_rgb_accel_get_results": (None, [p, u32]),
        "wait_for_ycbcr_to_rgb_finsh": (None, []),
        "wait_for_idct_y_finsh": (None, []),
        "mjpeg423_b200_accel_set_geometry": (i32, [u32, u32]),
        "ycbcr_to_rgb_accel_calculate_buffer": (None, [p, p, p, p, i32, i32, i32]),
        "mjpeg423_b200_index": (i32, [p, sz, p, u32, C.POINTER(u32), C.POINTER(i32)]),
        "mjpeg423_b200_seek_iframe": (i32, [p, u32, u32, i32]),
        "mjpeg423_b200_fast_forward": (i32, [p, u32, u32, u32]),
        "mjpeg423_b200_rewind": (i32, [p, u32, u32]),
        "mjpeg423_b200_decode_frames_multi": (i32, [p, i32, p, u32, u64, u64, p, p]),
        "mjpeg423_b200_create": (i32, [C.POINTER(p), i32]),
        "mjpeg423_b200_destroy": (None, [p]),
        "mjpeg423_b200_set_option": (i32, [p, i32, C.c_int64]),
        "mjpeg423_b200_last_error": (C.c_char_p, []),
        "mjpeg423_b200_probe": (i32, [p, sz, C.POINTER(_Info)]),
        "mjpeg423_b200_set_quant": (i32, [p, p, p]),
        "mjpeg423_b200_decode_frames": (i32, [p, p, sz, u32, u32, p, i32]),
        "mjpeg423_b200_upload": (i32, [p, p, sz, u32, u32]),
        "mjpeg423_b200_decode_resident": (i32, [p, p]),
        "mjpeg423_b200_get_stats": (i32, [p, C.POINTER(_Stats)]),
        "mjpeg423_b200_resident_entropy": (i32, [p, p]),
        "mjpeg423_b200_resident_idct": (i32, [p, p, p]),
        "mjpeg423_b200_resident_colour": (i32, [p, p, p]),
        "mjpeg423_b200_resident_idct_colour": (i32, [p, p, p]),
        "mjpeg423_b200_host_alloc": (p, [sz]),
        "mjpeg423_b200_host_free": (None, [p]),
        "mjpeg423_b200_device_alloc": (p, [p, sz]),
        "mjpeg423_b200_device_free": (None, [p, p]),
        "mjpeg423_b200_memcpy_d2h": (i32, [p, p, p, sz]),
        "mjpeg423_b200_memcpy_h2d": (i32, [p, p, p, sz]),
        "mjpeg423_b200_sync": (i32, [p]),
        "mjpeg423_b200_device_count": (i32, []),
        "mjpeg423_b200_hash_frames": (i32, [p, p, u64, u32, p]),
        "mjpeg423_b200_encode_bound": (sz, [u32, u32, u32]),
        "mjpeg423_b200_encode_frames": (i32, [p, p, i32, u32, u32, u32, u32, u32, p, sz, C.POINTER(sz)]),
        "mjpeg423_b200_display_init": (p, [i32, i32, i32]),
        "mjpeg423_b200_display_free": (None, [p]),
        "mjpeg423_b200_display_register_written_buffer": (None, [p]),
        "mjpeg423_b200_display_buffer_is_available": (i32, [p]),
        "mjpeg423_b200_display_switch_frames": (i32, [p]),
        "mjpeg423_b200_display_get_buffer": (p, [p]),
        "mjpeg423_b200_display_get_displayed_buffer": (p, [p]),
        "mjpeg423_b200_display_clear_screen": (None, [p, C.c_char]),
        "mjpeg423_b200_display_num_buffers": (i32, [p]),
        "mjpeg423_b200_play": (C.c_long, [p, p, sz, u32, u32, p, u32, DISPLAY_CB, p, C.POINTER(u32)]),
        "mjpeg423_b200_write_bmp": (i32, [C.c_char_p, p, u32, u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _err(lib) -> str:
    return lib.mjpeg423_b200_last_error().decode("utf-8", "replace")


def _check(lib, rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {_err(lib)}")


def _require_gpu(lib) -> None:
    if lib.mjpeg423_b200_device_count() <= 0:
        raise RuntimeError("no CUDA device visible: the MJPEG423 B200 decode path has no CPU fallback")


def _bytes_arr(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf.reshape(-1).view(np.uint8))
    return np.frombuffer(bytes(buf), dtype=np.uint8)


# ---- reference interface --------------------------------------------------------------------------------
def lossless_decode(num_blocks: int, bitstream, DCACq: np.ndarray | None, quant, P: int = 0) -> np.ndarray:
    """Entropy decode + dequantise one plane (LIB/decoder/lossless_decode.c:60-135) on the GPU.

    bitstream: bytes-like, the plane's stream (its exact length is passed to the library).
    DCACq:     (num_blocks, 8, 8) int16, in/out (P frames accumulate into it); None allocates zeros.
    quant:     64 x int16, natural order.
    """
    lib = load_library()
    _require_gpu(lib)
    bs = _bytes_arr(bitstream)
    if DCACq is None:
        DCACq = np.zeros((num_blocks, 8, 8), dtype=np.int16)
    if DCACq.dtype != np.int16 or not DCACq.flags.c_contiguous or DCACq.size != num_blocks * 64:
        raise ValueError("DCACq must be a C-contiguous int16 array of num_blocks*64 elements")
    q = np.ascontiguousarray(np.asarray(quant, dtype=np.int16).reshape(64))
    src = np.zeros(bs.size + 8, dtype=np.uint8)
    src[:bs.size] = bs
    rc = lib.mjpeg423_b200_lossless_decode(num_blocks, src.ctypes.data, bs.size, DCACq.ctypes.data, q.ctypes.data, int(P))
    _check(lib, rc, "lossless_decode")
    return DCACq


def idct(DCAC: np.ndarray, block: np.ndarray | None = None) -> np.ndarray:
    """8x8 integer IDCT (LIB/decoder/idct.c:22-181) of one block (8, 8) or many (n, 8, 8) on the GPU."""
    lib = load_library()
    _require_gpu(lib)
    coef = np.ascontiguousarray(DCAC, dtype=np.int16)
    n = coef.size // 64
    if block is None:
        block = np.empty(coef.shape, dtype=np.uint8)
    if block.dtype != np.uint8 or not block.flags.c_contiguous or block.size != coef.size:
        raise ValueError("block must be a C-contiguous uint8 array of the same element count")
    _check(lib, lib.mjpeg423_b200_idct_blocks(coef.ctypes.data, block.ctypes.data, n), "idct")
    return block


def ycbcr_to_rgb(h: int, w: int, w_size: int, Y: np.ndarray, Cb: np.ndarray, Cr: np.ndarray,
                 rgbblock: np.ndarray) -> np.ndarray:
    """Colour-convert one 8x8 block into the BGRA raster `rgbblock` ((H, w_size, 4) uint8) at row h,
    column w (LIB/decoder/ycbcr_to_rgb.c:26-49), through the exported reference symbol."""
    lib = load_library()
    _require_gpu(lib)
    Y, Cb, Cr = (np.ascontiguousarray(a, dtype=np.uint8).reshape(8, 8) for a in (Y, Cb, Cr))
    assert rgbblock.dtype == np.uint8 and rgbblock.flags.c_contiguous
    lib.ycbcr_to_rgb(h, w, w_size, Y.ctypes.data, Cb.ctypes.data, Cr.ctypes.data, rgbblock.ctypes.data)
    return rgbblock


def ycbcr_to_rgb_frame(Y: np.ndarray, Cb: np.ndarray, Cr: np.ndarray, w_size: int, h_size: int) -> np.ndarray:
    """Whole frame: block-major (nb, 8, 8) planes -> (h_size, w_size, 4) BGRA."""
    lib = load_library()
    _require_gpu(lib)
    Y, Cb, Cr = (np.ascontiguousarray(a, dtype=np.uint8) for a in (Y, Cb, Cr))
    out = np.empty((h_size, w_size, 4), dtype=np.uint8)
    rc = lib.mjpeg423_b200_ycbcr_to_rgb_frame(Y.ctypes.data, Cb.ctypes.data, Cr.ctypes.data, w_size, h_size, out.ctypes.data)
    _check(lib, rc, "ycbcr_to_rgb_frame")
    return out


def mjpeg423_decode(filename_in: str, filenamebase_out: str) -> None:
    """File-level decoder (LIB/decoder/mjpeg423_decoder.c:20-149): .mpg in, one BMP per frame out."""
    lib = load_library()
    _require_gpu(lib)
    lib.mjpeg423_decode(os.fsencode(filename_in), os.fsencode(filenamebase_out))


def probe(mpg) -> MpgInfo:
    """Parse the container header and frame chain of an in-memory .mpg (host only; no GPU needed)."""
    lib = load_library()
    a = _bytes_arr(mpg)
    info = _Info()
    _check(lib, lib.mjpeg423_b200_probe(a.ctypes.data, a.size, C.byref(info)), "probe")
    return MpgInfo(*(getattr(info, f[0]) for f in _Info._fields_))


# ---- batched frame-range decoder --------------------------------------------------------------------------
class _Shard(C.Structure):
    _fields_ = [("mpg", C.c_void_p), ("len", C.c_size_t)]


class IFrameIndex:
    """The I-frame index of a .mpg: `entries` is an (n, 2) uint32 array of {frame_index,
    frame_position}, the layout of the reference's iframe_trailer_t (LIB/common/mjpeg423_types.h:22-25);
    trailer_ok says whether the file's own trailer agrees with the header walk.  Host only: no GPU needed."""

    def __init__(self, mpg):
        lib = load_library()
        a = _bytes_arr(mpg)
        n, ok = C.c_uint32(0), C.c_int(0)
        _check(lib, lib.mjpeg423_b200_index(a.ctypes.data, a.size, None, 0, C.byref(n), C.byref(ok)), "index")
        self.entries = np.zeros((n.value, 2), dtype=np.uint32)
        _check(lib, lib.mjpeg423_b200_index(a.ctypes.data, a.size, self.entries.ctypes.data, n.value, C.byref(n), C.byref(ok)), "index")
        self.trailer_ok = bool(ok.value)
        self._lib = lib

    def __len__(self):
        return len(self.entries)

    def seek(self, frame: int, direction: int = 0) -> int:
        return self._lib.mjpeg423_b200_seek_iframe(self.entries.ctypes.data, len(self.entries), frame, direction)

    def fast_forward(self, num_frames: int, current: int) -> int:
        return self._lib.mjpeg423_b200_fast_forward(self.entries.ctypes.data, len(self.entries), num_frames, current)

    def rewind(self, current: int) -> int:
        return self._lib.mjpeg423_b200_rewind(self.entries.ctypes.data, len(self.entries), current)


def decode_frames_multi(files, devices, first: int = 0, n: int | None = None, out: np.ndarray | None = None):
    """Frame-range sharding over the GPUs of one box: `files` is one .mpg or a list of them forming one
    logical stream, `devices` a list of device ordinals (a device may appear twice).  Returns (frames, cuts)."""
    lib = load_library()
    _require_gpu(lib)
    if isinstance(files, (bytes, bytearray, np.ndarray)):
        files = [files]
    arrs = [_bytes_arr(f) for f in files]
    infos = [probe(a) for a in arrs]
    total = sum(i.num_frames for i in infos)
    if n is None:
        n = total - first
    shards = (_Shard * len(arrs))(*[_Shard(a.ctypes.data, a.size) for a in arrs])
    devs = (C.c_int * len(devices))(*devices)
    cuts = np.zeros(len(devices) + 1, dtype=np.uint64)
    if out is None:
        out = np.empty((n, infos[0].h_size, infos[0].w_size, 4), dtype=np.uint8)
    rc = lib.mjpeg423_b200_decode_frames_multi(devs, len(devices), shards, len(arrs), first, n, out.ctypes.data, cuts.ctypes.data)
    _check(lib, rc, "decode_frames_multi")
    return out, cuts


class PinnedBuffer:
    """Pinned host memory from the library, viewed as a numpy array."""

    def __init__(self, lib, nbytes: int):
        self._lib = lib
        self.ptr = lib.mjpeg423_b200_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError(f"cannot pin {nbytes} bytes of host memory")
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.mjpeg423_b200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        self.free()


class Decoder:
    """Batched MJPEG423 decoder on one GPU (C group 3 of include/mjpeg423_b200.h)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        _require_gpu(self.lib)
        h = C.c_void_p()
        _check(self.lib, self.lib.mjpeg423_b200_create(C.byref(h), device), "create")
        self.h = h
        self.device = device
        self._mpg = None
        self.info: MpgInfo | None = None
        self.n = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.mjpeg423_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, opt: int, value: int) -> None:
        _check(self.lib, self.lib.mjpeg423_b200_set_option(self.h, opt, value), "set_option")

    def set_quant(self, yq=None, cq=None) -> None:
        ya = None if yq is None else np.ascontiguousarray(np.asarray(yq, dtype=np.int16).reshape(64))
        ca = None if cq is None else np.ascontiguousarray(np.asarray(cq, dtype=np.int16).reshape(64))
        _check(self.lib, self.lib.mjpeg423_b200_set_quant(self.h, None if ya is None else ya.ctypes.data,
                                                         None if ca is None else ca.ctypes.data), "set_quant")

    def pinned(self, nbytes: int) -> PinnedBuffer:
        return PinnedBuffer(self.lib, nbytes)

    # -- end to end: host .mpg -> host frames -------------------------------------------------------------
    def decode_frames(self, mpg, first: int = 0, n: int | None = None, out=None) -> np.ndarray:
        """Decode frames [first, first+n) into (n, H, W, 4) BGRA. `out`: PinnedBuffer / numpy array / None."""
        a = _bytes_arr(mpg)
        info = probe(a)
        if n is None:
            n = info.num_frames - first
        nbytes = n * info.frame_bytes
        if out is None:
            out = np.empty(nbytes, dtype=np.uint8)
        arr = out.array if isinstance(out, PinnedBuffer) else out
        if arr.nbytes < nbytes:
            raise ValueError("output buffer too small")
        rc = self.lib.mjpeg423_b200_decode_frames(self.h, a.ctypes.data, a.size, first, n, arr.ctypes.data, 0)
        _check(self.lib, rc, "decode_frames")
        return arr.reshape(-1)[:nbytes].reshape(n, info.h_size, info.w_size, 4)

    # -- encoder: frames -> .mpg, the loop of LIB/encoder/mjpeg423_encoder.c:97-225 ----------
    def encode_frames(self, frames, max_I_interval: int = 1, fix_tail: bool = False, d_frames: int = 0,
                      shape: tuple | None = None, out=None) -> np.ndarray:
        """frames: (n, H, W, 4) uint8 BGRA on the host -- or d_frames = device pointer with shape = (n, H, W).
        `out`: optional PinnedBuffer / uint8 array that receives the file (pinned memory gives full PCIe speed; it must
        hold the result, mjpeg423_b200_encode_bound() is always enough).  Returns the .mpg bytes (byte-identical to
        the reference encoder up to its last 512 bytes)."""
        if d_frames:
            n, H, W = shape
            src, on_dev = d_frames, 1
        else:
            fr = np.ascontiguousarray(frames, dtype=np.uint8)
            n, H, W, _ = fr.shape
            src, on_dev = fr.ctypes.data, 0
        if out is None:
            out = np.empty(self.lib.mjpeg423_b200_encode_bound(n, W, H), dtype=np.uint8)
        elif isinstance(out, PinnedBuffer):
            out = out.array
        ln = C.c_size_t(0)
        rc = self.lib.mjpeg423_b200_encode_frames(self.h, src, on_dev, n, W, H, max_I_interval, 1 if fix_tail else 0,
                                                  out.ctypes.data, out.size, C.byref(ln))
        _check(self.lib, rc, "encode_frames")
        return out[:ln.value]

    def decode_frames_to_device(self, mpg, d_out: int, first: int = 0, n: int | None = None) -> None:
        a = _bytes_arr(mpg)
        if n is None:
            n = probe(a).num_frames - first
        _check(self.lib, self.lib.mjpeg423_b200_decode_frames(self.h, a.ctypes.data, a.size, first, n, d_out, 1),
               "decode_frames")

    # -- device resident ------------------------------------------------------------------------------------
    def upload(self, mpg, first: int = 0, n: int | None = None) -> MpgInfo:
        a = _bytes_arr(mpg)
        self.info = probe(a)
        if n is None:
            n = self.info.num_frames - first
        _check(self.lib, self.lib.mjpeg423_b200_upload(self.h, a.ctypes.data, a.size, first, n), "upload")
        self.n = n
        return self.info

    def decode_resident(self, d_out: int) -> None:
        _check(self.lib, self.lib.mjpeg423_b200_decode_resident(self.h, d_out), "decode_resident")

    def resident_entropy(self, d_coef: int) -> None:
        _check(self.lib, self.lib.mjpeg423_b200_resident_entropy(self.h, d_coef), "resident_entropy")

    def resident_idct(self, d_coef: int, d_samples: int) -> None:
        _check(self.lib, self.lib.mjpeg423_b200_resident_idct(self.h, d_coef, d_samples), "resident_idct")

    def resident_colour(self, d_samples: int, d_out: int) -> None:
        _check(self.lib, self.lib.mjpeg423_b200_resident_colour(self.h, d_samples, d_out), "resident_colour")

    def resident_idct_colour(self, d_coef: int, d_out: int) -> None:
        _check(self.lib, self.lib.mjpeg423_b200_resident_idct_colour(self.h, d_coef, d_out), "resident_idct_colour")

    def stats(self) -> dict:
        s = _Stats()
        _check(self.lib, self.lib.mjpeg423_b200_get_stats(self.h, C.byref(s)), "get_stats")
        return {f[0]: getattr(s, f[0]) for f in _Stats._fields_}

    # -- memory ---------------------------------------------------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = self.lib.mjpeg423_b200_device_alloc(self.h, nbytes)
        if not p:
            raise MemoryError(f"cannot allocate {nbytes} bytes on device {self.device}")
        return p

    def device_free(self, p: int) -> None:
        self.lib.mjpeg423_b200_device_free(self.h, p)

    def to_host(self, d_src: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        _check(self.lib, self.lib.mjpeg423_b200_memcpy_d2h(self.h, out.ctypes.data, d_src, nbytes), "memcpy_d2h")
        return out.view(dtype)

    def to_device(self, d_dst: int, arr: np.ndarray) -> None:
        a = np.ascontiguousarray(arr)
        _check(self.lib, self.lib.mjpeg423_b200_memcpy_h2d(self.h, d_dst, a.ctypes.data, a.nbytes), "memcpy_h2d")

    def sync(self) -> None:
        _check(self.lib, self.lib.mjpeg423_b200_sync(self.h), "sync")

    def hash_frames(self, d_frames: int, frame_bytes: int, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint64)
        _check(self.lib, self.lib.mjpeg423_b200_hash_frames(self.h, d_frames, frame_bytes, n, out.ctypes.data), "hash_frames")
        return out


def frame_hash_host(frames: np.ndarray) -> np.ndarray:
    """numpy twin of the library's per-frame checksum (k_hash_frames): frames (n, ...) uint8 -> (n,) uint64."""
    n = frames.shape[0]
    w = np.ascontiguousarray(frames).reshape(n, -1).view(np.uint64)
    idx = (np.arange(1, w.shape[1] + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    with np.errstate(over="ignore"):
        z = w ^ idx[None, :]
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return z.sum(axis=1, dtype=np.uint64)


# ---- display side: the reference's frame ring, ece423_vid_ctl.h:67-77 ---------------------------
class Display:
    """N-buffer frame ring (pinned host memory when a CUDA device exists, plain memory otherwise)."""

    def __init__(self, width: int, height: int, num_buffers: int = 4):
        self.lib = load_library()
        self.h = self.lib.mjpeg423_b200_display_init(width, height, num_buffers)
        if not self.h:
            raise RuntimeError("display_init failed: " + _err(self.lib))
        self.width, self.height = width, height
        self.num_buffers = self.lib.mjpeg423_b200_display_num_buffers(self.h)

    def _view(self, ptr: int) -> np.ndarray:
        buf = (C.c_uint8 * (self.width * self.height * 4)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.uint8).reshape(self.height, self.width, 4)

    def register_written_buffer(self) -> None:
        self.lib.mjpeg423_b200_display_register_written_buffer(self.h)

    def buffer_is_available(self) -> int:
        return self.lib.mjpeg423_b200_display_buffer_is_available(self.h)

    def switch_frames(self) -> int:
        return self.lib.mjpeg423_b200_display_switch_frames(self.h)

    def get_buffer(self) -> np.ndarray:
        return self._view(self.lib.mjpeg423_b200_display_get_buffer(self.h))

    def get_displayed_buffer(self) -> np.ndarray:
        return self._view(self.lib.mjpeg423_b200_display_get_displayed_buffer(self.h))

    def clear_screen(self, color: int = 0) -> None:
        self.lib.mjpeg423_b200_display_clear_screen(self.h, bytes([color & 255]))

    def close(self) -> None:
        if self.h:
            self.lib.mjpeg423_b200_display_free(self.h)
            self.h = None


def play(dec: "Decoder", mpg, display: Display, first: int = 0, n: int | None = None, frame_period_us: int = 0,
         on_display=None) -> tuple[int, int]:
    """C0/playback.c's loop: decode frames [first, first+n) through the ring; on_display(frame_index, frame ndarray)
    is called for every displayed frame.  Returns (frames displayed, timer ticks without a new frame)."""
    a = _bytes_arr(mpg)
    if n is None:
        n = probe(a).num_frames - first
    H, W = display.height, display.width

    def cb(_user, idx, ptr):
        if on_display is not None:
            buf = (C.c_uint8 * (W * H * 4)).from_address(ptr)
            on_display(idx, np.frombuffer(buf, dtype=np.uint8).reshape(H, W, 4))
    ccb = DISPLAY_CB(cb)
    dropped = C.c_uint32(0)
    rc = dec.lib.mjpeg423_b200_play(dec.h, a.ctypes.data, a.size, first, n, display.h, frame_period_us, ccb, None, C.byref(dropped))
    if rc < 0:
        _check(dec.lib, int(rc), "play")
    return int(rc), int(dropped.value)


def write_bmp(path: str, frame: np.ndarray) -> None:
    """encode_bmp(), LIB/libbmp/encode_bmp.c: 32-bpp BMP of one (H, W, 4) BGRA frame (host-only function)."""
    lib = load_library()
    fr = np.ascontiguousarray(frame, dtype=np.uint8)
    _check(lib, lib.mjpeg423_b200_write_bmp(os.fsencode(path), fr.ctypes.data, fr.shape[1], fr.shape[0]), "write_bmp")
