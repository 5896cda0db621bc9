"""ctypes front-end of the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under mjpeg423-video-decoder-software_b200/ does.

  port()  -> Checker over oracle/libmjpeg423_oracle.so   (the restatement, mjpeg423_oracle.c)
  ref()   -> Checker over oracle/_ref/libmjpeg423_ref.so (the reference's own C files, built by
             `make -C oracle ref` where /root/reference exists) or None when that file is absent.
Both expose the same methods, mirroring the reference entry points
(LIB/decoder/mjpeg423_decoder.h:14-17) on numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libmjpeg423_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmjpeg423_ref.so")

YQUANT = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
                   14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                   18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                   49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int16)
CQUANT = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                   24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32, dtype=np.int16)


def build(ref: bool = True) -> None:
    """Compile the checkers (gcc). `make ref` is a no-op message when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, "port"] + (["ref"] if ref else []), check=True)


def _padded(bitstream) -> np.ndarray:
    b = np.frombuffer(bytes(bitstream), dtype=np.uint8) if not isinstance(bitstream, np.ndarray) else bitstream
    out = np.zeros(b.size + 16, dtype=np.uint8)  # reference look-ahead
    out[:b.size] = b
    return out


class Checker:
    def __init__(self, path: str, is_ref: bool):
        self.lib = C.CDLL(path)
        self.is_ref = is_ref
        self.kind = "reference" if is_ref else "port"
        L = self.lib
        p = C.c_void_p
        if is_ref:
            L.lossless_decode.argtypes = [C.c_int, p, p, p, C.c_int]
            L.lossless_decode.restype = None
            L.idct.argtypes = [p, p]
            L.ycbcr_to_rgb.argtypes = [C.c_int, C.c_int, C.c_uint32, p, p, p, p]
            L.lossless_encode.argtypes = [C.c_int, p, p]
            L.lossless_encode.restype = C.c_uint32
            L.fdct.argtypes = [p, p]
            L.rgb_to_ycbcr.argtypes = [C.c_int, C.c_int, C.c_uint32, p, p, p, p]
            L.quantize_I.argtypes = [p, p, p, p, p]
            L.mjpeg423_decode.argtypes = [C.c_char_p, C.c_char_p]
            self._ld, self._idct, self._col = L.lossless_decode, L.idct, L.ycbcr_to_rgb
            self._frame, self._mpg = L.ref_decode_frame, L.ref_decode_mpg
        else:
            L.orc_lossless_decode.argtypes = [C.c_int, p, p, p, C.c_int]
            L.orc_lossless_decode.restype = C.c_uint64
            L.orc_idct.argtypes = [p, p]
            L.orc_ycbcr_to_rgb.argtypes = [C.c_int, C.c_int, C.c_uint32, p, p, p, p]
            self._ld, self._idct, self._col = L.orc_lossless_decode, L.orc_idct, L.orc_ycbcr_to_rgb
            self._frame, self._mpg = L.orc_decode_frame, L.orc_decode_mpg
        self._frame.argtypes = [C.c_uint32, C.c_uint32, p, C.c_uint32, C.c_uint32, C.c_int, p, p, p, p, p, p]
        self._frame.restype = None
        self._mpg.argtypes = [p, C.c_size_t, C.c_uint32, C.c_uint32, p, p, p, C.c_int, p]
        self._mpg.restype = C.c_int
        self._enc = L.ref_encode_mpg if is_ref else L.orc_encode_mpg
        self._enc.argtypes = [p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, p, p, C.c_int, p, C.c_size_t, p]
        self._enc.restype = C.c_int
        if is_ref:
            L.ref_encode_mpg_files.argtypes = [p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, p,
                                               C.c_size_t, p]
            L.ref_encode_mpg_files.restype = C.c_int

    # -- stage functions -------------------------------------------------------------------------
    def lossless_decode(self, num_blocks: int, bitstream, quant, P: int = 0, DCACq: np.ndarray | None = None):
        """-> (num_blocks, 8, 8) int16 dequantised coefficients. DCACq is the in/out state for P frames."""
        bs = _padded(bitstream)
        q = np.ascontiguousarray(np.asarray(quant, dtype=np.int16).reshape(64))
        if DCACq is None:
            DCACq = np.zeros((num_blocks, 8, 8), dtype=np.int16)
        assert DCACq.dtype == np.int16 and DCACq.flags.c_contiguous and DCACq.size == num_blocks * 64
        self._ld(num_blocks, bs.ctypes.data, DCACq.ctypes.data, q.ctypes.data, int(P))
        return DCACq

    def idct(self, coef: np.ndarray) -> np.ndarray:
        """(n, 8, 8) int16 -> (n, 8, 8) uint8, one reference idct() call per block."""
        coef = np.ascontiguousarray(coef, dtype=np.int16).reshape(-1, 8, 8)
        out = np.empty(coef.shape, dtype=np.uint8)
        cp, op = coef.ctypes.data, out.ctypes.data
        for b in range(coef.shape[0]):
            self._idct(cp + 128 * b, op + 64 * b)
        return out

    def ycbcr_to_rgb(self, Y: np.ndarray, Cb: np.ndarray, Cr: np.ndarray, W: int, H: int) -> np.ndarray:
        """Block-major (nb, 8, 8) uint8 planes -> (H, W, 4) BGRA raster, one call per block."""
        Y, Cb, Cr = (np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 8, 8) for a in (Y, Cb, Cr))
        rgb = np.zeros((H, W, 4), dtype=np.uint8)
        wb = W // 8
        for b in range(Y.shape[0]):
            self._col((b // wb) * 8, (b % wb) * 8, W, Y.ctypes.data + 64 * b, Cb.ctypes.data + 64 * b,
                      Cr.ctypes.data + 64 * b, rgb.ctypes.data)
        return rgb

    # -- frame / container level (harness.c) ------------------------------------------------------
    def decode_mpg(self, mpg: np.ndarray, first: int = 0, n: int | None = None, yq=None, cq=None,
                   nthreads: int = 1, stage_secs: np.ndarray | None = None) -> np.ndarray:
        mpg = _padded(mpg)
        nframes, W, H = (int(x) for x in mpg[:12].view("<u4"))
        if n is None:
            n = nframes - first
        out = np.empty((n, H, W, 4), dtype=np.uint8)
        yqa = None if yq is None else np.ascontiguousarray(np.asarray(yq, dtype=np.int16).reshape(64))
        cqa = None if cq is None else np.ascontiguousarray(np.asarray(cq, dtype=np.int16).reshape(64))
        sp = None if stage_secs is None else stage_secs.ctypes.data
        rc = self._mpg(mpg.ctypes.data, mpg.size - 16, first, n, None if yqa is None else yqa.ctypes.data,
                       None if cqa is None else cqa.ctypes.data, out.ctypes.data, nthreads, sp)
        if rc != 0:
            raise RuntimeError(f"{self.kind} decode_mpg failed rc={rc}")
        return out

    # -- encoder: the frame loop of LIB/encoder/mjpeg423_encoder.c:97-225 -------------
    @staticmethod
    def encode_bound(n: int, W: int, H: int) -> int:
        return 20 + n * (16 + 3 * (W // 8) * (H // 8) * 160 + 8 + 8) + 512

    def encode_mpg(self, frames: np.ndarray, max_I_interval: int = 1, yq=None, cq=None, fix_tail: bool = False) -> np.ndarray:
        """frames: (n, H, W, 4) uint8 BGRA -> the .mpg bytes (the last 512 are the zeroed pad)."""
        fr = np.ascontiguousarray(frames, dtype=np.uint8)
        n, H, W, _ = fr.shape
        out = np.zeros(self.encode_bound(n, W, H), dtype=np.uint8)
        ln = C.c_size_t(0)
        yqa = None if yq is None else np.ascontiguousarray(yq, dtype=np.int16)
        cqa = None if cq is None else np.ascontiguousarray(cq, dtype=np.int16)
        rc = self._enc(fr.ctypes.data, n, W, H, max_I_interval, None if yqa is None else yqa.ctypes.data,
                       None if cqa is None else cqa.ctypes.data, int(fix_tail), out.ctypes.data, out.size, C.byref(ln))
        if rc:
            raise RuntimeError(f"{self.kind} encode_mpg failed rc={rc}")
        return out[:ln.value].copy()

    def encode_mpg_files(self, frames: np.ndarray, max_I_interval: int, tmpdir: str) -> np.ndarray:
        """The reference's own mjpeg423_encode() through BMP files in tmpdir (ref only).  The last 512 bytes of
        the result are uninitialised stack of the reference: compare [:-512]."""
        assert self.is_ref
        fr = np.ascontiguousarray(frames, dtype=np.uint8)
        n, H, W, _ = fr.shape
        out = np.zeros(self.encode_bound(n, W, H), dtype=np.uint8)
        ln = C.c_size_t(0)
        rc = self.lib.ref_encode_mpg_files(fr.ctypes.data, n, W, H, max_I_interval, tmpdir.encode(), out.ctypes.data,
                                           out.size, C.byref(ln))
        if rc:
            raise RuntimeError(f"ref_encode_mpg_files failed rc={rc}")
        return out[:ln.value].copy()

    # -- reference encoder pieces (ref only; used to cross-check the from-spec encoder) ------------
    def lossless_encode(self, levels: np.ndarray) -> bytes:
        assert self.is_ref
        lv = np.ascontiguousarray(levels, dtype=np.int16).reshape(-1, 64)
        buf = np.zeros(lv.shape[0] * 160 + 16, dtype=np.uint8)
        n = self.lib.lossless_encode(lv.shape[0], lv.ctypes.data, buf.ctypes.data)
        return bytes(buf[:n])

    def fdct(self, block: np.ndarray) -> np.ndarray:
        assert self.is_ref
        b = np.ascontiguousarray(block, dtype=np.uint8).reshape(8, 8)
        out = np.zeros((8, 8), dtype=np.int16)
        self.lib.fdct(b.ctypes.data, out.ctypes.data)
        return out


_cache: dict[str, Checker | None] = {}


def port() -> Checker:
    if "port" not in _cache:
        if not os.path.exists(PORT_SO):
            build(ref=False)
        _cache["port"] = Checker(PORT_SO, False)
    return _cache["port"]


def ref() -> Checker | None:
    if "ref" not in _cache:
        if not os.path.exists(REF_SO) and os.path.isdir("/root/reference"):
            build(ref=True)
        _cache["ref"] = Checker(REF_SO, True) if os.path.exists(REF_SO) else None
    return _cache["ref"]


def best() -> Checker:
    """The strongest checker available: the compiled reference if present, else the port."""
    return ref() or port()
