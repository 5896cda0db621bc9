"""mjpeg423_b200 -- B200-native MJPEG423 decode hot path (entropy decode -> dequantise -> 8x8 integer
IDCT -> YCbCr->RGB) behind the reference's own entry points.

The product is the C-ABI library `libmjpeg423_b200.so` (include/mjpeg423_b200.h; CUDA for sm_100a +
C++ host runtime).  This package is the thin Python host mirror used by tests and bench:

  api      numpy-level mirror of the reference interface (lossless_decode, idct, ycbcr_to_rgb,
           mjpeg423_decode) and the batched frame-range Decoder, all calling the C-ABI via ctypes
  synth    from-spec stream producer (host only)
  build    in-tree nvcc / g++ build

There is no CPU fallback: every decode entry raises if the CUDA library is missing or no GPU is present.
"""
from . import build, synth  # noqa: F401
from .api import (  # noqa: F401
    CQUANT, YQUANT, ZIGZAG, Decoder, Display, IFrameIndex, MpgInfo, decode_frames_multi, idct, lossless_decode, load_library, mjpeg423_decode, play, probe,
    write_bmp, ycbcr_to_rgb,
)
