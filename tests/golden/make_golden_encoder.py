"""Regenerate tests/golden/golden_encoder.npz from the REFERENCE ENCODER itself: the real mjpeg423_encode()
(LIB/encoder/mjpeg423_encoder.c:20-230) of oracle/_ref, run end to end through BMP files written by the reference's
own encode_bmp().  Run here (where /root/reference exists):

    python tests/golden/make_golden_encoder.py

Contents:  enc_frames (6, 32, 48, 4) BGRA input frames; enc_mpg_gop24 / enc_mpg_gop1 / enc_mpg_gop3 = the reference's
.mpg files for max_I_interval 24 / 1 / 3 WITHOUT their last 512 bytes (uninitialised stack, mjpeg423_encoder.c:219-220).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


def main():
    R = oracle.ref()
    assert R is not None, "oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists"
    rng = np.random.default_rng(4230)
    n, H, W = 6, 32, 48
    y, x = np.mgrid[0:H, 0:W]
    fr = np.zeros((n, H, W, 4), np.uint8)
    for f in range(n):
        base = np.stack([(5 * x + f) & 255, (7 * y + 2 * f) & 255, (3 * (x + y)) & 255], -1)
        fr[f, ..., :3] = (base + rng.integers(0, 8, size=(H, W, 3))) & 255
    fr[3, 8:24, 8:40, :3] = rng.integers(0, 256, size=(16, 32, 3))        # a busy patch: this frame prefers I
    out = {"enc_frames": fr}
    for gop in (24, 1, 3):
        with tempfile.TemporaryDirectory() as d:
            out[f"enc_mpg_gop{gop}"] = R.encode_mpg_files(fr, gop, d)[:-512]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_encoder.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
