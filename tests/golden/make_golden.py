"""Regenerate tests/golden/golden_small.npz from the REFERENCE itself (oracle/_ref, i.e. the unmodified
C files of /root/reference compiled by oracle/Makefile).  Run here (where /root/reference exists):

    python tests/golden/make_golden.py

The fixture travels to the GPU box; nothing at test time reads /root/reference.
Contents (all produced by reference functions, inputs seeded):
  mpg_ip / frames_ip      64x48, 5 frames, GOP 3 (I P P I P) stream  -> BGRA frames (ref decode loop)
  mpg_dense / frames_dense 64x48, 2 intra frames, full-range noise, all-ones quant tables
  mpg_flat / frames_flat  64x48, 2 intra frames, top 24 rows flat (zero-run adversarial case)
  coef / samp             1024 random int16 blocks (mixed ranges)    -> reference idct()
  ycc / bgra              3 x 256 random sample blocks              -> reference ycbcr_to_rgb()
  ystream_levels / ystream / ystream_coef   random sparse levels, their from-spec stream, reference
                          lossless_decode() output with Yquant
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mjpeg423_b200  # noqa: E402
from mjpeg423_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    R = oracle.ref()
    assert R is not None, "oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists"
    rng = np.random.default_rng(423)
    out = {}
    W, H = 64, 48
    fr = np.stack([synth.synth_frame(W, H, i, 16) for i in range(5)])
    out["mpg_ip"] = synth.encode_mpg(fr, gop=3)
    out["frames_ip"] = R.decode_mpg(out["mpg_ip"])
    ones = np.ones(64, dtype=np.int16)
    fr = np.stack([synth.synth_frame(W, H, i, 256) for i in range(2)])
    out["mpg_dense"] = synth.encode_mpg(fr, yq=ones, cq=ones)
    out["frames_dense"] = R.decode_mpg(out["mpg_dense"], yq=ones, cq=ones)
    fr = np.stack([synth.synth_frame(W, H, i, 16, flat_rows=24) for i in range(2)])
    out["mpg_flat"] = synth.encode_mpg(fr)
    out["frames_flat"] = R.decode_mpg(out["mpg_flat"])
    coef = np.concatenate([
        rng.integers(-32768, 32768, size=(256, 8, 8)),
        rng.integers(-2048, 2048, size=(256, 8, 8)),
        rng.integers(-300, 300, size=(256, 8, 8)) * (rng.random((256, 8, 8)) < 0.15),
        rng.integers(-64, 1500, size=(256, 8, 8)) * (np.arange(64).reshape(8, 8) == 0),
    ]).astype(np.int16)
    out["coef"] = coef
    out["samp"] = R.idct(coef)
    ycc = rng.integers(0, 256, size=(3, 256, 8, 8)).astype(np.uint8)
    out["ycc"] = ycc
    out["bgra"] = R.ycbcr_to_rgb(ycc[0], ycc[1], ycc[2], 128, 128)  # 16x16 blocks
    nb = 300
    lv = (rng.integers(-40, 41, size=(nb, 64)) * (rng.random((nb, 64)) < 0.12)).astype(np.int16)
    lv[:, 0] = rng.integers(-200, 200, size=nb)
    lv[7] = 0                      # an all-zero block (12 bits)
    lv[11, 63] = 5                 # a block that ends on zig-zag 63 (no END symbol)
    lv[12] = rng.integers(-2047, 2048, size=64)   # fully dense block
    stream = R.lossless_encode(lv)
    # the reference encoder zeroes the last partial byte; levels chosen so the tail is END
    out["ystream_levels"] = lv
    out["ystream"] = np.frombuffer(stream, dtype=np.uint8)
    out["ystream_coef"] = R.lossless_decode(nb, stream, oracle.YQUANT)
    path = os.path.join(ROOT, "tests", "golden", "golden_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
