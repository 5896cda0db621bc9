"""Encoder timing: python tools/enc_run.py [--frames 64] [--w 1920 --h 1080] [--amp 16] [--max-i 24]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mjpeg423_b200
ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--amp", type=int, default=16)
ap.add_argument("--max-i", type=int, default=24)
ap.add_argument("--passes", type=int, default=3)
a = ap.parse_args()
rng = np.random.default_rng(0)
y, x = np.mgrid[0:a.h, 0:a.w]
fr = np.zeros((a.frames, a.h, a.w, 4), np.uint8)
for f in range(a.frames):
    base = np.stack([(255 * x // a.w + f) & 255, (255 * y // a.h + f) & 255, (255 * (x + y) // (a.w + a.h) + f) & 255], -1)
    fr[f, ..., :3] = (base + rng.integers(0, a.amp, size=(a.h, a.w, 3))) & 255
dec = mjpeg423_b200.Decoder(0)
d = dec.device_alloc(fr.nbytes)
dec.to_device(d, fr)
pin = dec.pinned(a.frames * (2 << 20))
for i in range(a.passes):
    t0 = time.time()
    mpg = dec.encode_frames(None, a.max_i, d_frames=d, shape=(a.frames, a.h, a.w), out=pin)
    t1 = time.time()
    st = dec.stats()
    print(f"pass {i}: device-resident frames: {st['total_ms']:.2f} ms device, {1e3*(t1-t0):.1f} ms wall, {a.frames/(t1-t0):.0f} fps, "
          f"{mpg.size/a.frames/1e3:.1f} KB/frame, P frames {mjpeg423_b200.probe(mpg).num_pframes}")
t0 = time.time(); mpg = dec.encode_frames(fr, a.max_i); t1 = time.time()
print(f"host frames: {1e3*(t1-t0):.1f} ms wall, {a.frames/(t1-t0):.0f} fps")
