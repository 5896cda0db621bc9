"""Small fixed workload for ncu: upload once, decode_resident twice (the second pass is the profiled one).

    python tools/profile_run.py [--workload 1080p] [--frames 256] [--staged 0|1] [--passes 2]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mjpeg423_b200  # noqa: E402
from mjpeg423_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="1080p")
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--staged", type=int, default=0)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--profile", type=int, default=1)
a = ap.parse_args()
geo = {"480p": (640, 480, 16, None), "1080p": (1920, 1080, 16, None), "4k": (3840, 2160, 256, None),
       "4k-q1": (3840, 2160, 256, np.ones(64, np.int16))}[a.workload]
W, H, amp, q = geo
mpg = synth.synth_mpg(W, H, a.frames, min(a.frames, 32), amp, 0, q, q)
dec = mjpeg423_b200.Decoder(0)
if q is not None:
    dec.set_quant(q, q)
dec.set_option(api.OPT_STAGED, a.staged)
dec.set_option(api.OPT_PROFILE, a.profile)
dec.upload(mpg)
d_out = dec.device_alloc(a.frames * W * H * 4)
for _ in range(a.passes):
    dec.decode_resident(d_out)
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in dec.stats().items()})
if a.profile:                      # the pipelined total (chunks overlapped on two streams), median of 5
    dec.set_option(api.OPT_PROFILE, 0)
    ms = []
    for _ in range(7):
        dec.decode_resident(d_out)
        ms.append(dec.stats()["total_ms"])
    print("pipelined_ms", round(sorted(ms)[3], 3), "chunks", len(ms) and dec.stats().get("kernel_launches"))
dec.device_free(d_out)
dec.close()
