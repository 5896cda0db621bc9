"""P-frame stream decode timing: encode procedural 1080p frames with the GPU encoder (GOP 24), decode resident, verify.
    python tools/pframe_run.py [frames] [--no-verify]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, mjpeg423_b200
from mjpeg423_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 240
W, H = 1920, 1080
rng = np.random.default_rng(1)
base = bench.procedural_frames(W, H, 1, 1, 16)[0]
fr = np.empty((n, H, W, 4), np.uint8)
for f in range(n):                       # a static picture with a moving patch: P frames win
    fr[f] = base
    x0 = (f * 16) % (W - 256)
    fr[f, 400:656, x0:x0 + 256, :3] = rng.integers(0, 256, size=(256, 256, 3), dtype=np.uint8)
dec = mjpeg423_b200.Decoder(0)
mpg = dec.encode_frames(fr, 24, fix_tail=True)
info = mjpeg423_b200.probe(mpg)
print("frames", info.num_frames, "P frames", info.num_pframes, "bytes/frame", mpg.size // n)
dec.set_option(api.OPT_PROFILE, 1)
dec.upload(mpg)
d_out = dec.device_alloc(n * W * H * 4)
for _ in range(3):
    dec.decode_resident(d_out)
st = dec.stats()
if "--no-verify" not in sys.argv:                      # every frame against the reference's own decoder (64-bit checksums)
    from oracle import oracle
    want = api.frame_hash_host(oracle.best().decode_mpg(mpg, nthreads=os.cpu_count() or 1))
    got = dec.hash_frames(d_out, W * H * 4, n)
    assert np.array_equal(got, want), "P-frame decode differs from the oracle"
    print("verified", n, "frames against the oracle")
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
print("fps", n / st["total_ms"] * 1e3)
