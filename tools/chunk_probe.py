"""Resident 1080p decode time against the pipeline chunk size: python tools/chunk_probe.py [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mjpeg423_b200
from mjpeg423_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
W, H = 1920, 1080
mpg = synth.synth_mpg(W, H, n, 32, 16, 0, None, None)
dec = mjpeg423_b200.Decoder(0)
dec.upload(mpg)
d_out = dec.device_alloc(n * W * H * 4)
for chunk in (0, 250, 400, 500, 667, 1000, 2000):
    dec.set_option(api.OPT_CHUNK_FRAMES, chunk)
    try:
        for _ in range(2):
            dec.decode_resident(d_out)
        ms = []
        for _ in range(5):
            dec.decode_resident(d_out)
            ms.append(dec.stats()["total_ms"])
        print("chunk", chunk, "ms", round(min(ms), 3), round(sorted(ms)[2], 3), "fps", round(n / sorted(ms)[2] * 1e3))
    except Exception as e:
        print("chunk", chunk, "failed", str(e)[:100])
