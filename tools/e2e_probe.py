"""End-to-end call (pinned host .mpg -> pinned host frames) against the pipeline chunk size: wall and event time per
call.  python tools/e2e_probe.py   (compare with tools/pcie_probe.py on the same box)"""
import sys, time, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mjpeg423_b200
from mjpeg423_b200 import api, synth
W,H=1920,1080
n=517
mpg=synth.synth_mpg(W,H,n,32,16,0,None,None)
dec=mjpeg423_b200.Decoder(0)
pin_in=dec.pinned(mpg.size+64); pin_in.array[:mpg.size]=mpg
pin_out=dec.pinned(n*W*H*4)
for chunk in (0, 16, 31, 124):
    dec.set_option(api.OPT_CHUNK_FRAMES, chunk)
    dec.decode_frames(pin_in.array[:mpg.size],0,n,out=pin_out)
    ws=[]; es=[]
    for _ in range(4):
        t0=time.perf_counter(); dec.decode_frames(pin_in.array[:mpg.size],0,n,out=pin_out); ws.append((time.perf_counter()-t0)*1e3); es.append(dec.stats()['total_ms'])
    print('chunk',chunk,'wall ms',[round(x,2) for x in ws],'event ms',[round(x,2) for x in es], 'fps', round(n/min(ws)*1e3))
