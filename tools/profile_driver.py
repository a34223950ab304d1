#!/usr/bin/env python3
"""Small fixed workload for ncu: one stream encodes a short 4K GOP (1 IDR + N-1 P) through the public API.
usage: profile_driver.py [nframes] [width height]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import gen_yuv
import ks265codec_b200 as ks

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160)
yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=1234), np.uint8)
cfg = ks.default_config(w, h, preset="veryfast", qp=27, iper=128, psnr=1)
with ks.Encoder(cfg) as e:
    e.set_profiling(True)
    bs, rec, st = e.encode_gop(yuv)
    print("frames", n, "bytes", bs.size, "launches", st.gpu_launches, "stage ms/launch",
          {k: round(v[0] / max(1, v[1]), 3) for k, v in e.stage_times().items()})
