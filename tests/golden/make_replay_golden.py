#!/usr/bin/env python3
"""Generates tests/golden/ref_intra_streams.json and ref_p_streams.json: I pictures and P-only sequences coded by the REFERENCE encoder (oracle/_ref/appencoder, needs /root/reference once)
at several presets / QPs, each with the MD5 of what the reference DECODER makes of it.  tests/test_replay.py re-creates those pictures from
the parsed decisions with the oracle's kernels (oracle/ora_replay.c) and must hit the same MD5 -- no reference binary needed at test time."""
import base64
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_yuv  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
nat = np.frombuffer(gzip.open(os.path.join(HERE, "nat_320x240_6f.yuv.gz"), "rb").read(), np.uint8)
CASES = [("nat320_veryfast_qp27", nat, 320, 240, "veryfast", 27, 0), ("nat320_slow_qp22", nat, 320, 240, "slow", 22, 0),
         ("nat320_ultrafast_qp37", nat, 320, 240, "ultrafast", 37, 3), ("nat320_placebo_qp32", nat, 320, 240, "placebo", 32, 5),
         ("syn200x120_medium_qp30", np.frombuffer(gen_yuv.make(200, 120, 1, seed=5), np.uint8), 200, 120, "medium", 30, 0),
         ("syn256x128_veryslow_qp18", np.frombuffer(gen_yuv.make(256, 128, 1, seed=9), np.uint8), 256, 128, "veryslow", 18, 0)]
# two textured crops of the reference repo's own 720p clip (read at generation time only)
big = "/root/reference/iOS_demo/resource/1280x720_15.yuv"
if os.path.exists(big):
    f0 = np.fromfile(big, np.uint8, count=1280 * 720 * 3 // 2)
    def crop(x, y, w, h):
        Y = f0[:1280 * 720].reshape(720, 1280)[y:y + h, x:x + w]
        U = f0[1280 * 720:1280 * 720 * 5 // 4].reshape(360, 640)[y // 2:(y + h) // 2, x // 2:(x + w) // 2]
        V = f0[1280 * 720 * 5 // 4:].reshape(360, 640)[y // 2:(y + h) // 2, x // 2:(x + w) // 2]
        return np.concatenate([Y.ravel(), U.ravel(), V.ravel()])
    CASES += [("crop720_veryfast_qp27", crop(448, 232, 384, 256), 384, 256, "veryfast", 27, 0), ("crop720_slow_qp24", crop(64, 360, 320, 192), 320, 192, "slow", 24, 0)]
out = []
for name, yuv, w, h, preset, qp, frame in CASES:
    fs = w * h * 3 // 2
    with tempfile.TemporaryDirectory() as d:
        clip, bs, dec = os.path.join(d, "i.yuv"), os.path.join(d, "o.265"), os.path.join(d, "d.yuv")
        open(clip, "wb").write(yuv[frame * fs:(frame + 1) * fs].tobytes())
        subprocess.run([os.path.join(REF, "appencoder"), "-i", clip, "-wdt", str(w), "-hgt", str(h), "-fr", "15", "-preset", preset, "-rc", "0", "-qp", str(qp),
                        "-iper", "128", "-frms", "1", "-threads", "1", "-b", bs], capture_output=True, check=True)
        subprocess.run([os.path.join(REF, "appdecoder"), "-b", bs, "-o", dec, "-threads", "1"], capture_output=True, check=True)
        stream, decoded = open(bs, "rb").read(), open(dec, "rb").read()
        assert len(decoded) == fs
        out.append({"name": name, "width": w, "height": h, "preset": preset, "qp": qp, "stream_b64": base64.b64encode(stream).decode(),
                    "decoded_md5": hashlib.md5(decoded).hexdigest()})
        print(name, len(stream), "bytes")
json.dump(out, open(os.path.join(HERE, "ref_intra_streams.json"), "w"), indent=0)

# P-only sequences (-bframes 0): every picture's MD5 as the reference decoder writes it (display order)
SEQS = [("nat320_veryfast_qp27_6f", nat, 320, 240, "veryfast", 27, 6), ("nat320_slow_qp32_6f", nat, 320, 240, "slow", 32, 6), ("nat320_ultrafast_qp22_6f", nat, 320, 240, "ultrafast", 22, 6)]
if os.path.exists(big):
    n = 8
    fr = np.fromfile(big, np.uint8, count=1280 * 720 * 3 // 2 * n).reshape(n, -1)
    def crop_seq(x, y, w, h):
        o = []
        for f in range(n):
            Y = fr[f][:1280 * 720].reshape(720, 1280)[y:y + h, x:x + w]
            U = fr[f][1280 * 720:1280 * 720 * 5 // 4].reshape(360, 640)[y // 2:(y + h) // 2, x // 2:(x + w) // 2]
            V = fr[f][1280 * 720 * 5 // 4:].reshape(360, 640)[y // 2:(y + h) // 2, x // 2:(x + w) // 2]
            o.append(np.concatenate([Y.ravel(), U.ravel(), V.ravel()]))
        return np.concatenate(o)
    SEQS += [("crop720_veryfast_qp27_8f", crop_seq(448, 232, 384, 256), 384, 256, "veryfast", 27, 8), ("crop720_medium_qp30_8f", crop_seq(640, 300, 320, 192), 320, 192, "medium", 30, 8)]
out = []
# ... and sequences with the reference's DEFAULT picture structure (hierarchical B pictures, decoding order != display order)
SEQS = [(a, b, c, d, e, f, g, ("-bframes", "0")) for a, b, c, d, e, f, g in SEQS]
SEQS += [("nat320_veryfast_qp27_6f_B", nat, 320, 240, "veryfast", 27, 6, ()), ("nat320_placebo_qp30_6f_B", nat, 320, 240, "placebo", 30, 6, ()),
         # rate-controlled: cu_qp_delta per CTB (the -rc 0 -qp arguments that follow are overridden by these)
         ("nat320_fast_crf26_6f_B", nat, 320, 240, "fast", 27, 6, ("-rc", "3", "-crf", "26")), ("nat320_veryfast_abr200_6f_B", nat, 320, 240, "veryfast", 27, 6, ("-rc", "2", "-br", "200"))]
if os.path.exists(big):
    SEQS += [("crop720_medium_qp27_8f_B", crop_seq(448, 232, 384, 256), 384, 256, "medium", 27, 8, ()), ("crop720_veryslow_qp32_8f_B", crop_seq(640, 300, 320, 192), 320, 192, "veryslow", 32, 8, ())]
for name, yuv, w, h, preset, qp, nf, extra in SEQS:
    fs = w * h * 3 // 2
    with tempfile.TemporaryDirectory() as d:
        clip, bs, dec = os.path.join(d, "i.yuv"), os.path.join(d, "o.265"), os.path.join(d, "d.yuv")
        open(clip, "wb").write(yuv[:nf * fs].tobytes())
        subprocess.run([os.path.join(REF, "appencoder"), "-i", clip, "-wdt", str(w), "-hgt", str(h), "-fr", "15", "-preset", preset, *(("-rc", "0", "-qp", str(qp)) if "-rc" not in extra else ()),
                        "-iper", "128", *extra, "-frms", str(nf), "-threads", "1", "-b", bs], capture_output=True, check=True)
        subprocess.run([os.path.join(REF, "appdecoder"), "-b", bs, "-o", dec, "-threads", "1"], capture_output=True, check=True)
        stream, decoded = open(bs, "rb").read(), open(dec, "rb").read()
        assert len(decoded) == fs * nf
        out.append({"name": name, "width": w, "height": h, "preset": preset, "qp": qp, "frames": nf, "stream_b64": base64.b64encode(stream).decode(),
                    "decoded_md5": [hashlib.md5(decoded[f * fs:(f + 1) * fs]).hexdigest() for f in range(nf)]})
        print(name, len(stream), "bytes")
json.dump(out, open(os.path.join(HERE, "ref_p_streams.json"), "w"), indent=0)
