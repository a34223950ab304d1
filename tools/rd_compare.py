#!/usr/bin/env python3
"""Rate/PSNR of the reference encoder vs this repo's CPU model (== the CUDA path bit for bit) on one I420 clip -- reproduces the P3 table of
DESIGN.md section 2.  Needs /root/reference (or oracle/_ref) for the reference binary; runs on CPU only.
usage: rd_compare.py clip.yuv width height frames [fps] [qp ...]"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from katlib import SeqCfg, oracle, ptr

REF = os.path.join(ROOT, "oracle", "_ref", "appencoder")


def psnr_planes(a, b, w, h, n):
    fs, out = w * h * 3 // 2, []
    for off, sz in ((0, w * h), (w * h, w * h // 4), (w * h * 5 // 4, w * h // 4)):
        sse = sum(float(((a[f * fs + off:f * fs + off + sz].astype(np.int32) - b[f * fs + off:f * fs + off + sz].astype(np.int32)) ** 2).sum()) for f in range(n))
        out.append(10 * np.log10(255.0 ** 2 * sz * n / max(sse, 1e-9)))
    return out


def ours(yuv, w, h, n, qp, fps, bframes=0, satd=0, iters=16, sao=3, me=0):
    O = oracle(); O.ora_encode_sequence.restype = C.c_long
    cfg = SeqCfg(w, h, n, qp, 128, 0, 64, iters, 2, 1, sao, 3, satd, bframes, me, 0, 2400)
    bs = np.zeros(w * h * 3 * n + 100000, np.uint8); rec = np.zeros(w * h * 3 // 2 * n, np.uint8)
    nb = O.ora_encode_sequence(C.byref(cfg), ptr(yuv), ptr(bs), C.c_size_t(bs.size), ptr(rec))
    assert nb > 0
    return nb * 8 * fps / n / 1000.0, psnr_planes(yuv, rec, w, h, n)


def reference(path, w, h, n, qp, fps, extra=()):
    with tempfile.TemporaryDirectory() as d:
        cmd = [REF, "-i", path, "-wdt", str(w), "-hgt", str(h), "-fr", str(fps), "-preset", "veryfast", "-rc", "0", "-qp", str(qp), "-iper", "128",
               "-frms", str(n), "-threads", "1", "-psnr", "1", "-b", os.path.join(d, "r.265"), *extra]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=3600).stdout
    m = re.search(r"bitrate, psnr:\s*([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", out)
    return float(m.group(1)), [float(m.group(i)) for i in (2, 3, 4)]


def main():
    path, w, h, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    fps = int(sys.argv[5]) if len(sys.argv) > 5 else 15
    qps = [int(q) for q in sys.argv[6:]] or [27, 32]
    yuv = np.fromfile(path, np.uint8)[:w * h * 3 // 2 * n]
    print("| | " + " | ".join("`-qp %d` kbps / PSNR-Y" % q for q in qps) + " |")
    print("|---|" + "---|" * len(qps))
    rows = [("reference, default GOP (hierarchical B)", lambda q: reference(path, w, h, n, q, fps)),
            ("reference `-bframes 0`", lambda q: reference(path, w, h, n, q, fps, ("-bframes", "0"))),
            ("ours", lambda q: ours(yuv, w, h, n, q, fps)),
            ("ours `-bframes 2`", lambda q: ours(yuv, w, h, n, q, fps, bframes=2))]
    for name, fn in rows:
        cells = []
        for q in qps:
            kbps, ps = fn(q)
            cells.append("%.0f / %.2f" % (kbps, ps[0]))
        print("| %s | %s |" % (name, " | ".join(cells)))


if __name__ == "__main__":
    main()
