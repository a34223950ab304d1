#!/usr/bin/env python3
"""SURVEY.md 8c tier P2, by hand: encode a clip with the REFERENCE encoder (oracle/_ref/appencoder, any options), then
  1. re-create its reconstruction from the parsed stream with the oracle's kernels (oracle/ora_replay.c) and compare every picture with what
     the reference DECODER writes, and
  2. hold the levels it coded against OUR forward transform + quantiser + sign-data hiding on the same residuals, and
  3. (-bframes 0 streams) hold its zero-block decisions against our RD zero-out on its own predictions, its vectors against our search on its
     own reference pictures, and its SAO parameters against our SAO decision on its own deblocked pictures.
usage: replay_check.py clip.yuv width height qp preset frames [extra appencoder options, e.g. -bframes 0 / -rc 3 -crf 26]"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import stream_stats  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    clip, w, h, qp, preset, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6])
    extra = sys.argv[7:]
    rc_args = [] if "-rc" in extra else ["-rc", "0", "-qp", str(qp)]
    with tempfile.TemporaryDirectory() as d:
        bsf, decf = os.path.join(d, "r.265"), os.path.join(d, "d.yuv")
        subprocess.run([os.path.join(REF, "appencoder"), "-i", clip, "-wdt", str(w), "-hgt", str(h), "-fr", "15", "-preset", preset, *rc_args, "-iper", "128",
                        "-frms", str(n), "-threads", "1", "-b", bsf, *extra], capture_output=True, check=True)
        subprocess.run([os.path.join(REF, "appdecoder"), "-b", bsf, "-o", decf, "-threads", "1"], capture_output=True, check=True)
        bs, dec = np.fromfile(bsf, np.uint8), np.fromfile(decf, np.uint8)
    O = C.CDLL(os.path.join(ROOT, "oracle", "libks_oracle.so"))
    O.ora_parse_stream.restype = C.c_void_p
    O.ora_parse_stream.argtypes = [C.c_void_p, C.c_size_t]
    O.ora_parse_num_pics.argtypes = [C.c_void_p]
    O.ora_parse_pic_stats.restype = C.POINTER(stream_stats.PicStats)
    O.ora_parse_pic_stats.argtypes = [C.c_void_p, C.c_int]
    O.ora_replay_pictures.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    O.ora_replay_compare_levels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    ps = O.ora_parse_stream(bs.ctypes.data, bs.size)
    npic, fs = O.ora_parse_num_pics(ps), w * h * 3 // 2
    out = np.zeros(fs * npic, np.uint8)
    rc = O.ora_replay_pictures(ps, 0, npic, out.ctypes.data)
    st = [O.ora_parse_pic_stats(ps, i).contents for i in range(npic)]
    order = sorted(range(npic), key=lambda i: st[i].poc)
    bad = sum(not np.array_equal(out[i * fs:(i + 1) * fs], dec[k * fs:(k + 1) * fs]) for k, i in enumerate(order))
    print("stream %d bytes, %d pictures (slice types in decoding order: %s); replay rc %d; pictures differing from the reference decoder: %d"
          % (bs.size, npic, "".join("BPI"[s.slice_type] for s in st), rc, bad))
    src = np.fromfile(clip, np.uint8, count=fs * n)
    cnt = (C.c_long * 20)()
    rc2 = O.ora_replay_compare_levels(ps, 0, npic, src.ctypes.data, cnt)
    print("levels the reference coded vs our transform-block coder on the same residuals (rc %d):" % rc2)
    for k, name in enumerate(("I-slice luma", "I-slice chroma", "intra blocks in P/B slices", "inter luma", "inter chroma")):
        a, b, c, e = cnt[4 * k:4 * k + 4]
        if a:
            print("  %-28s %7d blocks, %7d identical (%.1f %%); %8d coefficient positions, %7d differ (%.2f %%)" % (name, a, b, 100.0 * b / a, c, e, 100.0 * e / max(c, 1)))
    if "-bframes" in extra and extra[extra.index("-bframes") + 1] == "0":
        O.ora_replay_compare_zero_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        print("zero-block decisions on the reference's own predictions (luma blocks of its inter CUs, P pictures):")
        for delta in (0, 3):
            z = (C.c_long * 6)()
            O.ora_replay_compare_zero_blocks(ps, 0, npic, src.ctypes.data, delta, z)
            b, ref, plain, ours, both, neither = list(z)
            if b:
                print("  lambda of QP+%d on non-key pictures: %d blocks; reference codes %.1f %%, plain quantiser %.1f %%, ours %.1f %%; same decision on %.1f %%"
                      % (delta, b, 100.0 * ref / b, 100.0 * plain / b, 100.0 * ours / b, 100.0 * (both + neither) / b))
        O.ora_replay_compare_me.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        for method, name in ((0, "small diamond"), (1, "hexagon")):
            k = (C.c_long * 23)()
            O.ora_replay_compare_me(ps, 0, npic, src.ctypes.data, method, k)
            if k[0]:
                print("our search (%s) on the reference's own reference pictures: %d cells; same vector %.1f %%, within 1/4 sample %.1f %%, our SAD <= the SAD at its "
                      "vector %.1f %%; mean SAD ours %.1f, reference %.1f" % (name, k[0], 100.0 * k[1] / k[0], 100.0 * k[2] / k[0], 100.0 * k[3] / k[0], k[4] / k[0], k[5] / k[0]))
                for b, rng in enumerate(("< 2", "2-8", "8-16", "16-32", ">= 32")):
                    if k[8 + 3 * b]:
                        print("    reference vector %5s samples: %6d cells, mean SAD ours %.0f, reference %.0f" % (rng, k[8 + 3 * b], k[9 + 3 * b] / k[8 + 3 * b], k[10 + 3 * b] / k[8 + 3 * b]))
        O.ora_replay_compare_sao.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        z = (C.c_long * 16)()
        O.ora_replay_compare_sao(ps, 0, npic, src.ctypes.data, 3 if preset in ("ultrafast", "superfast", "veryfast", "fast") else 4, z)
        for g, name in ((0, "luma"), (1, "chroma")):
            k = z[8 * g:8 * g + 7]
            if k[0]:
                print("SAO %s: %d CTUs; reference on %.1f %%, ours on %.1f %%, same type %.1f %%; both on with the same type %d: same class/band %.1f %%, same offsets %.1f %%"
                      % (name, k[0], 100.0 * k[1] / k[0], 100.0 * k[2] / k[0], 100.0 * k[3] / k[0], k[4], 100.0 * k[5] / max(k[4], 1), 100.0 * k[6] / max(k[4], 1)))
    return 1 if rc or bad else 0


if __name__ == "__main__":
    sys.exit(main())
