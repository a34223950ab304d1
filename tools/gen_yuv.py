#!/usr/bin/env python3
"""Deterministic synthetic I420 generator (SURVEY.md 8d / Appendix A.4 recipe, vectorised).

multi-octave smoothed texture (cells 64/16/4/2 px, amplitudes 50/30/18/10, mean 128) on a (W+256)x(H+128) canvas;
frame t = bilinear sample at offset (2.5t, 1.0t) px; one W/8 x H/8 inverted box moving (+6,+3) px/frame;
+ N(0, 2^2) luma noise per frame; chroma same recipe x0.4.  usage: gen_yuv.py W H NFRAMES SEED OUT.yuv
"""
import sys
import numpy as np


def _box(a, o, axis):
    c = np.cumsum(np.insert(a, 0, 0, axis=axis), axis=axis, dtype=np.float64)
    n = a.shape[axis]
    lo = np.clip(np.arange(n) - o // 2, 0, n); hi = np.clip(np.arange(n) - o // 2 + o, 0, n)
    return ((np.take(c, hi, axis=axis) - np.take(c, lo, axis=axis)) / o).astype(np.float32)


def tex(rng, W, H):
    acc = np.zeros((H, W), np.float32)
    for o, amp in ((64, 50), (16, 30), (4, 18), (2, 10)):
        g = rng.standard_normal((H // o + 3, W // o + 3)).astype(np.float32)
        up = np.repeat(np.repeat(g, o, axis=0), o, axis=1)
        up = _box(_box(up, o, 1), o, 0)
        acc += amp * up[:H, :W] / max(float(up.std()), 1e-6) * 0.5
    return acc


def samp(P, x, y, W, H):
    x0, y0 = int(x), int(y); ax, ay = np.float32(x - x0), np.float32(y - y0)
    a = P[y0:y0 + H, x0:x0 + W]; b = P[y0:y0 + H, x0 + 1:x0 + 1 + W]
    c = P[y0 + 1:y0 + 1 + H, x0:x0 + W]; d = P[y0 + 1:y0 + 1 + H, x0 + 1:x0 + 1 + W]
    return (1 - ay) * ((1 - ax) * a + ax * b) + ay * ((1 - ax) * c + ax * d)


def frames(w, h, n, seed=1234):
    rng = np.random.default_rng(seed)
    PW, PH = w + 256, h + 128
    Y = tex(rng, PW, PH) + 128; U = tex(rng, PW // 2, PH // 2) * 0.4 + 128; V = tex(rng, PW // 2, PH // 2) * 0.4 + 128
    for t in range(n):
        fx, fy = (2.5 * t) % 200, (1.0 * t) % 100
        y = samp(Y, fx, fy, w, h).copy(); u = samp(U, fx / 2, fy / 2, w // 2, h // 2); v = samp(V, fx / 2, fy / 2, w // 2, h // 2)
        bx, by = (w // 4 + 6 * t) % (w - w // 8), (h // 3 + 3 * t) % (h - h // 8)
        y[by:by + h // 8, bx:bx + w // 8] = 200 - y[by:by + h // 8, bx:bx + w // 8] * 0.3
        y += rng.standard_normal(y.shape).astype(np.float32) * 2.0
        yield b"".join(np.clip(p + 0.5, 0, 255).astype(np.uint8).tobytes() for p in (y, u, v))


def make(w, h, n, seed=1234):
    return b"".join(frames(w, h, n, seed))


if __name__ == "__main__":
    w, h, n, seed, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    with open(out, "wb") as f:
        for fr in frames(w, h, n, seed):
            f.write(fr)
