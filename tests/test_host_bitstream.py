"""CPU-only: the host bitstream writer (product code) driven by the CPU picture model, arbitrated by the REFERENCE
DECODER (oracle/_ref/appdecoder; SURVEY.md 8c tier P1) and frozen by committed stream/recon digests."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from katlib import GOLDEN, ROOT, SeqCfg, oracle, ptr

sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_yuv  # noqa: E402

DEC = os.path.join(ROOT, "oracle", "_ref", "appdecoder")


def model_encode(yuv, w, h, n, qp, iper, sbh=1, sao=1, subpel=2, bframes=0, me=0, rc=0, crf=24.0):
    O = oracle()
    O.ora_encode_sequence.restype = C.c_long
    cfg = SeqCfg(w, h, n, qp, iper, 0, 64, 16, subpel, sbh, sao, 3, 0, bframes, me, rc, int(round(crf * 100)))
    bs = np.zeros(w * h * 3 * n + 100000, np.uint8); rec = np.zeros(w * h * 3 // 2 * n, np.uint8)
    nb = O.ora_encode_sequence(C.byref(cfg), ptr(yuv), ptr(bs), C.c_size_t(bs.size), ptr(rec))
    assert nb > 0
    return bs[:nb], rec


CASES = [("syn_192x112", 192, 112, 5, 32, 16, 1, 3, 2), ("syn_256x144_sao4", 256, 144, 4, 30, 16, 1, 4, 2), ("syn_200x120_pad", 200, 120, 4, 27, 2, 1, 1, 2),
         ("syn_320x240_nosbh", 320, 240, 4, 24, 8, 0, 0, 1), ("syn_416x240_intra", 416, 240, 2, 35, 1, 1, 1, 2),
         ("syn_192x112_b1", 192, 112, 6, 32, 16, 1, 3, 2, 1), ("syn_320x176_b3", 320, 176, 9, 28, 16, 1, 1, 2, 3), ("syn_200x120_b2_gops", 200, 120, 9, 30, 4, 1, 4, 2, 2),
         ("syn_256x144_hex", 256, 144, 5, 30, 16, 1, 3, 2, 0, 1), ("syn_320x176_crf26", 320, 176, 8, 0, 4, 1, 3, 2, 0, 0, 3, 26.0), ("syn_192x112_crf22_b2_hex", 192, 112, 7, 0, 16, 1, 1, 2, 2, 1, 3, 22.0)]


def _yuv(name, w, h, n):
    return np.frombuffer(gen_yuv.make(w, h, n, seed=21), np.uint8)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_closed_loop_with_reference_decoder(case):
    name, w, h, n, qp, iper, sbh, sao, subpel = case[:9]
    extra = case[9:]
    if not os.path.exists(DEC):
        pytest.skip("oracle/_ref/appdecoder not staged (needs /root/reference once: make -C oracle)")
    bs, rec = model_encode(_yuv(name, w, h, n), w, h, n, qp, iper, sbh, sao, subpel, *extra)
    with tempfile.TemporaryDirectory() as d:
        p, o = os.path.join(d, "t.265"), os.path.join(d, "t.yuv")
        open(p, "wb").write(bs.tobytes())
        r = subprocess.run([DEC, "-b", p, "-o", o, "-threads", "1"], capture_output=True, text=True, timeout=300)
        assert os.path.exists(o), r.stdout[-300:] + r.stderr[-300:]
        dec = np.fromfile(o, np.uint8)
    assert dec.size == rec.size and np.array_equal(dec, rec), "reference decoder output != model reconstruction"


def test_natural_clip_closed_loop():
    import gzip
    if not os.path.exists(DEC):
        pytest.skip("oracle/_ref/appdecoder not staged")
    yuv = np.frombuffer(gzip.open(os.path.join(GOLDEN, "nat_320x240_6f.yuv.gz"), "rb").read(), np.uint8)
    bs, rec = model_encode(yuv, 320, 240, 6, 30, 6)
    with tempfile.TemporaryDirectory() as d:
        p, o = os.path.join(d, "t.265"), os.path.join(d, "t.yuv")
        open(p, "wb").write(bs.tobytes())
        subprocess.run([DEC, "-b", p, "-o", o, "-threads", "1"], capture_output=True, text=True, timeout=300)
        dec = np.fromfile(o, np.uint8)
    assert np.array_equal(dec, rec)


def test_stream_digests_are_frozen():
    """golden digests (tests/golden/model_digests.json, generated while the decoder check above was green) keep the
    model + bitstream writer from drifting on boxes where the reference decoder is not available"""
    path = os.path.join(GOLDEN, "model_digests.json")
    got = {}
    for case in CASES:
        name, w, h, n, qp, iper, sbh, sao, subpel = case[:9]
        bs, rec = model_encode(_yuv(name, w, h, n), w, h, n, qp, iper, sbh, sao, subpel, *case[9:])
        got[name] = {"bs_md5": hashlib.md5(bs.tobytes()).hexdigest(), "rec_md5": hashlib.md5(rec.tobytes()).hexdigest(), "bytes": int(bs.size)}
    if os.environ.get("KS_WRITE_GOLDEN") == "1":
        json.dump(got, open(path, "w"), indent=1, sort_keys=True)
    want = json.load(open(path))
    assert got == want


def test_md5_matches_hashlib():
    O = oracle()
    rng = np.random.default_rng(0)
    for n in (0, 1, 55, 56, 63, 64, 65, 1000, 4096 + 7):
        b = rng.integers(0, 256, n, dtype=np.uint8)
        d = np.zeros(16, np.uint8)
        O.ks_md5(ptr(b) if n else None, C.c_size_t(n), ptr(d))
        assert d.tobytes() == hashlib.md5(b.tobytes()).digest()
