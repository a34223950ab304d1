"""SURVEY.md 8c tier P2 on the CPU: I pictures and P-only sequences coded by the REFERENCE encoder are re-created from their own parsed decisions (oracle/ora_parse.c)
with the oracle's leaf kernels (oracle/ora_replay.c) and must equal what the reference DECODER makes of them, byte for byte.  The streams and
the decoder's MD5 are committed (tests/golden/ref_intra_streams.json, made by tests/golden/make_replay_golden.py); where the reference decoder
is staged (oracle/_ref) the comparison is also made against a live decode."""
import base64
import ctypes as C
import hashlib
import json
import os
import subprocess

import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_intra_streams.json")))
SEQS = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_p_streams.json")))
DEC = os.path.join(ROOT, "oracle", "_ref", "appdecoder")
if os.path.join(ROOT, "tools") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tools"))
import stream_stats  # noqa: E402  (the ctypes mirror of the parser's per-picture record)


def _oracle():
    O = C.CDLL(os.path.join(ROOT, "oracle", "libks_oracle.so"))
    O.ora_parse_stream.restype = C.c_void_p
    O.ora_parse_stream.argtypes = [C.c_void_p, C.c_size_t]
    O.ora_parse_free.argtypes = [C.c_void_p]
    for f in (O.ora_parse_error, O.ora_parse_num_pics, O.ora_parse_width, O.ora_parse_height):
        f.argtypes = [C.c_void_p]
    O.ora_replay_intra_picture.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    O.ora_replay_pictures.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    O.ora_parse_pic_stats.restype = C.POINTER(stream_stats.PicStats)
    O.ora_parse_pic_stats.argtypes = [C.c_void_p, C.c_int]
    return O


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_intra_picture_is_recreated_from_its_parsed_decisions(case, tmp_path):
    O = _oracle()
    bs = np.frombuffer(base64.b64decode(case["stream_b64"]), np.uint8).copy()
    ps = O.ora_parse_stream(bs.ctypes.data, bs.size)
    try:
        assert O.ora_parse_error(ps) == 0 and O.ora_parse_num_pics(ps) == 1
        w, h = O.ora_parse_width(ps), O.ora_parse_height(ps)
        assert (w, h) == (case["width"], case["height"])            # sizes without a conformance window: coded == display
        out = np.zeros(w * h * 3 // 2, np.uint8)
        assert O.ora_replay_intra_picture(ps, 0, out.ctypes.data) == 0
    finally:
        O.ora_parse_free(ps)
    assert hashlib.md5(out.tobytes()).hexdigest() == case["decoded_md5"]
    if os.path.exists(DEC):
        p, o = tmp_path / "s.265", tmp_path / "d.yuv"
        p.write_bytes(bs.tobytes())
        subprocess.run([DEC, "-b", str(p), "-o", str(o), "-threads", "1"], capture_output=True, timeout=120)
        assert np.array_equal(np.fromfile(o, np.uint8), out)


@pytest.mark.parametrize("seq", SEQS, ids=[c["name"] for c in SEQS])
def test_reference_p_sequence_is_recreated_from_its_parsed_decisions(seq, tmp_path):
    """-bframes 0 and default-GOP (hierarchical B) streams of the reference encoder: skip / merge (spatial, temporal, combined bi-predictive
    candidates) / AMVP over several reference pictures in both lists, 8x8..64x64 inter CUs incl. 2NxN / Nx2N partitions, bi-prediction, intra
    CUs inside inter pictures, residuals, deblocking strengths from vectors and coefficients, SAO.  The replay runs in decoding order; the
    decoder writes display order."""
    O = _oracle()
    bs = np.frombuffer(base64.b64decode(seq["stream_b64"]), np.uint8).copy()
    ps = O.ora_parse_stream(bs.ctypes.data, bs.size)
    try:
        n = O.ora_parse_num_pics(ps)
        assert O.ora_parse_error(ps) == 0 and n == seq["frames"]
        w, h = O.ora_parse_width(ps), O.ora_parse_height(ps)
        fs = w * h * 3 // 2
        out = np.zeros(fs * n, np.uint8)
        assert O.ora_replay_pictures(ps, 0, n, out.ctypes.data) == 0
        # a later picture on its own: the pictures it references are replayed behind the scenes
        one = np.zeros(fs, np.uint8)
        assert O.ora_replay_pictures(ps, n - 1, 1, one.ctypes.data) == 0 and np.array_equal(one, out[(n - 1) * fs:])
        pocs = [O.ora_parse_pic_stats(ps, i).contents.poc for i in range(n)]
    finally:
        O.ora_parse_free(ps)
    out = np.concatenate([out[i * fs:(i + 1) * fs] for i in sorted(range(n), key=lambda i: pocs[i])])          # display order
    assert [hashlib.md5(out[f * fs:(f + 1) * fs].tobytes()).hexdigest() for f in range(n)] == seq["decoded_md5"]
    if os.path.exists(DEC):
        p, o = tmp_path / "s.265", tmp_path / "d.yuv"
        p.write_bytes(bs.tobytes())
        subprocess.run([DEC, "-b", str(p), "-o", str(o), "-threads", "1"], capture_output=True, timeout=120)
        assert np.array_equal(np.fromfile(o, np.uint8), out)


@pytest.mark.parametrize("seq", [c for c in SEQS if c["name"].startswith("nat320_")], ids=[c["name"] for c in SEQS if c["name"].startswith("nat320_")])
def test_reference_levels_against_our_transform_block_coder(seq):
    """Every transform block with coefficients in the reference's stream: source minus the reference's own prediction (re-created by the replay),
    through OUR forward transform + quantiser + sign-data hiding, against the levels the reference coded.  At the presets without RD
    quantisation (ultrafast .. veryfast, the north star's) every level must be identical; from `fast` up the reference switches its RDOQ on
    (SURVEY: rdoQuant E@0x4a3040), which this repo does not have, and a large share differs -- asserted too, so the check cannot pass vacuously."""
    import gzip
    O = _oracle()
    O.ora_replay_compare_levels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    src = np.frombuffer(gzip.open(os.path.join(ROOT, "tests", "golden", "nat_320x240_6f.yuv.gz"), "rb").read(), np.uint8).copy()
    bs = np.frombuffer(base64.b64decode(seq["stream_b64"]), np.uint8).copy()
    ps = O.ora_parse_stream(bs.ctypes.data, bs.size)
    cnt = (C.c_long * 20)()
    try:
        assert O.ora_replay_compare_levels(ps, 0, O.ora_parse_num_pics(ps), src.ctypes.data, cnt) == 0
    finally:
        O.ora_parse_free(ps)
    blocks, same = sum(cnt[4 * k] for k in range(5)), sum(cnt[4 * k + 1] for k in range(5))
    assert blocks > 300
    if seq["preset"] in ("ultrafast", "superfast", "veryfast"):
        assert same == blocks and sum(cnt[4 * k + 3] for k in range(5)) == 0
    else:
        assert same < 0.9 * blocks


def test_reference_zero_block_decisions_against_our_rd_zero_out():
    """Row a14 (the per-TU driver's zero-block decisions, closed in the reference): on the reference's OWN predictions, which luma transform
    blocks of its inter CUs carry levels, and which would ours keep?  The plain quantiser would code 75 % of them, the reference codes 38 %; our
    RD zero-out with the lambda schedule of ks_rc_lambda_qp (QP+3 on the non-key pictures of the P cascade) codes 34 % and agrees block by block
    on 88 % (without the raised lambda: 60 % coded, 77 % agreement) -- on 720p x 16 pictures the same comparison gives 16.2 % vs 16.1 % coded
    and 95.6 % agreement (tools/replay_check.py)."""
    import gzip
    O = _oracle()
    O.ora_replay_compare_zero_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    src = np.frombuffer(gzip.open(os.path.join(ROOT, "tests", "golden", "nat_320x240_6f.yuv.gz"), "rb").read(), np.uint8).copy()
    seq = [c for c in SEQS if c["name"] == "nat320_veryfast_qp27_6f"][0]
    bs = np.frombuffer(base64.b64decode(seq["stream_b64"]), np.uint8).copy()
    ps = O.ora_parse_stream(bs.ctypes.data, bs.size)
    res = {}
    try:
        for delta in (0, 3):
            cnt = (C.c_long * 6)()
            assert O.ora_replay_compare_zero_blocks(ps, 0, O.ora_parse_num_pics(ps), src.ctypes.data, delta, cnt) == 0
            res[delta] = list(cnt)
    finally:
        O.ora_parse_free(ps)
    blocks, ref, plain, ours, both, neither = res[3]
    assert blocks > 500 and plain > 1.5 * ref                                   # the reference drops about half of what a plain quantiser would code
    assert abs(ours - ref) < 0.2 * ref and both + neither > 0.85 * blocks        # ours: the same share, 85+ % the same blocks
    assert res[0][3] > 1.3 * ref and res[0][4] + res[0][5] < both + neither      # without the raised lambda: far more coded blocks, less agreement


def test_reference_sao_parameters_against_our_decision():
    """Row a18 (the SAO offset decision, closed in the reference; ours is our own): run OUR decision on the reference's deblocked pictures and
    compare with the parameters it coded.  The statistics and the apply stage are pinned elsewhere (KAT, replay); this only MEASURES how far the
    decision is from the reference's -- it switches SAO on for about twice as many CTUs as ours (720p x 16 pictures: 48 % vs 20 % of the luma
    CTUs) -- and keeps the comparison from rotting."""
    import gzip
    O = _oracle()
    O.ora_replay_compare_sao.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    src = np.frombuffer(gzip.open(os.path.join(ROOT, "tests", "golden", "nat_320x240_6f.yuv.gz"), "rb").read(), np.uint8).copy()
    seq = [c for c in SEQS if c["name"] == "nat320_veryfast_qp27_6f"][0]
    bs = np.frombuffer(base64.b64decode(seq["stream_b64"]), np.uint8).copy()
    ps = O.ora_parse_stream(bs.ctypes.data, bs.size)
    cnt = (C.c_long * 16)()
    try:
        assert O.ora_replay_compare_sao(ps, 0, O.ora_parse_num_pics(ps), src.ctypes.data, 3, cnt) == 0
    finally:
        O.ora_parse_free(ps)
    ctus, ref_on, ours_on, same_type = cnt[0], cnt[1], cnt[2], cnt[3]
    assert ctus == 6 * 5 * 4 and ref_on > 0 and 0 < ours_on <= ref_on and same_type >= 0.25 * ctus


def test_reference_vectors_against_our_search():
    """Rows a3-a6 (search + refinement; the reference's start-point logic is closed): OUR search runs, cell by cell, on the reference encoder's own
    reference pictures and is compared with the vectors it chose.  A measurement, kept alive by this test: on content with little motion ours finds
    an equal or lower SAD almost everywhere (720p natural: 96.7 % of 50 523 cells, same vector in 82 %; synthetic: 99.2 %), on content with real
    motion it loses (480p natural: mean SAD 752 vs 550, 35 % of the cells worse, growing with the vector length) -- DESIGN 9 ranks the fix."""
    import gzip
    O = _oracle()
    O.ora_replay_compare_me.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    src = np.frombuffer(gzip.open(os.path.join(ROOT, "tests", "golden", "nat_320x240_6f.yuv.gz"), "rb").read(), np.uint8).copy()
    seq = [c for c in SEQS if c["name"] == "nat320_veryfast_qp27_6f"][0]
    bs = np.frombuffer(base64.b64decode(seq["stream_b64"]), np.uint8).copy()
    ps = O.ora_parse_stream(bs.ctypes.data, bs.size)
    try:
        res = []
        for method in (0, 1):
            cnt = (C.c_long * 23)()
            assert O.ora_replay_compare_me(ps, 0, O.ora_parse_num_pics(ps), src.ctypes.data, method, cnt) == 0
            res.append(list(cnt))
    finally:
        O.ora_parse_free(ps)
    for k in res:
        assert k[0] > 300 and k[1] <= k[2] <= k[0] and k[6] + k[7] <= k[0] and k[4] > 0 and k[5] > 0
        assert sum(k[8 + 3 * b] for b in range(5)) == k[0]
        assert k[3] > 0.25 * k[0]                                               # (this crop of the 480p clip is the hard case: 39 %)


def test_replay_rejects_what_it_does_not_cover():
    O = _oracle()
    bs = np.frombuffer(base64.b64decode(CASES[0]["stream_b64"]), np.uint8).copy()
    ps = O.ora_parse_stream(bs.ctypes.data, bs.size)
    out = np.zeros(16, np.uint8)
    assert O.ora_replay_intra_picture(ps, 1, out.ctypes.data) == -1          # no such picture
    O.ora_parse_free(ps)
