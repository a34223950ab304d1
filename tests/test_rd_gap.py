"""CPU-only: rate / PSNR of this repo's encoder (the CPU model == the CUDA path bit for bit) against the REFERENCE encoder at matched picture
structure (`-bframes 0`) on the committed natural clip (tests/golden/nat_320x240_6f.yuv.gz, a crop of the reference repo's own 640x480 clip).

SURVEY.md 8c tier P3 suggests <= +10 % bits at >= -0.3 dB.  That is NOT met yet (DESIGN.md section 2 has the table for the full clips); this
test is the regression guard at the level actually reached, so that the gap can only shrink: round 1 stood at 2.1-2.8x the reference's bits on
natural content, the CU/merge decision + RD zero-out + raised lambda on non-key pictures + intra CUs in P pictures + 8x8 intra CUs brought
it to the bounds asserted here."""
import gzip
import os
import sys

import numpy as np
import pytest

from katlib import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import rd_compare  # noqa: E402


@pytest.mark.parametrize("qp,max_ratio,min_dpsnr", [(27, 1.45, -0.6), (32, 1.55, -0.55)])
def test_bits_and_psnr_against_reference_p_only(qp, max_ratio, min_dpsnr, tmp_path):
    if not os.path.exists(rd_compare.REF):
        pytest.skip("oracle/_ref/appencoder not staged (needs /root/reference once: make -C oracle)")
    yuv = np.frombuffer(gzip.open(os.path.join(GOLDEN, "nat_320x240_6f.yuv.gz"), "rb").read(), np.uint8)
    clip = tmp_path / "nat.yuv"
    clip.write_bytes(yuv.tobytes())
    kbps, psnr = rd_compare.ours(yuv, 320, 240, 6, qp, 15)
    rk, rp = rd_compare.reference(str(clip), 320, 240, 6, qp, 15, ("-bframes", "0"))
    assert kbps <= rk * max_ratio, "bitrate %.1f kbps vs reference %.1f (x%.2f, bound x%.2f)" % (kbps, rk, kbps / rk, max_ratio)
    assert psnr[0] - rp[0] >= min_dpsnr, "PSNR-Y %.2f vs reference %.2f" % (psnr[0], rp[0])
