"""Host-side logic that needs no GPU: the per-picture QP decision (ks_ratecontrol.c) and the coding schedule of a GOP shard."""
import ctypes as C

import pytest

import ks265codec_b200 as ks
from katlib import oracle

KS_SLICE_B, KS_SLICE_P, KS_SLICE_I = 0, 1, 2


class KsRc(C.Structure):
    """csrc/host/ks_ratecontrol.h: ks_rc"""
    _fields_ = [("mode", C.c_int), ("qp", C.c_int), ("fixqp", C.c_int), ("bframes", C.c_int), ("crf", C.c_double), ("cplx_sum", C.c_double),
                ("cplx_cnt", C.c_double), ("base_cplx", C.c_double), ("qp_min", C.c_int), ("qp_max", C.c_int)]


def _rc(lib, mode, qp=27, fixqp=0, crf=24.0, cells=32400, bframes=0):
    lib.ks_rc_init.argtypes = [C.POINTER(KsRc), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
    lib.ks_rc_picture_qp.argtypes = [C.POINTER(KsRc), C.c_int, C.c_int]
    lib.ks_rc_update.argtypes = [C.POINTER(KsRc), C.c_int, C.c_uint64]
    rc = KsRc()
    assert lib.ks_rc_init(C.byref(rc), mode, qp, fixqp, crf, cells, bframes) == (0 if mode in (0, 3) else -1)
    return rc


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_fixed_qp_follows_the_reference_cascade(which):
    L = ks.lib() if which == "product" else oracle()
    rc = _rc(L, 0, qp=27)
    assert L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_I, 0) == 27
    assert [L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_P, p) for p in range(1, 9)] == [30, 29, 30, 28, 30, 29, 30, 28]     # reference -bframes 0 -qp 27 [probe]
    rc = _rc(L, 0, qp=27, fixqp=1)
    assert {L.ks_rc_picture_qp(C.byref(rc), t, p) for t in (KS_SLICE_I, KS_SLICE_P, KS_SLICE_B) for p in range(5)} == {27}
    rc = _rc(L, 0, qp=30, bframes=2)
    assert (L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_P, 3), L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_B, 1)) == (31, 33)
    rc = _rc(L, 0, qp=50)
    assert L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_P, 1) == 51                                                          # clipped


def test_crf_tracks_the_search_cost():
    L = ks.lib()
    cells = 32400
    rc = _rc(L, 3, crf=24.0, cells=cells)
    assert L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_I, 0) == 21                      # no history yet: crf - 3
    assert L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_P, 4) == 24                      # poc 4 is the cascade's anchor (+0)
    L.ks_rc_update(C.byref(rc), KS_SLICE_P, 512 * cells)                             # exactly the calibration complexity
    assert L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_P, 8) == 24
    for _ in range(6):
        L.ks_rc_update(C.byref(rc), KS_SLICE_P, 4 * 512 * cells)                     # 4x the complexity: +2.4 * log2(4) = +4.8
    assert L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_P, 8) == 29
    assert L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_P, 9) == 31                      # cascade +2 on top
    L.ks_rc_update(C.byref(rc), KS_SLICE_I, 10 ** 12)                                # I and B pictures do not move the model
    assert L.ks_rc_picture_qp(C.byref(rc), KS_SLICE_P, 8) == 29
    assert _rc(L, 1) is not None                                                      # ABR is refused by ks_rc_init (checked inside _rc)


def test_gop_schedule_codes_anchors_before_their_b_pictures():
    O = oracle()
    n, bf = 10, 3
    arr = (C.c_int * (n + 1))
    order, typ, l0, l1 = arr(), arr(), arr(), arr()
    cnt = O.ora_gop_schedule(n, bf, order, typ, l0, l1)
    assert cnt == n and sorted(order[:cnt]) == list(range(n))
    assert list(order[:cnt]) == [0, 4, 1, 2, 3, 8, 5, 6, 7, 9]
    assert [typ[i] for i in range(cnt)] == [KS_SLICE_I, KS_SLICE_P, KS_SLICE_B, KS_SLICE_B, KS_SLICE_B, KS_SLICE_P, KS_SLICE_B, KS_SLICE_B, KS_SLICE_B, KS_SLICE_P]
    seen = set()
    for i in range(cnt):
        for ref in (l0[i], l1[i]):
            assert ref < 0 or ref in seen, "picture %d references %d before it is coded" % (order[i], ref)
        seen.add(order[i])
    assert O.ora_gop_schedule(5, 0, order, typ, l0, l1) == 5 and list(order[:5]) == [0, 1, 2, 3, 4]
