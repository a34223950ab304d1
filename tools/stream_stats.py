#!/usr/bin/env python3
"""Per-picture statistics of an HEVC stream (the reference encoder's or this repo's) from the oracle's slice-data parser (oracle/ora_parse.c):
where the bits go (SAO / split flags / CU headers / mvd / luma / chroma), CU sizes and kinds, non-zero levels.  CPU only; test infrastructure.
usage: stream_stats.py stream.265 [more.265 ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


class PicStats(C.Structure):
    """oracle/ora_parse.c: ora_pic_stats"""
    _fields_ = [("poc", C.c_int), ("slice_type", C.c_int), ("qp", C.c_int), ("nal_type", C.c_int), ("num_ref", C.c_int * 2)] + \
               [(n, C.c_long) for n in "bits_total bits_sao bits_split bits_cu_hdr bits_mvd bits_luma bits_chroma bits_intra_mode".split()] + \
               [("n_cu", C.c_long * 4), ("n_skip", C.c_long * 4), ("n_merge", C.c_long * 4), ("n_amvp", C.c_long * 4), ("n_intra", C.c_long * 4),
                ("n_intra_nxn", C.c_long), ("n_tu", C.c_long * 4), ("n_cbf_luma", C.c_long), ("n_cbf_chroma", C.c_long)] + \
               [(n, C.c_long) for n in "nz_luma nz_chroma sum_abs_luma sum_abs_chroma n_mvd_nonzero sao_on_luma sao_on_chroma sao_merge".split()]


def parse(path_or_bytes):
    """returns (error, [PicStats copies], [ok flags])"""
    from katlib import oracle
    O = oracle()
    O.ora_parse_stream.restype = C.c_void_p
    O.ora_parse_stream.argtypes = [C.c_char_p, C.c_size_t]
    O.ora_parse_pic_stats.restype = C.POINTER(PicStats)
    O.ora_parse_pic_stats.argtypes = [C.c_void_p, C.c_int]
    O.ora_parse_sizeof_stats.restype = C.c_size_t
    for f in ("ora_parse_error", "ora_parse_num_pics"):
        getattr(O, f).argtypes = [C.c_void_p]
    O.ora_parse_pic_ok.argtypes = [C.c_void_p, C.c_int]
    O.ora_parse_free.argtypes = [C.c_void_p]
    assert O.ora_parse_sizeof_stats() == C.sizeof(PicStats), "PicStats mirror is stale"
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    h = O.ora_parse_stream(bytes(data), len(data))
    n = O.ora_parse_num_pics(h)
    pics, oks = [], []
    for i in range(n):
        st = PicStats()
        C.memmove(C.byref(st), O.ora_parse_pic_stats(h, i), C.sizeof(PicStats))
        pics.append(st); oks.append(bool(O.ora_parse_pic_ok(h, i)))
    err = O.ora_parse_error(h)
    O.ora_parse_free(h)
    return err, pics, oks


def main():
    for path in sys.argv[1:]:
        err, pics, oks = parse(path)
        print("%s: %d pictures, parse error %d" % (path, len(pics), err))
        print("poc t qp   total |   sao split cuhdr   mvd  luma chroma | cu8/16/32/64 | skip merge amvp intra(nxn) | nzY nzC | sao on Y/C merge")
        for st, ok in zip(pics, oks):
            print("%3d %s %2d %7d | %5d %5d %5d %5d %6d %5d | %s | %4d %4d %4d %4d(%d) | %6d %5d | %d/%d %d%s" % (
                st.poc, "BPI"[st.slice_type], st.qp, st.bits_total, st.bits_sao, st.bits_split, st.bits_cu_hdr - st.bits_mvd, st.bits_mvd, st.bits_luma, st.bits_chroma,
                "/".join(str(v) for v in st.n_cu), sum(st.n_skip), sum(st.n_merge), sum(st.n_amvp), sum(st.n_intra), st.n_intra_nxn,
                st.nz_luma, st.nz_chroma, st.sao_on_luma, st.sao_on_chroma, st.sao_merge, "" if ok else "  <-- PARSE FAILED"))


if __name__ == "__main__":
    main()
