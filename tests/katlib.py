"""KAT record reader + oracle (ctypes) bindings shared by the tests.

Record format is written by oracle/kat/harvest.c (vectors harvested from the reference binary's own
scalar kernels, SURVEY.md 8c tier P0).  TEST INFRASTRUCTURE: imports oracle/libks_oracle.so.
"""
import ctypes as C
import gzip
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")


def read_kat(path):
    """yield (name, params:list[int], inputs:list[bytes], outputs:list[bytes])"""
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        assert data[off:off + 4] == b"KAT1", "bad KAT magic"
        name = data[off + 4:off + 36].split(b"\0")[0].decode()
        off += 36
        (npar,) = struct.unpack_from("<I", data, off); off += 4
        params = list(struct.unpack_from("<%di" % npar, data, off)); off += 4 * npar
        (nb,) = struct.unpack_from("<I", data, off); off += 4
        ins, outs = [], []
        for _ in range(nb):
            is_out, nbytes = struct.unpack_from("<II", data, off); off += 8
            (outs if is_out else ins).append(data[off:off + nbytes]); off += nbytes
        yield name, params, ins, outs


_oracle = None


def oracle():
    """build (if needed) and load oracle/libks_oracle.so"""
    global _oracle
    if _oracle is not None:
        return _oracle
    so = os.path.join(ORACLE_DIR, "libks_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.startswith("ora_") or f == "ks_oracle.h"]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, os.path.join(ORACLE_DIR, "libks_oracle.so")])
    lib = C.CDLL(so)
    lib.ora_sad.restype = C.c_uint32
    lib.ora_sad.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_long]
    lib.ora_satd.restype = C.c_uint32
    lib.ora_satd.argtypes = lib.ora_sad.argtypes
    lib.ora_sad4.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_long]
    lib.ora_sad3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_long]
    lib.ora_sse.restype = C.c_uint32
    lib.ora_sse.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    _oracle = lib
    return lib


def ptr(a, byte_off=0):
    return C.c_void_p(a.ctypes.data + byte_off)


# ---- ctypes mirrors of the CPU model's configuration structs (ONE definition: tests, smoke() and tools import these) ----
class OraCfg(C.Structure):
    """oracle/ora_frame.h: ora_cfg"""
    _fields_ = [(n, C.c_int) for n in ("width", "height", "me_range", "me_iters", "subpel", "sign_hiding", "sao", "strong_intra", "satd", "me_method")]


class SeqCfg(C.Structure):
    """oracle/ora_encoder.c: ora_seq_cfg"""
    _fields_ = [(n, C.c_int) for n in "width height nframes qp iper fixqp me_range me_iters subpel sign_hiding sao max_merge_cand satd bframes me_method rc crf_x100".split()]
