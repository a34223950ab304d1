"""The reference's public encoder ABI (Android_demo/prebuilt/include/qy265enc.h:196-233) served by libks265qy.so (include/ks265_qyabi.h).
CPU: exports, struct layout against the reference's own header (where it exists), configuration calls, loud failure without CUDA.
GPU: the shim's access units, concatenated, are byte-for-byte the GOP shards of the encoder API; a caller compiled against the
reference's header (tests/qy/qy_caller.c) produces the same file and the reference decoder accepts it."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
LIB = os.path.join(ROOT, "ks265codec_b200", "libks265qy.so")
REF_INC = "/root/reference/Android_demo/prebuilt/include"
CALLER = os.path.join(ROOT, "ks265codec_b200", "bin", "qy_caller")
SYMBOLS = ["QY265EncoderOpen", "QY265EncoderClose", "QY265EncoderReconfig", "QY265EncoderEncodeHeaders", "QY265EncoderEncodeFrame",
           "QY265EncoderKeyFrameRequest", "QY265EncoderDelayedFrames", "QY265ConfigDefault", "QY265ConfigDefaultPreset", "QY265ConfigParse",
           "QY265SetLogPrintf", "QY265SetAuthWarning", "strLibQy265Version"]
OK, FAIL, NOTSUPPORTED = 0, C.c_int(0x80000001).value, C.c_int(0x80000004).value


class Vui(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("signal_type_present", "video_format", "full_range", "colour_desc_present", "primaries", "transfer", "matrix")]


class Config(C.Structure):                      # ksqy_config
    _fields_ = ([("auth", C.c_void_p)] + [(n, C.c_int) for n in ("tune", "preset", "latency", "profile_id", "headers_before_keyframe", "width", "height")]
                + [("fps", C.c_double), ("bframes", C.c_int), ("temporal_layer", C.c_int)]
                + [(n, C.c_int) for n in ("vpp_denoise", "vpp_edge", "vpp_color", "vpp_hdr")] + [("vpp_hdr_strength", C.c_double), ("vpp_hdr_iter", C.c_int)]
                + [(n, C.c_double) for n in ("vpp_hdr_sigma_s", "vpp_hdr_sigma_r", "vpp_recur_filter")]
                + [(n, C.c_int) for n in ("rc", "bitrate_kbps", "vbv_buffer_size", "vbv_max_rate", "vbv_min_rate", "qp", "crf", "visual_quality", "intra_period",
                                          "qp_min", "qp_max", "frame_skip", "wavefront", "frame_parallel", "threads", "vui_present")]
                + [("vui", Vui)] + [(n, C.c_int) for n in ("log_level", "lookahead", "calc_psnr", "calc_ssim", "short_loading", "pass_")]
                + [("stat_file", C.c_char * 256), ("rate_tolerance", C.c_double)]
                + [(n, C.c_int) for n in ("rdoq", "me", "part", "do64", "tu_inter", "tu_intra", "smooth", "transskip", "subme", "satd_inter", "satd_intra",
                                          "search_range", "ref_num", "ref0", "sao", "long_term_ref", "aq_mode")]
                + [("aq_strength", C.c_double), ("rasl", C.c_int)])


class Yuv(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("plane", C.c_void_p * 3), ("stride", C.c_int * 3)]


class Picture(C.Structure):
    _fields_ = [("slice_type", C.c_int), ("poc", C.c_int), ("pts", C.c_longlong), ("dts", C.c_longlong), ("yuv", C.POINTER(Yuv))]


class Nal(C.Structure):
    _fields_ = [("nal_type", C.c_int), ("tid", C.c_int), ("size", C.c_int), ("pts", C.c_longlong), ("payload", C.POINTER(C.c_ubyte))]


def shim():
    L = C.CDLL(LIB)
    L.QY265EncoderOpen.restype = C.c_void_p
    L.QY265EncoderOpen.argtypes = [C.POINTER(Config), C.POINTER(C.c_int)]
    L.QY265EncoderClose.argtypes = [C.c_void_p]
    L.QY265EncoderEncodeHeaders.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Nal)), C.POINTER(C.c_int)]
    L.QY265EncoderEncodeFrame.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Nal)), C.POINTER(C.c_int), C.POINTER(Picture), C.POINTER(Picture), C.c_int]
    L.QY265EncoderKeyFrameRequest.argtypes = [C.c_void_p]
    L.QY265EncoderReconfig.argtypes = [C.c_void_p, C.POINTER(Config)]
    L.QY265EncoderDelayedFrames.argtypes = [C.c_void_p]
    L.QY265ConfigDefaultPreset.argtypes = [C.POINTER(Config), C.c_char_p, C.c_char_p, C.c_char_p]
    L.QY265ConfigParse.argtypes = [C.POINTER(Config), C.c_char_p, C.c_char_p]
    return L


def test_library_exports_the_reference_symbols():
    L = C.CDLL(LIB)
    assert [s for s in SYMBOLS if not hasattr(L, s)] == []
    assert C.sizeof(Config) == 584          # sizeof(QY265EncConfig) on x86-64, see tests/qy/qy_layout.c


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_INC, "qy265enc.h")), reason="the reference's header only exists in the authoring container")
def test_struct_layout_equals_the_reference_header(tmp_path):
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-std=gnu11", "-Wall", "-I" + REF_INC, "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "qy", "qy_layout.c"), "-o", exe])
    assert "layout ok" in subprocess.check_output([exe], text=True)
    # and a program written against that header links against the shim with nothing else
    subprocess.check_call(["gcc", "-std=gnu11", "-Wall", "-I" + REF_INC, os.path.join(ROOT, "tests", "qy", "qy_caller.c"), "-o", str(tmp_path / "caller"),
                           "-L" + os.path.dirname(LIB), "-lks265qy", "-Wl,-rpath," + os.path.dirname(LIB)])


def test_config_calls():
    L = shim()
    cfg = Config()
    assert L.QY265ConfigDefaultPreset(C.byref(cfg), b"veryfast", None, None) == OK
    assert (cfg.preset, cfg.tune, cfg.latency) == (2, 0, 3)
    assert (cfg.rc, cfg.qp, cfg.crf, cfg.intra_period, cfg.bframes, cfg.sao, cfg.subme, cfg.me) == (2, 26, 24, 128, -1, 3, 2, 0)
    assert L.QY265ConfigDefaultPreset(C.byref(cfg), b"warp9", None, None) < 0
    assert L.QY265ConfigDefaultPreset(C.byref(cfg), b"slow", b"game", b"lowdelay") == OK and (cfg.preset, cfg.tune, cfg.latency, cfg.satd_inter, cfg.me) == (5, 2, 1, 1, 1)
    assert L.QY265ConfigParse(C.byref(cfg), b"picWidth", b"640") == 0 and L.QY265ConfigParse(C.byref(cfg), b"hgt", b"480") == 0 and (cfg.width, cfg.height) == (640, 480)
    assert L.QY265ConfigParse(C.byref(cfg), b"frameRate", b"29.97") == 0 and abs(cfg.fps - 29.97) < 1e-9
    assert L.QY265ConfigParse(C.byref(cfg), b"preset", b"ultrafast") == 0 and (cfg.preset, cfg.subme, cfg.width) == (0, 1, 640)
    assert L.QY265ConfigParse(C.byref(cfg), b"nonsense", b"1") == -1 and L.QY265ConfigParse(C.byref(cfg), b"qp", b"abc") == -2


def test_open_rejects_what_the_device_path_lacks_and_fails_loudly_without_cuda():
    L = shim()
    cfg = Config(); err = C.c_int(0)
    L.QY265ConfigDefaultPreset(C.byref(cfg), b"veryfast", None, None)
    cfg.width, cfg.height = 192, 112                # rc stays at the reference's default, 2 (ABR)
    assert not L.QY265EncoderOpen(C.byref(cfg), C.byref(err)) and err.value == NOTSUPPORTED
    cfg.rc, cfg.width = 0, 191
    assert not L.QY265EncoderOpen(C.byref(cfg), C.byref(err)) and err.value == FAIL
    import torch
    if not torch.cuda.is_available():               # no CPU fallback behind the ABI either
        cfg.width = 192
        assert not L.QY265EncoderOpen(C.byref(cfg), C.byref(err)) and err.value != OK


def _drive(L, yuv, w, h, n, iper, qp, key_at=(), reconfig_at=None, reconfig_qp=None):
    """the demo's loop (encoderwrapper.c:366-400) through ctypes: one reused input buffer, flush while DelayedFrames"""
    cfg = Config(); err = C.c_int(0)
    assert L.QY265ConfigDefaultPreset(C.byref(cfg), b"veryfast", None, None) == OK
    cfg.width, cfg.height, cfg.rc, cfg.qp, cfg.intra_period, cfg.fps = w, h, 0, qp, iper, 30.0
    hnd = L.QY265EncoderOpen(C.byref(cfg), C.byref(err))
    assert hnd and err.value == OK
    fs = w * h * 3 // 2
    buf = np.empty(fs, np.uint8)
    y = Yuv(); y.width, y.height = w, h
    y.plane[0], y.plane[1], y.plane[2] = buf.ctypes.data, buf.ctypes.data + w * h, buf.ctypes.data + w * h * 5 // 4
    y.stride[0], y.stride[1], y.stride[2] = w, w // 2, w // 2
    pic, out = Picture(), Picture(); pic.yuv = C.pointer(y)
    nals, cnt = C.POINTER(Nal)(), C.c_int(0)
    stream, units = bytearray(), []

    def take(r):
        assert r >= 0
        if cnt.value:
            au = b"".join(bytes(bytearray(nals[i].payload[:nals[i].size])) for i in range(cnt.value))
            assert r == len(au)
            stream.extend(au); units.append((out.slice_type, out.poc, out.pts, [nals[i].nal_type for i in range(cnt.value)]))
    hn, hc = C.POINTER(Nal)(), C.c_int(0)
    assert L.QY265EncoderEncodeHeaders(hnd, C.byref(hn), C.byref(hc)) > 0 and [hn[i].nal_type for i in range(hc.value)] == [32, 33, 34]
    for f in range(n):
        if f in key_at:
            L.QY265EncoderKeyFrameRequest(hnd)
        if f == reconfig_at:
            cfg.qp = reconfig_qp
            L.QY265EncoderReconfig(hnd, C.byref(cfg))
        buf[:] = yuv[f * fs:(f + 1) * fs]
        pic.pts = 1000 + f
        take(L.QY265EncoderEncodeFrame(hnd, C.byref(nals), C.byref(cnt), C.byref(pic), C.byref(out), 0))
    while L.QY265EncoderDelayedFrames(hnd):
        take(L.QY265EncoderEncodeFrame(hnd, C.byref(nals), C.byref(cnt), None, C.byref(out), 0))
    L.QY265EncoderClose(hnd)
    return bytes(stream), units


@pytest.mark.gpu
def test_shim_output_equals_the_gop_shards_of_the_encoder_api(tmp_path):
    import gen_yuv
    import ks265codec_b200 as ks
    w, h, n, iper, qp = 192, 112, 10, 4, 30
    yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=11), np.uint8)
    fs = w * h * 3 // 2
    stream, units = _drive(shim(), yuv, w, h, n, iper, qp)
    assert len(units) == n and [u[2] for u in units] == [1000 + f for f in range(n)]
    assert [u[0] for u in units] == [2, 1, 1, 1, 2, 1, 1, 1, 2, 1] and [u[1] for u in units] == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1]
    assert units[0][3] == [32, 33, 34, 19] and units[1][3] == [1]
    want = bytearray()
    cfg = ks.default_config(w, h, preset="veryfast", qp=qp, iper=iper, fps=30.0)
    with ks.Encoder(cfg) as e:
        for s in range(0, n, iper):
            want.extend(bytes(e.encode_gop(yuv[s * fs:min(n, s + iper) * fs])[0]))
    assert stream == bytes(want)
    # a key-frame request cuts the shard short: pictures 0..2 | 3..6 | 7..9
    stream2, units2 = _drive(shim(), yuv, w, h, n, iper, qp, key_at=(3,))
    assert [u[0] for u in units2] == [2, 1, 1, 2, 1, 1, 1, 2, 1, 1] and len(stream2) > 0
    # Reconfig takes effect at the next shard boundary: pictures 0..3 at the old QP, 4..9 at the new one
    stream3, units3 = _drive(shim(), yuv, w, h, n, iper, qp, reconfig_at=1, reconfig_qp=qp + 6)
    want3 = bytearray()
    for q, (a, b) in ((qp, (0, 4)), (qp + 6, (4, 8)), (qp + 6, (8, 10))):
        with ks.Encoder(ks.default_config(w, h, preset="veryfast", qp=q, iper=iper, fps=30.0)) as e:
            want3.extend(bytes(e.encode_gop(yuv[a * fs:b * fs])[0]))
    assert len(units3) == n and stream3 == bytes(want3)
    # the caller compiled against the reference's own header (built where that header exists) writes the same file, and the reference decoder takes it
    if os.path.exists(CALLER):
        clip, out = tmp_path / "in.yuv", tmp_path / "out.265"
        clip.write_bytes(yuv.tobytes())
        r = subprocess.run([CALLER, str(clip), str(w), str(h), str(out), "rc=0", "qp=%d" % qp, "iper=%d" % iper, "fr=30"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert out.read_bytes() == stream
    dec = os.path.join(ROOT, "oracle", "_ref", "appdecoder")
    if os.path.exists(dec):
        bsf, yf = tmp_path / "s.265", tmp_path / "d.yuv"
        bsf.write_bytes(stream)
        subprocess.run([dec, "-b", str(bsf), "-o", str(yf)], capture_output=True, timeout=120)
        d = np.fromfile(yf, np.uint8)
        assert d.size == n * fs
        mse = float(((d.astype(np.int32) - yuv.astype(np.int32)) ** 2).mean())
        assert 10 * np.log10(255.0 ** 2 / mse) > 30.0
