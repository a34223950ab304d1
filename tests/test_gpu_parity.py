"""GPU parity tests: the CUDA hot path (through the C-ABI of include/ks265_gpu.h / ks265_enc.h) against the CPU
oracle (oracle/, pinned to the reference binary by tests/test_oracle_kat.py) and against the REFERENCE DECODER
(oracle/_ref/appdecoder, SURVEY.md 8c tier P1).  Integer work: every comparison is bit-exact."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from katlib import ROOT, OraCfg, SeqCfg, oracle, ptr

sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_yuv  # noqa: E402

pytestmark = pytest.mark.gpu

import ks265codec_b200 as ks  # noqa: E402

DEC = os.path.join(ROOT, "oracle", "_ref", "appdecoder")


def first_diff(a, b, what, shape=None):
    a = np.asarray(a).ravel(); b = np.asarray(b).ravel()
    assert a.size == b.size, "%s: size %d vs %d" % (what, a.size, b.size)
    d = np.nonzero(a != b)[0]
    if d.size:
        loc = int(d[0])
        where = divmod(loc, shape[1]) if shape else loc
        pytest.fail("%s: %d mismatches, first at %s: gpu=%s oracle=%s" % (what, d.size, where, a[loc], b[loc]))


def gpu_cfg(subpel=2, sbh=1, sao=1, iters=16, satd=0, me=0):
    g = ks.KsGpuCfg()
    g.me_range, g.me_iters, g.subpel, g.sign_hiding, g.sao, g.strong_intra = 64, iters, subpel, sbh, sao, 1
    g.n_src_slots, g.n_rec_slots, g.n_syn_slots, g.satd, g.me_method = 3, 2, 2, satd, me
    return g


# ---------------------------------------------------------------------------------------- leaf KAT replays
def test_kat_sad16():
    L, O = ks.lib(), oracle()
    rng = np.random.default_rng(1)
    for _ in range(8):
        a = rng.integers(0, 256, (16, 24), dtype=np.uint8); b = rng.integers(0, 256, (16, 40), dtype=np.uint8)
        out = np.zeros(1, np.uint32)
        assert L.ks_gpu_kat_sad16(ptr(a), ptr(b), 24, 40, ptr(out)) == 0
        assert int(out[0]) == O.ora_sad(ptr(a), ptr(b), 24, 40, 16, 16)


def test_kat_satd16():
    L, O = ks.lib(), oracle()
    rng = np.random.default_rng(3)
    for amp in (255, 40, 6, 1):
        a = rng.integers(0, 256, (16, 16), dtype=np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-amp, amp + 1, (16, 16)), 0, 255).astype(np.uint8)
        out = np.zeros(1, np.uint32)
        assert L.ks_gpu_kat_satd16(ptr(a), ptr(b), 16, 16, ptr(out)) == 0
        assert int(out[0]) == O.ora_satd(ptr(a), ptr(b), 16, 16, 16, 16)


def test_kat_interp_luma16_all_phases_and_borders():
    L, O = ks.lib(), oracle()
    rng = np.random.default_rng(2)
    w, h, pad = 64, 48, 96
    plane = rng.integers(0, 256, (h, w), dtype=np.uint8)
    padded = np.pad(plane, pad, mode="edge")
    for (x, y) in ((16, 16), (0, 0), (48, 32), (40, 8)):
        for mvx in (-9, -3, -2, -1, 0, 1, 2, 3, 6, 11, 70):
            for mvy in (-70, -5, -1, 0, 1, 2, 3, 7):
                got = np.zeros((16, 16), np.uint8); exp = np.zeros((16, 16), np.uint8)
                assert L.ks_gpu_kat_interp_luma16(ptr(plane), w, h, x, y, mvx, mvy, ptr(got)) == 0
                O.ora_mc_luma(ptr(exp), 16, ptr(padded, (pad + y) * padded.shape[1] + pad + x), padded.shape[1], 16, 16, mvx, mvy)
                first_diff(got, exp, "interp x=%d y=%d mv=(%d,%d)" % (x, y, mvx, mvy), (16, 16))


@pytest.mark.parametrize("log2n", [2, 3, 4, 5])
def test_kat_transform_block(log2n):
    """fdct -> quant -> (sign hiding) -> dequant -> idct+pred on the device == oracle chain (itself == reference KATs)"""
    L, O = ks.lib(), oracle()
    rng = np.random.default_rng(10 + log2n)
    n = 1 << log2n
    O.ora_quant.restype = C.c_int; O.ora_sign_hide.restype = C.c_int
    scan = build_scan(log2n)
    for trial in range(24):
        qp = int(rng.integers(10, 46)); intra = trial & 1; sbh = (trial >> 1) & 1
        pred = rng.integers(0, 256, (n, n), dtype=np.uint8)
        amp = (2, 8, 40, 120)[trial % 4]
        src = np.clip(pred.astype(int) + rng.integers(-amp, amp + 1, (n, n)), 0, 255).astype(np.uint8)
        lev = np.zeros((n, n), np.int16); rec = np.zeros((n, n), np.uint8); cbf = np.zeros(1, np.int32)
        assert L.ks_gpu_kat_tb(log2n, ptr(src), ptr(pred), qp, intra, sbh, ptr(lev), ptr(rec), ptr(cbf)) == 0
        res = np.zeros(n * n, np.int16); coef = np.zeros(n * n, np.int16); q = np.zeros(n * n, np.int16); du = np.zeros(n * n, np.int16)
        O.ora_residual(ptr(res), ptr(src), ptr(pred), n, n, n)
        O.ora_fdct(ptr(res), ptr(coef), n, n, log2n, 0)
        nnz = O.ora_quant(ptr(coef), ptr(q), n, qp, log2n, intra, ptr(du))
        if nnz and sbh:
            nnz = O.ora_sign_hide(ptr(coef), ptr(q), ptr(du), n, log2n, ptr(scan))
        exp = pred.copy()
        if nnz:
            deq = np.zeros(n * n, np.int16)
            O.ora_dequant(ptr(q), ptr(deq), n, qp, log2n)
            O.ora_idct_add(ptr(deq), ptr(exp), ptr(pred), n, n, n, log2n, 0)
        first_diff(lev, q, "levels log2n=%d trial=%d qp=%d" % (log2n, trial, qp), (n, n))
        first_diff(rec, exp, "recon log2n=%d trial=%d" % (log2n, trial), (n, n))
        assert int(cbf[0]) == int(nnz > 0)


def build_scan(log2n):
    def diag(n):
        out, x, y = [], 0, 0
        while len(out) < n * n:
            while y >= 0:
                if x < n and y < n:
                    out.append((x, y))
                y -= 1; x += 1
            y, x = x, 0
        return out
    d4, dcg = diag(4), diag(1 << (log2n - 2))
    return np.array([(((cy << 2) + py) << 8) | ((cx << 2) + px) for (cx, cy) in dcg for (px, py) in d4], np.uint16)


def test_harvested_reference_vectors_on_the_device():
    """the vectors harvested from the REFERENCE binary's own sad_c / had_c / interpLuma{Hor,Ver}8to8_c (tests/golden/kat_*.bin.gz) replayed
    on the device entry points: device output == the reference's output, with no oracle in between"""
    from katlib import GOLDEN, read_kat
    L = ks.lib()
    n_sad = n_had = n_int = 0
    for name, p, ins, outs in read_kat(os.path.join(GOLDEN, "kat_sad.bin.gz")):
        if name not in ("sad", "had") or tuple(p[:2]) != (16, 16):
            continue
        w, h, sa, sb = p
        a = np.frombuffer(ins[0], np.uint8).copy(); b = np.frombuffer(ins[1], np.uint8).copy()
        out = np.zeros(1, np.uint32)
        fn = L.ks_gpu_kat_sad16 if name == "sad" else L.ks_gpu_kat_satd16
        assert fn(ptr(a), ptr(b, sb + 1), sa, sb, ptr(out)) == 0
        assert int(out[0]) == int(np.frombuffer(outs[0], np.uint32)[0]), (name, p)
        n_sad += name == "sad"; n_had += name == "had"
    for name, p, ins, outs in read_kat(os.path.join(GOLDEN, "kat_interp.bin.gz")):
        if name not in ("luma_h_8to8", "luma_v_8to8") or p[0] != 32:
            continue
        w, h, frac, ss, ds = p
        src = np.frombuffer(ins[0], np.uint8).copy().reshape(-1, ss)              # 80 x 80 plane; the block's sample (0,0) sits at (4,4)
        ref = np.frombuffer(outs[0], np.uint8)[:h * ds].reshape(h, ds)[:, :w]
        plane = np.ascontiguousarray(src[:80, :80])
        for (bx, by) in ((0, 0), (16, 16), (8, 4)):           # 16x16 sub-blocks of the 32x32 result; all taps stay inside the harvested buffer
            got = np.zeros((16, 16), np.uint8)
            mvx, mvy = (frac, 0) if name == "luma_h_8to8" else (0, frac)
            assert L.ks_gpu_kat_interp_luma16(ptr(plane), 80, 80, 4 + bx, 4 + by, mvx, mvy, ptr(got)) == 0
            first_diff(got, ref[by:by + 16, bx:bx + 16], "%s frac %d block (%d,%d) vs the reference binary" % (name, frac, bx, by), (16, 16))
            n_int += 1
    assert n_sad >= 2 and n_had >= 2 and n_int >= 18


# ---------------------------------------------------------------------------------------- picture stages
def run_oracle_picture(cfg, slice_type, qp, boff, toff, src, ref, prev_cells, lambda_qp=None):
    O = oracle()
    O.ora_run_picture.restype = C.c_long
    W, H = cfg.width, cfg.height
    ncell, nctu = (W >> 4) * (H >> 4), ((W + 63) >> 6) * ((H + 63) >> 6)
    fsz = W * H * 3 // 2
    o = dict(pre=np.zeros(fsz, np.uint8), fin=np.zeros(fsz, np.uint8), cells=np.zeros(ncell * 8, np.uint8),
             lev=np.zeros(fsz, np.int16), ctus=np.zeros(nctu * 72, np.uint8), pool=np.zeros(fsz, np.int16))
    n = O.ora_run_picture(C.byref(cfg), slice_type, qp, min(51, qp if lambda_qp is None else lambda_qp), boff, toff, ptr(src), ptr(ref) if ref is not None else None,
                          ptr(prev_cells) if prev_cells is not None else None, ptr(o["pre"]), ptr(o["fin"]), ptr(o["cells"]),
                          ptr(o["lev"]), ptr(o["ctus"]), ptr(o["pool"]))
    assert n >= 0
    o["n_cg"] = n
    return o


@pytest.mark.parametrize("w,h,qp,sbh,sao,subpel,satd,me", [(192, 112, 32, 1, 1, 2, 0, 0), (200, 120, 27, 1, 1, 2, 0, 0), (320, 240, 24, 0, 0, 1, 0, 0), (256, 128, 37, 1, 4, 0, 0, 0), (272, 144, 29, 1, 3, 2, 1, 0), (336, 208, 30, 1, 4, 2, 0, 0),
                                                            (352, 208, 28, 1, 3, 2, 0, 1), (208, 128, 33, 1, 1, 2, 1, 1)])
def test_picture_stages_match_oracle(w, h, qp, sbh, sao, subpel, satd, me):
    """I picture then 3 P pictures: every stage output of the device (ME field, pre-filter recon, dense levels, final
    recon after deblock+SAO, SAO parameters, CG bitmaps, packed level pool) equals the CPU model."""
    L = ks.lib()
    nfr = 4
    yuv = np.frombuffer(gen_yuv.make(w, h, nfr, seed=7), np.uint8)
    g = gpu_cfg(subpel, sbh, sao, satd=satd, me=me)
    err = C.c_int(0)
    ctx = L.ks_gpu_open(0, w, h, C.byref(g), C.byref(err))
    assert ctx, "ks_gpu_open failed: %d" % err.value
    try:
        cw_, ch_ = C.c_int(0), C.c_int(0)
        L.ks_gpu_coded_size(ctx, C.byref(cw_), C.byref(ch_))
        W, H = cw_.value, ch_.value
        assert (W, H) == ((w + 15) & ~15, (h + 15) & ~15)
        cfg = OraCfg(W, H, 64, 16, subpel, sbh, sao, 1, satd, me)
        fsz, dsz = W * H * 3 // 2, w * h * 3 // 2
        ncell, nctu = (W >> 4) * (H >> 4), ((W + 63) >> 6) * ((H + 63) >> 6)
        ref_fin = None; prev_cells = None
        for f in range(nfr):
            fr = yuv[f * dsz:(f + 1) * dsz]
            y, u, v = fr[:w * h], fr[w * h:w * h * 5 // 4], fr[w * h * 5 // 4:]
            assert L.ks_gpu_upload_frame(ctx, f % 3, ptr(y), ptr(u), ptr(v), w, w // 2) == 0
            src = np.zeros(fsz, np.uint8)
            is_i = f == 0
            pp = ks.KsPicParams(ks.KS_SLICE_I if is_i else ks.KS_SLICE_P, qp if is_i else qp + 1, f % 3, -1 if is_i else (f & 1) ^ 1, f & 1, f & 1,
                                -1 if is_i else (f & 1) ^ 1, 0 if is_i else 2, 0 if is_i else 2, 1)
            pp.lambda_qp_delta = 3 if f == 2 else 0          # one P picture decides with the raised lambda (non-key pictures of the cascade)
            if not is_i:
                me = np.zeros(ncell * 8, np.uint8)
                assert L.ks_gpu_debug_me(ctx, C.byref(pp), ptr(me)) == 0
            out = ks.KsPicOut()
            assert L.ks_gpu_encode_picture(ctx, C.byref(pp), C.byref(out)) == 0
            assert L.ks_gpu_debug_fetch(ctx, 2, f % 3, ptr(src), fsz) == 0           # coded-size source as the device sees it
            o = run_oracle_picture(cfg, pp.slice_type, pp.qp, pp.beta_offset_div2, pp.tc_offset_div2, src, ref_fin, prev_cells, pp.qp + pp.lambda_qp_delta)
            tag = "%dx%d f%d " % (w, h, f)
            if not is_i:
                O = oracle()
                ome = np.zeros(ncell * 8, np.uint8)
                O.ora_run_me(C.byref(cfg), pp.qp, ptr(src), ptr(ref_fin), ptr(prev_cells) if prev_cells is not None else None, ptr(ome))
                first_diff(me.view(np.int16).reshape(-1, 4)[:, :2], ome.view(np.int16).reshape(-1, 4)[:, :2], tag + "ME field (cell, comp)", (ncell, 2))
            cells = np.ctypeslib.as_array(C.cast(out.cells, C.POINTER(C.c_uint8)), (ncell * 8,)).copy()
            first_diff(cells.reshape(-1, 8), o["cells"].reshape(-1, 8), tag + "cells (cell, byte)", (ncell, 8))
            pre = np.zeros(fsz, np.uint8)
            assert L.ks_gpu_debug_fetch(ctx, 0, 0, ptr(pre), fsz) == 0
            lev = np.zeros(fsz, np.int16)
            assert L.ks_gpu_debug_fetch(ctx, 1, 0, ptr(lev), fsz * 2) == 0
            first_diff(lev[:W * H], o["lev"][:W * H], tag + "luma levels (y,x)", (H, W))
            first_diff(lev[W * H:], o["lev"][W * H:], tag + "chroma levels", (H, W // 2))
            # the device deblocks in place, so its "pre" buffer holds the DEBLOCKED picture; compare the final instead
            fin = np.zeros(fsz, np.uint8)
            if (W, H) == (w, h):
                assert L.ks_gpu_fetch_recon(ctx, f & 1, ptr(fin), ptr(fin, W * H), ptr(fin, W * H * 5 // 4), W, W // 2) == 0
                first_diff(fin[:W * H], o["fin"][:W * H], tag + "final luma (y,x)", (H, W))
                first_diff(fin[W * H:], o["fin"][W * H:], tag + "final chroma", (H, W // 2))
            ctus = np.ctypeslib.as_array(C.cast(out.ctus, C.POINTER(C.c_uint8)), (nctu * 72,)).copy().reshape(-1, 72)
            octus = o["ctus"].reshape(-1, 72)
            first_diff(ctus[:, 52:70], octus[:, 52:70], tag + "SAO params (ctu, byte)", (nctu, 18))
            first_diff(ctus[:, :52], octus[:, :52], tag + "CG bitmaps/base (ctu, byte)", (nctu, 52))
            assert out.n_cg == o["n_cg"], tag + "n_cg %d vs %d" % (out.n_cg, o["n_cg"])
            pool = np.ctypeslib.as_array(out.levels, (max(out.n_cg, 1) * 16,)).copy()[:out.n_cg * 16]
            first_diff(pool, o["pool"][:out.n_cg * 16], tag + "level pool")
            sse = [int(out.sse[k]) for k in range(3)]
            # the device's SSE covers the DISPLAY area (the reference's PSNR does too), not the padded coded picture
            osse = []
            for (a, b), (pw, ph), (dw, dh) in zip(((0, W * H), (W * H, W * H * 5 // 4), (W * H * 5 // 4, fsz)), ((W, H), (W // 2, H // 2), (W // 2, H // 2)),
                                                  ((w, h), (w // 2, h // 2), (w // 2, h // 2))):
                d = (o["fin"][a:b].astype(np.int64) - src[a:b]).reshape(ph, pw)[:dh, :dw]
                osse.append(int((d ** 2).sum()))
            assert sse == osse, tag + "SSE %s vs %s" % (sse, osse)
            ref_fin = o["fin"]; prev_cells = o["cells"]
    finally:
        L.ks_gpu_close(ctx)


# ---------------------------------------------------------------------------------------- whole encoder
def oracle_encode(yuv, w, h, n, qp, iper, subpel=2, sbh=1, sao=1, iters=16, satd=0, bframes=0, me=0, rc=0, crf=24.0):
    O = oracle()
    O.ora_encode_sequence.restype = C.c_long
    cfg = SeqCfg(w, h, n, qp, iper, 0, 64, iters, subpel, sbh, sao, 3, satd, bframes, me, rc, int(round(crf * 100)))
    bs = np.zeros(w * h * 3 * n + 100000, np.uint8); rec = np.zeros(w * h * 3 // 2 * n, np.uint8)
    nb = O.ora_encode_sequence(C.byref(cfg), ptr(yuv), ptr(bs), C.c_size_t(bs.size), ptr(rec))
    assert nb > 0
    return bs[:nb], rec


def decode_with_reference(bs, nbytes_expected):
    if not os.path.exists(DEC):
        pytest.skip("oracle/_ref/appdecoder not staged")
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.265"); o = os.path.join(d, "t.yuv")
        open(p, "wb").write(bytes(bs))
        r = subprocess.run([DEC, "-b", p, "-o", o, "-threads", "1"], capture_output=True, text=True, timeout=600)
        assert os.path.exists(o), "reference decoder produced nothing: " + r.stdout[-300:] + r.stderr[-300:]
        return np.fromfile(o, np.uint8)


@pytest.mark.parametrize("w,h,n,qp,preset", [(192, 112, 5, 32, "veryfast"), (416, 240, 6, 27, "veryfast"), (200, 120, 4, 30, "superfast"), (1280, 720, 5, 32, "superfast"), (352, 288, 5, 28, "fast"), (320, 176, 4, 31, "medium"), (1920, 1080, 3, 27, "veryfast")])
def test_encoder_bitstream_equals_oracle_and_decodes(w, h, n, qp, preset):
    """ks265_encoder_encode_gop: (1) Annex-B bytes == CPU model's bytes, (2) recon == model recon,
    (3) the reference decoder's output of OUR stream == our recon (the vendor's own -hm style self-test)."""
    yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=11), np.uint8)
    cfg = ks.default_config(w, h, preset=preset, qp=qp, iper=n, psnr=1)
    with ks.Encoder(cfg) as e:
        bs, rec, st = e.encode_gop(yuv, want_recon=True)
    obs, orec = oracle_encode(yuv, w, h, n, qp, n, subpel=cfg.subpel, sbh=cfg.sign_hiding, sao=cfg.sao, iters=cfg.me_iters, satd=cfg.satd)
    first_diff(rec, orec, "recon vs oracle")
    assert bytes(bs) == bytes(obs), "bitstream differs from the CPU model (%d vs %d bytes)" % (bs.size, obs.size)
    dec = decode_with_reference(bs, rec.size)
    first_diff(dec, rec, "reference decoder output vs our recon")
    assert st.frames == n and st.gpu_launches > 0


@pytest.mark.parametrize("w,h,n,qp,preset,bf", [(192, 112, 6, 32, "veryfast", 1), (416, 240, 8, 27, "veryfast", 3), (320, 176, 7, 30, "medium", 2), (1280, 720, 6, 30, "veryfast", 2)])
def test_encoder_bframes_equal_oracle_and_decode(w, h, n, qp, preset, bf):
    """-bframes n (IDR, P anchors, non-reference B pictures between them): same three checks as the P-only streams.  The reference
    decoder reorders to display order, so its output is compared against our display-order reconstruction."""
    yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=5), np.uint8)
    cfg = ks.default_config(w, h, preset=preset, qp=qp, iper=n, psnr=1, bframes=bf)
    with ks.Encoder(cfg) as e:
        bs, rec, st = e.encode_gop(yuv, want_recon=True)
    obs, orec = oracle_encode(yuv, w, h, n, qp, n, subpel=cfg.subpel, sbh=cfg.sign_hiding, sao=cfg.sao, iters=cfg.me_iters, satd=cfg.satd, bframes=bf)
    first_diff(rec, orec, "recon vs oracle")
    assert bytes(bs) == bytes(obs), "bitstream differs from the CPU model (%d vs %d bytes)" % (bs.size, obs.size)
    dec = decode_with_reference(bs, rec.size)
    first_diff(dec, rec, "reference decoder output vs our recon")
    assert st.frames == n


@pytest.mark.parametrize("w,h,n,preset,bf,me,rc,crf", [(416, 240, 6, "veryfast", 0, 1, 0, 0.0), (320, 176, 8, "veryfast", 0, 0, 3, 26.0), (352, 288, 9, "slow", 2, 1, 3, 24.0), (1280, 720, 6, "slow", 0, 1, 3, 24.0)])
def test_encoder_hex_and_crf_equal_oracle_and_decode(w, h, n, preset, bf, me, rc, crf):
    """-me 1 (hexagon search) and -rc 3 -crf (host rate control fed by the device's search cost): stream bytes == CPU model,
    recon == model, reference decoder output == recon.  BASELINE configs[3] uses -preset slow -rc 3 -crf 24."""
    yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=9), np.uint8)
    cfg = ks.default_config(w, h, preset=preset, qp=30, iper=n, psnr=1, bframes=bf, me=me, rc=rc, crf=crf)
    with ks.Encoder(cfg) as e:
        bs, rec, st = e.encode_gop(yuv, want_recon=True)
    obs, orec = oracle_encode(yuv, w, h, n, 30, n, subpel=cfg.subpel, sbh=cfg.sign_hiding, sao=cfg.sao, iters=cfg.me_iters, satd=cfg.satd, bframes=bf, me=me, rc=rc, crf=crf)
    first_diff(rec, orec, "recon vs oracle")
    assert bytes(bs) == bytes(obs), "bitstream differs from the CPU model (%d vs %d bytes)" % (bs.size, obs.size)
    dec = decode_with_reference(bs, rec.size)
    first_diff(dec, rec, "reference decoder output vs our recon")


def test_natural_clip_closed_loop():
    """natural content (the reference repo's own 640x480 clip, centre 320x240 crop of frames 8..13, staged under tests/golden) through the decoder"""
    clip = os.path.join(ROOT, "tests", "golden", "nat_320x240_6f.yuv.gz")
    if not os.path.exists(clip):
        pytest.skip("natural clip fixture not present")
    import gzip
    yuv = np.frombuffer(gzip.open(clip, "rb").read(), np.uint8)
    w, h, n = 320, 240, 6
    cfg = ks.default_config(w, h, preset="veryfast", qp=30, iper=n)
    with ks.Encoder(cfg) as e:
        bs, rec, st = e.encode_gop(yuv, want_recon=True)
    dec = decode_with_reference(bs, rec.size)
    first_diff(dec, rec, "reference decoder output vs our recon (natural clip)")
    obs, orec = oracle_encode(yuv, w, h, n, 30, n)
    assert bytes(bs) == bytes(obs)


def test_4k_roundtrip_property():
    """BASELINE size (3840x2160): size-independent properties instead of the slow CPU model:
    our stream decodes with the reference decoder to exactly our reconstruction, and PSNR is sane."""
    w, h, n = 3840, 2160, 3
    yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=5), np.uint8)
    cfg = ks.default_config(w, h, preset="veryfast", qp=27, iper=128, psnr=1)
    with ks.Encoder(cfg) as e:
        bs, rec, st = e.encode_gop(yuv, want_recon=True)
        bs2, rec2, _ = e.encode_gop(yuv, want_recon=True)
    assert bytes(bs) == bytes(bs2) and np.array_equal(rec, rec2), "encoder is not deterministic run to run"
    dec = decode_with_reference(bs, rec.size)
    first_diff(dec, rec, "4K: reference decoder output vs our recon")
    mse = ((rec[:w * h].astype(np.float64) - yuv[:w * h]) ** 2).mean()
    assert 10 * np.log10(255 * 255 / mse) > 33.0
    sse_y = int(((rec.reshape(n, -1)[:, :w * h].astype(np.int64) - yuv.reshape(n, -1)[:, :w * h]) ** 2).sum())
    assert st.sse[0] == sse_y          # device-side SSE (PSNR line) agrees with the host recomputation


def _bench_sequence(w, h, n):
    """the benchmark's own input: bench.py's ping-pong shard of 16 generated pictures"""
    sys.path.insert(0, ROOT)
    import bench
    frames = [np.frombuffer(fr, np.uint8) for fr in gen_yuv.frames(w, h, bench.DISTINCT, seed=1234)]
    return np.concatenate([frames[i] for i in bench.shard_order(n)])


def test_4k_full_shard_decodes_to_recon():
    """the benchmarked workload at full size: one whole 128-picture 3840x2160 GOP shard (bench.py's sequence, POC lsb wraps at 256 not reached,
    QP cascade, raised-lambda pictures) -> reference decoder output == our reconstruction, byte for byte"""
    w, h, n = 3840, 2160, 128
    yuv = _bench_sequence(w, h, n)
    cfg = ks.default_config(w, h, preset="veryfast", qp=27, iper=128, psnr=1)
    with ks.Encoder(cfg) as e:
        bs, rec, st = e.encode_gop(yuv, want_recon=True)
    assert st.frames == n
    dec = decode_with_reference(bs, rec.size)
    assert dec.size == rec.size
    for f in range(n):           # picture by picture: a failure names the first bad POC
        a, b = dec[f * (w * h * 3 // 2):(f + 1) * (w * h * 3 // 2)], rec[f * (w * h * 3 // 2):(f + 1) * (w * h * 3 // 2)]
        assert np.array_equal(a, b), "4K shard: reference decoder output != our recon at POC %d" % f
    sse_y = sum(int(((rec[f * (w * h * 3 // 2):f * (w * h * 3 // 2) + w * h].astype(np.int64) - yuv[f * (w * h * 3 // 2):f * (w * h * 3 // 2) + w * h]) ** 2).sum()) for f in range(n))
    assert st.sse[0] == sse_y
    assert 10 * np.log10(255.0 ** 2 * w * h * n / sse_y) > 34.0


def test_1080p_long_shard_equals_oracle_and_decodes():
    """BASELINE configs[1] size: 128-picture 1920x1080 shard -> reference decoder == recon; its first 20 pictures == the CPU model's bytes
    (the model needs ~2.5 s per 1080p picture, so the bit-exact leg is bounded)"""
    w, h, n, m = 1920, 1080, 128, 20
    yuv = _bench_sequence(w, h, n)
    cfg = ks.default_config(w, h, preset="veryfast", qp=27, iper=128, psnr=1)
    with ks.Encoder(cfg) as e:
        bs, rec, st = e.encode_gop(yuv, want_recon=True)
        bsm, recm, _ = e.encode_gop(yuv[:m * w * h * 3 // 2], want_recon=True)
    dec = decode_with_reference(bs, rec.size)
    first_diff(dec, rec, "1080p shard: reference decoder output vs our recon")
    obs, orec = oracle_encode(yuv[:m * w * h * 3 // 2], w, h, m, 27, 128, subpel=cfg.subpel, sbh=cfg.sign_hiding, sao=cfg.sao, iters=cfg.me_iters, satd=cfg.satd)
    first_diff(recm, orec, "1080p recon vs oracle")
    assert bytes(bsm) == bytes(obs)
    assert np.array_equal(rec[:recm.size], recm), "a shard's first pictures do not depend on its length"


def test_encoder_recovers_after_output_overflow():
    """an encode_gop that fails with -28 (output buffer too small) must not poison the handle: the pictures in flight are dropped
    and the next call works"""
    w, h, n = 416, 240, 6
    yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=13), np.uint8)
    cfg = ks.default_config(w, h, preset="veryfast", qp=22, iper=n)
    with ks.Encoder(cfg) as e:
        good, rec, _ = e.encode_gop(yuv, want_recon=True)
        small = np.empty(max(4096, good.size // 3), np.uint8)
        with pytest.raises(RuntimeError):
            e.encode_gop(yuv, out=small)
        again, rec2, _ = e.encode_gop(yuv, want_recon=True)
    assert bytes(again) == bytes(good) and np.array_equal(rec, rec2)


def test_8k_roundtrip_property():
    """BASELINE configs[4] size (7680x4320): one IDR + one P picture decode with the reference decoder to exactly our recon"""
    w, h, n = 7680, 4320, 2
    yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=8), np.uint8)
    cfg = ks.default_config(w, h, preset="veryfast", qp=27, iper=128, psnr=1)
    with ks.Encoder(cfg) as e:
        bs, rec, st = e.encode_gop(yuv, want_recon=True)
    dec = decode_with_reference(bs, rec.size)
    first_diff(dec, rec, "8K: reference decoder output vs our recon")
    mse = ((rec[:w * h].astype(np.float64) - yuv[:w * h]) ** 2).mean()
    assert 10 * np.log10(255 * 255 / mse) > 33.0


def test_cli_matches_library_and_reference_decoder(tmp_path):
    """the appencoder-compatible CLI: flags, summary lines, -b stream decodable by the reference decoder == its own -o"""
    w, h, n = 416, 240, 9
    yuv = gen_yuv.make(w, h, n, seed=4)
    src, bs, rec = tmp_path / "in.yuv", tmp_path / "o.265", tmp_path / "o.yuv"
    src.write_bytes(yuv)
    r = subprocess.run([ks.CLI_PATH, "-i", str(src), "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-preset", "veryfast", "-rc", "0", "-qp", "30",
                        "-iper", "4", "-b", str(bs), "-o", str(rec), "-psnr", "1", "-md5", "1", "-threads", "1", "-streams", "2"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    assert "Total Frames: %d" % n in r.stdout and "bitrate, psnr:" in r.stdout and "H265 encoder passed!!!" in r.stdout and "POC 0 MD5" in r.stdout
    ours = np.fromfile(rec, np.uint8)
    dec = decode_with_reference(np.fromfile(bs, np.uint8), ours.size)
    first_diff(dec, ours, "CLI: reference decoder output vs -o recon")
    # three closed GOP shards (4+4+1 pictures) concatenated in order == the model's stream for the same settings
    obs, orec = oracle_encode(np.frombuffer(yuv, np.uint8), w, h, n, 30, 4)
    assert bytes(np.fromfile(bs, np.uint8)) == bytes(obs)
