"""Python mirror of the encoder API (include/ks265_enc.h), itself a mirror of the reference's qy265enc.h
(QY265ConfigDefaultPreset / QY265EncoderOpen / QY265EncoderEncodeFrame / QY265EncoderClose).  One Encoder = one
device context = one GOP shard in flight; run several (threads) per GPU for throughput."""
import ctypes as C

import numpy as np


def default_config(width, height, preset="veryfast", qp=27, iper=128, device=0, **kw):
    from . import Ks265Config, lib
    cfg = Ks265Config()
    cfg.width, cfg.height = width, height
    if lib().ks265_config_default_preset(C.byref(cfg), preset.encode()) != 0:
        raise ValueError("unknown preset %r" % preset)
    cfg.qp, cfg.iper, cfg.device = qp, iper, device
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError("unknown config field %r" % k)
        setattr(cfg, k, v)
    return cfg


class Encoder:
    """with Encoder(cfg) as e: bitstream, recon, stats = e.encode_gop(frames_u8)"""

    def __init__(self, cfg):
        from . import lib
        self._lib = lib()
        self.cfg = cfg
        err = C.c_int(0)
        self._h = self._lib.ks265_encoder_open(C.byref(cfg), C.byref(err))
        if not self._h:
            raise RuntimeError("ks265_encoder_open failed (error %d): the hot path needs a CUDA device, there is no CPU fallback" % err.value)
        self.frame_bytes = cfg.width * cfg.height * 3 // 2

    def close(self):
        if self._h:
            self._lib.ks265_encoder_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def encode_gop(self, frames, want_recon=False, device_ptr=None, nframes=None, out=None):
        """frames: bytes / uint8 ndarray of n display-size I420 pictures in HOST memory, or device_ptr (int) to the same
        layout in DEVICE memory.  Returns (bitstream bytes, recon ndarray|None, stats)."""
        from . import Ks265GopStats
        if device_ptr is None:
            arr = np.frombuffer(frames, dtype=np.uint8) if not isinstance(frames, np.ndarray) else frames
            n = arr.size // self.frame_bytes if nframes is None else nframes
            hp = C.c_void_p(arr.ctypes.data)
            dp = None
        else:
            n = nframes
            hp, dp = None, C.c_void_p(device_ptr)
        cap = self.frame_bytes * n * 2 + (1 << 20)      # CABAC worst case (noise at QP 0) is ~2.1 bytes per luma sample = 1.4 x the picture bytes
        bs = out if out is not None else np.empty(cap, np.uint8)
        rec = np.empty(self.frame_bytes * n, np.uint8) if want_recon else None
        st = Ks265GopStats()
        r = self._lib.ks265_encoder_encode_gop(self._h, hp, dp, n, C.c_void_p(bs.ctypes.data), bs.size,
                                               C.c_void_p(rec.ctypes.data) if want_recon else None, C.byref(st))
        if r < 0:
            raise RuntimeError("ks265_encoder_encode_gop failed: %d" % r)
        return bs[:r], rec, st

    def run_gop_device(self, device_ptr, nframes):
        """device pipeline only (no entropy coding): inputs already in HBM; returns stats"""
        from . import Ks265GopStats
        st = Ks265GopStats()
        r = self._lib.ks265_encoder_run_gop_device(self._h, C.c_void_p(device_ptr), nframes, C.byref(st))
        if r < 0:
            raise RuntimeError("ks265_encoder_run_gop_device failed: %d" % r)
        return st

    STAGES = ("me", "recon_inter", "recon_intra", "deblock", "sao", "pack", "decide", "intra_p")

    def set_profiling(self, on=True):
        self._lib.ks265_encoder_set_profiling(self._h, int(on))

    def stage_times(self):
        """{stage: (total ms, launches of the stage)} measured with CUDA events on the encoder's stream"""
        ms = (C.c_double * len(self.STAGES))(); n = (C.c_uint64 * len(self.STAGES))()
        self._lib.ks265_encoder_get_stage_times(self._h, ms, n)
        return {k: (ms[i], int(n[i])) for i, k in enumerate(self.STAGES)}
