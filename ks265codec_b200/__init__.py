"""ks265codec_b200 -- B200-native HEVC encode hot path behind the KSC265 (ksvc/ks265codec) appencoder surface.

Python is plumbing only: ctypes bindings of the C-ABI in include/ks265_gpu.h / include/ks265_enc.h
(libks265gpu.so: hand-written sm_100a CUDA kernels + the C host encoder), the GOP-shard scheduler and the
NCCL gather of NAL units for multi-GPU runs.  There is NO CPU fallback: opening an encoder without a CUDA
device raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libks265gpu.so")
CLI_PATH = os.path.join(_HERE, "bin", "appencoder")

KS_SLICE_B, KS_SLICE_P, KS_SLICE_I = 0, 1, 2


class KsCell(C.Structure):
    _fields_ = [("mvx", C.c_int16), ("mvy", C.c_int16), ("cu_log2", C.c_uint8), ("flags", C.c_uint8),
                ("intra_mode", C.c_uint8), ("rsv", C.c_uint8)]


class KsSaoParam(C.Structure):
    _fields_ = [("type", C.c_uint8), ("band_or_class", C.c_uint8), ("off", C.c_int8 * 4)]


class KsCtuSyn(C.Structure):
    _fields_ = [("cg_y", C.c_uint16 * 16), ("cg_cb", C.c_uint8 * 8), ("cg_cr", C.c_uint8 * 8), ("cg_base", C.c_uint32),
                ("sao", KsSaoParam * 3), ("rsv", C.c_uint8 * 2)]


class KsGpuCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("me_range", "me_iters", "subpel", "sign_hiding", "sao", "strong_intra",
                                       "n_src_slots", "n_rec_slots", "n_syn_slots", "satd", "me_method")]


class KsPicParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("slice_type", "qp", "src_slot", "ref_slot", "out_slot", "syn_slot", "prev_syn_slot",
                                       "beta_offset_div2", "tc_offset_div2", "want_sse", "ref1_slot", "dist_l0", "dist_anchor", "want_me_cost", "lambda_qp_delta")]


class KsCellB(C.Structure):
    _fields_ = [("mvx1", C.c_int16), ("mvy1", C.c_int16), ("dir", C.c_uint8), ("rsv", C.c_uint8 * 3)]


class KsPicOut(C.Structure):
    _fields_ = [("cells", C.POINTER(KsCell)), ("ctus", C.POINTER(KsCtuSyn)), ("levels", C.POINTER(C.c_int16)),
                ("n_cg", C.c_uint32), ("sse", C.c_uint64 * 3), ("cells_b", C.POINTER(KsCellB)), ("me_cost", C.c_uint64)]


class Ks265Config(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fps", C.c_double), ("preset", C.c_int), ("rc", C.c_int),
                ("qp", C.c_int), ("iper", C.c_int), ("fixqp", C.c_int), ("sao", C.c_int), ("sign_hiding", C.c_int),
                ("me_range", C.c_int), ("me_iters", C.c_int), ("subpel", C.c_int), ("satd", C.c_int), ("device", C.c_int), ("psnr", C.c_int), ("bframes", C.c_int), ("me", C.c_int), ("crf", C.c_double)]


class Ks265GopStats(C.Structure):
    _fields_ = [("frames", C.c_int), ("sse", C.c_uint64 * 3), ("bytes", C.c_uint64), ("gpu_launches", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("h2d_bytes", C.c_uint64)]


assert C.sizeof(KsCell) == 8 and C.sizeof(KsCtuSyn) == 72

GPU_SYMBOLS = ["ks_gpu_open", "ks_gpu_close", "ks_gpu_coded_size", "ks_gpu_upload_frame", "ks_gpu_upload_frame_device", "ks_gpu_stage_acquire", "ks_gpu_upload_staged",
               "ks_gpu_encode_picture_submit", "ks_gpu_encode_picture_finish", "ks_gpu_encode_picture", "ks_gpu_fetch_recon",
               "ks_gpu_launch_count", "ks_gpu_stream", "ks_gpu_set_profiling", "ks_gpu_get_stage_times", "ks_gpu_d2h_bytes", "ks_gpu_abi_sizeof", "ks_gpu_abort", "ks_gpu_debug_fetch", "ks_gpu_debug_me", "ks_gpu_kat_sad16", "ks_gpu_kat_satd16",
               "ks_gpu_kat_interp_luma16", "ks_gpu_kat_tb"]
ENC_SYMBOLS = ["ks265_config_default_preset", "ks265_preset_index", "ks265_encoder_open", "ks265_encoder_close",
               "ks265_encoder_encode_gop", "ks265_encoder_run_gop_device", "ks265_encoder_set_picture_stats", "ks265_encoder_encode_gop_cb", "ks265_alloc_host", "ks265_free_host", "ks265_encoder_set_profiling", "ks265_encoder_get_stage_times", "ks265_encoder_headers"]

_lib = None


def build(verbose=False):
    """compile libks265gpu.so + bin/appencoder in-tree (nvcc, sm_100a)"""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")], stdout=out)


def lib():
    """load the native library; raises (never falls back) if it is missing"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("ks265codec_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the hot path)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.ks_gpu_open.restype = C.c_void_p
    L.ks_gpu_open.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(KsGpuCfg), C.POINTER(C.c_int)]
    L.ks_gpu_close.argtypes = [C.c_void_p]
    L.ks_gpu_coded_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ks_gpu_upload_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.ks_gpu_upload_frame_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ks_gpu_encode_picture_submit.argtypes = [C.c_void_p, C.POINTER(KsPicParams)]
    L.ks_gpu_encode_picture_finish.argtypes = [C.c_void_p, C.c_int, C.POINTER(KsPicOut)]
    L.ks_gpu_encode_picture.argtypes = [C.c_void_p, C.POINTER(KsPicParams), C.POINTER(KsPicOut)]
    L.ks_gpu_fetch_recon.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.ks_gpu_launch_count.restype = C.c_uint64
    L.ks_gpu_launch_count.argtypes = [C.c_void_p]
    L.ks_gpu_abort.argtypes = [C.c_void_p]
    L.ks_gpu_stream.restype = C.c_void_p
    L.ks_gpu_stream.argtypes = [C.c_void_p]
    L.ks_gpu_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.ks_gpu_debug_me.argtypes = [C.c_void_p, C.POINTER(KsPicParams), C.c_void_p]
    L.ks_gpu_kat_sad16.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p]
    L.ks_gpu_kat_satd16.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p]
    L.ks_gpu_kat_interp_luma16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.ks_gpu_kat_tb.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ks265_config_default_preset.argtypes = [C.POINTER(Ks265Config), C.c_char_p]
    L.ks265_preset_index.argtypes = [C.c_char_p]
    L.ks265_encoder_open.restype = C.c_void_p
    L.ks265_encoder_open.argtypes = [C.POINTER(Ks265Config), C.POINTER(C.c_int)]
    L.ks265_encoder_close.argtypes = [C.c_void_p]
    L.ks265_encoder_encode_gop.restype = C.c_long
    L.ks265_encoder_encode_gop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(Ks265GopStats)]
    L.ks265_encoder_run_gop_device.restype = C.c_long
    L.ks265_encoder_run_gop_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Ks265GopStats)]
    L.ks265_encoder_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.ks265_encoder_get_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    _lib = L
    return L


from .encoder import Encoder, default_config  # noqa: E402,F401
