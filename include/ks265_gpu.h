/*
 * ks265_gpu.h -- C-ABI of the B200 hot path (libks265gpu.so).  Plain pointers and sizes only.
 *
 * The reference (ksvc/ks265codec, binary-only) has no external hook for its hot path: its internal boundary is
 * the set of global function-pointer tables filled by initEncGlobeVar (E@0x4738e0) / initDCT_Function (E@0x473830)
 * / initCommonGlobeVar (E@0x433ec0) / initDeblockFunc (E@0x433e40) / initSaoEncFunction (E@0x4a6b60) and driven
 * per CTU by CCtuEnc::processOneCtu (E@0x4692e0) -> motionSearchP (E@0x47d070) -> reconstruct (E@0x47d600) ->
 * CLoopFilterCtu::Execute (E@0x492b40).  SURVEY.md 8b therefore defines this new C-ABI between the host encoder
 * (CLI `appencoder`, qy265enc.h-style API) and the device; each entry point names what it replaces.
 *
 * Threading: a context is single-caller (one host thread per context, like a QY265 encoder handle); use one
 * context per concurrently encoded GOP shard.  All functions return 0 on success, a negative KS_E* code on error.
 */
#ifndef KS265_GPU_H
#define KS265_GPU_H
#include <stddef.h>
#include <stdint.h>
#include "ks265_syntax.h"
#ifdef __cplusplus
extern "C" {
#endif

#define KS_EINVAL  (-22)
#define KS_ENOMEM  (-12)
#define KS_ECUDA   (-5)
#define KS_ENODEV  (-19)

typedef struct ks_gpu_ctx ks_gpu_ctx;

typedef struct ks_gpu_cfg {
    int me_range;       /* -merange (reference constant 64) */
    int me_iters;       /* small-diamond steps (reference: tME.range >> shift, interMeDia E@0x4849d0) */
    int subpel;         /* 0 integer, 1 half, 2 quarter (reference -subme) */
    int sign_hiding;    /* PPS sign_data_hiding_enabled_flag (reference: 1 in every preset) */
    int sao;            /* reference -sao level: 0 off; 1..3 BO + EO 0/1 with `_fast` (every 2nd row) statistics; 4 BO + EO 0..3, all rows */
    int strong_intra;
    int n_src_slots;    /* source pictures resident on the device (>= 2) */
    int n_rec_slots;    /* reconstructed/reference pictures resident on the device (>= 2) */
    int n_syn_slots;    /* pictures in flight between submit and finish (>= 2) */
    int satd;           /* sub-pel cost = SATD (had_c) instead of SAD: reference `satdInter`, presets fast..placebo */
    int me_method;      /* integer search: 0 small diamond (interMeDia E@0x4849d0), 1 hexagon + square refine (interMeHex E@0x484c00) */
} ks_gpu_cfg;

typedef struct ks_pic_params {
    int slice_type;     /* KS_SLICE_I / KS_SLICE_P / KS_SLICE_B */
    int qp;
    int src_slot;       /* source picture (ks_gpu_upload_frame*) */
    int ref_slot;       /* reconstructed picture used as list-0 reference (-1 for I) */
    int out_slot;       /* where this picture's final reconstruction goes */
    int syn_slot;       /* syntax/output slot for this picture */
    int prev_syn_slot;  /* syntax slot of the previous coded picture: its MVs seed the search (-1: none) */
    int beta_offset_div2, tc_offset_div2;
    int want_sse;       /* accumulate per-plane SSE vs source (for -psnr) */
    /* B pictures (reference: motionSearchB E@0x47c710 / interMeBiFull_c E@0x480040 / DefaultWeightedBi_c E@0x4350f0): */
    int ref1_slot;      /* reconstructed picture used as list-1 reference (the LATER anchor); ref_slot is the earlier one */
    int dist_l0;        /* POC(cur) - POC(list-0 reference) > 0 */
    int dist_anchor;    /* POC(list-1 reference) - POC(list-0 reference); prev_syn_slot names the later anchor, whose vectors
                           (spanning dist_anchor pictures) are scaled to seed both searches */
    int want_me_cost;   /* P pictures: return the sum of the per-cell winning search costs (host rate control's complexity measure) */
    int lambda_qp_delta;/* >= 0: the CU/merge decision and the RD zero-out of residual blocks use the lambda of QP = qp + delta (the host raises it on
                           the non-key P pictures of the 4-picture cascade, ks_rc_lambda_qp) */
} ks_pic_params;

/* results of one picture: pointers into pinned host memory owned by the context, valid until the syntax slot is reused */
typedef struct ks_pic_out {
    const ks_cell    *cells;
    const ks_ctu_syn *ctus;
    const int16_t    *levels;
    uint32_t          n_cg;
    uint64_t          sse[3];
    const ks_cell_b  *cells_b;     /* B pictures, else NULL */
    uint64_t          me_cost;     /* want_me_cost: sum over cells of (SAD or SATD + lambda*mv bits) of the chosen vector, else 0 */
} ks_pic_out;

/* replaces: createHevcEncoder/createModules (E@0x4b44f0/0x4b3380) device-side state; width/height = display size */
ks_gpu_ctx *ks_gpu_open(int device, int width, int height, const ks_gpu_cfg *cfg, int *err);
void ks_gpu_close(ks_gpu_ctx *ctx);
int  ks_gpu_coded_size(const ks_gpu_ctx *ctx, int *width, int *height);
/* replaces: ctuCacheLoadSrcYuv (EncCtuCache.cpp) -- host planes -> device source slot (pinned staging + async H2D) */
int  ks_gpu_upload_frame(ks_gpu_ctx *ctx, int slot, const uint8_t *y, const uint8_t *u, const uint8_t *v, int stride_y, int stride_uv);
/* same without a caller-side picture buffer: acquire the context's next page-locked staging buffer (display-size I420, tightly packed), fill it
 * (e.g. read() a file straight into it), then start its H2D into `slot`.  Two buffers alternate; acquire waits for the buffer's previous copy. */
uint8_t *ks_gpu_stage_acquire(ks_gpu_ctx *ctx);
int  ks_gpu_upload_staged(ks_gpu_ctx *ctx, int slot);
/* same, but the I420 picture (display size, tightly packed) already lives in device memory.  When the display size is already a multiple of 16
 * (and the pointer 16-byte aligned) the kernels read it IN PLACE -- no copy; the caller keeps it unchanged until the picture has finished */
int  ks_gpu_upload_frame_device(ks_gpu_ctx *ctx, int slot, const void *dev_i420);
/* replaces: IEncTaskManage::executeTasks -> processOneCtu for a whole picture: ME + sub-pel (a1-a7), CU quadtree / merge decision
 * (processTree E@0x46b610), MC + residual DCT/quant/SBH/dequant/IDCT with RD zero-out (a8-a14), deblock (a16), SAO (a17-a20), level packing;
 * starts the D2H of the syntax */
int  ks_gpu_encode_picture_submit(ks_gpu_ctx *ctx, const ks_pic_params *pp);
/* waits for the picture submitted on `syn_slot` and returns its syntax */
int  ks_gpu_encode_picture_finish(ks_gpu_ctx *ctx, int syn_slot, ks_pic_out *out);
/* submit + finish */
int  ks_gpu_encode_picture(ks_gpu_ctx *ctx, const ks_pic_params *pp, ks_pic_out *out);
/* replaces: dumpYUVWithCrop (E@0x4b54c0): reconstructed picture, cropped to the display size */
int  ks_gpu_fetch_recon(ks_gpu_ctx *ctx, int rec_slot, uint8_t *y, uint8_t *u, uint8_t *v, int stride_y, int stride_uv);
/* number of kernel launches issued by this context so far (bench.py gpu_launches) */
uint64_t ks_gpu_launch_count(const ks_gpu_ctx *ctx);
/* raw CUDA stream of the context (cudaStream_t) so callers can time with events on the launching stream */
void *ks_gpu_stream(ks_gpu_ctx *ctx);

/* per-stage device timing (CUDA events on the context's stream, accumulated at finish): stage order
 * 0 motion search, 1 inter prediction+residual, 2 intra picture, 3 deblock, 4 SAO, 5 level packing, 6 CU/merge decision, 7 intra CUs of P pictures */
#define KS_NSTAGES 8
int  ks_gpu_set_profiling(ks_gpu_ctx *ctx, int on);
int  ks_gpu_get_stage_times(const ks_gpu_ctx *ctx, double ms[KS_NSTAGES], uint64_t launches[KS_NSTAGES]);
/* drop every picture still in flight (submitted, not finished): waits for the device, clears the pending marks.  For error paths. */
int  ks_gpu_abort(ks_gpu_ctx *ctx);
/* bytes copied device->host so far (syntax blocks) */
uint64_t ks_gpu_d2h_bytes(const ks_gpu_ctx *ctx);
/* sizeof() of the ABI structs as this library was built (binding self-check: 0 ks_gpu_cfg, 1 ks_pic_params, 2 ks_pic_out, 3 ks_cell,
 * 4 ks_cell_b, 5 ks_ctu_syn, 6 ks265_config, 7 ks265_gop_stats; 0 for an unknown index) */
size_t ks_gpu_abi_sizeof(int which);

/* ---- stage-level debug/test access (tests compare every stage with the oracle) ---- */
enum { KS_DBG_PRE_RECON = 0, KS_DBG_LEVELS = 1, KS_DBG_SRC = 2 };
/* copies coded-size planes (Y, U, V back to back; levels as int16) of the LAST submitted picture to `dst` */
int  ks_gpu_debug_fetch(ks_gpu_ctx *ctx, int what, int slot, void *dst, size_t bytes);
/* run only the motion search of a P picture and return the cells (16x16 MV field) */
int  ks_gpu_debug_me(ks_gpu_ctx *ctx, const ks_pic_params *pp, ks_cell *cells_out);

/* ---- known-answer entry points: the reference's leaf signatures replayed on the device ---- */
/* sad_c E@0x473db0 (a,b,strideA,strideB,h,w), w,h in {16} -- the ME kernel's VABSDIFF4 + shuffle path */
int  ks_gpu_kat_sad16(const uint8_t *a, const uint8_t *b, long stride_a, long stride_b, uint32_t *out);
/* had_c E@0x474500 (SATD, 8x8 Hadamard tiles) on a 16x16 block -- register butterflies + shuffle stages */
int  ks_gpu_kat_satd16(const uint8_t *a, const uint8_t *b, long stride_a, long stride_b, uint32_t *out);
/* interpLuma{Hor,Ver}8to8 / Hor8to16+Ver16to8 composition for a 16x16 block at quarter-sample (fx,fy);
 * `ref` points at integer sample (0,0) of a plane of size w x h (coordinates clamp at the borders) */
int  ks_gpu_kat_interp_luma16(const uint8_t *ref_plane, int w, int h, int x, int y, int mvx, int mvy, uint8_t *dst16x16);
/* H265_2dDct{4,8,16,32}_c + H265QuantBlock_c + H265DeQuantBlock_c + H265_2dIDct*_c chain on one block (log2n 2..5; the 4x4 DST is not on the device):
 * src/pred N x N (stride N); outputs levels (N x N int16) and reconstruction (N x N) */
int  ks_gpu_kat_tb(int log2n, const uint8_t *src, const uint8_t *pred, int qp, int intra_slice, int sign_hiding,
                   int16_t *levels, uint8_t *recon, int *cbf);

#ifdef __cplusplus
}
#endif
#endif
