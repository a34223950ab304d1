/*
 * ks265_enc.h -- host-side encoder API of the ks265 B200 encoder (libks265gpu.so).
 *
 * Mirrors the reference's public C API in Android_demo/prebuilt/include/qy265enc.h: QY265ConfigDefaultPreset (:226)
 * -> ks265_config_default_preset, QY265EncoderOpen (:196) -> ks265_encoder_open, QY265EncoderEncodeFrame (:215) ->
 * ks265_encoder_encode_gop (the unit of work here is a closed GOP shard: the device pipeline is picture-serial
 * inside a GOP and GOP shards are what spread over streams / GPUs), QY265EncoderClose (:198) -> ks265_encoder_close.
 * Error codes follow qy265def.h:7-22 in spirit: 0 = OK, negative = failure.
 */
#ifndef KS265_ENC_H
#define KS265_ENC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ks265_config {
    int width, height;          /* -wdt / -hgt */
    double fps;                 /* -fr */
    int preset;                 /* 0 ultrafast .. 8 placebo (README.md:20-24) */
    int rc;                     /* -rc: only 0 (fixed QP) runs on the device path */
    int qp;                     /* -qp */
    int iper;                   /* -iper: intra period = GOP shard length */
    int fixqp;                  /* -fixqp: 1 = same QP for I and P (default: P = QP+1 like the reference) */
    int sao;                    /* -sao level 0..4 (see ks_gpu_cfg.sao) */
    int sign_hiding;
    int me_range, me_iters, subpel;
    int satd;                   /* sub-pel cost metric: 0 SAD (ultrafast..veryfast), 1 SATD (fast..placebo), like the reference */
    int device;                 /* CUDA device ordinal */
    int psnr;                   /* compute per-plane SSE on the device */
    int bframes;                /* -bframes: B pictures between anchors (0 = IDR + P...; default 0 this round) */
    int me;                     /* -me: integer search, 0 small diamond (DIA), 1 hexagon (HEX); the reference's 2 (UMH) maps to 1 */
    double crf;                 /* -crf (used with -rc 3) */
} ks265_config;

typedef struct ks265_gop_stats {
    int frames;
    uint64_t sse[3];            /* summed over the shard (display area, like the reference's PSNR) */
    uint64_t bytes;
    uint64_t gpu_launches;
    uint64_t d2h_bytes;         /* syntax bytes copied device->host */
    uint64_t h2d_bytes;         /* picture bytes copied host->device */
} ks265_gop_stats;

/* per-picture record (the rows of the reference's `-psnr 2` table), coding order */
typedef struct ks265_pic_stat {
    int poc, slice_type, qp;    /* poc = display index inside the shard; slice_type KS_SLICE_* (0 B, 1 P, 2 I) */
    uint64_t bits;
    uint64_t sse[3];
} ks265_pic_stat;

typedef struct ks265_encoder ks265_encoder;

int  ks265_config_default_preset(ks265_config *cfg, const char *preset);     /* fills everything but width/height */
int  ks265_preset_index(const char *name);
ks265_encoder *ks265_encoder_open(const ks265_config *cfg, int *err);
void ks265_encoder_close(ks265_encoder *enc);
/* Encode `nframes` display-size I420 pictures (host memory, tightly packed) as one closed GOP: IDR + P...
 * Writes Annex-B NAL units (VPS/SPS/PPS first) to `bs`; optionally the reconstruction (display size I420) to `recon`.
 * `frames_dev` != NULL means the pictures already live in DEVICE memory (same layout) and `frames` is ignored.
 * Returns bytes written or a negative error. */
long ks265_encoder_encode_gop(ks265_encoder *enc, const uint8_t *frames, const void *frames_dev, int nframes,
                              uint8_t *bs, size_t bs_cap, uint8_t *recon, ks265_gop_stats *stats);
/* VPS + SPS + PPS of the stream (Annex-B), the same bytes every GOP shard starts with (reference: QY265EncoderEncodeHeaders, qy265enc.h:202) */
long ks265_encoder_headers(ks265_encoder *enc, uint8_t *out, size_t cap);
/* the next encode_gop calls also fill `stats[0..cap)` with one record per coded picture (NULL / 0 turns it off) */
void ks265_encoder_set_picture_stats(ks265_encoder *enc, ks265_pic_stat *stats, int cap);
/* page-locked host memory for picture buffers: pictures handed to encode_gop from such a buffer are DMA-ed in place (no staging copy) */
void *ks265_alloc_host(size_t bytes);
void ks265_free_host(void *p);
/* the same GOP shard with the pictures PULLED through a callback: `read_picture(opaque, display_index, dst)` fills dst with one display-size
 * I420 picture (dst is page-locked staging memory of the device context: a file read lands where the DMA engine picks it up) and returns 0.
 * The reference's counterpart is its reader thread (CInputYUV::startReadThread E@0x4cbe80) feeding QY265EncoderEncodeFrame. */
typedef int (*ks265_read_fn)(void *opaque, int display_index, uint8_t *dst);
long ks265_encoder_encode_gop_cb(ks265_encoder *enc, ks265_read_fn read_picture, void *opaque, int nframes,
                                 uint8_t *bs, size_t bs_cap, uint8_t *recon, ks265_gop_stats *stats);
/* per-stage device times accumulated since `on` (see ks_gpu_get_stage_times) */
int  ks265_encoder_set_profiling(ks265_encoder *enc, int on);
int  ks265_encoder_get_stage_times(ks265_encoder *enc, double ms[8], uint64_t launches[8]);   /* KS_NSTAGES entries */
/* device-only variant for measurement: runs the device pipeline of a GOP without entropy coding */
long ks265_encoder_run_gop_device(ks265_encoder *enc, const void *frames_dev, int nframes, ks265_gop_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
