/*
 * ks265_qyabi.h -- the reference's public encoder ABI, served by the B200 encoder (libks265qy.so).
 *
 * A program written against the reference's Android_demo/prebuilt/include/qy265enc.h + qy265def.h (the only public API source,
 * SURVEY.md 8b "Secondary: C API") links against libks265qy.so unchanged: same exported names, same argument meaning, same struct
 * layouts (x86-64 SysV; tests/test_qyabi.py compiles a caller against the reference's own header and checks sizeof/offsetof of
 * every mirrored struct against this file).  Callers that have the reference header keep including it; this header is for callers
 * that do not, and for the shim itself.  Field names here are ours; order, types and meaning are the ABI.
 *
 * What differs from the reference, all of it announced through the log callback:
 *   - only rate control 0 (fixed QP) and 3 (CRF) exist on the device path: any other `rc` makes Open fail with QY_NOTSUPPORTED;
 *   - the unit of device work is a closed GOP shard of `intra_period` pictures, so output lags input by up to one intra period
 *     (the reference lags by its lookahead + B reorder depth); QY265EncoderDelayedFrames counts both buffered inputs and
 *     finished-but-unfetched pictures, so the usual "while (DelayedFrames) EncodeFrame(NULL)" flush loop works as is;
 *   - `intra_period` <= 0 ("only the first picture is intra") becomes shards of 256 pictures, each starting with an IDR;
 *   - vpp_*, two-pass, VUI, tune, latency, AQ and the thread knobs are accepted and ignored (out of scope).
 */
#ifndef KS265_QYABI_H
#define KS265_QYABI_H
#ifdef __cplusplus
extern "C" {
#endif

/* result codes, qy265def.h:7-22 */
#define KSQY_OK            0
#define KSQY_FAIL          ((int)0x80000001)
#define KSQY_OUTOFMEMORY   ((int)0x80000002)
#define KSQY_POINTER       ((int)0x80000003)
#define KSQY_NOTSUPPORTED  ((int)0x80000004)
#define KSQY_BAD_NAME      (-1)          /* QY265_PARAM_BAD_NAME,  qy265enc.h:231 */
#define KSQY_BAD_VALUE     (-2)          /* QY265_PARAM_BAD_VALUE, qy265enc.h:232 */

/* QY265EncConfig, qy265enc.h:51-148 (enums are ints) */
typedef struct ksqy_config {
    void  *auth;
    int    tune, preset, latency, profile_id, headers_before_keyframe, width, height;
    double fps;
    int    bframes, temporal_layer;
    int    vpp_denoise, vpp_edge, vpp_color, vpp_hdr;
    double vpp_hdr_strength;
    int    vpp_hdr_iter;
    double vpp_hdr_sigma_s, vpp_hdr_sigma_r, vpp_recur_filter;
    int    rc, bitrate_kbps, vbv_buffer_size, vbv_max_rate, vbv_min_rate, qp, crf, visual_quality, intra_period, qp_min, qp_max, frame_skip;
    int    wavefront, frame_parallel, threads;
    int    vui_present;
    struct { int signal_type_present, video_format, full_range, colour_desc_present, primaries, transfer, matrix; } vui;
    int    log_level, lookahead, calc_psnr, calc_ssim, short_loading, pass;
    char   stat_file[256];
    double rate_tolerance;
    int    rdoq, me, part, do64, tu_inter, tu_intra, smooth, transskip, subme, satd_inter, satd_intra, search_range, ref_num, ref0, sao,
           long_term_ref, aq_mode;
    double aq_strength;
    int    rasl;
} ksqy_config;

typedef struct ksqy_yuv     { int width, height; unsigned char *plane[3]; int stride[3]; } ksqy_yuv;                 /* QY265YUV,     :160-165 */
typedef struct ksqy_picture { int slice_type, poc; long long pts, dts; ksqy_yuv *yuv; } ksqy_picture;                /* QY265Picture, :168-174 */
typedef struct ksqy_nal     { int nal_type, tid, size; long long pts; unsigned char *payload; } ksqy_nal;            /* QY265Nal,     :177-184 */

#ifndef KS265_QYABI_TYPES_ONLY   /* (the layout test includes this file next to the reference's header, whose prototypes use its own type names) */
/* qy265enc.h:196-233; arguments as there.  EncodeFrame returns the bytes of the access unit it hands out (0 = none yet) or a negative code. */
void *QY265EncoderOpen(ksqy_config *cfg, int *error_code);
void  QY265EncoderClose(void *enc);
void  QY265EncoderReconfig(void *enc, ksqy_config *cfg);                 /* takes effect at the next GOP shard; the picture size cannot change */
int   QY265EncoderEncodeHeaders(void *enc, ksqy_nal **nals, int *nal_count);
int   QY265EncoderEncodeFrame(void *enc, ksqy_nal **nals, int *nal_count, ksqy_picture *in, ksqy_picture *out, int force_logo);
void  QY265EncoderKeyFrameRequest(void *enc);
int   QY265EncoderDelayedFrames(void *enc);
int   QY265ConfigDefault(ksqy_config *cfg, int preset, int tune, int latency);
int   QY265ConfigDefaultPreset(ksqy_config *cfg, char *preset, char *tune, char *latency);
int   QY265ConfigParse(ksqy_config *cfg, const char *name, const char *value);
/* qy265def.h:178-196 */
void  QY265SetLogPrintf(void (*fn)(const char *msg));
void  QY265SetAuthWarning(void (*fn)(void));                             /* there is no licence check here: never called */
extern const char strLibQy265Version[];
#endif

#ifdef __cplusplus
}
#endif
#endif
