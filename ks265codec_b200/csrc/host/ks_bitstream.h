/*
 * ks_bitstream.h -- host-side HEVC bitstream writer (parameter sets, slice headers, CABAC slice data).
 * North star: "CABAC and rate-control stay on the host".  Counterpart of the reference's
 * EncParameterSetWrite.cpp (write_ParamSet<VPS/SPS/PPS>), EncSlice.cpp (write_slice_segment_header E@0x4aa090),
 * EncCtuSbac.cpp (CCtuSbac::processCtuSbac E@0x46dd00, encodeCoeffNxN E@0x46df40) and EncCabac.cpp
 * (CEncCabacEngine::EncodeBin/Bypass/Flush).  Written from the H.265 spec; correctness is arbitrated by the
 * reference's own decoder (centos_x64/appdecoder) in tests/ (SURVEY.md 8c tier P1).
 */
#ifndef KS_BITSTREAM_H
#define KS_BITSTREAM_H
#include <stddef.h>
#include <stdint.h>
#include "ks265_syntax.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ks_stream_params {
    int disp_width, disp_height;     /* -wdt / -hgt */
    int width, height;               /* coded size (multiple of 16) */
    int fps_num, fps_den;
    int sign_hiding;                 /* PPS sign_data_hiding_enabled_flag */
    int sao;                         /* SPS sample_adaptive_offset_enabled_flag */
    int max_merge_cand;              /* slice MaxNumMergeCand (reference: 3) */
    int pps_beta_offset_div2, pps_tc_offset_div2;
    int strong_intra_smoothing;
    int log2_max_poc_lsb;            /* 8 */
    int bframes;                     /* > 0: streams carry B pictures (DPB 3, one picture of reordering) */
} ks_stream_params;

typedef struct ks_slice_params {
    int nal_type;                    /* 19 IDR_W_RADL, 1 TRAIL_R, 0 TRAIL_N ... */
    int temporal_id;
    int slice_type;
    int poc;
    int qp;
    int num_neg_refs;                /* short-term RPS: negative pictures (delta POCs, all used by curr) */
    int neg_delta_poc[4];            /* e.g. {-1}; [0] is RefPicList0[0] */
    int num_pos_refs;                /* short-term RPS: positive pictures (B pictures) */
    int pos_delta_poc[4];            /* e.g. {+3}; [0] is RefPicList1[0] */
    int deblock_override;            /* slice-level override of beta/tc */
    int beta_offset_div2, tc_offset_div2;
    int sao_luma, sao_chroma;
} ks_slice_params;

/* each returns the number of bytes written (Annex-B: 00 00 00 01 + NAL), or -1 if `cap` is too small */
long ks_write_vps(const ks_stream_params *sp, uint8_t *out, size_t cap);
long ks_write_sps(const ks_stream_params *sp, uint8_t *out, size_t cap);
long ks_write_pps(const ks_stream_params *sp, uint8_t *out, size_t cap);
/* entropy-code one picture (one slice).  scratch: caller-provided work memory of ks_slice_scratch_bytes(). */
size_t ks_slice_scratch_bytes(const ks_stream_params *sp);
long ks_write_slice(const ks_stream_params *sp, const ks_slice_params *sl, const ks_frame_syn *syn,
                    void *scratch, uint8_t *out, size_t cap);

/* MD5 of a byte range (for -md5 1, reference calcMd5 / libmd5.cpp) */
void ks_md5(const uint8_t *data, size_t len, uint8_t digest[16]);
/* PSNR helper: sum of squared error between two planes */
uint64_t ks_plane_sse(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h);

#ifdef __cplusplus
}
#endif
#endif
