/* ora_parse.h -- records produced by the HEVC stream parser (ora_parse.c), shared with the replay (ora_replay.c).  TEST INFRASTRUCTURE. */
#ifndef ORA_PARSE_H
#define ORA_PARSE_H
#include <stddef.h>
#include <stdint.h>

typedef struct ora_sao_rec { uint8_t type[3], pos[3]; int8_t off[3][4]; } ora_sao_rec;   /* per CTU and component: 0 off / 1 band (pos = first band) / 2 edge (pos = class) */
typedef struct ora_cu_rec {
    uint16_t x, y; uint8_t log2, pred_mode /* 0 inter, 1 intra */, part_mode, skip;
    uint8_t merge[4], merge_idx[4], inter_dir[4], ref_idx[4][2], mvp[4][2];
    int16_t mvd[4][2][2];
    uint8_t intra_mode[4], chroma_mode;
    uint8_t root_cbf;
    uint32_t first_tu, n_tu;
} ora_cu_rec;
typedef struct ora_tu_rec { uint16_t x, y; uint8_t log2, cbf; /* bit0 Y, 1 Cb, 2 Cr */ int8_t qp_delta; uint32_t lev_off[3]; } ora_tu_rec;   /* lev_off: index into levels (raster NxN), ~0u = none */

typedef struct ora_pic_stats {
    int poc, slice_type, qp, nal_type, num_ref[2];
    long bits_total, bits_sao, bits_split, bits_cu_hdr, bits_mvd, bits_luma, bits_chroma, bits_intra_mode;
    long n_cu[4] /* by log2 3..6 */, n_skip[4], n_merge[4], n_amvp[4], n_intra[4], n_intra_nxn, n_tu[4] /* by log2 2..5 */, n_cbf_luma, n_cbf_chroma;
    long nz_luma, nz_chroma, sum_abs_luma, sum_abs_chroma, n_mvd_nonzero;
    long sao_on_luma, sao_on_chroma, sao_merge;
} ora_pic_stats;

typedef struct ora_parsed_pic {
    ora_pic_stats st;
    ora_cu_rec *cus; size_t n_cus;
    ora_tu_rec *tus; size_t n_tus;
    int16_t *lev; size_t n_lev;
    int ok;               /* slice ended exactly on end_of_slice_segment_flag after the last CTU */
    int ref_poc[2][16];
    ora_sao_rec *sao;     /* per CTU (raster) */
    int dbk_disabled, beta_off_div2, tc_off_div2, cb_qp_off, cr_qp_off, cu_qp_delta_enabled, any_qp_delta;
    int tmvp, col_ref_idx, max_merge, par_mrg_level;      /* slice_temporal_mvp_enabled_flag, collocated_ref_idx, MaxNumMergeCand, Log2ParMrgLevel */
    int n_list0, list0_poc[16];                           /* RefPicList0 as POCs */
    int n_list1, list1_poc[16], col_from_l0, mvd_l1_zero;
    int qg_depth;                                         /* diff_cu_qp_delta_depth */
    int sign_hiding;                                      /* pps sign_data_hiding_enabled_flag */
} ora_parsed_pic;

typedef struct ora_parsed_stream {
    int width, height, n_pics, log2_ctb, log2_min_cb, max_merge;
    ora_parsed_pic *pics;
    int error;            /* 0 = every slice parsed to its end */
    int strong_intra;     /* sps strong_intra_smoothing_enabled_flag */
} ora_parsed_stream;

#endif
