/*
 * ora_frame.h -- CPU model of the per-picture hot path (TEST INFRASTRUCTURE; see ks_oracle.h).
 *
 * The leaf kernels in ora_kernels.c restate the reference's primitives (rows a1..a19 of SURVEY.md 8a).
 * The reference's decision layer (processTree E@0x46b610 etc.) is closed, so the picture-level
 * driver here restates OUR decision algorithm (documented in DESIGN.md), composed from those primitives in
 * the order the reference's drivers use them: meInitPoint/interMeDia/subMeSquare (a3,a5,a6) -> interpolatePu*
 * (a7) -> reconstruct (a14: a8..a13) -> ctuDeblockFilterVer/Hor (a16) -> SAO stats/decide/apply (a17..a19,
 * sequencing a20).  The CUDA path must reproduce its outputs bit for bit.
 */
#ifndef ORA_FRAME_H
#define ORA_FRAME_H
#include <stdint.h>
#include "../include/ks265_syntax.h"
#ifdef __cplusplus
extern "C" {
#endif

#define ORA_PAD 96

typedef struct ora_cfg {
    int width, height;      /* coded size, multiples of 16 */
    int me_range;           /* +-integer search range */
    int me_iters;           /* max small-diamond steps */
    int subpel;             /* 0 integer, 1 half, 2 quarter */
    int sign_hiding;
    int sao;                /* 0 off; 1..3: BO + EO classes 0,1, statistics from every 2nd row (reference `_fast` statSao variants,
                               presets ultrafast..fast); 4: BO + all four EO classes, every row */
    int strong_intra;
    int satd;               /* sub-pel cost: SATD (had_c) instead of SAD */
    int me_method;          /* integer search: 0 small diamond (a3 interMeDia), 1 hexagon + square refine (a4 interMeHex) */
} ora_cfg;

typedef struct ora_plane { uint8_t *base, *p; int stride, w, h; } ora_plane;
typedef struct ora_pic { ora_plane c[3]; } ora_pic;

int  ora_pic_alloc(ora_pic *pic, int w, int h);
void ora_pic_free(ora_pic *pic);
void ora_pic_load(ora_pic *pic, const uint8_t *i420, int src_w, int src_h);   /* copies + edge-extends to coded size */
void ora_pic_extend(ora_pic *pic);                                              /* replicate borders into the pad */

/* dense level planes, same geometry as the pixel planes (no pad) */
typedef struct ora_levels { int16_t *c[3]; } ora_levels;

/* picture-level stages.  `cells` has (w/16)*(h/16) entries. */
void ora_intra_picture(const ora_cfg *cfg, int qp, const ora_pic *src, ora_pic *rec, ks_cell *cells, ora_levels *lv);
/* returns the sum of the per-cell winning search costs (the rate control's complexity measure) */
/* lambda_qp: the QP whose lambda drives the CU/merge decision and the RD zero-out of residual blocks (ks_pic_params.lambda_qp) */
uint64_t ora_inter_picture(const ora_cfg *cfg, int qp, int lambda_qp, const ora_pic *src, const ora_pic *ref, const ks_cell *prev_cells,
                       ora_pic *rec, ks_cell *cells, ora_levels *lv);
/* the motion search alone: per-cell vectors + the distortion of each winner */
uint64_t ora_me_field(const ora_cfg *cfg, int qp, const ora_pic *src, const ora_pic *ref, const ks_cell *prev_cells, ks_cell *cells, int *dist);
void ora_b_picture(const ora_cfg *cfg, int qp, int lambda_qp, const ora_pic *src, const ora_pic *ref0, const ora_pic *ref1, const ks_cell *anchor_cells,
                   int d0, int da, ora_pic *rec, ks_cell *cells, ks_cell_b *cells_b, ora_levels *lv);
void ora_deblock_picture(const ora_cfg *cfg, int qp, int beta_offset_div2, int tc_offset_div2, ora_pic *rec, const ks_cell *cells);
void ora_deblock_picture_b(const ora_cfg *cfg, int qp, int beta_offset_div2, int tc_offset_div2, ora_pic *rec, const ks_cell *cells, const ks_cell_b *cells_b);
void ora_sao_picture(const ora_cfg *cfg, int qp, const ora_pic *src, const ora_pic *deblocked, ora_pic *out, ks_ctu_syn *ctus);
/* pack dense levels into the boundary format (CG bitmaps + pool); returns number of CGs */
uint32_t ora_pack_levels(const ora_cfg *cfg, const ora_levels *lv, ks_ctu_syn *ctus, int16_t *pool);

/* lambda tables shared (by value) with the product: round(16*sqrt(0.85*2^((qp-12)/3))), round(16*0.85*2^((qp-12)/3)) */
extern const int ora_lambda_sad_q4[52];
extern const int ora_lambda_sse_q4[52];

#ifdef __cplusplus
}
#endif
#endif
