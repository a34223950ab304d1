/*
 * ks_oracle.h -- CPU restatement of the KSC265 (ks265codec v2.6.1.3) HEVC hot-path kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ks265codec_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the checker.
 *
 * The reference ships binaries only.  Every function here cites the symbol in
 * /root/reference/centos_x64/appencoder ("E@0x...") whose behaviour it restates; the restatement
 * is pinned bit-exactly against known-answer vectors harvested from that symbol by
 * oracle/kat/harvest.c (LD_PRELOAD shim) and stored under tests/golden/.
 */
#ifndef KS_ORACLE_H
#define KS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- constant tables (ora_tables.c) ---- */
extern const int8_t  ora_dct32[32][32];      /* == g_uiTr32  E@0x4d0740 */
extern const int8_t  ora_dst4[4][4];         /* DST-VII 4x4 */
extern const int8_t  ora_luma_filter[4][8];  /* == g_iLumaFilterCoeff E@0x4cc780 */
extern const int8_t  ora_chroma_filter[8][4];/* == g_iChromaFilterCoeff E@0x4cc7c0 */
extern const uint8_t ora_tc_table[54];       /* == uiTCTable E@0x4cc660 */
extern const uint8_t ora_beta_table[52];     /* == uiBetaTable E@0x4cc6a0 */
extern const uint8_t ora_chroma_qp[58];      /* == g_ucChromaScale E@0x4cfb40 */
extern const int     ora_quant_scales[6];    /* == g_quantScales E@0x4cfb14 */
extern const int     ora_inv_quant_scales[6];/* == g_invQuantScales E@0x4cfb20 */

/* ---- a1/a2/a15: block cost kernels ---- */
/* sad_c E@0x473db0 (a,b,strideA,strideB,h,w) */
uint32_t ora_sad(const uint8_t *a, const uint8_t *b, long sa, long sb, long h, long w);
/* sad4_c E@0x473e30: SADs at (0,-1),(0,+1),(-1,0),(+1,0) around ref, each << 4 */
void ora_sad4(const uint8_t *src, const uint8_t *ref, long ss, long sr, long h, uint32_t out[4], long w);
/* sad3_c E@0x474070: 3 explicit refs, unshifted */
void ora_sad3(const uint8_t *src, const uint8_t *r0, const uint8_t *r1, const uint8_t *r2,
              long ss, long sr, long h, uint32_t out[3], long w);
/* had_c E@0x474500: SATD, 8x8 Hadamard blocks (sum+2)>>2, 4x4 blocks (sum+1)>>1 */
uint32_t ora_satd(const uint8_t *a, const uint8_t *b, long sa, long sb, long h, long w);
/* sse_c<N> E@0x474d70.. */
uint32_t ora_sse(const uint8_t *a, const uint8_t *b, int sa, int sb, int n);

/* ---- a7: interpolation.  strides in elements. src points at sample (0,0). ---- */
void ora_interp_luma_h_8to8 (uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac); /* E@0x417600 */
void ora_interp_luma_h_8to16(int16_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac); /* E@0x417cd0 */
void ora_interp_luma_v_8to8 (uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac); /* E@0x418250 */
void ora_interp_luma_v_8to16(int16_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac); /* E@0x418b70 */
void ora_interp_luma_v_16to8(uint8_t *dst, int ds, const int16_t *src, int ss, int w, int h, int frac); /* E@0x419360 */
void ora_interp_luma_v_16to16(int16_t *dst, int ds, const int16_t *src, int ss, int w, int h, int frac);/* E@0x419ca0 */
void ora_interp_chroma_h_8to8 (uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac); /* E@0x41a4a0 */
void ora_interp_chroma_h_8to16(int16_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac); /* E@0x41a610 */
void ora_interp_chroma_v_8to8 (uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac); /* E@0x41a740 */
void ora_interp_chroma_v_8to16(int16_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac); /* E@0x41a8f0 */
void ora_interp_chroma_v_16to8(uint8_t *dst, int ds, const int16_t *src, int ss, int w, int h, int frac); /* E@0x41aa70 */
void ora_interp_chroma_v_16to16(int16_t *dst, int ds, const int16_t *src, int ss, int w, int h, int frac);/* E@0x41ac40 */
/* full quarter-pel motion-compensated prediction of a w x h block (spec 8.5.3.3.3) built from the
 * primitives above exactly as the reference composes them (H 8to16 on rows -3..h+4, then V 16to8). */
void ora_mc_luma(uint8_t *dst, int ds, const uint8_t *ref, int rs, int w, int h, int mvx, int mvy);
void ora_mc_chroma(uint8_t *dst, int ds, const uint8_t *ref, int rs, int w, int h, int mvx, int mvy);
/* 14-bit intermediate versions (for bi-pred) + DefaultWeightedBi_c E@0x4350f0 */
void ora_mc_luma_16(int16_t *dst, int ds, const uint8_t *ref, int rs, int w, int h, int mvx, int mvy);
void ora_mc_chroma_16(int16_t *dst, int ds, const uint8_t *ref, int rs, int w, int h, int mvx, int mvy);
void ora_weighted_bi(uint8_t *dst, int ds, const int16_t *p0, const int16_t *p1, int ps, int w, int h);

/* ---- a8..a13: residual path ---- */
void ora_residual(int16_t *res, const uint8_t *src, const uint8_t *pred, int ss, int ps, int n);
/* H265_2dDct{4,8,16,32}_c E@0x4b7600.. / H265_2dDst4x4_c E@0x4b7660.
 * stage-1 shift 2*log2N-2, stage-2 shift 7 (NOT the HM shifts). strides in int16 units. */
void ora_fdct(const int16_t *src, int16_t *dst, int src_stride, int dst_stride, int log2n, int is_dst);
/* H265QuantBlock_c E@0x4a2580 (coef,dst,stride,scale,add,qbits,N,deltaU) -> #nonzero */
int  ora_quant_block(const int16_t *coef, int16_t *dst, int stride, int scale, int add, int qbits, int n, int16_t *delta_u);
/* call-site parameter derivation inside `reconstruct` E@0x47da2f: qbits=21+qp/6-log2N, add=(I?171:85)<<(qbits-9) */
int  ora_quant(const int16_t *coef, int16_t *dst, int stride, int qp, int log2n, int is_intra_slice, int16_t *delta_u);
/* sign-data hiding (HM-style "HDQ" variant, signBitHidingHDQ E@0x4a29c0): adjusts levels so that the
 * parity of each 4x4 coefficient group's abs-sum carries the sign of its first coefficient.
 * scan = coefficient scan (positions in raster order of the NxN block, n*n entries). returns nnz. */
int  ora_sign_hide(const int16_t *coef, int16_t *level, const int16_t *delta_u, int stride, int log2n, const uint16_t *scan);
/* H265DeQuantBlock_c E@0x439540 (src,dst,stride,scale,add,shift,w,lastRow) */
void ora_dequant_block(const int16_t *src, int16_t *dst, int stride, int scale, int add, int shift, int w, int last_row);
void ora_dequant(const int16_t *level, int16_t *coef, int stride, int qp, int log2n);
/* H265_2dIDct*_c E@0x4417f0.. : normative inverse transform (shifts 7, 12) fused with pred add + clip */
void ora_idct_add(const int16_t *coef, uint8_t *dst, const uint8_t *pred, int coef_stride, int dst_stride,
                  int pred_stride, int log2n, int is_dst);

/* ---- intra prediction (spec 8.4.4.2). nb = 4n+1 reference samples after
 * substitution: nb[0]=p[-1][2n-1] ... nb[2n-1]=p[-1][0], nb[2n]=p[-1][-1], nb[2n+1+x]=p[x][-1] ---- */
void ora_intra_pred(uint8_t *dst, int ds, const uint8_t *nb, int log2n, int mode, int is_luma, int strong);

/* ---- a16: deblocking (spec 8.7.2). one 4-sample edge segment.  pix points at q0 of line 0;
 * xstride = step across the edge, ystride = step along it. returns 0 none / 1 weak / 2 strong ---- */
int  ora_deblock_luma_seg(uint8_t *pix, int xstride, int ystride, int beta, int tc);
void ora_deblock_chroma_seg(uint8_t *pix, int xstride, int ystride, int tc, int nlines);

/* ---- a17/a19: SAO ---- */
/* statSaoBoEo01_c E@0x4a6370: packed accumulators v=(d<<12)|1; bo[32], eo[8x8 joint table] */
void ora_sao_stat_boeo01(int *eo, int *bo, const uint8_t *org, const uint8_t *rec, int rec_stride,
                         int org_stride, int w, int h, int row_step);
/* normative statistics for all four EO classes + BO over a CTB region with picture-edge handling:
 * stats[0..3][cat 0..4] = {sum diff, count} for EO class c, stats[4][band 0..31] for BO */
typedef struct { int32_t eo_sum[4][5], eo_cnt[4][5], bo_sum[32], bo_cnt[32]; } ora_sao_stats;
void ora_sao_stats_ctb(ora_sao_stats *st, const uint8_t *org, int os, const uint8_t *rec, int rs,
                       int x0, int y0, int w, int h, int pic_w, int pic_h, int row_step, int n_classes);
/* normative SAO apply of one CTB (spec 8.7.3): src = deblocked picture, dst = output picture */
void ora_sao_apply_ctb(uint8_t *dst, int ds, const uint8_t *src, int ss, int x0, int y0, int w, int h,
                       int pic_w, int pic_h, int type, int band_pos_or_class, const int8_t off[4]);

#ifdef __cplusplus
}
#endif
#endif
