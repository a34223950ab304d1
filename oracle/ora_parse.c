/*
 * ora_parse.c -- HEVC (Main profile) bitstream PARSER for the streams the reference encoder emits: parameter sets, slice headers and the
 * CABAC slice data (H.265 7.3 / 9.3).  TEST INFRASTRUCTURE (SURVEY.md 8c tier P2, 8f row f4): it recovers the reference's own decisions
 * (CU quadtree, skip/merge/AMVP, intra modes, vectors, transform levels) from `appencoder -b` output so that
 *   - tools/stream_stats.py can put the reference's and this repo's decisions side by side (where do the bits go), and
 *   - a replay of those decisions through the device kernels can be compared with the reference's `-o` reconstruction.
 * Written from the spec; validated by construction: every slice must end exactly on end_of_slice_segment_flag = 1 after the last CTU, with
 * the arithmetic decoder in sync, for every picture of every reference stream under test (tests/test_parse.py).
 * Limits: 4:2:0 8-bit, one slice per picture, no tiles / WPP entry points / PCM / transquant bypass / scaling lists / long-term refs /
 * weighted prediction (the reference uses none of them at -threads 1).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/ks265_syntax.h"
#include "ora_parse.h"

/* ------------------------------------------------------------------ bit reader ---------------------- */
typedef struct { const uint8_t *b; size_t n, pos; } bitr;          /* pos in bits */
static uint32_t br_u(bitr *r, int n) { uint32_t v = 0; while (n--) { uint32_t bit = r->pos < r->n * 8 ? (r->b[r->pos >> 3] >> (7 - (r->pos & 7))) & 1 : 0; v = (v << 1) | bit; r->pos++; } return v; }
static uint32_t br_ue(bitr *r) { int z = 0; while (r->pos < r->n * 8 && !br_u(r, 1)) z++; return z ? ((1u << z) - 1 + br_u(r, z)) : 0; }
static int br_se(bitr *r) { uint32_t k = br_ue(r); return (k & 1) ? (int)((k + 1) >> 1) : -(int)(k >> 1); }

/* ------------------------------------------------------------------ parameter sets ------------------ */
typedef struct { int n_neg, n_pos; int dpoc[16]; int used[16]; } rps_t;       /* negatives first (closest first), then positives */
typedef struct {
    int w, h, log2_min_cb, log2_ctb, log2_min_tb, log2_max_tb, tu_depth_inter, tu_depth_intra, amp, sao, log2_max_poc, tmvp, strong_intra;
    int n_rps; rps_t rps[64]; int long_term, n_lt_sps;
} sps_t;
typedef struct {
    int dep_slices, output_flag_present, extra_bits, sign_hiding, cabac_init_present, ref_l0, ref_l1, init_qp, tskip, cu_qp_delta, diff_cu_qp_delta_depth,
        cb_off, cr_off, slice_chroma_off, wp, wbp, tq_bypass, tiles, wpp, lf_across, deblock_ctrl, deblock_override, deblock_disabled, beta, tc, lists_mod,
        par_mrg, slice_ext;
} pps_t;

static void parse_ptl(bitr *r, int max_sub) { br_u(r, 8); br_u(r, 32); br_u(r, 4); br_u(r, 32); br_u(r, 11); br_u(r, 1); br_u(r, 8);
    int pp[8], lp[8]; for (int i = 0; i < max_sub; i++) { pp[i] = br_u(r, 1); lp[i] = br_u(r, 1); }
    if (max_sub > 0) for (int i = max_sub; i < 8; i++) br_u(r, 2);
    for (int i = 0; i < max_sub; i++) { if (pp[i]) { br_u(r, 32); br_u(r, 32); br_u(r, 24); } if (lp[i]) br_u(r, 8); } }

static int parse_rps(bitr *r, sps_t *s, int idx, int n_sets, rps_t *out, int in_slice)
{
    int inter = idx ? br_u(r, 1) : 0;
    memset(out, 0, sizeof(*out));
    if (inter) {
        int delta_idx = in_slice ? (int)br_ue(r) + 1 : 1;
        const rps_t *ref = &s->rps[idx - delta_idx];
        int sign = br_u(r, 1), absd = (int)br_ue(r) + 1, drps = sign ? -absd : absd, nref = ref->n_neg + ref->n_pos;
        int used[17], usedelta[17];
        for (int j = 0; j <= nref; j++) { used[j] = br_u(r, 1); usedelta[j] = 1; if (!used[j]) usedelta[j] = br_u(r, 1); }
        /* 7.4.8 (7-61, 7-62) */
        int k = 0;
        for (int j = ref->n_pos - 1; j >= 0; j--) { int d = ref->dpoc[ref->n_neg + j] + drps; if (d < 0 && usedelta[ref->n_neg + j]) { out->dpoc[k] = d; out->used[k++] = used[ref->n_neg + j]; } }
        if (drps < 0 && usedelta[nref]) { out->dpoc[k] = drps; out->used[k++] = used[nref]; }
        for (int j = 0; j < ref->n_neg; j++) { int d = ref->dpoc[j] + drps; if (d < 0 && usedelta[j]) { out->dpoc[k] = d; out->used[k++] = used[j]; } }
        out->n_neg = k;
        for (int j = ref->n_neg - 1; j >= 0; j--) { int d = ref->dpoc[j] + drps; if (d > 0 && usedelta[j]) { out->dpoc[k] = d; out->used[k++] = used[j]; } }
        if (drps > 0 && usedelta[nref]) { out->dpoc[k] = drps; out->used[k++] = used[nref]; }
        for (int j = 0; j < ref->n_pos; j++) { int d = ref->dpoc[ref->n_neg + j] + drps; if (d > 0 && usedelta[ref->n_neg + j]) { out->dpoc[k] = d; out->used[k++] = used[ref->n_neg + j]; } }
        out->n_pos = k - out->n_neg;
    } else {
        out->n_neg = (int)br_ue(r); out->n_pos = (int)br_ue(r);
        if (out->n_neg + out->n_pos > 16) return -1;
        int prev = 0;
        for (int i = 0; i < out->n_neg; i++) { prev -= (int)br_ue(r) + 1; out->dpoc[i] = prev; out->used[i] = br_u(r, 1); }
        prev = 0;
        for (int i = 0; i < out->n_pos; i++) { prev += (int)br_ue(r) + 1; out->dpoc[out->n_neg + i] = prev; out->used[out->n_neg + i] = br_u(r, 1); }
    }
    (void)n_sets;
    return 0;
}
static int parse_sps(bitr *r, sps_t *s)
{
    memset(s, 0, sizeof(*s));
    br_u(r, 4); int max_sub = br_u(r, 3); br_u(r, 1); parse_ptl(r, max_sub);
    br_ue(r); if (br_ue(r) != 1) return -1;                     /* chroma_format_idc must be 4:2:0 */
    s->w = (int)br_ue(r); s->h = (int)br_ue(r);
    if (br_u(r, 1)) { br_ue(r); br_ue(r); br_ue(r); br_ue(r); }
    if (br_ue(r) || br_ue(r)) return -1;                        /* 8-bit only */
    s->log2_max_poc = (int)br_ue(r) + 4;
    int sub_info = br_u(r, 1);
    for (int i = sub_info ? 0 : max_sub; i <= max_sub; i++) { br_ue(r); br_ue(r); br_ue(r); }
    s->log2_min_cb = (int)br_ue(r) + 3; s->log2_ctb = s->log2_min_cb + (int)br_ue(r);
    s->log2_min_tb = (int)br_ue(r) + 2; s->log2_max_tb = s->log2_min_tb + (int)br_ue(r);
    s->tu_depth_inter = (int)br_ue(r); s->tu_depth_intra = (int)br_ue(r);
    if (br_u(r, 1)) return -2;                                  /* scaling lists */
    s->amp = br_u(r, 1); s->sao = br_u(r, 1);
    if (br_u(r, 1)) return -3;                                  /* pcm */
    s->n_rps = (int)br_ue(r);
    if (s->n_rps > 64) return -1;
    for (int i = 0; i < s->n_rps; i++) if (parse_rps(r, s, i, s->n_rps, &s->rps[i], 0)) return -1;
    s->long_term = br_u(r, 1);
    if (s->long_term) { s->n_lt_sps = (int)br_ue(r); for (int i = 0; i < s->n_lt_sps; i++) { br_u(r, s->log2_max_poc); br_u(r, 1); } }
    s->tmvp = br_u(r, 1); s->strong_intra = br_u(r, 1);
    return 0;
}
static int parse_pps(bitr *r, pps_t *p)
{
    memset(p, 0, sizeof(*p));
    br_ue(r); br_ue(r); p->dep_slices = br_u(r, 1); p->output_flag_present = br_u(r, 1); p->extra_bits = br_u(r, 3); p->sign_hiding = br_u(r, 1); p->cabac_init_present = br_u(r, 1);
    p->ref_l0 = (int)br_ue(r) + 1; p->ref_l1 = (int)br_ue(r) + 1; p->init_qp = br_se(r) + 26; br_u(r, 1); p->tskip = br_u(r, 1);
    p->cu_qp_delta = br_u(r, 1); if (p->cu_qp_delta) p->diff_cu_qp_delta_depth = (int)br_ue(r);
    p->cb_off = br_se(r); p->cr_off = br_se(r); p->slice_chroma_off = br_u(r, 1); p->wp = br_u(r, 1); p->wbp = br_u(r, 1); p->tq_bypass = br_u(r, 1);
    p->tiles = br_u(r, 1); p->wpp = br_u(r, 1);
    if (p->tiles || p->tq_bypass || p->wp || p->wbp) return -1;
    p->lf_across = br_u(r, 1); p->deblock_ctrl = br_u(r, 1);
    if (p->deblock_ctrl) { p->deblock_override = br_u(r, 1); p->deblock_disabled = br_u(r, 1); if (!p->deblock_disabled) { p->beta = br_se(r); p->tc = br_se(r); } }
    if (br_u(r, 1)) return -2;
    p->lists_mod = br_u(r, 1); p->par_mrg = (int)br_ue(r) + 2; p->slice_ext = br_u(r, 1);
    return 0;
}

/* ------------------------------------------------------------------ CABAC decoder (9.3.4.3) --------- */
enum {
    CX_SPLIT_CU = 0, CX_SKIP = 3, CX_MERGE_FLAG = 6, CX_MERGE_IDX = 7, CX_PART_MODE = 8, CX_PRED_MODE = 12,
    CX_PREV_INTRA = 13, CX_CHROMA_PRED = 14, CX_MVD = 15, CX_CBF_LUMA = 17, CX_CBF_CHROMA = 19, CX_ROOT_CBF = 24,
    CX_LAST_X = 25, CX_LAST_Y = 43, CX_CSBF = 61, CX_SIG = 65, CX_GT1 = 107, CX_GT2 = 131, CX_MVP_IDX = 137,
    CX_SAO_MERGE = 138, CX_SAO_TYPE = 139, CX_INTER_DIR = 140, CX_REF_IDX = 145, CX_SPLIT_TU = 147, CX_QP_DELTA = 150, CX_TSKIP = 152, CX_COUNT = 154
};
#define CNU 154
/* init values (Tables 9-5..9-37), rows: initType 0 (I), 1 (P), 2 (B); the first 145 entries are the layout of the product's writer */
static const uint8_t init_values[3][CX_COUNT] = {
 { 139,141,157, CNU,CNU,CNU, CNU, CNU, 184,CNU,CNU,CNU, CNU, 184, 63, CNU,CNU, 111,141, 94,138,182,154,154, CNU,
   110,110,124,125,140,153,125,127,140,109,111,143,127,111,79,108,123,63,
   110,110,124,125,140,153,125,127,140,109,111,143,127,111,79,108,123,63,
   91,171,134,141,
   111,111,125,110,110,94,124,108,124,107,125,141,179,153,125,107,125,141,179,153,125,107,125,141,179,153,125,
   140,139,182,182,152,136,152,136,153,136,139,111,136,139,111,
   140,92,137,138,140,152,138,139,153,74,149,92,139,107,122,152,140,179,166,182,140,227,122,197,
   138,153,136,167,152,152, CNU, 153, 200, CNU,CNU,CNU,CNU,CNU, CNU,CNU, 153,138,138, 154,154, 139,139 },
 { 107,139,126, 197,185,201, 110, 122, 154,139,154,154, 149, 154, 152, 140,198, 153,111, 149,107,167,154,154, 79,
   125,110,94,110,95,79,125,111,110,78,110,111,111,95,94,108,123,108,
   125,110,94,110,95,79,125,111,110,78,110,111,111,95,94,108,123,108,
   121,140,61,154,
   155,154,139,153,139,123,123,63,153,166,183,140,136,153,154,166,183,140,136,153,154,166,183,140,136,153,154,
   170,153,123,123,107,121,107,121,167,151,183,140,151,183,140,
   154,196,196,167,154,152,167,182,182,134,149,136,153,121,136,137,169,194,166,167,154,167,137,182,
   107,167,91,122,107,167, 168, 153, 185, 95,79,63,31,31, 153,153, 124,138,94, 154,154, 139,139 },
 { 107,139,126, 197,185,201, 154, 137, 154,139,154,154, 134, 183, 152, 169,198, 153,111, 149,92,167,154,154, 79,
   125,110,124,110,95,94,125,111,111,79,125,126,111,111,79,108,123,93,
   125,110,124,110,95,94,125,111,111,79,125,126,111,111,79,108,123,93,
   121,140,61,154,
   170,154,139,153,139,123,123,63,124,166,183,140,136,153,154,166,183,140,136,153,154,166,183,140,136,153,154,
   170,153,138,138,122,121,122,121,167,151,183,140,151,183,140,
   154,196,167,167,154,152,167,182,182,134,149,136,153,121,136,122,169,208,166,167,154,152,167,182,
   107,167,91,107,107,167, 168, 153, 160, 95,79,63,31,31, 153,153, 224,167,122, 154,154, 139,139 },
};
static const uint8_t range_lps[64][4] = {
    {128,176,208,240},{128,167,197,227},{128,158,187,216},{123,150,178,205},{116,142,169,195},{111,135,160,185},{105,128,152,175},{100,122,144,166},
    {95,116,137,158},{90,110,130,150},{85,104,123,142},{81,99,117,135},{77,94,111,128},{73,89,105,122},{69,85,100,116},{66,80,95,110},
    {62,76,90,104},{59,72,86,99},{56,69,81,94},{53,65,77,89},{51,62,73,85},{48,59,69,80},{46,56,66,76},{43,53,63,72},
    {41,50,59,69},{39,48,56,65},{37,45,54,62},{35,43,51,59},{33,41,48,56},{32,39,46,53},{30,37,43,50},{29,35,41,48},
    {27,33,39,45},{26,31,37,43},{24,30,35,41},{23,28,33,39},{22,27,32,37},{21,26,30,35},{20,24,29,33},{19,23,27,31},
    {18,22,26,30},{17,21,25,28},{16,20,23,27},{15,19,22,25},{14,18,21,24},{14,17,20,23},{13,16,19,22},{12,15,18,21},
    {12,14,17,20},{11,14,16,19},{11,13,15,18},{10,12,15,17},{10,12,14,16},{9,11,13,15},{9,11,12,14},{8,10,12,14},
    {8,9,11,13},{7,9,11,12},{7,9,10,12},{7,8,10,11},{6,8,9,11},{6,7,9,10},{6,7,8,9},{2,2,2,2}};
static const uint8_t next_lps[64] = {0,0,1,2,2,4,4,5,6,7,8,9,9,11,11,12,13,13,15,15,16,16,18,18,19,19,21,21,22,22,23,24,
    24,25,26,26,27,27,28,29,29,30,30,30,31,32,32,33,33,33,34,34,35,35,35,36,36,36,37,37,37,38,38,63};

typedef struct { const uint8_t *p, *end; uint32_t range, value; int bits_needed; uint8_t ctx[CX_COUNT]; long bins; } cabd;
static uint32_t cd_byte(cabd *c) { return c->p < c->end ? *c->p++ : 0; }
static void cd_init(cabd *c, const uint8_t *p, const uint8_t *end, int init_type, int qp)
{
    c->p = p; c->end = end; c->range = 510; c->bits_needed = -8; c->bins = 0;
    c->value = cd_byte(c) << 8; c->value |= cd_byte(c);
    if (qp < 0) qp = 0; if (qp > 51) qp = 51;
    for (int i = 0; i < CX_COUNT; i++) {
        int v = init_values[init_type][i], m = (v >> 4) * 5 - 45, n = ((v & 15) << 3) - 16, pre = ((m * qp) >> 4) + n;
        if (pre < 1) pre = 1; if (pre > 126) pre = 126;
        int mps = pre > 63;
        c->ctx[i] = (uint8_t)(((mps ? pre - 64 : 63 - pre) << 1) | mps);
    }
}
/* bits consumed so far, in 1/1 bit units (arithmetic-decoder position; used for the per-category statistics) */
static long cd_pos(const cabd *c, const uint8_t *base) { return (long)(c->p - base) * 8 + c->bits_needed; }
static int cd_bin(cabd *c, int ci)
{
    uint32_t s = c->ctx[ci], state = s >> 1, mps = s & 1, lps = range_lps[state][(c->range >> 6) - 4], bin;
    c->bins++;
    c->range -= lps;
    uint32_t scaled = c->range << 7;
    if (c->value < scaled) {
        bin = mps;
        c->ctx[ci] = (uint8_t)(((state < 62 ? state + 1 : state) << 1) | mps);
        if (scaled < (256u << 7)) { c->range = scaled >> 6; c->value += c->value; if (++c->bits_needed == 0) { c->bits_needed = -8; c->value += cd_byte(c); } }
    } else {
        bin = 1 - mps;
        int nb = 0; { uint32_t r = lps; while (r < 256) { r <<= 1; nb++; } }
        c->value = (c->value - scaled) << nb; c->range = lps << nb; c->bits_needed += nb;
        if (state == 0) mps ^= 1;
        c->ctx[ci] = (uint8_t)((next_lps[state] << 1) | mps);
        if (c->bits_needed >= 0) { c->value += cd_byte(c) << c->bits_needed; c->bits_needed -= 8; }
    }
    return (int)bin;
}
static int cd_bypass(cabd *c)
{
    c->value += c->value;
    if (++c->bits_needed >= 0) { c->bits_needed = -8; c->value += cd_byte(c); }
    uint32_t scaled = c->range << 7;
    if (c->value >= scaled) { c->value -= scaled; return 1; }
    return 0;
}
static uint32_t cd_bypass_bits(cabd *c, int n) { uint32_t v = 0; while (n--) v = (v << 1) | (uint32_t)cd_bypass(c); return v; }
static int cd_terminate(cabd *c)
{
    c->range -= 2;
    uint32_t scaled = c->range << 7;
    if (c->value >= scaled) return 1;
    if (scaled < (256u << 7)) { c->range = scaled >> 6; c->value += c->value; if (++c->bits_needed == 0) { c->bits_needed = -8; c->value += cd_byte(c); } }
    return 0;
}

/* ------------------------------------------------------------------ picture-level state ------------- */
/* one record per coding unit, in decoding order (the P2 replay input; also the raw material of the statistics) */

   /* per CTU and component: 0 off / 1 band (pos = first band) / 2 edge (pos = class) */


typedef struct {
    const sps_t *sps; const pps_t *pps; cabd cd; const uint8_t *base;
    int slice_type, qp, max_merge, num_ref[2], mvd_l1_zero, sao_luma, sao_chroma, cabac_init;
    int w, h, min_cb_w, min_cb_h, pu_w, pu_h;
    uint8_t *depth;       /* per min CB: coding quadtree depth */
    uint8_t *skipf;       /* per min CB */
    uint8_t *ipm;         /* per 4x4: luma intra mode (1 = DC for non-intra) */
    uint8_t *is_intra;    /* per 4x4 */
    ora_pic_stats st;
    ora_cu_rec *cus; size_t n_cus, cap_cus;
    ora_tu_rec *tus; size_t n_tus, cap_tus;
    int16_t *lev; size_t n_lev, cap_lev;
    int err;
    int cu_pred_mode, cu_intra_luma[4], cu_intra_chroma, cu_part;    /* of the CU being parsed */
    int is_cu_qp_delta_coded;
    ora_sao_rec *sao; int ctw;    /* per CTU, merges resolved */
} pctx;

static uint8_t scan_diag4[16], scan_diag8[64], scan_diag2[4], scan_hor4[16], scan_ver4[16], scan_hor2[4], scan_ver2[4], scan_hor8[64], scan_ver8[64];
static int scans_ready;
static void build_diag(uint8_t *dst, int n, int shift)
{
    int i = 0, x = 0, y = 0;
    for (;;) { while (y >= 0) { if (x < n && y < n) dst[i++] = (uint8_t)((y << shift) | x); y--; x++; } y = x; x = 0; if (i >= n * n) break; }
}
static void init_scans(void)
{
    if (scans_ready) return;
    build_diag(scan_diag2, 2, 3); build_diag(scan_diag4, 4, 2); build_diag(scan_diag8, 8, 3);
    for (int i = 0; i < 16; i++) { scan_hor4[i] = (uint8_t)(((i >> 2) << 2) | (i & 3)); scan_ver4[i] = (uint8_t)(((i & 3) << 2) | (i >> 2)); }
    for (int i = 0; i < 4; i++) { scan_hor2[i] = (uint8_t)(((i >> 1) << 3) | (i & 1)); scan_ver2[i] = (uint8_t)(((i & 1) << 3) | (i >> 1)); }
    for (int i = 0; i < 64; i++) { scan_hor8[i] = (uint8_t)(((i >> 3) << 3) | (i & 7)); scan_ver8[i] = (uint8_t)(((i & 7) << 3) | (i >> 3)); }
    scans_ready = 1;
}

static int zavail(const pctx *p, int xc, int yc, int xn, int yn)
{   /* 6.4.1 with one slice per picture: available iff inside the picture and earlier in decoding order (z-scan at min-TB... CTB raster + z inside) */
    if (xn < 0 || yn < 0 || xn >= p->w || yn >= p->h) return 0;
    int l = p->sps->log2_ctb, cw = (p->w + (1 << l) - 1) >> l;
    int ac = (yc >> l) * cw + (xc >> l), an = (yn >> l) * cw + (xn >> l);
    if (an != ac) return an < ac;
    /* z-order inside the CTB at 4x4 granularity */
    unsigned zc = 0, zn = 0;
    for (int b = 0; b < l - 2; b++) {
        zc |= (((unsigned)(xc >> 2) >> b) & 1u) << (2 * b) | (((unsigned)(yc >> 2) >> b) & 1u) << (2 * b + 1);
        zn |= (((unsigned)(xn >> 2) >> b) & 1u) << (2 * b) | (((unsigned)(yn >> 2) >> b) & 1u) << (2 * b + 1);
    }
    return zn < zc;
}

/* ---- 7.3.8.3 sao ---- */
static void parse_sao(pctx *p, int rx, int ry)
{
    cabd *c = &p->cd;
    long b0 = cd_pos(c, p->base);
    int merge_left = 0, merge_up = 0;
    ora_sao_rec *rec = &p->sao[ry * p->ctw + rx];
    memset(rec, 0, sizeof(*rec));
    if (rx > 0) merge_left = cd_bin(c, CX_SAO_MERGE);
    if (!merge_left && ry > 0) merge_up = cd_bin(c, CX_SAO_MERGE);
    if (merge_left || merge_up) {
        p->st.sao_merge++;
        *rec = merge_left ? p->sao[ry * p->ctw + rx - 1] : p->sao[(ry - 1) * p->ctw + rx];
        /* a merge copies the candidate's parameters of the components whose SAO is on in THIS slice (7.4.9.3.2 infers the rest as off) */
        if (!p->sao_luma) rec->type[0] = 0;
        if (!p->sao_chroma) rec->type[1] = rec->type[2] = 0;
    } else {
        int type = 0;
        for (int ci = 0; ci < 3; ci++) {
            if ((ci == 0 && !p->sao_luma) || (ci > 0 && !p->sao_chroma)) continue;
            if (ci < 2) { type = cd_bin(c, CX_SAO_TYPE); if (type) type = cd_bypass(c) ? 2 : 1; if (type) { if (ci == 0) p->st.sao_on_luma++; else p->st.sao_on_chroma++; } }
            rec->type[ci] = (uint8_t)type;
            if (!type) continue;
            int off[4];
            for (int k = 0; k < 4; k++) { int a = 0; while (a < 7 && cd_bypass(c)) a++; off[k] = a; }
            if (type == 1) { for (int k = 0; k < 4; k++) if (off[k] && cd_bypass(c)) off[k] = -off[k]; rec->pos[ci] = (uint8_t)cd_bypass_bits(c, 5); }
            else { if (ci < 2) rec->pos[ci] = (uint8_t)cd_bypass_bits(c, 2); else rec->pos[2] = rec->pos[1]; off[2] = -off[2]; off[3] = -off[3]; }
            for (int k = 0; k < 4; k++) rec->off[ci][k] = (int8_t)off[k];
        }
    }
    p->st.bits_sao += cd_pos(c, p->base) - b0;
}

/* ---- 7.3.8.11 residual_coding ---- */
static const uint8_t ctx_idx_map4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
static void parse_residual(pctx *p, int x0, int y0, int log2, int cidx, int16_t *out /* n*n raster, zeroed */)
{
    cabd *c = &p->cd; (void)x0; (void)y0;
    int n = 1 << log2;
    if (p->pps->tskip && log2 == 2) cd_bin(c, CX_TSKIP + (cidx ? 1 : 0));
    /* last significant position */
    int off, shift;
    if (cidx == 0) { off = 3 * (log2 - 2) + ((log2 - 1) >> 2); shift = (log2 + 1) >> 2; } else { off = 15; shift = log2 - 2; }
    int cmax = (log2 << 1) - 1, px = 0, py = 0;
    while (px < cmax && cd_bin(c, CX_LAST_X + off + (px >> shift))) px++;
    while (py < cmax && cd_bin(c, CX_LAST_Y + off + (py >> shift))) py++;
    int lx = px, ly = py;
    if (px > 3) { int nb = (px >> 1) - 1; lx = (1 << nb) * (2 + (px & 1)) + (int)cd_bypass_bits(c, nb); }
    if (py > 3) { int nb = (py >> 1) - 1; ly = (1 << nb) * (2 + (py & 1)) + (int)cd_bypass_bits(c, nb); }
    /* scanIdx (7.4.9.11): mode dependent for intra 4x4 / 8x8 luma and 4x4 chroma (of 8x8 luma) */
    int scan_idx = 0;
    if (p->cu_pred_mode == 1 && (log2 == 2 || (log2 == 3 && cidx == 0))) {
        int m = cidx == 0 ? p->cu_intra_luma[0] : p->cu_intra_chroma;     /* caller sets cu_intra_luma[0] to the mode of THIS TB's partition */
        if (m >= 6 && m <= 14) scan_idx = 2; else if (m >= 22 && m <= 30) scan_idx = 1;
    }
    if (scan_idx == 2) { int t = lx; lx = ly; ly = t; }
    const uint8_t *s4 = scan_idx == 0 ? scan_diag4 : (scan_idx == 1 ? scan_hor4 : scan_ver4);
    const uint8_t *scg; int cgw = n >> 2;
    if (log2 == 2) scg = (const uint8_t *)"\0";
    else if (log2 == 3) scg = scan_idx == 0 ? scan_diag2 : (scan_idx == 1 ? scan_hor2 : scan_ver2);
    else if (log2 == 4) { static uint8_t d4x[16]; static int rdy; if (!rdy) { build_diag(d4x, 4, 3); rdy = 1; } scg = d4x; }
    else scg = scan_diag8;
    /* horizontal scan of an 8x8 block walks the 4x4 groups row by row: groups (0,0),(1,0),(0,1),(1,1) -> that is scan_hor2 above (x | y << 3) */
    int last_cg = 0, last_pos = 0;
    {
        int cgx = lx >> 2, cgy = ly >> 2, pxy = ((ly & 3) << 2) | (lx & 3);
        for (int i = 0; i < cgw * cgw; i++) if ((scg[i] & 7) == cgx && (scg[i] >> 3) == cgy) { last_cg = i; break; }
        for (int i = 0; i < 16; i++) if (s4[i] == pxy) { last_pos = i; break; }
    }
    uint8_t csbf[8][8]; memset(csbf, 0, sizeof(csbf));
    int c1 = 1, is_luma = cidx == 0;
    long nz = 0, sabs = 0;
    for (int i = last_cg; i >= 0; i--) {
        int cx = scg[i] & 7, cy = scg[i] >> 3;
        int right = cx + 1 < cgw ? csbf[cy][cx + 1] : 0, below = cy + 1 < cgw ? csbf[cy + 1][cx] : 0;
        int coded = 1, infer_dc = 0;
        if (i < last_cg && i > 0) { coded = cd_bin(c, CX_CSBF + ((right | below) ? 1 : 0) + (cidx ? 2 : 0)); infer_dc = 1; }
        csbf[cy][cx] = (uint8_t)coded;
        if (!coded) continue;
        int prev = right | (below << 1);
        uint8_t sig[16]; memset(sig, 0, 16);
        int start = 15, nsig = 0;
        if (i == last_cg) { start = last_pos - 1; sig[last_pos] = 1; nsig = 1; }
        for (int k = start; k >= 0; k--) {
            int xp = s4[k] & 3, yp = s4[k] >> 2, sc;
            if (k == 0 && infer_dc && nsig == 0) { sig[0] = 1; nsig++; break; }
            /* 9.3.4.2.5 */
            if (log2 == 2) sc = ctx_idx_map4[(yp << 2) + xp];
            else if (cx == 0 && cy == 0 && xp == 0 && yp == 0) sc = 0;
            else {
                if (prev == 0) sc = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
                else if (prev == 1) sc = yp == 0 ? 2 : yp == 1 ? 1 : 0;
                else if (prev == 2) sc = xp == 0 ? 2 : xp == 1 ? 1 : 0;
                else sc = 2;
                if (is_luma) { if (cx || cy) sc += 3; sc += log2 == 3 ? (scan_idx == 0 ? 9 : 15) : 21; }
                else sc += log2 == 3 ? 9 : 12;
            }
            if (cd_bin(c, CX_SIG + sc + (is_luma ? 0 : 27))) { sig[k] = 1; nsig++; }
        }
        if (!nsig) continue;
        /* levels, high scan position first */
        int pos[16], np = 0;
        for (int k = 15; k >= 0; k--) if (sig[k]) pos[np++] = k;
        int absv[16];
        int ctx_set = (i > 0 && is_luma) ? 2 : 0;
        if (c1 == 0) ctx_set++;
        c1 = 1;
        int first_g1 = -1, ng1 = np < 8 ? np : 8;
        for (int k = 0; k < np; k++) absv[k] = 1;
        for (int k = 0; k < ng1; k++) {
            int g1 = cd_bin(c, CX_GT1 + (is_luma ? 0 : 16) + 4 * ctx_set + c1);
            if (g1) { absv[k] = 2; c1 = 0; if (first_g1 < 0) first_g1 = k; }
            else if (c1 < 3 && c1 > 0) c1++;
        }
        if (first_g1 >= 0) { if (cd_bin(c, CX_GT2 + (is_luma ? 0 : 4) + ctx_set)) absv[first_g1] = 3; }
        int hidden = p->pps->sign_hiding && (pos[0] - pos[np - 1] > 3);
        uint32_t signs = cd_bypass_bits(c, hidden ? np - 1 : np);
        if (hidden) signs <<= 1;
        int rice = 0, sum = 0;
        for (int k = 0; k < np; k++) {
            int base = k < 8 ? (k == first_g1 ? 3 : 2) : 1;
            if (absv[k] == base) {
                int pre = 0;
                while (pre < 4 && cd_bypass(c)) pre++;
                int rem;
                if (pre < 4) rem = (pre << rice) + (int)cd_bypass_bits(c, rice);
                else { int e = 0; while (e < 28 && cd_bypass(c)) e++; pre = 4 + e; rem = (((1 << (pre - 3)) + 3 - 1) << rice) + (int)cd_bypass_bits(c, pre - 3 + rice); }
                absv[k] = base + rem;
                if (absv[k] > 3 * (1 << rice) && rice < 4) rice++;
            }
            sum += absv[k];
        }
        for (int k = 0; k < np; k++) {
            int neg = (signs >> (np - 1 - k)) & 1;
            if (hidden && k == np - 1) neg = sum & 1;
            int xx = (cx << 2) + (s4[pos[k]] & 3), yy = (cy << 2) + (s4[pos[k]] >> 2);
            out[yy * n + xx] = (int16_t)(neg ? -absv[k] : absv[k]);
            nz++; sabs += absv[k];
        }
    }
    if (is_luma) { p->st.nz_luma += nz; p->st.sum_abs_luma += sabs; } else { p->st.nz_chroma += nz; p->st.sum_abs_chroma += sabs; }
}

static int16_t *lev_alloc(pctx *p, int n2, uint32_t *off)
{
    if (p->n_lev + (size_t)n2 > p->cap_lev) { p->cap_lev = (p->cap_lev + (size_t)n2) * 2; p->lev = (int16_t *)realloc(p->lev, p->cap_lev * 2); }
    *off = (uint32_t)p->n_lev; memset(p->lev + p->n_lev, 0, (size_t)n2 * 2); p->n_lev += (size_t)n2;
    return p->lev + *off;
}
static ora_tu_rec *tu_new(pctx *p)
{
    if (p->n_tus == p->cap_tus) { p->cap_tus = p->cap_tus ? p->cap_tus * 2 : 4096; p->tus = (ora_tu_rec *)realloc(p->tus, p->cap_tus * sizeof(ora_tu_rec)); }
    ora_tu_rec *t = &p->tus[p->n_tus++]; memset(t, 0, sizeof(*t)); t->lev_off[0] = t->lev_off[1] = t->lev_off[2] = ~0u;
    return t;
}

/* ---- 7.3.8.8 transform_tree / 7.3.8.10 transform_unit ---- */
static void parse_transform_tree(pctx *p, int x0, int y0, int xb, int yb, int log2, int depth, int blk, int cbf_cb_p, int cbf_cr_p, int max_depth, int inter_split)
{
    cabd *c = &p->cd; const sps_t *s = p->sps;
    int split;
    int intra_split = p->cu_pred_mode == 1 && p->cu_part == 3;
    if (log2 <= s->log2_max_tb && log2 > s->log2_min_tb && depth < max_depth && !(intra_split && depth == 0)) split = cd_bin(c, CX_SPLIT_TU + 5 - log2);
    else split = log2 > s->log2_max_tb || (intra_split && depth == 0) || (inter_split && depth == 0);
    int cbf_cb = 0, cbf_cr = 0;
    if (log2 > 2) {
        if (depth == 0 || cbf_cb_p) cbf_cb = cd_bin(c, CX_CBF_CHROMA + depth);
        if (depth == 0 || cbf_cr_p) cbf_cr = cd_bin(c, CX_CBF_CHROMA + depth);
    } else { cbf_cb = cbf_cb_p; cbf_cr = cbf_cr_p; }     /* 4x4 luma blocks inherit: the chroma block hangs on blkIdx 3 */
    if (split) {
        int h = 1 << (log2 - 1);
        for (int k = 0; k < 4; k++) parse_transform_tree(p, x0 + (k & 1) * h, y0 + (k >> 1) * h, x0, y0, log2 - 1, depth + 1, k, cbf_cb, cbf_cr, max_depth, 0);
        return;
    }
    int cbf_luma = 1;
    if (p->cu_pred_mode == 1 || depth != 0 || cbf_cb || cbf_cr) cbf_luma = cd_bin(c, CX_CBF_LUMA + (depth == 0 ? 1 : 0));
    ora_tu_rec *t = tu_new(p);
    t->x = (uint16_t)x0; t->y = (uint16_t)y0; t->log2 = (uint8_t)log2;
    int chroma_here = log2 > 2 || blk == 3, any_c = cbf_cb || cbf_cr;     /* 7.3.8.10: cbfChroma looks at the parent's flags for every 4x4 luma block, not only blkIdx 3 */
    if (cbf_luma || any_c) {
        if (p->pps->cu_qp_delta && !p->is_cu_qp_delta_coded) {
            int a = 0;
            while (a < 5 && cd_bin(c, CX_QP_DELTA + (a ? 1 : 0))) a++;
            if (a == 5) { int k = 0; while (k < 16 && cd_bypass(c)) { a += 1 << k; k++; } a += (int)cd_bypass_bits(c, k); }
            if (a && cd_bypass(c)) a = -a;
            t->qp_delta = (int8_t)a; p->is_cu_qp_delta_coded = 1;
        }
        long b0 = cd_pos(c, p->base);
        /* the luma mode of this TB's partition drives its scan: for NxN the partition index is the 8x8-relative quadrant of the TB */
        int save = p->cu_intra_luma[0];
        if (p->cu_pred_mode == 1 && p->cu_part == 3) {      /* NxN: the partition = quadrant of (x0,y0) inside the CU */
            const ora_cu_rec *cu = &p->cus[p->n_cus - 1];
            int hs = 1 << (cu->log2 - 1), q = ((x0 - cu->x) >= hs ? 1 : 0) | ((y0 - cu->y) >= hs ? 2 : 0);
            p->cu_intra_luma[0] = cu->intra_mode[q];
        }
        if (cbf_luma) { t->cbf |= 1; parse_residual(p, x0, y0, log2, 0, lev_alloc(p, 1 << (2 * log2), &t->lev_off[0])); p->st.n_cbf_luma++; }
        p->cu_intra_luma[0] = save;
        long b1 = cd_pos(c, p->base);
        p->st.bits_luma += b1 - b0;
        if (chroma_here) {
            int l2c = log2 > 2 ? log2 - 1 : 2, xc = log2 > 2 ? x0 : xb, yc = log2 > 2 ? y0 : yb;
            if (cbf_cb) { t->cbf |= 2; parse_residual(p, xc, yc, l2c, 1, lev_alloc(p, 1 << (2 * l2c), &t->lev_off[1])); p->st.n_cbf_chroma++; }
            if (cbf_cr) { t->cbf |= 4; parse_residual(p, xc, yc, l2c, 2, lev_alloc(p, 1 << (2 * l2c), &t->lev_off[2])); p->st.n_cbf_chroma++; }
        }
        p->st.bits_chroma += cd_pos(c, p->base) - b1;
    }
    p->st.n_tu[log2 - 2]++;
}

static void parse_mvd(pctx *p, int16_t out[2])
{
    cabd *c = &p->cd;
    int gx = cd_bin(c, CX_MVD), gy = cd_bin(c, CX_MVD), g1x = 0, g1y = 0;
    if (gx) g1x = cd_bin(c, CX_MVD + 1);
    if (gy) g1y = cd_bin(c, CX_MVD + 1);
    int v[2] = {0, 0};
    for (int k = 0; k < 2; k++) {
        int g = k ? gy : gx, g1 = k ? g1y : g1x;
        if (!g) continue;
        int a = 1;
        if (g1) { int kk = 1, base = 0; while (kk < 32 && cd_bypass(c)) { base += 1 << kk; kk++; } a = 2 + base + (int)cd_bypass_bits(c, kk); }
        v[k] = cd_bypass(c) ? -a : a;
    }
    out[0] = (int16_t)v[0]; out[1] = (int16_t)v[1];
    if (v[0] || v[1]) p->st.n_mvd_nonzero++;
}

/* ---- 7.3.8.5 coding_unit ---- */
static void parse_cu(pctx *p, int x0, int y0, int log2)
{
    cabd *c = &p->cd; const sps_t *s = p->sps;
    int size = 1 << log2, mcb = s->log2_min_cb, li = log2 - 3;
    if (p->n_cus == p->cap_cus) { p->cap_cus = p->cap_cus ? p->cap_cus * 2 : 4096; p->cus = (ora_cu_rec *)realloc(p->cus, p->cap_cus * sizeof(ora_cu_rec)); }
    ora_cu_rec *cu = &p->cus[p->n_cus++]; memset(cu, 0, sizeof(*cu));
    cu->x = (uint16_t)x0; cu->y = (uint16_t)y0; cu->log2 = (uint8_t)log2; cu->first_tu = (uint32_t)p->n_tus;
    long b0 = cd_pos(c, p->base), hdr_excl = 0;
    int skip = 0;
    p->st.n_cu[li]++;
    if (p->slice_type != KS_SLICE_I) {
        int ctx = 0, cxm = x0 >> mcb, cym = y0 >> mcb;
        if (x0 > 0) ctx += p->skipf[cym * p->min_cb_w + cxm - 1];
        if (y0 > 0) ctx += p->skipf[(cym - 1) * p->min_cb_w + cxm];
        skip = cd_bin(c, CX_SKIP + ctx);
    }
    for (int yy = y0 >> mcb; yy < ((y0 + size) >> mcb) && yy < p->min_cb_h; yy++) for (int xx = x0 >> mcb; xx < ((x0 + size) >> mcb) && xx < p->min_cb_w; xx++) p->skipf[yy * p->min_cb_w + xx] = (uint8_t)skip;
    int pred_intra = p->slice_type == KS_SLICE_I, part = 0;
    cu->skip = (uint8_t)skip;
    if (skip) {
        int idx = 0;
        if (p->max_merge > 1) { idx = cd_bin(c, CX_MERGE_IDX); if (idx) while (idx < p->max_merge - 1 && cd_bypass(c)) idx++; }
        cu->merge[0] = 1; cu->merge_idx[0] = (uint8_t)idx;
        p->st.n_skip[li]++;
        p->cu_pred_mode = 0;
    } else {
        if (p->slice_type != KS_SLICE_I) pred_intra = cd_bin(c, CX_PRED_MODE);
        p->cu_pred_mode = pred_intra;
        if (!pred_intra || log2 == mcb) {
            /* part_mode (9.3.4.2.x, Table 9-43): intra: 1 -> 2Nx2N, 0 -> NxN; inter: 1 -> 2Nx2N, 01 -> 2NxN, 001/00 -> Nx2N (/NxN) */
            if (cd_bin(c, CX_PART_MODE)) part = 0;
            else if (pred_intra) part = 3;
            else if (log2 == mcb) {
                if (cd_bin(c, CX_PART_MODE + 1)) part = 1;
                else if (log2 == 3) part = 2;
                else part = cd_bin(c, CX_PART_MODE + 2) ? 2 : 3;
            } else if (!s->amp) part = cd_bin(c, CX_PART_MODE + 1) ? 1 : 2;
            else {      /* AMP */
                int horiz = cd_bin(c, CX_PART_MODE + 1);
                if (cd_bin(c, CX_PART_MODE + 3)) part = horiz ? 1 : 2;
                else part = (horiz ? 4 : 6) + cd_bypass(c);
            }
        }
        p->cu_part = part;
        cu->pred_mode = (uint8_t)pred_intra; cu->part_mode = (uint8_t)part;
        if (pred_intra) {
            long bm0 = cd_pos(c, p->base);
            int np = part == 3 ? 4 : 1, prev[4], ps = part == 3 ? size >> 1 : size;
            for (int k = 0; k < np; k++) prev[k] = cd_bin(c, CX_PREV_INTRA);
            for (int k = 0; k < np; k++) {
                int xp = x0 + (k & 1) * ps, yp = y0 + (k >> 1) * ps;
                int ca = 1, cb = 1;
                if (zavail(p, xp, yp, xp - 1, yp) && p->is_intra[(yp >> 2) * p->pu_w + ((xp - 1) >> 2)]) ca = p->ipm[(yp >> 2) * p->pu_w + ((xp - 1) >> 2)];
                if (zavail(p, xp, yp, xp, yp - 1) && p->is_intra[((yp - 1) >> 2) * p->pu_w + (xp >> 2)] && ((yp - 1) >> s->log2_ctb) == (yp >> s->log2_ctb)) cb = p->ipm[((yp - 1) >> 2) * p->pu_w + (xp >> 2)];
                int mpm[3];
                if (ca == cb) { if (ca < 2) { mpm[0] = 0; mpm[1] = 1; mpm[2] = 26; } else { mpm[0] = ca; mpm[1] = 2 + ((ca + 29) & 31); mpm[2] = 2 + ((ca - 2 + 1) & 31); } }
                else { mpm[0] = ca; mpm[1] = cb; mpm[2] = (ca != 0 && cb != 0) ? 0 : ((ca != 1 && cb != 1) ? 1 : 26); }
                int mode;
                if (prev[k]) { int mi = cd_bypass(c); if (mi) mi += cd_bypass(c); mode = mpm[mi]; }
                else {
                    int rem = (int)cd_bypass_bits(c, 5);
                    if (mpm[0] > mpm[1]) { int t = mpm[0]; mpm[0] = mpm[1]; mpm[1] = t; }
                    if (mpm[0] > mpm[2]) { int t = mpm[0]; mpm[0] = mpm[2]; mpm[2] = t; }
                    if (mpm[1] > mpm[2]) { int t = mpm[1]; mpm[1] = mpm[2]; mpm[2] = t; }
                    mode = rem; for (int j = 0; j < 3; j++) if (mode >= mpm[j]) mode++;
                }
                cu->intra_mode[k] = (uint8_t)mode; p->cu_intra_luma[k] = mode;
                for (int yy = yp >> 2; yy < ((yp + ps) >> 2) && yy < p->pu_h; yy++) for (int xx = xp >> 2; xx < ((xp + ps) >> 2) && xx < p->pu_w; xx++) { p->ipm[yy * p->pu_w + xx] = (uint8_t)mode; p->is_intra[yy * p->pu_w + xx] = 1; }
            }
            int cm = 4;
            if (cd_bin(c, CX_CHROMA_PRED)) cm = (int)cd_bypass_bits(c, 2);
            static const int cmodes[4] = {0, 26, 10, 1};
            int lm = p->cu_intra_luma[0];
            p->cu_intra_chroma = cm == 4 ? lm : (cmodes[cm] == lm ? 34 : cmodes[cm]);
            cu->chroma_mode = (uint8_t)p->cu_intra_chroma;
            p->st.n_intra[li]++; if (part == 3) p->st.n_intra_nxn++;
            p->st.bits_intra_mode += cd_pos(c, p->base) - bm0;
        } else {
            int npu = part == 0 ? 1 : (part == 3 ? 4 : 2);
            for (int k = 0; k < npu; k++) {
                int mf = cd_bin(c, CX_MERGE_FLAG);
                cu->merge[k] = (uint8_t)mf;
                if (mf) {
                    int idx = 0;
                    if (p->max_merge > 1) { idx = cd_bin(c, CX_MERGE_IDX); if (idx) while (idx < p->max_merge - 1 && cd_bypass(c)) idx++; }
                    cu->merge_idx[k] = (uint8_t)idx;
                    if (k == 0) p->st.n_merge[li]++;
                } else {
                    long bv0 = cd_pos(c, p->base);
                    int dir = 1;
                    if (p->slice_type == KS_SLICE_B) {
                        int pw = part == 2 || part >= 6 ? size >> 1 : size, ph = part == 1 || (part >= 4 && part < 6) ? size >> 1 : size;
                        if (part == 3) { pw = ph = size >> 1; }
                        if (pw + ph != 12 && cd_bin(c, CX_INTER_DIR + (s->log2_ctb - log2))) dir = 3;
                        else dir = cd_bin(c, CX_INTER_DIR + 4) ? 2 : 1;
                    }
                    cu->inter_dir[k] = (uint8_t)dir;
                    for (int X = 0; X < 2; X++) {
                        if (!(dir & (1 << X))) continue;
                        int nr = p->num_ref[X], ri = 0;
                        if (nr > 1) { while (ri < nr - 1 && ri < 2 && cd_bin(c, CX_REF_IDX + ri)) ri++; if (ri == 2) while (ri < nr - 1 && cd_bypass(c)) ri++; }
                        cu->ref_idx[k][X] = (uint8_t)ri;
                        if (X == 1 && p->mvd_l1_zero && dir == 3) { cu->mvd[k][1][0] = cu->mvd[k][1][1] = 0; }
                        else parse_mvd(p, cu->mvd[k][X]);
                        cu->mvp[k][X] = (uint8_t)cd_bin(c, CX_MVP_IDX);
                    }
                    if (k == 0) p->st.n_amvp[li]++;
                    p->st.bits_mvd += cd_pos(c, p->base) - bv0;
                }
            }
        }
        int root = 1;
        if (!pred_intra && !(part == 0 && cu->merge[0])) root = cd_bin(c, CX_ROOT_CBF);
        cu->root_cbf = (uint8_t)root;
        if (root) {
            long bt0 = cd_pos(c, p->base), l0 = p->st.bits_luma + p->st.bits_chroma;
            int max_depth = pred_intra ? s->tu_depth_intra + (part == 3) : s->tu_depth_inter;
            int inter_split = s->tu_depth_inter == 0 && !pred_intra && part != 0;
            parse_transform_tree(p, x0, y0, x0, y0, log2, 0, 0, 0, 0, max_depth, inter_split);
            hdr_excl = (p->st.bits_luma + p->st.bits_chroma) - l0; (void)bt0;
        }
    }
    if (!pred_intra || skip) {
        for (int yy = y0 >> 2; yy < ((y0 + size) >> 2) && yy < p->pu_h; yy++) for (int xx = x0 >> 2; xx < ((x0 + size) >> 2) && xx < p->pu_w; xx++) { p->ipm[yy * p->pu_w + xx] = 1; p->is_intra[yy * p->pu_w + xx] = 0; }
    }
    cu->n_tu = (uint32_t)p->n_tus - cu->first_tu;
    p->st.bits_cu_hdr += cd_pos(c, p->base) - b0 - hdr_excl;
}

/* ---- 7.3.8.4 coding_quadtree ---- */
static void parse_quadtree(pctx *p, int x0, int y0, int log2, int depth)
{
    cabd *c = &p->cd; const sps_t *s = p->sps;
    int size = 1 << log2, split, mcb = s->log2_min_cb;
    if (p->err) return;
    if (p->pps->cu_qp_delta && log2 >= s->log2_ctb - p->pps->diff_cu_qp_delta_depth) p->is_cu_qp_delta_coded = 0;     /* start of a quantisation group */
    if (x0 + size <= p->w && y0 + size <= p->h && log2 > mcb) {
        int ctx = 0;
        if (x0 > 0 && zavail(p, x0, y0, x0 - 1, y0)) ctx += p->depth[(y0 >> mcb) * p->min_cb_w + ((x0 - 1) >> mcb)] > depth;
        if (y0 > 0 && zavail(p, x0, y0, x0, y0 - 1)) ctx += p->depth[((y0 - 1) >> mcb) * p->min_cb_w + (x0 >> mcb)] > depth;
        long b0 = cd_pos(c, p->base);
        split = cd_bin(c, CX_SPLIT_CU + ctx);
        p->st.bits_split += cd_pos(c, p->base) - b0;
    } else split = log2 > mcb;
    if (split) {
        int h = size >> 1;
        for (int k = 0; k < 4; k++) { int xx = x0 + (k & 1) * h, yy = y0 + (k >> 1) * h; if (xx < p->w && yy < p->h) parse_quadtree(p, xx, yy, log2 - 1, depth + 1); }
    } else {
        for (int yy = y0 >> mcb; yy < ((y0 + size) >> mcb) && yy < p->min_cb_h; yy++) for (int xx = x0 >> mcb; xx < ((x0 + size) >> mcb) && xx < p->min_cb_w; xx++) p->depth[yy * p->min_cb_w + xx] = (uint8_t)depth;
        parse_cu(p, x0, y0, log2);
    }
}

/* ------------------------------------------------------------------ stream level --------------------- */


static size_t unescape(const uint8_t *in, size_t n, uint8_t *out)
{
    size_t o = 0; int z = 0;
    for (size_t i = 0; i < n; i++) {
        if (z >= 2 && in[i] == 3) { z = 0; continue; }
        out[o++] = in[i]; z = in[i] == 0 ? z + 1 : 0;
    }
    return o;
}

static int parse_slice(const sps_t *sps, const pps_t *pps, int nal_type, const uint8_t *rb, size_t n, int *prev_poc_tid0, ora_parsed_pic *out)
{
    bitr r = {rb, n, 0};
    int first = br_u(&r, 1);
    if (nal_type >= 16 && nal_type <= 23) br_u(&r, 1);
    br_ue(&r);
    if (!first) return -10;                                      /* one slice per picture only */
    br_u(&r, pps->extra_bits);
    int st = (int)br_ue(&r);
    if (pps->output_flag_present) br_u(&r, 1);
    int poc = 0; rps_t rps; memset(&rps, 0, sizeof(rps));
    int tmvp = 0, n_lt_used = 0;
    if (nal_type != 19 && nal_type != 20) {
        int lsb = (int)br_u(&r, sps->log2_max_poc), maxl = 1 << sps->log2_max_poc, prev = *prev_poc_tid0, plsb = prev & (maxl - 1), pmsb = prev - plsb, msb;
        if (lsb < plsb && plsb - lsb >= maxl / 2) msb = pmsb + maxl; else if (lsb > plsb && lsb - plsb > maxl / 2) msb = pmsb - maxl; else msb = pmsb;
        poc = msb + lsb;
        if (!br_u(&r, 1)) { sps_t tmp = *sps; if (parse_rps(&r, &tmp, sps->n_rps, sps->n_rps, &rps, 1)) return -11; }
        else { int bits = 0; while ((1 << bits) < sps->n_rps) bits++; rps = sps->rps[bits ? br_u(&r, bits) : 0]; }
        if (sps->long_term) {
            int n_sps = sps->n_lt_sps > 0 ? (int)br_ue(&r) : 0, n_pics = (int)br_ue(&r), bits = 0;
            while ((1 << bits) < sps->n_lt_sps) bits++;
            for (int i = 0; i < n_sps + n_pics; i++) {
                if (i < n_sps) { if (sps->n_lt_sps > 1) br_u(&r, bits); } else { br_u(&r, sps->log2_max_poc); n_lt_used += (int)br_u(&r, 1); }
                if (br_u(&r, 1)) br_ue(&r);
            }
        }
        if (sps->tmvp) tmvp = br_u(&r, 1);
    }
    *prev_poc_tid0 = poc;
    int sao_l = 0, sao_c = 0;
    if (sps->sao) { sao_l = br_u(&r, 1); sao_c = br_u(&r, 1); }
    int nref[2] = {0, 0}, max_merge = 5, mvd_l1_zero = 0, cabac_init = 0, col_ref_idx = 0, col_from_l0 = 1;
    if (st != KS_SLICE_I) {
        nref[0] = pps->ref_l0; nref[1] = st == KS_SLICE_B ? pps->ref_l1 : 0;
        if (br_u(&r, 1)) { nref[0] = (int)br_ue(&r) + 1; if (st == KS_SLICE_B) nref[1] = (int)br_ue(&r) + 1; }
        int npt = n_lt_used; for (int i = 0; i < rps.n_neg + rps.n_pos; i++) npt += rps.used[i];
        if (pps->lists_mod && npt > 1) { int bits = 0; while ((1 << bits) < npt) bits++; for (int X = 0; X < (st == KS_SLICE_B ? 2 : 1); X++) if (br_u(&r, 1)) for (int i = 0; i < nref[X]; i++) br_u(&r, bits); }
        if (st == KS_SLICE_B) mvd_l1_zero = br_u(&r, 1);
        if (pps->cabac_init_present) cabac_init = br_u(&r, 1);
        if (tmvp) { if (st == KS_SLICE_B) col_from_l0 = br_u(&r, 1); if ((col_from_l0 && nref[0] > 1) || (!col_from_l0 && nref[1] > 1)) col_ref_idx = (int)br_ue(&r); }
        max_merge = 5 - (int)br_ue(&r);
    }
    int qp = pps->init_qp + br_se(&r);
    int cb_off = pps->cb_off, cr_off = pps->cr_off, beta_off = pps->beta, tc_off = pps->tc;
    if (pps->slice_chroma_off) { cb_off += br_se(&r); cr_off += br_se(&r); }
    int dbk_override = 0, dbk_disabled = pps->deblock_disabled;
    if (pps->deblock_override) dbk_override = br_u(&r, 1);
    if (dbk_override) { dbk_disabled = br_u(&r, 1); if (!dbk_disabled) { beta_off = br_se(&r); tc_off = br_se(&r); } }
    if (pps->lf_across && (sao_l || sao_c || !dbk_disabled)) br_u(&r, 1);
    if (pps->tiles || pps->wpp) { int ne = (int)br_ue(&r); if (ne) { int ol = (int)br_ue(&r) + 1; for (int i = 0; i < ne; i++) br_u(&r, ol); if (ne) return -12; } }
    if (pps->slice_ext) { int l = (int)br_ue(&r); for (int i = 0; i < l; i++) br_u(&r, 8); }
    br_u(&r, 1); while (r.pos & 7) br_u(&r, 1);                  /* byte_alignment */

    pctx p; memset(&p, 0, sizeof(p));
    p.sps = sps; p.pps = pps; p.slice_type = st; p.qp = qp; p.max_merge = max_merge; p.num_ref[0] = nref[0]; p.num_ref[1] = nref[1];
    p.mvd_l1_zero = mvd_l1_zero; p.sao_luma = sao_l; p.sao_chroma = sao_c;
    p.w = sps->w; p.h = sps->h;
    p.min_cb_w = (p.w + (1 << sps->log2_min_cb) - 1) >> sps->log2_min_cb; p.min_cb_h = (p.h + (1 << sps->log2_min_cb) - 1) >> sps->log2_min_cb;
    p.pu_w = (p.w + 3) >> 2; p.pu_h = (p.h + 3) >> 2;
    p.depth = (uint8_t *)calloc((size_t)p.min_cb_w * p.min_cb_h, 1); p.skipf = (uint8_t *)calloc((size_t)p.min_cb_w * p.min_cb_h, 1);
    p.ipm = (uint8_t *)malloc((size_t)p.pu_w * p.pu_h); memset(p.ipm, 1, (size_t)p.pu_w * p.pu_h); p.is_intra = (uint8_t *)calloc((size_t)p.pu_w * p.pu_h, 1);
    int init_type = st == KS_SLICE_I ? 0 : (st == KS_SLICE_P ? (cabac_init ? 2 : 1) : (cabac_init ? 1 : 2));
    p.base = rb + (r.pos >> 3);
    cd_init(&p.cd, p.base, rb + n, init_type, qp);
    p.st.poc = poc; p.st.slice_type = st; p.st.qp = qp; p.st.nal_type = nal_type; p.st.num_ref[0] = nref[0]; p.st.num_ref[1] = nref[1];
    int l = sps->log2_ctb, ctw = (p.w + (1 << l) - 1) >> l, cth = (p.h + (1 << l) - 1) >> l, ok = 1;
    p.sao = (ora_sao_rec *)calloc((size_t)ctw * cth, sizeof(ora_sao_rec)); p.ctw = ctw;
    for (int a = 0; a < ctw * cth && ok; a++) {
        int rx = a % ctw, ry = a / ctw;
        if (sao_l || sao_c) parse_sao(&p, rx, ry);
        parse_quadtree(&p, rx << l, ry << l, l, 0);
        int end = cd_terminate(&p.cd);
        if (end != (a == ctw * cth - 1)) { ok = 0; if (getenv("ORA_PARSE_DEBUG")) { fprintf(stderr, "poc %d: terminate bin %d at CTU %d of %d; CUs of this CTU:\n", poc, end, a, ctw * cth); for (size_t q = 0; q < p.n_cus; q++) if ((p.cus[q].x >> l) == rx && (p.cus[q].y >> l) == ry) fprintf(stderr, "  cu (%d,%d) log2 %d intra %d part %d skip %d modes %d %d %d %d chroma %d ntu %u\n", p.cus[q].x, p.cus[q].y, p.cus[q].log2, p.cus[q].pred_mode, p.cus[q].part_mode, p.cus[q].skip, p.cus[q].intra_mode[0], p.cus[q].intra_mode[1], p.cus[q].intra_mode[2], p.cus[q].intra_mode[3], p.cus[q].chroma_mode, p.cus[q].n_tu); } }
        if (p.cd.p > p.cd.end + 2) ok = 0;
    }
    /* after end_of_slice_segment_flag = 1 the decoder must sit within the last bytes of the RBSP (9.3.2.5 reads rbsp_trailing_bits) */
    if (ok && (size_t)(p.cd.p - rb) + 4 < n) ok = 0;
    p.st.bits_total = (long)(n - (r.pos >> 3)) * 8;
    out->sao = p.sao; out->dbk_disabled = dbk_disabled; out->beta_off_div2 = beta_off; out->tc_off_div2 = tc_off; out->cb_qp_off = cb_off; out->cr_qp_off = cr_off;
    out->cu_qp_delta_enabled = pps->cu_qp_delta; out->any_qp_delta = 0;
    out->tmvp = tmvp; out->col_ref_idx = col_ref_idx; out->max_merge = max_merge; out->par_mrg_level = pps->par_mrg; out->n_list0 = 0; out->n_list1 = 0;
    if (st != KS_SLICE_I) {       /* 8.3.4 without list modification: the used short-term pictures before, then after the current one, repeated cyclically */
        int cand[32], nc = 0;
        for (int i = 0; i < rps.n_neg + rps.n_pos; i++) if (rps.used[i]) cand[nc++] = poc + rps.dpoc[i];
        for (int i = 0; i < nref[0] && i < 16 && nc; i++) out->list0_poc[out->n_list0++] = cand[i % nc];
        /* list 1: the pictures after the current one first, then those before it */
        int c1[32], n1 = 0;
        for (int i = rps.n_neg; i < rps.n_neg + rps.n_pos; i++) if (rps.used[i]) c1[n1++] = poc + rps.dpoc[i];
        for (int i = 0; i < rps.n_neg; i++) if (rps.used[i]) c1[n1++] = poc + rps.dpoc[i];
        for (int i = 0; i < nref[1] && i < 16 && n1; i++) out->list1_poc[out->n_list1++] = c1[i % n1];
    }
    out->col_from_l0 = col_from_l0; out->mvd_l1_zero = mvd_l1_zero; out->qg_depth = pps->diff_cu_qp_delta_depth; out->sign_hiding = pps->sign_hiding;
    for (size_t q = 0; q < p.n_tus; q++) if (p.tus[q].qp_delta) out->any_qp_delta = 1;
    out->st = p.st; out->cus = p.cus; out->n_cus = p.n_cus; out->tus = p.tus; out->n_tus = p.n_tus; out->lev = p.lev; out->n_lev = p.n_lev; out->ok = ok;
    for (int i = 0; i < 16; i++) { out->ref_poc[0][i] = i < rps.n_neg ? poc + rps.dpoc[i] : 0; out->ref_poc[1][i] = i < rps.n_pos ? poc + rps.dpoc[rps.n_neg + i] : 0; }
    free(p.depth); free(p.skipf); free(p.ipm); free(p.is_intra);
    return ok ? 0 : -20;
}

ora_parsed_stream *ora_parse_stream(const uint8_t *bs, size_t n)
{
    init_scans();
    ora_parsed_stream *ps = (ora_parsed_stream *)calloc(1, sizeof(*ps));
    sps_t sps; pps_t pps; int have_sps = 0, have_pps = 0, prev_poc = 0, cap = 0;
    uint8_t *rb = (uint8_t *)malloc(n + 8);
    size_t i = 0;
    while (i + 3 < n) {
        if (!(bs[i] == 0 && bs[i + 1] == 0 && bs[i + 2] == 1)) { i++; continue; }
        size_t s = i + 3, e = s;
        while (e + 2 < n && !(bs[e] == 0 && bs[e + 1] == 0 && (bs[e + 2] == 1 || (bs[e + 2] == 0 && e + 3 < n && bs[e + 3] == 1)))) e++;
        if (e + 2 >= n) e = n;
        int nal_type = (bs[s] >> 1) & 63;
        size_t m = unescape(bs + s + 2, e - s - 2, rb);
        while (m && rb[m - 1] == 0) m--;                          /* trailing zero bytes belong to the next start code */
        bitr r = {rb, m, 0};
        if (nal_type == 33) { int k = parse_sps(&r, &sps); if (getenv("ORA_PARSE_DEBUG")) fprintf(stderr, "SPS %dx%d mincb %d ctb %d mintb %d maxtb %d depth inter %d intra %d amp %d sao %d nrps %d lt %d tmvp %d sis %d -> %d\n", sps.w, sps.h, sps.log2_min_cb, sps.log2_ctb, sps.log2_min_tb, sps.log2_max_tb, sps.tu_depth_inter, sps.tu_depth_intra, sps.amp, sps.sao, sps.n_rps, sps.long_term, sps.tmvp, sps.strong_intra, k); if (k) { ps->error = k; break; } have_sps = 1; ps->width = sps.w; ps->height = sps.h; ps->log2_ctb = sps.log2_ctb; ps->log2_min_cb = sps.log2_min_cb; ps->strong_intra = sps.strong_intra; }
        else if (nal_type == 34) { int k = parse_pps(&r, &pps); if (getenv("ORA_PARSE_DEBUG")) fprintf(stderr, "PPS sbh %d cabac_init %d refs %d %d tskip %d cuqpd %d/%d wpp %d dbk %d/%d/%d lists_mod %d par_mrg %d -> %d\n", pps.sign_hiding, pps.cabac_init_present, pps.ref_l0, pps.ref_l1, pps.tskip, pps.cu_qp_delta, pps.diff_cu_qp_delta_depth, pps.wpp, pps.deblock_ctrl, pps.deblock_override, pps.deblock_disabled, pps.lists_mod, pps.par_mrg, k); if (k) { ps->error = k; break; } have_pps = 1; }
        else if (nal_type <= 21 && have_sps && have_pps) {
            if (ps->n_pics == cap) { cap = cap ? cap * 2 : 64; ps->pics = (ora_parsed_pic *)realloc(ps->pics, (size_t)cap * sizeof(ora_parsed_pic)); }
            ora_parsed_pic *pic = &ps->pics[ps->n_pics]; memset(pic, 0, sizeof(*pic));
            int k = parse_slice(&sps, &pps, nal_type, rb, m, &prev_poc, pic);
            ps->n_pics++;
            if (k && !ps->error) ps->error = k * 1000 - (ps->n_pics - 1);
        }
        i = e;
    }
    free(rb);
    return ps;
}
void ora_parse_free(ora_parsed_stream *ps)
{
    if (!ps) return;
    for (int i = 0; i < ps->n_pics; i++) { free(ps->pics[i].cus); free(ps->pics[i].tus); free(ps->pics[i].lev); free(ps->pics[i].sao); }
    free(ps->pics); free(ps);
}
int ora_parse_error(const ora_parsed_stream *ps) { return ps->error; }
int ora_parse_num_pics(const ora_parsed_stream *ps) { return ps->n_pics; }
int ora_parse_width(const ora_parsed_stream *ps) { return ps->width; }      /* coded size (the SPS's, before the conformance window) */
int ora_parse_height(const ora_parsed_stream *ps) { return ps->height; }
const ora_pic_stats *ora_parse_pic_stats(const ora_parsed_stream *ps, int i) { return &ps->pics[i].st; }
size_t ora_parse_sizeof_stats(void) { return sizeof(ora_pic_stats); }
int ora_parse_pic_ok(const ora_parsed_stream *ps, int i) { return ps->pics[i].ok; }
/* flat copies of one picture's records for the Python tools / tests */
size_t ora_parse_num_cus(const ora_parsed_stream *ps, int i) { return ps->pics[i].n_cus; }
size_t ora_parse_num_tus(const ora_parsed_stream *ps, int i) { return ps->pics[i].n_tus; }
const ora_cu_rec *ora_parse_cus(const ora_parsed_stream *ps, int i) { return ps->pics[i].cus; }
const ora_tu_rec *ora_parse_tus(const ora_parsed_stream *ps, int i) { return ps->pics[i].tus; }
const int16_t *ora_parse_levels(const ora_parsed_stream *ps, int i) { return ps->pics[i].lev; }
size_t ora_parse_sizeof_cu(void) { return sizeof(ora_cu_rec); }
size_t ora_parse_sizeof_tu(void) { return sizeof(ora_tu_rec); }
