/*
 * ora_replay.c -- SURVEY.md 8c tier P2 on the CPU: re-create the reference ENCODER's reconstruction from its own parsed decisions
 * (ora_parse.c: CU quadtree, part modes, intra modes, merge indices / reference indices / vector differences, transform tree, levels, SAO
 * parameters, loop-filter settings), using only the oracle's leaf kernels (ora_kernels.c: intra prediction 4..32 incl. strong smoothing,
 * luma / chroma interpolation, dequantiser, IDCT 4..32 + IDST, deblocking segments, SAO apply) plus the normative derivations a decoder
 * needs (reference-sample availability, merge list, AMVP, temporal vector prediction with POC scaling, boundary strengths).  The result must
 * equal what the reference DECODER makes of the same stream, byte for byte (tests/test_replay.py): that pins those kernels against the
 * reference at every block size and in every combination ITS encoder uses -- 4x4 NxN partitions, 32x32 intra CUs, 8x8 inter CUs, bi-prediction,
 * several reference pictures per list -- not only at the sizes our own streams exercise (tier P1).  TEST INFRASTRUCTURE: nothing under ks265codec_b200/ links this.
 * Limits: 2Nx2N / 2NxN / Nx2N inter prediction blocks (no AMP, no NxN inter), one slice per
 * picture, cu_qp_delta with one quantisation group per CTB (the reference's rate-controlled streams), Log2ParMrgLevel 2, no PCM / transform skip / scaling lists / long-term pictures.
 */
#include <stdlib.h>
#include <string.h>
#include "ks_oracle.h"
#include "ora_frame.h"
#include "ora_parse.h"

int ora_me_probe(const ora_cfg *cfg, int qp, const ora_pic *src, const ora_pic *ref, int x0, int y0, int tpx, int tpy, int *mx, int *my, int *dist);   /* ora_frame.c */
int ora_tb_levels(int qp, int intra_slice, int log2, int is_luma, int intra_mode, int sign_hiding, const uint8_t *src, int ss, const uint8_t *pred, int ps, int16_t *lev);   /* ora_frame.c */

/* a transform block with coefficients, seen just before its residual is added: component, position and size in component samples, the
 * prediction, the block's QP, the intra mode (-1: inter block) and the levels the stream carries (raster n x n) */
typedef void (*tb_tap)(void *user, int ci, int x, int y, int log2, const uint8_t *pred, int pred_stride, int qp, int intra_mode, const int16_t *lev);
/* an inter CU seen right after its prediction: position, size, and per luma transform block (raster of min(size,32)-blocks) the stream's cbf */
typedef void (*cu_tap)(void *user, int x, int y, int log2, const uint8_t *pred, int pred_stride, int qp, const uint8_t *cbf_luma);
typedef struct { uint8_t *p[3]; int w, h; uint8_t *done; int dw; tb_tap tap; void *tap_user; cu_tap cutap; } rpic;       /* planes (pitch w, w/2); done: per 4x4 luma block, decoded */

/* 6.4.1 for one sample position in LUMA coordinates: inside the picture and already decoded (decoding order == z-scan order) */
static int sample_avail(const rpic *r, int x, int y) { return x >= 0 && y >= 0 && x < r->w && y < r->h && r->done[(y >> 2) * r->dw + (x >> 2)]; }

/* 8.4.4.2.2: neighbours of the n x n block at (x0, y0) of component ci (component samples), substituted; layout as ora_intra_pred expects */
static void neighbours(const rpic *r, int ci, int x0, int y0, int n, uint8_t *nb)
{
    const int sh = ci ? 1 : 0, pitch = r->w >> sh, tot = 4 * n + 1;
    uint8_t av[129]; int any = 0;
    for (int i = 0; i < tot; i++) {
        int xn, yn;
        if (i < 2 * n) { xn = x0 - 1; yn = y0 + 2 * n - 1 - i; }
        else if (i == 2 * n) { xn = x0 - 1; yn = y0 - 1; }
        else { xn = x0 + i - 2 * n - 1; yn = y0 - 1; }
        av[i] = (uint8_t)sample_avail(r, xn << sh, yn << sh);
        if (av[i]) { nb[i] = r->p[ci][(size_t)yn * pitch + xn]; any = 1; }
    }
    if (!any) { memset(nb, 128, (size_t)tot); return; }
    if (!av[0]) { int i = 1; while (!av[i]) i++; nb[0] = nb[i]; }
    for (int i = 1; i < tot; i++) if (!av[i]) nb[i] = nb[i - 1];
}

static void recon_block(rpic *r, int ci, int x0, int y0, int log2, int mode, int strong, int qp, const int16_t *lev)
{
    const int n = 1 << log2, pitch = r->w >> (ci ? 1 : 0);
    uint8_t nb[129], pred[32 * 32];
    int16_t coef[32 * 32];
    neighbours(r, ci, x0, y0, n, nb);
    ora_intra_pred(pred, n, nb, log2, mode, ci == 0, strong);
    uint8_t *dst = r->p[ci] + (size_t)y0 * pitch + x0;
    if (lev) {
        if (r->tap) r->tap(r->tap_user, ci, x0, y0, log2, pred, n, qp, mode, lev);
        ora_dequant(lev, coef, n, qp, log2);
        ora_idct_add(coef, dst, pred, n, pitch, n, log2, ci == 0 && log2 == 2);           /* 4x4 intra luma: DST */
    } else for (int y = 0; y < n; y++) memcpy(dst + (size_t)y * pitch, pred + y * n, (size_t)n);
}

/* ------------------------------------------------------------------ inter prediction ---- */
typedef struct { int16_t mv[2][2]; int ref_poc[2]; int8_t ref_idx[2]; uint8_t pf; } minfo;   /* per 4x4 luma block; pf bit X = predFlagLX, 0 = intra / not decoded */
typedef struct { int16_t mv[2][2]; int8_t ref_idx[2]; uint8_t pf; } mcand;
typedef struct { int poc, valid; uint8_t *pix; minfo *mv; } dpic;                             /* a decoded picture: output samples + motion field */
#define DPB_N 32            /* ring of decoded pictures (the reference keeps at most a GOP-of-8 anchor plus its neighbours) */
typedef struct {
    const ora_parsed_stream *ps; const ora_parsed_pic *pp;
    rpic r; minfo *mv; uint8_t *cbfy, *intra, *qpy; /* per 4x4: motion, "in a luma TB with coefficients", intra, QpY of the CU */
    dpic *dpb; int poc, no_backward;
    const uint8_t *sao_src; int sao_level; long *sao_counts;     /* ora_replay_compare_sao */
    const uint8_t *me_src; int me_method; long *me_counts;       /* ora_replay_compare_me */
    ora_pic me_cur, me_ref; int me_ref_poc, me_cur_poc, me_ready;
} rctx;

static int list_n(const rctx *c, int X) { return X ? c->pp->n_list1 : c->pp->n_list0; }
static int list_poc(const rctx *c, int X, int i) { return X ? c->pp->list1_poc[i] : c->pp->list0_poc[i]; }
static const minfo *nb_motion(const rctx *c, int x, int y)
{   /* 6.4.2: the prediction block covering (x, y) if it is decoded already and inter */
    if (!sample_avail(&c->r, x, y)) return NULL;
    const minfo *m = &c->mv[(y >> 2) * c->r.dw + (x >> 2)];
    return m->pf ? m : NULL;
}
static const dpic *find_pic(const rctx *c, int poc) { for (int i = 0; i < DPB_N; i++) if (c->dpb[i].valid && c->dpb[i].poc == poc) return &c->dpb[i]; return NULL; }
static int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static void scale_mv(int16_t *mv, int td, int tb)
{   /* 8.5.3.2.7 (8-179..8-183) */
    td = clip3i(-128, 127, td); tb = clip3i(-128, 127, tb);
    const int tx = (16384 + (abs(td) >> 1)) / td, ds = clip3i(-4096, 4095, (tb * tx + 32) >> 6);
    for (int k = 0; k < 2; k++) { const int v = ds * mv[k]; mv[k] = (int16_t)clip3i(-32768, 32767, (v < 0 ? -1 : 1) * ((abs(v) + 127) >> 8)); }
}
/* 8.5.3.2.8 / .9: temporal candidate for reference index `ref_idx` of list X */
static int temporal_mv(const rctx *c, int xp, int yp, int w, int h, int X, int ref_idx, int16_t *mv)
{
    const ora_parsed_pic *pp = c->pp;
    if (!pp->tmvp) return 0;
    const int cl = pp->st.slice_type == 0 && !pp->col_from_l0;                   /* the list the collocated picture comes from */
    if (pp->col_ref_idx >= list_n(c, cl) || ref_idx >= list_n(c, X)) return 0;
    const dpic *col = find_pic(c, list_poc(c, cl, pp->col_ref_idx));
    if (!col) return 0;
    const int l = c->ps->log2_ctb;
    for (int pass = 0; pass < 2; pass++) {
        int x, y;
        if (pass == 0) { x = xp + w; y = yp + h; if ((yp >> l) != (y >> l) || y >= c->r.h || x >= c->r.w) continue; }
        else { x = xp + (w >> 1); y = yp + (h >> 1); }
        const minfo *m = &col->mv[(((y >> 4) << 4) >> 2) * c->r.dw + (((x >> 4) << 4) >> 2)];
        if (!m->pf) continue;
        int lc;                                                                   /* which of the collocated block's vectors */
        if (!(m->pf & 1)) lc = 1;
        else if (!(m->pf & 2)) lc = 0;
        else lc = c->no_backward ? X : pp->col_from_l0;
        mv[0] = m->mv[lc][0]; mv[1] = m->mv[lc][1];
        const int cd = col->poc - m->ref_poc[lc], td = c->poc - list_poc(c, X, ref_idx);
        if (cd != td) scale_mv(mv, cd, td);
        return 1;
    }
    return 0;
}
static int same_motion(const minfo *a, const minfo *b)
{
    if (a->pf != b->pf) return 0;
    for (int X = 0; X < 2; X++) if ((a->pf >> X) & 1) if (a->mv[X][0] != b->mv[X][0] || a->mv[X][1] != b->mv[X][1] || a->ref_idx[X] != b->ref_idx[X]) return 0;
    return 1;
}
static mcand cand_of(const minfo *m) { mcand k; memcpy(k.mv, m->mv, sizeof(k.mv)); k.ref_idx[0] = m->ref_idx[0]; k.ref_idx[1] = m->ref_idx[1]; k.pf = m->pf; return k; }
/* 8.5.3.2.2 - .5: merge candidate `idx` of the prediction block (xp, yp, w, h), partition part_idx of a CU with part mode `part` (Log2ParMrgLevel 2) */
static mcand merge_candidate(const rctx *c, int xp, int yp, int w, int h, int part, int part_idx, int idx)
{
    const ora_parsed_pic *pp = c->pp;
    const int is_b = pp->st.slice_type == 0, maxc = pp->max_merge;
    mcand list[6]; int n = 0;
    const minfo *a1 = nb_motion(c, xp - 1, yp + h - 1), *b1r = nb_motion(c, xp + w - 1, yp - 1), *b0 = nb_motion(c, xp + w, yp - 1),
                *a0 = nb_motion(c, xp - 1, yp + h), *b2 = nb_motion(c, xp - 1, yp - 1);
    if (part_idx == 1 && (part == 2 || part == 6 || part == 7)) a1 = NULL;       /* Nx2N, nLx2N, nRx2N: the first partition would make the CU 2Nx2N */
    if (part_idx == 1 && (part == 1 || part == 4 || part == 5)) b1r = NULL;      /* 2NxN, 2NxnU, 2NxnD */
    const minfo *b1 = b1r;
    if (b1 && a1 && same_motion(b1, a1)) b1 = NULL;
    if (b0 && b1r && same_motion(b0, b1r)) b0 = NULL;
    if (a0 && a1 && same_motion(a0, a1)) a0 = NULL;
    if (b2 && ((a1 && same_motion(b2, a1)) || (b1r && same_motion(b2, b1r)))) b2 = NULL;
    if (b2 && (a1 != NULL) + (b1 != NULL) + (b0 != NULL) + (a0 != NULL) == 4) b2 = NULL;
    const minfo *sp[5] = {a1, b1, b0, a0, b2};
    for (int k = 0; k < 5; k++) if (sp[k]) list[n++] = cand_of(sp[k]);
    if (n < maxc) {
        mcand t; memset(&t, 0, sizeof(t));
        if (temporal_mv(c, xp, yp, w, h, 0, 0, t.mv[0])) t.pf |= 1;
        if (is_b && temporal_mv(c, xp, yp, w, h, 1, 0, t.mv[1])) t.pf |= 2;
        if (t.pf) list[n++] = t;
    }
    if (is_b && n > 1 && n < maxc) {                                             /* 8.5.3.2.4: combined bi-predictive candidates */
        static const uint8_t l0i[12] = {0, 1, 0, 2, 1, 2, 0, 3, 1, 3, 2, 3}, l1i[12] = {1, 0, 2, 0, 2, 1, 3, 0, 3, 1, 3, 2};
        const int norig = n;
        for (int k = 0; k < norig * (norig - 1) && n < maxc; k++) {
            const mcand *p0 = &list[l0i[k]], *p1 = &list[l1i[k]];
            if (!(p0->pf & 1) || !(p1->pf & 2)) continue;
            if (list_poc(c, 0, p0->ref_idx[0]) == list_poc(c, 1, p1->ref_idx[1]) && p0->mv[0][0] == p1->mv[1][0] && p0->mv[0][1] == p1->mv[1][1]) continue;
            mcand m; m.pf = 3; m.mv[0][0] = p0->mv[0][0]; m.mv[0][1] = p0->mv[0][1]; m.ref_idx[0] = p0->ref_idx[0];
            m.mv[1][0] = p1->mv[1][0]; m.mv[1][1] = p1->mv[1][1]; m.ref_idx[1] = p1->ref_idx[1];
            list[n++] = m;
        }
    }
    mcand out;
    if (idx < n) out = list[idx];
    else {                                                                       /* 8.5.3.2.5: zero candidates, the reference index counts up and then stays 0 */
        const int nref = is_b ? (pp->n_list0 < pp->n_list1 ? pp->n_list0 : pp->n_list1) : pp->n_list0, zi = idx - n;
        memset(&out, 0, sizeof(out));
        out.pf = is_b ? 3 : 1; out.ref_idx[0] = (int8_t)(zi < nref ? zi : 0); out.ref_idx[1] = is_b ? out.ref_idx[0] : -1;
    }
    if (out.pf == 3 && w + h == 12) { out.pf = 1; out.ref_idx[1] = -1; }          /* 8x4 / 4x8 blocks are never bi-predicted */
    return out;
}
/* 8.5.3.2.6 / .7: motion vector predictor of list X for reference index ref_idx */
static void amvp_predictor(const rctx *c, int xp, int yp, int w, int h, int X, int ref_idx, int mvp_idx, int16_t *pred)
{
    const int target = list_poc(c, X, ref_idx), Y = !X;
    const minfo *A[2] = {nb_motion(c, xp - 1, yp + h), nb_motion(c, xp - 1, yp + h - 1)};
    const minfo *B[3] = {nb_motion(c, xp + w, yp - 1), nb_motion(c, xp + w - 1, yp - 1), nb_motion(c, xp - 1, yp - 1)};
    const int scaled_flag = A[0] != NULL || A[1] != NULL;                         /* availability of A0 / A1 (6.4.2: decoded and not intra) */
    int have_a = 0, have_b = 0; int16_t a[2] = {0, 0}, b[2] = {0, 0};
#define SAME(m, L) (((m)->pf >> (L)) & 1 && (m)->ref_poc[L] == target)
#define TAKE(dst, m, L) { (dst)[0] = (m)->mv[L][0]; (dst)[1] = (m)->mv[L][1]; }
    for (int k = 0; k < 2 && !have_a; k++) if (A[k]) { if (SAME(A[k], X)) { TAKE(a, A[k], X) have_a = 1; } else if (SAME(A[k], Y)) { TAKE(a, A[k], Y) have_a = 1; } }
    for (int k = 0; k < 2 && !have_a; k++) if (A[k]) {
        const int L = ((A[k]->pf >> X) & 1) ? X : Y;
        TAKE(a, A[k], L) have_a = 1;
        if (A[k]->ref_poc[L] != target) scale_mv(a, c->poc - A[k]->ref_poc[L], c->poc - target);
    }
    for (int k = 0; k < 3 && !have_b; k++) if (B[k]) { if (SAME(B[k], X)) { TAKE(b, B[k], X) have_b = 1; } else if (SAME(B[k], Y)) { TAKE(b, B[k], Y) have_b = 1; } }
    if (!scaled_flag && have_b) { a[0] = b[0]; a[1] = b[1]; have_a = 1; }
    if (!scaled_flag) {
        have_b = 0;
        for (int k = 0; k < 3 && !have_b; k++) if (B[k]) {
            const int L = ((B[k]->pf >> X) & 1) ? X : Y;
            TAKE(b, B[k], L) have_b = 1;
            if (B[k]->ref_poc[L] != target) scale_mv(b, c->poc - B[k]->ref_poc[L], c->poc - target);
        }
    }
#undef SAME
#undef TAKE
    int16_t l[3][2]; int n = 0;
    if (have_a) { l[n][0] = a[0]; l[n][1] = a[1]; n++; }
    if (have_b && !(have_a && a[0] == b[0] && a[1] == b[1])) { l[n][0] = b[0]; l[n][1] = b[1]; n++; }
    if (n < 2) { int16_t t[2]; if (temporal_mv(c, xp, yp, w, h, X, ref_idx, t)) { l[n][0] = t[0]; l[n][1] = t[1]; n++; } }
    while (n < 2) { l[n][0] = 0; l[n][1] = 0; n++; }
    pred[0] = l[mvp_idx][0]; pred[1] = l[mvp_idx][1];
}
/* motion compensation of one block from a decoded picture, reference samples clamped to the picture (8.5.3.3.3); dst8 (final samples) or
 * dst16 (14-bit intermediate for bi-prediction) */
static void mc_block(const dpic *ref, int W, int H, int ci, int x0, int y0, int w, int h, int mvx, int mvy, uint8_t *dst8, int16_t *dst16, int ds)
{
    const int sh = ci ? 1 : 0, pw = W >> sh, ph = H >> sh, taps = ci ? 4 : 8, before = taps / 2 - 1, fb = ci ? 3 : 2;
    const uint8_t *plane = ref->pix + (ci == 0 ? 0 : (ci == 1 ? (size_t)W * H : (size_t)W * H + (size_t)W * H / 4));
    const int ix = x0 + (mvx >> fb), iy = y0 + (mvy >> fb), tw = w + taps - 1, th = h + taps - 1;
    uint8_t tmp[(64 + 7) * (64 + 7)];
    for (int y = 0; y < th; y++) for (int x = 0; x < tw; x++)
        tmp[y * tw + x] = plane[(size_t)clip3i(0, ph - 1, iy - before + y) * pw + clip3i(0, pw - 1, ix - before + x)];
    const int mask = (1 << fb) - 1;
    const uint8_t *org = tmp + before * tw + before;
    if (dst8) { if (ci) ora_mc_chroma(dst8, ds, org, tw, w, h, mvx & mask, mvy & mask); else ora_mc_luma(dst8, ds, org, tw, w, h, mvx & mask, mvy & mask); }
    else { if (ci) ora_mc_chroma_16(dst16, ds, org, tw, w, h, mvx & mask, mvy & mask); else ora_mc_luma_16(dst16, ds, org, tw, w, h, mvx & mask, mvy & mask); }
}
/* prediction of one block into the picture being reconstructed */
static int predict_pu(rctx *c, const mcand *m, int x, int y, int w, int h)
{
    const int W = c->r.w, H = c->r.h;
    const dpic *ref[2] = {NULL, NULL};
    for (int X = 0; X < 2; X++) if ((m->pf >> X) & 1) { if (m->ref_idx[X] < 0 || m->ref_idx[X] >= list_n(c, X)) return -7; ref[X] = find_pic(c, list_poc(c, X, m->ref_idx[X])); if (!ref[X]) return -8; }
    for (int ci = 0; ci < 3; ci++) {
        const int sh = ci ? 1 : 0, pw = W >> sh, bx = x >> sh, by = y >> sh, bw = w >> sh, bh = h >> sh;
        uint8_t *dst = c->r.p[ci] + (size_t)by * pw + bx;
        if (m->pf == 3) {
            int16_t p0[64 * 64], p1[64 * 64];
            mc_block(ref[0], W, H, ci, bx, by, bw, bh, m->mv[0][0], m->mv[0][1], NULL, p0, bw);
            mc_block(ref[1], W, H, ci, bx, by, bw, bh, m->mv[1][0], m->mv[1][1], NULL, p1, bw);
            uint8_t tmp[64 * 64];
            ora_weighted_bi(tmp, bw, p0, p1, bw, bw, bh);
            for (int r = 0; r < bh; r++) memcpy(dst + (size_t)r * pw, tmp + r * bw, (size_t)bw);
        } else {
            const int X = m->pf == 2;
            mc_block(ref[X], W, H, ci, bx, by, bw, bh, m->mv[X][0], m->mv[X][1], dst, NULL, pw);
        }
    }
    return 0;
}
/* 8.7.2.4, the motion part: 1 when the two blocks' motion differs enough for an edge to show */
static int motion_bs(const minfo *p, const minfo *q)
{
    const int np = (p->pf & 1) + (p->pf >> 1), nq = (q->pf & 1) + (q->pf >> 1);
    if (np != nq) return 1;
#define FAR(a, b) (abs((a)[0] - (b)[0]) >= 4 || abs((a)[1] - (b)[1]) >= 4)
    if (np == 1) {
        const int lp = p->pf == 2, lq = q->pf == 2;
        return p->ref_poc[lp] != q->ref_poc[lq] || FAR(p->mv[lp], q->mv[lq]);
    }
    const int p0 = p->ref_poc[0], p1 = p->ref_poc[1], q0 = q->ref_poc[0], q1 = q->ref_poc[1];
    if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;           /* different reference pictures */
    if (p0 != p1) {                                                              /* two different pictures: compare the vectors that point into the same one */
        if (p0 == q0) return FAR(p->mv[0], q->mv[0]) || FAR(p->mv[1], q->mv[1]);
        return FAR(p->mv[0], q->mv[1]) || FAR(p->mv[1], q->mv[0]);
    }
    return (FAR(p->mv[0], q->mv[0]) || FAR(p->mv[1], q->mv[1])) && (FAR(p->mv[0], q->mv[1]) || FAR(p->mv[1], q->mv[0]));
#undef FAR
}

/* one picture (I or P slice) into c->r (pre-filter reconstruction, then deblocked in place); `out` receives the SAO output */
static int replay_picture(rctx *c, uint8_t *out)
{
    const ora_parsed_stream *ps = c->ps; const ora_parsed_pic *pp = c->pp;
    if (!pp->ok) return -2;
    if (pp->any_qp_delta && pp->qg_depth != 0) return -3;                       /* cu_qp_delta with one quantisation group per CTB only (the reference's -rc 3) */
    c->no_backward = 1;                                                         /* NoBackwardPredFlag: no reference picture follows the current one in output order */
    for (int X = 0; X < 2; X++) for (int i = 0; i < list_n(c, X); i++) if (list_poc(c, X, i) > c->poc) c->no_backward = 0;
    rpic *r = &c->r;
    const int W = r->w, H = r->h, qp = pp->st.qp, dw = r->dw;
    const size_t ysz = (size_t)W * H;
    int qpc[3] = {qp, 0, 0}, qp_prev = qp, cur_ctu = -1, qg_delta = 0;
    memset(r->done, 0, (size_t)dw * ((H + 3) >> 2)); memset(c->cbfy, 0, (size_t)dw * ((H + 3) >> 2)); memset(c->intra, 0, (size_t)dw * ((H + 3) >> 2));
    memset(c->mv, 0, sizeof(minfo) * (size_t)dw * ((H + 3) >> 2));
    /* block edges on the 8x8 grid, one flag per 4-sample segment: vedge[(y/4) * ew + x/8], hedge[(y/8) * dw + x/4] */
    const int ew = (W + 7) >> 3, eh = (H + 7) >> 3;
    uint8_t *vedge = (uint8_t *)calloc((size_t)ew * ((H + 3) >> 2), 1), *hedge = (uint8_t *)calloc((size_t)dw * eh, 1);
    int rc = 0;
    for (size_t k = 0; k < pp->n_cus && !rc; k++) {
        const ora_cu_rec *cu = &pp->cus[k];
        const int S = 1 << cu->log2, half = S >> 1;
        {   /* 8.6.1 with one quantisation group per CTB: the prediction is the QpY of the last CU of the previous CTB (the slice QP at first); a
             * coded cu_qp_delta applies to its CU and to the CUs after it in the group */
            const int l = ps->log2_ctb, ctu = (cu->y >> l) * ((W + (1 << l) - 1) >> l) + (cu->x >> l);
            if (ctu != cur_ctu) { if (cur_ctu >= 0) qp_prev = qpc[0]; cur_ctu = ctu; qg_delta = 0; }
            for (uint32_t q = 0; q < cu->n_tu; q++) if (pp->tus[cu->first_tu + q].qp_delta) qg_delta = pp->tus[cu->first_tu + q].qp_delta;
            qpc[0] = (qp_prev + qg_delta + 52) % 52;
            qpc[1] = ora_chroma_qp[clip3i(0, 57, qpc[0] + pp->cb_qp_off)]; qpc[2] = ora_chroma_qp[clip3i(0, 57, qpc[0] + pp->cr_qp_off)];
            for (int y = cu->y >> 2; y < ((cu->y + S) >> 2) && y < ((H + 3) >> 2); y++) for (int x = cu->x >> 2; x < ((cu->x + S) >> 2) && x < dw; x++) c->qpy[y * dw + x] = (uint8_t)qpc[0];
        }
        if (cu->pred_mode == 0) {
            if (cu->part_mode > 2) { rc = -6; break; }                          /* 2Nx2N, 2NxN, Nx2N (the reference codes no AMP / NxN inter blocks) */
            const int npu = cu->part_mode ? 2 : 1;
            for (int k = 0; k < npu && !rc; k++) {
                const int pw_ = cu->part_mode == 2 ? half : S, ph_ = cu->part_mode == 1 ? half : S;
                const int px_ = cu->x + (cu->part_mode == 2 ? k * half : 0), py_ = cu->y + (cu->part_mode == 1 ? k * half : 0);
                mcand m;
                if (cu->merge[k]) m = merge_candidate(c, px_, py_, pw_, ph_, cu->part_mode, k, cu->merge_idx[k]);
                else {
                    memset(&m, 0, sizeof(m));
                    m.pf = pp->st.slice_type == 0 ? cu->inter_dir[k] : 1; m.ref_idx[0] = m.ref_idx[1] = -1;
                    for (int X = 0; X < 2 && !rc; X++) if ((m.pf >> X) & 1) {
                        int16_t pred[2];
                        m.ref_idx[X] = (int8_t)cu->ref_idx[k][X];
                        if (m.ref_idx[X] >= list_n(c, X)) { rc = -7; break; }
                        amvp_predictor(c, px_, py_, pw_, ph_, X, m.ref_idx[X], cu->mvp[k][X], pred);
                        m.mv[X][0] = (int16_t)(pred[0] + cu->mvd[k][X][0]); m.mv[X][1] = (int16_t)(pred[1] + cu->mvd[k][X][1]);
                    }
                }
                if (!rc) rc = predict_pu(c, &m, px_, py_, pw_, ph_);
                if (rc) break;
                for (int y = py_ >> 2; y < ((py_ + ph_) >> 2) && y < ((H + 3) >> 2); y++) for (int x = px_ >> 2; x < ((px_ + pw_) >> 2) && x < dw; x++) {
                    minfo *mi = &c->mv[y * dw + x]; memset(mi, 0, sizeof(*mi)); mi->pf = m.pf;
                    for (int X = 0; X < 2; X++) if ((m.pf >> X) & 1) { mi->mv[X][0] = m.mv[X][0]; mi->mv[X][1] = m.mv[X][1]; mi->ref_idx[X] = m.ref_idx[X]; mi->ref_poc[X] = list_poc(c, X, m.ref_idx[X]); }
                    r->done[y * dw + x] = 1;
                }
                /* the boundary between the two prediction blocks of a CU is a prediction edge */
                if (k == 1 && cu->part_mode == 2 && (px_ & 7) == 0) for (int y = py_ >> 2; y < ((py_ + ph_) >> 2); y++) vedge[y * ew + (px_ >> 3)] |= 1;
                if (k == 1 && cu->part_mode == 1 && (py_ & 7) == 0) for (int x = px_ >> 2; x < ((px_ + pw_) >> 2); x++) hedge[(py_ >> 3) * dw + x] |= 1;
            }
            if (rc) break;
            if (c->me_counts && pp->st.slice_type == 1 && cu->part_mode == 0 && cu->log2 >= 4) {
                /* OUR search, cell by cell, on the picture the reference predicted this CU from (only CUs that use the previous picture: what our
                 * encoder's single reference would be) */
                const minfo *mi = &c->mv[(cu->y >> 2) * dw + (cu->x >> 2)];
                const dpic *rp = find_pic(c, mi->ref_poc[0]);
                if (rp && mi->ref_poc[0] == c->poc - 1) {
                    if (!c->me_ready) { ora_pic_alloc(&c->me_cur, W, H); ora_pic_alloc(&c->me_ref, W, H); c->me_ready = 1; c->me_cur_poc = c->me_ref_poc = -1000000; }
                    if (c->me_cur_poc != c->poc) { ora_pic_load(&c->me_cur, c->me_src + ysz * 3 / 2 * (size_t)c->poc, W, H); c->me_cur_poc = c->poc; }
                    if (c->me_ref_poc != rp->poc) { ora_pic_load(&c->me_ref, rp->pix, W, H); c->me_ref_poc = rp->poc; }
                    ora_cfg cfg; memset(&cfg, 0, sizeof(cfg)); cfg.width = W; cfg.height = H; cfg.me_range = 64; cfg.me_iters = 16; cfg.subpel = 2; cfg.me_method = c->me_method;
                    for (int cy = cu->y; cy < cu->y + S && cy + 16 <= H; cy += 16) for (int cx = cu->x; cx < cu->x + S && cx + 16 <= W; cx += 16) {
                        const minfo *col = &rp->mv[(cy >> 2) * dw + (cx >> 2)];                 /* the co-located vector of the previous picture seeds the search */
                        int mx, my, dist;
                        int tpx = 0, tpy = 0;
                        if (col->pf & 1) { const int d = rp->poc - col->ref_poc[0]; tpx = d > 0 ? col->mv[0][0] / d : 0; tpy = d > 0 ? col->mv[0][1] / d : 0; }   /* per picture of distance */
                        ora_me_probe(&cfg, qp, &c->me_cur, &c->me_ref, cx, cy, tpx, tpy, &mx, &my, &dist);
                        uint8_t pr[256];
                        ora_mc_luma(pr, 16, c->me_ref.c[0].p + (size_t)cy * c->me_ref.c[0].stride + cx, c->me_ref.c[0].stride, 16, 16, mi->mv[0][0], mi->mv[0][1]);
                        const int rd = (int)ora_sad(c->me_cur.c[0].p + (size_t)cy * c->me_cur.c[0].stride + cx, pr, c->me_cur.c[0].stride, 16, 16, 16);
                        long *k = c->me_counts;
                        k[0]++; k[1] += mx == mi->mv[0][0] && my == mi->mv[0][1]; k[2] += abs(mx - mi->mv[0][0]) <= 1 && abs(my - mi->mv[0][1]) <= 1;
                        k[3] += dist <= rd; k[4] += dist; k[5] += rd; k[6] += dist < rd; k[7] += dist > rd;
                        {   /* the same by the size of the reference's vector (whole samples): < 2, < 8, < 16, < 32, >= 32 -> cells, our SAD sum, its SAD sum */
                            const int mag = (abs(mi->mv[0][0]) > abs(mi->mv[0][1]) ? abs(mi->mv[0][0]) : abs(mi->mv[0][1])) >> 2;
                            const int b = mag < 2 ? 0 : (mag < 8 ? 1 : (mag < 16 ? 2 : (mag < 32 ? 3 : 4)));
                            k[8 + 3 * b]++; k[9 + 3 * b] += dist; k[10 + 3 * b] += rd;
                        }
                    }
                }
            }
            if (r->cutap) {
                const int tl = cu->log2 > 5 ? 5 : cu->log2, nt = S >> tl; uint8_t cbf[4] = {0, 0, 0, 0};
                for (uint32_t q = 0; q < cu->n_tu; q++) { const ora_tu_rec *t = &pp->tus[cu->first_tu + q]; if (t->log2 == tl && (t->cbf & 1)) cbf[((t->y - cu->y) >> tl) * nt + ((t->x - cu->x) >> tl)] = 1; else if (t->log2 != tl) cbf[0] |= 2; }
                if (!(cbf[0] & 2)) r->cutap(r->tap_user, cu->x, cu->y, cu->log2, r->p[0] + (size_t)cu->y * W + cu->x, W, qpc[0], cbf);
            }
        } else
            for (int y = cu->y >> 2; y < ((cu->y + S) >> 2) && y < ((H + 3) >> 2); y++) for (int x = cu->x >> 2; x < ((cu->x + S) >> 2) && x < dw; x++) c->intra[y * dw + x] = 1;
        /* the CU boundary is a prediction-block edge AND a transform-block edge whatever its transform tree looks like (8.7.2.3 starts from the coding block) */
        if ((cu->x & 7) == 0 && cu->x > 0) for (int y = cu->y >> 2; y < ((cu->y + S) >> 2) && y < ((H + 3) >> 2); y++) vedge[y * ew + (cu->x >> 3)] = 3;
        if ((cu->y & 7) == 0 && cu->y > 0) for (int x = cu->x >> 2; x < ((cu->x + S) >> 2) && x < dw; x++) hedge[(cu->y >> 3) * dw + x] = 3;
        for (uint32_t q = 0; q < cu->n_tu; q++) {
            const ora_tu_rec *t = &pp->tus[cu->first_tu + q];
            const int n = 1 << t->log2;
            if (cu->pred_mode == 1) {
                const int part = cu->part_mode == 3 ? ((t->x - cu->x >= half) ? 1 : 0) | ((t->y - cu->y >= half) ? 2 : 0) : 0;
                recon_block(r, 0, t->x, t->y, t->log2, cu->intra_mode[part], ps->strong_intra, qpc[0], (t->cbf & 1) ? pp->lev + t->lev_off[0] : NULL);
            } else if (t->cbf & 1) {
                int16_t coef[32 * 32];
                uint8_t *dst = r->p[0] + (size_t)t->y * W + t->x;
                if (r->tap) r->tap(r->tap_user, 0, t->x, t->y, t->log2, dst, W, qpc[0], -1, pp->lev + t->lev_off[0]);
                ora_dequant(pp->lev + t->lev_off[0], coef, n, qpc[0], t->log2);
                ora_idct_add(coef, dst, dst, n, W, W, t->log2, 0);
            }
            for (int y = t->y >> 2; y < ((t->y + n) >> 2); y++) for (int x = t->x >> 2; x < ((t->x + n) >> 2); x++) { r->done[y * dw + x] = 1; if (t->cbf & 1) c->cbfy[y * dw + x] = 1; }
            if ((t->x & 7) == 0 && t->x > 0) for (int y = t->y >> 2; y < ((t->y + n) >> 2); y++) vedge[y * ew + (t->x >> 3)] |= 2;
            if ((t->y & 7) == 0 && t->y > 0) for (int x = t->x >> 2; x < ((t->x + n) >> 2); x++) hedge[(t->y >> 3) * dw + x] |= 2;
            /* chroma: with the luma block, or -- 4x4 luma blocks -- one 4x4 block per component after the fourth luma block of the 8x8 parent */
            int xc, yc, l2c;
            if (t->log2 > 2) { xc = t->x >> 1; yc = t->y >> 1; l2c = t->log2 - 1; }
            else if ((t->x & 4) && (t->y & 4)) { xc = (t->x - 4) >> 1; yc = (t->y - 4) >> 1; l2c = 2; }
            else continue;
            for (int ci = 1; ci < 3; ci++) {
                const int16_t *lev = (t->cbf & (1 << ci)) ? pp->lev + t->lev_off[ci] : NULL;
                if (cu->pred_mode == 1) recon_block(r, ci, xc, yc, l2c, cu->chroma_mode, 0, qpc[ci], lev);
                else if (lev) {
                    int16_t coef[32 * 32];
                    uint8_t *dst = r->p[ci] + (size_t)yc * (W >> 1) + xc;
                    if (r->tap) r->tap(r->tap_user, ci, xc, yc, l2c, dst, W >> 1, qpc[ci], -1, lev);
                    ora_dequant(lev, coef, 1 << l2c, qpc[ci], l2c);
                    ora_idct_add(coef, dst, dst, 1 << l2c, W >> 1, W >> 1, l2c, 0);
                }
            }
        }
        if (cu->pred_mode == 1 && cu->n_tu == 0) { rc = -9; break; }
    }
    /* 8.7.2: all vertical edges of the picture, then all horizontal ones */
    if (!rc && !pp->dbk_disabled) {
        for (int dir = 0; dir < 2; dir++)
            for (int e = 8; e < (dir ? H : W); e += 8)
                for (int s = 0; s < (dir ? W : H); s += 4) {
                    const int fl = dir ? hedge[(e >> 3) * dw + (s >> 2)] : vedge[(s >> 2) * ew + (e >> 3)];
                    if (!fl) continue;
                    const int xq = dir ? s : e, yq = dir ? e : s, xp = dir ? s : e - 1, yp = dir ? e - 1 : s;
                    const int iq = (yq >> 2) * dw + (xq >> 2), ip = (yp >> 2) * dw + (xp >> 2);
                    int bs = 0;
                    if (c->intra[iq] || c->intra[ip]) bs = 2;                   /* 8.7.2.4 */
                    else if ((fl & 2) && (c->cbfy[iq] || c->cbfy[ip])) bs = 1;
                    else bs = motion_bs(&c->mv[ip], &c->mv[iq]);
                    if (!bs) continue;
                    const int ql = (c->qpy[ip] + c->qpy[iq] + 1) >> 1;           /* 8.7.2.5.3: the edge's QP is the mean of the two CUs' */
                    const int beta = ora_beta_table[clip3i(0, 51, ql + 2 * pp->beta_off_div2)];
                    const int tc = ora_tc_table[clip3i(0, 53, ql + 2 * (bs - 1) + 2 * pp->tc_off_div2)];
                    ora_deblock_luma_seg(r->p[0] + (size_t)yq * W + xq, dir ? W : 1, dir ? 1 : W, beta, tc);
                    if (bs == 2 && !(e & 8)) for (int ci = 1; ci < 3; ci++) {    /* 8.7.2.5.5: QpC from that mean + the chroma offset */
                        const int tcc = ora_tc_table[clip3i(0, 53, ora_chroma_qp[clip3i(0, 57, ql + (ci == 1 ? pp->cb_qp_off : pp->cr_qp_off))] + 2 + 2 * pp->tc_off_div2)];
                        ora_deblock_chroma_seg(r->p[ci] + (size_t)(yq >> 1) * (W >> 1) + (xq >> 1), dir ? (W >> 1) : 1, dir ? 1 : (W >> 1), tcc, 2);
                    }
                }
    }
    if (!rc && c->sao_counts) {         /* OUR SAO decision on the reference's deblocked picture, against the parameters it coded */
        ora_cfg cfg; memset(&cfg, 0, sizeof(cfg)); cfg.width = W; cfg.height = H; cfg.sao = c->sao_level;
        ora_pic src, deb, tmp;
        const int l = ps->log2_ctb, ctw = (W + (1 << l) - 1) >> l, cth = (H + (1 << l) - 1) >> l;
        ks_ctu_syn *ctus = (ks_ctu_syn *)calloc((size_t)ctw * cth, sizeof(ks_ctu_syn));
        if (l == 6 && ctus && !ora_pic_alloc(&src, W, H) && !ora_pic_alloc(&deb, W, H) && !ora_pic_alloc(&tmp, W, H)) {
            ora_pic_load(&src, c->sao_src + ysz * 3 / 2 * (size_t)c->poc, W, H); ora_pic_load(&deb, r->p[0], W, H);
            ora_sao_picture(&cfg, qp, &src, &deb, &tmp, ctus);
            for (int i = 0; i < ctw * cth; i++) for (int grp = 0; grp < 2; grp++) {
                const int ci = grp ? 1 : 0, rt = pp->sao[i].type[ci], ot = ctus[i].sao[ci].type;
                long *k = c->sao_counts + grp * 8;
                k[0]++; k[1] += rt != 0; k[2] += ot != 0; k[3] += rt == ot;
                if (rt && rt == ot) {
                    int same_pos = pp->sao[i].pos[ci] == ctus[i].sao[ci].band_or_class && (!grp || rt == 2 || pp->sao[i].pos[2] == ctus[i].sao[2].band_or_class);
                    int same_off = same_pos && !memcmp(pp->sao[i].off[ci], ctus[i].sao[ci].off, 4) && (!grp || !memcmp(pp->sao[i].off[2], ctus[i].sao[2].off, 4));
                    k[4]++; k[5] += same_pos; k[6] += same_off;
                }
            }
            ora_pic_free(&src); ora_pic_free(&deb); ora_pic_free(&tmp);
        }
        free(ctus);
    }
    /* 8.7.3: SAO reads the deblocked picture and writes the output picture */
    if (!rc) {
        memcpy(out, r->p[0], ysz * 3 / 2);
        const int l = ps->log2_ctb, ctw = (W + (1 << l) - 1) >> l, cth = (H + (1 << l) - 1) >> l;
        for (int ry = 0; ry < cth; ry++) for (int rx = 0; rx < ctw; rx++) {
            const ora_sao_rec *sr = &pp->sao[ry * ctw + rx];
            for (int ci = 0; ci < 3; ci++) {
                if (!sr->type[ci]) continue;
                const int sh = ci ? 1 : 0, pw = W >> sh, ph = H >> sh, x0 = (rx << l) >> sh, y0 = (ry << l) >> sh, cs = (1 << l) >> sh;
                const int w = x0 + cs > pw ? pw - x0 : cs, h = y0 + cs > ph ? ph - y0 : cs;
                const size_t off = ci == 0 ? 0 : (ci == 1 ? ysz : ysz + ysz / 4);
                ora_sao_apply_ctb(out + off, pw, r->p[0] + off, pw, x0, y0, w, h, pw, ph, sr->type[ci], sr->pos[ci], sr->off[ci]);
            }
        }
    }
    free(vedge); free(hedge);
    return rc;
}

/* Replays pictures first .. first + count - 1 of the stream (decoding order; I and P slices, P-only streams come out in display order) into
 * `out` (count coded-size I420 pictures).  Returns 0, or a negative code at the first picture outside the limits in the header. */
typedef struct { tb_tap fn; void *user; int *cur_pic; cu_tap cufn; const uint8_t *sao_src; int sao_level; long *sao_counts; const uint8_t *me_src; int me_method; long *me_counts; } tap_cfg;
static int replay_run(const ora_parsed_stream *ps, int first, int count, uint8_t *out, const tap_cfg *tap)
{
    if (!ps || first < 0 || count < 1 || first + count > ps->n_pics) return -1;
    const int W = ps->width, H = ps->height, dw = (W + 3) >> 2, nb = dw * ((H + 3) >> 2);
    const size_t ysz = (size_t)W * H, fsz = ysz * 3 / 2;
    rctx c; memset(&c, 0, sizeof(c));
    c.ps = ps; c.r.w = W; c.r.h = H; c.r.dw = dw;
    uint8_t *pre = (uint8_t *)calloc(fsz, 1);
    c.r.p[0] = pre; c.r.p[1] = pre + ysz; c.r.p[2] = pre + ysz + ysz / 4;
    c.r.done = (uint8_t *)calloc((size_t)nb, 1); c.cbfy = (uint8_t *)calloc((size_t)nb, 1); c.intra = (uint8_t *)calloc((size_t)nb, 1); c.qpy = (uint8_t *)calloc((size_t)nb, 1);
    c.mv = (minfo *)calloc((size_t)nb, sizeof(minfo));
    dpic dpb[DPB_N]; memset(dpb, 0, sizeof(dpb)); c.dpb = dpb;
    int rc = 0, slot = 0;
    /* pictures before `first` that the requested ones may reference are replayed too (from the closest IDR back) */
    int start = first; while (start > 0 && ps->pics[start].st.nal_type != 19 && ps->pics[start].st.nal_type != 20) start--;
    uint8_t *scratch = (uint8_t *)malloc(fsz);
    for (int i = start; i < first + count && !rc; i++) {
        c.pp = &ps->pics[i]; c.poc = c.pp->st.poc;
        c.r.tap = tap && i >= first ? tap->fn : NULL; c.r.tap_user = tap ? tap->user : NULL; c.r.cutap = tap && i >= first ? tap->cufn : NULL;
        if (tap && tap->cur_pic) *tap->cur_pic = i;
        c.me_counts = tap && i >= first ? tap->me_counts : NULL; c.me_src = tap ? tap->me_src : NULL; c.me_method = tap ? tap->me_method : 0;
        c.sao_counts = tap && i >= first ? tap->sao_counts : NULL; c.sao_src = tap ? tap->sao_src : NULL; c.sao_level = tap ? tap->sao_level : 0;
        if (c.pp->st.nal_type == 19 || c.pp->st.nal_type == 20) for (int k = 0; k < DPB_N; k++) dpb[k].valid = 0;
        uint8_t *dst = i >= first ? out + fsz * (size_t)(i - first) : scratch;
        rc = replay_picture(&c, dst);
        if (rc) break;
        dpic *d = &dpb[slot]; slot = (slot + 1) % DPB_N;
        if (!d->pix) { d->pix = (uint8_t *)malloc(fsz); d->mv = (minfo *)malloc(sizeof(minfo) * (size_t)nb); }
        memcpy(d->pix, dst, fsz); memcpy(d->mv, c.mv, sizeof(minfo) * (size_t)nb); d->poc = c.poc; d->valid = 1;
    }
    for (int k = 0; k < DPB_N; k++) { free(dpb[k].pix); free(dpb[k].mv); }
    if (c.me_ready) { ora_pic_free(&c.me_cur); ora_pic_free(&c.me_ref); }
    free(scratch); free(pre); free(c.r.done); free(c.cbfy); free(c.intra); free(c.qpy); free(c.mv);
    return rc;
}

int ora_replay_pictures(const ora_parsed_stream *ps, int first, int count, uint8_t *out) { return replay_run(ps, first, count, out, NULL); }

/* ---- the reference's LEVELS against our transform-block coder ----
 * For every transform block with coefficients of the reference's stream: residual = source - the reference's own prediction (re-created above),
 * through OUR forward transform + quantiser + sign-data hiding (ora_tb_levels), compared with the levels the reference coded.  `src` holds the
 * source pictures in DISPLAY order (coded size, I420), valid for one IDR period.  counts[cat][0..3] = blocks compared, blocks with identical
 * levels, coefficients that are non-zero on either side, coefficients that differ; cat = 0 I-slice luma, 1 I-slice chroma, 2 inter-slice intra
 * blocks, 3 inter blocks luma, 4 inter blocks chroma.  Tells how much of the reference's quantiser (rounding offsets, sign hiding, any RD
 * quantisation) our restatement reproduces; the reference's decisions to drop whole blocks do not enter (those blocks carry no levels). */
typedef struct { const ora_parsed_stream *ps; const uint8_t *src; int cur; long (*counts)[4]; } lev_cmp;
static void compare_tap(void *user, int ci, int x, int y, int log2, const uint8_t *pred, int ps_, int qp, int intra_mode, const int16_t *lev)
{
    lev_cmp *u = (lev_cmp *)user;
    const ora_parsed_pic *pp = &u->ps->pics[u->cur];
    const int W = u->ps->width, H = u->ps->height, n = 1 << log2, islice = pp->st.slice_type == 2;
    const size_t fsz = (size_t)W * H * 3 / 2;
    const uint8_t *plane = u->src + fsz * (size_t)pp->st.poc + (ci == 0 ? 0 : (ci == 1 ? (size_t)W * H : (size_t)W * H * 5 / 4));
    const int pitch = ci ? W >> 1 : W;
    int16_t ours[1024];
    ora_tb_levels(qp, islice, log2, ci == 0, intra_mode, pp->sign_hiding, plane + (size_t)y * pitch + x, pitch, pred, ps_, ours);
    const int cat = islice ? (ci ? 1 : 0) : (intra_mode >= 0 ? 2 : (ci ? 4 : 3));
    long nz = 0, diff = 0;
    for (int i = 0; i < n * n; i++) { if (ours[i] || lev[i]) nz++; if (ours[i] != lev[i]) diff++; }
    u->counts[cat][0]++; u->counts[cat][1] += diff == 0; u->counts[cat][2] += nz; u->counts[cat][3] += diff;
}
int ora_replay_compare_levels(const ora_parsed_stream *ps, int first, int count, const uint8_t *src, long counts[5][4])
{
    if (!ps || !src || !counts) return -1;
    const size_t fsz = (size_t)ps->width * ps->height * 3 / 2;
    uint8_t *out = (uint8_t *)malloc(fsz * (size_t)(count > 0 ? count : 1));
    lev_cmp u; u.ps = ps; u.src = src; u.cur = 0; u.counts = counts;
    memset(counts, 0, sizeof(long) * 20);
    tap_cfg t; memset(&t, 0, sizeof(t)); t.fn = compare_tap; t.user = &u; t.cur_pic = &u.cur;
    const int rc = replay_run(ps, first, count, out, &t);
    free(out);
    return rc;
}

/* ---- the reference's zero-block decisions against ours ----
 * For every luma transform block of every 2Nx2N-transform inter CU of the P pictures: does the reference code levels (its cbf), would the plain
 * quantiser produce any, and does OUR coder keep them after its RD zero-out (lambda of QP + lambda_delta on the pictures with poc % 4 != 0,
 * as ks_rc_lambda_qp does for -bframes 0 streams)?  counts[0..5] = blocks, reference codes, plain quantiser non-zero, ours codes, both code,
 * neither codes. */
typedef struct { const ora_parsed_stream *ps; const uint8_t *src; int cur, lambda_delta; long *counts; } zb_cmp;
int ora_tb_decision(int qp, int lambda_qp, int log2, int sign_hiding, const uint8_t *src, int ss, const uint8_t *pred, int ps, int *plain_nnz);   /* ora_frame.c */
static void zero_tap(void *user, int x, int y, int log2, const uint8_t *pred, int ps_, int qp, const uint8_t *cbf)
{
    zb_cmp *u = (zb_cmp *)user;
    const ora_parsed_pic *pp = &u->ps->pics[u->cur];
    if (pp->st.slice_type != 1) return;
    const int W = u->ps->width, H = u->ps->height, tl = log2 > 5 ? 5 : log2, nt = (1 << log2) >> tl, n = 1 << tl;
    const uint8_t *plane = u->src + (size_t)W * H * 3 / 2 * (size_t)pp->st.poc;
    const int lq = qp + ((pp->st.poc & 3) ? (u->lambda_delta & 255) : (u->lambda_delta >> 8));     /* low byte: non-key pictures, next byte: key pictures */
    for (int j = 0; j < nt; j++) for (int i = 0; i < nt; i++) {
        int plain = 0;
        const int ours = ora_tb_decision(qp, lq, tl, pp->sign_hiding, plane + (size_t)(y + j * n) * W + x + i * n, W, pred + (size_t)(j * n) * ps_ + i * n, ps_, &plain);
        const int ref = cbf[j * nt + i] & 1;
        u->counts[0]++; u->counts[1] += ref; u->counts[2] += plain != 0; u->counts[3] += ours; u->counts[4] += ref && ours; u->counts[5] += !ref && !ours;
    }
}
int ora_replay_compare_zero_blocks(const ora_parsed_stream *ps, int first, int count, const uint8_t *src, int lambda_delta, long counts[6])
{
    if (!ps || !src || !counts) return -1;
    const size_t fsz = (size_t)ps->width * ps->height * 3 / 2;
    uint8_t *out = (uint8_t *)malloc(fsz * (size_t)(count > 0 ? count : 1));
    zb_cmp u; u.ps = ps; u.src = src; u.cur = 0; u.lambda_delta = lambda_delta; u.counts = counts;
    memset(counts, 0, sizeof(long) * 6);
    tap_cfg t; memset(&t, 0, sizeof(t)); t.user = &u; t.cur_pic = &u.cur; t.cufn = zero_tap;
    const int rc = replay_run(ps, first, count, out, &t);
    free(out);
    return rc;
}

/* ---- the reference's SAO parameters against our decision (row a18) ----
 * OUR SAO decision (ora_sao_picture at `sao_level`, the model the device kernel equals) runs on the reference's own deblocked pictures and
 * source; per CTU and component group (0 luma, 1 chroma) counts[grp * 8 + 0..6] = CTUs, reference has SAO on, ours has it on, same type,
 * both on with the same type, of those the same band position / edge class, of those identical offsets.  `src` in display order (one IDR period). */
int ora_replay_compare_sao(const ora_parsed_stream *ps, int first, int count, const uint8_t *src, int sao_level, long counts[16])
{
    if (!ps || !src || !counts) return -1;
    const size_t fsz = (size_t)ps->width * ps->height * 3 / 2;
    uint8_t *out = (uint8_t *)malloc(fsz * (size_t)(count > 0 ? count : 1));
    memset(counts, 0, sizeof(long) * 16);
    tap_cfg t; memset(&t, 0, sizeof(t)); t.sao_src = src; t.sao_level = sao_level; t.sao_counts = counts;
    const int rc = replay_run(ps, first, count, out, &t);
    free(out);
    return rc;
}

/* ---- the reference's vectors against our search (rows a3-a6) ----
 * For every 16x16 cell of the 2Nx2N inter CUs (>= 16x16, P pictures, predicted from the previous picture) of a -bframes 0 stream: OUR search
 * (small diamond / hexagon + half + quarter refinement, SAD) on the reference's own reference picture, seeded like our encoder seeds it (the
 * co-located vector of the previous picture).  counts[0..7] = cells, identical vector, within one quarter sample, our SAD <= the SAD at the
 * reference's vector, sum of our SADs, sum of the reference's, ours strictly lower, ours strictly higher; counts[8 + 3 b ..] = cells, our SAD sum, the reference's SAD sum for the cells whose
 * reference vector is < 2, < 8, < 16, < 32, >= 32 whole samples long (b = 0..4). */
int ora_me_probe(const ora_cfg *cfg, int qp, const ora_pic *src, const ora_pic *ref, int x0, int y0, int tpx, int tpy, int *mx, int *my, int *dist);   /* ora_frame.c */
int ora_replay_compare_me(const ora_parsed_stream *ps, int first, int count, const uint8_t *src, int me_method, long counts[23])
{
    if (!ps || !src || !counts) return -1;
    const size_t fsz = (size_t)ps->width * ps->height * 3 / 2;
    uint8_t *out = (uint8_t *)malloc(fsz * (size_t)(count > 0 ? count : 1));
    memset(counts, 0, sizeof(long) * 23);
    tap_cfg t; memset(&t, 0, sizeof(t)); t.me_src = src; t.me_method = me_method; t.me_counts = counts;
    const int rc = replay_run(ps, first, count, out, &t);
    free(out);
    return rc;
}

/* the intra-only entry point of the first version: one I picture */
int ora_replay_intra_picture(const ora_parsed_stream *ps, int pic, uint8_t *out)
{
    if (!ps || pic < 0 || pic >= ps->n_pics) return -1;
    if (ps->pics[pic].st.slice_type != 2) return -2;
    return ora_replay_pictures(ps, pic, 1, out);
}
