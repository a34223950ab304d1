/*
 * ora_replay.c -- SURVEY.md 8c tier P2 on the CPU: re-create the reference ENCODER's reconstruction from its own parsed decisions
 * (ora_parse.c: CU quadtree, part modes, intra modes, merge indices / reference indices / vector differences, transform tree, levels, SAO
 * parameters, loop-filter settings), using only the oracle's leaf kernels (ora_kernels.c: intra prediction 4..32 incl. strong smoothing,
 * luma / chroma interpolation, dequantiser, IDCT 4..32 + IDST, deblocking segments, SAO apply) plus the normative derivations a decoder
 * needs (reference-sample availability, merge list, AMVP, temporal vector prediction with POC scaling, boundary strengths).  The result must
 * equal what the reference DECODER makes of the same stream, byte for byte (tests/test_replay.py): that pins those kernels against the
 * reference at every block size and in every combination ITS encoder uses -- 4x4 NxN partitions, 32x32 intra CUs, 8x8 inter CUs, several
 * reference pictures -- not only at the sizes our own streams exercise (tier P1).  TEST INFRASTRUCTURE: nothing under ks265codec_b200/ links this.
 * Limits: I and P slices (the reference's -bframes 0 streams), 2Nx2N inter prediction blocks (presets ultrafast .. slow), one slice per
 * picture, cu_qp_delta all zero (the reference at -rc 0), Log2ParMrgLevel 2, no PCM / transform skip / scaling lists / long-term pictures.
 */
#include <stdlib.h>
#include <string.h>
#include "ks_oracle.h"
#include "ora_parse.h"

typedef struct { uint8_t *p[3]; int w, h; uint8_t *done; int dw; } rpic;       /* planes (pitch w, w/2); done: per 4x4 luma block, decoded */

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
        ora_dequant(lev, coef, n, qp, log2);
        ora_idct_add(coef, dst, pred, n, pitch, n, log2, ci == 0 && log2 == 2);           /* 4x4 intra luma: DST */
    } else for (int y = 0; y < n; y++) memcpy(dst + (size_t)y * pitch, pred + y * n, (size_t)n);
}

/* ------------------------------------------------------------------ inter prediction (P slices) ---- */
typedef struct { int16_t mvx, mvy; int ref_poc; int8_t ref_idx; uint8_t inter; } minfo;      /* per 4x4 luma block */
typedef struct { int poc, valid; uint8_t *pix; minfo *mv; } dpic;                             /* a decoded picture: output samples + motion field */
#define DPB_N 32            /* ring of decoded pictures (the reference keeps at most a GOP-of-8 anchor plus its neighbours) */
typedef struct {
    const ora_parsed_stream *ps; const ora_parsed_pic *pp;
    rpic r; minfo *mv; uint8_t *cbfy, *intra;     /* per 4x4: motion, "in a luma TB with coefficients", intra */
    dpic *dpb; int poc;
} rctx;

static const minfo *nb_motion(const rctx *c, int x, int y)
{   /* 6.4.2: the prediction block covering (x, y) if it is decoded already and inter */
    if (!sample_avail(&c->r, x, y)) return NULL;
    const minfo *m = &c->mv[(y >> 2) * c->r.dw + (x >> 2)];
    return m->inter ? m : NULL;
}
static const dpic *find_pic(const rctx *c, int poc) { for (int i = 0; i < DPB_N; i++) if (c->dpb[i].valid && c->dpb[i].poc == poc) return &c->dpb[i]; return NULL; }
static int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static void scale_mv(int16_t *mx, int16_t *my, int td, int tb)
{   /* 8.5.3.2.7 (8-179..8-183) */
    td = clip3i(-128, 127, td); tb = clip3i(-128, 127, tb);
    const int tx = (16384 + (abs(td) >> 1)) / td, ds = clip3i(-4096, 4095, (tb * tx + 32) >> 6);
    const int vx = ds * *mx, vy = ds * *my;
    *mx = (int16_t)clip3i(-32768, 32767, (vx < 0 ? -1 : 1) * ((abs(vx) + 127) >> 8));
    *my = (int16_t)clip3i(-32768, 32767, (vy < 0 ? -1 : 1) * ((abs(vy) + 127) >> 8));
}
/* 8.5.3.2.8 / .9: temporal candidate for reference index `ref_idx` of list 0 */
static int temporal_mv(const rctx *c, int xp, int yp, int w, int h, int ref_idx, int16_t *mx, int16_t *my)
{
    const ora_parsed_pic *pp = c->pp;
    if (!pp->tmvp || pp->col_ref_idx >= pp->n_list0) return 0;
    const dpic *col = find_pic(c, pp->list0_poc[pp->col_ref_idx]);
    if (!col) return 0;
    const int l = c->ps->log2_ctb;
    for (int pass = 0; pass < 2; pass++) {
        int x, y;
        if (pass == 0) { x = xp + w; y = yp + h; if ((yp >> l) != (y >> l) || y >= c->r.h || x >= c->r.w) continue; }
        else { x = xp + (w >> 1); y = yp + (h >> 1); }
        const minfo *m = &col->mv[(((y >> 4) << 4) >> 2) * c->r.dw + (((x >> 4) << 4) >> 2)];
        if (!m->inter) continue;
        *mx = m->mvx; *my = m->mvy;
        const int cd = col->poc - m->ref_poc, td = c->poc - pp->list0_poc[ref_idx];
        if (cd != td) scale_mv(mx, my, cd, td);
        return 1;
    }
    return 0;
}
typedef struct { int16_t mvx, mvy; int ref_idx; } mcand;
static int same_motion(const minfo *a, const minfo *b) { return a->mvx == b->mvx && a->mvy == b->mvy && a->ref_idx == b->ref_idx; }
/* 8.5.3.2.2 - .5 for a 2Nx2N prediction block of a P slice (Log2ParMrgLevel 2: no merge estimation regions) */
static mcand merge_candidate(const rctx *c, int xp, int yp, int w, int h, int idx)
{
    mcand list[6]; int n = 0;
    const minfo *a1 = nb_motion(c, xp - 1, yp + h - 1), *b1 = nb_motion(c, xp + w - 1, yp - 1), *b0 = nb_motion(c, xp + w, yp - 1),
                *a0 = nb_motion(c, xp - 1, yp + h), *b2 = nb_motion(c, xp - 1, yp - 1);
    if (b1 && a1 && same_motion(b1, a1)) b1 = NULL;
    if (b0 && nb_motion(c, xp + w - 1, yp - 1) && same_motion(b0, nb_motion(c, xp + w - 1, yp - 1))) b0 = NULL;
    if (a0 && a1 && same_motion(a0, a1)) a0 = NULL;
    if (b2 && ((a1 && same_motion(b2, a1)) || (nb_motion(c, xp + w - 1, yp - 1) && same_motion(b2, nb_motion(c, xp + w - 1, yp - 1))))) b2 = NULL;
    if (b2 && (a1 != NULL) + (b1 != NULL) + (b0 != NULL) + (a0 != NULL) == 4) b2 = NULL;
    const minfo *sp[5] = {a1, b1, b0, a0, b2};
    for (int k = 0; k < 5; k++) if (sp[k]) { list[n].mvx = sp[k]->mvx; list[n].mvy = sp[k]->mvy; list[n].ref_idx = sp[k]->ref_idx; n++; }
    if (n < c->pp->max_merge) { int16_t mx, my; if (temporal_mv(c, xp, yp, w, h, 0, &mx, &my)) { list[n].mvx = mx; list[n].mvy = my; list[n].ref_idx = 0; n++; } }
    mcand z = {0, 0, 0};
    if (idx < n && idx < c->pp->max_merge) return list[idx];
    const int zi = idx - (n < c->pp->max_merge ? n : c->pp->max_merge);       /* zero candidates: reference index counts up, then stays 0 */
    z.ref_idx = zi < c->pp->n_list0 ? zi : 0;
    return z;
}
/* 8.5.3.2.6 / .7: motion vector predictor of list 0 for reference index ref_idx */
static void amvp_predictor(const rctx *c, int xp, int yp, int w, int h, int ref_idx, int mvp_idx, int16_t *px, int16_t *py)
{
    const int target = c->pp->list0_poc[ref_idx];
    const minfo *A[2] = {nb_motion(c, xp - 1, yp + h), nb_motion(c, xp - 1, yp + h - 1)};
    const minfo *B[3] = {nb_motion(c, xp + w, yp - 1), nb_motion(c, xp + w - 1, yp - 1), nb_motion(c, xp - 1, yp - 1)};
    /* isScaledFlag looks at the AVAILABILITY of A0 / A1 (6.4.2: decoded and not intra) */
    const int scaled_flag = A[0] != NULL || A[1] != NULL;
    int have_a = 0, have_b = 0; int16_t ax = 0, ay = 0, bx = 0, by = 0;
    for (int k = 0; k < 2 && !have_a; k++) if (A[k] && A[k]->ref_poc == target) { ax = A[k]->mvx; ay = A[k]->mvy; have_a = 1; }
    for (int k = 0; k < 2 && !have_a; k++) if (A[k]) { ax = A[k]->mvx; ay = A[k]->mvy; have_a = 1; if (A[k]->ref_poc != target) scale_mv(&ax, &ay, c->poc - A[k]->ref_poc, c->poc - target); }
    for (int k = 0; k < 3 && !have_b; k++) if (B[k] && B[k]->ref_poc == target) { bx = B[k]->mvx; by = B[k]->mvy; have_b = 1; }
    if (!scaled_flag && have_b) { ax = bx; ay = by; have_a = 1; }
    if (!scaled_flag) {
        have_b = 0;
        for (int k = 0; k < 3 && !have_b; k++) if (B[k]) { bx = B[k]->mvx; by = B[k]->mvy; have_b = 1; if (B[k]->ref_poc != target) scale_mv(&bx, &by, c->poc - B[k]->ref_poc, c->poc - target); }
    }
    int16_t lx[3], ly[3]; int n = 0;
    if (have_a) { lx[n] = ax; ly[n] = ay; n++; }
    if (have_b && !(have_a && ax == bx && ay == by)) { lx[n] = bx; ly[n] = by; n++; }
    if (n < 2) { int16_t tx, ty; if (temporal_mv(c, xp, yp, w, h, ref_idx, &tx, &ty)) { lx[n] = tx; ly[n] = ty; n++; } }
    while (n < 2) { lx[n] = 0; ly[n] = 0; n++; }
    *px = lx[mvp_idx]; *py = ly[mvp_idx];
}
/* motion compensation of one block from a decoded picture, reference samples clamped to the picture (8.5.3.3.3) */
static void mc_block(const dpic *ref, int W, int H, int ci, int x0, int y0, int w, int h, int mvx, int mvy, uint8_t *dst, int ds)
{
    const int sh = ci ? 1 : 0, pw = W >> sh, ph = H >> sh, taps = ci ? 4 : 8, before = taps / 2 - 1, fb = ci ? 3 : 2;
    const uint8_t *plane = ref->pix + (ci == 0 ? 0 : (ci == 1 ? (size_t)W * H : (size_t)W * H + (size_t)W * H / 4));
    const int ix = x0 + (mvx >> fb), iy = y0 + (mvy >> fb), tw = w + taps - 1, th = h + taps - 1;
    uint8_t tmp[(64 + 7) * (64 + 7)];
    for (int y = 0; y < th; y++) for (int x = 0; x < tw; x++)
        tmp[y * tw + x] = plane[(size_t)clip3i(0, ph - 1, iy - before + y) * pw + clip3i(0, pw - 1, ix - before + x)];
    const int mask = (1 << fb) - 1;
    if (ci) ora_mc_chroma(dst, ds, tmp + before * tw + before, tw, w, h, mvx & mask, mvy & mask);
    else ora_mc_luma(dst, ds, tmp + before * tw + before, tw, w, h, mvx & mask, mvy & mask);
}

/* one picture (I or P slice) into c->r (pre-filter reconstruction, then deblocked in place); `out` receives the SAO output */
static int replay_picture(rctx *c, uint8_t *out)
{
    const ora_parsed_stream *ps = c->ps; const ora_parsed_pic *pp = c->pp;
    if (!pp->ok) return -2;
    if (pp->any_qp_delta) return -3;
    if (pp->st.slice_type == 0) return -5;                                      /* B slices: not covered */
    rpic *r = &c->r;
    const int W = r->w, H = r->h, qp = pp->st.qp, dw = r->dw;
    const size_t ysz = (size_t)W * H;
    const int qpc[3] = {qp, ora_chroma_qp[clip3i(0, 57, qp + pp->cb_qp_off)], ora_chroma_qp[clip3i(0, 57, qp + pp->cr_qp_off)]};
    memset(r->done, 0, (size_t)dw * ((H + 3) >> 2)); memset(c->cbfy, 0, (size_t)dw * ((H + 3) >> 2)); memset(c->intra, 0, (size_t)dw * ((H + 3) >> 2));
    memset(c->mv, 0, sizeof(minfo) * (size_t)dw * ((H + 3) >> 2));
    /* block edges on the 8x8 grid, one flag per 4-sample segment: vedge[(y/4) * ew + x/8], hedge[(y/8) * dw + x/4] */
    const int ew = (W + 7) >> 3, eh = (H + 7) >> 3;
    uint8_t *vedge = (uint8_t *)calloc((size_t)ew * ((H + 3) >> 2), 1), *hedge = (uint8_t *)calloc((size_t)dw * eh, 1);
    int rc = 0;
    for (size_t k = 0; k < pp->n_cus && !rc; k++) {
        const ora_cu_rec *cu = &pp->cus[k];
        const int S = 1 << cu->log2, half = S >> 1;
        if (cu->pred_mode == 0) {
            if (cu->part_mode != 0) { rc = -6; break; }                         /* only 2Nx2N inter prediction blocks */
            mcand m;
            if (cu->merge[0]) m = merge_candidate(c, cu->x, cu->y, S, S, cu->merge_idx[0]);
            else {
                int16_t px, py;
                m.ref_idx = cu->ref_idx[0][0];
                if (m.ref_idx >= pp->n_list0) { rc = -7; break; }
                amvp_predictor(c, cu->x, cu->y, S, S, m.ref_idx, cu->mvp[0][0], &px, &py);
                m.mvx = (int16_t)(px + cu->mvd[0][0][0]); m.mvy = (int16_t)(py + cu->mvd[0][0][1]);
            }
            if (m.ref_idx >= pp->n_list0) { rc = -7; break; }
            const dpic *ref = find_pic(c, pp->list0_poc[m.ref_idx]);
            if (!ref) { rc = -8; break; }
            mc_block(ref, W, H, 0, cu->x, cu->y, S, S, m.mvx, m.mvy, r->p[0] + (size_t)cu->y * W + cu->x, W);
            for (int ci = 1; ci < 3; ci++) mc_block(ref, W, H, ci, cu->x >> 1, cu->y >> 1, half, half, m.mvx, m.mvy, r->p[ci] + (size_t)(cu->y >> 1) * (W >> 1) + (cu->x >> 1), W >> 1);
            for (int y = cu->y >> 2; y < ((cu->y + S) >> 2) && y < ((H + 3) >> 2); y++) for (int x = cu->x >> 2; x < ((cu->x + S) >> 2) && x < dw; x++) {
                minfo *mi = &c->mv[y * dw + x]; mi->mvx = m.mvx; mi->mvy = m.mvy; mi->ref_idx = (int8_t)m.ref_idx; mi->ref_poc = pp->list0_poc[m.ref_idx]; mi->inter = 1;
                r->done[y * dw + x] = 1;
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
                    ora_dequant(lev, coef, 1 << l2c, qpc[ci], l2c);
                    ora_idct_add(coef, dst, dst, 1 << l2c, W >> 1, W >> 1, l2c, 0);
                }
            }
        }
        if (cu->pred_mode == 1 && cu->n_tu == 0) { rc = -9; break; }
    }
    /* 8.7.2: all vertical edges of the picture, then all horizontal ones */
    if (!rc && !pp->dbk_disabled) {
        const int beta = ora_beta_table[clip3i(0, 51, qp + 2 * pp->beta_off_div2)];
        int tcc[3] = {0, 0, 0};
        for (int ci = 1; ci < 3; ci++)             /* 8.7.2.5.5: QpC from the luma QP + the chroma offset, Bs 2 */
            tcc[ci] = ora_tc_table[clip3i(0, 53, ora_chroma_qp[clip3i(0, 57, qp + (ci == 1 ? pp->cb_qp_off : pp->cr_qp_off))] + 2 + 2 * pp->tc_off_div2)];
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
                    else {
                        const minfo *mq = &c->mv[iq], *mp = &c->mv[ip];
                        if (mq->ref_poc != mp->ref_poc || abs(mq->mvx - mp->mvx) >= 4 || abs(mq->mvy - mp->mvy) >= 4) bs = 1;
                    }
                    if (!bs) continue;
                    const int tc = ora_tc_table[clip3i(0, 53, qp + 2 * (bs - 1) + 2 * pp->tc_off_div2)];
                    ora_deblock_luma_seg(r->p[0] + (size_t)yq * W + xq, dir ? W : 1, dir ? 1 : W, beta, tc);
                    if (bs == 2 && !(e & 8)) for (int ci = 1; ci < 3; ci++)
                        ora_deblock_chroma_seg(r->p[ci] + (size_t)(yq >> 1) * (W >> 1) + (xq >> 1), dir ? (W >> 1) : 1, dir ? 1 : (W >> 1), tcc[ci], 2);
                }
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
int ora_replay_pictures(const ora_parsed_stream *ps, int first, int count, uint8_t *out)
{
    if (!ps || first < 0 || count < 1 || first + count > ps->n_pics) return -1;
    const int W = ps->width, H = ps->height, dw = (W + 3) >> 2, nb = dw * ((H + 3) >> 2);
    const size_t ysz = (size_t)W * H, fsz = ysz * 3 / 2;
    rctx c; memset(&c, 0, sizeof(c));
    c.ps = ps; c.r.w = W; c.r.h = H; c.r.dw = dw;
    uint8_t *pre = (uint8_t *)calloc(fsz, 1);
    c.r.p[0] = pre; c.r.p[1] = pre + ysz; c.r.p[2] = pre + ysz + ysz / 4;
    c.r.done = (uint8_t *)calloc((size_t)nb, 1); c.cbfy = (uint8_t *)calloc((size_t)nb, 1); c.intra = (uint8_t *)calloc((size_t)nb, 1);
    c.mv = (minfo *)calloc((size_t)nb, sizeof(minfo));
    dpic dpb[DPB_N]; memset(dpb, 0, sizeof(dpb)); c.dpb = dpb;
    int rc = 0, slot = 0;
    /* pictures before `first` that the requested ones may reference are replayed too (from the closest IDR back) */
    int start = first; while (start > 0 && ps->pics[start].st.nal_type != 19 && ps->pics[start].st.nal_type != 20) start--;
    uint8_t *scratch = (uint8_t *)malloc(fsz);
    for (int i = start; i < first + count && !rc; i++) {
        c.pp = &ps->pics[i]; c.poc = c.pp->st.poc;
        if (c.pp->st.nal_type == 19 || c.pp->st.nal_type == 20) for (int k = 0; k < DPB_N; k++) dpb[k].valid = 0;
        uint8_t *dst = i >= first ? out + fsz * (size_t)(i - first) : scratch;
        rc = replay_picture(&c, dst);
        if (rc) break;
        dpic *d = &dpb[slot]; slot = (slot + 1) % DPB_N;
        if (!d->pix) { d->pix = (uint8_t *)malloc(fsz); d->mv = (minfo *)malloc(sizeof(minfo) * (size_t)nb); }
        memcpy(d->pix, dst, fsz); memcpy(d->mv, c.mv, sizeof(minfo) * (size_t)nb); d->poc = c.poc; d->valid = 1;
    }
    for (int k = 0; k < DPB_N; k++) { free(dpb[k].pix); free(dpb[k].mv); }
    free(scratch); free(pre); free(c.r.done); free(c.cbfy); free(c.intra); free(c.mv);
    return rc;
}

/* the intra-only entry point of the first version: one I picture */
int ora_replay_intra_picture(const ora_parsed_stream *ps, int pic, uint8_t *out)
{
    if (!ps || pic < 0 || pic >= ps->n_pics) return -1;
    if (ps->pics[pic].st.slice_type != 2) return -2;
    return ora_replay_pictures(ps, pic, 1, out);
}
