/*
 * ora_replay.c -- SURVEY.md 8c tier P2 on the CPU: re-create the reference encoder's reconstruction of an INTRA picture from its own parsed
 * decisions (ora_parse.c: CU quadtree, part modes, intra modes, transform tree, levels, SAO parameters, loop-filter settings), using only the
 * oracle's leaf kernels (ora_kernels.c: intra prediction 4..32 incl. strong smoothing, dequantiser, IDCT 4..32 + IDST, deblocking segments,
 * SAO apply).  The result must equal what the reference DECODER makes of the same stream, byte for byte (tests/test_replay.py): that pins those
 * kernels against the reference at every block size its encoder uses -- 4x4 NxN partitions, 32x32 CUs -- not only at the sizes our own
 * streams exercise (tier P1).  TEST INFRASTRUCTURE: nothing under ks265codec_b200/ links this.
 * Limits: I slices, one slice per picture, cu_qp_delta all zero (the reference at -rc 0), no PCM / transform skip / scaling lists.
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

/* out: coded-size I420 (width x height of the SPS).  Returns 0, or a negative code for streams outside the limits above. */
int ora_replay_intra_picture(const ora_parsed_stream *ps, int pic, uint8_t *out)
{
    if (!ps || pic < 0 || pic >= ps->n_pics) return -1;
    const ora_parsed_pic *pp = &ps->pics[pic];
    if (pp->st.slice_type != 2 || !pp->ok) return -2;
    if (pp->any_qp_delta) return -3;
    const int W = ps->width, H = ps->height, qp = pp->st.qp;
    const int qpc[3] = {qp, ora_chroma_qp[qp + pp->cb_qp_off < 0 ? 0 : (qp + pp->cb_qp_off > 57 ? 57 : qp + pp->cb_qp_off)],
                        ora_chroma_qp[qp + pp->cr_qp_off < 0 ? 0 : (qp + pp->cr_qp_off > 57 ? 57 : qp + pp->cr_qp_off)]};
    rpic r; r.w = W; r.h = H; r.dw = (W + 3) >> 2;
    const size_t ysz = (size_t)W * H;
    uint8_t *pre = (uint8_t *)calloc(ysz * 3 / 2, 1);                          /* reconstruction before / after deblocking (in place) */
    r.p[0] = pre; r.p[1] = pre + ysz; r.p[2] = pre + ysz + ysz / 4;
    r.done = (uint8_t *)calloc((size_t)r.dw * ((H + 3) >> 2), 1);
    /* transform-block edges on the 8x8 grid, one flag per 4-sample segment: vedge[(y/4) * ew + x/8], hedge[(y/8) * dw + x/4] */
    const int ew = (W + 7) >> 3, eh = (H + 7) >> 3;
    uint8_t *vedge = (uint8_t *)calloc((size_t)ew * ((H + 3) >> 2), 1), *hedge = (uint8_t *)calloc((size_t)r.dw * eh, 1);
    int rc = 0;
    for (size_t c = 0; c < pp->n_cus && !rc; c++) {
        const ora_cu_rec *cu = &pp->cus[c];
        if (cu->pred_mode != 1) { rc = -4; break; }
        const int half = 1 << (cu->log2 - 1);
        for (uint32_t k = 0; k < cu->n_tu; k++) {
            const ora_tu_rec *t = &pp->tus[cu->first_tu + k];
            const int n = 1 << t->log2;
            const int part = cu->part_mode == 3 ? ((t->x - cu->x >= half) ? 1 : 0) | ((t->y - cu->y >= half) ? 2 : 0) : 0;
            recon_block(&r, 0, t->x, t->y, t->log2, cu->intra_mode[part], ps->strong_intra, qpc[0], (t->cbf & 1) ? pp->lev + t->lev_off[0] : NULL);
            for (int y = t->y >> 2; y < ((t->y + n) >> 2); y++) for (int x = t->x >> 2; x < ((t->x + n) >> 2); x++) r.done[y * r.dw + x] = 1;
            if ((t->x & 7) == 0 && t->x > 0) for (int y = t->y >> 2; y < ((t->y + n) >> 2); y++) vedge[y * ew + (t->x >> 3)] = 1;
            if ((t->y & 7) == 0 && t->y > 0) for (int x = t->x >> 2; x < ((t->x + n) >> 2); x++) hedge[(t->y >> 3) * r.dw + x] = 1;
            /* chroma: with the luma block, or -- 4x4 luma blocks -- one 4x4 block per component after the fourth luma block of the 8x8 parent */
            int xc, yc, l2c;
            if (t->log2 > 2) { xc = t->x >> 1; yc = t->y >> 1; l2c = t->log2 - 1; }
            else if ((t->x & 4) && (t->y & 4)) { xc = (t->x - 4) >> 1; yc = (t->y - 4) >> 1; l2c = 2; }
            else continue;
            for (int ci = 1; ci < 3; ci++)
                recon_block(&r, ci, xc, yc, l2c, cu->chroma_mode, 0, qpc[ci], (t->cbf & (1 << ci)) ? pp->lev + t->lev_off[ci] : NULL);
        }
    }
    /* 8.7.2: all vertical edges of the picture, then all horizontal ones; every edge of an intra picture has Bs 2 */
    if (!rc && !pp->dbk_disabled) {
        const int beta = ora_beta_table[qp + 2 * pp->beta_off_div2 < 0 ? 0 : (qp + 2 * pp->beta_off_div2 > 51 ? 51 : qp + 2 * pp->beta_off_div2)];
        int ti = qp + 2 + 2 * pp->tc_off_div2; ti = ti < 0 ? 0 : (ti > 53 ? 53 : ti);
        const int tc = ora_tc_table[ti];
        int tcc[3] = {0, 0, 0};
        for (int ci = 1; ci < 3; ci++) {            /* 8.7.2.5.5: QpC from the luma QP + the PPS offset (cQpPicOffset), not the slice's */
            int q = ora_chroma_qp[qpc[0] + (ci == 1 ? pp->cb_qp_off : pp->cr_qp_off) < 0 ? 0 : qpc[0] + (ci == 1 ? pp->cb_qp_off : pp->cr_qp_off)];
            int i2 = q + 2 + 2 * pp->tc_off_div2; i2 = i2 < 0 ? 0 : (i2 > 53 ? 53 : i2);
            tcc[ci] = ora_tc_table[i2];
        }
        for (int dir = 0; dir < 2; dir++)
            for (int e = 8; e < (dir ? H : W); e += 8)
                for (int s = 0; s < (dir ? W : H); s += 4) {
                    if (!(dir ? hedge[(e >> 3) * r.dw + (s >> 2)] : vedge[(s >> 2) * ew + (e >> 3)])) continue;
                    const int xq = dir ? s : e, yq = dir ? e : s;
                    ora_deblock_luma_seg(r.p[0] + (size_t)yq * W + xq, dir ? W : 1, dir ? 1 : W, beta, tc);
                    if (!(e & 8)) for (int ci = 1; ci < 3; ci++)
                        ora_deblock_chroma_seg(r.p[ci] + (size_t)(yq >> 1) * (W >> 1) + (xq >> 1), dir ? (W >> 1) : 1, dir ? 1 : (W >> 1), tcc[ci], 2);
                }
    }
    /* 8.7.3: SAO reads the deblocked picture and writes the output picture */
    if (!rc) {
        memcpy(out, pre, ysz * 3 / 2);
        const int l = ps->log2_ctb, ctw = (W + (1 << l) - 1) >> l, cth = (H + (1 << l) - 1) >> l;
        for (int ry = 0; ry < cth; ry++) for (int rx = 0; rx < ctw; rx++) {
            const ora_sao_rec *sr = &pp->sao[ry * ctw + rx];
            for (int ci = 0; ci < 3; ci++) {
                if (!sr->type[ci]) continue;
                const int sh = ci ? 1 : 0, pw = W >> sh, ph = H >> sh, x0 = (rx << l) >> sh, y0 = (ry << l) >> sh, cs = (1 << l) >> sh;
                const int w = x0 + cs > pw ? pw - x0 : cs, h = y0 + cs > ph ? ph - y0 : cs;
                const size_t off = ci == 0 ? 0 : (ci == 1 ? ysz : ysz + ysz / 4);
                ora_sao_apply_ctb(out + off, pw, pre + off, pw, x0, y0, w, h, pw, ph, sr->type[ci], sr->pos[ci], sr->off[ci]);
            }
        }
    }
    free(pre); free(r.done); free(vedge); free(hedge);
    return rc;
}
