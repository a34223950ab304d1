/* ora_frame.c -- CPU model of the picture-level hot path (TEST INFRASTRUCTURE; see ora_frame.h). */
#include "ora_frame.h"
#include "ks_oracle.h"
#include <stdlib.h>
#include <string.h>

const int ora_lambda_sad_q4[52] = {4,4,5,5,6,7,7,8,9,10,12,13,15,17,19,21,23,26,30,33,37,42,47,53,59,66,74,83,94,105,118,132,149,167,187,210,236,265,297,334,375,421,472,530,595,668,749,841,944,1060,1189,1335};
const int ora_lambda_sse_q4[52] = {1,1,1,2,2,3,3,4,5,7,9,11,14,17,22,27,34,43,54,69,86,109,137,173,218,274,345,435,548,691,870,1097,1382,1741,2193,2763,3482,4387,5527,6963,8773,11053,13926,17546,22107,27853,35092,44214,55706,70185,88427,111411};

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------ pictures --------------------- */
int ora_pic_alloc(ora_pic *pic, int w, int h)
{
    for (int i = 0; i < 3; i++) {
        int pw = i ? w / 2 : w, ph = i ? h / 2 : h;
        ora_plane *p = &pic->c[i];
        p->w = pw; p->h = ph; p->stride = pw + 2 * ORA_PAD;
        p->base = (uint8_t *)calloc((size_t)p->stride * (ph + 2 * ORA_PAD), 1);
        if (!p->base) return -1;
        p->p = p->base + (size_t)ORA_PAD * p->stride + ORA_PAD;
    }
    return 0;
}
void ora_pic_free(ora_pic *pic) { for (int i = 0; i < 3; i++) { free(pic->c[i].base); pic->c[i].base = NULL; } }
void ora_pic_extend(ora_pic *pic)
{
    for (int i = 0; i < 3; i++) {
        ora_plane *p = &pic->c[i];
        for (int y = 0; y < p->h; y++) {
            uint8_t *r = p->p + (size_t)y * p->stride;
            memset(r - ORA_PAD, r[0], ORA_PAD); memset(r + p->w, r[p->w - 1], ORA_PAD);
        }
        for (int y = 1; y <= ORA_PAD; y++) {
            memcpy(p->p - (size_t)y * p->stride - ORA_PAD, p->p - ORA_PAD, (size_t)p->stride);
            memcpy(p->p + (size_t)(p->h - 1 + y) * p->stride - ORA_PAD, p->p + (size_t)(p->h - 1) * p->stride - ORA_PAD, (size_t)p->stride);
        }
    }
}
void ora_pic_load(ora_pic *pic, const uint8_t *i420, int sw, int sh)
{
    const uint8_t *s = i420;
    for (int i = 0; i < 3; i++) {
        ora_plane *p = &pic->c[i];
        int w = i ? sw / 2 : sw, h = i ? sh / 2 : sh;
        for (int y = 0; y < p->h; y++) {
            const uint8_t *sr = s + (size_t)imin(y, h - 1) * w;
            uint8_t *d = p->p + (size_t)y * p->stride;
            memcpy(d, sr, (size_t)w);
            for (int x = w; x < p->w; x++) d[x] = sr[w - 1];
        }
        s += (size_t)w * h;
    }
    ora_pic_extend(pic);
}

/* ------------------------------------------------------------------ helpers ---------------------- */
static uint16_t scan_tb[4][1024];    /* per log2 (2..5): scan position -> (y<<8)|x, diag CG order x diag 4x4 */
static uint16_t scan_hv[2][2][64];   /* [log2 - 2][0 horizontal, 1 vertical]: the mode-dependent scans of intra 4x4 / 8x8 blocks (7.4.9.11, 6.5.4/6.5.5) */
static int scan_ready;
static void build_scans(void)
{
    if (scan_ready) return;
    uint8_t d4[16], dcg[64];
    for (int l = 2; l <= 5; l++) {
        int ncg = 1 << (l - 2);
        for (int pass = 0; pass < 2; pass++) {
            int n = pass ? ncg : 4, i = 0, x = 0, y = 0; uint8_t *dst = pass ? dcg : d4;
            for (;;) { while (y >= 0) { if (x < n && y < n) dst[i++] = (uint8_t)((y << 3) | x); y--; x++; } y = x; x = 0; if (i >= n * n) break; }
        }
        for (int c = 0; c < ncg * ncg; c++)
            for (int k = 0; k < 16; k++) {
                int xx = ((dcg[c] & 7) << 2) + (d4[k] & 7), yy = ((dcg[c] >> 3) << 2) + (d4[k] >> 3);
                scan_tb[l - 2][c * 16 + k] = (uint16_t)((yy << 8) | xx);
            }
    }
    for (int l = 2; l <= 3; l++) {
        int ncg = 1 << (l - 2);
        for (int c = 0; c < ncg * ncg; c++) for (int k = 0; k < 16; k++) {
            /* horizontal: groups row by row, samples row by row; vertical: the transpose */
            int gx = c % ncg, gy = c / ncg, px = k & 3, py = k >> 2;
            scan_hv[l - 2][0][c * 16 + k] = (uint16_t)((((gy << 2) + py) << 8) | ((gx << 2) + px));
            scan_hv[l - 2][1][c * 16 + k] = (uint16_t)((((gx << 2) + px) << 8) | ((gy << 2) + py));
        }
    }
    scan_ready = 1;
}
/* scan of an intra transform block (7.4.9.11): luma 4x4 / 8x8 and chroma 4x4 blocks scan horizontally for the near-vertical modes 22..30,
 * vertically for the near-horizontal modes 6..14, diagonally otherwise */
static const uint16_t *intra_scan(int log2, int is_luma, int mode)
{
    if (log2 == 2 || (log2 == 3 && is_luma)) {
        if (mode >= 22 && mode <= 30) return scan_hv[log2 - 2][0];
        if (mode >= 6 && mode <= 14) return scan_hv[log2 - 2][1];
    }
    return scan_tb[log2 - 2];
}
static inline int zidx(int x, int y)
{   /* z-scan index of the 8x8 block containing (x,y) inside its CTB (equivalent to the 16x16 order for blocks of 16 and up) */
    int cx = (x >> 3) & 7, cy = (y >> 3) & 7;
    return (cx & 1) | ((cy & 1) << 1) | ((cx & 2) << 1) | ((cy & 2) << 2) | ((cx & 4) << 2) | ((cy & 4) << 3);
}
static int avail(const ora_cfg *cfg, int xc, int yc, int xn, int yn)
{   /* H.265 6.4.1, one slice, CTB 64, 8x8 granularity */
    if (xn < 0 || yn < 0 || xn >= cfg->width || yn >= cfg->height) return 0;
    int cw = (cfg->width + 63) >> 6;
    int ac = (yc >> 6) * cw + (xc >> 6), an = (yn >> 6) * cw + (xn >> 6);
    if (an != ac) return an < ac;
    return zidx(xn, yn) < zidx(xc, yc);
}
static inline int mvbits(int d) { int a = iabs(d); return a ? 2 * (32 - __builtin_clz((unsigned)a)) + 1 : 1; }

/* estimated bits of a transform block's levels (fit against the host CABAC on natural clips: 3 per non-zero level + 2 per magnitude
 * doubling + 4 per coded 4x4 group + the anti-diagonal of the outermost level); 0 for an all-zero block */
static int level_bits_est(const int16_t *q, int n)
{
    int nnz = 0, slog = 0, ncg = 0, maxd = 0;
    for (int gy = 0; gy < n; gy += 4) for (int gx = 0; gx < n; gx += 4) {
        int any = 0;
        for (int y = gy; y < gy + 4; y++) for (int x = gx; x < gx + 4; x++) {
            int a = iabs(q[y * n + x]);
            if (!a) continue;
            any = 1; nnz++; slog += 31 - __builtin_clz((unsigned)a); if (x + y > maxd) maxd = x + y;
        }
        ncg += any;
    }
    return nnz ? 3 * nnz + 2 * slog + 4 * ncg + maxd : 0;
}
/* one transform block: residual -> fdct -> quant (-> sign hiding) -> dequant -> idct+pred.  returns cbf.
 * rdz: RD zero-out (inter blocks; reference: the zero-block / skip decisions of tuDecision E@0x47e2f0 and skipFastDecision E@0x47f720 -- closed,
 * this is our rule): drop the levels when SSE(src,pred) <= SSE(src,rec) + lambda * bits. */
static int64_t tb_d1; static int tb_bits;      /* of the last code_tb call: SSE(src, rec) and level_bits_est of the block as coded */
static int code_tb(const ora_cfg *cfg, int qp, int intra_slice, int log2, int is_dst,
                   const uint8_t *src, int ss, const uint8_t *pred, int ps, uint8_t *rec, int rs, int16_t *lev, int ls, int rdz_lambda_q4, const uint16_t *scan)
{
    int n = 1 << log2;
    int16_t res[1024], coef[1024], q[1024], du[1024], deq[1024];
    ora_residual(res, src, pred, ss, ps, n);
    ora_fdct(res, coef, n, n, log2, is_dst);
    int nnz = ora_quant(coef, q, n, qp, log2, intra_slice, du);
    if (nnz && cfg->sign_hiding) nnz = ora_sign_hide(coef, q, du, n, log2, scan ? scan : scan_tb[log2 - 2]);
    if (nnz) {
        ora_dequant(q, deq, n, qp, log2);
        ora_idct_add(deq, rec, pred, n, rs, ps, log2, is_dst);
        if (rdz_lambda_q4) {
            int64_t d0 = 0, d1 = 0;
            for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) {
                int a = res[y * n + x], b = (int)src[y * ss + x] - (int)rec[y * rs + x];
                d0 += a * a; d1 += b * b;
            }
            if (d0 * 16 <= d1 * 16 + (int64_t)rdz_lambda_q4 * level_bits_est(q, n)) { nnz = 0; memset(q, 0, sizeof(int16_t) * (size_t)n * n); }
        }
    }
    for (int y = 0; y < n; y++) memcpy(lev + (size_t)y * ls, q + y * n, (size_t)n * 2);
    if (!nnz) for (int y = 0; y < n; y++) memcpy(rec + (size_t)y * rs, pred + (size_t)y * ps, (size_t)n);
    tb_d1 = 0; tb_bits = nnz ? level_bits_est(q, n) : 0;
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) { int e = (int)src[y * ss + x] - (int)rec[y * rs + x]; tb_d1 += e * e; }
    return nnz != 0;
}

/* The levels OUR transform-block coder makes of (src, pred): forward transform, quantiser, sign-data hiding -- no RD zero-out.  ora_replay.c
 * holds them against the levels the reference encoder coded for the same block (its prediction re-created from its own stream).
 * intra_mode < 0: inter block (diagonal scan, DCT). */
int ora_tb_levels(int qp, int intra_slice, int log2, int is_luma, int intra_mode, int sign_hiding, const uint8_t *src, int ss, const uint8_t *pred, int ps, int16_t *lev)
{
    int n = 1 << log2;
    int16_t res[1024], coef[1024], du[1024];
    build_scans();
    ora_residual(res, src, pred, ss, ps, n);
    ora_fdct(res, coef, n, n, log2, intra_mode >= 0 && is_luma && log2 == 2);
    int nnz = ora_quant(coef, lev, n, qp, log2, intra_slice, du);
    if (nnz && sign_hiding) nnz = ora_sign_hide(coef, lev, du, n, log2, intra_mode >= 0 ? intra_scan(log2, is_luma, intra_mode) : scan_tb[log2 - 2]);
    return nnz;
}

/* Would OUR inter transform-block coder keep levels for (src, pred)?  Returns 1 / 0 after the RD zero-out at lambda_qp; *plain_nnz = non-zero
 * levels of the plain quantiser + sign hiding before that decision.  ora_replay.c holds this against the reference's coded-block flags. */
int ora_tb_decision(int qp, int lambda_qp, int log2, int sign_hiding, const uint8_t *src, int ss, const uint8_t *pred, int ps, int *plain_nnz)
{
    ora_cfg cfg; memset(&cfg, 0, sizeof(cfg)); cfg.sign_hiding = sign_hiding;
    uint8_t rec[32 * 32]; int16_t lev[32 * 32];
    build_scans();
    if (plain_nnz) { *plain_nnz = 0; code_tb(&cfg, qp, 0, log2, 0, src, ss, pred, ps, rec, 32, lev, 32, 0, NULL); for (int y = 0; y < (1 << log2); y++) for (int x = 0; x < (1 << log2); x++) *plain_nnz += lev[y * 32 + x] != 0; }
    return code_tb(&cfg, qp, 0, log2, 0, src, ss, pred, ps, rec, 32, lev, 32, ora_lambda_sse_q4[clip3(0, 51, lambda_qp)], NULL);
}

/* ------------------------------------------------------------------ intra picture ---------------- */
static void build_nb(const ora_cfg *cfg, const ora_plane *rec, int comp, int x0, int y0, int n, uint8_t *nb)
{   /* 8.4.4.2.2 reference sample availability + substitution; (x0,y0) in component samples */
    int sh = comp ? 1 : 0, lx = x0 << sh, ly = y0 << sh, tot = 4 * n + 1;
    uint8_t av[129];
    int any = 0;
    for (int i = 0; i < tot; i++) {
        int xn, yn;
        if (i < 2 * n) { xn = x0 - 1; yn = y0 + 2 * n - 1 - i; }
        else if (i == 2 * n) { xn = x0 - 1; yn = y0 - 1; }
        else { xn = x0 + i - 2 * n - 1; yn = y0 - 1; }
        av[i] = (uint8_t)avail(cfg, lx, ly, xn * (1 << sh), yn * (1 << sh));
        if (av[i]) { nb[i] = rec->p[(size_t)yn * rec->stride + xn]; any = 1; }
    }
    if (!any) { memset(nb, 128, (size_t)tot); return; }
    if (!av[0]) { int i = 1; while (!av[i]) i++; nb[0] = nb[i]; }
    for (int i = 1; i < tot; i++) if (!av[i]) nb[i] = nb[i - 1];
}
/* intra ESTIMATE of an n x n block (n = 16 or 8) in the search metric: best of DC / horizontal / vertical / planar predicted from the SOURCE
 * picture's neighbours (left column, top row, the samples right of / below them clamped to the picture; 128 where the picture ends; no
 * boundary smoothing).  It decides inter vs intra (stage D) and 16x16 vs four 8x8 intra CUs; the real mode search runs on reconstructed
 * neighbours (intra_block). */
static int intra_estimate(const ora_cfg *cfg, const ora_plane *src, int x0, int y0, int n)
{
    const uint8_t *s = src->p + (size_t)y0 * src->stride + x0;
    int W = cfg->width, H = cfg->height, st = src->stride, lg = n == 16 ? 4 : 3;
    int left[16], top[16], hl = x0 > 0, ht = y0 > 0, sl = 0, stp = 0;
    for (int i = 0; i < n; i++) { left[i] = hl ? s[i * st - 1] : 128; top[i] = ht ? s[i - st] : 128; sl += left[i]; stp += top[i]; }
    int tr = ht ? src->p[(size_t)(y0 - 1) * st + imin(x0 + n, W - 1)] : 128, bl = hl ? src->p[(size_t)imin(y0 + n, H - 1) * st + x0 - 1] : 128;
    int dc = (hl && ht) ? (sl + stp + n) >> (lg + 1) : (hl ? (sl + n / 2) >> lg : (ht ? (stp + n / 2) >> lg : 128));
    int best = 0x7fffffff;
    for (int m = 0; m < 4; m++) {
        uint8_t pred[256];
        for (int y = 0; y < n; y++) for (int x = 0; x < n; x++)
            pred[y * n + x] = (uint8_t)(m == 0 ? dc : (m == 1 ? left[y] : (m == 2 ? top[x] : ((n - 1 - x) * left[y] + (x + 1) * tr + (n - 1 - y) * top[x] + (y + 1) * bl + n) >> (lg + 1))));
        int c = (int)((cfg->satd && cfg->subpel > 0) ? ora_satd(s, pred, st, n, n, n) : ora_sad(s, pred, st, n, n, n));
        if (c < best) best = c;
    }
    return best;
}
/* one intra CU of 16x16 (lg = 4) or 8x8 (lg = 3): 35-mode decision by SAD + lambda*bits (source neighbours), then prediction from the reconstructed neighbours, luma +
 * chroma (DM) residual coding with the mode-dependent scans.  cost_q4 = 16 * SSE(Y,Cb,Cr) + lambda_sse * (estimated level bits + 1 per block) */
typedef struct { int mode, cbf, luma_bits; int64_t cost_q4; } intra_res;
static intra_res intra_block(const ora_cfg *cfg, int qp, int intra_slice, const ora_pic *src, ora_pic *rec, ora_levels *lv, int x0, int y0, int lg)
{
    int W = cfg->width, n = 1 << lg;
    int lam = ora_lambda_sad_q4[qp], lamq = ora_lambda_sse_q4[qp], qpc = ora_chroma_qp[qp];
    uint8_t nb[65], pred[256];
    const uint8_t *s = src->c[0].p + (size_t)y0 * src->c[0].stride + x0;
    /* the mode is chosen against the SOURCE picture's neighbours: that takes the 35-mode search off the reconstruction dependency chain (every block
     * of the picture can search at once); the prediction itself uses the reconstructed neighbours.  Costs 1-2 % bits [measured]. */
    build_nb(cfg, &src->c[0], 0, x0, y0, n, nb);
    intra_res r; r.mode = 0; r.cbf = 0; r.cost_q4 = 0;
    int best_cost = 0x7fffffff;
    for (int m = 0; m < 35; m++) {
        ora_intra_pred(pred, n, nb, lg, m, 1, cfg->strong_intra);
        int bits = (m == 0 || m == 1 || m == 10 || m == 26) ? 3 : 6;
        int cost = (int)ora_sad(s, pred, src->c[0].stride, n, n, n) + ((lam * bits) >> 4);
        if (cost < best_cost) { best_cost = cost; r.mode = m; }
    }
    build_nb(cfg, &rec->c[0], 0, x0, y0, n, nb);
    ora_intra_pred(pred, n, nb, lg, r.mode, 1, cfg->strong_intra);
    uint8_t *rp = rec->c[0].p + (size_t)y0 * rec->c[0].stride + x0;
    if (code_tb(cfg, qp, intra_slice, lg, 0, s, src->c[0].stride, pred, n, rp, rec->c[0].stride, lv->c[0] + (size_t)y0 * W + x0, W, 0, intra_scan(lg, 1, r.mode))) r.cbf |= KS_F_CBF_Y;
    r.luma_bits = tb_bits; r.cost_q4 += tb_d1 * 16 + (int64_t)lamq * (tb_bits + 1);
    for (int ci = 1; ci < 3; ci++) {
        int xc = x0 >> 1, yc = y0 >> 1, nc = n >> 1;
        uint8_t nbc[33], pc[64];
        build_nb(cfg, &rec->c[ci], ci, xc, yc, nc, nbc);
        ora_intra_pred(pc, nc, nbc, lg - 1, r.mode, 0, 0);
        const uint8_t *sc = src->c[ci].p + (size_t)yc * src->c[ci].stride + xc;
        uint8_t *rc = rec->c[ci].p + (size_t)yc * rec->c[ci].stride + xc;
        if (code_tb(cfg, qpc, intra_slice, lg - 1, 0, sc, src->c[ci].stride, pc, nc, rc, rec->c[ci].stride, lv->c[ci] + (size_t)yc * (W / 2) + xc, W / 2, 0, intra_scan(lg - 1, 0, r.mode)))
            r.cbf |= ci == 1 ? KS_F_CBF_CB : KS_F_CBF_CR;
        r.cost_q4 += tb_d1 * 16 + (int64_t)lamq * (tb_bits + 1);
    }
    return r;
}
/* one 16x16 intra cell: a 16x16 CU, or four 8x8 CUs coded in z-order when that is cheaper in J = SSE + lambda * bits (both are really coded;
 * the 8x8 alternative is only tried in I pictures and when the 16x16 luma block costs at least ORA_SPLIT8_MIN_BITS estimated bits: smooth or
 * noise-like blocks practically never gain -- 3 % of the attempts succeed on the synthetic clip against 70 % on natural content [measured]).  Reference: its intra CUs go down to 8x8 / 4x4 partitions; 8x8 CUs dominate its I pictures on natural content
 * [probe: tools/stream_stats.py].  Header bits: 8 for a 16x16 CU, 4 x 8 + 2 for the four. */
#define ORA_SPLIT8_MIN_BITS 200
static void intra_cell(const ora_cfg *cfg, int qp, int intra_slice, const ora_pic *src, ora_pic *rec, ks_cell *cells, ora_levels *lv, int x0, int y0)
{
    int cw = cfg->width >> 4, W = cfg->width, lamq = ora_lambda_sse_q4[qp];
    ks_cell *c = &cells[(y0 >> 4) * cw + (x0 >> 4)];
    memset(c, 0, sizeof(*c));
    intra_res r16 = intra_block(cfg, qp, intra_slice, src, rec, lv, x0, y0, 4);
    c->cu_log2 = 4; c->flags = (uint8_t)(KS_F_INTRA | r16.cbf); c->intra_mode = (uint8_t)r16.mode;
    if (!intra_slice || r16.luma_bits < ORA_SPLIT8_MIN_BITS) return;
    int64_t j16 = r16.cost_q4 + (int64_t)lamq * 8, j8 = (int64_t)lamq * (8 * 4 + 2);
    /* keep the 16x16 result, then code the four 8x8 CUs over it */
    uint8_t srec[3][256]; int16_t slev[3][256];
    for (int ci = 0; ci < 3; ci++) {
        int sh = ci ? 1 : 0, m = 16 >> sh, pw = W >> sh;
        for (int y = 0; y < m; y++) {
            memcpy(srec[ci] + y * m, rec->c[ci].p + (size_t)((y0 >> sh) + y) * rec->c[ci].stride + (x0 >> sh), (size_t)m);
            memcpy(slev[ci] + y * m, lv->c[ci] + (size_t)((y0 >> sh) + y) * pw + (x0 >> sh), (size_t)m * 2);
        }
    }
    int modes[4], cy = 0, cb = 0, cr = 0, any = 0;
    for (int k = 0; k < 4; k++) {
        intra_res r = intra_block(cfg, qp, intra_slice, src, rec, lv, x0 + 8 * (k & 1), y0 + 8 * (k >> 1), 3);
        modes[k] = r.mode; any |= r.cbf; j8 += r.cost_q4;
        if (r.cbf & KS_F_CBF_Y) cy |= 1 << k; if (r.cbf & KS_F_CBF_CB) cb |= 1 << k; if (r.cbf & KS_F_CBF_CR) cr |= 1 << k;
    }
    if (j8 < j16) {
        c->cu_log2 = 3; c->flags = (uint8_t)(KS_F_INTRA | any);
        c->mvx = (int16_t)(uint16_t)(modes[0] | (modes[1] << 8)); c->mvy = (int16_t)(uint16_t)(modes[2] | (modes[3] << 8));
        c->intra_mode = (uint8_t)(cy | (cb << 4)); c->rsv = (uint8_t)cr;
        return;
    }
    for (int ci = 0; ci < 3; ci++) {
        int sh = ci ? 1 : 0, m = 16 >> sh, pw = W >> sh;
        for (int y = 0; y < m; y++) {
            memcpy(rec->c[ci].p + (size_t)((y0 >> sh) + y) * rec->c[ci].stride + (x0 >> sh), srec[ci] + y * m, (size_t)m);
            memcpy(lv->c[ci] + (size_t)((y0 >> sh) + y) * pw + (x0 >> sh), slev[ci] + y * m, (size_t)m * 2);
        }
    }
}
void ora_intra_picture(const ora_cfg *cfg, int qp, const ora_pic *src, ora_pic *rec, ks_cell *cells, ora_levels *lv)
{
    build_scans();
    int W = cfg->width, H = cfg->height;
    int ctus_w = (W + 63) >> 6, ctus_h = (H + 63) >> 6;
    for (int cty = 0; cty < ctus_h; cty++) for (int ctx = 0; ctx < ctus_w; ctx++)
        for (int z = 0; z < 16; z++) {
            int cx = (z & 1) | ((z >> 1) & 2), cy = ((z >> 1) & 1) | ((z >> 2) & 2);
            int x0 = (ctx << 6) + (cx << 4), y0 = (cty << 6) + (cy << 4);
            if (x0 >= W || y0 >= H) continue;
            intra_cell(cfg, qp, 1, src, rec, cells, lv, x0, y0);
        }
}

/* ------------------------------------------------------------------ inter picture ---------------- */
static const int8_t dia_dx[4] = {0, 0, -1, 1}, dia_dy[4] = {-1, 1, 0, 0};
static const int8_t sq_dx[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, sq_dy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};

static int me_cell(const ora_cfg *cfg, int lam, const ora_plane *src, const ora_plane *ref, int x0, int y0, int tpx, int tpy, int *omx, int *omy, int *odist)
{   /* a5 (start point) + a3 (small diamond, x264 DIA) + a6 (half/quarter refinement with real interpolation, SAD cost) */
    const uint8_t *s = src->p + (size_t)y0 * src->stride + x0;
    const uint8_t *r0 = ref->p + (size_t)y0 * ref->stride + x0;
    int rs = ref->stride, ss = src->stride, R = cfg->me_range;
#define MVCOST(qx, qy) ((lam * (mvbits((qx) - tpx) + mvbits((qy) - tpy))) >> 4)
#define ICOST(ix, iy) ((int)ora_sad(s, r0 + (iy) * rs + (ix), ss, rs, 16, 16) + MVCOST((ix) * 4, (iy) * 4))
    int bx = 0, by = 0, bc = ICOST(0, 0);
    int cx = clip3(-R, R, (tpx + 2) >> 2), cy = clip3(-R, R, (tpy + 2) >> 2);
    if (cx || cy) { int c = ICOST(cx, cy); if (c < bc) { bc = c; bx = cx; by = cy; } }
    if (cfg->me_method == 0) {
        for (int it = 0; it < cfg->me_iters; it++) {
            int bk = -1, lc = bc;
            for (int k = 0; k < 4; k++) {
                int nx = bx + dia_dx[k], ny = by + dia_dy[k];
                if (iabs(nx) > R || iabs(ny) > R) continue;
                int c = ICOST(nx, ny);
                if (c < lc) { lc = c; bk = k; }
            }
            if (bk < 0) break;
            bx += dia_dx[bk]; by += dia_dy[bk]; bc = lc;
        }
    } else {
        /* a4: x264 hexagon search (me.c X264_ME_HEX): all six points once, then only the three new points of the hexagon
         * moved in the winning direction, rotating with mod6m1; finally the 8-neighbour square.  Ties keep the earlier point. */
        static const int hx[8] = {-1, -2, -1, 1, 2, 1, -1, -2}, hy[8] = {-2, 0, 2, 2, 0, -2, -2, 0}, mod6m1[8] = {5, 0, 1, 2, 3, 4, 5, 0};
        int dir = -1, lc = bc;
        for (int k = 0; k < 6; k++) {
            int nx = bx + hx[k + 1], ny = by + hy[k + 1];
            if (iabs(nx) > R || iabs(ny) > R) continue;
            int c = ICOST(nx, ny);
            if (c < lc) { lc = c; dir = k; }
        }
        if (dir >= 0) {
            bx += hx[dir + 1]; by += hy[dir + 1]; bc = lc;
            for (int it = 1; it < cfg->me_iters; it++) {
                int bk = -1; lc = bc;
                for (int j = 0; j < 3; j++) {
                    int nx = bx + hx[dir + j], ny = by + hy[dir + j];
                    if (iabs(nx) > R || iabs(ny) > R) continue;
                    int c = ICOST(nx, ny);
                    if (c < lc) { lc = c; bk = j; }
                }
                if (bk < 0) break;
                dir = mod6m1[dir + bk - 1 + 1];
                bx += hx[dir + 1]; by += hy[dir + 1]; bc = lc;
            }
        }
        static const int qx8[8] = {0, 0, -1, 1, -1, -1, 1, 1}, qy8[8] = {-1, 1, 0, 0, -1, 1, -1, 1};
        int bk = -1; lc = bc;
        for (int k = 0; k < 8; k++) {
            int nx = bx + qx8[k], ny = by + qy8[k];
            if (iabs(nx) > R || iabs(ny) > R) continue;
            int c = ICOST(nx, ny);
            if (c < lc) { lc = c; bk = k; }
        }
        if (bk >= 0) { bx += qx8[bk]; by += qy8[bk]; bc = lc; }
    }
    int mx = bx * 4, my = by * 4;
    if (cfg->satd && cfg->subpel > 0) bc = (int)ora_satd(s, r0 + by * rs + bx, ss, rs, 16, 16) + MVCOST(mx, my);
    for (int step = 2; step >= 1; step--) {
        if (cfg->subpel < (step == 2 ? 1 : 2)) break;
        int bk = -1, lc = bc;
        for (int k = 0; k < 8; k++) {
            int qx = mx + sq_dx[k] * step, qy = my + sq_dy[k] * step;
            uint8_t pred[256];
            ora_mc_luma(pred, 16, r0, rs, 16, 16, qx, qy);
            int c = (int)(cfg->satd ? ora_satd(s, pred, ss, 16, 16, 16) : ora_sad(s, pred, ss, 16, 16, 16)) + MVCOST(qx, qy);
            if (c < lc) { lc = c; bk = k; }
        }
        if (bk >= 0) { mx += sq_dx[bk] * step; my += sq_dy[bk] * step; bc = lc; }
    }
    *omx = mx; *omy = my;
    if (odist) *odist = bc - MVCOST(mx, my);
    return bc;
#undef ICOST
#undef MVCOST
}

static void recon_inter_cu(const ora_cfg *cfg, int qp, int qpc, int rdz, const ora_pic *src, const ora_pic *ref, ora_pic *rec,
                           ks_cell *cells, ora_levels *lv, int x0, int y0, int log2, int mvx, int mvy)
{
    int S = 1 << log2, W = cfg->width, cw = W >> 4;
    static uint8_t pred[3][64 * 64];
    ora_mc_luma(pred[0], 64, ref->c[0].p + (size_t)y0 * ref->c[0].stride + x0, ref->c[0].stride, S, S, mvx, mvy);
    for (int ci = 1; ci < 3; ci++)
        ora_mc_chroma(pred[ci], 64, ref->c[ci].p + (size_t)(y0 / 2) * ref->c[ci].stride + x0 / 2, ref->c[ci].stride, S / 2, S / 2, mvx, mvy);
    int T = imin(S, 32), tl = log2 > 5 ? 5 : log2;
    for (int ty = 0; ty < S; ty += T) for (int tx = 0; tx < S; tx += T) {
        int x = x0 + tx, y = y0 + ty, f = 0;
        if (code_tb(cfg, qp, 0, tl, 0, src->c[0].p + (size_t)y * src->c[0].stride + x, src->c[0].stride, pred[0] + ty * 64 + tx, 64,
                    rec->c[0].p + (size_t)y * rec->c[0].stride + x, rec->c[0].stride, lv->c[0] + (size_t)y * W + x, W, rdz, NULL)) f |= KS_F_CBF_Y;
        for (int ci = 1; ci < 3; ci++) {
            int xc = x / 2, yc = y / 2;
            if (code_tb(cfg, qpc, 0, tl - 1, 0, src->c[ci].p + (size_t)yc * src->c[ci].stride + xc, src->c[ci].stride,
                        pred[ci] + (ty / 2) * 64 + tx / 2, 64, rec->c[ci].p + (size_t)yc * rec->c[ci].stride + xc, rec->c[ci].stride,
                        lv->c[ci] + (size_t)yc * (W / 2) + xc, W / 2, 0, NULL)) f |= ci == 1 ? KS_F_CBF_CB : KS_F_CBF_CR;
        }
        for (int yy = y; yy < y + T; yy += 16) for (int xx = x; xx < x + T; xx += 16) {
            ks_cell *c = &cells[(yy >> 4) * cw + (xx >> 4)];
            c->mvx = (int16_t)mvx; c->mvy = (int16_t)mvy; c->cu_log2 = (uint8_t)log2; c->flags = (uint8_t)f; c->intra_mode = 0; c->rsv = 0;
        }
    }
}

/* ---- mode decision of a P picture (reference: processTree E@0x46b610 / checkInterPu2Nx2N / skipFullMergeDecision E@0x47f720 /
 * GetMergeCandsForP -- closed; this is OUR algorithm, designed so that every heavy step is independent per cell):
 *   stage E  per CTU: one candidate list = the search results of the CTU's own cells (z-order), zero, and of the cells bordering it on
 *            the left / above (distinct vectors, <= ORA_NCAND); per cell the distortion (search metric) of EVERY list entry;
 *   stage D  per CTU, in coding order: 64 -> 32 -> 16 quadtree by J = distortion + lambda * bits, bits from the 2Nx2N merge list /
 *            AMVP predictors of the vectors decided so far (cells of other CTUs count with their search results). ---- */
#define ORA_NCAND 16
#define ORA_INTRA_HDR_BITS 10
#define ORA_INTRA_FLOOR 64          /* x lambda (SAD domain): ~23 x Qstep per cell, i.e. a mean absolute error of ~0.09 Qstep */
typedef struct { int n; int16_t mvx[ORA_NCAND], mvy[ORA_NCAND]; int dist[16][ORA_NCAND]; int intra[16]; } ora_ctu_cands;   /* dist[j * 4 + i][k]; intra[j * 4 + i] */

static int mvd_bits_est(int d) { int a = iabs(d); if (a == 0) return 1; if (a == 1) return 3; int v = a - 2, k = 1, b = 3; while (v >= (1 << k)) { v -= 1 << k; k++; b++; } return b + k + 1; }

static void cand_add(ora_ctu_cands *t, const ks_cell *mv0, int cw, int ch, int cx, int cy, int zero)
{
    if (!zero && (cx < 0 || cy < 0 || cx >= cw || cy >= ch)) return;
    int mvx = zero ? 0 : mv0[cy * cw + cx].mvx, mvy = zero ? 0 : mv0[cy * cw + cx].mvy;
    for (int k = 0; k < t->n; k++) if (t->mvx[k] == mvx && t->mvy[k] == mvy) return;
    if (t->n >= ORA_NCAND) return;
    t->mvx[t->n] = (int16_t)mvx; t->mvy[t->n] = (int16_t)mvy; t->n++;
}
static void decide_candidates(const ora_cfg *cfg, const ora_plane *src, const ora_plane *ref, const ks_cell *mv0, const int *dist0, int X, int Y, ora_ctu_cands *t)
{   /* (X,Y) = cell coordinates of the CTU */
    int cw = cfg->width >> 4, ch = cfg->height >> 4;
    t->n = 0;
    for (int z = 0; z < 16; z++) cand_add(t, mv0, cw, ch, X + ((z & 1) | ((z >> 1) & 2)), Y + (((z >> 1) & 1) | ((z >> 2) & 2)), 0);
    cand_add(t, mv0, cw, ch, 0, 0, 1);
    for (int b = 3; b >= 0; b--) cand_add(t, mv0, cw, ch, X - 1, Y + b, 0);
    for (int a = 0; a < 4; a++) cand_add(t, mv0, cw, ch, X + a, Y - 1, 0);
    cand_add(t, mv0, cw, ch, X + 4, Y - 1, 0); cand_add(t, mv0, cw, ch, X - 1, Y + 4, 0); cand_add(t, mv0, cw, ch, X - 1, Y - 1, 0);
    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
        int cx = X + i, cy = Y + j;
        if (cx >= cw || cy >= ch) continue;
        const uint8_t *s0 = src->p + (size_t)(cy << 4) * src->stride + (cx << 4), *r0 = ref->p + (size_t)(cy << 4) * ref->stride + (cx << 4);
        t->intra[j * 4 + i] = intra_estimate(cfg, src, cx << 4, cy << 4, 16);
        for (int k = 0; k < t->n; k++) {
            if (t->mvx[k] == mv0[cy * cw + cx].mvx && t->mvy[k] == mv0[cy * cw + cx].mvy) { t->dist[j * 4 + i][k] = dist0[cy * cw + cx]; continue; }
            uint8_t pred[256];
            ora_mc_luma(pred, 16, r0, ref->stride, 16, 16, t->mvx[k], t->mvy[k]);
            t->dist[j * 4 + i][k] = (int)((cfg->satd && cfg->subpel > 0) ? ora_satd(s0, pred, src->stride, 16, 16, 16) : ora_sad(s0, pred, src->stride, 16, 16, 16));
        }
    }
}

/* OUR search for one 16x16 cell against an arbitrary reference picture (ora_replay.c runs it on the reference encoder's own reference pictures and
 * compares vector and distortion with what the reference chose): returns the winner's cost, *dist its distortion in the search metric */
int ora_me_probe(const ora_cfg *cfg, int qp, const ora_pic *src, const ora_pic *ref, int x0, int y0, int tpx, int tpy, int *mx, int *my, int *dist)
{
    return me_cell(cfg, ora_lambda_sad_q4[qp], &src->c[0], &ref->c[0], x0, y0, tpx, tpy, mx, my, dist);
}

/* stage D state of one CTU: vectors of the 4x4 cells plus a one-cell border (row -1: above CTUs incl. above-left / above-right,
 * column -1: left CTU); index [j + 1][i + 1], i = -1..4, j = -1..4 */
typedef struct {
    int16_t mvx[6][6], mvy[6][6];
    uint8_t ok[6][6];            /* the cell exists, precedes the current block in coding order and carries a vector */
    uint8_t log2[4][4], intra[4][4];
} ora_ctu_state;
static inline int zcell(int i, int j) { return (i & 1) | ((j & 1) << 1) | ((i & 2) << 1) | ((j & 2) << 2); }
static int nb_ok(const ora_ctu_state *st, int i, int j, int zcur)
{   /* (i,j) CTU-local cell coordinates of the neighbour; zcur = z-index of the current block's first cell */
    if (i < -1 || j < -1 || i > 4 || j > 3) return 0;
    if (!st->ok[j + 1][i + 1]) return 0;
    if (j == -1 || i == -1) return 1;
    if (i > 3) return 0;
    return zcell(i, j) < zcur;
}
/* bits to code vector (mx,my) for the 2Nx2N CU at local cell (i,j), size s cells: merge index, or AMVP + mvd (an estimate: the host codes the real thing) */
static int motion_bits(const ora_ctu_state *st, int i, int j, int s, int maxc, int mx, int my)
{
    int zc = zcell(i, j);
    const int ci[5] = {i - 1, i + s - 1, i + s, i - 1, i - 1}, cj[5] = {j + s - 1, j - 1, j - 1, j + s, j - 1};   /* A1 B1 B0 A0 B2 */
    int av[5], vx[5], vy[5];
    for (int k = 0; k < 5; k++) {
        av[k] = nb_ok(st, ci[k], cj[k], zc);
        vx[k] = av[k] ? st->mvx[cj[k] + 1][ci[k] + 1] : 0; vy[k] = av[k] ? st->mvy[cj[k] + 1][ci[k] + 1] : 0;
    }
#define SAME(a, b) (vx[a] == vx[b] && vy[a] == vy[b])
    int use[5] = {av[0], av[1] && !(av[0] && SAME(0, 1)), av[2] && !(av[1] && SAME(1, 2)), av[3] && !(av[0] && SAME(0, 3)),
                  av[4] && !(av[0] && SAME(0, 4)) && !(av[1] && SAME(1, 4))};
    if (use[0] + use[1] + use[2] + use[3] == 4) use[4] = 0;
#undef SAME
    int n = 0, idx = -1;
    for (int k = 0; k < 5 && n < maxc && idx < 0; k++) if (use[k]) { if (vx[k] == mx && vy[k] == my) idx = n; n++; }
    if (idx < 0 && n < maxc && mx == 0 && my == 0) idx = n;
    if (idx >= 0) return 1 + (maxc > 1 ? (idx < maxc - 1 ? idx + 1 : maxc - 1) : 0);
    /* AMVP: a = first of (A0, A1), b = first of (B0, B1, B2) */
    int fa = av[3] ? 3 : (av[0] ? 0 : -1), fb = av[2] ? 2 : (av[1] ? 1 : (av[4] ? 4 : -1));
    int best = 0x7fffffff;
    if (fa >= 0) best = imin(best, mvd_bits_est(mx - vx[fa]) + mvd_bits_est(my - vy[fa]));
    if (fb >= 0) best = imin(best, mvd_bits_est(mx - vx[fb]) + mvd_bits_est(my - vy[fb]));
    if (fa < 0 || fb < 0 || (vx[fa] == vx[fb] && vy[fa] == vy[fb])) best = imin(best, mvd_bits_est(mx) + mvd_bits_est(my));
    return 5 + best;
}
static int decide_block(const ora_ctu_cands *t, int ncx, int ncy, int lam, int maxc, ora_ctu_state *st, int i, int j, int s)
{   /* ncx/ncy = cells of the CTU inside the picture */
    if (i >= ncx || j >= ncy) return 0;
    int inside = i + s <= ncx && j + s <= ncy, jsplit = 0;
    if (s > 1) {
        int h = s >> 1;
        jsplit = inside ? (lam >> 4) : 0;          /* split_cu_flag = 1 */
        for (int k = 0; k < 4; k++) jsplit += decide_block(t, ncx, ncy, lam, maxc, st, i + (k & 1) * h, j + (k >> 1) * h, h);
        if (!inside) return jsplit;
    }
    int best = 0x7fffffff, bk = 0;
    for (int k = 0; k < t->n; k++) {
        int sum = 0;
        for (int b = 0; b < s; b++) for (int a = 0; a < s; a++) sum += t->dist[(j + b) * 4 + i + a][k];
        int c = sum + ((lam * (motion_bits(st, i, j, s, maxc, t->mvx[k], t->mvy[k]) + (s > 1))) >> 4);     /* + split_cu_flag = 0 above the minimum size */
        if (c < best) { best = c; bk = k; }
    }
    if (s > 1 && best > jsplit) return jsplit;
    if (s == 1) {       /* intra 16x16 CU (reference: intra CUs in P slices are a third of the CUs on natural clips [probe]): ~10 bits of header */
        int ji = t->intra[j * 4 + i] + ((lam * ORA_INTRA_HDR_BITS) >> 4);
        /* ...and only above the quantisation-noise floor: the estimate predicts from SOURCE neighbours while the inter distortion carries the
         * reference picture's coding noise, so on static smooth areas intra "won" cells the reference simply skips (720p natural: -4.2 % bits, same PSNR) */
        if (ji < best && t->dist[j * 4 + i][bk] > ((ORA_INTRA_FLOOR * lam) >> 4)) {
            st->mvx[j + 1][i + 1] = 0; st->mvy[j + 1][i + 1] = 0; st->ok[j + 1][i + 1] = 0; st->log2[j][i] = 4; st->intra[j][i] = 1;
            return ji;
        }
    }
    for (int b = 0; b < s; b++) for (int a = 0; a < s; a++) { st->intra[j + b][i + a] = 0;
        st->mvx[j + b + 1][i + a + 1] = t->mvx[bk]; st->mvy[j + b + 1][i + a + 1] = t->mvy[bk]; st->ok[j + b + 1][i + a + 1] = 1;
        st->log2[j + b][i + a] = (uint8_t)(s == 4 ? 6 : (s == 2 ? 5 : 4));
    }
    return best;
}
static void decide_ctu(const ora_cfg *cfg, int lam, int maxc, const ks_cell *mv0, const ora_ctu_cands *t, int X, int Y, ks_cell *cells)
{
    int cw = cfg->width >> 4, ch = cfg->height >> 4;
    ora_ctu_state st; memset(&st, 0, sizeof(st));
    for (int j = -1; j <= 4; j++) for (int i = -1; i <= 4; i++) {
        int cx = X + i, cy = Y + j;
        if (cx < 0 || cy < 0 || cx >= cw || cy >= ch) continue;
        if (j == -1 || (i == -1 && j <= 3)) { st.ok[j + 1][i + 1] = 1; st.mvx[j + 1][i + 1] = mv0[cy * cw + cx].mvx; st.mvy[j + 1][i + 1] = mv0[cy * cw + cx].mvy; }
    }
    int ncx = imin(4, cw - X), ncy = imin(4, ch - Y);
    decide_block(t, ncx, ncy, lam, maxc, &st, 0, 0, 4);
    for (int j = 0; j < ncy; j++) for (int i = 0; i < ncx; i++) {
        ks_cell *c = &cells[(Y + j) * cw + X + i];
        memset(c, 0, sizeof(*c)); c->mvx = st.mvx[j + 1][i + 1]; c->mvy = st.mvy[j + 1][i + 1]; c->cu_log2 = st.log2[j][i];
        if (st.intra[j][i]) c->flags = KS_F_INTRA;
    }
}

/* motion search of every 16x16 cell (independent of neighbours: predictor = co-located MV of the previous picture); returns the cost sum */
uint64_t ora_me_field(const ora_cfg *cfg, int qp, const ora_pic *src, const ora_pic *ref, const ks_cell *prev_cells, ks_cell *cells, int *dist)
{
    int W = cfg->width, H = cfg->height, cw = W >> 4, ch = H >> 4, lam = ora_lambda_sad_q4[qp];
    uint64_t cost_sum = 0;
    for (int cy = 0; cy < ch; cy++) for (int cx = 0; cx < cw; cx++) {
        int tpx = 0, tpy = 0, mx, my, d;
        if (prev_cells && !(prev_cells[cy * cw + cx].flags & KS_F_INTRA)) { tpx = prev_cells[cy * cw + cx].mvx; tpy = prev_cells[cy * cw + cx].mvy; }
        cost_sum += (uint64_t)me_cell(cfg, lam, &src->c[0], &ref->c[0], cx << 4, cy << 4, tpx, tpy, &mx, &my, &d);
        ks_cell *c = &cells[cy * cw + cx];
        memset(c, 0, sizeof(*c)); c->mvx = (int16_t)mx; c->mvy = (int16_t)my; c->cu_log2 = 4;
        if (dist) dist[cy * cw + cx] = d;
    }
    return cost_sum;
}

/* lambda_qp: the QP whose lambda drives the mode decision and the RD zero-out (>= qp; the encoder raises it on the non-key P pictures) */
uint64_t ora_inter_picture(const ora_cfg *cfg, int qp, int lambda_qp, const ora_pic *src, const ora_pic *ref, const ks_cell *prev_cells,
                       ora_pic *rec, ks_cell *cells, ora_levels *lv)
{
    build_scans();
    int W = cfg->width, H = cfg->height, cw = W >> 4, ch = H >> 4;
    lambda_qp = clip3(0, 51, lambda_qp);
    int lam = ora_lambda_sad_q4[lambda_qp], rdz = ora_lambda_sse_q4[lambda_qp], qpc = ora_chroma_qp[qp];
    /* 1. motion search per 16x16 cell */
    ks_cell *mv0 = malloc(sizeof(ks_cell) * (size_t)cw * ch);
    int *dist0 = malloc(sizeof(int) * (size_t)cw * ch);
    uint64_t cost_sum = ora_me_field(cfg, qp, src, ref, prev_cells, mv0, dist0);
    /* 2. candidate distortions, then the CU quadtree / merge decision, CTU by CTU */
    for (int Y = 0; Y < ch; Y += 4) for (int X = 0; X < cw; X += 4) {
        ora_ctu_cands t;
        decide_candidates(cfg, &src->c[0], &ref->c[0], mv0, dist0, X, Y, &t);
        decide_ctu(cfg, lam, getenv("MAXC") ? atoi(getenv("MAXC")) : 3, mv0, &t, X, Y, cells);
    }
    free(mv0); free(dist0);
    /* 3. prediction + residual coding + reconstruction per CU */
    for (int y = 0; y < H; y += 16) for (int x = 0; x < W; x += 16) {
        ks_cell c = cells[(y >> 4) * cw + (x >> 4)];
        int S = 1 << c.cu_log2;
        if ((x & (S - 1)) || (y & (S - 1)) || (c.flags & KS_F_INTRA)) continue;
        recon_inter_cu(cfg, qp, qpc, rdz, src, ref, rec, cells, lv, x, y, c.cu_log2, c.mvx, c.mvy);
    }
    /* 4. intra CUs, in coding order: their neighbours (the inter CUs of step 3, earlier intra CUs) are reconstructed */
    for (int cty = 0; cty < (H + 63) >> 6; cty++) for (int ctx = 0; ctx < (W + 63) >> 6; ctx++)
        for (int z = 0; z < 16; z++) {
            int x0 = (ctx << 6) + (((z & 1) | ((z >> 1) & 2)) << 4), y0 = (cty << 6) + ((((z >> 1) & 1) | ((z >> 2) & 2)) << 4);
            if (x0 >= W || y0 >= H || !(cells[(y0 >> 4) * cw + (x0 >> 4)].flags & KS_F_INTRA)) continue;
            intra_cell(cfg, qp, 0, src, rec, cells, lv, x0, y0);
        }
    return cost_sum;
}

/* ------------------------------------------------------------------ B picture (a7 bi-prediction) - */
static int scale_pred(int mv, int num, int den) { return den ? (mv * num) / den : 0; }       /* C division: truncates toward zero */

/* residual coding of one CU whose prediction is already assembled in pred[3] (pitch 64) */
static void recon_cu_from_pred(const ora_cfg *cfg, int qp, int qpc, int rdz, const ora_pic *src, uint8_t (*pred)[64 * 64], ora_pic *rec,
                               ks_cell *cells, ora_levels *lv, int x0, int y0, int log2)
{
    int S = 1 << log2, W = cfg->width, cw = W >> 4;
    int T = imin(S, 32), tl = log2 > 5 ? 5 : log2;
    for (int ty = 0; ty < S; ty += T) for (int tx = 0; tx < S; tx += T) {
        int x = x0 + tx, y = y0 + ty, f = 0;
        if (code_tb(cfg, qp, 0, tl, 0, src->c[0].p + (size_t)y * src->c[0].stride + x, src->c[0].stride, pred[0] + ty * 64 + tx, 64,
                    rec->c[0].p + (size_t)y * rec->c[0].stride + x, rec->c[0].stride, lv->c[0] + (size_t)y * W + x, W, rdz, NULL)) f |= KS_F_CBF_Y;
        for (int ci = 1; ci < 3; ci++) {
            int xc = x / 2, yc = y / 2;
            if (code_tb(cfg, qpc, 0, tl - 1, 0, src->c[ci].p + (size_t)yc * src->c[ci].stride + xc, src->c[ci].stride,
                        pred[ci] + (ty / 2) * 64 + tx / 2, 64, rec->c[ci].p + (size_t)yc * rec->c[ci].stride + xc, rec->c[ci].stride,
                        lv->c[ci] + (size_t)yc * (W / 2) + xc, W / 2, 0, NULL)) f |= ci == 1 ? KS_F_CBF_CB : KS_F_CBF_CR;
        }
        for (int yy = y; yy < y + T; yy += 16) for (int xx = x; xx < x + T; xx += 16) {
            ks_cell *c = &cells[(yy >> 4) * cw + (xx >> 4)];
            c->cu_log2 = (uint8_t)log2; c->flags = (uint8_t)f; c->intra_mode = 0; c->rsv = 0;
        }
    }
}
static void predict_cell(const ora_pic *ref0, const ora_pic *ref1, int x0, int y0, int n, int dir, int mx0, int my0, int mx1, int my1,
                         uint8_t (*pred)[64 * 64], int px, int py)
{   /* n x n luma (+ n/2 chroma) prediction of a block at picture (x0,y0), written at (px,py) of the CU buffers */
    for (int ci = 0; ci < 3; ci++) {
        int sh = ci ? 1 : 0, w = n >> sh, bx = x0 >> sh, by = y0 >> sh;
        uint8_t *d = pred[ci] + (py >> sh) * 64 + (px >> sh);
        const ora_plane *r0 = &ref0->c[ci], *r1 = &ref1->c[ci];
        if (dir == 3) {
            int16_t a[64 * 64], b[64 * 64];
            if (ci == 0) { ora_mc_luma_16(a, 64, r0->p + (size_t)by * r0->stride + bx, r0->stride, w, w, mx0, my0); ora_mc_luma_16(b, 64, r1->p + (size_t)by * r1->stride + bx, r1->stride, w, w, mx1, my1); }
            else { ora_mc_chroma_16(a, 64, r0->p + (size_t)by * r0->stride + bx, r0->stride, w, w, mx0, my0); ora_mc_chroma_16(b, 64, r1->p + (size_t)by * r1->stride + bx, r1->stride, w, w, mx1, my1); }
            ora_weighted_bi(d, 64, a, b, 64, w, w);
        } else {
            const ora_plane *r = dir == 1 ? r0 : r1; int mx = dir == 1 ? mx0 : mx1, my = dir == 1 ? my0 : my1;
            if (ci == 0) ora_mc_luma(d, 64, r->p + (size_t)by * r->stride + bx, r->stride, w, w, mx, my);
            else ora_mc_chroma(d, 64, r->p + (size_t)by * r->stride + bx, r->stride, w, w, mx, my);
        }
    }
}
/* B picture between two anchors: list 0 = ref0 (earlier), list 1 = ref1 (later).  anchor_cells = motion field of the later
 * anchor (a P picture predicted from ref0 over `da` pictures); d0 = POC(cur) - POC(ref0).  Predictors: the anchor's vector
 * scaled to each list.  Per cell: best of list 0 / list 1 / bi-prediction (DefaultWeightedBi_c) by SAD(or SATD)+lambda*bits. */
void ora_b_picture(const ora_cfg *cfg, int qp, int lambda_qp, const ora_pic *src, const ora_pic *ref0, const ora_pic *ref1, const ks_cell *anchor_cells,
                   int d0, int da, ora_pic *rec, ks_cell *cells, ks_cell_b *cells_b, ora_levels *lv)
{
    build_scans();
    int W = cfg->width, H = cfg->height, cw = W >> 4, ch = H >> 4;
    int lam = ora_lambda_sad_q4[qp], qpc = ora_chroma_qp[qp];
    static uint8_t pred[3][64 * 64];
    for (int cy = 0; cy < ch; cy++) for (int cx = 0; cx < cw; cx++) {
        int ax = 0, ay = 0;
        if (anchor_cells && !(anchor_cells[cy * cw + cx].flags & KS_F_INTRA)) { ax = anchor_cells[cy * cw + cx].mvx; ay = anchor_cells[cy * cw + cx].mvy; }
        int t0x = scale_pred(ax, d0, da), t0y = scale_pred(ay, d0, da), t1x = scale_pred(ax, d0 - da, da), t1y = scale_pred(ay, d0 - da, da);
        int m0x, m0y, m1x, m1y;
        int c0 = me_cell(cfg, lam, &src->c[0], &ref0->c[0], cx << 4, cy << 4, t0x, t0y, &m0x, &m0y, NULL);
        int c1 = me_cell(cfg, lam, &src->c[0], &ref1->c[0], cx << 4, cy << 4, t1x, t1y, &m1x, &m1y, NULL);
        predict_cell(ref0, ref1, cx << 4, cy << 4, 16, 3, m0x, m0y, m1x, m1y, pred, 0, 0);
        const uint8_t *s = src->c[0].p + (size_t)(cy << 4) * src->c[0].stride + (cx << 4);
        int cb = (int)(cfg->satd && cfg->subpel > 0 ? ora_satd(s, pred[0], src->c[0].stride, 64, 16, 16) : ora_sad(s, pred[0], src->c[0].stride, 64, 16, 16))
                 + ((lam * (mvbits(m0x - t0x) + mvbits(m0y - t0y))) >> 4) + ((lam * (mvbits(m1x - t1x) + mvbits(m1y - t1y))) >> 4);
        int dir = 1, best = c0;
        if (c1 < best) { best = c1; dir = 2; }
        if (cb < best) { best = cb; dir = 3; }
        ks_cell *c = &cells[cy * cw + cx]; ks_cell_b *b = &cells_b[cy * cw + cx];
        memset(c, 0, sizeof(*c)); memset(b, 0, sizeof(*b));
        c->cu_log2 = 4; b->dir = (uint8_t)dir;
        if (dir & 1) { c->mvx = (int16_t)m0x; c->mvy = (int16_t)m0y; }
        if (dir & 2) { b->mvx1 = (int16_t)m1x; b->mvy1 = (int16_t)m1y; }
    }
#define SAMEM(i, j) (cells[i].mvx == cells[j].mvx && cells[i].mvy == cells[j].mvy && cells_b[i].dir == cells_b[j].dir && cells_b[i].mvx1 == cells_b[j].mvx1 && cells_b[i].mvy1 == cells_b[j].mvy1)
    for (int y = 0; y + 32 <= H; y += 32) for (int x = 0; x + 32 <= W; x += 32) {
        int a = (y >> 4) * cw + (x >> 4);
        if (SAMEM(a, a + 1) && SAMEM(a, a + cw) && SAMEM(a, a + cw + 1)) cells[a].cu_log2 = cells[a + 1].cu_log2 = cells[a + cw].cu_log2 = cells[a + cw + 1].cu_log2 = 5;
    }
    for (int y = 0; y + 64 <= H; y += 64) for (int x = 0; x + 64 <= W; x += 64) {
        int a = (y >> 4) * cw + (x >> 4), ok = 1;
        for (int j = 0; j < 4 && ok; j++) for (int i = 0; i < 4; i++) { int t = a + j * cw + i; if (cells[t].cu_log2 != 5 || !SAMEM(a, t)) { ok = 0; break; } }
        if (ok) for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) cells[a + j * cw + i].cu_log2 = 6;
    }
#undef SAMEM
    for (int y = 0; y < H; y += 16) for (int x = 0; x < W; x += 16) {
        int i = (y >> 4) * cw + (x >> 4);
        int log2 = cells[i].cu_log2, S = 1 << log2;
        if ((x & (S - 1)) || (y & (S - 1))) continue;
        predict_cell(ref0, ref1, x, y, S, cells_b[i].dir, cells[i].mvx, cells[i].mvy, cells_b[i].mvx1, cells_b[i].mvy1, pred, 0, 0);
        recon_cu_from_pred(cfg, qp, qpc, ora_lambda_sse_q4[clip3(0, 51, lambda_qp)], src, pred, rec, cells, lv, x, y, log2);
    }
}

/* ------------------------------------------------------------------ deblocking (a16) ------------- */
static int is_tu_edge(const ks_cell *p, const ks_cell *q, int xp, int yp, int xq, int yq, int pos)
{   /* pos = coordinate (x for vertical edges, y for horizontal) of the edge, a multiple of 16 */
    int sp = 1 << p->cu_log2, sq = 1 << q->cu_log2;
    int same = p->cu_log2 == q->cu_log2 && (xp & ~(sp - 1)) == (xq & ~(sq - 1)) && (yp & ~(sp - 1)) == (yq & ~(sq - 1));
    if (!same) return 1;
    return p->cu_log2 == 6 && (pos & 31) == 0;
}
static int edge_bs(const ks_cell *p, const ks_cell *q, const ks_cell_b *pb, const ks_cell_b *qb)
{   /* spec 8.7.2.4; with B pictures the two lists always name different pictures, so motion compares list by list */
    if ((p->flags | q->flags) & KS_F_INTRA) return 2;
    if ((p->flags | q->flags) & KS_F_CBF_Y) return 1;
    int dp = pb ? pb->dir : 1, dq = qb ? qb->dir : 1;
    if (dp != dq) return 1;                       /* different reference pictures or number of motion vectors */
    if ((dp & 1) && (iabs(p->mvx - q->mvx) >= 4 || iabs(p->mvy - q->mvy) >= 4)) return 1;
    if ((dp & 2) && (iabs(pb->mvx1 - qb->mvx1) >= 4 || iabs(pb->mvy1 - qb->mvy1) >= 4)) return 1;
    return 0;
}
void ora_deblock_picture_b(const ora_cfg *cfg, int qp, int beta_off, int tc_off, ora_pic *rec, const ks_cell *cells, const ks_cell_b *cells_b);
void ora_deblock_picture(const ora_cfg *cfg, int qp, int beta_off, int tc_off, ora_pic *rec, const ks_cell *cells)
{
    ora_deblock_picture_b(cfg, qp, beta_off, tc_off, rec, cells, NULL);
}
void ora_deblock_picture_b(const ora_cfg *cfg, int qp, int beta_off, int tc_off, ora_pic *rec, const ks_cell *cells, const ks_cell_b *cells_b)
{
    int W = cfg->width, H = cfg->height, cw = W >> 4;
    int beta = ora_beta_table[clip3(0, 51, qp + (beta_off << 1))];
    int qpc = ora_chroma_qp[qp];
    for (int dir = 0; dir < 2; dir++)
        for (int e = 8; e < (dir ? H : W); e += 8)
            for (int t = 0; t < (dir ? W : H); t += 4) {
                int xq = dir ? t : e, yq = dir ? e : t, xp = dir ? t : e - 1, yp = dir ? e - 1 : t;
                const ks_cell *p = &cells[(yp >> 4) * cw + (xp >> 4)], *q = &cells[(yq >> 4) * cw + (xq >> 4)];
                int bs;
                if (e & 8) {                        /* inside a cell: only the CU/TU boundaries of a cell split into four 8x8 intra CUs */
                    if (p->cu_log2 != 3) continue;
                    bs = 2;
                } else {
                    if (!is_tu_edge(p, q, xp, yp, xq, yq, e)) continue;
                    bs = edge_bs(p, q, cells_b ? &cells_b[(yp >> 4) * cw + (xp >> 4)] : NULL, cells_b ? &cells_b[(yq >> 4) * cw + (xq >> 4)] : NULL);
                    if (!bs) continue;
                }
                int tc = ora_tc_table[clip3(0, 53, qp + 2 * (bs - 1) + (tc_off << 1))];
                ora_plane *pl = &rec->c[0];
                ora_deblock_luma_seg(pl->p + (size_t)yq * pl->stride + xq, dir ? pl->stride : 1, dir ? 1 : pl->stride, beta, tc);
                if (bs == 2 && !(e & 8)) {          /* chroma edges lie on the 8-sample chroma grid = multiples of 16 luma samples */
                    int tcc = ora_tc_table[clip3(0, 53, qpc + 2 + (tc_off << 1))];
                    for (int ci = 1; ci < 3; ci++) {
                        ora_plane *pc = &rec->c[ci];
                        ora_deblock_chroma_seg(pc->p + (size_t)(yq >> 1) * pc->stride + (xq >> 1), dir ? pc->stride : 1, dir ? 1 : pc->stride, tcc, 2);
                    }
                }
            }
}

/* ------------------------------------------------------------------ SAO (a17..a19) --------------- */
static int sao_offset_rd(int sum, int cnt, int signc, int lam, int is_bo, int *best_o)
{   /* signc: +1 offsets must be >=0, -1 must be <=0, 0 free.  returns cost (delta SSE + lambda*bits), Q0 */
    *best_o = 0;
    if (!cnt) return 0;
    int o = sum >= 0 ? (sum + cnt / 2) / cnt : -((-sum + cnt / 2) / cnt);
    o = clip3(-7, 7, o);
    if ((signc > 0 && o < 0) || (signc < 0 && o > 0)) o = 0;
    int best = (lam * 1) >> 4;                 /* offset 0: one bin */
    int step = o > 0 ? -1 : 1;
    for (int v = o; v != 0; v += step) {
        int bits = iabs(v) + 1 + (is_bo ? 1 : 0);
        int c = cnt * v * v - 2 * v * sum + ((lam * bits) >> 4);
        if (c < best || (c == best && iabs(v) < iabs(*best_o))) { best = c; *best_o = v; }
    }
    return best;
}
void ora_sao_picture(const ora_cfg *cfg, int qp, const ora_pic *src, const ora_pic *deb, ora_pic *out, ks_ctu_syn *ctus)
{
    int W = cfg->width, H = cfg->height, ctus_w = (W + 63) >> 6, ctus_h = (H + 63) >> 6;
    int lam = ora_lambda_sse_q4[qp];
    for (int ry = 0; ry < ctus_h; ry++) for (int rx = 0; rx < ctus_w; rx++) {
        ks_ctu_syn *ct = &ctus[ry * ctus_w + rx];
        memset(ct->sao, 0, sizeof(ct->sao));
        ora_sao_stats st[3];
        for (int ci = 0; ci < 3; ci++) {
            int sh = ci ? 1 : 0, x0 = (rx << 6) >> sh, y0 = (ry << 6) >> sh, pw = W >> sh, ph = H >> sh;
            int w = imin(64 >> sh, pw - x0), h = imin(64 >> sh, ph - y0);
            ora_sao_stats_ctb(&st[ci], src->c[ci].p, src->c[ci].stride, deb->c[ci].p, deb->c[ci].stride, x0, y0, w, h, pw, ph, cfg->sao >= 4 ? 1 : 2, cfg->sao >= 4 ? 4 : 2);
        }
        for (int grp = 0; grp < 2; grp++) {
            int c0 = grp ? 1 : 0, c1 = grp ? 2 : 0;
            int best_cost = 0, best_type = 0, best_class = 0, best_off[3][4], best_band[3];
            memset(best_off, 0, sizeof(best_off)); memset(best_band, 0, sizeof(best_band));
            if (cfg->sao) {
                for (int k = 0; k < (cfg->sao >= 4 ? 4 : 2); k++) {           /* edge classes */
                    int total = (lam * 4) >> 4, off[3][4];
                    for (int ci = c0; ci <= c1; ci++)
                        for (int cat = 1; cat <= 4; cat++)
                            total += sao_offset_rd(st[ci].eo_sum[k][cat], st[ci].eo_cnt[k][cat], cat <= 2 ? 1 : -1, lam, 0, &off[ci][cat - 1]);
                    if (total < best_cost) { best_cost = total; best_type = 2; best_class = k; for (int ci = c0; ci <= c1; ci++) memcpy(best_off[ci], off[ci], sizeof(off[ci])); }
                }
                {                                       /* band offset: best 4 consecutive bands per component */
                    int total = (lam * 7) >> 4, off[3][4], band[3];
                    for (int ci = c0; ci <= c1; ci++) {
                        int bc[32], bo[32], bestc = 0x7fffffff, bs = 0;
                        for (int b = 0; b < 32; b++) bc[b] = sao_offset_rd(st[ci].bo_sum[b], st[ci].bo_cnt[b], 0, lam, 1, &bo[b]);
                        for (int s = 0; s <= 28; s++) { int c = bc[s] + bc[s + 1] + bc[s + 2] + bc[s + 3]; if (c < bestc) { bestc = c; bs = s; } }
                        total += bestc; band[ci] = bs;
                        for (int j = 0; j < 4; j++) off[ci][j] = bo[bs + j];
                    }
                    if (total < best_cost) { best_cost = total; best_type = 1; for (int ci = c0; ci <= c1; ci++) { memcpy(best_off[ci], off[ci], sizeof(off[ci])); best_band[ci] = band[ci]; } }
                }
                int nz = 0;
                for (int ci = c0; ci <= c1; ci++) for (int j = 0; j < 4; j++) nz |= best_off[ci][j];
                if (!nz) best_type = 0;
            }
            for (int ci = c0; ci <= c1; ci++) {
                ks_sao_param *p = &ct->sao[ci];
                p->type = (uint8_t)best_type;
                if (best_type) {
                    p->band_or_class = (uint8_t)(best_type == 2 ? best_class : best_band[ci]);
                    for (int j = 0; j < 4; j++) p->off[j] = (int8_t)best_off[ci][j];
                }
            }
        }
        for (int ci = 0; ci < 3; ci++) {
            int sh = ci ? 1 : 0, x0 = (rx << 6) >> sh, y0 = (ry << 6) >> sh, pw = W >> sh, ph = H >> sh;
            int w = imin(64 >> sh, pw - x0), h = imin(64 >> sh, ph - y0);
            const ks_sao_param *p = &ct->sao[ci];
            ora_sao_apply_ctb(out->c[ci].p, out->c[ci].stride, deb->c[ci].p, deb->c[ci].stride, x0, y0, w, h, pw, ph, p->type, p->band_or_class, p->off);
        }
    }
}

/* ------------------------------------------------------------------ level packing ---------------- */
uint32_t ora_pack_levels(const ora_cfg *cfg, const ora_levels *lv, ks_ctu_syn *ctus, int16_t *pool)
{
    int W = cfg->width, H = cfg->height, ctus_w = (W + 63) >> 6, ctus_h = (H + 63) >> 6;
    uint32_t n = 0;
    for (int ry = 0; ry < ctus_h; ry++) for (int rx = 0; rx < ctus_w; rx++) {
        ks_ctu_syn *ct = &ctus[ry * ctus_w + rx];
        memset(ct->cg_y, 0, sizeof(ct->cg_y)); memset(ct->cg_cb, 0, 8); memset(ct->cg_cr, 0, 8);
        ct->cg_base = n; ct->rsv[0] = ct->rsv[1] = 0;
        for (int ci = 0; ci < 3; ci++) {
            int sh = ci ? 1 : 0, pw = W >> sh, ph = H >> sh, ng = 16 >> sh;
            for (int gy = 0; gy < ng; gy++) for (int gx = 0; gx < ng; gx++) {
                int x = ((rx << 6) >> sh) + gx * 4, y = ((ry << 6) >> sh) + gy * 4;
                if (x >= pw || y >= ph) continue;
                const int16_t *s = lv->c[ci] + (size_t)y * pw + x;
                int nz = 0;
                for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) nz |= s[j * pw + i];
                if (!nz) continue;
                if (ci == 0) ct->cg_y[gy] |= (uint16_t)(1u << gx); else if (ci == 1) ct->cg_cb[gy] |= (uint8_t)(1u << gx); else ct->cg_cr[gy] |= (uint8_t)(1u << gx);
                for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) pool[(size_t)n * 16 + j * 4 + i] = s[j * pw + i];
                n++;
            }
        }
    }
    return n;
}
