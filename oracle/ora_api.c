/* ora_api.c -- flat-buffer entry point of the CPU picture model for the GPU parity tests (TEST INFRASTRUCTURE).
 * All pictures are coded-size I420 (Y, U, V back to back, no padding). */
#include "ora_frame.h"
#include "ks_oracle.h"
#include <stdlib.h>
#include <string.h>

static void pic_from_flat(ora_pic *p, const uint8_t *flat, int W, int H)
{
    ora_pic_load(p, flat, W, H);
}
static void pic_to_flat(const ora_pic *p, uint8_t *flat)
{
    for (int ci = 0; ci < 3; ci++)
        for (int y = 0; y < p->c[ci].h; y++) { memcpy(flat, p->c[ci].p + (size_t)y * p->c[ci].stride, (size_t)p->c[ci].w); flat += p->c[ci].w; }
}
/* one picture through the model.  ref/prev_cells NULL for I pictures.  outputs (any may be NULL):
 * pre (pre-filter reconstruction), fin (final), cells, lev (dense levels Y,U,V), ctus, pool; returns n_cg or <0 */
long ora_run_picture(const ora_cfg *cfg, int slice_type, int qp, int lambda_qp, int beta_off, int tc_off,
                     const uint8_t *src, const uint8_t *ref, const ks_cell *prev_cells,
                     uint8_t *pre_out, uint8_t *fin_out, ks_cell *cells_out, int16_t *lev_out, ks_ctu_syn *ctus_out, int16_t *pool_out)
{
    int W = cfg->width, H = cfg->height, cw = W >> 4, ch = H >> 4, ctw = (W + 63) >> 6, cth = (H + 63) >> 6;
    ora_pic s, r, pre, deb, fin;
    if (ora_pic_alloc(&s, W, H) || ora_pic_alloc(&r, W, H) || ora_pic_alloc(&pre, W, H) || ora_pic_alloc(&deb, W, H) || ora_pic_alloc(&fin, W, H)) return -1;
    ks_cell *cells = calloc((size_t)cw * ch, sizeof(ks_cell));
    ks_ctu_syn *ctus = calloc((size_t)ctw * cth, sizeof(ks_ctu_syn));
    ora_levels lv; lv.c[0] = calloc((size_t)W * H, 2); lv.c[1] = calloc((size_t)W * H / 4, 2); lv.c[2] = calloc((size_t)W * H / 4, 2);
    int16_t *pool = malloc((size_t)W * H * 3);
    pic_from_flat(&s, src, W, H);
    if (slice_type == KS_SLICE_I) ora_intra_picture(cfg, qp, &s, &pre, cells, &lv);
    else { pic_from_flat(&r, ref, W, H); ora_inter_picture(cfg, qp, lambda_qp, &s, &r, prev_cells, &pre, cells, &lv); }
    for (int ci = 0; ci < 3; ci++) memcpy(deb.c[ci].base, pre.c[ci].base, (size_t)pre.c[ci].stride * (pre.c[ci].h + 2 * ORA_PAD));
    ora_deblock_picture(cfg, qp, beta_off, tc_off, &deb, cells);
    ora_sao_picture(cfg, qp, &s, &deb, &fin, ctus);
    uint32_t n = ora_pack_levels(cfg, &lv, ctus, pool);
    if (pre_out) pic_to_flat(&pre, pre_out);
    if (fin_out) pic_to_flat(&fin, fin_out);
    if (cells_out) memcpy(cells_out, cells, (size_t)cw * ch * sizeof(ks_cell));
    if (lev_out) { memcpy(lev_out, lv.c[0], (size_t)W * H * 2); memcpy(lev_out + (size_t)W * H, lv.c[1], (size_t)W * H / 2); memcpy(lev_out + (size_t)W * H * 5 / 4, lv.c[2], (size_t)W * H / 2); }
    if (ctus_out) memcpy(ctus_out, ctus, (size_t)ctw * cth * sizeof(ks_ctu_syn));
    if (pool_out) memcpy(pool_out, pool, (size_t)n * 32);
    ora_pic_free(&s); ora_pic_free(&r); ora_pic_free(&pre); ora_pic_free(&deb); ora_pic_free(&fin);
    free(cells); free(ctus); free(lv.c[0]); free(lv.c[1]); free(lv.c[2]); free(pool);
    return (long)n;
}
/* motion search only (for ks_gpu_debug_me) */
void ora_run_me(const ora_cfg *cfg, int qp, const uint8_t *src, const uint8_t *ref, const ks_cell *prev_cells, ks_cell *cells_out)
{
    int W = cfg->width, H = cfg->height;
    ora_pic s, r; ora_pic_alloc(&s, W, H); ora_pic_alloc(&r, W, H);
    pic_from_flat(&s, src, W, H); pic_from_flat(&r, ref, W, H);
    ora_me_field(cfg, qp, &s, &r, prev_cells, cells_out, NULL);
    ora_pic_free(&s); ora_pic_free(&r);
}

/* sizeof() of the configuration structs (binding self-check for the ctypes mirrors in tests/katlib.py): 0 ora_cfg, 1 ora_seq_cfg */
size_t ora_sizeof_seq_cfg(void);
size_t ora_abi_sizeof(int which) { return which == 0 ? sizeof(ora_cfg) : (which == 1 ? ora_sizeof_seq_cfg() : 0); }
