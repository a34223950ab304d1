/* ora_encoder.c -- CPU model of the whole encoder: oracle picture pipeline + the PRODUCT's host bitstream
 * writer (ks_bitstream.c).  TEST INFRASTRUCTURE: lets the closed-loop decoder test (SURVEY 8c tier P1) run
 * without a GPU, and gives the GPU tests a byte-exact bitstream + reconstruction to compare against. */
#include "ora_frame.h"
#include "ks_oracle.h"
#include "../ks265codec_b200/csrc/host/ks_bitstream.h"
#include <stdlib.h>
#include <string.h>

typedef struct ora_seq_cfg {
    int width, height;        /* display size */
    int nframes, qp, iper, fixqp;
    int me_range, me_iters, subpel, sign_hiding, sao, max_merge_cand, satd;
} ora_seq_cfg;

static void store_cropped(const ora_pic *p, int w, int h, uint8_t *dst)
{
    for (int ci = 0; ci < 3; ci++) {
        int pw = ci ? w / 2 : w, ph = ci ? h / 2 : h;
        for (int y = 0; y < ph; y++) { memcpy(dst, p->c[ci].p + (size_t)y * p->c[ci].stride, (size_t)pw); dst += pw; }
    }
}

/* returns bitstream bytes (or <0); recon_out receives nframes display-size I420 pictures (may be NULL).
 * If dump_* are non-NULL they receive, for the LAST frame, the cells / ctus / pool (for GPU-vs-oracle tests). */
long ora_encode_sequence(const ora_seq_cfg *sc, const uint8_t *yuv, uint8_t *bs, size_t bs_cap, uint8_t *recon_out)
{
    int W = (sc->width + 15) & ~15, H = (sc->height + 15) & ~15;
    ora_cfg cfg = {W, H, sc->me_range, sc->me_iters, sc->subpel, sc->sign_hiding, sc->sao, 1, sc->satd};
    ks_stream_params sp; memset(&sp, 0, sizeof(sp));
    sp.disp_width = sc->width; sp.disp_height = sc->height; sp.width = W; sp.height = H; sp.fps_num = 30; sp.fps_den = 1;
    sp.sign_hiding = sc->sign_hiding; sp.sao = sc->sao != 0; sp.max_merge_cand = sc->max_merge_cand;
    sp.pps_beta_offset_div2 = 2; sp.pps_tc_offset_div2 = 2; sp.strong_intra_smoothing = 1; sp.log2_max_poc_lsb = 8;
    int cw = W >> 4, ch = H >> 4, ctw = (W + 63) >> 6, cth = (H + 63) >> 6;
    size_t fsz = (size_t)sc->width * sc->height * 3 / 2;
    ora_pic src, pre, deb, fin[2];
    if (ora_pic_alloc(&src, W, H) || ora_pic_alloc(&pre, W, H) || ora_pic_alloc(&deb, W, H) || ora_pic_alloc(&fin[0], W, H) || ora_pic_alloc(&fin[1], W, H)) return -2;
    ks_cell *cells[2] = {calloc((size_t)cw * ch, sizeof(ks_cell)), calloc((size_t)cw * ch, sizeof(ks_cell))};
    ks_ctu_syn *ctus = calloc((size_t)ctw * cth, sizeof(ks_ctu_syn));
    ora_levels lv; lv.c[0] = calloc((size_t)W * H, 2); lv.c[1] = calloc((size_t)W * H / 4, 2); lv.c[2] = calloc((size_t)W * H / 4, 2);
    int16_t *pool = malloc((size_t)W * H * 3);
    void *scratch = malloc(ks_slice_scratch_bytes(&sp));
    long pos = 0, n;
    int poc = 0;
    for (int f = 0; f < sc->nframes; f++) {
        int is_i = (f % sc->iper) == 0;
        ora_pic_load(&src, yuv + fsz * f, sc->width, sc->height);
        int qp = is_i || sc->fixqp ? sc->qp : sc->qp + 1;
        if (qp > 51) qp = 51;
        ks_cell *cur = cells[f & 1], *prev = cells[(f & 1) ^ 1];
        ora_pic *out = &fin[f & 1], *ref = &fin[(f & 1) ^ 1];
        memset(lv.c[0], 0, (size_t)W * H * 2); memset(lv.c[1], 0, (size_t)W * H / 2); memset(lv.c[2], 0, (size_t)W * H / 2);
        if (is_i) { poc = 0; ora_intra_picture(&cfg, qp, &src, &pre, cur, &lv); }
        else ora_inter_picture(&cfg, qp, &src, ref, prev, &pre, cur, &lv);
        for (int ci = 0; ci < 3; ci++) memcpy(deb.c[ci].base, pre.c[ci].base, (size_t)pre.c[ci].stride * (pre.c[ci].h + 2 * ORA_PAD));
        int boff = is_i ? 0 : 2, toff = is_i ? 0 : 2;
        ora_deblock_picture(&cfg, qp, boff, toff, &deb, cur);
        ora_sao_picture(&cfg, qp, &src, &deb, out, ctus);
        ora_pic_extend(out);
        uint32_t ncg = ora_pack_levels(&cfg, &lv, ctus, pool);
        if (is_i) {
            if ((n = ks_write_vps(&sp, bs + pos, bs_cap - pos)) < 0) return -1; pos += n;
            if ((n = ks_write_sps(&sp, bs + pos, bs_cap - pos)) < 0) return -1; pos += n;
            if ((n = ks_write_pps(&sp, bs + pos, bs_cap - pos)) < 0) return -1; pos += n;
        }
        ks_frame_syn syn = {W, H, cw, ch, ctw, cth, is_i ? KS_SLICE_I : KS_SLICE_P, qp, poc, cur, ctus, pool, ncg};
        ks_slice_params sl; memset(&sl, 0, sizeof(sl));
        sl.nal_type = is_i ? 19 : 1; sl.slice_type = syn.slice_type; sl.poc = poc; sl.qp = qp;
        sl.num_neg_refs = is_i ? 0 : 1; sl.neg_delta_poc[0] = -1;
        sl.deblock_override = is_i; sl.beta_offset_div2 = boff; sl.tc_offset_div2 = toff;
        sl.sao_luma = sl.sao_chroma = sc->sao != 0;
        if ((n = ks_write_slice(&sp, &sl, &syn, scratch, bs + pos, bs_cap - pos)) < 0) return -1;
        pos += n;
        if (recon_out) store_cropped(out, sc->width, sc->height, recon_out + fsz * f);
        poc++;
    }
    ora_pic_free(&src); ora_pic_free(&pre); ora_pic_free(&deb); ora_pic_free(&fin[0]); ora_pic_free(&fin[1]);
    free(cells[0]); free(cells[1]); free(ctus); free(lv.c[0]); free(lv.c[1]); free(lv.c[2]); free(pool); free(scratch);
    return pos;
}
