/* ora_encoder.c -- CPU model of the whole encoder: oracle picture pipeline + the PRODUCT's host bitstream
 * writer (ks_bitstream.c).  TEST INFRASTRUCTURE: lets the closed-loop decoder test (SURVEY 8c tier P1) run
 * without a GPU, and gives the GPU tests a byte-exact bitstream + reconstruction to compare against. */
#include "ora_frame.h"
#include "ks_oracle.h"
#include "../ks265codec_b200/csrc/host/ks_bitstream.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "../ks265codec_b200/csrc/host/ks_ratecontrol.h"

typedef struct ora_seq_cfg {
    int width, height;        /* display size */
    int nframes, qp, iper, fixqp;
    int me_range, me_iters, subpel, sign_hiding, sao, max_merge_cand, satd;
    int bframes;              /* B pictures between anchors (0 = IPPP) */
    int me_method;            /* 0 diamond, 1 hexagon */
    int rc, crf_x100;         /* rc 0: fixed QP; rc 3: CRF (crf * 100), QP per picture from the host rate control (ks_ratecontrol.c) */
} ora_seq_cfg;

size_t ora_sizeof_seq_cfg(void) { return sizeof(ora_seq_cfg); }

static void store_cropped(const ora_pic *p, int w, int h, uint8_t *dst)
{
    for (int ci = 0; ci < 3; ci++) {
        int pw = ci ? w / 2 : w, ph = ci ? h / 2 : h;
        for (int y = 0; y < ph; y++) { memcpy(dst, p->c[ci].p + (size_t)y * p->c[ci].stride, (size_t)pw); dst += pw; }
    }
}

/* Coding schedule of one closed GOP shard of n pictures with `bf` B pictures between anchors (shared by the model and, in
 * ks_encoder.c, by the product): anchors at 0, bf+1, 2(bf+1), ... and the last picture; each anchor is coded before the
 * B pictures that precede it in display order.  order[i] = display index, type[i] = KS_SLICE_*, l0[i]/l1[i] = display
 * index of the list-0 / list-1 reference (-1 none). */
int ora_gop_schedule(int n, int bf, int *order, int *type, int *l0, int *l1)
{
    int k = 0, prev = 0;
    order[k] = 0; type[k] = KS_SLICE_I; l0[k] = l1[k] = -1; k++;
    while (prev < n - 1) {
        int next = prev + bf + 1 < n - 1 ? prev + bf + 1 : n - 1;
        order[k] = next; type[k] = KS_SLICE_P; l0[k] = prev; l1[k] = -1; k++;
        for (int b = prev + 1; b < next; b++) { order[k] = b; type[k] = KS_SLICE_B; l0[k] = prev; l1[k] = next; k++; }
        prev = next;
    }
    return k;
}

/* returns bitstream bytes (or <0); recon_out receives nframes display-size I420 pictures in DISPLAY order (may be NULL). */
long ora_encode_sequence(const ora_seq_cfg *sc, const uint8_t *yuv, uint8_t *bs, size_t bs_cap, uint8_t *recon_out)
{
    int W = (sc->width + 15) & ~15, H = (sc->height + 15) & ~15;
    ora_cfg cfg = {W, H, sc->me_range, sc->me_iters, sc->subpel, sc->sign_hiding, sc->sao, 1, sc->satd, sc->me_method};
    ks_stream_params sp; memset(&sp, 0, sizeof(sp));
    sp.disp_width = sc->width; sp.disp_height = sc->height; sp.width = W; sp.height = H; sp.fps_num = 30; sp.fps_den = 1;
    sp.sign_hiding = sc->sign_hiding; sp.sao = sc->sao != 0; sp.max_merge_cand = sc->max_merge_cand;
    sp.pps_beta_offset_div2 = 2; sp.pps_tc_offset_div2 = 2; sp.strong_intra_smoothing = 1; sp.log2_max_poc_lsb = 8;
    sp.bframes = sc->bframes;
    int cw = W >> 4, ch = H >> 4, ctw = (W + 63) >> 6, cth = (H + 63) >> 6;
    size_t fsz = (size_t)sc->width * sc->height * 3 / 2;
    ora_pic src, pre, deb, fin[3];
    if (ora_pic_alloc(&src, W, H) || ora_pic_alloc(&pre, W, H) || ora_pic_alloc(&deb, W, H) || ora_pic_alloc(&fin[0], W, H) || ora_pic_alloc(&fin[1], W, H) || ora_pic_alloc(&fin[2], W, H)) return -2;
    ks_cell *cells[3] = {calloc((size_t)cw * ch, sizeof(ks_cell)), calloc((size_t)cw * ch, sizeof(ks_cell)), calloc((size_t)cw * ch, sizeof(ks_cell))};
    ks_cell_b *cells_b = calloc((size_t)cw * ch, sizeof(ks_cell_b));
    ks_ctu_syn *ctus = calloc((size_t)ctw * cth, sizeof(ks_ctu_syn));
    ora_levels lv; lv.c[0] = calloc((size_t)W * H, 2); lv.c[1] = calloc((size_t)W * H / 4, 2); lv.c[2] = calloc((size_t)W * H / 4, 2);
    int16_t *pool = malloc((size_t)W * H * 3);
    void *scratch = malloc(ks_slice_scratch_bytes(&sp));
    int maxn = sc->iper < sc->nframes ? sc->iper : sc->nframes;
    int *order = malloc(sizeof(int) * 4 * (size_t)(maxn + 1)), *type = order + maxn + 1, *l0 = type + maxn + 1, *l1 = l0 + maxn + 1;
    long pos = 0, n;
    for (int g0 = 0; g0 < sc->nframes; g0 += sc->iper) {
        int gn = sc->nframes - g0 < sc->iper ? sc->nframes - g0 : sc->iper;
        int cnt = ora_gop_schedule(gn, sc->bframes, order, type, l0, l1);
        ks_rc rc;                 /* one rate-control state per closed-GOP shard, like the product */
        if (ks_rc_init(&rc, sc->rc, sc->qp, sc->fixqp, sc->crf_x100 / 100.0, cw * ch, sc->bframes)) return -3;
        /* anchors alternate between fin[0]/fin[1] (+cells[0]/[1]); B pictures reconstruct into fin[2] (+cells[2]) */
        int slot_of_prev = -1, slot_of_next = -1, anchor_idx = 0;
        for (int i = 0; i < cnt; i++) {
            int f = order[i], t = type[i];
            ora_pic_load(&src, yuv + fsz * (size_t)(g0 + f), sc->width, sc->height);
            int qp = ks_rc_picture_qp(&rc, t, f), lqp = ks_rc_lambda_qp(&rc, t, f, qp);
            uint64_t me_cost = 0;
            memset(lv.c[0], 0, (size_t)W * H * 2); memset(lv.c[1], 0, (size_t)W * H / 2); memset(lv.c[2], 0, (size_t)W * H / 2);
            int slot; ks_cell *cur; ora_pic *out;
            if (t == KS_SLICE_B) { slot = 2; }
            else { slot = anchor_idx & 1; anchor_idx++; slot_of_prev = slot_of_next; slot_of_next = slot; }
            cur = cells[slot]; out = &fin[slot];
            if (t == KS_SLICE_I) ora_intra_picture(&cfg, qp, &src, &pre, cur, &lv);
            else if (t == KS_SLICE_P) me_cost = ora_inter_picture(&cfg, qp, lqp, &src, &fin[slot_of_prev], slot_of_prev >= 0 && i > 1 ? cells[slot_of_prev] : cells[slot_of_prev], &pre, cur, &lv);
            else ora_b_picture(&cfg, qp, lqp, &src, &fin[slot_of_prev], &fin[slot_of_next], cells[slot_of_next], f - l0[i], l1[i] - l0[i], &pre, cur, cells_b, &lv);
            for (int ci = 0; ci < 3; ci++) memcpy(deb.c[ci].base, pre.c[ci].base, (size_t)pre.c[ci].stride * (pre.c[ci].h + 2 * ORA_PAD));
            int boff = t == KS_SLICE_I ? 0 : 2, toff = boff;
            ora_deblock_picture_b(&cfg, qp, boff, toff, &deb, cur, t == KS_SLICE_B ? cells_b : NULL);
            ora_sao_picture(&cfg, qp, &src, &deb, out, ctus);
            ora_pic_extend(out);
            uint32_t ncg = ora_pack_levels(&cfg, &lv, ctus, pool);
            if (t == KS_SLICE_I) {
                if ((n = ks_write_vps(&sp, bs + pos, bs_cap - pos)) < 0) return -1; pos += n;
                if ((n = ks_write_sps(&sp, bs + pos, bs_cap - pos)) < 0) return -1; pos += n;
                if ((n = ks_write_pps(&sp, bs + pos, bs_cap - pos)) < 0) return -1; pos += n;
            }
            ks_frame_syn syn = {W, H, cw, ch, ctw, cth, t, qp, f, cur, ctus, pool, ncg, t == KS_SLICE_B ? cells_b : NULL};
            ks_slice_params sl; memset(&sl, 0, sizeof(sl));
            sl.nal_type = t == KS_SLICE_I ? 19 : (t == KS_SLICE_P ? 1 : 0); sl.slice_type = t; sl.poc = f; sl.qp = qp;
            if (t != KS_SLICE_I) { sl.num_neg_refs = 1; sl.neg_delta_poc[0] = l0[i] - f; }
            if (t == KS_SLICE_B) { sl.num_pos_refs = 1; sl.pos_delta_poc[0] = l1[i] - f; }
            sl.deblock_override = t == KS_SLICE_I; sl.beta_offset_div2 = boff; sl.tc_offset_div2 = toff;
            sl.sao_luma = sl.sao_chroma = sc->sao != 0;
            if ((n = ks_write_slice(&sp, &sl, &syn, scratch, bs + pos, bs_cap - pos)) < 0) return -1;
            pos += n;
            if (recon_out) store_cropped(out, sc->width, sc->height, recon_out + fsz * (size_t)(g0 + f));
            if (getenv("ORA_RC_DEBUG")) fprintf(stderr, "poc %d type %d qp %d me_cost/cell %.1f bytes %ld\n", f, t, qp, (double)me_cost / (cw * ch), n);
            ks_rc_update(&rc, t, me_cost);
        }
    }
    ora_pic_free(&src); ora_pic_free(&pre); ora_pic_free(&deb); ora_pic_free(&fin[0]); ora_pic_free(&fin[1]); ora_pic_free(&fin[2]);
    free(cells[0]); free(cells[1]); free(cells[2]); free(cells_b); free(ctus); free(lv.c[0]); free(lv.c[1]); free(lv.c[2]); free(pool); free(scratch); free(order);
    return pos;
}
