/*
 * ks_encoder.c -- host-side encoder: drives the device hot path (ks265_gpu.h) picture by picture and entropy-codes
 * the returned frame syntax (ks_bitstream.c).  Reference counterpart: CHevcEncode::encodeFrame (E@0x4b5050) /
 * encodeOneFrame (E@0x4b4980) minus lookahead/rate control (north star: only -rc 0 is on the device path).
 * The device works on picture f+1 while this thread entropy-codes picture f.
 */
#include "ks265_enc.h"
#include "ks265_gpu.h"
#include "ks_bitstream.h"
#include "ks_ratecontrol.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct ks265_encoder {
    ks265_config cfg;
    ks_gpu_ctx *gpu;
    ks_stream_params sp;
    void *scratch;
    int W, H;
    ks265_pic_stat *pic_stats; int pic_stats_cap;
    ks265_read_fn read_fn; void *read_opaque;     /* encode_gop_cb: pictures are pulled into the context's pinned staging */
};

static const char *const k_presets[] = {"ultrafast", "superfast", "veryfast", "fast", "medium", "slow", "slower", "veryslow", "placebo"};
int ks265_preset_index(const char *name)
{
    for (int i = 0; i < 9; i++) if (!strcmp(name, k_presets[i])) return i;
    return -1;
}
int ks265_config_default_preset(ks265_config *cfg, const char *preset)
{
    int p = ks265_preset_index(preset ? preset : "veryfast");
    if (p < 0) return -1;
    int w = cfg->width, h = cfg->height;
    memset(cfg, 0, sizeof(*cfg));
    cfg->width = w; cfg->height = h; cfg->fps = 30.0; cfg->preset = p; cfg->rc = 0; cfg->qp = 27; cfg->iper = 128;
    cfg->sao = p <= 3 ? 3 : 4;       /* reference -sao per preset: 1,1,3,3,4,4,4,4; 1 and 3 behave alike here */
    cfg->sign_hiding = 1; cfg->me_range = 64;
    cfg->me_iters = p == 0 ? 8 : (p == 1 ? 12 : (p == 2 ? 16 : 32));
    cfg->subpel = p == 0 ? 1 : 2;
    cfg->satd = p >= 3;
    cfg->me = p >= 5 ? 1 : 0;        /* small diamond, the search the north star names (-me 1 = HEX, the reference's default up to medium); slow.. use HEX,
                                        the closest implemented search to the reference's UMH */
    cfg->crf = 24.0;
    return 0;
}

ks265_encoder *ks265_encoder_open(const ks265_config *cfg, int *err)
{
    int e = 0;
    ks265_encoder *enc = (ks265_encoder *)calloc(1, sizeof(*enc));
    if (!enc) { if (err) *err = -12; return NULL; }
    enc->cfg = *cfg;
    if (cfg->rc != 0 && cfg->rc != 3) { fprintf(stderr, "ks265: -rc %d is not implemented (0 = fixed QP, 3 = CRF)\n", cfg->rc); if (err) *err = -22; free(enc); return NULL; }
    ks_gpu_cfg g; memset(&g, 0, sizeof(g));
    g.me_range = cfg->me_range; g.me_iters = cfg->me_iters; g.subpel = cfg->subpel; g.sign_hiding = cfg->sign_hiding; g.sao = cfg->sao; g.satd = cfg->satd; g.me_method = cfg->me;
    g.strong_intra = 1; g.n_src_slots = 3; g.n_rec_slots = cfg->bframes ? 4 : 2; g.n_syn_slots = cfg->bframes ? 4 : 2;
    enc->gpu = ks_gpu_open(cfg->device, cfg->width, cfg->height, &g, &e);
    if (!enc->gpu) { if (err) *err = e; free(enc); return NULL; }
    ks_gpu_coded_size(enc->gpu, &enc->W, &enc->H);
    ks_stream_params *sp = &enc->sp;
    sp->disp_width = cfg->width; sp->disp_height = cfg->height; sp->width = enc->W; sp->height = enc->H;
    sp->fps_num = (int)(cfg->fps * 1000 + 0.5); sp->fps_den = 1000;
    sp->sign_hiding = cfg->sign_hiding; sp->sao = cfg->sao != 0; sp->max_merge_cand = 3;
    sp->pps_beta_offset_div2 = 2; sp->pps_tc_offset_div2 = 2; sp->strong_intra_smoothing = 1; sp->log2_max_poc_lsb = 8;
    sp->bframes = cfg->bframes;
    enc->scratch = malloc(ks_slice_scratch_bytes(sp));
    if (!enc->scratch) { ks_gpu_close(enc->gpu); free(enc); if (err) *err = -12; return NULL; }
    if (err) *err = 0;
    return enc;
}
void ks265_encoder_close(ks265_encoder *enc)
{
    if (!enc) return;
    ks_gpu_close(enc->gpu); free(enc->scratch); free(enc);
}

/* Coding schedule of one closed GOP shard of n pictures with bf B pictures between anchors: anchors at 0, bf+1, 2(bf+1), ...
 * and the last picture; each anchor is coded before the B pictures that precede it in display order (reference:
 * GopStructure::fillRpsInGop, EncGopStruct.cpp -- here non-hierarchical, non-reference B pictures).
 * order[i] = display index, type[i] = KS_SLICE_*, l0/l1 = display index of the list-0 / list-1 reference (-1 none). */
static int gop_schedule(int n, int bf, int *order, int *type, int *l0, int *l1)
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

typedef struct { ks_pic_params pp; int disp, type, l0, l1; } coded_pic;

/* device slots: anchors alternate reconstruction/syntax slots 0/1, B pictures alternate 2/3; sources rotate over 3 slots */
static void plan_pictures(const ks265_encoder *enc, int n, coded_pic *cp, int *count)
{
    const ks265_config *c = &enc->cfg;
    int *order = (int *)malloc(sizeof(int) * 4 * (size_t)(n + 1)), *type = order + n + 1, *l0 = type + n + 1, *l1 = l0 + n + 1;
    int cnt = gop_schedule(n, c->bframes, order, type, l0, l1);
    int anchors = 0, bs = 0, slot_prev = -1, slot_next = -1;
    for (int i = 0; i < cnt; i++) {
        ks_pic_params *pp = &cp[i].pp;
        memset(pp, 0, sizeof(*pp));
        cp[i].disp = order[i]; cp[i].type = type[i]; cp[i].l0 = l0[i]; cp[i].l1 = l1[i];
        pp->slice_type = type[i];
        pp->qp = c->qp;               /* the rate control sets the real value right before the picture is submitted */
        pp->want_me_cost = c->rc == 3 && type[i] == KS_SLICE_P;
        pp->src_slot = i % 3;
        pp->beta_offset_div2 = type[i] == KS_SLICE_I ? 0 : 2; pp->tc_offset_div2 = pp->beta_offset_div2;   /* reference: I slices override to 0/0, others PPS 2/2 */
        pp->want_sse = c->psnr;
        pp->ref_slot = pp->ref1_slot = pp->prev_syn_slot = -1;
        if (type[i] == KS_SLICE_B) {
            pp->out_slot = pp->syn_slot = 2 + (bs & 1); bs++;
            pp->ref_slot = slot_prev; pp->ref1_slot = slot_next; pp->prev_syn_slot = slot_next;
            pp->dist_l0 = order[i] - l0[i]; pp->dist_anchor = l1[i] - l0[i];
        } else {
            int slot = anchors & 1; anchors++;
            slot_prev = slot_next; slot_next = slot;
            pp->out_slot = pp->syn_slot = slot;
            if (type[i] == KS_SLICE_P) { pp->ref_slot = slot_prev; pp->prev_syn_slot = slot_prev; }
        }
    }
    *count = cnt;
    free(order);
}
static int upload(ks265_encoder *enc, int slot, int disp, const uint8_t *frames, const void *frames_dev)
{
    size_t fsz = (size_t)enc->cfg.width * enc->cfg.height * 3 / 2;
    int w = enc->cfg.width, h = enc->cfg.height;
    if (frames_dev) return ks_gpu_upload_frame_device(enc->gpu, slot, (const uint8_t *)frames_dev + fsz * disp);
    if (enc->read_fn) {
        uint8_t *d = ks_gpu_stage_acquire(enc->gpu);
        if (!d) return KS_ECUDA;
        if (enc->read_fn(enc->read_opaque, disp, d)) return -5;
        return ks_gpu_upload_staged(enc->gpu, slot);
    }
    const uint8_t *y = frames + fsz * disp;
    return ks_gpu_upload_frame(enc->gpu, slot, y, y + (size_t)w * h, y + (size_t)w * h * 5 / 4, w, w / 2);
}

/* bs == NULL: device pipeline only (no entropy coding) */
static long encode_gop_impl(ks265_encoder *enc, const uint8_t *frames, const void *frames_dev, int nframes,
                            uint8_t *bs, size_t cap, uint8_t *recon, ks265_gop_stats *stats)
{
    const ks_stream_params *sp = &enc->sp;
    size_t fsz = (size_t)enc->cfg.width * enc->cfg.height * 3 / 2;
    int w = enc->cfg.width, h = enc->cfg.height, r, cnt = 0;
    long pos = 0, n = 0;
    uint64_t l0c = ks_gpu_launch_count(enc->gpu), d0 = ks_gpu_d2h_bytes(enc->gpu), cg = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    coded_pic *cp = (coded_pic *)malloc(sizeof(coded_pic) * (size_t)(nframes + 1));
    if (!cp) return -12;
    plan_pictures(enc, nframes, cp, &cnt);
    ks_rc rc;                          /* one rate-control state per closed-GOP shard: shards stay independent */
    if (ks_rc_init(&rc, enc->cfg.rc, enc->cfg.qp, enc->cfg.fixqp, enc->cfg.crf, (enc->W >> 4) * (enc->H >> 4), enc->cfg.bframes)) { free(cp); return -22; }
    cp[0].pp.qp = ks_rc_picture_qp(&rc, cp[0].type, cp[0].disp);
    cp[0].pp.lambda_qp_delta = ks_rc_lambda_qp(&rc, cp[0].type, cp[0].disp, cp[0].pp.qp) - cp[0].pp.qp;
    if (bs) {
        if ((n = ks_write_vps(sp, bs + pos, cap - pos)) < 0) { free(cp); return -28; } pos += n;
        if ((n = ks_write_sps(sp, bs + pos, cap - pos)) < 0) { free(cp); return -28; } pos += n;
        if ((n = ks_write_pps(sp, bs + pos, cap - pos)) < 0) { free(cp); return -28; } pos += n;
    }
#define FAIL(code) do { ks_gpu_abort(enc->gpu); free(cp); return (code); } while (0)      /* pictures still in flight are dropped: the handle stays usable */
    if ((r = upload(enc, cp[0].pp.src_slot, cp[0].disp, frames, frames_dev))) FAIL(r);
    if ((r = ks_gpu_encode_picture_submit(enc->gpu, &cp[0].pp))) FAIL(r);
    for (int i = 0; i < cnt; i++) {
        /* keep the device busy: picture i+1 is uploaded and submitted before picture i is entropy-coded */
        ks_pic_out out;
        if (i + 1 < cnt && (r = upload(enc, cp[i + 1].pp.src_slot, cp[i + 1].disp, frames, frames_dev))) FAIL(r);
        if ((r = ks_gpu_encode_picture_finish(enc->gpu, cp[i].pp.syn_slot, &out))) FAIL(r);
        ks_rc_update(&rc, cp[i].type, out.me_cost);
        if (i + 1 < cnt) {
            cp[i + 1].pp.qp = ks_rc_picture_qp(&rc, cp[i + 1].type, cp[i + 1].disp);
            cp[i + 1].pp.lambda_qp_delta = ks_rc_lambda_qp(&rc, cp[i + 1].type, cp[i + 1].disp, cp[i + 1].pp.qp) - cp[i + 1].pp.qp;
            if ((r = ks_gpu_encode_picture_submit(enc->gpu, &cp[i + 1].pp))) FAIL(r);
        }
        cg += out.n_cg;
        if (recon) {      /* picture i's reconstruction slot is not rewritten before picture i+2 is submitted */
            uint8_t *y = recon + fsz * cp[i].disp;
            if ((r = ks_gpu_fetch_recon(enc->gpu, cp[i].pp.out_slot, y, y + (size_t)w * h, y + (size_t)w * h * 5 / 4, w, w / 2))) FAIL(r);
        }
        if (bs) {
            const ks_pic_params *p = &cp[i].pp;
            ks_frame_syn syn = {enc->W, enc->H, enc->W >> 4, enc->H >> 4, (enc->W + 63) >> 6, (enc->H + 63) >> 6, p->slice_type, p->qp, cp[i].disp,
                                out.cells, out.ctus, out.levels, out.n_cg, out.cells_b};
            ks_slice_params sl; memset(&sl, 0, sizeof(sl));
            sl.nal_type = cp[i].type == KS_SLICE_I ? 19 : (cp[i].type == KS_SLICE_P ? 1 : 0); sl.slice_type = p->slice_type; sl.poc = cp[i].disp; sl.qp = p->qp;
            if (cp[i].type != KS_SLICE_I) { sl.num_neg_refs = 1; sl.neg_delta_poc[0] = cp[i].l0 - cp[i].disp; }
            if (cp[i].type == KS_SLICE_B) { sl.num_pos_refs = 1; sl.pos_delta_poc[0] = cp[i].l1 - cp[i].disp; }
            sl.deblock_override = cp[i].type == KS_SLICE_I; sl.beta_offset_div2 = p->beta_offset_div2; sl.tc_offset_div2 = p->tc_offset_div2;
            sl.sao_luma = sl.sao_chroma = sp->sao;
            if ((n = ks_write_slice(sp, &sl, &syn, enc->scratch, bs + pos, cap - pos)) < 0) FAIL(-28);
            pos += n;
        }
        if (stats) { stats->sse[0] += out.sse[0]; stats->sse[1] += out.sse[1]; stats->sse[2] += out.sse[2]; }
        if (enc->pic_stats && i < enc->pic_stats_cap) {
            ks265_pic_stat *ps = &enc->pic_stats[i];
            ps->poc = cp[i].disp; ps->slice_type = cp[i].type; ps->qp = cp[i].pp.qp; ps->bits = bs ? (uint64_t)n * 8 : 0;
            ps->sse[0] = out.sse[0]; ps->sse[1] = out.sse[1]; ps->sse[2] = out.sse[2];
        }
    }
#undef FAIL
    free(cp);
    if (stats) {
        stats->frames = nframes; stats->bytes = bs ? (uint64_t)pos : cg * 32; stats->gpu_launches = ks_gpu_launch_count(enc->gpu) - l0c;
        stats->d2h_bytes = ks_gpu_d2h_bytes(enc->gpu) - d0; stats->h2d_bytes = frames_dev ? 0 : (uint64_t)fsz * nframes;
    }
    return bs ? pos : (long)nframes;
}

long ks265_encoder_encode_gop(ks265_encoder *enc, const uint8_t *frames, const void *frames_dev, int nframes,
                              uint8_t *bs, size_t cap, uint8_t *recon, ks265_gop_stats *stats)
{
    if (!enc || (!frames && !frames_dev) || nframes < 1 || !bs) return -22;
    return encode_gop_impl(enc, frames, frames_dev, nframes, bs, cap, recon, stats);
}

long ks265_encoder_encode_gop_cb(ks265_encoder *enc, ks265_read_fn read_picture, void *opaque, int nframes,
                                 uint8_t *bs, size_t cap, uint8_t *recon, ks265_gop_stats *stats)
{
    if (!enc || !read_picture || nframes < 1 || !bs) return -22;
    enc->read_fn = read_picture; enc->read_opaque = opaque;
    long r = encode_gop_impl(enc, NULL, NULL, nframes, bs, cap, recon, stats);
    enc->read_fn = NULL; enc->read_opaque = NULL;
    if (r >= 0 && stats) stats->h2d_bytes = (uint64_t)enc->cfg.width * enc->cfg.height * 3 / 2 * (uint64_t)nframes;
    return r;
}

long ks265_encoder_run_gop_device(ks265_encoder *enc, const void *frames_dev, int nframes, ks265_gop_stats *stats)
{
    if (!enc || !frames_dev || nframes < 1) return -22;
    return encode_gop_impl(enc, NULL, frames_dev, nframes, NULL, 0, NULL, stats);
}

long ks265_encoder_headers(ks265_encoder *enc, uint8_t *out, size_t cap)
{
    if (!enc || !out) return -22;
    long pos = 0, n;
    if ((n = ks_write_vps(&enc->sp, out + pos, cap - pos)) < 0) return -28; pos += n;
    if ((n = ks_write_sps(&enc->sp, out + pos, cap - pos)) < 0) return -28; pos += n;
    if ((n = ks_write_pps(&enc->sp, out + pos, cap - pos)) < 0) return -28; pos += n;
    return pos;
}
void ks265_encoder_set_picture_stats(ks265_encoder *enc, ks265_pic_stat *stats, int cap) { if (enc) { enc->pic_stats = stats; enc->pic_stats_cap = stats ? cap : 0; } }

int ks265_encoder_set_profiling(ks265_encoder *enc, int on) { return enc ? ks_gpu_set_profiling(enc->gpu, on) : -22; }
int ks265_encoder_get_stage_times(ks265_encoder *enc, double ms[KS_NSTAGES], uint64_t launches[KS_NSTAGES]) { return enc ? ks_gpu_get_stage_times(enc->gpu, ms, launches) : -22; }
