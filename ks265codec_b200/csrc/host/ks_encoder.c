/*
 * ks_encoder.c -- host-side encoder: drives the device hot path (ks265_gpu.h) picture by picture and entropy-codes
 * the returned frame syntax (ks_bitstream.c).  Reference counterpart: CHevcEncode::encodeFrame (E@0x4b5050) /
 * encodeOneFrame (E@0x4b4980) minus lookahead/rate control (north star: only -rc 0 is on the device path).
 * The device works on picture f+1 while this thread entropy-codes picture f.
 */
#include "ks265_enc.h"
#include "ks265_gpu.h"
#include "ks_bitstream.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct ks265_encoder {
    ks265_config cfg;
    ks_gpu_ctx *gpu;
    ks_stream_params sp;
    void *scratch;
    int W, H;
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
    return 0;
}

ks265_encoder *ks265_encoder_open(const ks265_config *cfg, int *err)
{
    int e = 0;
    ks265_encoder *enc = (ks265_encoder *)calloc(1, sizeof(*enc));
    if (!enc) { if (err) *err = -12; return NULL; }
    enc->cfg = *cfg;
    if (cfg->rc != 0) { fprintf(stderr, "ks265: only -rc 0 (fixed QP) is implemented on the device path\n"); if (err) *err = -22; free(enc); return NULL; }
    ks_gpu_cfg g; memset(&g, 0, sizeof(g));
    g.me_range = cfg->me_range; g.me_iters = cfg->me_iters; g.subpel = cfg->subpel; g.sign_hiding = cfg->sign_hiding; g.sao = cfg->sao; g.satd = cfg->satd;
    g.strong_intra = 1; g.n_src_slots = 3; g.n_rec_slots = 2; g.n_syn_slots = 2;
    enc->gpu = ks_gpu_open(cfg->device, cfg->width, cfg->height, &g, &e);
    if (!enc->gpu) { if (err) *err = e; free(enc); return NULL; }
    ks_gpu_coded_size(enc->gpu, &enc->W, &enc->H);
    ks_stream_params *sp = &enc->sp;
    sp->disp_width = cfg->width; sp->disp_height = cfg->height; sp->width = enc->W; sp->height = enc->H;
    sp->fps_num = (int)(cfg->fps * 1000 + 0.5); sp->fps_den = 1000;
    sp->sign_hiding = cfg->sign_hiding; sp->sao = cfg->sao != 0; sp->max_merge_cand = 3;
    sp->pps_beta_offset_div2 = 2; sp->pps_tc_offset_div2 = 2; sp->strong_intra_smoothing = 1; sp->log2_max_poc_lsb = 8;
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

static void pic_setup(const ks265_encoder *enc, int f, ks_pic_params *pp)
{
    const ks265_config *c = &enc->cfg;
    int is_i = f == 0;
    memset(pp, 0, sizeof(*pp));
    pp->slice_type = is_i ? KS_SLICE_I : KS_SLICE_P;
    pp->qp = (is_i || c->fixqp) ? c->qp : c->qp + 1;
    if (pp->qp > 51) pp->qp = 51;
    pp->src_slot = f % 3; pp->out_slot = f & 1; pp->ref_slot = is_i ? -1 : ((f & 1) ^ 1);
    pp->syn_slot = f & 1; pp->prev_syn_slot = is_i ? -1 : ((f & 1) ^ 1);
    pp->beta_offset_div2 = is_i ? 0 : 2; pp->tc_offset_div2 = is_i ? 0 : 2;      /* reference: I slices override to 0/0, P use PPS 2/2 */
    pp->want_sse = c->psnr;
}
static int upload(ks265_encoder *enc, int f, const uint8_t *frames, const void *frames_dev)
{
    size_t fsz = (size_t)enc->cfg.width * enc->cfg.height * 3 / 2;
    int w = enc->cfg.width, h = enc->cfg.height;
    if (frames_dev) return ks_gpu_upload_frame_device(enc->gpu, f % 3, (const uint8_t *)frames_dev + fsz * f);
    const uint8_t *y = frames + fsz * f;
    return ks_gpu_upload_frame(enc->gpu, f % 3, y, y + (size_t)w * h, y + (size_t)w * h * 5 / 4, w, w / 2);
}

long ks265_encoder_encode_gop(ks265_encoder *enc, const uint8_t *frames, const void *frames_dev, int nframes,
                              uint8_t *bs, size_t cap, uint8_t *recon, ks265_gop_stats *stats)
{
    if (!enc || (!frames && !frames_dev) || nframes < 1 || !bs) return -22;
    const ks_stream_params *sp = &enc->sp;
    size_t fsz = (size_t)enc->cfg.width * enc->cfg.height * 3 / 2;
    int w = enc->cfg.width, h = enc->cfg.height, r;
    long pos = 0, n;
    uint64_t l0 = ks_gpu_launch_count(enc->gpu), d0 = ks_gpu_d2h_bytes(enc->gpu);
    if (stats) memset(stats, 0, sizeof(*stats));
    if ((n = ks_write_vps(sp, bs + pos, cap - pos)) < 0) return -28; pos += n;
    if ((n = ks_write_sps(sp, bs + pos, cap - pos)) < 0) return -28; pos += n;
    if ((n = ks_write_pps(sp, bs + pos, cap - pos)) < 0) return -28; pos += n;
    ks_pic_params pp[2];
    if ((r = upload(enc, 0, frames, frames_dev))) return r;
    pic_setup(enc, 0, &pp[0]);
    if ((r = ks_gpu_encode_picture_submit(enc->gpu, &pp[0]))) return r;
    for (int f = 0; f < nframes; f++) {
        /* keep the device busy: upload + submit picture f+1 before entropy-coding picture f.
         * (recon fetch of f, when requested, must precede submit(f+1)? no: f+1 writes the OTHER recon slot) */
        ks_pic_out out;
        if (f + 1 < nframes) {
            if ((r = upload(enc, f + 1, frames, frames_dev))) return r;
            /* syntax slot (f+1)&1 == (f-1)&1 was finished in the previous iteration; recon slot (f+1)&1 held picture f-1 */
            pic_setup(enc, f + 1, &pp[(f + 1) & 1]);
        }
        if ((r = ks_gpu_encode_picture_finish(enc->gpu, f & 1, &out))) return r;
        if (f + 1 < nframes && (r = ks_gpu_encode_picture_submit(enc->gpu, &pp[(f + 1) & 1]))) return r;
        if (recon) {
            /* picture f's reconstruction stays valid until picture f+2 is submitted */
            uint8_t *y = recon + fsz * f;
            if ((r = ks_gpu_fetch_recon(enc->gpu, f & 1, y, y + (size_t)w * h, y + (size_t)w * h * 5 / 4, w, w / 2))) return r;
        }
        const ks_pic_params *p = &pp[f & 1];
        ks_frame_syn syn = {enc->W, enc->H, enc->W >> 4, enc->H >> 4, (enc->W + 63) >> 6, (enc->H + 63) >> 6, p->slice_type, p->qp, f,
                            out.cells, out.ctus, out.levels, out.n_cg};
        ks_slice_params sl; memset(&sl, 0, sizeof(sl));
        sl.nal_type = f == 0 ? 19 : 1; sl.slice_type = p->slice_type; sl.poc = f; sl.qp = p->qp;
        sl.num_neg_refs = f == 0 ? 0 : 1; sl.neg_delta_poc[0] = -1;
        sl.deblock_override = f == 0; sl.beta_offset_div2 = p->beta_offset_div2; sl.tc_offset_div2 = p->tc_offset_div2;
        sl.sao_luma = sl.sao_chroma = sp->sao;
        if ((n = ks_write_slice(sp, &sl, &syn, enc->scratch, bs + pos, cap - pos)) < 0) return -28;
        pos += n;
        if (stats) { stats->sse[0] += out.sse[0]; stats->sse[1] += out.sse[1]; stats->sse[2] += out.sse[2]; }
    }
    if (stats) { stats->frames = nframes; stats->bytes = (uint64_t)pos; stats->gpu_launches = ks_gpu_launch_count(enc->gpu) - l0;
                 stats->d2h_bytes = ks_gpu_d2h_bytes(enc->gpu) - d0; stats->h2d_bytes = frames_dev ? 0 : (uint64_t)fsz * nframes; }
    return pos;
}

long ks265_encoder_run_gop_device(ks265_encoder *enc, const void *frames_dev, int nframes, ks265_gop_stats *stats)
{
    if (!enc || !frames_dev || nframes < 1) return -22;
    int r;
    uint64_t l0 = ks_gpu_launch_count(enc->gpu), d0 = ks_gpu_d2h_bytes(enc->gpu);
    ks_pic_params pp;
    ks_pic_out out;
    uint64_t cg = 0;
    for (int f = 0; f < nframes; f++) {
        if ((r = upload(enc, f, NULL, frames_dev))) return r;
        pic_setup(enc, f, &pp);
        if ((r = ks_gpu_encode_picture_submit(enc->gpu, &pp))) return r;
        if (f > 0) { if ((r = ks_gpu_encode_picture_finish(enc->gpu, (f - 1) & 1, &out))) return r; cg += out.n_cg; }
    }
    if ((r = ks_gpu_encode_picture_finish(enc->gpu, (nframes - 1) & 1, &out))) return r;
    cg += out.n_cg;
    if (stats) { memset(stats, 0, sizeof(*stats)); stats->frames = nframes; stats->bytes = cg * 32; stats->gpu_launches = ks_gpu_launch_count(enc->gpu) - l0; stats->d2h_bytes = ks_gpu_d2h_bytes(enc->gpu) - d0; }
    return (long)nframes;
}

int ks265_encoder_set_profiling(ks265_encoder *enc, int on) { return enc ? ks_gpu_set_profiling(enc->gpu, on) : -22; }
int ks265_encoder_get_stage_times(ks265_encoder *enc, double ms[6], uint64_t launches[6]) { return enc ? ks_gpu_get_stage_times(enc->gpu, ms, launches) : -22; }
