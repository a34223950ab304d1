/*
 * ks_qyapi.c -- libks265qy.so: the reference's public encoder API (Android_demo/prebuilt/include/qy265enc.h:196-233, qy265def.h:178-196)
 * on top of the B200 encoder (ks265_enc.h).  See include/ks265_qyabi.h for the contract and the deviations.
 *
 * Shape: pictures handed to QY265EncoderEncodeFrame are copied into a page-locked shard buffer (the caller may reuse its planes at
 * once, as the reference's own demo does: encoderwrapper.c:366-379).  When the shard holds one intra period -- or on flush / key-frame
 * request -- it is encoded as a closed GOP on the device (ks265_encoder_encode_gop) and the Annex-B output is cut into access units;
 * every call then hands out at most one access unit, oldest first, so the queue drains exactly as fast as the next shard fills.
 */
#include "ks265_qyabi.h"
#include "ks265_enc.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

const char strLibQy265Version[] = "ks265codec_b200 (qy265enc.h ABI shim) 2.6.1.3-compatible";

static void (*g_log)(const char *) = NULL;
static void (*g_auth_warning)(void) = NULL;
void QY265SetLogPrintf(void (*fn)(const char *)) { g_log = fn; }
void QY265SetAuthWarning(void (*fn)(void)) { g_auth_warning = fn; (void)g_auth_warning; }
static void logf_(int level, int threshold, const char *fmt, ...)
{
    if (level < threshold) return;
    char msg[512];
    va_list ap; va_start(ap, fmt); vsnprintf(msg, sizeof(msg), fmt, ap); va_end(ap);
    if (g_log) g_log(msg); else fputs(msg, stderr);
}

static const char *const k_preset_names[] = {"ultrafast", "superfast", "veryfast", "fast", "medium", "slow", "slower", "veryslow", "placebo"};
static const char *const k_tune_names[] = {"default", "selfshow", "game", "movie", "screen"};
static const char *const k_latency_names[] = {"zerolatency", "lowdelay", "livestreaming", "default"};
static int name_index(const char *const *names, int n, const char *s)
{
    for (int i = 0; i < n; i++) if (!strcmp(names[i], s)) return i;
    return -1;
}

/* ---------------------------------------------------------------- configuration ---- */
int QY265ConfigDefault(ksqy_config *c, int preset, int tune, int latency)
{
    if (!c) return KSQY_POINTER;
    if (preset < 0 || preset > 8 || tune < 0 || tune > 4 || latency < 0 || latency > 3) return KSQY_FAIL;
    ks265_config k; memset(&k, 0, sizeof(k));
    if (ks265_config_default_preset(&k, k_preset_names[preset])) return KSQY_FAIL;
    memset(c, 0, sizeof(*c));
    c->tune = tune; c->preset = preset; c->latency = latency;
    c->profile_id = 1; c->headers_before_keyframe = 1; c->fps = 25.0; c->bframes = -1;
    c->rc = 2; c->bitrate_kbps = 512; c->qp = 26; c->crf = 24; c->visual_quality = 95; c->intra_period = 128; c->qp_max = 51;      /* header comments :75-87; the CLI echo */
    c->wavefront = 1; c->frame_parallel = 1;
    c->vui.video_format = 5; c->vui.primaries = c->vui.transfer = c->vui.matrix = 2;
    c->lookahead = -1; c->rate_tolerance = 2.0;
    c->me = k.me; c->do64 = 1; c->tu_inter = c->tu_intra = -1; c->smooth = 1; c->subme = k.subpel; c->satd_inter = k.satd; c->satd_intra = k.satd;
    c->search_range = k.me_range; c->ref_num = 1; c->sao = k.sao; c->aq_strength = 1.0; c->rasl = 1;
    return KSQY_OK;
}

int QY265ConfigDefaultPreset(ksqy_config *c, char *preset, char *tune, char *latency)
{
    const int p = preset ? name_index(k_preset_names, 9, preset) : 2;
    const int t = tune ? name_index(k_tune_names, 5, tune) : 0;
    const int l = latency ? name_index(k_latency_names, 4, latency) : 3;
    if (p < 0 || t < 0 || l < 0) return KSQY_FAIL;
    return QY265ConfigDefault(c, p, t, l);
}

/* names: the struct fields of qy265enc.h and the appencoder option names of the same knobs (README.md option table) */
int QY265ConfigParse(ksqy_config *c, const char *name, const char *value)
{
    if (!c || !name || !value) return KSQY_BAD_NAME;
    if (*name == '-') name++;
    if (!strcmp(name, "preset")) { int p = name_index(k_preset_names, 9, value); if (p < 0) return KSQY_BAD_VALUE; const ksqy_config keep = *c; QY265ConfigDefault(c, p, keep.tune, keep.latency);
        c->auth = keep.auth; c->width = keep.width; c->height = keep.height; c->fps = keep.fps; return 0; }
    if (!strcmp(name, "tune")) { int t = name_index(k_tune_names, 5, value); if (t < 0) return KSQY_BAD_VALUE; c->tune = t; return 0; }
    if (!strcmp(name, "latency")) { int l = name_index(k_latency_names, 4, value); if (l < 0) return KSQY_BAD_VALUE; c->latency = l; return 0; }
    if (!strcmp(name, "statFileName") || !strcmp(name, "stat")) { if (strlen(value) >= sizeof(c->stat_file)) return KSQY_BAD_VALUE; strcpy(c->stat_file, value); return 0; }
    char *end = NULL;
    const double d = strtod(value, &end);
    if (end == value || *end) return KSQY_BAD_VALUE;
    const int i = (int)d;
    static const struct { const char *field, *opt; size_t off; int is_double; } tab[] = {
#define F(field_, opt_, m_, dbl_) {field_, opt_, offsetof(ksqy_config, m_), dbl_}
        F("picWidth", "wdt", width, 0), F("picHeight", "hgt", height, 0), F("frameRate", "fr", fps, 1), F("bframes", "bframes", bframes, 0),
        F("profileId", "profile", profile_id, 0), F("bHeaderBeforeKeyframe", "hdrbeforekey", headers_before_keyframe, 0), F("temporalLayer", "temporallayer", temporal_layer, 0),
        F("rc", "rc", rc, 0), F("bitrateInkbps", "br", bitrate_kbps, 0), F("vbv_buffer_size", "vbvbuf", vbv_buffer_size, 0), F("vbv_max_rate", "vbvmax", vbv_max_rate, 0),
        F("vbv_min_rate", "vbvmin", vbv_min_rate, 0), F("qp", "qp", qp, 0), F("crf", "crf", crf, 0), F("visual_quality", "vq", visual_quality, 0), F("iIntraPeriod", "iper", intra_period, 0),
        F("qpmin", "qpmin", qp_min, 0), F("qpmax", "qpmax", qp_max, 0), F("enFrameSkip", "frameskip", frame_skip, 0), F("enWavefront", "wpp", wavefront, 0),
        F("enFrameParallel", "fpp", frame_parallel, 0), F("threads", "threads", threads, 0), F("logLevel", "v", log_level, 0), F("lookahead", "lookahead", lookahead, 0),
        F("calcPsnr", "psnr", calc_psnr, 0), F("calcSsim", "ssim", calc_ssim, 0), F("shortLoadingForPlayer", "shortload", short_loading, 0), F("iPass", "pass", pass, 0),
        F("fRateTolerance", "ratetol", rate_tolerance, 1), F("rdoq", "rdoq", rdoq, 0), F("me", "me", me, 0), F("part", "part", part, 0), F("do64", "do64", do64, 0),
        F("tuInter", "tuinter", tu_inter, 0), F("tuIntra", "tuintra", tu_intra, 0), F("smooth", "smooth", smooth, 0), F("transskip", "transskip", transskip, 0),
        F("subme", "subme", subme, 0), F("satdInter", "satdinter", satd_inter, 0), F("satdIntra", "satdintra", satd_intra, 0), F("searchrange", "sr", search_range, 0),
        F("refnum", "ref", ref_num, 0), F("ref0", "ref0", ref0, 0), F("sao", "sao", sao, 0), F("longTermRef", "ltr", long_term_ref, 0), F("iAqMode", "aqmode", aq_mode, 0),
        F("fAqStrength", "aqstrength", aq_strength, 1), F("rasl", "rasl", rasl, 0),
#undef F
    };
    for (size_t k = 0; k < sizeof(tab) / sizeof(tab[0]); k++)
        if (!strcmp(name, tab[k].field) || !strcmp(name, tab[k].opt)) {
            if (tab[k].is_double) *(double *)((char *)c + tab[k].off) = d; else *(int *)((char *)c + tab[k].off) = i;
            return 0;
        }
    return KSQY_BAD_NAME;
}

/* ---------------------------------------------------------------- the encoder handle ---- */
typedef struct au_rec { size_t off, len; int first_nal, nal_count, slice_type, poc; long long pts, dts; } au_rec;
typedef struct qy_enc {
    ksqy_config   cfg, pending;           /* `pending` replaces `cfg` at the next shard start after a Reconfig */
    int           have_pending, log_threshold;
    ks265_config  k;
    ks265_encoder *enc;
    int           shard_len, cap_len;     /* pictures per GOP shard; what the per-picture arrays are sized for */
    size_t        frame_bytes;
    uint8_t      *frames;                 /* page-locked: shard_len pictures, tightly packed I420 */
    long long    *pts_in;                 /* pts of the buffered pictures, display order */
    int           buffered, key_request;
    uint8_t      *bs; size_t bs_cap;      /* Annex-B output of the last encoded shard */
    au_rec       *aus; int au_count, au_next;
    ksqy_nal     *nals; int nal_cap, nal_total;      /* NAL table of the whole shard; an access unit is a slice of it */
    ks265_pic_stat *pstat;
    uint8_t       hdr[512]; ksqy_nal hdr_nals[3];
    long long     frames_in, frames_out;
} qy_enc;

static int map_config(const ksqy_config *c, ks265_config *k, int log_threshold)
{
    if (c->width <= 0 || c->height <= 0 || (c->width & 1) || (c->height & 1)) { logf_(2, log_threshold, "ks265qy: picture size %dx%d is not usable (even, positive)\n", c->width, c->height); return KSQY_FAIL; }
    if (c->preset < 0 || c->preset > 8) return KSQY_FAIL;
    if (c->rc != 0 && c->rc != 3) {
        logf_(2, log_threshold, "ks265qy: rc %d is not implemented on the device path (0 = fixed QP, 3 = CRF)\n", c->rc);
        return KSQY_NOTSUPPORTED;
    }
    memset(k, 0, sizeof(*k));
    k->width = c->width; k->height = c->height;
    if (ks265_config_default_preset(k, k_preset_names[c->preset])) return KSQY_FAIL;
    k->fps = c->fps > 0 ? c->fps : 25.0; k->rc = c->rc; k->qp = c->qp; k->crf = (double)c->crf;
    k->iper = c->intra_period > 0 ? c->intra_period : 256;
    k->bframes = c->bframes > 0 ? c->bframes : 0;
    k->sao = c->sao; k->me = c->me > 1 ? 1 : (c->me < 0 ? 0 : c->me); k->subpel = c->subme < 0 ? 0 : (c->subme > 2 ? 2 : c->subme); k->satd = c->satd_inter != 0;
    if (c->search_range > 0) k->me_range = c->search_range;
    k->psnr = c->calc_psnr != 0;
    if (c->rdoq || c->part || c->transskip || c->long_term_ref || c->aq_mode || c->pass || c->vpp_denoise || c->vpp_edge || c->vpp_color || c->vpp_hdr)
        logf_(1, log_threshold, "ks265qy: rdoq / part / transskip / longTermRef / AQ / two-pass / vpp are not on the device path and are ignored\n");
    return KSQY_OK;
}

static void free_shard(qy_enc *q)
{
    if (q->enc) ks265_encoder_close(q->enc);
    if (q->frames) ks265_free_host(q->frames);
    free(q->pts_in); free(q->bs); free(q->aus); free(q->nals); free(q->pstat);
    q->enc = NULL; q->frames = NULL; q->pts_in = NULL; q->bs = NULL; q->aus = NULL; q->nals = NULL; q->pstat = NULL;
}

static int open_shard(qy_enc *q)
{
    int err = 0, r = map_config(&q->cfg, &q->k, q->log_threshold);
    if (r) return r;
    q->enc = ks265_encoder_open(&q->k, &err);
    if (!q->enc) { logf_(2, q->log_threshold, "ks265qy: encoder open failed (%d)\n", err); return err == -12 ? KSQY_OUTOFMEMORY : KSQY_FAIL; }
    q->shard_len = q->cap_len = q->k.iper;
    q->frame_bytes = (size_t)q->k.width * q->k.height * 3 / 2;
    q->frames = (uint8_t *)ks265_alloc_host(q->frame_bytes * q->shard_len);
    q->pts_in = (long long *)malloc(sizeof(long long) * q->shard_len);
    q->bs_cap = q->frame_bytes * q->shard_len / 4 + (4u << 20);
    q->bs = (uint8_t *)malloc(q->bs_cap);
    q->aus = (au_rec *)malloc(sizeof(au_rec) * q->shard_len);
    q->nal_cap = 4 * q->shard_len + 8;
    q->nals = (ksqy_nal *)malloc(sizeof(ksqy_nal) * q->nal_cap);
    q->pstat = (ks265_pic_stat *)malloc(sizeof(ks265_pic_stat) * q->shard_len);
    if (!q->frames || !q->pts_in || !q->bs || !q->aus || !q->nals || !q->pstat) { free_shard(q); return KSQY_OUTOFMEMORY; }
    return KSQY_OK;
}

/* a Reconfig takes effect here, between two shards (nothing buffered).  Only the encoder is re-opened: the output of the previous shard
 * (bs / aus / nals) may still be queued and stays where it is; the per-picture arrays only ever grow. */
static int reconfigure(qy_enc *q)
{
    ks265_config k; int err = 0;
    q->have_pending = 0;
    if (map_config(&q->pending, &k, q->log_threshold)) return KSQY_OK;                   /* (checked in Reconfig already) the old configuration stays */
    ks265_encoder *e = ks265_encoder_open(&k, &err);
    if (!e) { logf_(2, q->log_threshold, "ks265qy: Reconfig: encoder open failed (%d); the old configuration stays\n", err); return KSQY_OK; }
    if (k.iper > q->cap_len) {
        uint8_t *f = (uint8_t *)ks265_alloc_host(q->frame_bytes * k.iper);
        long long *pi = (long long *)malloc(sizeof(long long) * k.iper);
        ks265_pic_stat *ps = (ks265_pic_stat *)malloc(sizeof(ks265_pic_stat) * k.iper);
        au_rec *au = f && pi && ps ? (au_rec *)realloc(q->aus, sizeof(au_rec) * k.iper) : NULL;
        if (au) q->aus = au;
        ksqy_nal *nl = au ? (ksqy_nal *)realloc(q->nals, sizeof(ksqy_nal) * (4 * k.iper + 8)) : NULL;
        if (!nl) { if (f) ks265_free_host(f); free(pi); free(ps); ks265_encoder_close(e); return KSQY_OUTOFMEMORY; }
        ks265_free_host(q->frames); free(q->pts_in); free(q->pstat);
        q->frames = f; q->pts_in = pi; q->pstat = ps; q->nals = nl; q->nal_cap = 4 * k.iper + 8; q->cap_len = k.iper;
    }
    ks265_encoder_close(q->enc);
    q->enc = e; q->k = k; q->cfg = q->pending; q->shard_len = k.iper; q->log_threshold = q->cfg.log_level;
    return KSQY_OK;
}

void *QY265EncoderOpen(ksqy_config *cfg, int *error_code)
{
    int dummy; if (!error_code) error_code = &dummy;
    if (!cfg) { *error_code = KSQY_POINTER; return NULL; }
    qy_enc *q = (qy_enc *)calloc(1, sizeof(*q));
    if (!q) { *error_code = KSQY_OUTOFMEMORY; return NULL; }
    q->cfg = *cfg; q->log_threshold = cfg->log_level;
    if ((*error_code = open_shard(q))) { free(q); return NULL; }
    logf_(0, q->log_threshold, "ks265qy: %dx%d %.3f fps preset %s rc %d qp %d crf %d, GOP shards of %d pictures on CUDA device %d\n", q->k.width, q->k.height, q->k.fps,
          k_preset_names[q->k.preset], q->k.rc, q->k.qp, (int)q->k.crf, q->shard_len, q->k.device);
    return q;
}

void QY265EncoderClose(void *h)
{
    qy_enc *q = (qy_enc *)h;
    if (!q) return;
    free_shard(q); free(q);
}

void QY265EncoderReconfig(void *h, ksqy_config *cfg)
{
    qy_enc *q = (qy_enc *)h;
    if (!q || !cfg) return;
    if (cfg->width != q->cfg.width || cfg->height != q->cfg.height) { logf_(2, q->log_threshold, "ks265qy: Reconfig cannot change the picture size; ignored\n"); return; }
    ks265_config k;
    if (map_config(cfg, &k, q->log_threshold)) { logf_(2, q->log_threshold, "ks265qy: Reconfig rejected (unsupported settings); the old configuration stays\n"); return; }
    q->pending = *cfg; q->have_pending = 1;
}

/* split Annex-B bytes into NAL records (payload keeps its start code, the way the reference's callers fwrite it) */
static int split_nals(uint8_t *p, size_t n, ksqy_nal *out, int cap)
{
    int cnt = 0; size_t i = 0, start = (size_t)-1;
    while (i + 3 <= n) {
        const int sc3 = p[i] == 0 && p[i + 1] == 0 && p[i + 2] == 1;
        const int sc4 = i + 4 <= n && p[i] == 0 && p[i + 1] == 0 && p[i + 2] == 0 && p[i + 3] == 1;
        if (sc3 || sc4) {
            if (start != (size_t)-1 && cnt < cap) { out[cnt - 1].size = (int)(i - start); }
            if (cnt >= cap) return -1;
            start = i;
            const size_t hdr = i + (sc4 ? 4 : 3);
            out[cnt].payload = p + i; out[cnt].nal_type = hdr < n ? (p[hdr] >> 1) & 63 : 0; out[cnt].tid = hdr + 1 < n ? (p[hdr + 1] & 7) - 1 : 0; out[cnt].pts = 0; out[cnt].size = 0;
            cnt++;
            i = hdr;
        } else i++;
    }
    if (cnt) out[cnt - 1].size = (int)(n - start);
    return cnt;
}

static int encode_shard(qy_enc *q)
{
    ks265_gop_stats st;
    ks265_encoder_set_picture_stats(q->enc, q->pstat, q->shard_len);
    long n;
    for (;;) {
        n = ks265_encoder_encode_gop(q->enc, q->frames, NULL, q->buffered, q->bs, q->bs_cap, NULL, &st);
        if (n != -28) break;                                   /* output buffer too small: grow and encode the shard again */
        uint8_t *nb = (uint8_t *)realloc(q->bs, q->bs_cap * 2);
        if (!nb) return KSQY_OUTOFMEMORY;
        q->bs = nb; q->bs_cap *= 2;
    }
    if (n < 0) { logf_(2, q->log_threshold, "ks265qy: device encode failed (%ld)\n", n); return KSQY_FAIL; }
    q->nal_total = split_nals(q->bs, (size_t)n, q->nals, q->nal_cap);
    if (q->nal_total < 0) return KSQY_FAIL;
    /* access units: parameter sets ride with the slice that follows them; one slice NAL per picture, coding order == pstat order */
    int au = 0, first = 0;
    for (int i = 0; i < q->nal_total && au < q->buffered; i++) {
        if (q->nals[i].nal_type >= 32) continue;
        au_rec *a = &q->aus[au];
        a->first_nal = first; a->nal_count = i - first + 1;
        a->off = (size_t)(q->nals[first].payload - q->bs); a->len = (size_t)(q->nals[i].payload - q->bs) + q->nals[i].size - a->off;
        a->slice_type = q->pstat[au].slice_type; a->poc = q->pstat[au].poc;
        a->pts = q->pts_in[a->poc]; a->dts = q->pts_in[au];
        for (int j = first; j <= i; j++) q->nals[j].pts = a->pts;
        first = i + 1; au++;
    }
    if (au != q->buffered) { logf_(2, q->log_threshold, "ks265qy: %d pictures in, %d access units out\n", q->buffered, au); return KSQY_FAIL; }
    q->au_count = au; q->au_next = 0; q->buffered = 0;
    if (q->cfg.calc_psnr) {
        const double px = (double)q->k.width * q->k.height * au;
        logf_(0, q->log_threshold, "ks265qy: shard of %d pictures, %llu bytes, sse Y %llu (%.4f per sample)\n", au, (unsigned long long)st.bytes, (unsigned long long)st.sse[0], (double)st.sse[0] / px);
    }
    return KSQY_OK;
}

int QY265EncoderEncodeHeaders(void *h, ksqy_nal **nals, int *nal_count)
{
    qy_enc *q = (qy_enc *)h;
    if (!q || !nals || !nal_count) return KSQY_POINTER;
    const long n = ks265_encoder_headers(q->enc, q->hdr, sizeof(q->hdr));
    if (n < 0) return KSQY_FAIL;
    const int cnt = split_nals(q->hdr, (size_t)n, q->hdr_nals, 3);
    if (cnt < 0) return KSQY_FAIL;
    *nals = q->hdr_nals; *nal_count = cnt;
    return (int)n;
}

int QY265EncoderEncodeFrame(void *h, ksqy_nal **nals, int *nal_count, ksqy_picture *in, ksqy_picture *out, int force_logo)
{
    qy_enc *q = (qy_enc *)h;
    (void)force_logo;                                          /* no licence, no logo */
    if (!q || !nals || !nal_count) return KSQY_POINTER;
    *nals = NULL; *nal_count = 0;
    int r;
    /* nothing may overwrite the output of the previous shard while access units of it are still owed to the caller: the queue drains one
     * unit per call and a shard needs shard_len calls to fill, so this only bites after key-frame requests in quick succession */
    const int queue_busy = q->au_next < q->au_count;
    if (in) {
        if (!in->yuv || !in->yuv->plane[0] || !in->yuv->plane[1] || !in->yuv->plane[2]) return KSQY_POINTER;
        if (in->yuv->width != q->k.width || in->yuv->height != q->k.height) return KSQY_FAIL;
        if (q->buffered && q->key_request && !queue_busy) { if ((r = encode_shard(q))) return r; q->key_request = 0; }
        if (q->buffered == q->shard_len) {                     /* only when key-frame requests kept the queue busy for a whole shard */
            if (queue_busy) { logf_(2, q->log_threshold, "ks265qy: output queue not drained; fetch the pending access units first\n"); return KSQY_FAIL; }
            if ((r = encode_shard(q))) return r;
        }
        if (!q->buffered && q->have_pending && (r = reconfigure(q))) return r;      /* shard boundary */
        if (!q->buffered) q->key_request = 0;                  /* the picture that opens a shard is the key frame */
        uint8_t *dst = q->frames + q->frame_bytes * q->buffered;
        for (int c = 0; c < 3; c++) {
            const int w = c ? q->k.width / 2 : q->k.width, hh = c ? q->k.height / 2 : q->k.height;
            const unsigned char *src = in->yuv->plane[c];
            for (int y = 0; y < hh; y++) { memcpy(dst, src, (size_t)w); dst += w; src += in->yuv->stride[c]; }
        }
        q->pts_in[q->buffered++] = in->pts;
        q->frames_in++;
        if (q->buffered == q->shard_len && !queue_busy) { if ((r = encode_shard(q))) return r; }
    } else if (q->buffered && !queue_busy) {                   /* flush */
        if ((r = encode_shard(q))) return r;
    }
    if (q->au_next >= q->au_count) return 0;
    const au_rec *a = &q->aus[q->au_next++];
    *nals = q->nals + a->first_nal; *nal_count = a->nal_count;
    if (out) { out->slice_type = a->slice_type; out->poc = a->poc; out->pts = a->pts; out->dts = a->dts; out->yuv = NULL; }
    q->frames_out++;
    return (int)a->len;
}

void QY265EncoderKeyFrameRequest(void *h) { qy_enc *q = (qy_enc *)h; if (q) q->key_request = 1; }

int QY265EncoderDelayedFrames(void *h)
{
    qy_enc *q = (qy_enc *)h;
    return q ? q->buffered + (q->au_count - q->au_next) : 0;
}
