/*
 * ks_bitstream.c -- HEVC Main-profile bitstream writer for the ks265 B200 encoder (host side).
 * See ks_bitstream.h for the reference counterparts.  Section numbers refer to ITU-T H.265 (v3+).
 */
#include "ks_bitstream.h"
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <stdlib.h>
#include <string.h>

/* =========================================================== raw bit writer ======================== */
typedef struct { uint8_t *buf; size_t cap, pos; uint32_t cur; int nbits; int overflow; } bitw;

static void bw_init(bitw *b, uint8_t *buf, size_t cap) { b->buf = buf; b->cap = cap; b->pos = 0; b->cur = 0; b->nbits = 0; b->overflow = 0; }
static void bw_byte(bitw *b, uint8_t v) { if (b->pos < b->cap) b->buf[b->pos++] = v; else b->overflow = 1; }
static void bw_put(bitw *b, uint32_t v, int n)
{
    while (n > 0) {
        int take = 8 - b->nbits; if (take > n) take = n;
        b->cur = (b->cur << take) | ((v >> (n - take)) & ((1u << take) - 1));
        b->nbits += take; n -= take;
        if (b->nbits == 8) { bw_byte(b, (uint8_t)b->cur); b->cur = 0; b->nbits = 0; }
    }
}
static void bw_ue(bitw *b, uint32_t v)
{
    uint32_t x = v + 1; int len = 0;
    while ((x >> len) > 1) len++;
    bw_put(b, 0, len); bw_put(b, x, len + 1);
}
static void bw_se(bitw *b, int v) { bw_ue(b, v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
static void bw_trailing(bitw *b) { bw_put(b, 1, 1); if (b->nbits) bw_put(b, 0, 8 - b->nbits); }

/* Annex-B NAL: start code + 2-byte header + emulation-prevented payload (7.4.2 / Annex B) */
static long nal_emit(uint8_t *out, size_t cap, int nal_type, int tid, const uint8_t *rbsp, size_t n)
{
    size_t o = 0; int zeros = 0;
    if (cap < 6) return -1;
    out[o++] = 0; out[o++] = 0; out[o++] = 0; out[o++] = 1;
    out[o++] = (uint8_t)(nal_type << 1); out[o++] = (uint8_t)(tid + 1);
    for (size_t i = 0; i < n; i++) {
        if (o + 2 > cap) return -1;
        if (zeros >= 2 && rbsp[i] <= 3) { out[o++] = 3; zeros = 0; }
        out[o++] = rbsp[i];
        zeros = rbsp[i] == 0 ? zeros + 1 : 0;
    }
    return (long)o;
}

/* =========================================================== parameter sets ======================== */
static int level_idc(const ks_stream_params *sp)
{   /* Table A.8 by luma picture size (reference emits 93/120/150 for 720p/1080p/2160p) */
    long ps = (long)sp->width * sp->height;
    if (ps <= 552960) return 90; if (ps <= 983040) return 93; if (ps <= 2228224) return 120;
    if (ps <= 8912896) return 150; return 180;
}
static void write_ptl(bitw *b, const ks_stream_params *sp)
{   /* 7.3.3 profile_tier_level(1, 0): Main profile, main tier */
    bw_put(b, 0, 2); bw_put(b, 0, 1); bw_put(b, 1, 5);
    bw_put(b, 0x60000000u, 32);            /* compatibility flags: profiles 1 and 2 */
    bw_put(b, 1, 1); bw_put(b, 0, 1); bw_put(b, 0, 1); bw_put(b, 1, 1);   /* progressive, !interlaced, !non-packed, frame-only */
    bw_put(b, 0, 32); bw_put(b, 0, 12);    /* 43 reserved zero bits + inbld/reserved */
    bw_put(b, (uint32_t)level_idc(sp), 8);
}
long ks_write_vps(const ks_stream_params *sp, uint8_t *out, size_t cap)
{
    uint8_t tmp[128]; bitw b; bw_init(&b, tmp, sizeof(tmp));
    bw_put(&b, 0, 4); bw_put(&b, 3, 2); bw_put(&b, 0, 6); bw_put(&b, 0, 3); bw_put(&b, 1, 1); bw_put(&b, 0xffff, 16);
    write_ptl(&b, sp);
    bw_put(&b, 1, 1); bw_ue(&b, sp->bframes ? 2 : 1); bw_ue(&b, sp->bframes ? 1 : 0); bw_ue(&b, 0);      /* sub_layer_ordering_info: dpb 2 (3 with B), reorder 0 (1) */
    bw_put(&b, 0, 6); bw_ue(&b, 0); bw_put(&b, 0, 1); bw_put(&b, 0, 1);
    bw_trailing(&b);
    return nal_emit(out, cap, 32, 0, tmp, b.pos);
}
long ks_write_sps(const ks_stream_params *sp, uint8_t *out, size_t cap)
{
    uint8_t tmp[256]; bitw b; bw_init(&b, tmp, sizeof(tmp));
    bw_put(&b, 0, 4); bw_put(&b, 0, 3); bw_put(&b, 1, 1);
    write_ptl(&b, sp);
    bw_ue(&b, 0); bw_ue(&b, 1);
    bw_ue(&b, (uint32_t)sp->width); bw_ue(&b, (uint32_t)sp->height);
    if (sp->width != sp->disp_width || sp->height != sp->disp_height) {
        bw_put(&b, 1, 1); bw_ue(&b, 0); bw_ue(&b, (uint32_t)(sp->width - sp->disp_width) / 2);
        bw_ue(&b, 0); bw_ue(&b, (uint32_t)(sp->height - sp->disp_height) / 2);
    } else bw_put(&b, 0, 1);
    bw_ue(&b, 0); bw_ue(&b, 0);
    bw_ue(&b, (uint32_t)sp->log2_max_poc_lsb - 4);
    bw_put(&b, 1, 1); bw_ue(&b, sp->bframes ? 2 : 1); bw_ue(&b, sp->bframes ? 1 : 0); bw_ue(&b, 0);
    bw_ue(&b, KS_MIN_CB_LOG2 - 3);               /* log2_min_luma_coding_block_size_minus3: 8x8 (used by intra CUs only) */
    bw_ue(&b, KS_CTU_LOG2 - KS_MIN_CB_LOG2);     /* log2_diff_max_min_luma_coding_block_size */
    bw_ue(&b, 0); bw_ue(&b, KS_MAX_TB_LOG2 - 2); /* TB 4..32 */
    bw_ue(&b, 0); bw_ue(&b, 0);                  /* max_transform_hierarchy_depth inter/intra = 0 (reference: 0/0) */
    bw_put(&b, 0, 1); bw_put(&b, 0, 1);          /* scaling lists off, AMP off */
    bw_put(&b, (uint32_t)sp->sao, 1); bw_put(&b, 0, 1);
    bw_ue(&b, 0); bw_put(&b, 0, 1);              /* no SPS RPS candidates, no long-term */
    bw_put(&b, 0, 1);                            /* sps_temporal_mvp_enabled_flag = 0 */
    bw_put(&b, (uint32_t)sp->strong_intra_smoothing, 1);
    bw_put(&b, 0, 1); bw_put(&b, 0, 1);          /* no VUI, no extension */
    bw_trailing(&b);
    return nal_emit(out, cap, 33, 0, tmp, b.pos);
}
long ks_write_pps(const ks_stream_params *sp, uint8_t *out, size_t cap)
{
    uint8_t tmp[128]; bitw b; bw_init(&b, tmp, sizeof(tmp));
    bw_ue(&b, 0); bw_ue(&b, 0); bw_put(&b, 0, 1); bw_put(&b, 0, 1); bw_put(&b, 0, 3);
    bw_put(&b, (uint32_t)sp->sign_hiding, 1); bw_put(&b, 0, 1);
    bw_ue(&b, 0); bw_ue(&b, 0); bw_se(&b, 0);
    bw_put(&b, 0, 1); bw_put(&b, 0, 1); bw_put(&b, 0, 1);    /* constrained intra, transform skip, cu_qp_delta */
    bw_se(&b, 0); bw_se(&b, 0); bw_put(&b, 0, 1);
    bw_put(&b, 0, 1); bw_put(&b, 0, 1); bw_put(&b, 0, 1);    /* weighted pred / bipred, transquant bypass */
    bw_put(&b, 0, 1); bw_put(&b, 0, 1);                      /* tiles, entropy_coding_sync */
    bw_put(&b, 1, 1);                                        /* loop filter across slices */
    bw_put(&b, 1, 1); bw_put(&b, 1, 1); bw_put(&b, 0, 1);    /* deblocking control present, override enabled, not disabled */
    bw_se(&b, sp->pps_beta_offset_div2); bw_se(&b, sp->pps_tc_offset_div2);
    bw_put(&b, 0, 1); bw_put(&b, 0, 1); bw_ue(&b, 0); bw_put(&b, 0, 1); bw_put(&b, 0, 1);
    bw_trailing(&b);
    return nal_emit(out, cap, 34, 0, tmp, b.pos);
}

/* =========================================================== CABAC engine (9.3.4) ================== */
static const uint8_t range_lps[64][4] = {
    {128,176,208,240},{128,167,197,227},{128,158,187,216},{123,150,178,205},{116,142,169,195},{111,135,160,185},{105,128,152,175},{100,122,144,166},
    {95,116,137,158},{90,110,130,150},{85,104,123,142},{81,99,117,135},{77,94,111,128},{73,89,105,122},{69,85,100,116},{66,80,95,110},
    {62,76,90,104},{59,72,86,99},{56,69,81,94},{53,65,77,89},{51,62,73,85},{48,59,69,80},{46,56,66,76},{43,53,63,72},
    {41,50,59,69},{39,48,56,65},{37,45,54,62},{35,43,51,59},{33,41,48,56},{32,39,46,53},{30,37,43,50},{29,35,41,48},
    {27,33,39,45},{26,31,37,43},{24,30,35,41},{23,28,33,39},{22,27,32,37},{21,26,30,35},{20,24,29,33},{19,23,27,31},
    {18,22,26,30},{17,21,25,28},{16,20,23,27},{15,19,22,25},{14,18,21,24},{14,17,20,23},{13,16,19,22},{12,15,18,21},
    {12,14,17,20},{11,14,16,19},{11,13,15,18},{10,12,15,17},{10,12,14,16},{9,11,13,15},{9,11,12,14},{8,10,12,14},
    {8,9,11,13},{7,9,11,12},{7,9,10,12},{7,8,10,11},{6,8,9,11},{6,7,9,10},{6,7,8,9},{2,2,2,2}};
static const uint8_t next_lps[64] = {0,0,1,2,2,4,4,5,6,7,8,9,9,11,11,12,13,13,15,15,16,16,18,18,19,19,21,21,22,22,23,24,
    24,25,26,26,27,27,28,29,29,30,30,30,31,32,32,33,33,33,34,34,35,35,35,36,36,36,37,37,37,38,38,63};

/* context layout */
enum {
    CX_SPLIT_CU = 0, CX_SKIP = 3, CX_MERGE_FLAG = 6, CX_MERGE_IDX = 7, CX_PART_MODE = 8, CX_PRED_MODE = 12,
    CX_PREV_INTRA = 13, CX_CHROMA_PRED = 14, CX_MVD = 15, CX_CBF_LUMA = 17, CX_CBF_CHROMA = 19, CX_ROOT_CBF = 24,
    CX_LAST_X = 25, CX_LAST_Y = 43, CX_CSBF = 61, CX_SIG = 65, CX_GT1 = 107, CX_GT2 = 131, CX_MVP_IDX = 137,
    CX_SAO_MERGE = 138, CX_SAO_TYPE = 139, CX_INTER_DIR = 140, CX_COUNT = 145
};
#define CNU 154
/* init values (Tables 9-5..9-37), rows: initType 0 (I), 1 (P), 2 (B) */
static const uint8_t init_values[3][CX_COUNT] = {
 { /* I */
   139,141,157, CNU,CNU,CNU, CNU, CNU, 184,CNU,CNU,CNU, CNU, 184, 63, CNU,CNU, 111,141, 94,138,182,154,154, CNU,
   110,110,124,125,140,153,125,127,140,109,111,143,127,111,79,108,123,63,
   110,110,124,125,140,153,125,127,140,109,111,143,127,111,79,108,123,63,
   91,171,134,141,
   111,111,125,110,110,94,124,108,124,107,125,141,179,153,125,107,125,141,179,153,125,107,125,141,179,153,125,
   140,139,182,182,152,136,152,136,153,136,139,111,136,139,111,
   140,92,137,138,140,152,138,139,153,74,149,92,139,107,122,152,140,179,166,182,140,227,122,197,
   138,153,136,167,152,152, CNU, 153, 200, CNU,CNU,CNU,CNU,CNU },
 { /* P */
   107,139,126, 197,185,201, 110, 122, 154,139,154,154, 149, 154, 152, 140,198, 153,111, 149,107,167,154,154, 79,
   125,110,94,110,95,79,125,111,110,78,110,111,111,95,94,108,123,108,
   125,110,94,110,95,79,125,111,110,78,110,111,111,95,94,108,123,108,
   121,140,61,154,
   155,154,139,153,139,123,123,63,153,166,183,140,136,153,154,166,183,140,136,153,154,166,183,140,136,153,154,
   170,153,123,123,107,121,107,121,167,151,183,140,151,183,140,
   154,196,196,167,154,152,167,182,182,134,149,136,153,121,136,137,169,194,166,167,154,167,137,182,
   107,167,91,122,107,167, 168, 153, 185, 95,79,63,31,31 },
 { /* B */
   107,139,126, 197,185,201, 154, 137, 154,139,154,154, 134, 183, 152, 169,198, 153,111, 149,92,167,154,154, 79,
   125,110,124,110,95,94,125,111,111,79,125,126,111,111,79,108,123,93,
   125,110,124,110,95,94,125,111,111,79,125,126,111,111,79,108,123,93,
   121,140,61,154,
   170,154,139,153,139,123,123,63,124,166,183,140,136,153,154,166,183,140,136,153,154,166,183,140,136,153,154,
   170,153,138,138,122,121,122,121,167,151,183,140,151,183,140,
   154,196,167,167,154,152,167,182,182,134,149,136,153,121,136,122,169,208,166,167,154,152,167,182,
   107,167,91,107,107,167, 168, 153, 160, 95,79,63,31,31 },
};

typedef struct {
    uint32_t low, range; int bits_left, buffered, num_buffered;
    uint8_t *buf; size_t pos, cap; int overflow;
    uint8_t ctx[CX_COUNT];         /* (pStateIdx<<1)|valMps */
} cabac;

static void cb_byte(cabac *c, int v) { if (c->pos < c->cap) c->buf[c->pos++] = (uint8_t)v; else c->overflow = 1; }
static void cb_init(cabac *c, uint8_t *buf, size_t cap, int init_type, int qp)
{
    c->low = 0; c->range = 510; c->bits_left = 23; c->buffered = 0xff; c->num_buffered = 0;
    c->buf = buf; c->cap = cap; c->pos = 0; c->overflow = 0;
    if (qp < 0) qp = 0; if (qp > 51) qp = 51;
    for (int i = 0; i < CX_COUNT; i++) {      /* 9.3.2.2 */
        int v = init_values[init_type][i];
        int m = (v >> 4) * 5 - 45, n = ((v & 15) << 3) - 16;
        int pre = ((m * qp) >> 4) + n;
        if (pre < 1) pre = 1; if (pre > 126) pre = 126;
        int mps = pre > 63;
        c->ctx[i] = (uint8_t)(((mps ? pre - 64 : 63 - pre) << 1) | mps);
    }
}
static void cb_write_out(cabac *c)
{
    uint32_t lead = c->low >> (24 - c->bits_left);
    c->bits_left += 8;
    c->low &= 0xffffffffu >> c->bits_left;
    if (lead == 0xff) c->num_buffered++;
    else if (c->num_buffered > 0) {
        uint32_t carry = lead >> 8;
        cb_byte(c, (int)(c->buffered + carry));
        c->buffered = (int)(lead & 0xff);
        int byte = (int)((0xff + carry) & 0xff);
        while (c->num_buffered > 1) { cb_byte(c, byte); c->num_buffered--; }
    } else { c->num_buffered = 1; c->buffered = (int)lead; }
}
/* context-coded bin (9.3.4.3.2 restated without the MPS/LPS branch): the LPS path is selected with a mask, the state
 * transition and the renormalisation shift come from tables (g_cb_next / g_cb_shift, built once) */
static uint8_t g_cb_next[128][2];      /* [(pStateIdx<<1)|valMps][bin] */
static uint8_t g_cb_shift[64];         /* [range >> 3] -> left shifts until range >= 256 */
static uint32_t g_cb_lps4[128];        /* [(pStateIdx<<1)|valMps] -> the four rangeTabLps entries as bytes: the load depends on the context only,
                                          the range then just selects a byte (keeps a memory access out of the bin-to-bin dependency chain) */
__attribute__((constructor)) static void init_cb_tables(void)
{
    for (int s = 0; s < 128; s++) for (int bin = 0; bin < 2; bin++) {
        int state = s >> 1, mps = s & 1;
        if (bin != mps) { if (state == 0) mps ^= 1; g_cb_next[s][bin] = (uint8_t)((next_lps[state] << 1) | mps); }
        else g_cb_next[s][bin] = (uint8_t)(((state < 62 ? state + 1 : state) << 1) | mps);
    }
    for (int k = 0; k < 64; k++) { int r = k << 3 | 7, n = 0; while (r < 256) { r <<= 1; n++; } g_cb_shift[k] = (uint8_t)n; }
    for (int s = 0; s < 128; s++) g_cb_lps4[s] = (uint32_t)range_lps[s >> 1][0] | ((uint32_t)range_lps[s >> 1][1] << 8) | ((uint32_t)range_lps[s >> 1][2] << 16) | ((uint32_t)range_lps[s >> 1][3] << 24);
}
static inline void cb_bin(cabac *c, int ctx, int bin)
{
    uint32_t s = c->ctx[ctx], range = c->range, low = c->low;
    uint32_t lps = (g_cb_lps4[s] >> ((range >> 3) & 24)) & 255u;
    uint32_t is_lps = 0u - (uint32_t)((uint32_t)bin != (s & 1));
    range -= lps;
    low += range & is_lps;
    range = (range & ~is_lps) | (lps & is_lps);
    c->ctx[ctx] = g_cb_next[s][bin];
    uint32_t nb = g_cb_shift[range >> 3];
    c->low = low << nb; c->range = range << nb;
    if ((c->bits_left -= (int)nb) < 12) cb_write_out(c);
}
static inline void cb_bypass(cabac *c, int bin)
{
    c->low <<= 1; if (bin) c->low += c->range;
    if (--c->bits_left < 12) cb_write_out(c);
}
static void cb_bypass_bins(cabac *c, uint32_t bins, int n)
{
    while (n > 8) {
        n -= 8; uint32_t pat = bins >> n;
        c->low <<= 8; c->low += c->range * pat; bins -= pat << n;
        c->bits_left -= 8; if (c->bits_left < 12) cb_write_out(c);
    }
    c->low <<= n; c->low += c->range * bins; c->bits_left -= n;
    if (c->bits_left < 12) cb_write_out(c);
}
static void cb_terminate(cabac *c, int bin)
{
    c->range -= 2;
    if (bin) { c->low += c->range; c->low <<= 7; c->range = 2 << 7; c->bits_left -= 7; }
    else if (c->range >= 256) return;
    else { c->low <<= 1; c->range <<= 1; c->bits_left--; }
    if (c->bits_left < 12) cb_write_out(c);
}
static void cb_finish(cabac *c)
{   /* flush (9.3.4.5) then rbsp_slice_segment_trailing_bits */
    if (c->low >> (32 - c->bits_left)) {
        cb_byte(c, c->buffered + 1);
        while (c->num_buffered > 1) { cb_byte(c, 0x00); c->num_buffered--; }
        c->low -= 1u << (32 - c->bits_left);
    } else {
        if (c->num_buffered > 0) cb_byte(c, c->buffered);
        while (c->num_buffered > 1) { cb_byte(c, 0xff); c->num_buffered--; }
    }
    /* write (low >> 8) using 24 - bits_left bits, then stop bit + alignment */
    int n = 24 - c->bits_left;
    uint64_t v = ((uint64_t)(c->low >> 8) & ((1ull << n) - 1));
    v = (v << 1) | 1; n += 1;
    int pad = (8 - (n & 7)) & 7;
    v <<= pad; n += pad;
    for (int i = n - 8; i >= 0; i -= 8) cb_byte(c, (int)((v >> i) & 0xff));
}

/* =========================================================== slice data ============================ */
typedef struct {
    const ks_stream_params *sp; const ks_slice_params *sl; const ks_frame_syn *syn;
    cabac cb;
    uint8_t *skip;       /* per cell: coded as skip */
    uint8_t *scan4;      /* diag scan of a 4x4: pos -> (y<<2)|x */
    uint8_t scan_cg[4][64]; /* diag scan of CG grid for log2 tb 2..5: idx -> (y<<3)|x */
    uint16_t cg_prefix[KS_CTU / 4 + 2 * KS_CTU / 8 + 1]; /* per-CTU prefix popcounts of CG rows */
    int cur_ctu;
} slice_enc;

static uint8_t g_scan4[16];
static uint8_t g_scan_cg[4][64];
static int g_scan_ready;
static void build_diag(uint8_t *dst, int n, int shift)
{   /* 6.5.3 up-right diagonal scan */
    int i = 0, x = 0, y = 0;
    for (;;) {
        while (y >= 0) { if (x < n && y < n) dst[i++] = (uint8_t)((y << shift) | x); y--; x++; }
        y = x; x = 0;
        if (i >= n * n) break;
    }
}
/* built when the library is loaded (like the bin-coder tables): shard threads only ever read them */
__attribute__((constructor)) static void scans_init(void)
{
    if (g_scan_ready) return;
    build_diag(g_scan4, 4, 2);
    for (int l = 2; l <= 5; l++) build_diag(g_scan_cg[l - 2], 1 << (l - 2), 3);
    g_scan_ready = 1;
}

static inline const ks_cell *cell_at(const ks_frame_syn *s, int x, int y) { return &s->cells[(y >> KS_CELL_LOG2) * s->cells_w + (x >> KS_CELL_LOG2)]; }
static inline int zidx(int x, int y)
{
    int cx = (x >> KS_CELL_LOG2) & 3, cy = (y >> KS_CELL_LOG2) & 3;
    return (cx & 1) | ((cy & 1) << 1) | ((cx & 2) << 1) | ((cy & 2) << 2);
}
/* 6.4.1 z-scan availability of the block containing (xn,yn) for a block whose origin is (xc,yc) */
static int avail(const ks_frame_syn *s, int xc, int yc, int xn, int yn)
{
    if (xn < 0 || yn < 0 || xn >= s->width || yn >= s->height) return 0;
    int ac = (yc >> KS_CTU_LOG2) * s->ctus_w + (xc >> KS_CTU_LOG2), an = (yn >> KS_CTU_LOG2) * s->ctus_w + (xn >> KS_CTU_LOG2);
    if (an != ac) return an < ac;
    return zidx(xn, yn) < zidx(xc, yc);
}

/* ---- coefficient access: CG (cgx,cgy in 4x4 units of the component plane) -> 16 levels or NULL ---- */
static void ctu_prefix(slice_enc *e, int ctu)
{
    const ks_ctu_syn *c = &e->syn->ctus[ctu];
    int acc = 0, k = 0;
    for (int i = 0; i < 16; i++) { e->cg_prefix[k++] = (uint16_t)acc; acc += __builtin_popcount(c->cg_y[i]); }
    for (int i = 0; i < 8; i++) { e->cg_prefix[k++] = (uint16_t)acc; acc += __builtin_popcount(c->cg_cb[i]); }
    for (int i = 0; i < 8; i++) { e->cg_prefix[k++] = (uint16_t)acc; acc += __builtin_popcount(c->cg_cr[i]); }
    e->cg_prefix[k] = (uint16_t)acc;
    e->cur_ctu = ctu;
}
/* sig_coeff_flag context increments (9.3.4.2.5), tabulated: [chroma][8x8 TB][CG != (0,0)][right | below<<1][raster pos in CG] */
static uint8_t g_sig_ctx[2][2][2][4][16];
static uint8_t g_sig_ctx_scan[2][2][2][4][16];      /* the same, indexed by scan position inside the CG */
static uint16_t g_r2s_lo[256], g_r2s_hi[256];      /* raster -> scan bit permutation of a 4x4 group, low / high raster byte */
__attribute__((constructor)) static void init_sig_ctx(void)
{
    for (int ch = 0; ch < 2; ch++) for (int l3 = 0; l3 < 2; l3++) for (int nz = 0; nz < 2; nz++) for (int prev = 0; prev < 4; prev++) for (int p = 0; p < 16; p++) {
        int xp = p & 3, yp = p >> 2, sc;
        if (!nz && p == 0) sc = 0;
        else {
            if (prev == 0) sc = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
            else if (prev == 1) sc = yp == 0 ? 2 : yp == 1 ? 1 : 0;
            else if (prev == 2) sc = xp == 0 ? 2 : xp == 1 ? 1 : 0;
            else sc = 2;
            if (!ch) { if (nz) sc += 3; sc += l3 ? 9 : 21; }
            else sc += l3 ? 9 : 12;
        }
        g_sig_ctx[ch][l3][nz][prev][p] = (uint8_t)(sc + (ch ? 27 : 0));
    }
    /* scan-ordered copies + the raster->scan bit permutation of a 4x4 group (cg_scan_mask) */
    scans_init();
    for (int ch = 0; ch < 2; ch++) for (int l3 = 0; l3 < 2; l3++) for (int nz = 0; nz < 2; nz++) for (int prev = 0; prev < 4; prev++) for (int n = 0; n < 16; n++) {
        int p = g_scan4[n], yx = (p >> 2) * 4 + (p & 3);
        g_sig_ctx_scan[ch][l3][nz][prev][n] = g_sig_ctx[ch][l3][nz][prev][yx];
    }
    for (int v = 0; v < 256; v++) {
        uint16_t lo = 0, hi = 0;
        for (int n = 0; n < 16; n++) {
            int p = g_scan4[n], r = (p >> 2) * 4 + (p & 3);          /* raster index of scan position n */
            if (r < 8) { if ((v >> r) & 1) lo |= (uint16_t)(1u << n); }
            else if ((v >> (r - 8)) & 1) hi |= (uint16_t)(1u << n);
        }
        g_r2s_lo[v] = lo; g_r2s_hi[v] = hi;
    }
}

/* non-zero mask of one 4x4 coefficient group in SCAN order (bit n = scan position n): raster-order compare mask (SSE2 where available),
 * then the raster->scan bit permutation through two 256-entry tables */
static inline uint32_t cg_scan_mask(const int16_t *lv)
{
    uint32_t raster;
#if defined(__SSE2__)
    const __m128i z = _mm_setzero_si128();
    const __m128i a = _mm_cmpeq_epi16(_mm_loadu_si128((const __m128i *)lv), z), b = _mm_cmpeq_epi16(_mm_loadu_si128((const __m128i *)(lv + 8)), z);
    raster = ~(uint32_t)_mm_movemask_epi8(_mm_packs_epi16(a, b)) & 0xffffu;
#else
    raster = 0;
    for (int i = 0; i < 16; i++) raster |= (uint32_t)(lv[i] != 0) << i;
#endif
    return (uint32_t)g_r2s_lo[raster & 255] | g_r2s_hi[raster >> 8];
}

/* ---- 7.3.8.11 residual_coding for one transform block (diagonal scan only: TB >= 8 or inter) ---- */
static void code_residual(slice_enc *e, int comp, int x0c, int y0c, int log2)
{   /* x0c,y0c: TB origin inside the CTU in component samples */
    cabac *c = &e->cb;
    int ncg = 1 << (log2 - 2);           /* CGs per side */
    const uint8_t *scg = g_scan_cg[log2 - 2];
    /* coded-CG bitmap of this TB, one byte per CG row, straight from the CTU record; level pointers are resolved lazily */
    const ks_ctu_syn *cs = &e->syn->ctus[e->cur_ctu];
    int cgx0 = x0c >> 2, cgy0 = y0c >> 2, row0 = comp == 0 ? 0 : (comp == 1 ? 16 : 24);
    uint32_t rows[8], rmask = (1u << ncg) - 1, any = 0;
    for (int r = 0; r < ncg; r++) {
        uint32_t bits = comp == 0 ? cs->cg_y[cgy0 + r] : (comp == 1 ? cs->cg_cb[cgy0 + r] : cs->cg_cr[cgy0 + r]);
        rows[r] = (bits >> cgx0) & rmask; any |= rows[r];
    }
    if (!any) return;                    /* caller guarantees cbf=1 */
    /* the arithmetic coder's range/low live in locals for the whole block: the bin-to-bin dependency then runs through
     * registers instead of store-to-load forwarding (context bytes alias the engine struct) */
    uint32_t low = c->low, range = c->range; int bl = c->bits_left; uint8_t *ctxs = c->ctx;
#define RFLUSH() do { c->low = low; c->bits_left = bl; cb_write_out(c); low = c->low; bl = c->bits_left; } while (0)
#define RBIN(ci, bv) do { \
        uint32_t s_ = ctxs[ci], b_ = (uint32_t)(bv), l_ = (g_cb_lps4[s_] >> ((range >> 3) & 24)) & 255u, m_ = 0u - (uint32_t)(b_ != (s_ & 1)); \
        range -= l_; low += range & m_; range = (range & ~m_) | (l_ & m_); \
        ctxs[ci] = g_cb_next[s_][b_]; \
        uint32_t n_ = (uint32_t)__builtin_clz(range) - 23u; low <<= n_; range <<= n_; \
        if ((bl -= (int)n_) < 12) RFLUSH(); \
    } while (0)
#define RBYP(bits, nbits) do { \
        uint32_t v_ = (uint32_t)(bits); int k_ = (int)(nbits); \
        while (k_ > 8) { k_ -= 8; uint32_t p_ = v_ >> k_; low = (low << 8) + range * p_; v_ -= p_ << k_; if ((bl -= 8) < 12) RFLUSH(); } \
        low = (low << k_) + range * v_; if ((bl -= k_) < 12) RFLUSH(); \
    } while (0)
#define CG_CODED(cx, cy) ((rows[cy] >> (cx)) & 1u)
#define CG_PTR(cx, cy) (e->syn->levels + (size_t)(cs->cg_base + e->cg_prefix[row0 + cgy0 + (cy)] + \
        (uint32_t)__builtin_popcount((comp == 0 ? cs->cg_y[cgy0 + (cy)] : (comp == 1 ? cs->cg_cb[cgy0 + (cy)] : cs->cg_cr[cgy0 + (cy)])) & ((1u << (cgx0 + (cx))) - 1))) * 16)
    int last_cg = ncg * ncg - 1;
    while (!CG_CODED(scg[last_cg] & 7, scg[last_cg] >> 3)) last_cg--;
    const int16_t *lastp = CG_PTR(scg[last_cg] & 7, scg[last_cg] >> 3);
    int last_pos = 15;
    while (!lastp[g_scan4[last_pos]]) last_pos--;
    int lx = ((scg[last_cg] & 7) << 2) + (g_scan4[last_pos] & 3), ly = ((scg[last_cg] >> 3) << 2) + (g_scan4[last_pos] >> 2);
    /* last_sig_coeff_{x,y}_prefix / suffix (9.3.4.2.3) */
    static const uint8_t group_idx[32] = {0,1,2,3,4,4,5,5,6,6,6,6,7,7,7,7,8,8,8,8,8,8,8,8,9,9,9,9,9,9,9,9};
    static const uint8_t min_in_group[10] = {0,1,2,3,4,6,8,12,16,24};
    int off, shift;
    if (comp == 0) { off = 3 * (log2 - 2) + ((log2 - 1) >> 2); shift = (log2 + 1) >> 2; }
    else { off = 15; shift = log2 - 2; }
    int gx = group_idx[lx], gy = group_idx[ly], cmax = (log2 << 1) - 1, i;
    for (i = 0; i < gx; i++) RBIN(CX_LAST_X + off + (i >> shift), 1);
    if (gx < cmax) RBIN(CX_LAST_X + off + (i >> shift), 0);
    for (i = 0; i < gy; i++) RBIN(CX_LAST_Y + off + (i >> shift), 1);
    if (gy < cmax) RBIN(CX_LAST_Y + off + (i >> shift), 0);
    if (gx > 3) RBYP((uint32_t)(lx - min_in_group[gx]), (gx - 2) >> 1);
    if (gy > 3) RBYP((uint32_t)(ly - min_in_group[gy]), (gy - 2) >> 1);

    uint8_t csbf[8][8]; memset(csbf, 0, sizeof(csbf));
    int c1 = 1;
    for (int i2 = last_cg; i2 >= 0; i2--) {
        int cx = scg[i2] & 7, cy = scg[i2] >> 3;
        int right = cx + 1 < ncg ? csbf[cy][cx + 1] : 0, below = cy + 1 < ncg ? csbf[cy + 1][cx] : 0;
        int present = (int)CG_CODED(cx, cy), coded = present, infer_dc = 0;
        if (i2 < last_cg && i2 > 0) {
            RBIN(CX_CSBF + ((right | below) ? 1 : 0) + (comp ? 2 : 0), coded);
            infer_dc = 1;
        } else coded = 1;                 /* last and DC CGs are inferred coded */
        csbf[cy][cx] = (uint8_t)coded;
        if (!coded) continue;
        int prev = right | (below << 1);
        const uint8_t *sctx = g_sig_ctx_scan[comp != 0][log2 == 3][(cx | cy) != 0][prev];      /* indexed by scan position */
        const int16_t *lv = NULL;
        uint32_t nz = 0;                   /* bit n: the coefficient at scan position n of this CG is non-zero */
        if (present) { lv = i2 == last_cg ? lastp : CG_PTR(cx, cy); nz = cg_scan_mask(lv); }
        int start = i2 == last_cg ? last_pos - 1 : 15;
        for (int n = start; n > 0; n--) RBIN(CX_SIG + sctx[n], (nz >> n) & 1);
        if (start >= 0 && !(infer_dc && !(nz >> 1))) RBIN(CX_SIG + sctx[0], nz & 1);      /* DC flag unless inferred (all others zero in a coded CG) */
        if (!nz) continue;                 /* inferred-coded CG without levels (only the DC group can be one) */
        /* greater1 / greater2 / signs / remaining, visiting only the non-zero positions (high scan position first) */
        int ctx_set = (i2 > 0 && comp == 0) ? 2 : 0;
        if (c1 == 0) ctx_set++;
        c1 = 1;
        int nsig = 0, first_g1 = -1;
        int absv[16];
        uint32_t signs = 0;
        const int last_sig = 31 - __builtin_clz(nz), first_sig = __builtin_ctz(nz);
        for (uint32_t m = nz; m; ) {
            const int n = 31 - __builtin_clz(m); m &= ~(1u << n);
            const int v = lv[g_scan4[n]];
            absv[nsig++] = v < 0 ? -v : v;
            signs = (signs << 1) | (uint32_t)(v < 0);
        }
        int ng1 = nsig < 8 ? nsig : 8;
        for (int k = 0; k < ng1; k++) {
            int g1 = absv[k] > 1;
            RBIN(CX_GT1 + (comp ? 16 : 0) + 4 * ctx_set + c1, g1);
            if (g1) { c1 = 0; if (first_g1 < 0) first_g1 = k; }
            else if (c1 < 3 && c1 > 0) c1++;
        }
        if (c1 == 0 && first_g1 >= 0) RBIN(CX_GT2 + (comp ? 4 : 0) + ctx_set, absv[first_g1] > 2);
        int hidden = e->sp->sign_hiding && (last_sig - first_sig > 3);
        if (hidden) RBYP(signs >> 1, nsig - 1); else RBYP(signs, nsig);
        int rice = 0;
        for (int k = 0; k < nsig; k++) {
            /* baseLevel the flags can express: 3 for the coefficient that carried greater2, 2 for the other
             * greater1-coded ones (first 8), 1 beyond */
            int base = k < 8 ? (k == first_g1 ? 3 : 2) : 1;
            if (absv[k] >= base) {
                int rem = absv[k] - base;
                /* 9.3.3.11 coeff_abs_level_remaining: TR prefix (cMax 4<<rice) + EGk suffix */
                if (rem < (3 << rice)) {
                    int len = rem >> rice;
                    RBYP((1u << (len + 1)) - 2, len + 1);
                    RBYP((uint32_t)rem & ((1u << rice) - 1), rice);
                } else {
                    int len = rice, code = rem - (3 << rice);
                    while (code >= (1 << len)) { code -= 1 << len; len++; }
                    RBYP((1u << (3 + len + 1 - rice)) - 2, 3 + len + 1 - rice);
                    RBYP((uint32_t)code, len);
                }
                if (absv[k] > 3 * (1 << rice) && rice < 4) rice++;
            }
        }
    }
    c->low = low; c->range = range; c->bits_left = bl;
#undef RBIN
#undef RBYP
#undef RFLUSH
#undef CG_CODED
#undef CG_PTR
}

/* ---- residual_coding of the SMALL intra blocks (luma 8x8, chroma 4x4 of an 8x8 intra CU): any scanIdx (7.4.9.11: 0 diagonal, 1 horizontal,
 *      2 vertical), 4x4 blocks with the ctxIdxMap contexts (9.3.4.2.5).  Straightforward (these blocks are a small share of the bins). ---- */
static const uint8_t *cg_levels(const slice_enc *e, int comp, int cgx, int cgy)
{   /* (cgx, cgy): coefficient group inside the CTU, in 4x4 units of the component plane; NULL when the group holds no level */
    const ks_ctu_syn *cs = &e->syn->ctus[e->cur_ctu];
    uint32_t bits = comp == 0 ? cs->cg_y[cgy] : (comp == 1 ? cs->cg_cb[cgy] : cs->cg_cr[cgy]);
    if (!((bits >> cgx) & 1)) return NULL;
    int row0 = comp == 0 ? 0 : (comp == 1 ? 16 : 24);
    return (const uint8_t *)(e->syn->levels + (size_t)(cs->cg_base + e->cg_prefix[row0 + cgy] + (uint32_t)__builtin_popcount(bits & ((1u << cgx) - 1))) * 16);
}
static void code_residual_small(slice_enc *e, int comp, int x0c, int y0c, int log2, int scan_idx)
{
    cabac *c = &e->cb;
    const int ncg = 1 << (log2 - 2), is_luma = comp == 0;
    int16_t blk[64]; memset(blk, 0, sizeof(blk));                 /* the block's levels, raster n x n */
    const int n = 1 << log2;
    int any = 0;
    for (int gy = 0; gy < ncg; gy++) for (int gx = 0; gx < ncg; gx++) {
        const int16_t *lv = (const int16_t *)cg_levels(e, comp, (x0c >> 2) + gx, (y0c >> 2) + gy);
        if (!lv) continue;
        any = 1;
        for (int k = 0; k < 16; k++) blk[((gy << 2) + (k >> 2)) * n + (gx << 2) + (k & 3)] = lv[k];
    }
    if (!any) return;
    /* scan tables: position i -> (x, y) */
    uint8_t sx[64], sy[64];
    for (int cgi = 0; cgi < ncg * ncg; cgi++) {
        int gx, gy;
        if (ncg == 1) gx = gy = 0;
        else if (scan_idx == 0) { static const uint8_t dx[4] = {0, 0, 1, 1}, dy[4] = {0, 1, 0, 1}; gx = dx[cgi]; gy = dy[cgi]; }
        else if (scan_idx == 1) { gx = cgi & 1; gy = cgi >> 1; }
        else { gx = cgi >> 1; gy = cgi & 1; }
        for (int k = 0; k < 16; k++) {
            int px, py;
            if (scan_idx == 0) { px = g_scan4[k] & 3; py = g_scan4[k] >> 2; }
            else if (scan_idx == 1) { px = k & 3; py = k >> 2; }
            else { px = k >> 2; py = k & 3; }
            sx[cgi * 16 + k] = (uint8_t)((gx << 2) + px); sy[cgi * 16 + k] = (uint8_t)((gy << 2) + py);
        }
    }
    int last = n * n - 1;
    while (!blk[sy[last] * n + sx[last]]) last--;
    int lx = sx[last], ly = sy[last];
    if (scan_idx == 2) { int t = lx; lx = ly; ly = t; }          /* 7.3.8.11: coordinates are swapped for the vertical scan */
    static const uint8_t group_idx[32] = {0,1,2,3,4,4,5,5,6,6,6,6,7,7,7,7,8,8,8,8,8,8,8,8,9,9,9,9,9,9,9,9};
    static const uint8_t min_in_group[10] = {0,1,2,3,4,6,8,12,16,24};
    int off, shift;
    if (is_luma) { off = 3 * (log2 - 2) + ((log2 - 1) >> 2); shift = (log2 + 1) >> 2; } else { off = 15; shift = log2 - 2; }
    int gx = group_idx[lx], gy = group_idx[ly], cmax = (log2 << 1) - 1, i;
    for (i = 0; i < gx; i++) cb_bin(c, CX_LAST_X + off + (i >> shift), 1);
    if (gx < cmax) cb_bin(c, CX_LAST_X + off + (i >> shift), 0);
    for (i = 0; i < gy; i++) cb_bin(c, CX_LAST_Y + off + (i >> shift), 1);
    if (gy < cmax) cb_bin(c, CX_LAST_Y + off + (i >> shift), 0);
    if (gx > 3) cb_bypass_bins(c, (uint32_t)(lx - min_in_group[gx]), (gx - 2) >> 1);
    if (gy > 3) cb_bypass_bins(c, (uint32_t)(ly - min_in_group[gy]), (gy - 2) >> 1);
    static const uint8_t ctx_idx_map4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
    uint8_t csbf[2][2] = {{0, 0}, {0, 0}};
    int last_cg = last >> 4, last_pos = last & 15, c1 = 1;
    for (int cgi = last_cg; cgi >= 0; cgi--) {
        const int cx = sx[cgi * 16] >> 2, cy = sy[cgi * 16] >> 2;
        const int right = cx + 1 < ncg ? csbf[cy][cx + 1] : 0, below = cy + 1 < ncg ? csbf[cy + 1][cx] : 0;
        int nzc = 0;
        for (int k = 0; k < 16; k++) nzc += blk[sy[cgi * 16 + k] * n + sx[cgi * 16 + k]] != 0;
        int coded = nzc != 0, infer_dc = 0;
        if (cgi < last_cg && cgi > 0) { cb_bin(c, CX_CSBF + ((right | below) ? 1 : 0) + (is_luma ? 0 : 2), coded); infer_dc = 1; }
        else coded = 1;
        csbf[cy][cx] = (uint8_t)coded;
        if (!coded) continue;
        const int prev = right | (below << 1);
        int start = cgi == last_cg ? last_pos - 1 : 15, seen = cgi == last_cg;
        for (int k = start; k >= 0; k--) {
            const int x = sx[cgi * 16 + k], y = sy[cgi * 16 + k], xp = x & 3, yp = y & 3, sig = blk[y * n + x] != 0;
            if (k == 0 && infer_dc && !seen) break;           /* inferred significant */
            int sc;
            if (log2 == 2) sc = ctx_idx_map4[(yp << 2) + xp];
            else if (cx == 0 && cy == 0 && xp == 0 && yp == 0) sc = 0;
            else {
                if (prev == 0) sc = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
                else if (prev == 1) sc = yp == 0 ? 2 : yp == 1 ? 1 : 0;
                else if (prev == 2) sc = xp == 0 ? 2 : xp == 1 ? 1 : 0;
                else sc = 2;
                if (is_luma) { if (cx || cy) sc += 3; sc += scan_idx == 0 ? 9 : 15; } else sc += 9;
            }
            cb_bin(c, CX_SIG + sc + (is_luma ? 0 : 27), sig);
            seen |= sig;
        }
        if (!nzc) continue;
        int absv[16], npos[16], nsig = 0;
        uint32_t signs = 0;
        for (int k = 15; k >= 0; k--) {
            const int v = blk[sy[cgi * 16 + k] * n + sx[cgi * 16 + k]];
            if (!v) continue;
            npos[nsig] = k; absv[nsig++] = v < 0 ? -v : v; signs = (signs << 1) | (uint32_t)(v < 0);
        }
        int ctx_set = (cgi > 0 && is_luma) ? 2 : 0;
        if (c1 == 0) ctx_set++;
        c1 = 1;
        int first_g1 = -1, ng1 = nsig < 8 ? nsig : 8;
        for (int k = 0; k < ng1; k++) {
            const int g1 = absv[k] > 1;
            cb_bin(c, CX_GT1 + (is_luma ? 0 : 16) + 4 * ctx_set + c1, g1);
            if (g1) { c1 = 0; if (first_g1 < 0) first_g1 = k; } else if (c1 < 3 && c1 > 0) c1++;
        }
        if (first_g1 >= 0) cb_bin(c, CX_GT2 + (is_luma ? 0 : 4) + ctx_set, absv[first_g1] > 2);
        const int hidden = e->sp->sign_hiding && (npos[0] - npos[nsig - 1] > 3);
        if (hidden) cb_bypass_bins(c, signs >> 1, nsig - 1); else cb_bypass_bins(c, signs, nsig);
        int rice = 0;
        for (int k = 0; k < nsig; k++) {
            const int base = k < 8 ? (k == first_g1 ? 3 : 2) : 1;
            if (absv[k] >= base) {
                const int rem = absv[k] - base;
                if (rem < (3 << rice)) { const int len = rem >> rice; cb_bypass_bins(c, (1u << (len + 1)) - 2, len + 1); cb_bypass_bins(c, (uint32_t)rem & ((1u << rice) - 1), rice); }
                else {
                    int len = rice, code = rem - (3 << rice);
                    while (code >= (1 << len)) { code -= 1 << len; len++; }
                    cb_bypass_bins(c, (1u << (3 + len + 1 - rice)) - 2, 3 + len + 1 - rice);
                    cb_bypass_bins(c, (uint32_t)code, len);
                }
                if (absv[k] > 3 * (1 << rice) && rice < 4) rice++;
            }
        }
    }
}
static inline int intra_scan_idx(int mode) { return (mode >= 22 && mode <= 30) ? 1 : ((mode >= 6 && mode <= 14) ? 2 : 0); }

/* ---- merge / AMVP candidates (8.5.3.2.2-.7; one reference picture per list, no TMVP) ---- */
typedef struct { int16_t x, y; } mv_t;
typedef struct { int dir; mv_t mv[2]; } motion_t;          /* dir: bit0 list 0 used, bit1 list 1 used; unused MVs are 0 */
static inline motion_t cell_motion(const ks_frame_syn *s, int x, int y)
{
    int i = (y >> KS_CELL_LOG2) * s->cells_w + (x >> KS_CELL_LOG2);
    const ks_cell *c = &s->cells[i];
    motion_t m; m.dir = 1; m.mv[0].x = c->mvx; m.mv[0].y = c->mvy; m.mv[1].x = m.mv[1].y = 0;
    if (s->cells_b) {
        const ks_cell_b *b = &s->cells_b[i];
        m.dir = b->dir;
        if (!(m.dir & 1)) m.mv[0].x = m.mv[0].y = 0;
        if (m.dir & 2) { m.mv[1].x = b->mvx1; m.mv[1].y = b->mvy1; }
    }
    return m;
}
static inline int motion_eq(const motion_t *a, const motion_t *b)
{
    return a->dir == b->dir && a->mv[0].x == b->mv[0].x && a->mv[0].y == b->mv[0].y && a->mv[1].x == b->mv[1].x && a->mv[1].y == b->mv[1].y;
}
/* the five spatial neighbours of a 2Nx2N PU, fetched once per CU and shared by the merge and AMVP derivations.
 * A1, B1 and B2 lie left of / above the CU and always precede it in z-scan order; A0 and B0 need the 6.4.1 test. */
enum { NB_A0, NB_A1, NB_B0, NB_B1, NB_B2 };
typedef struct { motion_t m[5]; int av[5]; } nb_set;
static void fetch_nb(const ks_frame_syn *s, int x, int y, int size, nb_set *nb)
{
    const int nx[5] = {x - 1, x - 1, x + size, x + size - 1, x - 1}, ny[5] = {y + size, y + size - 1, y - 1, y - 1, y - 1};
    nb->av[NB_A0] = avail(s, x, y, nx[0], ny[0]);
    nb->av[NB_A1] = x > 0;
    nb->av[NB_B0] = avail(s, x, y, nx[2], ny[2]);
    nb->av[NB_B1] = y > 0;
    nb->av[NB_B2] = x > 0 && y > 0;
    for (int k = 0; k < 5; k++) {
        if (!nb->av[k]) continue;
        if (cell_at(s, nx[k], ny[k])->flags & KS_F_INTRA) nb->av[k] = 0;
        else nb->m[k] = cell_motion(s, nx[k], ny[k]);
    }
}
static int merge_list(const ks_frame_syn *s, const nb_set *nb, int maxc, motion_t *list)
{
    const motion_t *a1 = &nb->m[NB_A1], *b1 = &nb->m[NB_B1], *b0 = &nb->m[NB_B0], *a0 = &nb->m[NB_A0], *b2 = &nb->m[NB_B2];
    int n = 0, fa1 = nb->av[NB_A1], ab1 = nb->av[NB_B1], fb1 = ab1;
    if (fb1 && fa1 && motion_eq(a1, b1)) fb1 = 0;
    int fb0 = nb->av[NB_B0];
    if (fb0 && ab1 && motion_eq(b1, b0)) fb0 = 0;
    int fa0 = nb->av[NB_A0];
    if (fa0 && fa1 && motion_eq(a1, a0)) fa0 = 0;
    int fb2 = nb->av[NB_B2];
    if (fb2 && fa1 && motion_eq(a1, b2)) fb2 = 0;
    if (fb2 && ab1 && motion_eq(b1, b2)) fb2 = 0;
    if (fa0 + fa1 + fb0 + fb1 == 4) fb2 = 0;
    if (fa1 && n < maxc) list[n++] = *a1;
    if (fb1 && n < maxc) list[n++] = *b1;
    if (fb0 && n < maxc) list[n++] = *b0;
    if (fa0 && n < maxc) list[n++] = *a0;
    if (fb2 && n < maxc) list[n++] = *b2;
    if (s->slice_type == KS_SLICE_B && n > 1 && n < maxc) {        /* 8.5.3.2.4 combined bi-predictive candidates */
        static const uint8_t l0i[12] = {0, 1, 0, 2, 1, 2, 0, 3, 1, 3, 2, 3}, l1i[12] = {1, 0, 2, 0, 2, 1, 3, 0, 3, 1, 3, 2};
        int orig = n;
        for (int k = 0; k < orig * (orig - 1) && n < maxc; k++) {
            const motion_t *c0 = &list[l0i[k]], *c1 = &list[l1i[k]];
            if ((c0->dir & 1) && (c1->dir & 2)) {                /* the two lists reference different pictures: always distinct */
                motion_t m; m.dir = 3; m.mv[0] = c0->mv[0]; m.mv[1] = c1->mv[1];
                list[n++] = m;
            }
        }
    }
    while (n < maxc) { memset(&list[n], 0, sizeof(list[n])); list[n].dir = s->slice_type == KS_SLICE_B ? 3 : 1; n++; }   /* 8.5.3.2.5 zero candidates */
    return n;
}
static mv_t scale_mv(mv_t mv, int tb, int td)
{   /* 8.5.3.2.7 (8-179..8-183) */
    if (td < -128) td = -128; if (td > 127) td = 127; if (tb < -128) tb = -128; if (tb > 127) tb = 127;
    int tx = (16384 + (abs(td) >> 1)) / td;
    int dsf = (tb * tx + 32) >> 6; if (dsf < -4096) dsf = -4096; if (dsf > 4095) dsf = 4095;
    mv_t r; int v;
    v = dsf * mv.x; v = (v < 0 ? -1 : 1) * ((abs(v) + 127) >> 8); r.x = (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v);
    v = dsf * mv.y; v = (v < 0 ? -1 : 1) * ((abs(v) + 127) >> 8); r.y = (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v);
    return r;
}
/* AMVP list of list X for the 2Nx2N PU whose neighbours are in nbs; dpoc[l] = POC(cur) - POC(RefPicList_l[0]) */
static void amvp_list(const nb_set *nbs, int X, const int dpoc[2], mv_t list[2])
{
    const int Y = 1 - X;
    const motion_t *nb = nbs->m; const int *av = nbs->av;      /* A0, A1, B0, B1, B2 */
    mv_t a = {0, 0}, b = {0, 0}; int fa = 0, fb = 0, n = 0;
    for (int k = 0; k < 2 && !fa; k++) if (av[k] && (nb[k].dir & (1 << X))) { a = nb[k].mv[X]; fa = 1; }
    for (int k = 0; k < 2 && !fa; k++) if (av[k]) { a = scale_mv(nb[k].mv[Y], dpoc[X], dpoc[Y]); fa = 1; }       /* only list Y is left: scaled */
    int scaled = av[0] || av[1];
    for (int k = 2; k < 5 && !fb; k++) if (av[k] && (nb[k].dir & (1 << X))) { b = nb[k].mv[X]; fb = 1; }
    if (!scaled && fb) { a = b; fa = 1; }
    if (!scaled) {
        fb = 0;
        for (int k = 2; k < 5 && !fb; k++) if (av[k]) { b = (nb[k].dir & (1 << X)) ? nb[k].mv[X] : scale_mv(nb[k].mv[Y], dpoc[X], dpoc[Y]); fb = 1; }
    }
    if (fa) list[n++] = a;
    if (fb && !(fa && a.x == b.x && a.y == b.y)) list[n++] = b;
    while (n < 2) { list[n].x = 0; list[n].y = 0; n++; }
}
static int mvd_bits(int d) { int a = d < 0 ? -d : d; if (a == 0) return 1; if (a == 1) return 3; int v = a - 2, k = 1, b = 3; while (v >= (1 << k)) { v -= 1 << k; k++; b++; } return b + k + 1; }

static void code_mvd(cabac *c, int dx, int dy)
{   /* 7.3.8.9 */
    int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    cb_bin(c, CX_MVD, ax > 0); cb_bin(c, CX_MVD, ay > 0);
    if (ax > 0) cb_bin(c, CX_MVD + 1, ax > 1);
    if (ay > 0) cb_bin(c, CX_MVD + 1, ay > 1);
    for (int k = 0; k < 2; k++) {
        int a = k ? ay : ax, d = k ? dy : dx;
        if (a > 0) {
            if (a > 1) {  /* abs_mvd_minus2: EG1 */
                int v = a - 2, kk = 1;
                while (v >= (1 << kk)) { cb_bypass(c, 1); v -= 1 << kk; kk++; }
                cb_bypass(c, 0);
                cb_bypass_bins(c, (uint32_t)v, kk);
            }
            cb_bypass(c, d < 0);
        }
    }
}

/* ---- transform tree for one CU (max_transform_hierarchy_depth 0: TU = min(CU,32)) ---- */
static void code_tu(slice_enc *e, int x, int y, int log2, int f)
{
    int xc = x & (KS_CTU - 1), yc = y & (KS_CTU - 1);
    if (f & KS_F_CBF_Y) code_residual(e, 0, xc, yc, log2);
    if (f & KS_F_CBF_CB) code_residual(e, 1, xc >> 1, yc >> 1, log2 - 1);
    if (f & KS_F_CBF_CR) code_residual(e, 2, xc >> 1, yc >> 1, log2 - 1);
}
static void code_transform_tree(slice_enc *e, int x, int y, int cu_log2, int intra)
{
    cabac *c = &e->cb; const ks_frame_syn *s = e->syn;
    if (cu_log2 <= KS_MAX_TB_LOG2) {
        int f = cell_at(s, x, y)->flags;
        cb_bin(c, CX_CBF_CHROMA + 0, (f & KS_F_CBF_CB) != 0);
        cb_bin(c, CX_CBF_CHROMA + 0, (f & KS_F_CBF_CR) != 0);
        if (intra || (f & (KS_F_CBF_CB | KS_F_CBF_CR))) cb_bin(c, CX_CBF_LUMA + 1, (f & KS_F_CBF_Y) != 0);
        code_tu(e, x, y, cu_log2, f);
    } else {                                  /* 64x64 CU: split inferred, four 32x32 TUs */
        int fq[4], cb_any = 0, cr_any = 0;
        for (int k = 0; k < 4; k++) {
            fq[k] = cell_at(s, x + (k & 1) * 32, y + (k >> 1) * 32)->flags;
            cb_any |= fq[k] & KS_F_CBF_CB; cr_any |= fq[k] & KS_F_CBF_CR;
        }
        cb_bin(c, CX_CBF_CHROMA + 0, cb_any != 0);
        cb_bin(c, CX_CBF_CHROMA + 0, cr_any != 0);
        for (int k = 0; k < 4; k++) {
            int xx = x + (k & 1) * 32, yy = y + (k >> 1) * 32;
            if (xx >= s->width || yy >= s->height) continue;   /* cannot happen for a 64 CU inside the picture */
            if (cb_any) cb_bin(c, CX_CBF_CHROMA + 1, (fq[k] & KS_F_CBF_CB) != 0);
            if (cr_any) cb_bin(c, CX_CBF_CHROMA + 1, (fq[k] & KS_F_CBF_CR) != 0);
            cb_bin(c, CX_CBF_LUMA + 0, (fq[k] & KS_F_CBF_Y) != 0);
            code_tu(e, xx, yy, 5, fq[k]);
        }
    }
}

static int cu_cbf_any(const ks_frame_syn *s, int x, int y, int size)
{
    int any = 0;
    for (int yy = y; yy < y + size; yy += KS_CELL) for (int xx = x; xx < x + size; xx += KS_CELL)
        any |= cell_at(s, xx, yy)->flags & (KS_F_CBF_Y | KS_F_CBF_CB | KS_F_CBF_CR);
    return any != 0;
}

/* P slices: merge_idx of the CU's own vector without materialising the candidate list.  A motion is one 32-bit word (mvx | mvy << 16, one
 * list, one reference), candidates are produced in list order A1, B1, B0, A0, B2, zero... with the 8.5.3.2.3 prunings, and the walk stops at
 * the first candidate equal to `cur` (typically A1) or when maxc candidates exist.  Same result as merge_list() + search. */
static inline uint32_t cell_mvw(const ks_cell *c) { uint32_t w; memcpy(&w, &c->mvx, 4); return w; }
static int merge_idx_p(const ks_frame_syn *s, int x, int y, int size, int maxc, uint32_t cur)
{
    int n = 0;
    uint32_t a1 = 0, b1 = 0, b0 = 0, a0 = 0, b2;
    int fa1 = 0, ab1 = 0, fb1, fb0 = 0, fa0 = 0, fb2 = 0;
    if (x > 0) { const ks_cell *c = cell_at(s, x - 1, y + size - 1); if (!(c->flags & KS_F_INTRA)) { fa1 = 1; a1 = cell_mvw(c); } }
    if (fa1) { if (a1 == cur) return n; if (++n == maxc) return -1; }
    if (y > 0) { const ks_cell *c = cell_at(s, x + size - 1, y - 1); if (!(c->flags & KS_F_INTRA)) { ab1 = 1; b1 = cell_mvw(c); } }
    fb1 = ab1 && !(fa1 && a1 == b1);
    if (fb1) { if (b1 == cur) return n; if (++n == maxc) return -1; }
    if (avail(s, x, y, x + size, y - 1)) { const ks_cell *c = cell_at(s, x + size, y - 1); if (!(c->flags & KS_F_INTRA)) { fb0 = 1; b0 = cell_mvw(c); } }
    if (fb0 && ab1 && b1 == b0) fb0 = 0;
    if (fb0) { if (b0 == cur) return n; if (++n == maxc) return -1; }
    if (avail(s, x, y, x - 1, y + size)) { const ks_cell *c = cell_at(s, x - 1, y + size); if (!(c->flags & KS_F_INTRA)) { fa0 = 1; a0 = cell_mvw(c); } }
    if (fa0 && fa1 && a1 == a0) fa0 = 0;
    if (fa0) { if (a0 == cur) return n; if (++n == maxc) return -1; }
    if (x > 0 && y > 0 && fa0 + fa1 + fb0 + fb1 != 4) {
        const ks_cell *c = cell_at(s, x - 1, y - 1);
        if (!(c->flags & KS_F_INTRA)) { b2 = cell_mvw(c); fb2 = !(fa1 && a1 == b2) && !(ab1 && b1 == b2); if (fb2) { if (b2 == cur) return n; if (++n == maxc) return -1; } }
    }
    return cur == 0 ? n : -1;             /* zero candidates fill the rest of the list: the first of them sits at index n */
}

/* luma intra mode of the prediction block covering (x,y) as a NEIGHBOUR sees it (8.4.2): DC for non-intra blocks */
static inline int intra_mode_at(const ks_frame_syn *s, int x, int y)
{
    const ks_cell *n = cell_at(s, x, y);
    if (!(n->flags & KS_F_INTRA)) return 1;
    return n->cu_log2 == 3 ? KS_SUB_MODE(n, ((x >> 3) & 1) | (((y >> 3) & 1) << 1)) : n->intra_mode;
}
/* prev_intra_luma_pred_flag / mpm_idx / rem_intra_luma_pred_mode of the block at (x,y) (7.3.8.5, 8.4.2) */
static void code_intra_mode(slice_enc *e, int x, int y, int mode)
{
    cabac *c = &e->cb; const ks_frame_syn *s = e->syn;
    int cand_a = 1, cand_b = 1;
    if (x > 0) cand_a = intra_mode_at(s, x - 1, y);
    if (y > 0 && ((y - 1) >> KS_CTU_LOG2) == (y >> KS_CTU_LOG2)) cand_b = intra_mode_at(s, x, y - 1);
    int mpm[3];
    if (cand_a == cand_b) {
        if (cand_a < 2) { mpm[0] = 0; mpm[1] = 1; mpm[2] = 26; }
        else { mpm[0] = cand_a; mpm[1] = 2 + ((cand_a + 29) & 31); mpm[2] = 2 + ((cand_a - 2 + 1) & 31); }
    } else {
        mpm[0] = cand_a; mpm[1] = cand_b;
        mpm[2] = (cand_a != 0 && cand_b != 0) ? 0 : (cand_a != 1 && cand_b != 1) ? 1 : 26;
    }
    int mi = -1;
    for (int k = 0; k < 3; k++) if (mpm[k] == mode) mi = k;
    cb_bin(c, CX_PREV_INTRA, mi >= 0);
    if (mi >= 0) { cb_bypass(c, mi > 0); if (mi > 0) cb_bypass(c, mi > 1); }
    else {
        if (mpm[0] > mpm[1]) { int t = mpm[0]; mpm[0] = mpm[1]; mpm[1] = t; }
        if (mpm[0] > mpm[2]) { int t = mpm[0]; mpm[0] = mpm[2]; mpm[2] = t; }
        if (mpm[1] > mpm[2]) { int t = mpm[1]; mpm[1] = mpm[2]; mpm[2] = t; }
        int rem = mode;
        for (int k = 2; k >= 0; k--) if (rem > mpm[k]) rem--;
        cb_bypass_bins(c, (uint32_t)rem, 5);
    }
}

/* ---- 7.3.8.5 coding_unit ---- */
static void code_cu(slice_enc *e, int x, int y, int log2)
{
    cabac *c = &e->cb; const ks_frame_syn *s = e->syn;
    const ks_cell *cu = cell_at(s, x, y);
    int size = 1 << log2, intra = cu->flags & KS_F_INTRA;
    if (s->slice_type != KS_SLICE_I) {
        motion_t ml[5], cur; nb_set nbs; int merge_idx = -1;
        if (!intra) {
            cur = cell_motion(s, x, y);
            if (!s->cells_b) merge_idx = merge_idx_p(s, x, y, size, e->sp->max_merge_cand, cell_mvw(cu));     /* P slice fast path */
            else {
                fetch_nb(s, x, y, size, &nbs);
                int n = merge_list(s, &nbs, e->sp->max_merge_cand, ml);
                for (int k = 0; k < n; k++) if (motion_eq(&ml[k], &cur)) { merge_idx = k; break; }
            }
        }
        int any = intra ? 1 : cu_cbf_any(s, x, y, size);
        int skip = !intra && merge_idx >= 0 && !any;
        int ctx = 0;
        if (x > 0) ctx += e->skip[(y >> KS_CELL_LOG2) * s->cells_w + ((x - 1) >> KS_CELL_LOG2)];      /* left / above always precede in z-scan */
        if (y > 0) ctx += e->skip[((y - 1) >> KS_CELL_LOG2) * s->cells_w + (x >> KS_CELL_LOG2)];
        cb_bin(c, CX_SKIP + ctx, skip);
        for (int yy = y; yy < y + size; yy += KS_CELL) for (int xx = x; xx < x + size; xx += KS_CELL)
            e->skip[(yy >> KS_CELL_LOG2) * s->cells_w + (xx >> KS_CELL_LOG2)] = (uint8_t)skip;
        if (skip || (!intra && merge_idx >= 0)) {
            if (!skip) { cb_bin(c, CX_PRED_MODE, 0); cb_bin(c, CX_PART_MODE, 1); cb_bin(c, CX_MERGE_FLAG, 1); }
            if (e->sp->max_merge_cand > 1) {         /* merge_idx: TR, first bin context coded */
                cb_bin(c, CX_MERGE_IDX, merge_idx != 0);
                if (merge_idx != 0) for (int k = 1; k < e->sp->max_merge_cand - 1; k++) { int b = merge_idx != k; cb_bypass(c, b); if (!b) break; }
            }
            if (skip) return;
            code_transform_tree(e, x, y, log2, 0);   /* merge 2Nx2N: rqt_root_cbf inferred 1 */
            return;
        }
        cb_bin(c, CX_PRED_MODE, intra != 0);
        if (!intra) {
            cb_bin(c, CX_PART_MODE, 1);              /* PART_2Nx2N */
            cb_bin(c, CX_MERGE_FLAG, 0);
            if (s->slice_type == KS_SLICE_B) {       /* inter_pred_idc (9.3.4.2.2): nPbW+nPbH != 12 */
                int depth = KS_CTU_LOG2 - log2;
                cb_bin(c, CX_INTER_DIR + depth, cur.dir == 3);
                if (cur.dir != 3) cb_bin(c, CX_INTER_DIR + 4, cur.dir == 2);
            }
            const int dpoc[2] = {-e->sl->neg_delta_poc[0], -e->sl->pos_delta_poc[0]};
            if (!s->cells_b) fetch_nb(s, x, y, size, &nbs);      /* the P fast path above did not need the neighbour set */
            for (int X = 0; X < 2; X++) {
                if (!(cur.dir & (1 << X))) continue;
                mv_t pl[2]; amvp_list(&nbs, X, dpoc, pl);
                int c0 = mvd_bits(cur.mv[X].x - pl[0].x) + mvd_bits(cur.mv[X].y - pl[0].y);
                int c1 = mvd_bits(cur.mv[X].x - pl[1].x) + mvd_bits(cur.mv[X].y - pl[1].y);
                int idx = c1 < c0;
                code_mvd(c, cur.mv[X].x - pl[idx].x, cur.mv[X].y - pl[idx].y);
                cb_bin(c, CX_MVP_IDX, idx);
            }
            cb_bin(c, CX_ROOT_CBF, any);
            if (any) code_transform_tree(e, x, y, log2, 0);
            return;
        }
    }
    /* intra 2Nx2N (part_mode is only present at the minimum CU size, 8x8: code_cu8) */
    code_intra_mode(e, x, y, cu->intra_mode);
    cb_bin(c, CX_CHROMA_PRED, 0);                     /* intra_chroma_pred_mode = 4 (DM) */
    code_transform_tree(e, x, y, log2, 1);
}

/* ---- one 8x8 intra CU (sub-block k of an intra cell with cu_log2 == 3): 2Nx2N, luma TB 8x8, chroma TBs 4x4, mode-dependent scans ---- */
static void code_cu8(slice_enc *e, int x, int y, int k)
{
    cabac *c = &e->cb; const ks_frame_syn *s = e->syn;
    const ks_cell *cu = cell_at(s, x, y);
    if (s->slice_type != KS_SLICE_I) {
        int ctx = 0;          /* the cell's skip mark is 0 (intra): neighbours inside the cell count as not skipped */
        if (x > 0) ctx += e->skip[(y >> KS_CELL_LOG2) * s->cells_w + ((x - 1) >> KS_CELL_LOG2)];
        if (y > 0) ctx += e->skip[((y - 1) >> KS_CELL_LOG2) * s->cells_w + (x >> KS_CELL_LOG2)];
        cb_bin(c, CX_SKIP + ctx, 0);
        e->skip[(y >> KS_CELL_LOG2) * s->cells_w + (x >> KS_CELL_LOG2)] = 0;
        cb_bin(c, CX_PRED_MODE, 1);
    }
    cb_bin(c, CX_PART_MODE, 1);                       /* PART_2Nx2N */
    const int mode = KS_SUB_MODE(cu, k), scan_idx = intra_scan_idx(mode);
    code_intra_mode(e, x, y, mode);
    cb_bin(c, CX_CHROMA_PRED, 0);
    const int fy = KS_SUB_CBF_Y(cu, k), fcb = KS_SUB_CBF_CB(cu, k), fcr = KS_SUB_CBF_CR(cu, k);
    cb_bin(c, CX_CBF_CHROMA + 0, fcb);
    cb_bin(c, CX_CBF_CHROMA + 0, fcr);
    cb_bin(c, CX_CBF_LUMA + 1, fy);
    const int xc = x & (KS_CTU - 1), yc = y & (KS_CTU - 1);
    if (fy) code_residual_small(e, 0, xc, yc, 3, scan_idx);
    if (fcb) code_residual_small(e, 1, xc >> 1, yc >> 1, 2, scan_idx);
    if (fcr) code_residual_small(e, 2, xc >> 1, yc >> 1, 2, scan_idx);
}

/* ---- 7.3.8.4 coding_quadtree ---- */
static void code_quadtree(slice_enc *e, int x, int y, int log2)
{
    const ks_frame_syn *s = e->syn; cabac *c = &e->cb;
    int size = 1 << log2, split;
    if (x + size <= s->width && y + size <= s->height && log2 > KS_MIN_CB_LOG2) {
        split = cell_at(s, x, y)->cu_log2 < log2;
        int depth = KS_CTU_LOG2 - log2, ctx = 0;
        if (x > 0) ctx += (KS_CTU_LOG2 - cell_at(s, x - 1, y)->cu_log2) > depth;
        if (y > 0) ctx += (KS_CTU_LOG2 - cell_at(s, x, y - 1)->cu_log2) > depth;
        cb_bin(c, CX_SPLIT_CU + ctx, split);
    } else split = log2 > KS_CELL_LOG2;              /* coded sizes are multiples of 16: a CU of 16 or less never crosses the picture edge */
    if (split && log2 == KS_CELL_LOG2) {
        for (int k = 0; k < 4; k++) code_cu8(e, x + (k & 1) * 8, y + (k >> 1) * 8, k);
    } else if (split) {
        int h = size >> 1;
        for (int k = 0; k < 4; k++) {
            int xx = x + (k & 1) * h, yy = y + (k >> 1) * h;
            if (xx < s->width && yy < s->height) code_quadtree(e, xx, yy, log2 - 1);
        }
    } else code_cu(e, x, y, log2);
}

/* ---- 7.3.8.3 sao ---- */
static int sao_equal(const ks_ctu_syn *a, const ks_ctu_syn *b) { return memcmp(a->sao, b->sao, sizeof(a->sao)) == 0; }
static void code_sao(slice_enc *e, int rx, int ry)
{
    cabac *c = &e->cb; const ks_frame_syn *s = e->syn; const ks_slice_params *sl = e->sl;
    const ks_ctu_syn *ct = &s->ctus[ry * s->ctus_w + rx];
    if (!sl->sao_luma && !sl->sao_chroma) return;
    if (rx > 0) { int m = sao_equal(ct, ct - 1); cb_bin(c, CX_SAO_MERGE, m); if (m) return; }
    if (ry > 0) { int m = sao_equal(ct, ct - s->ctus_w); cb_bin(c, CX_SAO_MERGE, m); if (m) return; }
    for (int ci = 0; ci < 3; ci++) {
        if ((ci == 0 && !sl->sao_luma) || (ci > 0 && !sl->sao_chroma)) continue;
        const ks_sao_param *p = &ct->sao[ci];
        int type = ci == 2 ? ct->sao[1].type : p->type;
        if (ci < 2) { cb_bin(c, CX_SAO_TYPE, type != 0); if (type) cb_bypass(c, type == 2); }
        if (!type) continue;
        for (int k = 0; k < 4; k++) {     /* sao_offset_abs: TR cMax 7 */
            int a = p->off[k] < 0 ? -p->off[k] : p->off[k];
            for (int j = 0; j < a; j++) cb_bypass(c, 1);
            if (a < 7) cb_bypass(c, 0);
        }
        if (type == 1) {
            for (int k = 0; k < 4; k++) if (p->off[k]) cb_bypass(c, p->off[k] < 0);
            cb_bypass_bins(c, p->band_or_class, 5);
        } else if (ci < 2) cb_bypass_bins(c, ci == 0 ? p->band_or_class : ct->sao[1].band_or_class, 2);
    }
}

size_t ks_slice_scratch_bytes(const ks_stream_params *sp)
{
    size_t cells = (size_t)(sp->width >> KS_CELL_LOG2) * (sp->height >> KS_CELL_LOG2);
    /* slice RBSP scratch: 3 bytes per luma sample (CABAC on noise at QP 0 measures 2.1) */
    return sizeof(slice_enc) + cells + (size_t)sp->width * sp->height * 3 + 65536;
}

long ks_write_slice(const ks_stream_params *sp, const ks_slice_params *sl, const ks_frame_syn *syn,
                    void *scratch, uint8_t *out, size_t cap)
{
    scans_init();
    slice_enc *e = (slice_enc *)scratch;
    memset(e, 0, sizeof(*e));
    e->sp = sp; e->sl = sl; e->syn = syn;
    size_t cells = (size_t)syn->cells_w * syn->cells_h;
    e->skip = (uint8_t *)(e + 1);
    memset(e->skip, 0, cells);
    uint8_t *rbsp = e->skip + cells;
    size_t rcap = (size_t)sp->width * sp->height * 3 + 65536 - 1024;

    /* ---- 7.3.6.1 slice_segment_header ---- */
    bitw b; bw_init(&b, rbsp, 512);
    int idr = sl->nal_type == 19 || sl->nal_type == 20;
    bw_put(&b, 1, 1);
    if (sl->nal_type >= 16 && sl->nal_type <= 23) bw_put(&b, 0, 1);
    bw_ue(&b, 0);
    bw_ue(&b, (uint32_t)sl->slice_type);
    if (!idr) {
        bw_put(&b, (uint32_t)sl->poc & ((1u << sp->log2_max_poc_lsb) - 1), sp->log2_max_poc_lsb);
        bw_put(&b, 0, 1);                        /* short_term_ref_pic_set_sps_flag = 0: explicit RPS (like the reference) */
        bw_ue(&b, (uint32_t)sl->num_neg_refs); bw_ue(&b, (uint32_t)sl->num_pos_refs);
        int prev = 0;
        for (int i = 0; i < sl->num_neg_refs; i++) { bw_ue(&b, (uint32_t)(prev - sl->neg_delta_poc[i] - 1)); bw_put(&b, 1, 1); prev = sl->neg_delta_poc[i]; }
        prev = 0;
        for (int i = 0; i < sl->num_pos_refs; i++) { bw_ue(&b, (uint32_t)(sl->pos_delta_poc[i] - prev - 1)); bw_put(&b, 1, 1); prev = sl->pos_delta_poc[i]; }
    }
    if (sp->sao) { bw_put(&b, (uint32_t)sl->sao_luma, 1); bw_put(&b, (uint32_t)sl->sao_chroma, 1); }
    if (sl->slice_type != KS_SLICE_I) {
        bw_put(&b, 1, 1); bw_ue(&b, 0);          /* num_ref_idx_active_override: 1 active ref (reference does the same) */
        if (sl->slice_type == KS_SLICE_B) { bw_ue(&b, 0); bw_put(&b, 0, 1); }      /* num_ref_idx_l1_active_minus1 = 0, mvd_l1_zero_flag = 0 */
        bw_ue(&b, (uint32_t)(5 - sp->max_merge_cand));
    }
    bw_se(&b, sl->qp - 26);
    bw_put(&b, (uint32_t)sl->deblock_override, 1);
    if (sl->deblock_override) { bw_put(&b, 0, 1); bw_se(&b, sl->beta_offset_div2); bw_se(&b, sl->tc_offset_div2); }
    bw_put(&b, 1, 1);                            /* slice_loop_filter_across_slices_enabled_flag */
    bw_trailing(&b);                             /* byte_alignment() has the same form */
    size_t hdr = b.pos;

    /* ---- slice_segment_data ---- */
    int init_type = sl->slice_type == KS_SLICE_I ? 0 : (sl->slice_type == KS_SLICE_P ? 1 : 2);
    cb_init(&e->cb, rbsp + hdr, rcap - hdr, init_type, sl->qp);
    int nctu = syn->ctus_w * syn->ctus_h;
    for (int a = 0; a < nctu; a++) {
        int rx = a % syn->ctus_w, ry = a / syn->ctus_w;
        ctu_prefix(e, a);
        code_sao(e, rx, ry);
        code_quadtree(e, rx << KS_CTU_LOG2, ry << KS_CTU_LOG2, KS_CTU_LOG2);
        cb_terminate(&e->cb, a == nctu - 1);
    }
    cb_finish(&e->cb);
    if (e->cb.overflow) return -1;
    return nal_emit(out, cap, sl->nal_type, sl->temporal_id, rbsp, hdr + e->cb.pos);
}

uint64_t ks_plane_sse(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
    uint64_t s = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++) { int d = (int)a[x] - (int)b[x]; s += (uint64_t)(d * d); }
    return s;
}
