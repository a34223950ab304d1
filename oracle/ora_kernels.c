/* ora_kernels.c -- CPU restatement of the reference's leaf kernels (TEST INFRASTRUCTURE, see ks_oracle.h).
 * Each function cites the symbol in /root/reference/centos_x64/appencoder it restates. */
#include "ks_oracle.h"
#include <stdlib.h>
#include <string.h>

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* ------------------------------------------------------------------ a1/a2 SAD -------------------- */
uint32_t ora_sad(const uint8_t *a, const uint8_t *b, long sa, long sb, long h, long w)
{   /* sad_c E@0x473db0 */
    uint32_t s = 0;
    for (long y = 0; y < h; y++, a += sa, b += sb)
        for (long x = 0; x < w; x++) s += (uint32_t)iabs((int)a[x] - (int)b[x]);
    return s;
}
void ora_sad4(const uint8_t *src, const uint8_t *ref, long ss, long sr, long h, uint32_t out[4], long w)
{   /* sad4_c E@0x473e30: x264 COST_MV_X4_DIR order (0,-1),(0,+1),(-1,0),(+1,0); results << 4 */
    out[0] = ora_sad(src, ref - sr, ss, sr, h, w) << 4;
    out[1] = ora_sad(src, ref + sr, ss, sr, h, w) << 4;
    out[2] = ora_sad(src, ref - 1, ss, sr, h, w) << 4;
    out[3] = ora_sad(src, ref + 1, ss, sr, h, w) << 4;
}
void ora_sad3(const uint8_t *src, const uint8_t *r0, const uint8_t *r1, const uint8_t *r2,
              long ss, long sr, long h, uint32_t out[3], long w)
{   /* sad3_c E@0x474070 */
    out[0] = ora_sad(src, r0, ss, sr, h, w);
    out[1] = ora_sad(src, r1, ss, sr, h, w);
    out[2] = ora_sad(src, r2, ss, sr, h, w);
}
uint32_t ora_sse(const uint8_t *a, const uint8_t *b, int sa, int sb, int n)
{   /* sse_c<N> E@0x474d70.. */
    uint32_t s = 0;
    for (int y = 0; y < n; y++, a += sa, b += sb)
        for (int x = 0; x < n; x++) { int d = (int)a[x] - (int)b[x]; s += (uint32_t)(d * d); }
    return s;
}
/* ------------------------------------------------------------------ a15 SATD --------------------- */
static uint32_t hadamard_abs_sum(const uint8_t *a, const uint8_t *b, long sa, long sb, int n)
{
    int m[64], t[64];
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) m[y * n + x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    /* rows */
    for (int y = 0; y < n; y++) {
        int *r = m + y * n;
        for (int len = 1; len < n; len <<= 1)
            for (int i = 0; i < n; i += len << 1)
                for (int j = i; j < i + len; j++) { int u = r[j], v = r[j + len]; r[j] = u + v; r[j + len] = u - v; }
    }
    /* columns */
    for (int x = 0; x < n; x++) {
        for (int y = 0; y < n; y++) t[y] = m[y * n + x];
        for (int len = 1; len < n; len <<= 1)
            for (int i = 0; i < n; i += len << 1)
                for (int j = i; j < i + len; j++) { int u = t[j], v = t[j + len]; t[j] = u + v; t[j + len] = u - v; }
        for (int y = 0; y < n; y++) m[y * n + x] = t[y];
    }
    uint32_t s = 0;
    for (int i = 0; i < n * n; i++) s += (uint32_t)iabs(m[i]);
    return s;
}
uint32_t ora_satd(const uint8_t *a, const uint8_t *b, long sa, long sb, long h, long w)
{   /* had_c E@0x474500 -> xCalcHADs8x8 E@0x474200 (HM xCalcHADs): 8x8 tiles when both dims are
     * multiples of 8 ((sum+2)>>2 each), else 4x4 tiles ((sum+1)>>1 each) */
    uint32_t s = 0;
    if ((w & 7) == 0 && (h & 7) == 0) {
        for (long y = 0; y < h; y += 8)
            for (long x = 0; x < w; x += 8) s += (hadamard_abs_sum(a + y * sa + x, b + y * sb + x, sa, sb, 8) + 2) >> 2;
    } else {
        for (long y = 0; y < h; y += 4)
            for (long x = 0; x < w; x += 4) s += (hadamard_abs_sum(a + y * sa + x, b + y * sb + x, sa, sb, 4) + 1) >> 1;
    }
    return s;
}

/* ------------------------------------------------------------------ a7 interpolation ------------- */
#define TAPS_H8(c, s, x)  ((c)[0]*(s)[(x)-3] + (c)[1]*(s)[(x)-2] + (c)[2]*(s)[(x)-1] + (c)[3]*(s)[(x)] + \
                           (c)[4]*(s)[(x)+1] + (c)[5]*(s)[(x)+2] + (c)[6]*(s)[(x)+3] + (c)[7]*(s)[(x)+4])
#define TAPS_V8(c, s, x, st) ((c)[0]*(s)[(x)-3*(st)] + (c)[1]*(s)[(x)-2*(st)] + (c)[2]*(s)[(x)-(st)] + (c)[3]*(s)[(x)] + \
                           (c)[4]*(s)[(x)+(st)] + (c)[5]*(s)[(x)+2*(st)] + (c)[6]*(s)[(x)+3*(st)] + (c)[7]*(s)[(x)+4*(st)])
#define TAPS_H4(c, s, x)  ((c)[0]*(s)[(x)-1] + (c)[1]*(s)[(x)] + (c)[2]*(s)[(x)+1] + (c)[3]*(s)[(x)+2])
#define TAPS_V4(c, s, x, st) ((c)[0]*(s)[(x)-(st)] + (c)[1]*(s)[(x)] + (c)[2]*(s)[(x)+(st)] + (c)[3]*(s)[(x)+2*(st)])

void ora_interp_luma_h_8to8(uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_luma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = clip8((TAPS_H8(c, src, x) + 32) >> 6); }
void ora_interp_luma_h_8to16(int16_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_luma_filter[frac];   /* raw 14-bit sum, no -8192 offset */
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = (int16_t)TAPS_H8(c, src, x); }
void ora_interp_luma_v_8to8(uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_luma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = clip8((TAPS_V8(c, src, x, ss) + 32) >> 6); }
void ora_interp_luma_v_8to16(int16_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_luma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = (int16_t)TAPS_V8(c, src, x, ss); }
void ora_interp_luma_v_16to8(uint8_t *dst, int ds, const int16_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_luma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = clip8((TAPS_V8(c, src, x, ss) + 2048) >> 12); }
void ora_interp_luma_v_16to16(int16_t *dst, int ds, const int16_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_luma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = (int16_t)(TAPS_V8(c, src, x, ss) >> 6); }
void ora_interp_chroma_h_8to8(uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_chroma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = clip8((TAPS_H4(c, src, x) + 32) >> 6); }
void ora_interp_chroma_h_8to16(int16_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_chroma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = (int16_t)TAPS_H4(c, src, x); }
void ora_interp_chroma_v_8to8(uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_chroma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = clip8((TAPS_V4(c, src, x, ss) + 32) >> 6); }
void ora_interp_chroma_v_8to16(int16_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_chroma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = (int16_t)TAPS_V4(c, src, x, ss); }
void ora_interp_chroma_v_16to8(uint8_t *dst, int ds, const int16_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_chroma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = clip8((TAPS_V4(c, src, x, ss) + 2048) >> 12); }
void ora_interp_chroma_v_16to16(int16_t *dst, int ds, const int16_t *src, int ss, int w, int h, int frac)
{   const int8_t *c = ora_chroma_filter[frac];
    for (int y = 0; y < h; y++, dst += ds, src += ss) for (int x = 0; x < w; x++) dst[x] = (int16_t)(TAPS_V4(c, src, x, ss) >> 6); }

/* composition used by interpolatePuLxLuma E@0x487260 (and spec 8.5.3.3.3.1) */
void ora_mc_luma(uint8_t *dst, int ds, const uint8_t *ref, int rs, int w, int h, int mvx, int mvy)
{
    int fx = mvx & 3, fy = mvy & 3;
    const uint8_t *p = ref + (mvy >> 2) * rs + (mvx >> 2);
    if (!fx && !fy) { for (int y = 0; y < h; y++) memcpy(dst + y * ds, p + y * rs, (size_t)w); return; }
    if (!fy) { ora_interp_luma_h_8to8(dst, ds, p, rs, w, h, fx); return; }
    if (!fx) { ora_interp_luma_v_8to8(dst, ds, p, rs, w, h, fy); return; }
    int16_t tmp[(64 + 7) * 64];
    ora_interp_luma_h_8to16(tmp, 64, p - 3 * rs, rs, w, h + 7, fx);
    ora_interp_luma_v_16to8(dst, ds, tmp + 3 * 64, 64, w, h, fy);
}
void ora_mc_chroma(uint8_t *dst, int ds, const uint8_t *ref, int rs, int w, int h, int mvx, int mvy)
{   /* chroma mv in 1/8 sample units of the chroma plane (4:2:0: same number as the luma quarter-pel mv) */
    int fx = mvx & 7, fy = mvy & 7;
    const uint8_t *p = ref + (mvy >> 3) * rs + (mvx >> 3);
    if (!fx && !fy) { for (int y = 0; y < h; y++) memcpy(dst + y * ds, p + y * rs, (size_t)w); return; }
    if (!fy) { ora_interp_chroma_h_8to8(dst, ds, p, rs, w, h, fx); return; }
    if (!fx) { ora_interp_chroma_v_8to8(dst, ds, p, rs, w, h, fy); return; }
    int16_t tmp[(32 + 3) * 32];
    ora_interp_chroma_h_8to16(tmp, 32, p - rs, rs, w, h + 3, fx);
    ora_interp_chroma_v_16to8(dst, ds, tmp + 32, 32, w, h, fy);
}
void ora_mc_luma_16(int16_t *dst, int ds, const uint8_t *ref, int rs, int w, int h, int mvx, int mvy)
{
    int fx = mvx & 3, fy = mvy & 3;
    const uint8_t *p = ref + (mvy >> 2) * rs + (mvx >> 2);
    if (!fx && !fy) { for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) dst[y * ds + x] = (int16_t)(p[y * rs + x] << 6); return; } /* InterpolateCopy8to16_c E@0x435010 */
    if (!fy) { ora_interp_luma_h_8to16(dst, ds, p, rs, w, h, fx); return; }
    if (!fx) { ora_interp_luma_v_8to16(dst, ds, p, rs, w, h, fy); return; }
    int16_t tmp[(64 + 7) * 64];
    ora_interp_luma_h_8to16(tmp, 64, p - 3 * rs, rs, w, h + 7, fx);
    ora_interp_luma_v_16to16(dst, ds, tmp + 3 * 64, 64, w, h, fy);
}
void ora_mc_chroma_16(int16_t *dst, int ds, const uint8_t *ref, int rs, int w, int h, int mvx, int mvy)
{
    int fx = mvx & 7, fy = mvy & 7;
    const uint8_t *p = ref + (mvy >> 3) * rs + (mvx >> 3);
    if (!fx && !fy) { for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) dst[y * ds + x] = (int16_t)(p[y * rs + x] << 6); return; }
    if (!fy) { ora_interp_chroma_h_8to16(dst, ds, p, rs, w, h, fx); return; }
    if (!fx) { ora_interp_chroma_v_8to16(dst, ds, p, rs, w, h, fy); return; }
    int16_t tmp[(32 + 3) * 32];
    ora_interp_chroma_h_8to16(tmp, 32, p - rs, rs, w, h + 3, fx);
    ora_interp_chroma_v_16to16(dst, ds, tmp + 32, 32, w, h, fy);
}
void ora_weighted_bi(uint8_t *dst, int ds, const int16_t *p0, const int16_t *p1, int ps, int w, int h)
{   /* DefaultWeightedBi_c E@0x4350f0: (p0+p1+64)>>7 */
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) dst[y * ds + x] = clip8((p0[y * ps + x] + p1[y * ps + x] + 64) >> 7);
}

/* ------------------------------------------------------------------ a8..a13 residual path -------- */
void ora_residual(int16_t *res, const uint8_t *src, const uint8_t *pred, int ss, int ps, int n)
{   /* calc_residual_N_sse2 / H265_CalResidual<N> */
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) res[y * n + x] = (int16_t)((int)src[y * ss + x] - (int)pred[y * ps + x]);
}
static inline int tcoef(int log2n, int is_dst, int k, int i)
{
    return is_dst ? ora_dst4[k][i] : ora_dct32[k << (5 - log2n)][i];
}
void ora_fdct(const int16_t *src, int16_t *dst, int src_stride, int dst_stride, int log2n, int is_dst)
{   /* H265_2dDct{4,8,16,32}_c E@0x4b7600/76c0/7720/7780, H265_2dDst4x4_c E@0x4b7660.
     * pass 1 transforms rows (horizontal), shift 2*log2N-2; pass 2 columns, shift 7; each result
     * rounded half-up and stored as int16 (NOT the HM shifts log2N-1 / log2N+6). */
    int n = 1 << log2n, s1 = 2 * log2n - 2, s2 = 7;
    int16_t tmp[32 * 32];
    for (int y = 0; y < n; y++)
        for (int u = 0; u < n; u++) {
            int acc = 0;
            for (int x = 0; x < n; x++) acc += tcoef(log2n, is_dst, u, x) * src[y * src_stride + x];
            tmp[u * n + y] = (int16_t)((acc + (1 << (s1 - 1))) >> s1);
        }
    for (int u = 0; u < n; u++)
        for (int v = 0; v < n; v++) {
            int acc = 0;
            for (int y = 0; y < n; y++) acc += tcoef(log2n, is_dst, v, y) * tmp[u * n + y];
            dst[v * dst_stride + u] = (int16_t)((acc + (1 << (s2 - 1))) >> s2);
        }
}
int ora_quant_block(const int16_t *coef, int16_t *dst, int stride, int scale, int add, int qbits, int n, int16_t *delta_u)
{   /* H265QuantBlock_c E@0x4a2580 (HM xQuant) */
    int nnz = 0;
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
            int c = coef[y * stride + x], a = iabs(c);
            int t = a * scale;
            int level = (t + add) >> qbits;
            if (delta_u) delta_u[y * stride + x] = (int16_t)((t - (level << qbits)) >> (qbits - 8));
            if (level > 32767) level = 32767;
            if (level) nnz++;
            dst[y * stride + x] = (int16_t)(c < 0 ? -level : level);
        }
    return nnz;
}
int ora_quant(const int16_t *coef, int16_t *dst, int stride, int qp, int log2n, int is_intra_slice, int16_t *delta_u)
{   /* parameter derivation at the call site in `reconstruct` E@0x47da2f..0x47da92 + H265_GetBaseQuantParam E@0x4a2520 */
    int qbits = 21 + qp / 6 - log2n;
    int add = (is_intra_slice ? 171 : 85) << (qbits - 9);
    return ora_quant_block(coef, dst, stride, ora_quant_scales[qp % 6], add, qbits, 1 << log2n, delta_u);
}
int ora_sign_hide(const int16_t *coef, int16_t *level, const int16_t *delta_u, int stride, int log2n, const uint16_t *scan)
{   /* signBitHidingHDQ E@0x4a29c0 (HM TComTrQuant::signBitHidingHDQ); scan[i] = (y<<8)|x of scan position i */
    int n = 1 << log2n, last_cg = -1;
#define POS(i) ((scan[i] >> 8) * stride + (scan[i] & 255))
    for (int sub = (n * n - 1) >> 4; sub >= 0; sub--) {
        int sp = sub << 4, first = 16, last = -1, sum = 0;
        for (int i = 15; i >= 0; i--) if (level[POS(sp + i)]) { last = i; break; }
        for (int i = 0; i < 16; i++) if (level[POS(sp + i)]) { first = i; break; }
        for (int i = first; i <= last; i++) sum += level[POS(sp + i)];
        if (last >= 0 && last_cg == -1) last_cg = 1;
        if (last - first >= 4) {
            int signbit = level[POS(sp + first)] > 0 ? 0 : 1;
            if (signbit != (sum & 1)) {
                int min_cost = 0x7fffffff, min_pos = -1, final_change = 0;
                for (int i = (last_cg == 1 ? last : 15); i >= 0; i--) {
                    int p = POS(sp + i), cost, change = 0;
                    if (level[p] != 0) {
                        if (delta_u[p] > 0) { cost = -delta_u[p]; change = 1; }
                        else if (i == first && iabs(level[p]) == 1) cost = 0x7fffffff;
                        else { cost = delta_u[p]; change = -1; }
                    } else if (i < first) {
                        int this_sign = coef[p] >= 0 ? 0 : 1;
                        if (this_sign != signbit) cost = 0x7fffffff;
                        else { cost = -delta_u[p]; change = 1; }
                    } else { cost = -delta_u[p]; change = 1; }
                    if (cost < min_cost) { min_cost = cost; final_change = change; min_pos = p; }
                }
                if (level[min_pos] == 32767 || level[min_pos] == -32768) final_change = -1;
                if (coef[min_pos] >= 0) level[min_pos] = (int16_t)(level[min_pos] + final_change);
                else level[min_pos] = (int16_t)(level[min_pos] - final_change);
            }
        }
        if (last_cg == 1) last_cg = 0;
    }
#undef POS
    int nnz = 0;
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) nnz += level[y * stride + x] != 0;
    return nnz;
}
void ora_dequant_block(const int16_t *src, int16_t *dst, int stride, int scale, int add, int shift, int w, int last_row)
{   /* H265DeQuantBlock_c E@0x439540 */
    for (int y = 0; y <= last_row; y++)
        for (int x = 0; x < w; x++) {
            int v = (src[y * stride + x] * scale + add) >> shift;
            dst[y * stride + x] = (int16_t)clip3(-32768, 32767, v);
        }
}
void ora_dequant(const int16_t *level, int16_t *coef, int stride, int qp, int log2n)
{   /* H265_GetBaseDeQuantParam E@0x439500: flat scaling list m=16 folded in: shift = log2N-1 (bitDepth 8) */
    int shift = log2n - 1, scale = ora_inv_quant_scales[qp % 6] << (qp / 6);
    ora_dequant_block(level, coef, stride, scale, 1 << (shift - 1), shift, 1 << log2n, (1 << log2n) - 1);
}
void ora_idct_add(const int16_t *coef, uint8_t *dst, const uint8_t *pred, int coef_stride, int dst_stride,
                  int pred_stride, int log2n, int is_dst)
{   /* H265_2dIDct{4,8,16,32}_c E@0x4417f0/446900/441ad0/447030, H265_2dIDst4x4_c E@0x441450:
     * spec 8.6.4.2 -- columns first (shift 7, clip int16), then rows (shift 12), + pred, clip u8 */
    int n = 1 << log2n;
    int16_t tmp[32 * 32];
    for (int x = 0; x < n; x++)
        for (int y = 0; y < n; y++) {
            int acc = 0;
            for (int k = 0; k < n; k++) acc += tcoef(log2n, is_dst, k, y) * coef[k * coef_stride + x];
            tmp[y * n + x] = (int16_t)clip3(-32768, 32767, (acc + 64) >> 7);
        }
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
            int acc = 0;
            for (int k = 0; k < n; k++) acc += tcoef(log2n, is_dst, k, x) * tmp[y * n + k];
            dst[y * dst_stride + x] = clip8(pred[y * pred_stride + x] + ((acc + 2048) >> 12));
        }
}

/* ------------------------------------------------------------------ intra prediction ------------- */
static const int8_t intra_angle[35] = {0,0,32,26,21,17,13,9,5,2,0,-2,-5,-9,-13,-17,-21,-26,-32,-26,-21,-17,-13,-9,-5,-2,0,2,5,9,13,17,21,26,32};
static const int16_t intra_inv_angle[35] = {0,0,0,0,0,0,0,0,0,0,0,-4096,-1638,-910,-630,-482,-390,-315,-256,-315,-390,-482,-630,-910,-1638,-4096,0,0,0,0,0,0,0,0,0};

void ora_intra_pred(uint8_t *dst, int ds, const uint8_t *nb, int log2n, int mode, int is_luma, int strong)
{   /* spec 8.4.4.2.3 (filtering), .4 planar, .5 DC, .6 angular.  (reference: IntraPred*_c / IntraPredFilterRef_*,
     * ComIntraPrediction.cpp; verified end-to-end through the reference decoder, not by leaf KAT) */
    int n = 1 << log2n;
    uint8_t fb[4 * 32 + 1];
    const uint8_t *p = nb;               /* p[2n] = corner, p[2n-1-y] = left[y], p[2n+1+x] = top[x] */
    if (is_luma && mode != 1 && n != 4) {
        int d1 = iabs(mode - 26), d2 = iabs(mode - 10), md = d1 < d2 ? d1 : d2;
        int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
        if (md > thr) {
            int c = nb[2 * n];
            if (strong && n == 32 && iabs(c + nb[4 * n] - 2 * nb[3 * n]) < 8 && iabs(c + nb[0] - 2 * nb[n]) < 8) {
                int bl = nb[0], tr = nb[4 * n];
                fb[0] = (uint8_t)bl; fb[2 * n] = (uint8_t)c; fb[4 * n] = (uint8_t)tr;
                for (int i = 0; i < 63; i++) {
                    fb[2 * n - 1 - i] = (uint8_t)(((63 - i) * c + (i + 1) * bl + 32) >> 6);
                    fb[2 * n + 1 + i] = (uint8_t)(((63 - i) * c + (i + 1) * tr + 32) >> 6);
                }
            } else {
                fb[0] = nb[0]; fb[4 * n] = nb[4 * n];
                for (int i = 1; i < 4 * n; i++) fb[i] = (uint8_t)((nb[i - 1] + 2 * nb[i] + nb[i + 1] + 2) >> 2);
            }
            p = fb;
        }
    }
    const uint8_t *left = p + 2 * n - 1;   /* left[-y] */
    const uint8_t *top = p + 2 * n + 1;    /* top[x]; top[-1] = corner */
#define L(y) left[-(y)]
    if (mode == 0) {
        for (int y = 0; y < n; y++) for (int x = 0; x < n; x++)
            dst[y * ds + x] = (uint8_t)(((n - 1 - x) * L(y) + (x + 1) * top[n] + (n - 1 - y) * top[x] + (y + 1) * L(n) + n) >> (log2n + 1));
    } else if (mode == 1) {
        int s = n;
        for (int i = 0; i < n; i++) s += top[i] + L(i);
        int dc = s >> (log2n + 1);
        for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) dst[y * ds + x] = (uint8_t)dc;
        if (is_luma && n < 32) {
            dst[0] = (uint8_t)((L(0) + 2 * dc + top[0] + 2) >> 2);
            for (int x = 1; x < n; x++) dst[x] = (uint8_t)((top[x] + 3 * dc + 2) >> 2);
            for (int y = 1; y < n; y++) dst[y * ds] = (uint8_t)((L(y) + 3 * dc + 2) >> 2);
        }
    } else {
        int ang = intra_angle[mode], inv = intra_inv_angle[mode];
        uint8_t refb[3 * 32 + 2]; uint8_t *ref = refb + 32;
        int vert = mode >= 18;
        /* ref[x] for x=-n..2n; main = top for vertical, left for horizontal */
        for (int x = 0; x <= n; x++) ref[x] = vert ? top[x - 1] : (x == 0 ? top[-1] : L(x - 1));
        if (ang < 0) {
            int lastx = (n * ang) >> 5;
            if (lastx < -1)
                for (int x = lastx; x <= -1; x++) { int i = -1 + ((x * inv + 128) >> 8); ref[x] = vert ? (i < 0 ? top[-1] : L(i)) : (i < 0 ? top[-1] : top[i]); }
        } else {
            for (int x = n + 1; x <= 2 * n; x++) ref[x] = vert ? top[x - 1] : L(x - 1);
        }
        for (int j = 0; j < n; j++) {       /* j = along the prediction direction axis (y for vertical) */
            int idx = ((j + 1) * ang) >> 5, f = ((j + 1) * ang) & 31;
            for (int i = 0; i < n; i++) {
                int v = f ? ((32 - f) * ref[i + idx + 1] + f * ref[i + idx + 2] + 16) >> 5 : ref[i + idx + 1];
                if (vert) dst[j * ds + i] = (uint8_t)v; else dst[i * ds + j] = (uint8_t)v;
            }
        }
        if (ang == 0 && is_luma && n < 32) {
            if (vert) for (int y = 0; y < n; y++) dst[y * ds] = clip8(top[0] + ((L(y) - top[-1]) >> 1));
            else      for (int x = 0; x < n; x++) dst[x] = clip8(L(0) + ((top[x] - top[-1]) >> 1));
        }
    }
#undef L
}

/* ------------------------------------------------------------------ a16 deblocking --------------- */
int ora_deblock_luma_seg(uint8_t *pix, int xs, int ys, int beta, int tc)
{   /* spec 8.7.2.5.3 decisions + 8.7.2.5.7 filter == EdgeFilterLuma{Ver,Hor}_c E@0x413100/0x4133f0 */
#define P(i, l) pix[-(i + 1) * xs + (l) * ys]
#define Q(i, l) pix[(i) * xs + (l) * ys]
    int dp0 = iabs(P(2,0) - 2 * P(1,0) + P(0,0)), dp3 = iabs(P(2,3) - 2 * P(1,3) + P(0,3));
    int dq0 = iabs(Q(2,0) - 2 * Q(1,0) + Q(0,0)), dq3 = iabs(Q(2,3) - 2 * Q(1,3) + Q(0,3));
    int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = dpq0 + dpq3;
    if (d >= beta) return 0;
    int s0 = 2 * dpq0 < (beta >> 2) && iabs(P(3,0) - P(0,0)) + iabs(Q(0,0) - Q(3,0)) < (beta >> 3) && iabs(P(0,0) - Q(0,0)) < ((5 * tc + 1) >> 1);
    int s3 = 2 * dpq3 < (beta >> 2) && iabs(P(3,3) - P(0,3)) + iabs(Q(0,3) - Q(3,3)) < (beta >> 3) && iabs(P(0,3) - Q(0,3)) < ((5 * tc + 1) >> 1);
    int strong = s0 && s3;
    int dep = dp < ((beta + (beta >> 1)) >> 3), deq = dq < ((beta + (beta >> 1)) >> 3);
    for (int l = 0; l < 4; l++) {
        int p0 = P(0,l), p1 = P(1,l), p2 = P(2,l), p3 = P(3,l), q0 = Q(0,l), q1 = Q(1,l), q2 = Q(2,l), q3 = Q(3,l);
        if (strong) {
            P(0,l) = (uint8_t)clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            P(1,l) = (uint8_t)clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
            P(2,l) = (uint8_t)clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
            Q(0,l) = (uint8_t)clip3(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            Q(1,l) = (uint8_t)clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
            Q(2,l) = (uint8_t)clip3(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
        } else {
            int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (iabs(delta) < tc * 10) {
                delta = clip3(-tc, tc, delta);
                P(0,l) = clip8(p0 + delta); Q(0,l) = clip8(q0 - delta);
                if (dep) P(1,l) = clip8(p1 + clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
                if (deq) Q(1,l) = clip8(q1 + clip3(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
            }
        }
    }
    return strong ? 2 : 1;
#undef P
#undef Q
}
void ora_deblock_chroma_seg(uint8_t *pix, int xs, int ys, int tc, int nlines)
{   /* spec 8.7.2.5.8 == PixelFilterChroma{Ver,Hor}_c E@0x4137d0/0x4138a0 */
    for (int l = 0; l < nlines; l++) {
        uint8_t *q = pix + l * ys;
        int p0 = q[-xs], p1 = q[-2 * xs], q0 = q[0], q1 = q[xs];
        int delta = clip3(-tc, tc, ((((q0 - p0) << 2) + p1 - q1 + 4) >> 3));
        q[-xs] = clip8(p0 + delta); q[0] = clip8(q0 - delta);
    }
}

/* ------------------------------------------------------------------ a17/a19 SAO ------------------ */
static inline int sgn(int v) { return (v > 0) - (v < 0); }
void ora_sao_stat_boeo01(int *eo, int *bo, const uint8_t *org, const uint8_t *rec, int rec_stride,
                         int org_stride, int w, int h, int row_step)
{   /* statSaoBoEo01_c E@0x4a6370: d=(int8)(org-rec); v=(d<<12)|1;
     * bo[rec>>3]+=v; eo[((2+sgn(c-up)+sgn(c-down))<<3)|(2+sgn(c-left)+sgn(c-right))]+=v */
    for (int y = 0; y < h; y += row_step) {
        const uint8_t *r = rec + y * rec_stride, *o = org + y * org_stride;
        for (int x = 0; x < w; x++) {
            int d = (int8_t)(uint8_t)(o[x] - r[x]);
            int v = (int)((uint32_t)d << 12) | 1;
            int c = r[x];
            int c0 = 2 + sgn(c - r[x - 1]) + sgn(c - r[x + 1]);
            int c1 = 2 + sgn(c - r[x - rec_stride]) + sgn(c - r[x + rec_stride]);
            bo[c >> 3] += v;
            eo[(c1 << 3) | c0] += v;
        }
    }
}
static const int8_t sao_dx[4][2] = {{-1, 1}, {0, 0}, {-1, 1}, {1, -1}};
static const int8_t sao_dy[4][2] = {{0, 0}, {-1, 1}, {-1, 1}, {-1, 1}};
static const uint8_t sao_cat[5] = {1, 2, 0, 3, 4};
void ora_sao_stats_ctb(ora_sao_stats *st, const uint8_t *org, int os, const uint8_t *rec, int rs,
                       int x0, int y0, int w, int h, int pic_w, int pic_h, int row_step, int n_classes)
{   /* row_step > 1 = the reference's `_fast` statistics (statSao*_fast_*: rowStep argument of statSaoBoEo01_c) */
    memset(st, 0, sizeof(*st));
    for (int y = y0; y < y0 + h; y += row_step)
        for (int x = x0; x < x0 + w; x++) {
            int c = rec[y * rs + x], d = (int)org[y * os + x] - c;
            st->bo_sum[c >> 3] += d; st->bo_cnt[c >> 3]++;
            for (int k = 0; k < n_classes; k++) {
                int xa = x + sao_dx[k][0], ya = y + sao_dy[k][0], xb = x + sao_dx[k][1], yb = y + sao_dy[k][1];
                if (xa < 0 || xb < 0 || ya < 0 || yb < 0 || xa >= pic_w || xb >= pic_w || ya >= pic_h || yb >= pic_h) continue;
                int cat = sao_cat[2 + sgn(c - rec[ya * rs + xa]) + sgn(c - rec[yb * rs + xb])];
                st->eo_sum[k][cat] += d; st->eo_cnt[k][cat]++;
            }
        }
}
void ora_sao_apply_ctb(uint8_t *dst, int ds, const uint8_t *src, int ss, int x0, int y0, int w, int h,
                       int pic_w, int pic_h, int type, int bp_or_class, const int8_t off[4])
{   /* spec 8.7.3 == qy265SaoApplyComponent E@0x43e770 / SaoApplyOffset{Bo,Eo0..3}_c */
    for (int y = y0; y < y0 + h; y++)
        for (int x = x0; x < x0 + w; x++) {
            int c = src[y * ss + x], v = c;
            if (type == 1) {
                int k = ((c >> 3) - bp_or_class) & 31;
                if (k < 4) v = c + off[k];
            } else if (type == 2) {
                int k = bp_or_class;
                int xa = x + sao_dx[k][0], ya = y + sao_dy[k][0], xb = x + sao_dx[k][1], yb = y + sao_dy[k][1];
                if (!(xa < 0 || xb < 0 || ya < 0 || yb < 0 || xa >= pic_w || xb >= pic_w || ya >= pic_h || yb >= pic_h)) {
                    int cat = sao_cat[2 + sgn(c - src[ya * ss + xa]) + sgn(c - src[yb * ss + xb])];
                    if (cat) v = c + off[cat - 1];
                }
            }
            dst[y * ds + x] = clip8(v);
        }
}
