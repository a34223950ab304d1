/*
 * ks_me.cuh -- motion estimation + interpolation for the ks265 B200 hot path.
 *
 * One warp owns one 16x16 cell.  The reference window (40 rows x 52 bytes) is staged in shared memory once
 * and re-centred only when the search walks out of it; every SAD is 2 x VABSDIFF4-accumulate per lane
 * (8 pixels per lane) and a packed butterfly shuffle reduction for 2 candidates at a time.
 * Sub-pel candidates use the real 8-tap interpolation (reference: subMeHpel_RealInterp E@0x4ac3b0 /
 * subMeQpel_8Sad_* E@0x4acb90.., kernels interpLuma{Hor,Ver}* E@0x417600..): horizontal taps run on
 * dp4a (u8 x s8), the 2-D case keeps the raw 14-bit row sums (no offset, like interpLumaHor8to16_c) in
 * shared memory and finishes with (sum+2048)>>12 like interpLumaVer16to8_c.
 * Search order / tie-breaking mirror oracle/ora_frame.c:me_cell bit for bit.
 */
#pragma once
#include "ks_common.cuh"

#define KS_WIN_H 40
#define KS_WIN_WW 13          /* words per window row (52 bytes; odd word pitch -> conflict-free rows) */
#define KS_WIN_MARGIN 12

struct KsWarpScratch {
    uint32_t win[KS_WIN_H][KS_WIN_WW];
    int16_t  tmp[24][16];     /* raw horizontal 8-tap sums, rows -3..+19 (single-block interpolation, ks_interp16) */
    uint32_t pl[2][12][16];   /* + two more planes shared by the sub-pel candidates of one cell (tmp doubles as plane 0); plane layout:
                                 [row pair][column] = rows (2k, 2k+1) packed lo/hi, so a vertical tap pair is ONE dp2a */
};

/* stage the window whose top-left luma sample is (wx0, wy0); wx0 % 4 == 0; coordinates clamp to the picture
 * (equivalent to the reference's padded reference planes, expandPicture_*) */
__device__ __forceinline__ void ks_load_window(uint32_t (*win)[KS_WIN_WW], const uint8_t *__restrict__ ref, int W, int H,
                                               int wx0, int wy0, int lane)
{
    /* lanes 0..12 / 13..25 own one word column of an even / odd row (26 of 32 lanes busy): the column, its bounds test and the
     * shared-memory address are loop invariants, a row costs a clamp, one multiply-add and the load.  All 20 loads of a lane are issued
     * before the first store (one L2 round trip per window); words that straddle the picture border are patched afterwards with
     * clamped byte loads (rare) */
    constexpr int PER = KS_WIN_H / 2;
    const int half = lane >= KS_WIN_WW ? 1 : 0, c = lane - half * KS_WIN_WW, gx = wx0 + 4 * c;
    const bool act = lane < 2 * KS_WIN_WW, in = act && gx >= 0 && gx + 3 < W;
    const uint8_t *col = ref + gx;
    uint32_t v[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int gy = min(max(wy0 + 2 * k + half, 0), H - 1);
        v[k] = in ? __ldg(reinterpret_cast<const uint32_t *>(col + (size_t)gy * W)) : 0u;
    }
    if (act) {
        uint32_t *dst = &win[half][c];
#pragma unroll
        for (int k = 0; k < PER; k++) dst[2 * k * KS_WIN_WW] = v[k];
    }
    if (__any_sync(0xffffffffu, act && !in)) {
        if (act && !in) {
#pragma unroll 1
            for (int k = 0; k < PER; k++) {
                const int gy = min(max(wy0 + 2 * k + half, 0), H - 1);
                const uint8_t *row = ref + (size_t)gy * W;
                uint32_t w = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) w |= (uint32_t)row[min(max(gx + b, 0), W - 1)] << (8 * b);
                win[2 * k + half][c] = w;
            }
        }
    }
    __syncwarp();
}

/* compact window for motion COMPENSATION of one 16x16 block with a known vector: 23 rows x 7 words (28 bytes) starting at
 * (wx0 % 4 == 0, wy0); the block's integer origin then sits at window (3..6, 3).  Same storage as the search window. */
#define KS_MCWIN_ROWS 23
#define KS_MCWIN_WORDS 7
__device__ __forceinline__ void ks_load_window_mc(uint32_t (*win)[KS_WIN_WW], const uint8_t *__restrict__ ref, int W, int H,
                                                  int wx0, int wy0, int lane)
{
    constexpr int NW = KS_MCWIN_ROWS * KS_MCWIN_WORDS, PER = (NW + KS_WARP - 1) / KS_WARP;
    uint32_t v[PER];
    unsigned border = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        int idx = lane + k * KS_WARP;
        int r = idx / KS_MCWIN_WORDS, c = idx - r * KS_MCWIN_WORDS;
        int gy = min(max(wy0 + r, 0), H - 1), gx = wx0 + 4 * c;
        bool in = idx < NW && gx >= 0 && gx + 3 < W;
        v[k] = in ? __ldg(reinterpret_cast<const uint32_t *>(ref + (size_t)gy * W + gx)) : 0u;
        if (idx < NW && !in) border |= 1u << k;
    }
#pragma unroll
    for (int k = 0; k < PER; k++) {
        int idx = lane + k * KS_WARP;
        int r = idx / KS_MCWIN_WORDS, c = idx - r * KS_MCWIN_WORDS;
        if (idx < NW) win[r][c] = v[k];
    }
    if (__any_sync(0xffffffffu, border != 0)) {
#pragma unroll 1
        for (int k = 0; k < PER; k++) {
            if (!((border >> k) & 1)) continue;
            int idx = lane + k * KS_WARP;
            int r = idx / KS_MCWIN_WORDS, c = idx - r * KS_MCWIN_WORDS;
            int gy = min(max(wy0 + r, 0), H - 1), gx = wx0 + 4 * c;
            const uint8_t *row = ref + (size_t)gy * W;
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) w |= (uint32_t)row[min(max(gx + b, 0), W - 1)] << (8 * b);
            win[r][c] = w;
        }
    }
    __syncwarp();
}

/* lane's 8 reference pixels of the 16x16 block whose origin is (bxw, byw) in window coordinates */
__device__ __forceinline__ void ks_win_px8(const uint32_t (*win)[KS_WIN_WW], int bxw, int byw, int lane, uint32_t &a, uint32_t &b)
{
    int wx = bxw + 8 * (lane & 1);
    const uint32_t *r = win[byw + (lane >> 1)] + (wx >> 2);
    unsigned sh = (wx & 3) * 8;
    a = __funnelshift_r(r[0], r[1], sh);
    b = __funnelshift_r(r[1], r[2], sh);
}
__device__ __forceinline__ unsigned ks_sad_partial(const uint32_t (*win)[KS_WIN_WW], int bxw, int byw, int lane, uint32_t s0, uint32_t s1)
{
    uint32_t a, b;
    ks_win_px8(win, bxw, byw, lane, a, b);
    return __vsadu4(a, s0) + __vsadu4(b, s1);
}

/* SATD of a 16x16 block (reference had_c E@0x474500 -> xCalcHADs8x8 E@0x474200): four 8x8 Hadamard tiles, each
 * (sum|coef| + 2) >> 2.  Lane holds 8 differences of row lane>>1, columns 8*(lane&1)..+7 (a = ref/pred, s = source).
 * Horizontal butterflies stay in registers, vertical ones are 3 shuffle stages across the 8 rows of a tile.
 * Warp-collective; every lane returns the block SATD.  Used as the sub-pel cost when the preset asks for SATD
 * (reference `satdInter`, qy265enc.h:138: fast..placebo), SAD otherwise (ultrafast..veryfast). */
__device__ __forceinline__ unsigned ks_satd16(uint32_t a0, uint32_t a1, uint32_t s0, uint32_t s1, int lane)
{
    int d[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        d[j] = (int)((s0 >> (8 * j)) & 255) - (int)((a0 >> (8 * j)) & 255);
        d[4 + j] = (int)((s1 >> (8 * j)) & 255) - (int)((a1 >> (8 * j)) & 255);
    }
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
        for (int i = 0; i < 8; i += 2 * len)
#pragma unroll
            for (int j = i; j < i + len; j++) { int u = d[j], v = d[j + len]; d[j] = u + v; d[j + len] = u - v; }
    /* rows of one 8x8 tile are lanes {2r + half}: row bit k of the tile <-> lane bit k+1 */
#pragma unroll
    for (int bit = 2; bit <= 8; bit <<= 1) {
        const bool hi = (lane & bit) != 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { int o = __shfl_xor_sync(0xffffffffu, d[j], bit); d[j] = hi ? o - d[j] : d[j] + o; }
    }
    unsigned t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) t += (unsigned)abs(d[j]);
    t += __shfl_xor_sync(0xffffffffu, t, 2); t += __shfl_xor_sync(0xffffffffu, t, 4); t += __shfl_xor_sync(0xffffffffu, t, 8);
    t = (t + 2) >> 2;                                   /* per 8x8 tile; tiles: lane bit 0 (left/right), lane bit 4 (top/bottom) */
    t += __shfl_xor_sync(0xffffffffu, t, 1); t += __shfl_xor_sync(0xffffffffu, t, 16);
    return t;
}

/* 16 bytes starting at window byte column p of row `row`, as 4 words */
__device__ __forceinline__ void ks_row16(const uint32_t (*win)[KS_WIN_WW], int row, int p, uint32_t n[4])
{
    const uint32_t *r = win[row] + (p >> 2);
    unsigned sh = (p & 3) * 8;
#pragma unroll
    for (int i = 0; i < 4; i++) n[i] = __funnelshift_r(r[i], r[i + 1], sh);
}
/* raw 8-tap horizontal sums for 8 consecutive outputs; n = 16 bytes starting 3 left of the first output */
__device__ __forceinline__ void ks_htaps8(const uint32_t n[4], int tlo, int thi, int out[8])
{
#pragma unroll
    for (int j = 0; j < 8; j++) {
        unsigned sh = (j & 3) * 8;
        uint32_t lo = __funnelshift_r(n[j >> 2], n[(j >> 2) + 1], sh);
        uint32_t hi = (j < 4) ? __funnelshift_r(n[1 + (j >> 2)], n[2 + (j >> 2)], sh)
                              : ((j == 4) ? n[2] : __funnelshift_r(n[2], n[3], sh));
        out[j] = ks_dp4a_us(hi, thi, ks_dp4a_us(lo, tlo, 0));
    }
}

/* quarter-sample prediction of the 16x16 block at integer window origin (bxw, byw) with fractions (fx, fy):
 * returns the lane's 8 predicted pixels (row lane>>1, columns 8*(lane&1)..+7) packed in (o0, o1).
 * Matches ora_mc_luma (spec 8.5.3.3.3.1) exactly. Warp-collective (uses scratch->tmp for the 2-D case). */
__device__ __forceinline__ void ks_interp16(KsWarpScratch *sc, int bxw, int byw, int fx, int fy, int lane, uint32_t &o0, uint32_t &o1)
{
    const int row = lane >> 1, half = lane & 1;
    if (fx == 0 && fy == 0) { ks_win_px8(sc->win, bxw, byw, lane, o0, o1); return; }
    int v[8];
    if (fy == 0) {
        uint32_t n[4];
        ks_row16(sc->win, byw + row, bxw + 8 * half - 3, n);
        ks_htaps8(n, c_luma_taps_packed[fx][0], c_luma_taps_packed[fx][1], v);
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (v[j] + 32) >> 6;
    } else if (fx == 0) {
        /* vertical taps straight from the samples, two 16-bit lanes per multiply: partial sums stay inside [-4080, 20400] */
        int e0 = 0, e1 = 0, e2 = 0, e3 = 0;
        const int wx = bxw + 8 * half;
        const unsigned sh = (wx & 3) * 8;
#pragma unroll 2
        for (int t = 0; t < 8; t++) {
            const int c = c_luma_taps[fy][t];
            const uint32_t *r = sc->win[byw + row - 3 + t] + (wx >> 2);
            const uint32_t a = __funnelshift_r(r[0], r[1], sh), b = __funnelshift_r(r[1], r[2], sh);
            e0 += c * (int)__byte_perm(a, 0, 0x4240); e1 += c * (int)__byte_perm(a, 0, 0x4341);
            e2 += c * (int)__byte_perm(b, 0, 0x4240); e3 += c * (int)__byte_perm(b, 0, 0x4341);
        }
        v[0] = (int)(short)(e0 & 0xffff); v[2] = (e0 + 0x8000) >> 16; v[1] = (int)(short)(e1 & 0xffff); v[3] = (e1 + 0x8000) >> 16;
        v[4] = (int)(short)(e2 & 0xffff); v[6] = (e2 + 0x8000) >> 16; v[5] = (int)(short)(e3 & 0xffff); v[7] = (e3 + 0x8000) >> 16;
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (v[j] + 32) >> 6;
    } else {
        const int tlo = c_luma_taps_packed[fx][0], thi = c_luma_taps_packed[fx][1];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            int rr = row + 16 * k;
            if (rr < 23) {
                uint32_t n[4]; int h[8];
                ks_row16(sc->win, byw - 3 + rr, bxw + 8 * half - 3, n);
                ks_htaps8(n, tlo, thi, h);
                uint32_t *d = reinterpret_cast<uint32_t *>(&sc->tmp[rr][8 * half]);
#pragma unroll
                for (int j = 0; j < 4; j++) d[j] = ((uint32_t)h[2 * j] & 0xffffu) | ((uint32_t)h[2 * j + 1] << 16);
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 0;
#pragma unroll 2
        for (int t = 0; t < 8; t++) {
            int c = c_luma_taps[fy][t];
            uint4 q = *reinterpret_cast<const uint4 *>(&sc->tmp[row + t][8 * half]);
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) { v[2 * j] += c * (int)(short)(w[j] & 0xffffu); v[2 * j + 1] += c * ((int)w[j] >> 16); }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (v[j] + 2048) >> 12;
        __syncwarp();
    }
    o0 = ks_pack_sat4(v[0], v[1], v[2], v[3]);
    o1 = ks_pack_sat4(v[4], v[5], v[6], v[7]);
}

/* ---- shared intermediates for the sub-pel search (reference: subMeQpel_8Sad_* pick a variant "so H/V intermediate
 * rows are shared").  A plane holds, for 24 window rows x 16 columns starting at (wxb, wyb), the raw
 * horizontal 8-tap sums for fraction fx (14-bit, like interpLumaHor8to16_c) or sample<<6 for fx == 0
 * (InterpolateCopy8to16_c), rows interleaved in pairs: word [r>>1][c] = (row r even | row r odd << 16).
 * Every candidate of the stage is then one vertical pass over a plane: 5 dp2a (s16 x s8 pairs) per sample. ---- */
typedef uint32_t (*KsPlane)[16];
/* physical 16-byte segment of logical segment `seg` (4 columns) in pair-row `pr`: every other PAIR of pair-rows has its segments swapped
 * pairwise, so the quarter-warp that spans pair-rows p, p+1, p+2 (odd first row) reads six distinct bank groups instead of colliding on p / p+2 */
#define KS_PLANE_SEG(pr, seg) ((seg) ^ (((pr) >> 1) & 1))
__device__ __forceinline__ void ks_make_plane(KsPlane pl, const uint32_t (*win)[KS_WIN_WW], int wxb, int wyb, int fx, int lane)
{
    const int tlo = c_luma_taps_packed[fx][0], thi = c_luma_taps_packed[fx][1];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int u = lane + 32 * k;               /* 48 units: (row, 8-column group); the lane two over holds the other row of the pair */
        const int r = u >> 1, g = u & 1;
        int h[8];
        if (u < 48) {
            uint32_t n[4];
            if (fx) {
                ks_row16(win, wyb + r, wxb + 8 * g - 3, n);
                ks_htaps8(n, tlo, thi, h);
            } else {
                const int wx = wxb + 8 * g;
                const uint32_t *rr = win[wyb + r] + (wx >> 2);
                const unsigned sh = (wx & 3) * 8;
                const uint32_t a = __funnelshift_r(rr[0], rr[1], sh), b = __funnelshift_r(rr[1], rr[2], sh);
#pragma unroll
                for (int j = 0; j < 4; j++) { h[j] = (int)((a >> (8 * j)) & 255) << 6; h[4 + j] = (int)((b >> (8 * j)) & 255) << 6; }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) h[j] = 0;
        }
        /* pair interleave in registers: the even row keeps columns 0..3 and receives the odd row's, the odd row keeps 4..7 and receives the
         * even row's; each lane then stores four complete (even | odd << 16) words with one 128-bit store (was eight 16-bit stores) */
        const int odd = r & 1;
        const uint32_t s0 = ((uint32_t)h[odd ? 0 : 4] & 0xffffu) | ((uint32_t)h[odd ? 1 : 5] << 16);
        const uint32_t s1 = ((uint32_t)h[odd ? 2 : 6] & 0xffffu) | ((uint32_t)h[odd ? 3 : 7] << 16);
        const uint32_t g0 = __shfl_xor_sync(0xffffffffu, s0, 2), g1 = __shfl_xor_sync(0xffffffffu, s1, 2);
        if (u < 48) {
            uint4 w;
            if (!odd) {
                w.x = ((uint32_t)h[0] & 0xffffu) | (g0 << 16);          w.y = ((uint32_t)h[1] & 0xffffu) | (g0 & 0xffff0000u);
                w.z = ((uint32_t)h[2] & 0xffffu) | (g1 << 16);          w.w = ((uint32_t)h[3] & 0xffffu) | (g1 & 0xffff0000u);
            } else {
                w.x = (g0 & 0xffffu) | ((uint32_t)h[4] << 16);          w.y = (g0 >> 16) | ((uint32_t)h[5] << 16);
                w.z = (g1 & 0xffffu) | ((uint32_t)h[6] << 16);          w.w = (g1 >> 16) | ((uint32_t)h[7] << 16);
            }
            const int pr = r >> 1;
            *reinterpret_cast<uint4 *>(&pl[pr][KS_PLANE_SEG(pr, 2 * g + odd) * 4]) = w;
        }
    }
    __syncwarp();
}
__device__ __forceinline__ int ks_dp2a_lo(int a, int b, int c) { int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int ks_dp2a_hi(int a, int b, int c) { int d; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
/* lane's 8 predicted samples (row lane>>1, columns 8*(lane&1)..+7) from a plane: vertical 8-tap with fraction fy over
 * plane rows roff+row .. roff+row+7 ((sum+2048)>>12, == interpLumaVer16to8_c), or the centre row rounded ((v+32)>>6) for fy == 0 */
__device__ __forceinline__ void ks_plane_pred(const KsPlane pl, int roff, int fy, int lane, uint32_t &o0, uint32_t &o1)
{
    const int row = lane >> 1, half = lane & 1;
    int v[8];
    if (fy == 0) {
        const int R = roff + 3 + row;
        const int prc = R >> 1;
        const uint4 qa = *reinterpret_cast<const uint4 *>(&pl[prc][KS_PLANE_SEG(prc, 2 * half) * 4]), qb = *reinterpret_cast<const uint4 *>(&pl[prc][KS_PLANE_SEG(prc, 2 * half + 1) * 4]);
        const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int j = 0; j < 8; j++) { int t = (R & 1) ? ((int)w[j] >> 16) : (int)(short)(w[j] & 0xffffu); v[j] = (t + 32) >> 6; }
    } else {
        const int R0 = roff + row, par = R0 & 1;
        const int k0 = c_vtaps_pk[fy][par][0], k1 = c_vtaps_pk[fy][par][1], k2 = c_vtaps_pk[fy][par][2];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 2048;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const int pr = min((R0 >> 1) + i, 11);      /* the 5th pair only matters for odd R0; clamp keeps the even case in bounds (its taps are 0) */
            const uint4 qa = *reinterpret_cast<const uint4 *>(&pl[pr][KS_PLANE_SEG(pr, 2 * half) * 4]), qb = *reinterpret_cast<const uint4 *>(&pl[pr][KS_PLANE_SEG(pr, 2 * half + 1) * 4]);
            const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const int kk = i < 2 ? k0 : (i < 4 ? k1 : k2);
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (i & 1) ? ks_dp2a_hi((int)w[j], kk, v[j]) : ks_dp2a_lo((int)w[j], kk, v[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] >>= 12;
    }
    o0 = ks_pack_sat4(v[0], v[1], v[2], v[3]);      /* clip to 8 bits and pack in one step */
    o1 = ks_pack_sat4(v[4], v[5], v[6], v[7]);
}

/* window origin that centres integer offset (cx, cy) of the cell at (x0, y0) */
__device__ __forceinline__ void ks_center_window(int x0, int y0, int cx, int cy, int &wx0, int &wy0)
{
    wx0 = (x0 + cx - KS_WIN_MARGIN) & ~3;
    wy0 = y0 + cy - KS_WIN_MARGIN;
}

/* ------------------------------------------------------------------ chroma motion compensation --- */
/* 8x8 chroma block, 4-tap filters, eighth-sample mv (spec 8.5.3.3.3.2 == ora_mc_chroma).  Warp-collective.
 * cwin: 12x12 byte window (rows/cols -1..+10 of the integer position), tmp: >= 11x8 int16. */
__device__ __forceinline__ void ks_mc_chroma8(uint8_t *cwin, int16_t *tmp, const uint8_t *__restrict__ ref, int PW, int PH,
                                              int xc, int yc, int mvx, int mvy, uint8_t *dst, int dpitch, int lane)
{
    const int ix = xc + (mvx >> 3) - 1, iy = yc + (mvy >> 3) - 1, fx = mvx & 7, fy = mvy & 7;
    {
        uint8_t b[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {                          /* 5 independent loads in flight per lane */
            int idx = min(lane + k * KS_WARP, 143), r = idx / 12, c = idx - r * 12;
            b[k] = __ldg(ref + (size_t)min(max(iy + r, 0), PH - 1) * PW + min(max(ix + c, 0), PW - 1));
        }
#pragma unroll
        for (int k = 0; k < 5; k++) if (lane + k * KS_WARP < 144) cwin[lane + k * KS_WARP] = b[k];
    }
    __syncwarp();
    const int row = lane >> 2, col = (lane & 3) * 2;
    int v0, v1;
    if (fx == 0 && fy == 0) { v0 = cwin[(row + 1) * 12 + col + 1]; v1 = cwin[(row + 1) * 12 + col + 2]; }
    else if (fy == 0) {
        const uint8_t *p = cwin + (row + 1) * 12 + col;
        int a = 0, b = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) { int c = c_chroma_taps[fx][t]; a += c * p[t]; b += c * p[t + 1]; }
        v0 = ks_clip8((a + 32) >> 6); v1 = ks_clip8((b + 32) >> 6);
    } else if (fx == 0) {
        const uint8_t *p = cwin + row * 12 + col + 1;
        int a = 0, b = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) { int c = c_chroma_taps[fy][t]; a += c * p[t * 12]; b += c * p[t * 12 + 1]; }
        v0 = ks_clip8((a + 32) >> 6); v1 = ks_clip8((b + 32) >> 6);
    } else {
        for (int idx = lane; idx < 88; idx += KS_WARP) {      /* raw horizontal sums, rows -1..+9 */
            int r = idx >> 3, c = idx & 7;
            const uint8_t *p = cwin + r * 12 + c;
            int a = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) a += c_chroma_taps[fx][t] * p[t];
            tmp[idx] = (int16_t)a;
        }
        __syncwarp();
        int a = 0, b = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) { int c = c_chroma_taps[fy][t]; a += c * tmp[(row + t) * 8 + col]; b += c * tmp[(row + t) * 8 + col + 1]; }
        v0 = ks_clip8((a + 2048) >> 12); v1 = ks_clip8((b + 2048) >> 12);
    }
    dst[row * dpitch + col] = (uint8_t)v0; dst[row * dpitch + col + 1] = (uint8_t)v1;
    __syncwarp();
}

#define KS_ME_WARPS 8
/* METHOD (0 diamond / 1 hexagon) and SATD are compile-time: each instantiation carries only the code it runs, which keeps the
 * hot configuration (diamond + SAD, veryfast) inside the instruction cache */
template <int METHOD, bool SATD>
__global__ void __launch_bounds__(KS_ME_WARPS * KS_WARP, 4)
ks_me_kernel(KsPicParams pp, const uint8_t *__restrict__ srcY, KsPlanes ref, const ks_cell *__restrict__ prev_cells,
             ks_cell *__restrict__ cells, KsPlanes pred, int *__restrict__ costs, int *__restrict__ dists, unsigned long long *__restrict__ cost_sum)
{
    __shared__ __align__(16) KsWarpScratch scratch[KS_ME_WARPS];
    __shared__ uint8_t cwins[KS_ME_WARPS][144];
    const uint8_t *__restrict__ refY = ref.p[0];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * KS_ME_WARPS + warp;
    if (cell >= pp.cw * pp.ch) return;
    KsWarpScratch *sc = &scratch[warp];
    const int cyc = cell / pp.cw, cxc = cell - cyc * pp.cw, x0 = cxc << 4, y0 = cyc << 4;
    const int W = pp.W, H = pp.H, R = pp.me_range, lam = pp.lambda_sad_q4;
    /* source pixels of this lane: row lane>>1, 8 bytes */
    const uint2 s = *reinterpret_cast<const uint2 *>(srcY + (size_t)(y0 + (lane >> 1)) * W + x0 + 8 * (lane & 1));
    int tpx = 0, tpy = 0;
    if (prev_cells) { ks_cell pc = prev_cells[cell]; if (!(pc.flags & KS_F_INTRA)) { tpx = pc.mvx; tpy = pc.mvy; } }
    if (pp.pred_den) { tpx = (tpx * pp.pred_num) / pp.pred_den; tpy = (tpy * pp.pred_num) / pp.pred_den; }    /* B pictures: anchor vector scaled to this list */
#define MVCOST(qx, qy) ((lam * (ks_mvbits((qx) - tpx) + ks_mvbits((qy) - tpy))) >> 4)

    /* ---- start point: zero vs rounded temporal predictor (reference: meInitPoint E@0x4818b0) ---- */
    int wx0, wy0;
    ks_center_window(x0, y0, 0, 0, wx0, wy0);
    ks_load_window(sc->win, refY, W, H, wx0, wy0, lane);
    int bx = 0, by = 0;
    int bc = (int)ks_warp_sum(ks_sad_partial(sc->win, x0 - wx0, y0 - wy0, lane, s.x, s.y)) + MVCOST(0, 0);
    {
        int cx = ks_clip3(-R, R, (tpx + 2) >> 2), cy = ks_clip3(-R, R, (tpy + 2) >> 2);
        if (cx | cy) {
            int bxw = x0 + cx - wx0, byw = y0 + cy - wy0;
            if (bxw < 0 || bxw > 35 || byw < 0 || byw > 24) {
                ks_center_window(x0, y0, cx, cy, wx0, wy0);
                ks_load_window(sc->win, refY, W, H, wx0, wy0, lane);
                bxw = x0 + cx - wx0; byw = y0 + cy - wy0;
            }
            int c = (int)ks_warp_sum(ks_sad_partial(sc->win, bxw, byw, lane, s.x, s.y)) + MVCOST(cx * 4, cy * 4);
            if (c < bc) { bc = c; bx = cx; by = cy; }
        }
    }
    if (METHOD == 0) {
    /* ---- small diamond (reference: interMeDia E@0x4849d0, x264 DIA with sad4 order up,down,left,right) ---- */
    for (int it = 0; it < pp.me_iters; it++) {
        int bxw = x0 + bx - wx0, byw = y0 + by - wy0;
        if (bxw < 1 || bxw > 34 || byw < 1 || byw > 23) {
            ks_center_window(x0, y0, bx, by, wx0, wy0);
            ks_load_window(sc->win, refY, W, H, wx0, wy0, lane);
            bxw = x0 + bx - wx0; byw = y0 + by - wy0;
        }
        unsigned p01 = ks_sad_partial(sc->win, bxw, byw - 1, lane, s.x, s.y) | (ks_sad_partial(sc->win, bxw, byw + 1, lane, s.x, s.y) << 16);
        unsigned p23 = ks_sad_partial(sc->win, bxw - 1, byw, lane, s.x, s.y) | (ks_sad_partial(sc->win, bxw + 1, byw, lane, s.x, s.y) << 16);
        p01 = ks_warp_sum(p01); p23 = ks_warp_sum(p23);
        int sad[4] = {(int)(p01 & 0xffffu), (int)(p01 >> 16), (int)(p23 & 0xffffu), (int)(p23 >> 16)};
        const int dx[4] = {0, 0, -1, 1}, dy[4] = {-1, 1, 0, 0};
        int bk = -1, lc = bc;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int nx = bx + dx[k], ny = by + dy[k];
            if (abs(nx) > R || abs(ny) > R) continue;
            int c = sad[k] + MVCOST(nx * 4, ny * 4);
            if (c < lc) { lc = c; bk = k; }
        }
        if (bk < 0) break;
        bx += dx[bk]; by += dy[bk]; bc = lc;
    }
    } else {
    /* ---- hexagon + square refine (reference: interMeHex E@0x484c00 = x264 HEX: six points once, then the three new points of
     *      the hexagon moved in the winning direction (mod6m1 rotation), then the eight square neighbours) ---- */
#define KS_RECENTER(m, xhi, yhi) { int bxw_ = x0 + bx - wx0, byw_ = y0 + by - wy0; \
        if (bxw_ < (m) || bxw_ > (xhi) || byw_ < (m) || byw_ > (yhi)) { ks_center_window(x0, y0, bx, by, wx0, wy0); ks_load_window(sc->win, refY, W, H, wx0, wy0, lane); } }
#define KS_SAD2(ax, ay, cx_, cy_) ks_warp_sum(ks_sad_partial(sc->win, x0 + bx - wx0 + (ax), y0 + by - wy0 + (ay), lane, s.x, s.y) | \
                                              (ks_sad_partial(sc->win, x0 + bx - wx0 + (cx_), y0 + by - wy0 + (cy_), lane, s.x, s.y) << 16))
#define KS_TRY(sadv, ox, oy, tag) { int nx = bx + (ox), ny = by + (oy); \
        if (abs(nx) <= R && abs(ny) <= R) { int c = (int)(sadv) + MVCOST(nx * 4, ny * 4); if (c < lc) { lc = c; bk = (tag); } } }
        int dir, bk = -1, lc = bc;
        KS_RECENTER(2, 33, 22)
        {
            const unsigned a = KS_SAD2(-2, 0, -1, 2), b = KS_SAD2(1, 2, 2, 0), c2 = KS_SAD2(1, -2, -1, -2);
            KS_TRY(a & 0xffffu, -2, 0, 0) KS_TRY(a >> 16, -1, 2, 1) KS_TRY(b & 0xffffu, 1, 2, 2)
            KS_TRY(b >> 16, 2, 0, 3) KS_TRY(c2 & 0xffffu, 1, -2, 4) KS_TRY(c2 >> 16, -1, -2, 5)
        }
        if (bk >= 0) {
            dir = bk; bx += c_hex_dx[dir + 1]; by += c_hex_dy[dir + 1]; bc = lc;
            for (int it = 1; it < pp.me_iters; it++) {
                KS_RECENTER(2, 33, 22)
                const int ax = c_hex_dx[dir], ay = c_hex_dy[dir], mx_ = c_hex_dx[dir + 1], my_ = c_hex_dy[dir + 1], ex = c_hex_dx[dir + 2], ey = c_hex_dy[dir + 2];
                const unsigned a = KS_SAD2(ax, ay, mx_, my_);
                const unsigned b = ks_warp_sum(ks_sad_partial(sc->win, x0 + bx - wx0 + ex, y0 + by - wy0 + ey, lane, s.x, s.y));
                bk = -1; lc = bc;
                KS_TRY(a & 0xffffu, ax, ay, 0) KS_TRY(a >> 16, mx_, my_, 1) KS_TRY(b & 0xffffu, ex, ey, 2)
                if (bk < 0) break;
                dir += bk - 1; dir = dir < 0 ? 5 : (dir > 5 ? 0 : dir);
                bx += c_hex_dx[dir + 1]; by += c_hex_dy[dir + 1]; bc = lc;
            }
        }
        KS_RECENTER(1, 34, 23)
        {
            const unsigned a = KS_SAD2(0, -1, 0, 1), b = KS_SAD2(-1, 0, 1, 0), c2 = KS_SAD2(-1, -1, -1, 1), d = KS_SAD2(1, -1, 1, 1);
            bk = -1; lc = bc;
            KS_TRY(a & 0xffffu, 0, -1, 0) KS_TRY(a >> 16, 0, 1, 1) KS_TRY(b & 0xffffu, -1, 0, 2) KS_TRY(b >> 16, 1, 0, 3)
            KS_TRY(c2 & 0xffffu, -1, -1, 4) KS_TRY(c2 >> 16, -1, 1, 5) KS_TRY(d & 0xffffu, 1, -1, 6) KS_TRY(d >> 16, 1, 1, 7)
            if (bk >= 0) { bx += c_sq_dx[bk]; by += c_sq_dy[bk]; bc = lc; }
        }
#undef KS_RECENTER
#undef KS_SAD2
#undef KS_TRY
    }
    /* ---- half then quarter refinement, 8 neighbours each (reference: subMeSquare E@0x4aee80).  The candidates of a
     *      stage share planes of raw horizontal sums: 2 planes + 2 vertical-only blocks for the half stage, 3 planes (one
     *      per x offset) for the quarter stage; each candidate is then a single vertical pass + SAD. ---- */
    int mx = bx * 4, my = by * 4;
    uint32_t best0, best1;                      /* the winning candidate's prediction (this lane's 8 samples) */
    ks_win_px8(sc->win, x0 + bx - wx0, y0 + by - wy0, lane, best0, best1);
    if (pp.subpel > 0) {
        int bxw = x0 + bx - wx0, byw = y0 + by - wy0;
        if (bxw < 4 || bxw > 27 || byw < 4 || byw > 19) {
            ks_center_window(x0, y0, bx, by, wx0, wy0);
            ks_load_window(sc->win, refY, W, H, wx0, wy0, lane);
            bxw = x0 + bx - wx0; byw = y0 + by - wy0;
        }
        const int sqx[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, sqy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
        const bool satd = SATD;
#define KS_SUBCOST(o0, o1) (satd ? ks_satd16(o0, o1, s.x, s.y, lane) : ks_warp_sum(__vsadu4(o0, s.x) + __vsadu4(o1, s.y)))
        if (satd) {     /* the metric changes for the sub-pel stages: re-cost the integer winner with SATD (x264 does the same) */
            uint32_t c0, c1;
            ks_win_px8(sc->win, bxw, byw, lane, c0, c1);
            bc = (int)ks_satd16(c0, c1, s.x, s.y, lane) + MVCOST(mx, my);
        }
        KsPlane P0 = reinterpret_cast<KsPlane>(sc->tmp), P1 = sc->pl[0], P2 = sc->pl[1];
        {   /* half-sample stage: planes for x-1/2 (P0) and x+1/2 (P2), rows -4..+19 of the integer position */
            const int wyb = byw - 4;
            ks_make_plane(P0, sc->win, bxw - 1, wyb, 2, lane);
            ks_make_plane(P2, sc->win, bxw, wyb, 2, lane);
            int bk = -1, lc = bc;
            /* the 8 candidates share 3 x and 3 y vector components: their bit costs are computed once */
            const int hbx0 = ks_mvbits(mx - 2 - tpx), hbx1 = ks_mvbits(mx - tpx), hbx2 = ks_mvbits(mx + 2 - tpx);
            const int hby0 = ks_mvbits(my - 2 - tpy), hby1 = ks_mvbits(my - tpy), hby2 = ks_mvbits(my + 2 - tpy);
#pragma unroll 1
            for (int k = 0; k < 8; k++) {
                const int dx = sqx[k], dy = sqy[k];
                uint32_t o0, o1;
                if (dx == 0) ks_interp16(sc, bxw, byw + (dy < 0 ? -1 : 0), 0, 2, lane, o0, o1);      /* vertical-only, straight from the samples */
                else ks_plane_pred(dx < 0 ? P0 : P2, dy < 0 ? 0 : 1, dy ? 2 : 0, lane, o0, o1);
                int c = (int)KS_SUBCOST(o0, o1) + ((lam * ((dx < 0 ? hbx0 : (dx == 0 ? hbx1 : hbx2)) + (dy < 0 ? hby0 : (dy == 0 ? hby1 : hby2)))) >> 4);
                if (c < lc) { lc = c; bk = k; best0 = o0; best1 = o1; }
            }
            if (bk >= 0) { mx += sqx[bk] * 2; my += sqy[bk] * 2; bc = lc; }
        }
        if (pp.subpel > 1) {   /* quarter-sample stage around (mx, my): one plane per x offset */
            const int iym = (my - 1) >> 2, wyb = y0 + iym - wy0 - 3;
            ks_make_plane(P0, sc->win, x0 + ((mx - 1) >> 2) - wx0, wyb, (mx - 1) & 3, lane);
            ks_make_plane(P1, sc->win, x0 + (mx >> 2) - wx0, wyb, mx & 3, lane);
            ks_make_plane(P2, sc->win, x0 + ((mx + 1) >> 2) - wx0, wyb, (mx + 1) & 3, lane);
            int bk = -1, lc = bc;
            const int qbx0 = ks_mvbits(mx - 1 - tpx), qbx1 = ks_mvbits(mx - tpx), qbx2 = ks_mvbits(mx + 1 - tpx);
            const int qby0 = ks_mvbits(my - 1 - tpy), qby1 = ks_mvbits(my - tpy), qby2 = ks_mvbits(my + 1 - tpy);
#pragma unroll 1
            for (int k = 0; k < 8; k++) {
                const int dx = sqx[k], dy = sqy[k], qy = my + dy;
                uint32_t o0, o1;
                ks_plane_pred(dx < 0 ? P0 : (dx == 0 ? P1 : P2), (qy >> 2) - iym, qy & 3, lane, o0, o1);
                int c = (int)KS_SUBCOST(o0, o1) + ((lam * ((dx < 0 ? qbx0 : (dx == 0 ? qbx1 : qbx2)) + (dy < 0 ? qby0 : (dy == 0 ? qby1 : qby2)))) >> 4);
                if (c < lc) { lc = c; bk = k; best0 = o0; best1 = o1; }
            }
            if (bk >= 0) { mx += sqx[bk]; my += sqy[bk]; bc = lc; }
        }
#undef KS_SUBCOST
    }
    const int dist = bc - MVCOST(mx, my);
#undef MVCOST
    if (lane == 0) {
        ks_cell c; c.mvx = (int16_t)mx; c.mvy = (int16_t)my; c.cu_log2 = 4; c.flags = 0; c.intra_mode = 0; c.rsv = 0;
        cells[cell] = c;
        if (costs) costs[cell] = bc;
        if (dists) dists[cell] = dist;
        if (cost_sum) atomicAdd(cost_sum, (unsigned long long)bc);      /* integer sum: order-independent */
    }
    /* ---- the search already holds the winner's luma prediction: emit it (and the chroma prediction for the same vector) so
     *      the residual kernel needs no motion compensation pass of its own (reference: getReusSubMePred E@0x486770 re-uses the
     *      sub-ME interpolation as the final prediction) ---- */
    if (pred.p[0]) {
        *reinterpret_cast<uint2 *>(pred.p[0] + (size_t)(y0 + (lane >> 1)) * W + x0 + 8 * (lane & 1)) = make_uint2(best0, best1);
        const int CW = W >> 1, CH = H >> 1;
#pragma unroll 1
        for (int ci = 0; ci < 2; ci++)
            ks_mc_chroma8(cwins[warp], &sc->tmp[0][0], ref.p[1 + ci], CW, CH, x0 >> 1, y0 >> 1, mx, my,
                          pred.p[1 + ci] + (size_t)(y0 >> 1) * CW + (x0 >> 1), CW, lane);
    }
}

void ks_launch_me(const KsPicParams &pp, const uint8_t *srcY, KsPlanes ref, const ks_cell *prev_cells, ks_cell *cells, KsPlanes pred, int *costs, int *dists, unsigned long long *cost_sum, cudaStream_t st)
{
    const int ncell = pp.cw * pp.ch;
    const dim3 grid((ncell + KS_ME_WARPS - 1) / KS_ME_WARPS), block(KS_ME_WARPS * KS_WARP);
    if (pp.me_method == 0) {
        if (pp.satd) ks_me_kernel<0, true><<<grid, block, 0, st>>>(pp, srcY, ref, prev_cells, cells, pred, costs, dists, cost_sum);
        else ks_me_kernel<0, false><<<grid, block, 0, st>>>(pp, srcY, ref, prev_cells, cells, pred, costs, dists, cost_sum);
    } else {
        if (pp.satd) ks_me_kernel<1, true><<<grid, block, 0, st>>>(pp, srcY, ref, prev_cells, cells, pred, costs, dists, cost_sum);
        else ks_me_kernel<1, false><<<grid, block, 0, st>>>(pp, srcY, ref, prev_cells, cells, pred, costs, dists, cost_sum);
    }
}

/* ------------------------------------------------------------------ bi-prediction (B pictures) ---- */
/* 14-bit intermediate prediction of the 16x16 block (reference interpolatePuBi E@0x488020: InterpolateCopy8to16 / interpLumaHor8to16 /
 * interpLumaVer8to16 / Hor8to16+Ver16to16): lane's 8 samples.  == ora_mc_luma_16 */
__device__ __forceinline__ void ks_interp16_raw(KsWarpScratch *sc, int bxw, int byw, int fx, int fy, int lane, int v[8])
{
    const int row = lane >> 1, half = lane & 1;
    if (fx == 0 && fy == 0) {
        uint32_t a, b; ks_win_px8(sc->win, bxw, byw, lane, a, b);
#pragma unroll
        for (int j = 0; j < 4; j++) { v[j] = (int)((a >> (8 * j)) & 255) << 6; v[4 + j] = (int)((b >> (8 * j)) & 255) << 6; }
    } else if (fy == 0) {
        uint32_t n[4];
        ks_row16(sc->win, byw + row, bxw + 8 * half - 3, n);
        ks_htaps8(n, c_luma_taps_packed[fx][0], c_luma_taps_packed[fx][1], v);
    } else if (fx == 0) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 0;
#pragma unroll 2
        for (int t = 0; t < 8; t++) {
            const int c = c_luma_taps[fy][t], wx = bxw + 8 * half;
            const uint32_t *r = sc->win[byw + row - 3 + t] + (wx >> 2);
            const unsigned sh = (wx & 3) * 8;
            const uint32_t a = __funnelshift_r(r[0], r[1], sh), b = __funnelshift_r(r[1], r[2], sh);
#pragma unroll
            for (int j = 0; j < 4; j++) { v[j] += c * (int)((a >> (8 * j)) & 255); v[4 + j] += c * (int)((b >> (8 * j)) & 255); }
        }
    } else {
        const int tlo = c_luma_taps_packed[fx][0], thi = c_luma_taps_packed[fx][1];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            int rr = row + 16 * k;
            if (rr < 23) {
                uint32_t n[4]; int h[8];
                ks_row16(sc->win, byw - 3 + rr, bxw + 8 * half - 3, n);
                ks_htaps8(n, tlo, thi, h);
                uint32_t *d = reinterpret_cast<uint32_t *>(&sc->tmp[rr][8 * half]);
#pragma unroll
                for (int j = 0; j < 4; j++) d[j] = ((uint32_t)h[2 * j] & 0xffffu) | ((uint32_t)h[2 * j + 1] << 16);
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 0;
#pragma unroll 2
        for (int t = 0; t < 8; t++) {
            const int c = c_luma_taps[fy][t];
            const uint4 q = *reinterpret_cast<const uint4 *>(&sc->tmp[row + t][8 * half]);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) { v[2 * j] += c * (int)(short)(w[j] & 0xffffu); v[2 * j + 1] += c * ((int)w[j] >> 16); }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] >>= 6;
        __syncwarp();
    }
}
/* 14-bit chroma prediction (2 samples per lane: row lane>>2, columns 2*(lane&3)..+1) of the 8x8 block; == ora_mc_chroma_16 */
__device__ __forceinline__ void ks_mc_chroma8_raw(uint8_t *cwin, int16_t *tmp, const uint8_t *__restrict__ ref, int PW, int PH,
                                                  int xc, int yc, int mvx, int mvy, int lane, int &v0, int &v1)
{
    const int ix = xc + (mvx >> 3) - 1, iy = yc + (mvy >> 3) - 1, fx = mvx & 7, fy = mvy & 7;
    {
        uint8_t b[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            int idx = min(lane + k * KS_WARP, 143), r = idx / 12, c = idx - r * 12;
            b[k] = __ldg(ref + (size_t)min(max(iy + r, 0), PH - 1) * PW + min(max(ix + c, 0), PW - 1));
        }
#pragma unroll
        for (int k = 0; k < 5; k++) if (lane + k * KS_WARP < 144) cwin[lane + k * KS_WARP] = b[k];
    }
    __syncwarp();
    const int row = lane >> 2, col = (lane & 3) * 2;
    if (fx == 0 && fy == 0) { v0 = cwin[(row + 1) * 12 + col + 1] << 6; v1 = cwin[(row + 1) * 12 + col + 2] << 6; }
    else if (fy == 0) {
        const uint8_t *p = cwin + (row + 1) * 12 + col;
        v0 = v1 = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) { int c = c_chroma_taps[fx][t]; v0 += c * p[t]; v1 += c * p[t + 1]; }
    } else if (fx == 0) {
        const uint8_t *p = cwin + row * 12 + col + 1;
        v0 = v1 = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) { int c = c_chroma_taps[fy][t]; v0 += c * p[t * 12]; v1 += c * p[t * 12 + 1]; }
    } else {
        for (int idx = lane; idx < 88; idx += KS_WARP) {
            int r = idx >> 3, c = idx & 7;
            const uint8_t *p = cwin + r * 12 + c;
            int a = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) a += c_chroma_taps[fx][t] * p[t];
            tmp[idx] = (int16_t)a;
        }
        __syncwarp();
        v0 = v1 = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) { int c = c_chroma_taps[fy][t]; v0 += c * tmp[(row + t) * 8 + col]; v1 += c * tmp[(row + t) * 8 + col + 1]; }
        v0 >>= 6; v1 >>= 6;
    }
    __syncwarp();
}

/* One warp per cell of a B picture: cost of bi-prediction (DefaultWeightedBi_c E@0x4350f0: (p0+p1+64)>>7 on the 14-bit
 * predictions) against the two single-list winners found by ks_me_kernel; writes the final motion (ks_cell + ks_cell_b) and
 * patches the prediction planes (list-0 prediction is already there).  Mirror of ora_b_picture's decision. */
__global__ void __launch_bounds__(KS_ME_WARPS * KS_WARP)
ks_bidir_kernel(KsPicParams pp, const uint8_t *__restrict__ srcY, KsPlanes ref0, KsPlanes ref1, const ks_cell *__restrict__ anchor_cells,
                int num0, int num1, int den, const ks_cell *__restrict__ cells1, const int *__restrict__ cost0, const int *__restrict__ cost1,
                KsPlanes pred1, ks_cell *__restrict__ cells, ks_cell_b *__restrict__ cells_b, KsPlanes pred)
{
    __shared__ __align__(16) KsWarpScratch scratch[KS_ME_WARPS];
    __shared__ uint8_t cwins[KS_ME_WARPS][144];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * KS_ME_WARPS + warp;
    if (cell >= pp.cw * pp.ch) return;
    KsWarpScratch *sc = &scratch[warp];
    const int cyc = cell / pp.cw, cxc = cell - cyc * pp.cw, x0 = cxc << 4, y0 = cyc << 4;
    const int W = pp.W, H = pp.H, CW = W >> 1, CH = H >> 1, lam = pp.lambda_sad_q4;
    const uint2 s = *reinterpret_cast<const uint2 *>(srcY + (size_t)(y0 + (lane >> 1)) * W + x0 + 8 * (lane & 1));
    const ks_cell c0 = cells[cell], c1 = cells1[cell];
    int ax = 0, ay = 0;
    if (anchor_cells) { ks_cell a = anchor_cells[cell]; if (!(a.flags & KS_F_INTRA)) { ax = a.mvx; ay = a.mvy; } }
    const int t0x = den ? (ax * num0) / den : 0, t0y = den ? (ay * num0) / den : 0, t1x = den ? (ax * num1) / den : 0, t1y = den ? (ay * num1) / den : 0;
    int p0[8], p1[8];
    {
        const int wx0 = (x0 + (c0.mvx >> 2) - 3) & ~3, wy0 = y0 + (c0.mvy >> 2) - 3;
        ks_load_window_mc(sc->win, ref0.p[0], W, H, wx0, wy0, lane);
        ks_interp16_raw(sc, x0 + (c0.mvx >> 2) - wx0, 3, c0.mvx & 3, c0.mvy & 3, lane, p0);
        const int wx1 = (x0 + (c1.mvx >> 2) - 3) & ~3, wy1 = y0 + (c1.mvy >> 2) - 3;
        __syncwarp();
        ks_load_window_mc(sc->win, ref1.p[0], W, H, wx1, wy1, lane);
        ks_interp16_raw(sc, x0 + (c1.mvx >> 2) - wx1, 3, c1.mvx & 3, c1.mvy & 3, lane, p1);
    }
    uint32_t b0 = 0, b1 = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { b0 |= (uint32_t)ks_clip8((p0[j] + p1[j] + 64) >> 7) << (8 * j); b1 |= (uint32_t)ks_clip8((p0[4 + j] + p1[4 + j] + 64) >> 7) << (8 * j); }
    const int metric = (pp.satd && pp.subpel > 0) ? (int)ks_satd16(b0, b1, s.x, s.y, lane) : (int)ks_warp_sum(__vsadu4(b0, s.x) + __vsadu4(b1, s.y));
    const int cb = metric + ((lam * (ks_mvbits(c0.mvx - t0x) + ks_mvbits(c0.mvy - t0y))) >> 4) + ((lam * (ks_mvbits(c1.mvx - t1x) + ks_mvbits(c1.mvy - t1y))) >> 4);
    int dir = 1, best = cost0[cell];
    if (cost1[cell] < best) { best = cost1[cell]; dir = 2; }
    if (cb < best) { best = cb; dir = 3; }
    if (lane == 0) {
        ks_cell c; c.mvx = (dir & 1) ? c0.mvx : 0; c.mvy = (dir & 1) ? c0.mvy : 0; c.cu_log2 = 4; c.flags = 0; c.intra_mode = 0; c.rsv = 0;
        ks_cell_b b; b.mvx1 = (dir & 2) ? c1.mvx : 0; b.mvy1 = (dir & 2) ? c1.mvy : 0; b.dir = (uint8_t)dir; b.rsv[0] = b.rsv[1] = b.rsv[2] = 0;
        cells[cell] = c; cells_b[cell] = b;
    }
    if (dir == 1) return;                                   /* prediction planes already hold the list-0 prediction */
    uint8_t *py = pred.p[0] + (size_t)(y0 + (lane >> 1)) * W + x0 + 8 * (lane & 1);
    const int crow = lane >> 2, ccol = (lane & 3) * 2;
    if (dir == 2) {
        *reinterpret_cast<uint2 *>(py) = *reinterpret_cast<const uint2 *>(pred1.p[0] + (size_t)(y0 + (lane >> 1)) * W + x0 + 8 * (lane & 1));
#pragma unroll
        for (int ci = 1; ci < 3; ci++) {
            size_t o = (size_t)((y0 >> 1) + crow) * CW + (x0 >> 1) + ccol;
            *reinterpret_cast<uint16_t *>(pred.p[ci] + o) = *reinterpret_cast<const uint16_t *>(pred1.p[ci] + o);
        }
        return;
    }
    *reinterpret_cast<uint2 *>(py) = make_uint2(b0, b1);
#pragma unroll 1
    for (int ci = 1; ci < 3; ci++) {
        int a0, a1, q0, q1;
        ks_mc_chroma8_raw(cwins[warp], &sc->tmp[0][0], ref0.p[ci], CW, CH, x0 >> 1, y0 >> 1, c0.mvx, c0.mvy, lane, a0, a1);
        ks_mc_chroma8_raw(cwins[warp], &sc->tmp[0][0], ref1.p[ci], CW, CH, x0 >> 1, y0 >> 1, c1.mvx, c1.mvy, lane, q0, q1);
        uint8_t *d = pred.p[ci] + (size_t)((y0 >> 1) + crow) * CW + (x0 >> 1) + ccol;
        d[0] = (uint8_t)ks_clip8((a0 + q0 + 64) >> 7); d[1] = (uint8_t)ks_clip8((a1 + q1 + 64) >> 7);
    }
}

void ks_launch_bidir(const KsPicParams &pp, const uint8_t *srcY, KsPlanes ref0, KsPlanes ref1, const ks_cell *anchor_cells, int num0, int num1, int den,
                     const ks_cell *cells1, const int *cost0, const int *cost1, KsPlanes pred1, ks_cell *cells, ks_cell_b *cells_b, KsPlanes pred, cudaStream_t st)
{
    int ncell = pp.cw * pp.ch;
    ks_bidir_kernel<<<(ncell + KS_ME_WARPS - 1) / KS_ME_WARPS, KS_ME_WARPS * KS_WARP, 0, st>>>(pp, srcY, ref0, ref1, anchor_cells, num0, num1, den, cells1, cost0, cost1, pred1, cells, cells_b, pred);
}
