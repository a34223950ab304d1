/*
 * ks_common.cuh -- device-side constants and helpers of the ks265 B200 hot path (sm_100a).  Included once,
 * by ks_kernels.cu (single translation unit: no relocatable device code needed).
 * Tables are the H.265 spec constants (identical to the reference's rodata: g_uiTr32 E@0x4d0740,
 * g_iLumaFilterCoeff E@0x4cc780, g_iChromaFilterCoeff E@0x4cc7c0, g_quantScales E@0x4cfb14,
 * g_invQuantScales E@0x4cfb20, uiTCTable E@0x4cc660, uiBetaTable E@0x4cc6a0, g_ucChromaScale E@0x4cfb40).
 */
#pragma once
#include <cstring>
#include "ks_launch.h"
#define KS_WARP 32

__constant__ int      c_dct[32][32];
__constant__ int      c_luma_taps[4][8];
__constant__ int      c_chroma_taps[8][4];
__constant__ int      c_luma_taps_packed[4][2];   /* taps 0..3 / 4..7 as 4 x s8 (dp4a operand) */
__constant__ int      c_chroma_taps_packed[8];
__constant__ int      c_vtaps_pk[4][2][3];        /* vertical luma taps as s8 pairs for dp2a over row-pair planes: [fy][first-row parity][word] */
/* x264 hexagon pattern hex2[8] (wraps so that dir-1..dir+1 index without a modulo) and the square refinement order */
__constant__ int      c_hex_dx[8] = {-1, -2, -1, 1, 2, 1, -1, -2};
__constant__ int      c_hex_dy[8] = {-2, 0, 2, 2, 0, -2, -2, 0};
__constant__ int      c_sq_dx[8] = {0, 0, -1, 1, -1, -1, 1, 1};
__constant__ int      c_sq_dy[8] = {-1, 1, 0, 0, -1, 1, -1, 1};
__constant__ uint8_t  c_tc_table[54];
__constant__ uint8_t  c_beta_table[52];
__constant__ uint8_t  c_chroma_qp[58];
__constant__ int      c_quant_scales[6];
__constant__ int      c_inv_quant_scales[6];
__constant__ uint16_t c_scan_tb[4][1024];         /* per log2 (2..5): scan position -> (y<<8)|x */
/* the residual kernel's shared-memory tables as one 16-byte-aligned global image (scan 8x8 | 16x16 | 32x32 | t0): every lane of a
 * CTA reads a different entry, which the constant cache would serialise 32 ways; from global memory it is 2 coalesced loads */
#define KS_RECON_TAB_U4 ((64 + 256 + 1024) * 2 / 16 + 256 * 4 / 16)
__device__ uint4 g_recon_tab[KS_RECON_TAB_U4];
__constant__ int8_t   c_intra_angle[35];
__constant__ int16_t  c_intra_inv_angle[35];

__device__ __forceinline__ int ks_clip3(int lo, int hi, int v) { return min(max(v, lo), hi); }
__device__ __forceinline__ int ks_clip8(int v) { return min(max(v, 0), 255); }
/* four s32 -> four u8 with saturation, v0 in the low byte (cvt.pack: d = c << 16 | sat(a) << 8 | sat(b)) */
__device__ __forceinline__ uint32_t ks_pack_sat4(int v0, int v1, int v2, int v3)
{
    uint32_t t, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(t));
    return d;
}
__device__ __forceinline__ int ks_mvbits(int d) { int a = abs(d); return a ? 2 * (32 - __clz(a)) + 1 : 1; }

/* u8 x s8 dot product with accumulate (dp4a.u32.s32): a = 4 unsigned bytes, b = 4 signed bytes */
__device__ __forceinline__ int ks_dp4a_us(unsigned a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
/* warp-wide integer sum, result in every lane: one REDUX instead of five shuffle+add steps (callers pack two 16-bit sums per word) */
__device__ __forceinline__ unsigned ks_warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ int ks_zidx(int x, int y)
{   /* z-scan index of the 8x8 block containing (x,y) inside its CTB (same order as the 16x16 one for blocks of 16 and up) */
    int cx = (x >> 3) & 7, cy = (y >> 3) & 7;
    return (cx & 1) | ((cy & 1) << 1) | ((cx & 2) << 1) | ((cy & 2) << 2) | ((cx & 4) << 2) | ((cy & 4) << 3);
}
/* H.265 6.4.1 z-scan availability at 8x8 granularity, one slice per picture */
__device__ __forceinline__ bool ks_avail(int W, int H, int ctw, int xc, int yc, int xn, int yn)
{
    if (xn < 0 || yn < 0 || xn >= W || yn >= H) return false;
    int ac = (yc >> 6) * ctw + (xc >> 6), an = (yn >> 6) * ctw + (xn >> 6);
    if (an != ac) return an < ac;
    return ks_zidx(xn, yn) < ks_zidx(xc, yc);
}

static void ks_upload_tables_impl()
{
    static const int8_t cosv[33] = {64,90,90,90,89,88,87,85,83,82,80,78,75,73,70,67,64,61,57,54,50,46,43,38,36,31,25,22,18,13,9,4,0};
    int dct[32][32];
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) {
            int m = (k * (2 * n + 1)) & 127;
            dct[k][n] = m <= 32 ? cosv[m] : m <= 64 ? -cosv[64 - m] : m <= 96 ? -cosv[m - 64] : cosv[128 - m];
        }
    cudaMemcpyToSymbol(c_dct, dct, sizeof(dct));
    static const int lt[4][8] = {{0,0,0,64,0,0,0,0},{-1,4,-10,58,17,-5,1,0},{-1,4,-11,40,40,-11,4,-1},{0,1,-5,17,58,-10,4,-1}};
    static const int ct[8][4] = {{0,64,0,0},{-2,58,10,-2},{-4,54,16,-2},{-6,46,28,-4},{-4,36,36,-4},{-4,28,46,-6},{-2,16,54,-4},{-2,10,58,-2}};
    cudaMemcpyToSymbol(c_luma_taps, lt, sizeof(lt));
    cudaMemcpyToSymbol(c_chroma_taps, ct, sizeof(ct));
    int ltp[4][2], ctp[8];
    for (int f = 0; f < 4; f++) for (int h = 0; h < 2; h++) {
        unsigned v = 0; for (int i = 0; i < 4; i++) v |= (unsigned)(lt[f][4 * h + i] & 255) << (8 * i); ltp[f][h] = (int)v; }
    for (int f = 0; f < 8; f++) { unsigned v = 0; for (int i = 0; i < 4; i++) v |= (unsigned)(ct[f][i] & 255) << (8 * i); ctp[f] = (int)v; }
    cudaMemcpyToSymbol(c_luma_taps_packed, ltp, sizeof(ltp));
    {   /* even first row: pairs (c0,c1)(c2,c3)(c4,c5)(c6,c7)(0,0); odd: (0,c0)(c1,c2)(c3,c4)(c5,c6)(c7,0) */
        int vt[4][2][3];
        for (int f = 0; f < 4; f++) for (int par = 0; par < 2; par++) {
            int seq[10];
            for (int i = 0; i < 10; i++) { int t = par ? i - 1 : i; seq[i] = (t >= 0 && t < 8) ? lt[f][t] : 0; }
            for (int w = 0; w < 3; w++) {
                unsigned v = 0;
                for (int b = 0; b < 4; b++) { int i = 4 * w + b; v |= (unsigned)((i < 10 ? seq[i] : 0) & 255) << (8 * b); }
                vt[f][par][w] = (int)v;
            }
        }
        cudaMemcpyToSymbol(c_vtaps_pk, vt, sizeof(vt));
    }
    cudaMemcpyToSymbol(c_chroma_taps_packed, ctp, sizeof(ctp));
    static const uint8_t tc[54] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24};
    static const uint8_t bt[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64};
    static const uint8_t cq[58] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51};
    cudaMemcpyToSymbol(c_tc_table, tc, sizeof(tc));
    cudaMemcpyToSymbol(c_beta_table, bt, sizeof(bt));
    cudaMemcpyToSymbol(c_chroma_qp, cq, sizeof(cq));
    static const int qs[6] = {26214,23302,20560,18396,16384,14564}, iq[6] = {40,45,51,57,64,72};
    cudaMemcpyToSymbol(c_quant_scales, qs, sizeof(qs));
    cudaMemcpyToSymbol(c_inv_quant_scales, iq, sizeof(iq));
    /* coefficient scan: diagonal over 4x4 coefficient groups x diagonal inside each group (H.265 6.5.3) */
    static uint16_t scan[4][1024];
    uint8_t d4[16], dcg[64];
    for (int l = 2; l <= 5; l++) {
        int ncg = 1 << (l - 2);
        for (int pass = 0; pass < 2; pass++) {
            int n = pass ? ncg : 4, i = 0, x = 0, y = 0; uint8_t *dst = pass ? dcg : d4;
            for (;;) { while (y >= 0) { if (x < n && y < n) dst[i++] = (uint8_t)((y << 3) | x); y--; x++; } y = x; x = 0; if (i >= n * n) break; }
        }
        for (int c = 0; c < ncg * ncg; c++)
            for (int k = 0; k < 16; k++)
                scan[l - 2][c * 16 + k] = (uint16_t)(((((dcg[c] >> 3) << 2) + (d4[k] >> 3)) << 8) | (((dcg[c] & 7) << 2) + (d4[k] & 7)));
    }
    cudaMemcpyToSymbol(c_scan_tb, scan, sizeof(scan));
    {
        static uint4 img[KS_RECON_TAB_U4];
        uint16_t *sp = reinterpret_cast<uint16_t *>(img);
        memcpy(sp, scan[1], 64 * 2); memcpy(sp + 64, scan[2], 256 * 2); memcpy(sp + 320, scan[3], 1024 * 2);
        int *tp = reinterpret_cast<int *>(sp + 1344);
        for (int i = 0; i < 256; i++) tp[i] = dct[2 * (i >> 4) + 1][i & 15];
        cudaMemcpyToSymbol(g_recon_tab, img, sizeof(img));
    }
    static const int8_t ang[35] = {0,0,32,26,21,17,13,9,5,2,0,-2,-5,-9,-13,-17,-21,-26,-32,-26,-21,-17,-13,-9,-5,-2,0,2,5,9,13,17,21,26,32};
    static const int16_t inv[35] = {0,0,0,0,0,0,0,0,0,0,0,-4096,-1638,-910,-630,-482,-390,-315,-256,-315,-390,-482,-630,-910,-1638,-4096,0,0,0,0,0,0,0,0,0};
    cudaMemcpyToSymbol(c_intra_angle, ang, sizeof(ang));
    cudaMemcpyToSymbol(c_intra_inv_angle, inv, sizeof(inv));
}
