/* ora_tables.c -- constant tables of the HEVC hot path (TEST INFRASTRUCTURE, see ks_oracle.h).
 * Values are the H.265 spec tables; tests/test_oracle_tables.py checks them against the dumps of the
 * reference's rodata (g_uiTr32 E@0x4d0740, g_iLumaFilterCoeff E@0x4cc780, ...) stored in tests/golden/. */
#include "ks_oracle.h"

/* cos(m*pi/64) magnitudes used by the HEVC core transform, m = 0..32 */
static const int8_t cosv[33] = {64,90,90,90,89,88,87,85,83,82,80,78,75,73,70,67,64,61,57,54,50,46,43,38,36,31,25,22,18,13,9,4,0};

int8_t ora_dct32_storage[32][32];
static int tables_ready;
static void __attribute__((constructor)) ora_tables_init(void)
{
    if (tables_ready) return;
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) {
            int m = (k * (2 * n + 1)) & 127, v;
            if (m <= 32) v = cosv[m];
            else if (m <= 64) v = -cosv[64 - m];
            else if (m <= 96) v = -cosv[m - 64];
            else v = cosv[128 - m];
            ora_dct32_storage[k][n] = (int8_t)v;
        }
    tables_ready = 1;
}
/* exported under the const name through an alias so users see a const table */
extern const int8_t ora_dct32[32][32] __attribute__((alias("ora_dct32_storage")));

const int8_t ora_dst4[4][4] = {{29,55,74,84},{74,74,0,-74},{84,-29,-74,55},{55,-84,74,-29}};
const int8_t ora_luma_filter[4][8] = {
    {0,0,0,64,0,0,0,0},{-1,4,-10,58,17,-5,1,0},{-1,4,-11,40,40,-11,4,-1},{0,1,-5,17,58,-10,4,-1}};
const int8_t ora_chroma_filter[8][4] = {
    {0,64,0,0},{-2,58,10,-2},{-4,54,16,-2},{-6,46,28,-4},{-4,36,36,-4},{-4,28,46,-6},{-2,16,54,-4},{-2,10,58,-2}};
const uint8_t ora_tc_table[54] = {
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24};
const uint8_t ora_beta_table[52] = {
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64};
const uint8_t ora_chroma_qp[58] = {
    0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,
    29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51};
const int ora_quant_scales[6] = {26214,23302,20560,18396,16384,14564};
const int ora_inv_quant_scales[6] = {40,45,51,57,64,72};
