/* ks_kat.h -- device-pointer launchers of the known-answer kernels (internal; see ks265_gpu.h ks_gpu_kat_*) */
#pragma once
#include <stdint.h>
int ks_kat_sad16_dev(const uint8_t *a, const uint8_t *b16, uint32_t *out);
int ks_kat_satd16_dev(const uint8_t *a, const uint8_t *b16, uint32_t *out);
int ks_kat_interp_dev(const uint8_t *plane, int w, int h, int x, int y, int mvx, int mvy, uint8_t *dst);
int ks_kat_tb_dev(int log2n, const uint8_t *src, const uint8_t *pred, int qp, int intra_slice, int sign_hiding, int16_t *levels, uint8_t *recon, int *cbf);
