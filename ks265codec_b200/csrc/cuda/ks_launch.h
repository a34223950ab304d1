/*
 * ks_launch.h -- host-visible launch interface of the CUDA hot path (internal to libks265gpu.so).
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "ks265_syntax.h"

/* one picture's planes: Y (pitch = W), U, V (pitch = W/2); tightly packed, W and H multiples of 16 */
struct KsPlanes { uint8_t *p[3]; };
struct KsLevels { int16_t *p[3]; };

/* per-picture launch parameters (passed by value to kernels) */
struct KsPicParams {
    int W, H;               /* coded luma size */
    int dW, dH;             /* display luma size (PSNR is taken over the display area, like the reference's) */
    int cw, ch;             /* 16x16 cells */
    int ctw, cth;           /* 64x64 CTUs */
    int slice_type, qp, qpc;
    int lambda_sad_q4, lambda_sse_q4;
    int lambda_dec_q4, rdz_lambda_q4;   /* lambdas of ks_pic_params.lambda_qp_delta: CU/merge decision (SAD domain) and RD zero-out of luma blocks (SSE domain; 0 = off) */
    int me_range, me_iters, subpel, satd, me_method;
    int sign_hiding, sao, strong_intra;
    int beta_offset_div2, tc_offset_div2;
    int pred_num, pred_den;  /* motion-search predictor = co-located vector * pred_num / pred_den (C division); den 0 = vector as is */
};

/* all launches are asynchronous on `st` */
/* once per device (thread-safe): constant tables + the kernels' dynamic shared-memory attributes; device < 0 = the current device */
int ks_init_device(int device);
/* costs: per-cell winning cost incl. vector bits (B pictures); dists: per-cell distortion of the winner (P pictures, input of ks_launch_decide) */
void ks_launch_me(const KsPicParams &pp, const uint8_t *srcY, KsPlanes ref, const ks_cell *prev_cells, ks_cell *cells, KsPlanes pred, int *costs, int *dists, unsigned long long *cost_sum, cudaStream_t st);
/* P pictures: candidate distortions + CU quadtree / merge decision per CTU (reads the search field mv0/dist0, writes the final cells and
 * re-predicts the cells whose vector changed) */
/* n_intra (device int): number of cells the decision made intra CUs (their prediction + residual: ks_launch_recon_intra in masked mode) */
void ks_launch_decide(const KsPicParams &pp, const uint8_t *srcY, KsPlanes ref, const ks_cell *mv0, const int *dist0, void *cands_ws, ks_cell *cells, KsPlanes pred, int *n_intra, cudaStream_t st);
size_t ks_decide_workspace_bytes(int nctu);
/* B pictures: per cell best of list 0 / list 1 / bi-prediction; finalises cells, cells_b and the prediction planes */
void ks_launch_bidir(const KsPicParams &pp, const uint8_t *srcY, KsPlanes ref0, KsPlanes ref1, const ks_cell *anchor_cells, int num0, int num1, int den,
                     const ks_cell *cells1, const int *cost0, const int *cost1, KsPlanes pred1, ks_cell *cells, ks_cell_b *cells_b, KsPlanes pred, cudaStream_t st);
void ks_launch_recon_inter(const KsPicParams &pp, KsPlanes src, KsPlanes pred, KsPlanes rec, KsLevels lv, ks_cell *cells, const ks_cell_b *cells_b, cudaStream_t st);
/* n_intra == NULL: I picture, every cell; else P picture: only the cells flagged KS_F_INTRA (nothing if *n_intra == 0), inter-slice rounding */
void ks_launch_recon_intra(const KsPicParams &pp, KsPlanes src, KsPlanes rec, KsLevels lv, ks_cell *cells, int *sync_ws, const int *n_intra, void *modes_ws, cudaStream_t st);
size_t ks_intra_workspace_bytes(int ncell);
void ks_launch_deblock(const KsPicParams &pp, KsPlanes rec, const ks_cell *cells, const ks_cell_b *cells_b, cudaStream_t st);
/* tm[3]: tensor maps of the three `deb` planes (box 96x66 / 64x34 bytes); tma_mask bit c = component c is staged by TMA */
/* sse_ctu: per-CTU squared error of the three planes (3 x u32 per CTU) or NULL; ks_launch_pack sums them into the picture SSE */
void ks_launch_sao(const KsPicParams &pp, KsPlanes src, KsPlanes deb, KsPlanes out, ks_ctu_syn *ctus, uint32_t *sse_ctu,
                   const CUtensorMap *tm, int tma_mask, cudaStream_t st);
void ks_launch_pack(const KsPicParams &pp, KsLevels lv, ks_ctu_syn *ctus, int16_t *pool, uint32_t *n_cg, uint32_t *scan_ws, const uint32_t *sse_ctu, unsigned long long *sse_out, cudaStream_t st);
/* number of kernel launches each stage issues (for bench.py's gpu_launches accounting) */
enum { KS_LAUNCHES_DECIDE = 3, KS_LAUNCHES_ME = 1, KS_LAUNCHES_RECON = 1, KS_LAUNCHES_INTRA = 2, KS_LAUNCHES_DEBLOCK = 2, KS_LAUNCHES_SAO = 2, KS_LAUNCHES_PACK = 3 };
