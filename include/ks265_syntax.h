/*
 * ks265_syntax.h -- the per-picture "frame syntax" block: everything the device hot path decides for one
 * picture, in the layout the CUDA kernels write to (pinned) host memory and the host entropy coder reads.
 *
 * This is the data format on the boundary between the B200 hot path (ME / transform+quant / loop filter)
 * and the host-side stages the north star keeps on the CPU (CABAC, headers, rate control).
 * In the reference the same information lives in TCodingUnit/TPredUnit/TTransUnit/TNborData objects walked by
 * CCtuSbac::processCtuSbac (E@0x46dd00) and encodeCoeffNxN (E@0x46df40); we do not mirror those layouts
 * ("design device structs fresh").
 */
#ifndef KS265_SYNTAX_H
#define KS265_SYNTAX_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KS_CTU_LOG2   6      /* CTB 64x64 (reference: CTB 64) */
#define KS_CTU        64
#define KS_CELL_LOG2  4      /* side-info granularity = 16x16 luma (inter CUs are 16..64) */
#define KS_CELL       16
#define KS_MIN_CB_LOG2 3     /* minimum coding block 8x8: an INTRA cell may be four 8x8 CUs (cu_log2 == 3, see ks_cell) */
#define KS_MAX_TB_LOG2 5     /* TB 4..32 */

enum { KS_SLICE_B = 0, KS_SLICE_P = 1, KS_SLICE_I = 2 };

/* one 16x16 luma cell (8 bytes).  A CU of size 2^cu_log2 covers (2^cu_log2/16)^2 cells, all carrying the same
 * cu_log2/mv/mode; cbf bits are those of the transform unit (min(CU,32)) covering the cell.
 * cu_log2 == 3 (intra only): the cell holds FOUR 8x8 intra CUs (2Nx2N, luma TB 8x8, chroma TBs 4x4), sub-block k = (x half) | (y half) << 1:
 *   luma modes in the vector bytes (KS_SUB_MODE), per-sub-block cbf bits in intra_mode / rsv (KS_SUB_CBF_*), flags = KS_F_INTRA | OR of them. */
typedef struct ks_cell {
    int16_t mvx, mvy;        /* quarter-sample luma MV, list 0 (inter CUs) */
    uint8_t cu_log2;         /* 4, 5 or 6 */
    uint8_t flags;           /* bit0 intra; bit1 cbf_luma; bit2 cbf_cb; bit3 cbf_cr */
    uint8_t intra_mode;      /* luma intra mode 0..34 (chroma uses DM = same mode) */
    uint8_t rsv;
} ks_cell;
#define KS_SUB_MODE(c, k)   ((int)(((k) & 2 ? (uint16_t)(c)->mvy : (uint16_t)(c)->mvx) >> (((k) & 1) * 8)) & 255)
#define KS_SUB_CBF_Y(c, k)  (((c)->intra_mode >> (k)) & 1)
#define KS_SUB_CBF_CB(c, k) (((c)->intra_mode >> (4 + (k))) & 1)
#define KS_SUB_CBF_CR(c, k) (((c)->rsv >> (k)) & 1)
#define KS_F_INTRA 1
#define KS_F_CBF_Y 2
#define KS_F_CBF_CB 4
#define KS_F_CBF_CR 8

/* B pictures only: list-1 motion + prediction direction of the cell (same raster as ks_cell) */
typedef struct ks_cell_b {
    int16_t mvx1, mvy1;      /* quarter-sample luma MV, list 1 (0 when unused) */
    uint8_t dir;             /* 1 = list 0 only (MV in ks_cell), 2 = list 1 only, 3 = bi-prediction */
    uint8_t rsv[3];
} ks_cell_b;

typedef struct ks_sao_param {
    uint8_t type;            /* 0 off, 1 band, 2 edge */
    uint8_t band_or_class;   /* band position (0..31) or EO class (0..3) */
    int8_t  off[4];          /* signed offsets as applied (EO: +,+,-,-) */
} ks_sao_param;

/* per-CTU record (72 bytes): which 4x4 coefficient groups are non-zero + where their levels sit in the pool */
typedef struct ks_ctu_syn {
    uint16_t cg_y[16];       /* bit x of row y: luma CG (x,y) of this CTU has a non-zero level */
    uint8_t  cg_cb[8];
    uint8_t  cg_cr[8];
    uint32_t cg_base;        /* index (in CG units of 16 int16) of this CTU's first CG in the level pool;
                                CGs are stored Y rows, Cb rows, Cr rows, each row left to right */
    ks_sao_param sao[3];     /* Y, Cb, Cr (Cb/Cr share type and EO class) */
    uint8_t  rsv[2];
} ks_ctu_syn;

/* view of one picture's syntax (pointers into one contiguous block) */
typedef struct ks_frame_syn {
    int width, height;       /* coded (padded to a multiple of 16) luma size */
    int cells_w, cells_h;    /* width/16, height/16 */
    int ctus_w, ctus_h;
    int slice_type;          /* KS_SLICE_* */
    int qp;                  /* slice QP (flat inside the picture) */
    int poc;
    const ks_cell    *cells; /* cells_w*cells_h, raster */
    const ks_ctu_syn *ctus;  /* ctus_w*ctus_h, raster */
    const int16_t    *levels;/* pool: n_cg * 16 int16, each CG row-major 4x4 */
    uint32_t          n_cg;
    const ks_cell_b  *cells_b;/* B pictures: cells_w*cells_h entries, else NULL */
} ks_frame_syn;

#ifdef __cplusplus
}
#endif
#endif
