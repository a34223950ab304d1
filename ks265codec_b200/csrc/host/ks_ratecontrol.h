/*
 * ks_ratecontrol.h -- host-side per-picture QP decision (north star: "CABAC and rate-control stay on the host").
 * Counterpart of the reference's EncRateControl.cpp entry points used by CHevcEncode::encodeFrame (E@0x4b5050):
 *   -rc 0  fixed QP: I = -qp; P-only streams cascade P = -qp + {1,3,2,3}[poc % 4] exactly like the reference at -bframes 0
 *          ([probe] -psnr 2 tables at veryfast/superfast, several -qp/-iper); with B pictures P = -qp+1, B = -qp+3; -fixqp 1 = flat
 *   -rc 3  CRF: x264/x265-style constant rate factor, qscale ~ complexity^(1-qcomp); the complexity measure is the device
 *          motion search's own cost sum of the preceding P pictures of the shard (the reference uses its lookahead's half-resolution
 *          SATD; a lookahead kernel is SURVEY 8(f) row f2).  One state per closed-GOP shard, so shards stay independent (8e).
 * -rc 1/2 (ABR/CBR with -br) are not implemented.
 */
#ifndef KS_RATECONTROL_H
#define KS_RATECONTROL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ks_rc {
    int mode, qp, fixqp, bframes;
    double crf;
    double cplx_sum, cplx_cnt;      /* exponentially decayed sum / count of per-picture search cost */
    double base_cplx;               /* cost of a picture that gets exactly QP = crf */
    int qp_min, qp_max;
} ks_rc;

/* cells = number of 16x16 cells of the coded picture; returns 0, or -1 for an unsupported mode */
int  ks_rc_init(ks_rc *rc, int mode, int qp, int fixqp, double crf, int cells, int bframes);
/* QP of the next picture to be coded (poc = display index inside the shard), given everything fed to ks_rc_update so far */
int  ks_rc_picture_qp(const ks_rc *rc, int slice_type, int poc);
/* QP whose lambda the device's mode decision / RD zero-out use for this picture (ks_pic_params.lambda_qp), >= qp */
int  ks_rc_lambda_qp(const ks_rc *rc, int slice_type, int poc, int qp);
/* feed a finished picture: me_cost = ks_pic_out.me_cost (0 for I and B pictures, which do not update the model) */
void ks_rc_update(ks_rc *rc, int slice_type, uint64_t me_cost);

#ifdef __cplusplus
}
#endif
#endif
