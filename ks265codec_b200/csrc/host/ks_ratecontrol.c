/* ks_ratecontrol.c -- see ks_ratecontrol.h */
#include "ks_ratecontrol.h"
#include "ks265_syntax.h"
#include <math.h>
#include <string.h>

#define KS_RC_QCOMP        0.6      /* x264 default */
#define KS_RC_BASE_COST    512.0    /* search cost per 16x16 cell (2 per luma sample) that maps to QP = crf */
#define KS_RC_IP_OFFSET    3        /* 6*log2(1.4): I pictures below the P level */
#define KS_RC_PB_OFFSET    2        /* 6*log2(1.3): non-reference B pictures above it */

int ks_rc_init(ks_rc *rc, int mode, int qp, int fixqp, double crf, int cells, int bframes)
{
    memset(rc, 0, sizeof(*rc));
    if (mode != 0 && mode != 3) return -1;
    rc->mode = mode; rc->qp = qp; rc->fixqp = fixqp; rc->crf = crf; rc->bframes = bframes;
    rc->base_cplx = KS_RC_BASE_COST * (double)cells;
    rc->qp_min = 0; rc->qp_max = 51;
    return 0;
}

static int clip_qp(const ks_rc *rc, int q) { return q < rc->qp_min ? rc->qp_min : (q > rc->qp_max ? rc->qp_max : q); }

/* P-only streams: the reference's 4-picture QP cascade (every 4th P is the better-quality anchor of the next three) */
static int p_cascade(const ks_rc *rc, int poc)
{
    static const int off[4] = {1, 3, 2, 3};
    return rc->bframes ? 1 : off[poc & 3];
}

int ks_rc_picture_qp(const ks_rc *rc, int slice_type, int poc)
{
    if (rc->mode == 0) {
        int q = rc->qp;
        if (!rc->fixqp) q += slice_type == KS_SLICE_I ? 0 : (slice_type == KS_SLICE_P ? p_cascade(rc, poc) : 3);
        return clip_qp(rc, q);
    }
    double q = rc->crf;
    if (rc->cplx_cnt > 0.0) {
        double cplx = rc->cplx_sum / rc->cplx_cnt;
        if (cplx < 1.0) cplx = 1.0;
        q += 6.0 * (1.0 - KS_RC_QCOMP) * log2(cplx / rc->base_cplx);
    }
    if (slice_type == KS_SLICE_I) q -= KS_RC_IP_OFFSET;
    else if (slice_type == KS_SLICE_B) q += KS_RC_PB_OFFSET;
    else q += p_cascade(rc, poc) - 1;
    return clip_qp(rc, (int)floor(q + 0.5));
}

/* The QP whose lambda drives the device's CU/merge decision and RD zero-out.  P-only streams: the three pictures between two
 * better-quality anchors of the 4-picture cascade decide with the lambda of QP+3 (twice the Lagrangian; HM's low-delay configuration scales
 * lambda by 2..4 on exactly those pictures, and the reference's per-picture sizes and PSNRs at -bframes 0 show the same pattern [probe]). */
int ks_rc_lambda_qp(const ks_rc *rc, int slice_type, int poc, int qp)
{
    int d = (slice_type == KS_SLICE_P && !rc->bframes && !rc->fixqp && (poc & 3)) ? 3 : 0;
    return qp + d > 51 ? 51 : qp + d;
}

void ks_rc_update(ks_rc *rc, int slice_type, uint64_t me_cost)
{
    if (rc->mode != 3 || slice_type != KS_SLICE_P) return;
    rc->cplx_sum = rc->cplx_sum * 0.5 + (double)me_cost;       /* x264's short-term blur: half-life of one picture */
    rc->cplx_cnt = rc->cplx_cnt * 0.5 + 1.0;
}
