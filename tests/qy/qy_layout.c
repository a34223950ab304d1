/* Layout check of include/ks265_qyabi.h against the reference's own public header (compiled only where /root/reference exists).
 * Every mirrored struct must have the reference's size and every field the reference's offset. */
#include <stddef.h>
#include <stdio.h>
#include "qy265enc.h"                 /* the reference's header, found through -I */
#define KS265_QYABI_TYPES_ONLY
#include "ks265_qyabi.h"

#define SAME(rt, rf, kt, kf) _Static_assert(offsetof(rt, rf) == offsetof(kt, kf), #rt "." #rf " vs " #kt "." #kf)
_Static_assert(sizeof(QY265EncConfig) == sizeof(ksqy_config), "config size");
_Static_assert(sizeof(QY265YUV) == sizeof(ksqy_yuv), "yuv size");
_Static_assert(sizeof(QY265Picture) == sizeof(ksqy_picture), "picture size");
_Static_assert(sizeof(QY265Nal) == sizeof(ksqy_nal), "nal size");
#define C(rf, kf) SAME(QY265EncConfig, rf, ksqy_config, kf)
C(pAuth, auth); C(tune, tune); C(preset, preset); C(latency, latency); C(profileId, profile_id); C(bHeaderBeforeKeyframe, headers_before_keyframe);
C(picWidth, width); C(picHeight, height); C(frameRate, fps); C(bframes, bframes); C(temporalLayer, temporal_layer);
C(vpp_denoise, vpp_denoise); C(vpp_edge, vpp_edge); C(vpp_color, vpp_color); C(vpp_hdr, vpp_hdr); C(vpp_hdr_strength, vpp_hdr_strength); C(vpp_hdr_iter, vpp_hdr_iter);
C(vpp_hdr_sigma_s, vpp_hdr_sigma_s); C(vpp_hdr_sigma_r, vpp_hdr_sigma_r); C(vpp_recur_filter, vpp_recur_filter);
C(rc, rc); C(bitrateInkbps, bitrate_kbps); C(vbv_buffer_size, vbv_buffer_size); C(vbv_max_rate, vbv_max_rate); C(vbv_min_rate, vbv_min_rate); C(qp, qp); C(crf, crf);
C(visual_quality, visual_quality); C(iIntraPeriod, intra_period); C(qpmin, qp_min); C(qpmax, qp_max); C(enFrameSkip, frame_skip);
C(enWavefront, wavefront); C(enFrameParallel, frame_parallel); C(threads, threads); C(vui_parameters_present_flag, vui_present);
C(vui.video_signal_type_present_flag, vui.signal_type_present); C(vui.video_format, vui.video_format); C(vui.video_full_range_flag, vui.full_range);
C(vui.colour_description_present_flag, vui.colour_desc_present); C(vui.colour_primaries, vui.primaries); C(vui.transfer_characteristics, vui.transfer); C(vui.matrix_coeffs, vui.matrix);
C(logLevel, log_level); C(lookahead, lookahead); C(calcPsnr, calc_psnr); C(calcSsim, calc_ssim); C(shortLoadingForPlayer, short_loading); C(iPass, pass);
C(statFileName, stat_file); C(fRateTolerance, rate_tolerance); C(rdoq, rdoq); C(me, me); C(part, part); C(do64, do64); C(tuInter, tu_inter); C(tuIntra, tu_intra);
C(smooth, smooth); C(transskip, transskip); C(subme, subme); C(satdInter, satd_inter); C(satdIntra, satd_intra); C(searchrange, search_range); C(refnum, ref_num);
C(ref0, ref0); C(sao, sao); C(longTermRef, long_term_ref); C(iAqMode, aq_mode); C(fAqStrength, aq_strength); C(rasl, rasl);
SAME(QY265YUV, iWidth, ksqy_yuv, width); SAME(QY265YUV, iHeight, ksqy_yuv, height); SAME(QY265YUV, pData, ksqy_yuv, plane); SAME(QY265YUV, iStride, ksqy_yuv, stride);
SAME(QY265Picture, iSliceType, ksqy_picture, slice_type); SAME(QY265Picture, poc, ksqy_picture, poc); SAME(QY265Picture, pts, ksqy_picture, pts);
SAME(QY265Picture, dts, ksqy_picture, dts); SAME(QY265Picture, yuv, ksqy_picture, yuv);
SAME(QY265Nal, naltype, ksqy_nal, nal_type); SAME(QY265Nal, tid, ksqy_nal, tid); SAME(QY265Nal, iSize, ksqy_nal, size); SAME(QY265Nal, pts, ksqy_nal, pts);
SAME(QY265Nal, pPayload, ksqy_nal, payload);
_Static_assert(QY_OK == KSQY_OK && (int)QY_FAIL == KSQY_FAIL && (int)QY_OUTOFMEMORY == KSQY_OUTOFMEMORY && (int)QY_POINTER == KSQY_POINTER && (int)QY_NOTSUPPORTED == KSQY_NOTSUPPORTED, "codes");
_Static_assert(QY265_PARAM_BAD_NAME == KSQY_BAD_NAME && QY265_PARAM_BAD_VALUE == KSQY_BAD_VALUE, "parse codes");
int main(void) { printf("layout ok: config %zu bytes\n", sizeof(QY265EncConfig)); return 0; }
