/* A caller of the reference's encoder API, written against the reference's own header (qy265enc.h, found through -I) the way its
 * Android demo drives it (Android_demo/.../encoderwrapper.c:296-414: ConfigDefaultPreset, Open, one EncodeFrame per picture read into
 * the SAME buffer, flush while DelayedFrames, Close).  Linked against libks265qy.so it must run unchanged.
 * usage: qy_caller in.yuv width height out.265 [name=value ...] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "qy265enc.h"

static void on_log(const char *msg) { fputs(msg, stderr); }

static int put(QY265Nal *nal, int n, FILE *f)
{
    for (int i = 0; i < n; i++) if (fwrite(nal[i].pPayload, (size_t)nal[i].iSize, 1, f) != 1) return -1;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s in.yuv width height out.265 [name=value ...]\n", argv[0]); return 2; }
    FILE *in = fopen(argv[1], "rb"), *out = fopen(argv[4], "wb");
    if (!in || !out) { perror("open"); return 2; }
    QY265EncConfig param;
    char preset[] = "veryfast", latency[] = "default";
    if (QY265ConfigDefaultPreset(&param, preset, NULL, latency) < 0) return 3;
    param.picWidth = atoi(argv[2]); param.picHeight = atoi(argv[3]);
    for (int i = 5; i < argc; i++) {
        char *eq = strchr(argv[i], '=');
        if (!eq) return 2;
        *eq = 0;
        if (QY265ConfigParse(&param, argv[i], eq + 1)) { fprintf(stderr, "bad option %s\n", argv[i]); return 3; }
    }
    QY265SetLogPrintf(on_log);
    QY265YUV yuv;
    const size_t luma = (size_t)param.picWidth * param.picHeight, chroma = luma / 4;
    yuv.pData[0] = (unsigned char *)malloc(luma * 3 / 2); yuv.pData[1] = yuv.pData[0] + luma; yuv.pData[2] = yuv.pData[0] + luma * 5 / 4;
    yuv.iWidth = param.picWidth; yuv.iHeight = param.picHeight; yuv.iStride[0] = yuv.iWidth; yuv.iStride[1] = yuv.iStride[2] = yuv.iWidth / 2;
    int err = 0;
    void *h = QY265EncoderOpen(&param, &err);
    if (!h) { fprintf(stderr, "open failed: 0x%x\n", (unsigned)err); return 4; }
    QY265Picture pic, pic_out;
    QY265Nal *nal; int n_nal, frames = 0, got = 0;
    memset(&pic, 0, sizeof(pic)); memset(&pic_out, 0, sizeof(pic_out));
    pic.yuv = &yuv;
    for (;; frames++) {
        if (fread(yuv.pData[0], 1, luma, in) != luma || fread(yuv.pData[1], 1, chroma, in) != chroma || fread(yuv.pData[2], 1, chroma, in) != chroma) break;
        pic.pts = frames;
        if (QY265EncoderEncodeFrame(h, &nal, &n_nal, &pic, &pic_out, 0) < 0 || put(nal, n_nal, out)) return 5;
        got += n_nal > 0;
    }
    while (QY265EncoderDelayedFrames(h)) {
        if (QY265EncoderEncodeFrame(h, &nal, &n_nal, NULL, &pic_out, 0) < 0 || put(nal, n_nal, out)) return 5;
        got += n_nal > 0;
    }
    QY265EncoderClose(h);
    fclose(out); fclose(in);
    printf("%d frames in, %d access units out\n", frames, got);
    return frames == got ? 0 : 6;
}
