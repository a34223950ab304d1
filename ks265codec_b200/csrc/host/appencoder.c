/*
 * appencoder.c -- command-line front end with the reference AppEncoder's flag surface (README.md:8-82 of
 * ksvc/ks265codec; AppEncCfg::ParseCfg E@0x4c4a40): -i -b -o -wdt -hgt -fr -frms -preset -rc -qp -iper -threads -psnr
 * -md5 -fixqp ... and the same summary lines ("Total Frames: N, test time: T ms, FPS: F", "bitrate, psnr: ...",
 * "H265 encoder passed!!!") so log parsers written for the reference (encoderwrapper.c:249-277) keep working.
 * New: -gpus N (GOP shards round-robin over N devices) and -streams S (concurrent GOP shards per device).
 * Flags of the reference that have no meaning on the device path are accepted and ignored with a warning.
 */
#define _GNU_SOURCE
#include "ks265_enc.h"
#include "ks_bitstream.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

typedef struct {
    const char *in, *bs, *rec, *preset;
    int w, h, frms, psnr, md5, gpus, streams;
    double fr;
    ks265_config cfg;
} app_cfg;

typedef struct {
    uint8_t *bs; long bs_bytes; ks265_gop_stats st; int first, n; int err, done;
    ks265_pic_stat *pics;          /* -psnr 2: per-picture table */
} shard_t;

typedef struct {
    app_cfg *a; shard_t *shards; int nshards; int next; pthread_mutex_t mu;
    int in_fd, rec_fd; size_t fsz;
    int live_workers;              /* workers that opened an encoder */
    double enc_ms;                 /* time inside ks265_encoder_encode_gop, summed over workers */
} job_t;

typedef struct { job_t *job; int device; } worker_arg;
typedef struct { job_t *job; shard_t *sh; } read_arg;

static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

/* picture source of one shard (reference: CInputYUV::startReadThread E@0x4cbe80 + readOneFrame): the encoder pulls picture i+1 while the device
 * works on picture i, straight into page-locked staging memory -- no shard-sized buffers, no extra copy */
static int read_picture(void *opaque, int disp, uint8_t *dst)
{
    read_arg *r = (read_arg *)opaque; job_t *j = r->job;
    size_t got = 0;
    while (got < j->fsz) {
        ssize_t k = pread(j->in_fd, dst + got, j->fsz - got, (off_t)(j->fsz * (size_t)(r->sh->first + disp) + got));
        if (k <= 0) return -1;
        got += (size_t)k;
    }
    return 0;
}
static int claim(job_t *j) { pthread_mutex_lock(&j->mu); int s = j->next < j->nshards ? j->next++ : -1; pthread_mutex_unlock(&j->mu); return s; }

static void *worker(void *argp)
{
    worker_arg *wa = (worker_arg *)argp; job_t *j = wa->job; app_cfg *a = j->a;
    ks265_config cfg = a->cfg; cfg.device = wa->device;
    int err = 0;
    ks265_encoder *enc = ks265_encoder_open(&cfg, &err);
    /* a worker that cannot open its encoder (e.g. out of device memory on the Nth stream) just leaves: the others take its shards */
    if (!enc) { fprintf(stderr, "appencoder: cannot open an encoder on device %d (error %d)\n", wa->device, err); return NULL; }
    pthread_mutex_lock(&j->mu); j->live_workers++; pthread_mutex_unlock(&j->mu);
    int maxn = j->shards[0].n;     /* shard 0 is the longest */
    uint8_t *recon = j->rec_fd >= 0 ? (uint8_t *)malloc(j->fsz * (size_t)maxn) : NULL;
    size_t cap = j->fsz * (size_t)maxn / 2 + (4 << 20);                 /* grown and retried on -28 (CABAC worst case is ~1.4 x the picture bytes) */
    uint8_t *bsbuf = (uint8_t *)malloc(cap);
    double enc_ms = 0;
    if (!bsbuf || (j->rec_fd >= 0 && !recon)) { fprintf(stderr, "appencoder: out of memory for %d-picture shard buffers\n", maxn); goto out; }
    for (int cur = claim(j); cur >= 0; cur = claim(j)) {
        shard_t *sh = &j->shards[cur];
        read_arg ra = {j, sh};
        if (a->psnr >= 2) sh->pics = (ks265_pic_stat *)calloc((size_t)sh->n, sizeof(ks265_pic_stat));
        ks265_encoder_set_picture_stats(enc, sh->pics, sh->pics ? sh->n : 0);
        double t0 = now_ms();
        long n = ks265_encoder_encode_gop_cb(enc, read_picture, &ra, sh->n, bsbuf, cap, recon, &sh->st);
        if (n == -28) {             /* output buffer too small: the handle stays usable (pictures in flight are dropped), retry with the worst-case size */
            size_t big = j->fsz * (size_t)maxn * 2 + (4 << 20);
            uint8_t *nb = (uint8_t *)realloc(bsbuf, big);
            if (nb) { bsbuf = nb; cap = big; n = ks265_encoder_encode_gop_cb(enc, read_picture, &ra, sh->n, bsbuf, cap, recon, &sh->st); }
        }
        enc_ms += now_ms() - t0;
        if (n < 0) { sh->err = (int)n; continue; }
        sh->bs = (uint8_t *)malloc((size_t)n);
        if (!sh->bs) { sh->err = -12; continue; }
        memcpy(sh->bs, bsbuf, (size_t)n);
        sh->bs_bytes = n; sh->done = 1;
        if (recon && pwrite(j->rec_fd, recon, j->fsz * (size_t)sh->n, (off_t)(j->fsz * (size_t)sh->first)) < 0) sh->err = -5;
    }
out:
    pthread_mutex_lock(&j->mu); j->enc_ms += enc_ms; pthread_mutex_unlock(&j->mu);
    free(bsbuf); free(recon);
    ks265_encoder_close(enc);
    return NULL;
}

static void usage(void)
{
    printf("ks265 B200 appencoder (AppEncoder-compatible)\n"
           "  -i <file.yuv> -wdt <w> -hgt <h> [-fr fps] [-frms n] [-b out.265] [-o recon.yuv]\n"
           "  [-preset ultrafast|superfast|veryfast|fast|medium|slow|veryslow|placebo] [-rc 0] [-qp q] [-iper n] [-fixqp 0|1]\n"
           "  [-sao 0..4] [-subme 0..2] [-merange n] [-bframes n] [-me 0|1] [-rc 0|3 -crf x] [-psnr 0|1|2] [-md5 0|1] [-threads n] [-gpus n] [-streams n]\n");
}

int main(int argc, char **argv)
{
    app_cfg a; memset(&a, 0, sizeof(a));
    a.fr = 30.0; a.preset = "veryfast"; a.gpus = 1; a.streams = 4; a.frms = -1;
    int qp = 27, iper = 128, fixqp = 0, rc = 0, sao = -1, subme = -1, merange = -1, bframes = -1, me = -1, threads = 0; double crf = -1;
    for (int i = 1; i < argc; i++) {
        const char *k = argv[i], *v = i + 1 < argc ? argv[i + 1] : NULL;
        if (!strcmp(k, "-v") || !strcmp(k, "-h") || !strcmp(k, "--help")) { usage(); return 0; }
        if (!v) { fprintf(stderr, "appencoder: option %s needs a value\n", k); return 2; }
        i++;
        if (!strcmp(k, "-i")) a.in = v; else if (!strcmp(k, "-b")) a.bs = v; else if (!strcmp(k, "-o")) a.rec = v;
        else if (!strcmp(k, "-wdt")) a.w = atoi(v); else if (!strcmp(k, "-hgt")) a.h = atoi(v);
        else if (!strcmp(k, "-fr")) a.fr = atof(v); else if (!strcmp(k, "-frms")) a.frms = atoi(v);
        else if (!strcmp(k, "-preset")) a.preset = v; else if (!strcmp(k, "-rc")) rc = atoi(v);
        else if (!strcmp(k, "-qp")) qp = atoi(v); else if (!strcmp(k, "-iper")) iper = atoi(v);
        else if (!strcmp(k, "-fixqp")) fixqp = atoi(v); else if (!strcmp(k, "-psnr")) a.psnr = atoi(v);
        else if (!strcmp(k, "-md5")) a.md5 = atoi(v); else if (!strcmp(k, "-gpus")) a.gpus = atoi(v);
        else if (!strcmp(k, "-streams")) a.streams = atoi(v); else if (!strcmp(k, "-sao")) sao = atoi(v);
        else if (!strcmp(k, "-bframes")) bframes = atoi(v);
        else if (!strcmp(k, "-me")) me = atoi(v);
        else if (!strcmp(k, "-crf")) crf = atof(v);
        else if (!strcmp(k, "-subme")) subme = atoi(v); else if (!strcmp(k, "-merange")) merange = atoi(v);
        else if (!strcmp(k, "-threads")) threads = atoi(v);      /* host worker count is -streams x -gpus here; 1 additionally prints the reference's `pure encoding time` line */
        else fprintf(stderr, "appencoder: warning: option %s %s is accepted for compatibility and ignored on the device path\n", k, v);
    }
    if (!a.in || a.w <= 0 || a.h <= 0) { usage(); return 2; }
    if (rc != 0 && rc != 3) { fprintf(stderr, "appencoder: -rc %d (ABR/CBR) is not implemented; use -rc 0 (fixed QP) or -rc 3 (CRF)\n", rc); return 2; }
    a.cfg.width = a.w; a.cfg.height = a.h;
    if (ks265_config_default_preset(&a.cfg, a.preset)) { fprintf(stderr, "appencoder: unknown preset %s\n", a.preset); return 2; }
    a.cfg.fps = a.fr; a.cfg.qp = qp; a.cfg.iper = iper < 1 ? 1 : iper; a.cfg.fixqp = fixqp;
    a.cfg.psnr = 1;         /* like the reference, the summary line always carries the PSNR (it prints it without -psnr too [probe]) */
    if (sao >= 0) a.cfg.sao = sao > 4 ? 4 : sao; if (subme >= 0) a.cfg.subpel = subme > 2 ? 2 : subme; if (merange > 0) a.cfg.me_range = merange;
    if (bframes >= 0) a.cfg.bframes = bframes > 7 ? 7 : bframes;
    if (me >= 0) a.cfg.me = me > 0;
    a.cfg.rc = rc; if (crf >= 0) a.cfg.crf = crf;
    if (a.gpus < 1) a.gpus = 1; if (a.streams < 1) a.streams = 1;
    if (a.gpus > 1 || a.streams > 16) setenv("KS_BLOCKING_SYNC", "1", 0);     /* many shard threads: they sleep instead of spinning while they wait (bench.py has the numbers) */

    job_t job; memset(&job, 0, sizeof(job));
    job.a = &a; job.fsz = (size_t)a.w * a.h * 3 / 2; pthread_mutex_init(&job.mu, NULL);
    FILE *fi = fopen(a.in, "rb");
    if (!fi) { perror(a.in); return 1; }
    struct stat sb; fstat(fileno(fi), &sb);
    long total_l = (long)(sb.st_size / (off_t)job.fsz);
    int total = total_l > 0x7fffffff ? 0x7fffffff : (int)total_l;
    if (a.frms > 0 && a.frms < total) total = a.frms;
    if (total < 1) { fprintf(stderr, "appencoder: input holds no complete %dx%d frame\n", a.w, a.h); return 1; }
    job.in_fd = fileno(fi); job.rec_fd = -1;
    FILE *fr = NULL;
    if (a.rec) { fr = fopen(a.rec, "wb"); if (!fr) { perror(a.rec); return 1; } job.rec_fd = fileno(fr); }
    job.nshards = (total + a.cfg.iper - 1) / a.cfg.iper;
    job.shards = (shard_t *)calloc(job.nshards, sizeof(shard_t));
    for (int s = 0; s < job.nshards; s++) { job.shards[s].first = s * a.cfg.iper; job.shards[s].n = total - s * a.cfg.iper < a.cfg.iper ? total - s * a.cfg.iper : a.cfg.iper; }
    /* echo of the effective configuration, `name: value` like the reference's start-up table */
    printf("ks265 B200 encoder (AppEncoder-compatible front end)\n");
    printf("preset: %-16s SourceWidth: %-11d SourceHeight: %-10d\nFrameRate: %-13.3f FrameToBeEncoded: %-6d IntraPeriod: %-11d\n", a.preset, a.w, a.h, a.fr, total, a.cfg.iper);
    printf("RCType: %-16d QP: %-20d Crf: %-19.2f\nFixedQp: %-15d BiPredFrames: %-10d IntMeSearchMethod: %-5d\n", rc, qp, a.cfg.crf, fixqp, a.cfg.bframes, a.cfg.me);
    printf("SubMeSearchMethod: %-5d SearchRange: %-11d SAO: %-19d\nActiveRefNum: 1          GopShards: %-13d Gpus x Streams: %d x %d\n", a.cfg.subpel, a.cfg.me_range, a.cfg.sao, job.nshards, a.gpus, a.streams);
    printf("encoder test start: %s\t res %dx%d\n", a.in, a.w, a.h);
    int nw = a.gpus * a.streams; if (nw > job.nshards) nw = job.nshards;
    pthread_t *th = (pthread_t *)calloc(nw, sizeof(pthread_t)); worker_arg *wa = (worker_arg *)calloc(nw, sizeof(worker_arg));
    double t0 = now_ms();
    for (int i = 0; i < nw; i++) { wa[i].job = &job; wa[i].device = i % a.gpus; pthread_create(&th[i], NULL, worker, &wa[i]); }
    for (int i = 0; i < nw; i++) pthread_join(th[i], NULL);
    double t1 = now_ms();
    FILE *fb = a.bs ? fopen(a.bs, "wb") : NULL;
    uint64_t bytes = 0, sse[3] = {0, 0, 0}, launches = 0; int rcode = 0;
    for (int s = 0; s < job.nshards; s++) {
        shard_t *sh = &job.shards[s];
        if (!sh->done || !sh->bs) { fprintf(stderr, "appencoder: GOP shard %d failed (error %d)\n", s, sh->err ? sh->err : -1); rcode = 1; continue; }
        if (sh->err) { fprintf(stderr, "appencoder: GOP shard %d: reconstruction not written (error %d)\n", s, sh->err); rcode = 1; }
        if (fb) fwrite(sh->bs, 1, (size_t)sh->bs_bytes, fb);
        bytes += (uint64_t)sh->bs_bytes; launches += sh->st.gpu_launches;
        for (int k = 0; k < 3; k++) sse[k] += sh->st.sse[k];
        free(sh->bs);
    }
    if (fb) fclose(fb);
    if (fr) fclose(fr);
    if (a.md5 && a.rec) {
        FILE *f = fopen(a.rec, "rb"); uint8_t *buf = (uint8_t *)malloc(job.fsz), d[3][16];
        for (int n = 0; f && n < total && fread(buf, 1, job.fsz, f) == job.fsz; n++) {
            ks_md5(buf, (size_t)a.w * a.h, d[0]); ks_md5(buf + (size_t)a.w * a.h, (size_t)a.w * a.h / 4, d[1]); ks_md5(buf + (size_t)a.w * a.h * 5 / 4, (size_t)a.w * a.h / 4, d[2]);
            printf("POC %d MD5 ", n);
            for (int k = 0; k < 3; k++) { for (int b = 0; b < 16; b++) printf("%02x", d[k][b]); printf(k < 2 ? "," : "\n"); }
        }
        if (f) fclose(f); free(buf);
    }
    double ms = t1 - t0, W = (double)((a.w + 15) & ~15), H = (double)((a.h + 15) & ~15), psnr[3];
    if (a.psnr >= 2) {       /* the reference's per-picture table (coding order inside each shard) */
        printf("poc\tslice\tbits\tpsnr\t\t\tqp\n");
        for (int s = 0; s < job.nshards; s++) {
            shard_t *sh = &job.shards[s];
            for (int i = 0; sh->pics && i < sh->n; i++) {
                const ks265_pic_stat *p = &sh->pics[i]; double q[3];
                for (int k = 0; k < 3; k++) { double npx = k ? W * H / 4 : W * H; q[k] = p->sse[k] ? 10.0 * log10(255.0 * 255.0 * npx / (double)p->sse[k]) : 99.99; }
                printf("%d\t%c\t%llu\t%.4f\t%.4f\t%.4f\t%d\n", sh->first + p->poc, p->slice_type == 2 ? 'I' : (p->slice_type == 1 ? 'P' : 'B'), (unsigned long long)p->bits, q[0], q[1], q[2], p->qp);
            }
            free(sh->pics);
        }
    }
    for (int k = 0; k < 3; k++) { double npx = (k ? W * H / 4 : W * H) * total; psnr[k] = sse[k] ? 10.0 * log10(255.0 * 255.0 * npx / (double)sse[k]) : 99.99; }
    printf("Total Frames: %d, test time: %.0fms, FPS: %.4f\n", total, ms, total * 1000.0 / ms);
    if (threads == 1 && job.enc_ms > 0) printf("Total Frames: %d, pure encoding time: %.0fms, %.4f fps\n", total, job.enc_ms, total * 1000.0 / job.enc_ms);
    printf("gpu kernel launches: %llu\n", (unsigned long long)launches);
    printf("bitrate, psnr: %.4f\t%.4f\t%.4f\t%.4f\n", bytes * 8.0 * a.fr / total / 1000.0, psnr[0], psnr[1], psnr[2]);
    if (!rcode) printf("H265 encoder passed!!!\n");
    fclose(fi);
    return rcode;
}
