/*
 * ks_gpu_api.cu -- implementation of the C-ABI in include/ks265_gpu.h: device/pinned memory, stream, the
 * per-picture launch sequence and the asynchronous syntax download.
 */
#include "ks265_gpu.h"
#include "ks265_enc.h"
#include "ks_launch.h"
#include "ks_kat.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "ks265gpu: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return KS_ECUDA; } } while (0)

static const int k_lambda_sad_q4[52] = {4,4,5,5,6,7,7,8,9,10,12,13,15,17,19,21,23,26,30,33,37,42,47,53,59,66,74,83,94,105,118,132,149,167,187,210,236,265,297,334,375,421,472,530,595,668,749,841,944,1060,1189,1335};
static const int k_lambda_sse_q4[52] = {1,1,1,2,2,3,3,4,5,7,9,11,14,17,22,27,34,43,54,69,86,109,137,173,218,274,345,435,548,691,870,1097,1382,1741,2193,2763,3482,4387,5527,6963,8773,11053,13926,17546,22107,27853,35092,44214,55706,70185,88427,111411};
static const uint8_t k_chroma_qp[58] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51};

#define KS_NSTAGE KS_NSTAGES   /* me, recon_inter, recon_intra, deblock, sao, pack, decide */
struct ks_syn_slot {
    ks_cell *d_cells; ks_ctu_syn *d_ctus; int16_t *d_pool; uint32_t *d_ncg; unsigned long long *d_sse, *d_mecost; ks_cell_b *d_cells_b;
    ks_cell_b *h_cells_b; int is_b;
    ks_cell *h_cells; ks_ctu_syn *h_ctus; int16_t *h_pool; uint32_t *h_ncg; unsigned long long *h_sse, *h_mecost;
    cudaEvent_t done; int pending;
    cudaEvent_t ev[KS_NSTAGE + 1]; int stage_of[KS_NSTAGE + 1]; int nev; size_t d2h_bytes;
};
struct ks_gpu_ctx {
    int device, dw, dh, W, H, cw, ch, ctw, cth;
    ks_gpu_cfg cfg;
    cudaStream_t st;
    cudaStream_t st_up;         /* source uploads run beside the previous picture's kernels of the same shard */
    cudaEvent_t *ev_up;         /* per source slot: upload finished (the picture's kernels wait for it) */
    cudaEvent_t *ev_src_read;   /* per source slot: the last picture that read it has finished (the next upload waits for it) */
    size_t fsz;                 /* bytes of one coded picture (W*H*3/2) */
    uint8_t **d_src, **d_rec;   /* slots */
    uint8_t **src_cur;          /* where each source slot's picture currently lives: d_src[slot], or the caller's device buffer (zero-copy) */
    uint8_t *d_pre;             /* pre-filter reconstruction / deblocked in place */
    CUtensorMap tm_pre[3];      /* TMA descriptors of d_pre's planes for the SAO tile staging */
    int tma_mask;               /* bit c: plane c qualifies (pitch multiple of 16 bytes) */
    uint8_t *d_pred;            /* inter prediction planes written by the motion search */
    uint8_t *d_pred1;           /* B pictures: list-1 prediction planes */
    ks_cell *d_cells1; int *d_cost0, *d_cost1;   /* d_cells1: list-1 field (B) / search field before the CU decision (P); d_cost0 doubles as its distortions */
    int16_t *d_lev;
    uint32_t *d_counts;
    int *d_nintra;               /* number of intra CUs the decision placed in the current P picture */
    void *d_imodes;              /* per-cell intra luma modes (mode search kernel -> dependent intra pass) */
    void *d_cands;               /* per-CTU candidate tables of the CU decision (stage E -> stage D) */
    uint32_t *d_sse_ctu;         /* per-CTU squared error partial sums (SAO kernel -> pack scan kernel) */
    int *d_sync;
    uint8_t *d_stage;           /* display-size I420 staging on device (upload + edge extension) */
    uint8_t *h_stage[2];        /* pinned, double-buffered */
    cudaEvent_t ev_stage[2]; int stage_idx;
    ks_syn_slot *syn;
    uint64_t launches;
    int profiling; double stage_ms[KS_NSTAGE]; uint64_t stage_n[KS_NSTAGE]; uint64_t d2h_total;
};

static KsPlanes planes_of(const ks_gpu_ctx *c, uint8_t *base) { KsPlanes p; p.p[0] = base; p.p[1] = base + (size_t)c->W * c->H; p.p[2] = p.p[1] + (size_t)c->W * c->H / 4; return p; }

/* edge-extend a display-size I420 picture (device) into a coded-size slot */
__global__ void ks_extend_kernel(const uint8_t *in, int dw, int dh, uint8_t *out, int W, int H)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    int plane = blockIdx.z, sh = plane ? 1 : 0, pw = W >> sh, ph = H >> sh, sw = dw >> sh, shh = dh >> sh;
    if (x >= pw || y >= ph) return;
    const uint8_t *ip = in + (plane == 0 ? 0 : (size_t)dw * dh + (plane == 2 ? (size_t)sw * shh : 0));
    uint8_t *op = out + (plane == 0 ? 0 : (size_t)W * H + (plane == 2 ? (size_t)pw * ph : 0));
    op[(size_t)y * pw + x] = ip[(size_t)min(y, shh - 1) * sw + min(x, sw - 1)];
}

/* 2-D u8 tensor map of one picture plane with a (bw x bh)-byte box, no swizzle, zero fill outside the picture.  The driver entry point is
 * resolved at run time (cudaGetDriverEntryPoint), so the library links against the runtime only. */
typedef CUresult (*ks_tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_plane_map(CUtensorMap *tm, void *ptr, int pw, int ph, int bw, int bh)
{
    static ks_tmap_encode_fn fn = NULL;
    if (!fn) {
        void *p = NULL; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return -1;
        fn = (ks_tmap_encode_fn)p;
    }
    if ((pw & 15) || ((uintptr_t)ptr & 15)) return 1;             /* row pitch / base must be multiples of 16 bytes: caller falls back */
    cuuint64_t dims[2] = {(cuuint64_t)pw, (cuuint64_t)ph}, strides[1] = {(cuuint64_t)pw};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, estr[2] = {1, 1};
    alignas(64) CUtensorMap tmp;                                  /* the encoder wants a 64-byte aligned destination; the context is calloc'ed */
    CUresult r = fn(&tmp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "ks265gpu: cuTensorMapEncodeTiled failed (%d) for a %dx%d plane\n", (int)r, pw, ph); return -1; }
    memcpy(tm, &tmp, sizeof(tmp));
    return 0;
}

extern "C" ks_gpu_ctx *ks_gpu_open(int device, int width, int height, const ks_gpu_cfg *cfg, int *err)
{
    int e = 0, ndev = 0;
    ks_gpu_ctx *c = NULL;
    if (err) *err = 0;
    if (width < 16 || height < 16 || (width & 1) || (height & 1) || !cfg) { e = KS_EINVAL; goto fail; }
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) { e = KS_ENODEV; fprintf(stderr, "ks265gpu: no CUDA device %d (the hot path has no CPU fallback)\n", device); goto fail; }
    if (cudaSetDevice(device) != cudaSuccess) { e = KS_ECUDA; goto fail; }
    c = (ks_gpu_ctx *)calloc(1, sizeof(*c));
    if (!c) { e = KS_ENOMEM; goto fail; }
    c->device = device; c->dw = width; c->dh = height; c->W = (width + 15) & ~15; c->H = (height + 15) & ~15;
    c->cw = c->W >> 4; c->ch = c->H >> 4; c->ctw = (c->W + 63) >> 6; c->cth = (c->H + 63) >> 6;
    c->cfg = *cfg;
    if (c->cfg.n_src_slots < 2) c->cfg.n_src_slots = 2;
    if (c->cfg.n_rec_slots < 2) c->cfg.n_rec_slots = 2;
    if (c->cfg.n_syn_slots < 2) c->cfg.n_syn_slots = 2;
    c->fsz = (size_t)c->W * c->H * 3 / 2;
    if (ks_init_device(device)) { e = KS_ECUDA; goto fail; }
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) { e = KS_ECUDA; goto fail; }
    if (cudaStreamCreateWithFlags(&c->st_up, cudaStreamNonBlocking) != cudaSuccess) { e = KS_ECUDA; goto fail; }
    c->ev_up = (cudaEvent_t *)calloc(c->cfg.n_src_slots, sizeof(cudaEvent_t));
    c->ev_src_read = (cudaEvent_t *)calloc(c->cfg.n_src_slots, sizeof(cudaEvent_t));
    for (int i = 0; i < c->cfg.n_src_slots; i++)
        if (cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_src_read[i], cudaEventDisableTiming) != cudaSuccess) { e = KS_ECUDA; goto fail; }
    c->d_src = (uint8_t **)calloc(c->cfg.n_src_slots, sizeof(uint8_t *));
    c->src_cur = (uint8_t **)calloc(c->cfg.n_src_slots, sizeof(uint8_t *));
    c->d_rec = (uint8_t **)calloc(c->cfg.n_rec_slots, sizeof(uint8_t *));
    c->syn = (ks_syn_slot *)calloc(c->cfg.n_syn_slots, sizeof(ks_syn_slot));
    {
        bool ok = true;
        for (int i = 0; i < c->cfg.n_src_slots; i++) { ok = ok && cudaMalloc(&c->d_src[i], c->fsz) == cudaSuccess; c->src_cur[i] = c->d_src[i]; }
        for (int i = 0; i < c->cfg.n_rec_slots; i++) ok = ok && cudaMalloc(&c->d_rec[i], c->fsz) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_pre, c->fsz) == cudaSuccess;
        if (ok) {   /* TMA descriptors for the SAO tile staging: Y box 96x66, chroma 64x34 (ks_loopfilter.cuh KS_SAO_PITCH_*) */
            uint8_t *pl[3] = {c->d_pre, c->d_pre + (size_t)c->W * c->H, c->d_pre + (size_t)c->W * c->H * 5 / 4};
            c->tma_mask = 0;
            for (int ci = 0; ci < 3 && ok; ci++) {
                int r = make_plane_map(&c->tm_pre[ci], pl[ci], c->W >> (ci ? 1 : 0), c->H >> (ci ? 1 : 0), ci ? 64 : 96, ci ? 34 : 66);
                if (r < 0) ok = false; else if (r == 0) c->tma_mask |= 1 << ci;
            }
            if (const char *e = getenv("KS_TMA_MASK")) c->tma_mask &= atoi(e);      /* debugging aid: 0 = stage every SAO tile with plain loads */
        }
        ok = ok && cudaMalloc(&c->d_pred, c->fsz) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_pred1, c->fsz) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_cells1, (size_t)c->cw * c->ch * sizeof(ks_cell)) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_cost0, (size_t)c->cw * c->ch * sizeof(int)) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_cost1, (size_t)c->cw * c->ch * sizeof(int)) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_lev, c->fsz * 2) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_counts, sizeof(uint32_t) * c->ctw * c->cth) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_nintra, sizeof(int)) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_imodes, ks_intra_workspace_bytes(c->cw * c->ch)) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_cands, ks_decide_workspace_bytes(c->ctw * c->cth)) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_sse_ctu, sizeof(uint32_t) * 3 * c->ctw * c->cth) == cudaSuccess;
        ok = ok && cudaMalloc(&c->d_sync, sizeof(int) * ((size_t)c->cw * c->ch + 1)) == cudaSuccess;
        size_t dsz = (size_t)width * height * 3 / 2;
        ok = ok && cudaMalloc(&c->d_stage, dsz) == cudaSuccess;
        for (int i = 0; i < 2; i++) { ok = ok && cudaHostAlloc(&c->h_stage[i], dsz, cudaHostAllocDefault) == cudaSuccess; ok = ok && cudaEventCreateWithFlags(&c->ev_stage[i], cudaEventDisableTiming) == cudaSuccess; }
        size_t ncell = (size_t)c->cw * c->ch, nctu = (size_t)c->ctw * c->cth, poolb = c->fsz * 2;
        for (int i = 0; i < c->cfg.n_syn_slots && ok; i++) {
            ks_syn_slot *s = &c->syn[i];
            ok = ok && cudaMalloc(&s->d_cells, ncell * sizeof(ks_cell)) == cudaSuccess;
            ok = ok && cudaMalloc(&s->d_cells_b, ncell * sizeof(ks_cell_b)) == cudaSuccess;
            ok = ok && cudaHostAlloc(&s->h_cells_b, ncell * sizeof(ks_cell_b), cudaHostAllocDefault) == cudaSuccess;
            ok = ok && cudaMalloc(&s->d_ctus, nctu * sizeof(ks_ctu_syn)) == cudaSuccess;
            ok = ok && cudaMalloc(&s->d_pool, poolb) == cudaSuccess;
            ok = ok && cudaMalloc(&s->d_ncg, sizeof(uint32_t)) == cudaSuccess;
            ok = ok && cudaMalloc(&s->d_sse, 3 * sizeof(unsigned long long)) == cudaSuccess;
            ok = ok && cudaMalloc(&s->d_mecost, sizeof(unsigned long long)) == cudaSuccess;
            ok = ok && cudaHostAlloc(&s->h_cells, ncell * sizeof(ks_cell), cudaHostAllocDefault) == cudaSuccess;
            ok = ok && cudaHostAlloc(&s->h_ctus, nctu * sizeof(ks_ctu_syn), cudaHostAllocDefault) == cudaSuccess;
            ok = ok && cudaHostAlloc(&s->h_pool, poolb, cudaHostAllocDefault) == cudaSuccess;
            ok = ok && cudaHostAlloc(&s->h_ncg, sizeof(uint32_t), cudaHostAllocDefault) == cudaSuccess;
            ok = ok && cudaHostAlloc(&s->h_sse, 3 * sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
            ok = ok && cudaHostAlloc(&s->h_mecost, sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
            /* KS_BLOCKING_SYNC=1: the shard's host thread sleeps instead of spinning while it waits for a picture (frees its core for the
             * entropy coders of other shards when every logical core is taken, e.g. 8 GPUs x 16 shards on a 128-thread host) */
            static const bool blocking = getenv("KS_BLOCKING_SYNC") && atoi(getenv("KS_BLOCKING_SYNC")) != 0;
            ok = ok && cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming | (blocking ? cudaEventBlockingSync : 0)) == cudaSuccess;
            for (int k = 0; k <= KS_NSTAGE; k++) ok = ok && cudaEventCreate(&s->ev[k]) == cudaSuccess;
            if (ok) cudaMemset(s->d_ctus, 0, nctu * sizeof(ks_ctu_syn));
        }
        if (!ok) { e = KS_ENOMEM; fprintf(stderr, "ks265gpu: allocation failed: %s\n", cudaGetErrorString(cudaGetLastError())); goto fail; }
    }
    return c;
fail:
    if (err) *err = e;
    if (c) ks_gpu_close(c);
    return NULL;
}

extern "C" void ks_gpu_close(ks_gpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->d_src) for (int i = 0; i < c->cfg.n_src_slots; i++) cudaFree(c->d_src[i]);
    if (c->d_rec) for (int i = 0; i < c->cfg.n_rec_slots; i++) cudaFree(c->d_rec[i]);
    cudaFree(c->d_pre); cudaFree(c->d_pred); cudaFree(c->d_pred1); cudaFree(c->d_cells1); cudaFree(c->d_cost0); cudaFree(c->d_cost1); cudaFree(c->d_lev); cudaFree(c->d_counts); cudaFree(c->d_sse_ctu); cudaFree(c->d_cands); cudaFree(c->d_imodes); cudaFree(c->d_nintra); cudaFree(c->d_sync); cudaFree(c->d_stage);
    for (int i = 0; i < 2; i++) { if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]); if (c->ev_stage[i]) cudaEventDestroy(c->ev_stage[i]); }
    if (c->syn) for (int i = 0; i < c->cfg.n_syn_slots; i++) {
        ks_syn_slot *s = &c->syn[i];
        cudaFree(s->d_cells); cudaFree(s->d_cells_b); if (s->h_cells_b) cudaFreeHost(s->h_cells_b); cudaFree(s->d_ctus); cudaFree(s->d_pool); cudaFree(s->d_ncg); cudaFree(s->d_sse); cudaFree(s->d_mecost); if (s->h_mecost) cudaFreeHost(s->h_mecost);
        if (s->h_cells) cudaFreeHost(s->h_cells); if (s->h_ctus) cudaFreeHost(s->h_ctus); if (s->h_pool) cudaFreeHost(s->h_pool);
        if (s->h_ncg) cudaFreeHost(s->h_ncg); if (s->h_sse) cudaFreeHost(s->h_sse);
        if (s->done) cudaEventDestroy(s->done);
        for (int k = 0; k <= KS_NSTAGE; k++) if (s->ev[k]) cudaEventDestroy(s->ev[k]);
    }
    if (c->st) cudaStreamDestroy(c->st);
    if (c->st_up) cudaStreamDestroy(c->st_up);
    for (int i = 0; i < c->cfg.n_src_slots; i++) { if (c->ev_up && c->ev_up[i]) cudaEventDestroy(c->ev_up[i]); if (c->ev_src_read && c->ev_src_read[i]) cudaEventDestroy(c->ev_src_read[i]); }
    free(c->ev_up); free(c->ev_src_read);
    free(c->d_src); free(c->src_cur); free(c->d_rec); free(c->syn); free(c);
}

extern "C" void *ks265_alloc_host(size_t bytes) { void *p = NULL; return cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess ? p : NULL; }
extern "C" void ks265_free_host(void *p) { if (p) cudaFreeHost(p); }
extern "C" int ks_gpu_coded_size(const ks_gpu_ctx *c, int *w, int *h) { if (!c) return KS_EINVAL; if (w) *w = c->W; if (h) *h = c->H; return 0; }
extern "C" uint64_t ks_gpu_launch_count(const ks_gpu_ctx *c) { return c ? c->launches : 0; }
extern "C" void *ks_gpu_stream(ks_gpu_ctx *c) { return c ? (void *)c->st : NULL; }
extern "C" int ks_gpu_set_profiling(ks_gpu_ctx *c, int on)
{
    if (!c) return KS_EINVAL;
    c->profiling = on != 0;
    for (int k = 0; k < KS_NSTAGE; k++) { c->stage_ms[k] = 0; c->stage_n[k] = 0; }
    return 0;
}
extern "C" int ks_gpu_get_stage_times(const ks_gpu_ctx *c, double ms[KS_NSTAGES], uint64_t n[KS_NSTAGES])
{
    if (!c) return KS_EINVAL;
    for (int k = 0; k < KS_NSTAGE; k++) { ms[k] = c->stage_ms[k]; n[k] = c->stage_n[k]; }
    return 0;
}
extern "C" uint64_t ks_gpu_d2h_bytes(const ks_gpu_ctx *c) { return c ? c->d2h_total : 0; }
extern "C" size_t ks_gpu_abi_sizeof(int which)
{
    static const size_t sz[8] = {sizeof(ks_gpu_cfg), sizeof(ks_pic_params), sizeof(ks_pic_out), sizeof(ks_cell), sizeof(ks_cell_b), sizeof(ks_ctu_syn),
                                 sizeof(ks265_config), sizeof(ks265_gop_stats)};
    return which >= 0 && which < 8 ? sz[which] : 0;
}

static int extend_into_slot(ks_gpu_ctx *c, const uint8_t *dev_i420, int slot)
{
    dim3 grid((c->W + 255) / 256, c->H, 3);
    ks_extend_kernel<<<grid, 256, 0, c->st_up>>>(dev_i420, c->dw, c->dh, c->d_src[slot], c->W, c->H);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ks_gpu_upload_frame(ks_gpu_ctx *c, int slot, const uint8_t *y, const uint8_t *u, const uint8_t *v, int sy, int suv)
{
    if (!c || slot < 0 || slot >= c->cfg.n_src_slots || !y || !u || !v) return KS_EINVAL;
    CK(cudaSetDevice(c->device));
    const size_t dsz = (size_t)c->dw * c->dh * 3 / 2;
    const bool tight = sy == c->dw && suv == c->dw / 2 && u == y + (size_t)c->dw * c->dh && v == u + (size_t)c->dw * c->dh / 4;
    const bool same = c->dw == c->W && c->dh == c->H;
    uint8_t *dst = same ? c->d_src[slot] : c->d_stage;          /* no padding needed: land directly in the slot */
    CK(cudaStreamWaitEvent(c->st_up, c->ev_src_read[slot], 0)); /* whoever last read this slot is done (no-op if nobody did) */
    cudaPointerAttributes at;
    if (tight && cudaPointerGetAttributes(&at, y) == cudaSuccess && at.type == cudaMemoryTypeHost) {
        /* caller's buffer is page-locked: DMA straight out of it (caller keeps it alive until the picture is finished,
         * the same ownership rule as QY265Picture, qy265enc.h:153-157) */
        CK(cudaMemcpyAsync(dst, y, dsz, cudaMemcpyHostToDevice, c->st_up));
    } else {
        (void)cudaGetLastError();
        const int si = c->stage_idx; c->stage_idx ^= 1;
        CK(cudaEventSynchronize(c->ev_stage[si]));  /* this pinned staging buffer's previous H2D must be done */
        uint8_t *d = c->h_stage[si];
        for (int r = 0; r < c->dh; r++) memcpy(d + (size_t)r * c->dw, y + (size_t)r * sy, c->dw);
        d += (size_t)c->dw * c->dh;
        for (int r = 0; r < c->dh / 2; r++) memcpy(d + (size_t)r * (c->dw / 2), u + (size_t)r * suv, c->dw / 2);
        d += (size_t)c->dw * c->dh / 4;
        for (int r = 0; r < c->dh / 2; r++) memcpy(d + (size_t)r * (c->dw / 2), v + (size_t)r * suv, c->dw / 2);
        CK(cudaMemcpyAsync(dst, c->h_stage[si], dsz, cudaMemcpyHostToDevice, c->st_up));
        CK(cudaEventRecord(c->ev_stage[si], c->st_up));
    }
    c->src_cur[slot] = c->d_src[slot];
    int r = same ? 0 : extend_into_slot(c, c->d_stage, slot);
    if (r) return r;
    CK(cudaEventRecord(c->ev_up[slot], c->st_up));
    return 0;
}
extern "C" uint8_t *ks_gpu_stage_acquire(ks_gpu_ctx *c)
{
    if (!c || cudaSetDevice(c->device) != cudaSuccess) return NULL;
    if (cudaEventSynchronize(c->ev_stage[c->stage_idx]) != cudaSuccess) return NULL;      /* this buffer's previous H2D must be done */
    return c->h_stage[c->stage_idx];
}
extern "C" int ks_gpu_upload_staged(ks_gpu_ctx *c, int slot)
{
    if (!c || slot < 0 || slot >= c->cfg.n_src_slots) return KS_EINVAL;
    CK(cudaSetDevice(c->device));
    const size_t dsz = (size_t)c->dw * c->dh * 3 / 2;
    const bool same = c->dw == c->W && c->dh == c->H;
    const int si = c->stage_idx; c->stage_idx ^= 1;
    CK(cudaStreamWaitEvent(c->st_up, c->ev_src_read[slot], 0));
    CK(cudaMemcpyAsync(same ? c->d_src[slot] : c->d_stage, c->h_stage[si], dsz, cudaMemcpyHostToDevice, c->st_up));
    CK(cudaEventRecord(c->ev_stage[si], c->st_up));
    c->src_cur[slot] = c->d_src[slot];
    int r = same ? 0 : extend_into_slot(c, c->d_stage, slot);
    if (r) return r;
    CK(cudaEventRecord(c->ev_up[slot], c->st_up));
    return 0;
}
extern "C" int ks_gpu_upload_frame_device(ks_gpu_ctx *c, int slot, const void *dev_i420)
{
    if (!c || slot < 0 || slot >= c->cfg.n_src_slots || !dev_i420) return KS_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamWaitEvent(c->st_up, c->ev_src_read[slot], 0));
    /* coded size == display size: the kernels read the caller's buffer in place (no copy; the caller keeps it unchanged until the picture that
     * uses the slot has finished); otherwise the edge-extension kernel writes the padded picture into the slot */
    if (c->dw == c->W && c->dh == c->H && !((uintptr_t)dev_i420 & 15)) c->src_cur[slot] = (uint8_t *)dev_i420;
    else { c->src_cur[slot] = c->d_src[slot]; int r = extend_into_slot(c, (const uint8_t *)dev_i420, slot); if (r) return r; }
    CK(cudaEventRecord(c->ev_up[slot], c->st_up));
    return 0;
}

static int fill_params(const ks_gpu_ctx *c, const ks_pic_params *p, KsPicParams *pp)
{
    if (p->qp < 0 || p->qp > 51) return KS_EINVAL;
    if (p->src_slot < 0 || p->src_slot >= c->cfg.n_src_slots || p->out_slot < 0 || p->out_slot >= c->cfg.n_rec_slots) return KS_EINVAL;
    if (p->syn_slot < 0 || p->syn_slot >= c->cfg.n_syn_slots || p->prev_syn_slot >= c->cfg.n_syn_slots) return KS_EINVAL;
    if (p->slice_type != KS_SLICE_I && (p->ref_slot < 0 || p->ref_slot >= c->cfg.n_rec_slots || p->ref_slot == p->out_slot)) return KS_EINVAL;
    if (p->slice_type == KS_SLICE_B && (p->ref1_slot < 0 || p->ref1_slot >= c->cfg.n_rec_slots || p->ref1_slot == p->out_slot || p->ref1_slot == p->ref_slot
                                         || p->dist_l0 <= 0 || p->dist_anchor <= p->dist_l0)) return KS_EINVAL;
    memset(pp, 0, sizeof(*pp));
    pp->W = c->W; pp->H = c->H; pp->dW = c->dw; pp->dH = c->dh; pp->cw = c->cw; pp->ch = c->ch; pp->ctw = c->ctw; pp->cth = c->cth;
    pp->slice_type = p->slice_type; pp->qp = p->qp; pp->qpc = k_chroma_qp[p->qp];
    pp->lambda_sad_q4 = k_lambda_sad_q4[p->qp]; pp->lambda_sse_q4 = k_lambda_sse_q4[p->qp];
    if (p->lambda_qp_delta < 0) return KS_EINVAL;
    { const int lqp = p->qp + p->lambda_qp_delta > 51 ? 51 : p->qp + p->lambda_qp_delta;
      pp->lambda_dec_q4 = k_lambda_sad_q4[lqp]; pp->rdz_lambda_q4 = p->slice_type == KS_SLICE_I ? 0 : k_lambda_sse_q4[lqp]; }
    pp->me_range = c->cfg.me_range; pp->me_iters = c->cfg.me_iters; pp->subpel = c->cfg.subpel; pp->satd = c->cfg.satd; pp->me_method = c->cfg.me_method;
    pp->sign_hiding = c->cfg.sign_hiding; pp->sao = c->cfg.sao; pp->strong_intra = c->cfg.strong_intra;
    pp->beta_offset_div2 = p->beta_offset_div2; pp->tc_offset_div2 = p->tc_offset_div2;
    return 0;
}

extern "C" int ks_gpu_encode_picture_submit(ks_gpu_ctx *c, const ks_pic_params *p)
{
    KsPicParams pp;
    if (!c || !p) return KS_EINVAL;
    int r = fill_params(c, p, &pp);
    if (r) return r;
    CK(cudaSetDevice(c->device));
    ks_syn_slot *s = &c->syn[p->syn_slot];
    if (s->pending) return KS_EINVAL;
    KsPlanes src = planes_of(c, c->src_cur[p->src_slot]), pre = planes_of(c, c->d_pre), out = planes_of(c, c->d_rec[p->out_slot]);
    KsLevels lv; lv.p[0] = c->d_lev; lv.p[1] = c->d_lev + (size_t)c->W * c->H; lv.p[2] = lv.p[1] + (size_t)c->W * c->H / 4;
    CK(cudaStreamWaitEvent(c->st, c->ev_up[p->src_slot], 0));  /* the source picture's upload (separate stream) */
    s->nev = 0;
#define MARK(stage) do { if (c->profiling) { cudaEventRecord(s->ev[s->nev], c->st); s->stage_of[s->nev++] = (stage); } } while (0)
    if (p->slice_type == KS_SLICE_I) {
        MARK(2);
        ks_launch_recon_intra(pp, src, pre, lv, s->d_cells, c->d_sync, NULL, c->d_imodes, c->st); c->launches += KS_LAUNCHES_INTRA;
    } else {
        KsPlanes ref = planes_of(c, c->d_rec[p->ref_slot]);
        const ks_cell *prev = p->prev_syn_slot >= 0 ? c->syn[p->prev_syn_slot].d_cells : NULL;
        MARK(0);
        KsPlanes pred = planes_of(c, c->d_pred);
        const ks_cell_b *cb = NULL;
        if (p->slice_type == KS_SLICE_B) {
            /* list 0 and list 1 searches seeded by the later anchor's vectors scaled to each list, then the bi-prediction decision */
            KsPlanes ref1 = planes_of(c, c->d_rec[p->ref1_slot]), pred1 = planes_of(c, c->d_pred1);
            KsPicParams p0 = pp, p1 = pp;
            p0.pred_num = p->dist_l0; p0.pred_den = p->dist_anchor; p1.pred_num = p->dist_l0 - p->dist_anchor; p1.pred_den = p->dist_anchor;
            ks_launch_me(p0, src.p[0], ref, prev, s->d_cells, pred, c->d_cost0, NULL, NULL, c->st);
            ks_launch_me(p1, src.p[0], ref1, prev, c->d_cells1, pred1, c->d_cost1, NULL, NULL, c->st);
            ks_launch_bidir(pp, src.p[0], ref, ref1, prev, p0.pred_num, p1.pred_num, p->dist_anchor, c->d_cells1, c->d_cost0, c->d_cost1, pred1, s->d_cells, s->d_cells_b, pred, c->st);
            c->launches += 2 * KS_LAUNCHES_ME + 1;
            cb = s->d_cells_b;
        } else {
            const bool mc = p->want_me_cost != 0;
            if (mc) CK(cudaMemsetAsync(s->d_mecost, 0, sizeof(unsigned long long), c->st));
            /* search field -> d_cells1 (+ distortions), then the CU quadtree / merge decision writes the final cells */
            ks_launch_me(pp, src.p[0], ref, prev, c->d_cells1, pred, NULL, c->d_cost0, mc ? s->d_mecost : NULL, c->st); c->launches += KS_LAUNCHES_ME;
            MARK(6);
            ks_launch_decide(pp, src.p[0], ref, c->d_cells1, c->d_cost0, c->d_cands, s->d_cells, pred, c->d_nintra, c->st); c->launches += KS_LAUNCHES_DECIDE;
        }
        MARK(1);
        ks_launch_recon_inter(pp, src, pred, pre, lv, s->d_cells, cb, c->st); c->launches += KS_LAUNCHES_RECON;
        if (p->slice_type == KS_SLICE_P) {      /* the intra CUs the decision placed: their neighbours' inter reconstruction now exists */
            MARK(7);
            ks_launch_recon_intra(pp, src, pre, lv, s->d_cells, c->d_sync, c->d_nintra, c->d_imodes, c->st); c->launches += KS_LAUNCHES_INTRA;
        }
    }
    s->is_b = p->slice_type == KS_SLICE_B;
    MARK(3);
    ks_launch_deblock(pp, pre, s->d_cells, s->is_b ? s->d_cells_b : NULL, c->st); c->launches += KS_LAUNCHES_DEBLOCK;
    MARK(4);
    ks_launch_sao(pp, src, pre, out, s->d_ctus, p->want_sse ? c->d_sse_ctu : NULL, c->tm_pre, c->tma_mask, c->st); c->launches += KS_LAUNCHES_SAO - 1;
    CK(cudaEventRecord(c->ev_src_read[p->src_slot], c->st));   /* SAO was the last stage to read the source picture */
    MARK(5);
    ks_launch_pack(pp, lv, s->d_ctus, s->d_pool, s->d_ncg, c->d_counts, p->want_sse ? c->d_sse_ctu : NULL, s->d_sse, c->st); c->launches += KS_LAUNCHES_PACK;
    MARK(-1);
#undef MARK
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(s->h_cells, s->d_cells, (size_t)c->cw * c->ch * sizeof(ks_cell), cudaMemcpyDeviceToHost, c->st));
    if (s->is_b) CK(cudaMemcpyAsync(s->h_cells_b, s->d_cells_b, (size_t)c->cw * c->ch * sizeof(ks_cell_b), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(s->h_ctus, s->d_ctus, (size_t)c->ctw * c->cth * sizeof(ks_ctu_syn), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(s->h_ncg, s->d_ncg, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
    if (p->want_sse) CK(cudaMemcpyAsync(s->h_sse, s->d_sse, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
    else { s->h_sse[0] = s->h_sse[1] = s->h_sse[2] = 0; }
    if (p->slice_type == KS_SLICE_P && p->want_me_cost) CK(cudaMemcpyAsync(s->h_mecost, s->d_mecost, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
    else *s->h_mecost = 0;
    CK(cudaEventRecord(s->done, c->st));
    s->pending = 1;
    return 0;
}

extern "C" int ks_gpu_encode_picture_finish(ks_gpu_ctx *c, int syn_slot, ks_pic_out *out)
{
    if (!c || !out || syn_slot < 0 || syn_slot >= c->cfg.n_syn_slots) return KS_EINVAL;
    ks_syn_slot *s = &c->syn[syn_slot];
    if (!s->pending) return KS_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(s->done));
    uint32_t n = *s->h_ncg;
    if ((size_t)n * 32 > c->fsz * 2) return KS_ECUDA;
    if (n) {
        /* the pool of this slot is not rewritten until the slot is reused, so a second small copy on the same stream is safe */
        CK(cudaMemcpyAsync(s->h_pool, s->d_pool, (size_t)n * 32, cudaMemcpyDeviceToHost, c->st));
        CK(cudaEventRecord(s->done, c->st));
        CK(cudaEventSynchronize(s->done));
    }
    if (c->profiling)
        for (int k = 0; k + 1 < s->nev; k++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s->ev[k], s->ev[k + 1]) == cudaSuccess) { c->stage_ms[s->stage_of[k]] += ms; c->stage_n[s->stage_of[k]]++; }
        }
    c->d2h_total += (size_t)c->cw * c->ch * sizeof(ks_cell) + (size_t)c->ctw * c->cth * sizeof(ks_ctu_syn) + 4 + (size_t)n * 32;
    s->pending = 0;
    out->cells = s->h_cells; out->ctus = s->h_ctus; out->levels = s->h_pool; out->n_cg = n;
    out->sse[0] = s->h_sse[0]; out->sse[1] = s->h_sse[1]; out->sse[2] = s->h_sse[2];
    out->me_cost = *s->h_mecost;
    out->cells_b = s->is_b ? s->h_cells_b : NULL;
    return 0;
}
extern "C" int ks_gpu_abort(ks_gpu_ctx *c)
{
    if (!c) return KS_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st_up));
    CK(cudaStreamSynchronize(c->st));
    for (int i = 0; i < c->cfg.n_syn_slots; i++) c->syn[i].pending = 0;
    return 0;
}
extern "C" int ks_gpu_encode_picture(ks_gpu_ctx *c, const ks_pic_params *p, ks_pic_out *out)
{
    int r = ks_gpu_encode_picture_submit(c, p);
    if (r) return r;
    return ks_gpu_encode_picture_finish(c, p->syn_slot, out);
}

extern "C" int ks_gpu_fetch_recon(ks_gpu_ctx *c, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int sy, int suv)
{
    if (!c || slot < 0 || slot >= c->cfg.n_rec_slots) return KS_EINVAL;
    CK(cudaSetDevice(c->device));
    KsPlanes p = planes_of(c, c->d_rec[slot]);
    CK(cudaMemcpy2DAsync(y, sy, p.p[0], c->W, c->dw, c->dh, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpy2DAsync(u, suv, p.p[1], c->W / 2, c->dw / 2, c->dh / 2, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpy2DAsync(v, suv, p.p[2], c->W / 2, c->dw / 2, c->dh / 2, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int ks_gpu_debug_fetch(ks_gpu_ctx *c, int what, int slot, void *dst, size_t bytes)
{
    if (!c || !dst) return KS_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st));
    const void *srcp; size_t need;
    if (what == KS_DBG_PRE_RECON) { srcp = c->d_pre; need = c->fsz; }
    else if (what == KS_DBG_LEVELS) { srcp = c->d_lev; need = c->fsz * 2; }
    else if (what == KS_DBG_SRC) { if (slot < 0 || slot >= c->cfg.n_src_slots) return KS_EINVAL; srcp = c->src_cur[slot]; need = c->fsz; }
    else return KS_EINVAL;
    if (bytes < need) return KS_EINVAL;
    CK(cudaMemcpy(dst, srcp, need, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int ks_gpu_debug_me(ks_gpu_ctx *c, const ks_pic_params *p, ks_cell *cells_out)
{
    KsPicParams pp;
    if (!c || !p || !cells_out || p->slice_type != KS_SLICE_P) return KS_EINVAL;
    int r = fill_params(c, p, &pp);
    if (r) return r;
    CK(cudaSetDevice(c->device));
    ks_syn_slot *s = &c->syn[p->syn_slot];
    KsPlanes src = planes_of(c, c->src_cur[p->src_slot]), ref = planes_of(c, c->d_rec[p->ref_slot]);
    const ks_cell *prev = p->prev_syn_slot >= 0 ? c->syn[p->prev_syn_slot].d_cells : NULL;
    KsPlanes nopred; nopred.p[0] = nopred.p[1] = nopred.p[2] = NULL;
    CK(cudaStreamWaitEvent(c->st, c->ev_up[p->src_slot], 0));
    ks_launch_me(pp, src.p[0], ref, prev, s->d_cells, nopred, NULL, NULL, NULL, c->st); c->launches += KS_LAUNCHES_ME;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cells_out, s->d_cells, (size_t)c->cw * c->ch * sizeof(ks_cell), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

/* ------------------------------------------------------------------ KAT entry points (host buffers) - */
template <typename T> struct dev_buf {
    T *p; dev_buf(size_t n) : p(NULL) { cudaMalloc(&p, n * sizeof(T)); } ~dev_buf() { cudaFree(p); }
};
extern "C" int ks_gpu_kat_sad16(const uint8_t *a, const uint8_t *b, long sa, long sb, uint32_t *out)
{
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return KS_ENODEV;
    if (ks_init_device(-1)) return KS_ECUDA;
    uint8_t ha[256], hb[256];
    for (int y = 0; y < 16; y++) { memcpy(ha + 16 * y, a + y * sa, 16); memcpy(hb + 16 * y, b + y * sb, 16); }
    dev_buf<uint8_t> da(256), db(256); dev_buf<uint32_t> dout(1);
    if (!da.p || !db.p || !dout.p) return KS_ENOMEM;
    CK(cudaMemcpy(da.p, ha, 256, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db.p, hb, 256, cudaMemcpyHostToDevice));
    if (ks_kat_sad16_dev(da.p, db.p, dout.p)) return KS_ECUDA;
    CK(cudaMemcpy(out, dout.p, 4, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int ks_gpu_kat_satd16(const uint8_t *a, const uint8_t *b, long sa, long sb, uint32_t *out)
{
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return KS_ENODEV;
    uint8_t ha[256], hb[256];
    for (int y = 0; y < 16; y++) { memcpy(ha + 16 * y, a + y * sa, 16); memcpy(hb + 16 * y, b + y * sb, 16); }
    dev_buf<uint8_t> da(256), db(256); dev_buf<uint32_t> dout(1);
    if (!da.p || !db.p || !dout.p) return KS_ENOMEM;
    CK(cudaMemcpy(da.p, ha, 256, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db.p, hb, 256, cudaMemcpyHostToDevice));
    if (ks_kat_satd16_dev(da.p, db.p, dout.p)) return KS_ECUDA;
    CK(cudaMemcpy(out, dout.p, 4, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int ks_gpu_kat_interp_luma16(const uint8_t *plane, int w, int h, int x, int y, int mvx, int mvy, uint8_t *dst)
{
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return KS_ENODEV;
    if ((w & 3) || w < 16 || h < 16) return KS_EINVAL;
    if (ks_init_device(-1)) return KS_ECUDA;
    dev_buf<uint8_t> dp((size_t)w * h), dd(256);
    if (!dp.p || !dd.p) return KS_ENOMEM;
    CK(cudaMemcpy(dp.p, plane, (size_t)w * h, cudaMemcpyHostToDevice));
    if (ks_kat_interp_dev(dp.p, w, h, x, y, mvx, mvy, dd.p)) return KS_ECUDA;
    CK(cudaMemcpy(dst, dd.p, 256, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int ks_gpu_kat_tb(int log2n, const uint8_t *src, const uint8_t *pred, int qp, int intra_slice, int sign_hiding,
                             int16_t *levels, uint8_t *recon, int *cbf)
{
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return KS_ENODEV;
    if (log2n < 2 || log2n > 5 || qp < 0 || qp > 51) return KS_EINVAL;
    if (ks_init_device(-1)) return KS_ECUDA;
    size_t nn = (size_t)1 << (2 * log2n);
    dev_buf<uint8_t> ds(nn), dp(nn), dr(nn); dev_buf<int16_t> dl(nn); dev_buf<int> dc(1);
    if (!ds.p || !dp.p || !dr.p || !dl.p || !dc.p) return KS_ENOMEM;
    CK(cudaMemcpy(ds.p, src, nn, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp.p, pred, nn, cudaMemcpyHostToDevice));
    if (ks_kat_tb_dev(log2n, ds.p, dp.p, qp, intra_slice, sign_hiding, dl.p, dr.p, dc.p)) return KS_ECUDA;
    CK(cudaMemcpy(levels, dl.p, nn * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(recon, dr.p, nn, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cbf, dc.p, 4, cudaMemcpyDeviceToHost));
    return 0;
}
