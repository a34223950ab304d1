/*
 * ks_recon.cuh -- prediction + residual path of the ks265 B200 hot path (SURVEY.md 8a rows a7-a14 + intra, f1).
 *
 *  ks_tb_code<N>   one warp codes 32/N transform blocks at once: residual -> forward DCT (reference
 *                  H265_2dDct*_c E@0x4b7600.., stage shifts 2*log2N-2 and 7) -> quantiser (H265QuantBlock_c
 *                  E@0x4a2580) -> sign-data hiding (signBitHidingHDQ E@0x4a29c0) -> dequantiser
 *                  (H265DeQuantBlock_c E@0x439540) -> inverse DCT + prediction add (H265_2dIDct*_c E@0x4417f0..).
 *                  int16/int32 butterfly work on the CUDA cores: lane = one row/column, the transform matrix is
 *                  fully unrolled partial butterflies with IMMEDIATE coefficients (ks_dct_gen.cuh; constant-bank
 *                  operands thrashed the immediate-constant cache),
 *                  transposes go through padded shared memory.  No tensor cores (north star).
 *  ks_recon_inter_kernel   one CTA per 64x64 CTU: CU-size decision, then the transform tasks (the reference's `reconstruct`
 *                  E@0x47d600 driver).  The prediction comes from the motion-search kernel, which already holds the winner's
 *                  interpolated block (reference: getReusSubMePred E@0x486770).
 *  ks_recon_intra_kernel   I pictures: CTU rows run as a wavefront (reference: WPP, CCtuEncWpp::waitForTopRightCtu
 *                  E@0x468d40) inside ONE launch, rows handed out by a ticket counter, progress flags in HBM.
 * Bit-exact mirror of oracle/ora_frame.c.
 */
#pragma once
#include "ks_me.cuh"

/* ------------------------------------------------------------------ transform-block coder -------- */
struct KsTbScratch {
    int16_t S[32 * 40];        /* transpose buffer, rows padded by 8 (conflict-free 128-bit reads) */
    int16_t C[32 * 32];        /* coefficients  [v][u] (per TB group: G blocks of N x N, consecutive) */
    int16_t L[32 * 32];        /* levels */
    int16_t D[32 * 32];        /* deltaU */
};

template <int N> struct KsLog2 { static const int v = N == 32 ? 5 : (N == 16 ? 4 : (N == 8 ? 3 : 2)); };
/* per-block statistics of ks_tb_code (the intra 16x16 vs 8x8 decision): SSE(src, rec) and the estimated level bits as coded */
struct KsTbStat { int d1, bits; };

#ifdef KS_INTRA_TIMING
__device__ long long g_tbt[8];
#define KS_TBT(k) do { if (N == 16 && lane == 0) { long long t_ = clock64(); atomicAdd((unsigned long long *)&g_tbt[k], (unsigned long long)(t_ - tb_t0)); tb_t0 = t_; } } while (0)
#else
#define KS_TBT(k) do {} while (0)
#endif
/* transform passes: generated partial butterflies with immediate coefficients (tools/gen_dct.py) */
#include "ks_dct_gen.cuh"

/* sign-data hiding for one coefficient group (16 scan positions starting at scan index sp) of a TB whose
 * coefficient/level/deltaU arrays are N x N row-major.  Mirrors ora_sign_hide's per-CG body.  The group's 16 positions, levels and deltas
 * are gathered into registers with independent (pipelined) shared loads and the decision runs on bit masks: the dependent pass of the intra
 * kernel waits on this code, so its latency -- not its instruction count -- is what matters.  One copy per kernel (not inlined). */
__device__ __noinline__ void ks_sbh_cg(const int16_t *C, int16_t *L, const int16_t *D, const uint16_t *scan, int sp, bool is_last_cg, int n)
{
    int pos[16], lv[16];
    unsigned nzm = 0, negm = 0, onem = 0, par = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { const int s = scan[sp + i]; pos[i] = (s >> 8) * n + (s & 255); }
#pragma unroll
    for (int i = 0; i < 16; i++) lv[i] = L[pos[i]];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        nzm |= (unsigned)(lv[i] != 0) << i; negm |= (unsigned)(lv[i] < 0) << i; onem |= (unsigned)(lv[i] == 1 || lv[i] == -1) << i; par ^= (unsigned)lv[i];
    }
    if (!nzm) return;
    const int first = __ffs(nzm) - 1, last = 31 - __clz(nzm);
    if (last - first < 4) return;
    const unsigned signbit = (negm >> first) & 1u;
    if (signbit == (par & 1u)) return;                     /* parity of the sum == parity of the number of odd levels */
    const int start = is_last_cg ? last : 15;
    int min_cost = 0x7fffffff, min_pos = 0, min_lv = 0, final_change = 0; bool min_neg = false;
#pragma unroll
    for (int i = 15; i >= 0; i--) {
        const int du = D[pos[i]];
        const bool nz = (nzm >> i) & 1u;
        /* the coefficient's sign: the level's when it has one, the coefficient's own otherwise */
        const bool neg = nz ? ((negm >> i) & 1u) != 0 : C[pos[i]] < 0;
        int cost = -du, change = 1;
        if (nz) {
            if (du <= 0) { change = -1; cost = (i == first && ((onem >> i) & 1u)) ? 0x7fffffff : du; }
        } else if (i < first && (unsigned)neg != signbit) cost = 0x7fffffff;
        if (i <= start && cost < min_cost) { min_cost = cost; final_change = change; min_pos = pos[i]; min_lv = lv[i]; min_neg = neg; }
    }
    if (min_lv == 32767 || min_lv == -32768) final_change = -1;
    L[min_pos] = (int16_t)(min_neg ? min_lv - final_change : min_lv + final_change);
}

/*
 * One warp, G = 32/N transform blocks: lane -> (g = lane / N, r = lane % N).
 *   src_row : pointer (global or shared) to row r of the lane's source block (N bytes, 4-byte aligned, valid lanes only)
 *   pred_row: shared  pointer to row r of the lane's prediction block (N bytes)
 *   rec_row : global pointer to row r of the lane's reconstruction
 *   lev_row : global pointer to row r of the lane's level block (dense int16 plane)
 *   t0      : shared 16x16 int table, t0[j][x] = M32[2j+1][x] (level-0 odd part of the 32-point butterflies)
 *   rdz_lambda_q4 > 0: RD zero-out (our restatement of the reference's zero-block decisions in tuDecision E@0x47e2f0 / skipFastDecision
 *             E@0x47f720): the block's levels are dropped when SSE(src,pred) <= SSE(src,rec) + lambda * bits, bits estimated as
 *             3 per level + 2 per magnitude doubling + 4 per coded 4x4 group + the anti-diagonal of the outermost level (== ora level_bits_est)
 * returns, per lane, whether the lane's block has any non-zero level (cbf).
 */
template <int N>
__device__ __noinline__ bool ks_tb_code(KsTbScratch *sc, const uint16_t *scan, const int *t0, bool valid,
                                        const uint8_t *src_row, const uint8_t *pred_row,
                                        uint8_t *__restrict__ rec_row, int16_t *__restrict__ lev_row,
                                        int qp, int intra_slice, int sign_hiding, int lane, int rdz_lambda_q4 = 0, KsTbStat *stat_out = nullptr)
{
    constexpr int LOG2 = KsLog2<N>::v, G = 32 / N, SP = N + 8;
    const int g = lane / N, r = lane % N;
    const unsigned gmask = (N == 32) ? 0xffffffffu : (((1u << N) - 1u) << (g * N));
    int16_t *S = sc->S + g * N * SP, *C = sc->C + g * N * N, *L = sc->L + g * N * N, *D = sc->D + g * N * N;
    int res[N];
    uint8_t pred[N];
    uint32_t srcw[N / 4];
#ifdef KS_INTRA_TIMING
    long long tb_t0 = clock64();
#endif
    /* a. residual row (the source row stays in registers for the statistics of step h) */
#pragma unroll
    for (int x = 0; x < N; x += 4) {
        uint32_t p4 = *reinterpret_cast<const uint32_t *>(pred_row + x);
        uint32_t s4 = valid ? *reinterpret_cast<const uint32_t *>(src_row + x) : p4;
        srcw[x >> 2] = s4;
#pragma unroll
        for (int b = 0; b < 4; b++) { pred[x + b] = (uint8_t)(p4 >> (8 * b)); res[x + b] = (int)((s4 >> (8 * b)) & 255) - (int)pred[x + b]; }
    }
    /* b. forward pass 1 (rows), stage shift 2*log2N-2; results stored transposed: S[u][r] */
    ks_fwd_pass<N>(res, 2 * LOG2 - 2, t0, [&](int u, int v) { S[u * SP + r] = (int16_t)v; });
    __syncwarp();
    KS_TBT(0);
    /* c. forward pass 2 (columns): lane = horizontal frequency u = r; d. quantise each coefficient (v, u=r) as it appears */
    if constexpr (N == 4) {
        const uint2 q = *reinterpret_cast<const uint2 *>(&S[r * SP]);
        res[0] = (int)(short)(q.x & 0xffffu); res[1] = (int)q.x >> 16; res[2] = (int)(short)(q.y & 0xffffu); res[3] = (int)q.y >> 16;
    } else {
#pragma unroll
        for (int y = 0; y < N; y += 8) {
            uint4 q = *reinterpret_cast<const uint4 *>(&S[r * SP + y]);
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) { res[y + 2 * j] = (int)(short)(w[j] & 0xffffu); res[y + 2 * j + 1] = (int)w[j] >> 16; }
        }
    }
    const int qbits = 21 + qp / 6 - LOG2, scale = c_quant_scales[qp % 6];
    const int add = (intra_slice ? 171 : 85) << (qbits - 9);
    bool nz = false;
    ks_fwd_pass<N>(res, 7, t0, [&](int v, int coef) {
        int c = (int)(short)coef, a = abs(c), m = a * scale, lv = (m + add) >> qbits;
        int du = (m - (lv << qbits)) >> (qbits - 8);
        lv = min(lv, 32767);
        nz |= lv != 0;
        C[v * N + r] = (int16_t)c; L[v * N + r] = (int16_t)(c < 0 ? -lv : lv); D[v * N + r] = (int16_t)du;
    });
    unsigned nzb = __ballot_sync(0xffffffffu, nz && valid);
    bool cbf = (nzb & gmask) != 0;
    KS_TBT(1);
    /* e. sign-data hiding, one coefficient group per lane (uniform ballots first, then the per-CG pass) */
    if (sign_hiding && nzb) {
        constexpr int NCG = (N / 4) * (N / 4);
        constexpr int REPS = (G * NCG + 31) / 32;
        unsigned ball[REPS];
#pragma unroll
        for (int rep = 0; rep < REPS; rep++) {
            int id = lane + 32 * rep, tb = id / NCG, cg = id % NCG;
            bool cgnz = false;
            if (id < G * NCG) {
                const int16_t *Lt = sc->L + tb * N * N;
#pragma unroll
                for (int i = 0; i < 16; i++) { int s = scan[cg * 16 + i]; cgnz |= Lt[(s >> 8) * N + (s & 255)] != 0; }
            }
            ball[rep] = __ballot_sync(0xffffffffu, cgnz);
        }
#pragma unroll
        for (int rep = 0; rep < REPS; rep++) {
            int id = lane + 32 * rep, tb = id / NCG, cg = id % NCG;
            if (id < G * NCG && ((ball[rep] >> lane) & 1)) {
                unsigned long long m;
                if (REPS == 2) m = ((unsigned long long)ball[REPS - 1] << 32) | ball[0];
                else m = ball[0];
                constexpr unsigned long long TBMASK = NCG == 64 ? ~0ull : ((1ull << (NCG & 63)) - 1ull);
                unsigned long long tbm = (m >> ((tb * NCG) & 63)) & TBMASK;
                int top = 63 - __clzll((long long)tbm);
                ks_sbh_cg(sc->C + tb * N * N, sc->L + tb * N * N, sc->D + tb * N * N, scan, cg * 16, cg == top, N);
            }
        }
    }
    __syncwarp();
    KS_TBT(2);
    /* g. reconstruction */
    if (nzb) {
        const int shift = LOG2 - 1, dq = c_inv_quant_scales[qp % 6] << (qp / 6), rnd = 1 << (shift - 1);
        int t[N];
        /* inverse pass 1 (columns, lane = column x = r): inputs are dequantised on the fly */
        ks_inv_pass<N>([&](int k) { return ks_clip3(-32768, 32767, ((int)L[k * N + r] * dq + rnd) >> shift); }, t, 7, true, t0);
        __syncwarp();
#pragma unroll
        for (int y = 0; y < N; y++) S[y * SP + r] = (int16_t)t[y];
        __syncwarp();
        /* inverse pass 2 (rows, lane = row y = r) */
        ks_inv_pass<N>([&](int k) { return (int)S[r * SP + k]; }, t, 12, false, t0);
#pragma unroll
        for (int x = 0; x < N; x++) pred[x] = (uint8_t)ks_clip8((int)pred[x] + t[x]);
    }
    KS_TBT(3);
    /* h. RD zero-out and/or per-block statistics: sums over the N lanes of the group (xor butterflies stay inside the aligned group) */
    if ((rdz_lambda_q4 && nzb) || stat_out) {
        int nnz = 0, slog = 0, maxd = 0, d0 = 0, d1 = 0; unsigned cgm = 0;
#pragma unroll
        for (int x = 0; x < N; x += 4) {
            const uint2 l4 = *reinterpret_cast<const uint2 *>(&L[r * N + x]);
            const uint32_t p4 = *reinterpret_cast<const uint32_t *>(pred_row + x);
            const uint32_t s4 = srcw[x >> 2];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int l = (int)(short)(((b & 2) ? l4.y : l4.x) >> (16 * (b & 1))), a = abs(l);
                if (a) { nnz++; slog += 31 - __clz(a); maxd = max(maxd, x + b + r); cgm |= 1u << (x >> 2); }
                const int sv = (int)((s4 >> (8 * b)) & 255), e0 = sv - (int)((p4 >> (8 * b)) & 255), e1 = valid ? sv - (int)pred[x + b] : 0;
                d0 += e0 * e0; d1 += e1 * e1;
            }
        }
        cgm |= __shfl_xor_sync(0xffffffffu, cgm, 1); cgm |= __shfl_xor_sync(0xffffffffu, cgm, 2);
        int ncg = (r & 3) == 0 ? __popc(cgm) : 0;
        /* nnz <= 1024, slog <= 15 * 1024, ncg <= 64: one packed sum (redux.sync over sub-warp masks measured slower than the butterflies) */
        unsigned pk = (unsigned)nnz | ((unsigned)slog << 11) | ((unsigned)ncg << 25);
#pragma unroll
        for (int o = 1; o < N; o <<= 1) {
            pk += __shfl_xor_sync(0xffffffffu, pk, o); d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            maxd = max(maxd, __shfl_xor_sync(0xffffffffu, maxd, o));
        }
        nnz = (int)(pk & 2047u); slog = (int)((pk >> 11) & 16383u); ncg = (int)(pk >> 25);
        int bits = nnz ? 3 * nnz + 2 * slog + 4 * ncg + maxd : 0;
        if (rdz_lambda_q4 && nnz && (long long)d0 * 16 <= (long long)d1 * 16 + (long long)rdz_lambda_q4 * bits) {
            cbf = false; d1 = d0; bits = 0;
#pragma unroll
            for (int x = 0; x < N; x += 4) {
                *reinterpret_cast<uint2 *>(&L[r * N + x]) = make_uint2(0u, 0u);
                const uint32_t p4 = *reinterpret_cast<const uint32_t *>(pred_row + x);
#pragma unroll
                for (int b = 0; b < 4; b++) pred[x + b] = (uint8_t)(p4 >> (8 * b));
            }
        }
        if (stat_out && r == 0 && valid) { stat_out[g].d1 = d1; stat_out[g].bits = bits; }
    }
    KS_TBT(4);
    /* f. store the level row (dense plane) and the reconstruction */
    if (valid) {
        if constexpr (N == 4) *reinterpret_cast<uint2 *>(lev_row) = *reinterpret_cast<const uint2 *>(&L[r * N]);
        else {
#pragma unroll
            for (int x = 0; x < N; x += 8) *reinterpret_cast<uint4 *>(lev_row + x) = *reinterpret_cast<const uint4 *>(&L[r * N + x]);
        }
#pragma unroll
        for (int x = 0; x < N; x += 4)
            *reinterpret_cast<uint32_t *>(rec_row + x) = (uint32_t)pred[x] | ((uint32_t)pred[x + 1] << 8) | ((uint32_t)pred[x + 2] << 16) | ((uint32_t)pred[x + 3] << 24);
    }
    __syncwarp();
    KS_TBT(5);
    return cbf;
}

/* ------------------------------------------------------------------ inter picture: one CTA per CTU - */
#define KS_RECON_WARPS 4                       /* one warp per 32x32 quadrant of the CTU: every warp gets the same mix of work */
struct KsReconSmem {
    KsTbScratch tb[KS_RECON_WARPS];
    uint16_t scan[64 + 256 + 1024];            /* scan tables for 8x8, 16x16, 32x32 */
    int      t0[256];                          /* M32[2j+1][x], j,x < 16: level-0 odd part of the 32-point butterflies */
    int16_t  mvx[16], mvy[16];
    int      m1[16];                           /* B pictures: list-1 vector (x | y << 16) and direction folded into one word pair */
    uint8_t  dirv[16];
    uint8_t  valid[16];
    uint8_t  clog2[16];                        /* P pictures: CU size chosen by ks_decide_tree_kernel */
    uint8_t  intra[16];                        /* P pictures: intra CU (coded by ks_recon_intra_kernel afterwards): nothing to do here */
    unsigned cbf[16];                          /* KS_F_CBF_* bits per cell, OR-ed by the transform tasks */
};
__device__ __forceinline__ const uint16_t *ks_scan_ptr(const KsReconSmem *sm, int n) { return sm->scan + (n == 8 ? 0 : (n == 16 ? 64 : 320)); }
__device__ __forceinline__ void ks_load_t0(int *t0, int tid, int nthreads)
{
    for (int i = tid; i < 256; i += nthreads) t0[i] = c_dct[2 * (i >> 4) + 1][i & 15];
}
__device__ __forceinline__ void ks_load_scans(uint16_t *scan, int tid, int nthreads)
{
    for (int i = tid; i < 64 + 256 + 1024; i += nthreads) scan[i] = i < 64 ? c_scan_tb[1][i] : (i < 320 ? c_scan_tb[2][i - 64] : c_scan_tb[3][i - 320]);
}

template <int MINB>
__global__ void __launch_bounds__(KS_RECON_WARPS * KS_WARP, MINB)
ks_recon_inter_kernel(KsPicParams pp, KsPlanes src, KsPlanes pred, KsPlanes rec, KsLevels lv, ks_cell *__restrict__ cells, const ks_cell_b *__restrict__ cells_b)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KsReconSmem *sm = reinterpret_cast<KsReconSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ctx = blockIdx.x, cty = blockIdx.y, X0 = ctx << 6, Y0 = cty << 6;
    const int W = pp.W, H = pp.H, CW = W >> 1, CH = H >> 1;
    for (int i = tid; i < KS_RECON_TAB_U4; i += KS_RECON_WARPS * KS_WARP) reinterpret_cast<uint4 *>(sm->scan)[i] = __ldg(&g_recon_tab[i]);   /* scan[] and t0[] are adjacent */
    if (tid < 16) {
        int cx = tid & 3, cy = tid >> 2, x = X0 + (cx << 4), y = Y0 + (cy << 4);
        bool v = x < W && y < H;
        sm->valid[tid] = v; sm->cbf[tid] = 0;
        sm->m1[tid] = 0; sm->dirv[tid] = 1; sm->clog2[tid] = 4; sm->intra[tid] = 0;
        if (v) {
            ks_cell c = cells[(y >> 4) * pp.cw + (x >> 4)]; sm->mvx[tid] = c.mvx; sm->mvy[tid] = c.mvy; sm->clog2[tid] = c.cu_log2; sm->intra[tid] = (c.flags & KS_F_INTRA) != 0;
            if (cells_b) { ks_cell_b b = cells_b[(y >> 4) * pp.cw + (x >> 4)]; sm->m1[tid] = (int)(uint16_t)b.mvx1 | ((int)b.mvy1 << 16); sm->dirv[tid] = b.dir; }
        } else { sm->mvx[tid] = 0; sm->mvy[tid] = 0; }
    }
    __syncthreads();
    /* CU size.  P pictures: decided by ks_decide_kernel (cu_log2 of the cells).  B pictures: four siblings with equal motion merge upward
     * (16 -> 32 -> 64), mirror of ora_b_picture.  Every thread derives the same flags from shared memory: no serial section, no task list. */
    bool q32[4], c64 = true;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int b0 = (q & 1) * 2 + (q >> 1) * 8;
        if (!cells_b) { q32[q] = sm->valid[b0] && sm->clog2[b0] >= 5; c64 = c64 && q32[q] && sm->clog2[b0] == 6; continue; }
        bool ok = sm->valid[b0] && sm->valid[b0 + 1] && sm->valid[b0 + 4] && sm->valid[b0 + 5];
        const int mx = sm->mvx[b0], my = sm->mvy[b0], m1 = sm->m1[b0], dr = sm->dirv[b0];
        ok = ok && sm->mvx[b0 + 1] == mx && sm->mvx[b0 + 4] == mx && sm->mvx[b0 + 5] == mx
                && sm->mvy[b0 + 1] == my && sm->mvy[b0 + 4] == my && sm->mvy[b0 + 5] == my
                && sm->m1[b0 + 1] == m1 && sm->m1[b0 + 4] == m1 && sm->m1[b0 + 5] == m1
                && sm->dirv[b0 + 1] == dr && sm->dirv[b0 + 4] == dr && sm->dirv[b0 + 5] == dr;
        q32[q] = ok;
        c64 = c64 && ok && mx == sm->mvx[0] && my == sm->mvy[0] && m1 == sm->m1[0] && dr == sm->dirv[0];
    }
    /* ---- transform tasks: 16 slots (k, q) = (kind 0..3, quadrant); a quadrant coded with one 32x32 TU uses kinds
     *      0 (luma 32) and 1 (Cb+Cr 16), otherwise kinds 0,1 (two luma 16 pairs) and 2,3 (four Cb 8 / four Cr 8).
     *      Warp w owns quadrant w (slots w, w+4, w+8, w+12), so the warps of a CTA finish together whatever the CU sizes. ---- */
#pragma unroll 1
    for (int sl = warp; sl < 16; sl += KS_RECON_WARPS) {
        const int k = sl >> 2, q = sl & 3;
        const int qx = (q & 1) * 32, qy = (q >> 1) * 32;             /* quadrant origin inside the CTU (luma) */
        const int b0 = (q & 1) * 2 + (q >> 1) * 8;
        if (!sm->valid[b0]) continue;
        KsTbScratch *ts = &sm->tb[warp];
        if (q32[q]) {
            if (k == 0) {
                int r = lane, x = X0 + qx, y = Y0 + qy + r;
                bool cbf = ks_tb_code<32>(ts, ks_scan_ptr(sm, 32), sm->t0, true, src.p[0] + (size_t)y * W + x, pred.p[0] + (size_t)y * W + x,
                                          rec.p[0] + (size_t)y * W + x, lv.p[0] + (size_t)y * W + x, pp.qp, 0, pp.sign_hiding, lane, pp.rdz_lambda_q4);
                if (lane == 0 && cbf) { atomicOr(&sm->cbf[b0], KS_F_CBF_Y); atomicOr(&sm->cbf[b0 + 1], KS_F_CBF_Y); atomicOr(&sm->cbf[b0 + 4], KS_F_CBF_Y); atomicOr(&sm->cbf[b0 + 5], KS_F_CBF_Y); }
            } else if (k == 1) {
                int g = lane >> 4, r = lane & 15, x = (X0 + qx) >> 1, y = ((Y0 + qy) >> 1) + r;
                bool cbf = ks_tb_code<16>(ts, ks_scan_ptr(sm, 16), sm->t0, true, src.p[1 + g] + (size_t)y * CW + x, pred.p[1 + g] + (size_t)y * CW + x,
                                          rec.p[1 + g] + (size_t)y * CW + x, lv.p[1 + g] + (size_t)y * CW + x, pp.qpc, 0, pp.sign_hiding, lane);
                if (r == 0 && cbf) { unsigned f = g ? KS_F_CBF_CR : KS_F_CBF_CB; atomicOr(&sm->cbf[b0], f); atomicOr(&sm->cbf[b0 + 1], f); atomicOr(&sm->cbf[b0 + 4], f); atomicOr(&sm->cbf[b0 + 5], f); }
            }
        } else if (k < 2) {
            int g = lane >> 4, r = lane & 15, kc = k * 2 + g;                 /* cell inside the quadrant */
            int cidx = b0 + (kc & 1) + (kc >> 1) * 4;
            bool v = sm->valid[cidx] && !sm->intra[cidx];
            int lx = qx + (kc & 1) * 16, ly = qy + (kc >> 1) * 16 + r, x = X0 + lx, y = Y0 + ly;
            bool cbf = ks_tb_code<16>(ts, ks_scan_ptr(sm, 16), sm->t0, v, src.p[0] + (size_t)y * W + x, pred.p[0] + (size_t)(v ? y : Y0) * W + (v ? x : X0),
                                      rec.p[0] + (size_t)y * W + x, lv.p[0] + (size_t)y * W + x, pp.qp, 0, pp.sign_hiding, lane, pp.rdz_lambda_q4);
            if (r == 0 && cbf) atomicOr(&sm->cbf[cidx], KS_F_CBF_Y);
        } else {
            int g = lane >> 3, r = lane & 7, kc = g, ci = k - 2;
            int cidx = b0 + (kc & 1) + (kc >> 1) * 4;
            bool v = sm->valid[cidx] && !sm->intra[cidx];
            int lx = (qx >> 1) + (kc & 1) * 8, ly = (qy >> 1) + (kc >> 1) * 8 + r, x = (X0 >> 1) + lx, y = (Y0 >> 1) + ly;
            bool cbf = ks_tb_code<8>(ts, ks_scan_ptr(sm, 8), sm->t0, v, src.p[1 + ci] + (size_t)y * CW + x, pred.p[1 + ci] + (size_t)(v ? y : (Y0 >> 1)) * CW + (v ? x : (X0 >> 1)),
                                     rec.p[1 + ci] + (size_t)y * CW + x, lv.p[1 + ci] + (size_t)y * CW + x, pp.qpc, 0, pp.sign_hiding, lane);
            if (r == 0 && cbf) atomicOr(&sm->cbf[cidx], ci ? KS_F_CBF_CR : KS_F_CBF_CB);
        }
    }
    __syncthreads();
    if (tid < 16 && sm->valid[tid] && !sm->intra[tid]) {
        int cx = tid & 3, cy = tid >> 2;
        ks_cell c; c.mvx = sm->mvx[tid]; c.mvy = sm->mvy[tid]; c.cu_log2 = (uint8_t)(c64 ? 6 : (q32[(cx >> 1) + (cy >> 1) * 2] ? 5 : 4)); c.flags = (uint8_t)sm->cbf[tid]; c.intra_mode = 0; c.rsv = 0;
        cells[((Y0 >> 4) + cy) * pp.cw + (X0 >> 4) + cx] = c;
    }
}

void ks_launch_recon_inter(const KsPicParams &pp, KsPlanes src, KsPlanes pred, KsPlanes rec, KsLevels lv, ks_cell *cells, const ks_cell_b *cells_b, cudaStream_t st)
{
    static const int minb = getenv("KS_RECON_MINB") ? atoi(getenv("KS_RECON_MINB")) : 4;       /* tuning knob: 4 (128 registers) or 5 (102) CTAs per SM */
    dim3 grid(pp.ctw, pp.cth);
    if (minb >= 5) ks_recon_inter_kernel<5><<<grid, KS_RECON_WARPS * KS_WARP, sizeof(KsReconSmem), st>>>(pp, src, pred, rec, lv, cells, cells_b);
    else ks_recon_inter_kernel<4><<<grid, KS_RECON_WARPS * KS_WARP, sizeof(KsReconSmem), st>>>(pp, src, pred, rec, lv, cells, cells_b);
}

/* ------------------------------------------------------------------ intra picture (wavefront) ---- */
/* predicted sample (x,y) of an n x n block for `mode` (spec 8.4.4.2.4-.6 == ora_intra_pred); p = reference
 * array (already filtered if the mode asks for it): p[2n-1-i] = left[i], p[2n] = corner, p[2n+1+i] = top[i] */
__device__ __forceinline__ int ks_intra_sample(const uint8_t *p, int n, int log2n, int mode, int x, int y, int dc, bool edge)
{
    const int n2 = 2 * n;
    if (mode == 0)
        return ((n - 1 - x) * p[n2 - 1 - y] + (x + 1) * p[n2 + 1 + n] + (n - 1 - y) * p[n2 + 1 + x] + (y + 1) * p[n2 - 1 - n] + n) >> (log2n + 1);
    if (mode == 1) {
        if (edge) {
            if (x == 0 && y == 0) return (p[n2 - 1] + 2 * dc + p[n2 + 1] + 2) >> 2;
            if (y == 0) return (p[n2 + 1 + x] + 3 * dc + 2) >> 2;
            if (x == 0) return (p[n2 - 1 - y] + 3 * dc + 2) >> 2;
        }
        return dc;
    }
    const int ang = c_intra_angle[mode], inv = c_intra_inv_angle[mode];
    const bool vert = mode >= 18;
    const int i = vert ? x : y, j = vert ? y : x;
    if (ang == 0 && edge && i == 0) {
        /* mode 26: pred[0][y] = top[0] + ((left[y]-corner)>>1); mode 10: pred[x][0] = left[0] + ((top[x]-corner)>>1) */
        return vert ? ks_clip8(p[n2 + 1] + ((p[n2 - 1 - j] - p[n2]) >> 1)) : ks_clip8(p[n2 - 1] + ((p[n2 + 1 + j] - p[n2]) >> 1));
    }
    const int idx = ((j + 1) * ang) >> 5, f = ((j + 1) * ang) & 31;
    int k0 = i + idx + 1, k1 = k0 + 1;
    auto refv = [&](int k) -> int {
        if (k >= 0) return vert ? p[n2 + k] : p[n2 - k];
        int i2 = -1 + ((k * inv + 128) >> 8);
        return vert ? p[n2 - 1 - i2] : p[n2 + 1 + i2];
    };
    int a = refv(k0);
    if (!f) return a;
    return ((32 - f) * a + f * refv(k1) + 16) >> 5;
}

#define KS_INTRA_WARPS 4              /* CTA of the dependent block coder: warps 0..2 substitute the three components, warps 0/1 code luma/chroma */
#define KS_MODES_WARPS 4              /* CTA of the (independent) mode search: the 35 modes are spread over the warps */
#define KS_SPLIT8_MIN_BITS 200          /* == ORA_SPLIT8_MIN_BITS */
/* luma modes of one cell, chosen by ks_intra_modes_kernel: the 16x16 CU's and the four 8x8 CUs' */
struct KsIntraModes { uint8_t m16, m8[4], rsv[3]; };

struct KsIntraSmem {
    KsTbScratch tb[2];
    uint16_t scan[64 + 256 + 1024];
    uint16_t scan4[3][16];       /* 4x4 blocks: diagonal, horizontal, vertical (7.4.9.11 scanIdx 0, 1, 2) */
    uint16_t scan8hv[2][64];     /* 8x8 blocks: horizontal, vertical (groups in the same order, 6.5.4 / 6.5.5) */
    uint8_t  nb[3][72];          /* substituted reference samples: luma 4n+1, chroma 2n+1 */
    uint8_t  raw[3][72];         /* as loaded (before substitution) */
    uint8_t  fb[72];             /* [1 2 1]-filtered luma references */
    uint8_t  av[3][72];
    uint8_t  predY[16 * 16];
    uint8_t  predC[2][8 * 8];
    alignas(16) uint8_t srcY[16 * 16];  /* the cell's source block, prefetched while the previous cell is coded */
    alignas(16) uint8_t srcC[2][8 * 8];
    int      dc[3];
    int      ticket;
    unsigned cbf;
    KsTbStat stat_y[4], stat_c[8];      /* ks_tb_code statistics of the block just coded: luma [0], Cb [0], Cr [1] */
    uint8_t  save_rec[256 + 64 + 64];   /* the 16x16 CU's result while the four 8x8 CUs are tried over it */
    int16_t  save_lev[256 + 64 + 64];
    int      try8, use8;
    long long j16, j8;
    int      cbf16, cbf8[3];
    alignas(8) KsIntraModes modes;
    long long t0, acc[8];
};
/* working set of the mode search (one cell per CTA) */
struct KsModesSmem {
    uint8_t  nb[72], raw[72], fb[72], av[72];
    uint8_t  mref[KS_MODES_WARPS][64];  /* per-warp extended main reference of the mode under test: index k+n, k = -n..2n */
    unsigned best_key;
    int      dc;
};

/* reference-sample substitution (spec 8.4.4.2.2) for one component by one warp: every entry takes the nearest available
 * entry at or before it in scan order (bottom-left -> corner -> top-right), leading unavailable entries take the first
 * available one, 128 if nothing is available.  Also the DC value.  tot = 4n+1 (<= 65). */
__device__ __forceinline__ void ks_intra_substitute(const uint8_t *raw, const uint8_t *av, uint8_t *nb, int tot, int n, int *dc_out, int lane)
{
    const unsigned m0 = __ballot_sync(0xffffffffu, lane < tot && av[lane]);
    const unsigned m1 = __ballot_sync(0xffffffffu, 32 + lane < tot && av[32 + lane]);
    const unsigned m2 = __ballot_sync(0xffffffffu, 64 + lane < tot && av[64 + lane]);
    const int first = m0 ? __ffs(m0) - 1 : (m1 ? 32 + __ffs(m1) - 1 : (m2 ? 64 + __ffs(m2) - 1 : -1));
    int dc = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int i = lane + 32 * k;
        if (i < tot) {
            const unsigned own = k == 0 ? m0 : (k == 1 ? m1 : m2), cand = own & (0xffffffffu >> (31 - lane));
            int j;
            if (cand) j = 32 * k + 31 - __clz(cand);
            else if (k >= 1 && (k == 2 ? m1 : m0)) { unsigned w = k == 2 ? m1 : m0; j = 32 * (k - 1) + 31 - __clz(w); }
            else if (k == 2 && m0) j = 31 - __clz(m0);
            else j = first;
            const int v = j < 0 ? 128 : raw[j];
            nb[i] = (uint8_t)v;
            if ((i >= n && i < 2 * n) || (i > 2 * n && i <= 3 * n)) dc += v;       /* left[0..n) and top[0..n) */
        }
    }
    dc = (int)ks_warp_sum((unsigned)dc);
    if (lane == 0) *dc_out = (dc + n) >> (n == 16 ? 5 : (n == 8 ? 4 : 3));
    __syncwarp();
}

__device__ __forceinline__ void ks_intra_load_small_scans(KsIntraSmem &sm, int tid)
{
    if (tid < 16) { sm.scan4[0][tid] = c_scan_tb[0][tid]; sm.scan4[1][tid] = (uint16_t)(((tid >> 2) << 8) | (tid & 3)); sm.scan4[2][tid] = (uint16_t)(((tid & 3) << 8) | (tid >> 2)); }
    if (tid < 64) {
        const int c = tid >> 4, k = tid & 15, gx = c & 1, gy = c >> 1, px = k & 3, py = k >> 2;
        sm.scan8hv[0][tid] = (uint16_t)((((gy << 2) + py) << 8) | ((gx << 2) + px));
        sm.scan8hv[1][tid] = (uint16_t)((((gx << 2) + px) << 8) | ((gy << 2) + py));
    }
}

/* ---- the 35-mode search of one intra block (LG = 4: 16x16, 3: 8x8) against the SOURCE picture's neighbours, by the KS_MODES_WARPS warps of a CTA:
 *      SAD + lambda * bits; warp w evaluates modes w, w+4, ... on all N*N samples.  Angular modes first project the (possibly filtered) references
 *      onto one extended main-reference array per mode (spec 8.4.4.2.6 ref[]), so a sample costs two shared loads + one interpolation.
 *      Taking the neighbours from the source (not the reconstruction) makes every block of the picture independent: the search leaves the
 *      wavefront's dependency chain.  Returns the mode (in every thread); ends with a CTA barrier.  Mirror of ora intra_block's decision. ---- */
template <int LG>
__device__ __forceinline__ int ks_intra_search(KsModesSmem &sm, const KsPicParams &pp, const uint8_t *__restrict__ srcY, int x0, int y0, int tid, int warp, int lane)
{
    constexpr int N = 1 << LG, TL = 4 * N + 1, FTHR = N == 16 ? 1 : 7, PER = N * N / 32, STEP = 32 / N;
    const int W = pp.W, H = pp.H;
    if (tid < TL) {
        int xn, yn;
        if (tid < 2 * N) { xn = x0 - 1; yn = y0 + 2 * N - 1 - tid; }
        else if (tid == 2 * N) { xn = x0 - 1; yn = y0 - 1; }
        else { xn = x0 + tid - 2 * N - 1; yn = y0 - 1; }
        const bool a = ks_avail(W, H, pp.ctw, x0, y0, xn, yn);
        sm.av[tid] = a;
        sm.raw[tid] = a ? srcY[(size_t)yn * W + xn] : 0;
    }
    if (tid == 127) sm.best_key = 0xffffffffu;
    __syncthreads();
    if (warp == 0) {
        ks_intra_substitute(sm.raw, sm.av, sm.nb, TL, N, &sm.dc, lane);
        for (int i = lane; i < TL; i += 32)
            sm.fb[i] = (i == 0 || i == 4 * N) ? sm.nb[i] : (uint8_t)((sm.nb[i - 1] + 2 * sm.nb[i] + sm.nb[i + 1] + 2) >> 2);
    }
    __syncthreads();
    {
        unsigned best = 0xffffffffu;
        const int px = lane & (N - 1), py0 = lane >> LG;
        uint8_t s[PER];
#pragma unroll
        for (int j = 0; j < PER; j++) s[j] = srcY[(size_t)(y0 + py0 + STEP * j) * W + x0 + px];
        uint8_t *mref = sm.mref[warp];
#pragma unroll 1
        for (int m = warp; m < 35; m += KS_MODES_WARPS) {
            int d1 = abs(m - 26), d2 = abs(m - 10);
            bool filt = m != 1 && min(d1, d2) > FTHR;            /* intraHorVerDistThres[nTbS] */
            const uint8_t *p = filt ? sm.fb : sm.nb;
            unsigned sad = 0;
            if (m < 2) {
#pragma unroll
                for (int j = 0; j < PER; j++) sad += abs(ks_intra_sample(p, N, LG, m, px, py0 + STEP * j, sm.dc, true) - (int)s[j]);
            } else {
                const int ang = c_intra_angle[m], inv = c_intra_inv_angle[m];
                const bool vert = m >= 18;
                __syncwarp();
                for (int e = lane; e < 3 * N + 1; e += 32) {           /* k = e - N in -N..2N */
                    int k = e - N, v;
                    if (k >= 0) v = vert ? p[2 * N + k] : p[2 * N - k];
                    else { int i2 = -1 + ((k * inv + 128) >> 8); i2 = min(max(i2, -1), 2 * N - 1); v = vert ? p[2 * N - 1 - i2] : p[2 * N + 1 + i2]; }
                    mref[e] = (uint8_t)v;
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < PER; j++) {
                    const int x = px, y = py0 + STEP * j, ii = vert ? x : y, jj = vert ? y : x;
                    int v;
                    if (ang == 0 && ii == 0)
                        v = vert ? ks_clip8(p[2 * N + 1] + ((p[2 * N - 1 - jj] - p[2 * N]) >> 1)) : ks_clip8(p[2 * N - 1] + ((p[2 * N + 1 + jj] - p[2 * N]) >> 1));
                    else {
                        const int idx = ((jj + 1) * ang) >> 5, f = ((jj + 1) * ang) & 31;
                        const int a = mref[N + ii + idx + 1], b2 = mref[N + ii + idx + 2];
                        v = f ? ((32 - f) * a + f * b2 + 16) >> 5 : a;
                    }
                    sad += abs(v - (int)s[j]);
                }
            }
            sad = ks_warp_sum(sad);
            int bits = (m == 0 || m == 1 || m == 10 || m == 26) ? 3 : 6;
            unsigned key = ((sad + ((pp.lambda_sad_q4 * bits) >> 4)) << 6) | (unsigned)m;
            best = min(best, key);
        }
        if (lane == 0) atomicMin(&sm.best_key, best);
    }
    __syncthreads();
    const int mode = (int)(sm.best_key & 63);
    __syncthreads();                                     /* best_key is reset by the next call */
    return mode;
}

/* the luma modes of every intra cell of the picture (I pictures: all cells incl. the four 8x8 alternatives; P pictures: the cells flagged
 * KS_F_INTRA, 16x16 only): one CTA per cell, no dependencies between cells */
__global__ void __launch_bounds__(KS_MODES_WARPS * KS_WARP)
ks_intra_modes_kernel(KsPicParams pp, const uint8_t *__restrict__ srcY, const ks_cell *__restrict__ cells, const int *__restrict__ n_intra, KsIntraModes *__restrict__ modes)
{
    __shared__ KsModesSmem sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cell = blockIdx.x, cy = cell / pp.cw, cx = cell - cy * pp.cw, x0 = cx << 4, y0 = cy << 4;
    const bool masked = n_intra != nullptr;
    if (masked && (*n_intra == 0 || !(cells[cell].flags & KS_F_INTRA))) return;
    KsIntraModes r; r.rsv[0] = r.rsv[1] = r.rsv[2] = 0; r.m8[0] = r.m8[1] = r.m8[2] = r.m8[3] = 0;
    r.m16 = (uint8_t)ks_intra_search<4>(sm, pp, srcY, x0, y0, tid, warp, lane);
    if (!masked) {
#pragma unroll 1
        for (int k = 0; k < 4; k++) r.m8[k] = (uint8_t)ks_intra_search<3>(sm, pp, srcY, x0 + 8 * (k & 1), y0 + 8 * (k >> 1), tid, warp, lane);
    }
    if (tid == 0) modes[cell] = r;
}

/* one intra CU of 16x16 (LG = 4) or 8x8 (LG = 3) with a GIVEN luma mode, by a CTA of KS_INTRA_WARPS warps: reference samples (reconstruction,
 * with availability), prediction, luma + chroma (DM) residual coding with the mode-dependent scans.  Results in shared memory: cbf,
 * stat_y[0] / stat_c[0..1].  Ends with a CTA barrier.  Mirror of ora intra_block after its decision. */
#ifdef KS_INTRA_TIMING
#include <cstdio>
__device__ long long g_it[8];
#define KS_T(k) do { if (tid == 0) { long long t_ = clock64(); sm.acc[k] += t_ - sm.t0; sm.t0 = t_; } } while (0)
#else
#define KS_T(k) do {} while (0)
#endif
template <int LG>
__device__ __forceinline__ void ks_intra_code_block(KsIntraSmem &sm, const KsPicParams &pp, const KsPlanes &src, const KsPlanes &rec, const KsLevels &lv,
                                                    int x0, int y0, int mode, int intra_slice, int tid, int warp, int lane)
{
    constexpr int N = 1 << LG, NC = N / 2, TL = 4 * N + 1, TC = 2 * N + 1, FTHR = N == 16 ? 1 : 7, NT = KS_INTRA_WARPS * KS_WARP;
    const int W = pp.W, H = pp.H, CW = W >> 1;
    /* 1. reference samples with availability (8.4.4.2.2); neighbours come from the pre-filter reconstruction.
     *    __ldcg: other CTAs (or earlier blocks of this one) wrote these lines, bypass the (non-coherent) L1 */
    for (int idx = tid; idx < TL + 2 * TC; idx += NT) {
        int ci = idx < TL ? 0 : (idx < TL + TC ? 1 : 2), i = idx - (ci == 0 ? 0 : (ci == 1 ? TL : TL + TC));
        int n = ci ? NC : N, sh = ci ? 1 : 0, bx = x0 >> sh, by = y0 >> sh, xn, yn;
        if (i < 2 * n) { xn = bx - 1; yn = by + 2 * n - 1 - i; }
        else if (i == 2 * n) { xn = bx - 1; yn = by - 1; }
        else { xn = bx + i - 2 * n - 1; yn = by - 1; }
        bool a = ks_avail(W, H, pp.ctw, x0, y0, xn << sh, yn << sh);
        sm.av[ci][i] = a;
        sm.raw[ci][i] = a ? __ldcg(rec.p[ci] + (size_t)yn * (W >> sh) + xn) : 0;
    }
    if (tid == 0) sm.cbf = 0;
    __syncthreads();
    KS_T(1);
    if (warp < 3) {
        ks_intra_substitute(sm.raw[warp], sm.av[warp], sm.nb[warp], warp ? TC : TL, warp ? NC : N, &sm.dc[warp], lane);
        if (warp == 0)
            for (int i = lane; i < TL; i += 32)
                sm.fb[i] = (i == 0 || i == 4 * N) ? sm.nb[0][i] : (uint8_t)((sm.nb[0][i - 1] + 2 * sm.nb[0][i] + sm.nb[0][i + 1] + 2) >> 2);
    }
    __syncthreads();
    KS_T(2);
    const int scan_idx = (mode >= 22 && mode <= 30) ? 1 : ((mode >= 6 && mode <= 14) ? 2 : 0);
    /* 2. prediction blocks: N*N luma samples, then 2 x NC*NC chroma */
    {
        const int d1 = abs(mode - 26), d2 = abs(mode - 10);
        const bool filt = mode != 1 && min(d1, d2) > FTHR;
        for (int t = tid; t < N * N + 2 * NC * NC; t += NT) {
            if (t < N * N) sm.predY[t] = (uint8_t)ks_intra_sample(filt ? sm.fb : sm.nb[0], N, LG, mode, t & (N - 1), t >> LG, sm.dc[0], true);
            else {
                const int u = t - N * N, ci = u / (NC * NC), k = u % (NC * NC);
                sm.predC[ci][k] = (uint8_t)ks_intra_sample(sm.nb[1 + ci], NC, LG - 1, mode, k & (NC - 1), k / NC, sm.dc[1 + ci], false);
            }
        }
    }
    __syncthreads();
    KS_T(3);
    /* 3. residual coding: warp 0 luma (lanes 0..N-1), warp 1 Cb + Cr (lanes 0..2*NC-1) */
    if (warp == 0) {
        int g = lane / N, r = lane % N, y = y0 + r;
        const uint16_t *scan = LG == 4 ? sm.scan + 64 : (scan_idx == 0 ? sm.scan : sm.scan8hv[scan_idx - 1]);
        bool cbf = ks_tb_code<N>(&sm.tb[0], scan, nullptr, g == 0, &sm.srcY[((y0 & 15) + r) * 16 + (x0 & 15)], &sm.predY[r * N],
                                 rec.p[0] + (size_t)y * W + x0, lv.p[0] + (size_t)y * W + x0, pp.qp, intra_slice, pp.sign_hiding, lane, 0, intra_slice ? sm.stat_y : nullptr);
        if (lane == 0 && cbf) atomicOr(&sm.cbf, KS_F_CBF_Y);
    } else if (warp == 1) {
        int g = lane / NC, r = lane % NC, ci = g & 1, x = x0 >> 1, y = (y0 >> 1) + r;
        const uint16_t *scan = LG == 4 ? sm.scan : sm.scan4[scan_idx];
        bool cbf = ks_tb_code<NC>(&sm.tb[1], scan, nullptr, g < 2, &sm.srcC[ci][((((y0 >> 1) & 7) + r) & 7) * 8 + ((x0 >> 1) & 7)], &sm.predC[ci][r * NC],
                                  rec.p[1 + ci] + (size_t)y * CW + x, lv.p[1 + ci] + (size_t)y * CW + x, pp.qpc, intra_slice, pp.sign_hiding, lane, 0, intra_slice ? sm.stat_c : nullptr);
        if (r == 0 && g < 2 && cbf) atomicOr(&sm.cbf, ci ? KS_F_CBF_CR : KS_F_CBF_CB);
    }
    __syncthreads();
    KS_T(4);
}

/* one 16x16 intra cell by a CTA: a 16x16 CU, or four 8x8 CUs coded in z-order when that is cheaper in J = 16 * SSE + lambda * bits (both
 * are really coded; the 8x8 alternative is only tried in I pictures and when the 16x16 luma block costs at least KS_SPLIT8_MIN_BITS estimated bits).
 * The luma modes come from ks_intra_modes_kernel.  Writes the cell record.  Mirror of ora intra_cell. */
__device__ __forceinline__ void ks_intra_code_cell(KsIntraSmem &sm, const KsPicParams &pp, const KsPlanes &src, const KsPlanes &rec, const KsLevels &lv,
                                                   ks_cell *__restrict__ cells, const KsIntraModes *__restrict__ modes, int x0, int y0, int intra_slice, int tid, int warp, int lane)
{
    const int W = pp.W, NT = KS_INTRA_WARPS * KS_WARP;
    const long long lamq = pp.lambda_sse_q4;
    const KsIntraModes md = sm.modes;                /* staged, like the source block, by the caller */
    ks_intra_code_block<4>(sm, pp, src, rec, lv, x0, y0, md.m16, intra_slice, tid, warp, lane);
    if (tid == 0) {                                  /* (the block statistics only exist in I pictures: P-picture intra cells stay 16x16) */
        sm.cbf16 = (int)sm.cbf;
        if (intra_slice) sm.j16 = 16ll * sm.stat_y[0].d1 + lamq * (sm.stat_y[0].bits + 1) + 16ll * sm.stat_c[0].d1 + lamq * (sm.stat_c[0].bits + 1)
               + 16ll * sm.stat_c[1].d1 + lamq * (sm.stat_c[1].bits + 1) + lamq * 8;
        sm.try8 = intra_slice ? sm.stat_y[0].bits >= KS_SPLIT8_MIN_BITS : 0;       /* I pictures only: the intra CUs of P pictures stay 16x16 */
        sm.j8 = lamq * (8 * 4 + 2);
        sm.cbf8[0] = sm.cbf8[1] = sm.cbf8[2] = 0;
    }
    __syncthreads();
    if (sm.try8) {
        /* keep the 16x16 result (its stores are visible after the barrier above), then code the four 8x8 CUs over it */
        for (int t = tid; t < 384; t += NT) {
            const int ci = t < 256 ? 0 : (t < 320 ? 1 : 2), k = t - (ci == 0 ? 0 : (ci == 1 ? 256 : 320)), m = ci ? 8 : 16, sh = ci ? 1 : 0;
            const size_t o = (size_t)((y0 >> sh) + k / m) * (W >> sh) + (x0 >> sh) + k % m;
            sm.save_rec[t] = __ldcg(rec.p[ci] + o); sm.save_lev[t] = __ldcg(lv.p[ci] + o);
        }
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            ks_intra_code_block<3>(sm, pp, src, rec, lv, x0 + 8 * (k & 1), y0 + 8 * (k >> 1), md.m8[k], intra_slice, tid, warp, lane);
            if (tid == 0) {
                if (sm.cbf & KS_F_CBF_Y) sm.cbf8[0] |= 1 << k; if (sm.cbf & KS_F_CBF_CB) sm.cbf8[1] |= 1 << k; if (sm.cbf & KS_F_CBF_CR) sm.cbf8[2] |= 1 << k;
                sm.j8 += 16ll * sm.stat_y[0].d1 + lamq * (sm.stat_y[0].bits + 1) + 16ll * sm.stat_c[0].d1 + lamq * (sm.stat_c[0].bits + 1)
                       + 16ll * sm.stat_c[1].d1 + lamq * (sm.stat_c[1].bits + 1);
            }
            __syncthreads();
        }
        if (tid == 0) sm.use8 = sm.j8 < sm.j16;
        __syncthreads();
        if (sm.use8) {
            if (tid == 0) {
                ks_cell c; c.cu_log2 = 3;
                c.flags = (uint8_t)(KS_F_INTRA | (sm.cbf8[0] ? KS_F_CBF_Y : 0) | (sm.cbf8[1] ? KS_F_CBF_CB : 0) | (sm.cbf8[2] ? KS_F_CBF_CR : 0));
                c.mvx = (int16_t)(uint16_t)(md.m8[0] | (md.m8[1] << 8)); c.mvy = (int16_t)(uint16_t)(md.m8[2] | (md.m8[3] << 8));
                c.intra_mode = (uint8_t)(sm.cbf8[0] | (sm.cbf8[1] << 4)); c.rsv = (uint8_t)sm.cbf8[2];
                cells[(y0 >> 4) * pp.cw + (x0 >> 4)] = c;
            }
            __syncthreads();
            return;
        }
        for (int t = tid; t < 384; t += NT) {          /* the 16x16 CU wins: put its reconstruction and levels back */
            const int ci = t < 256 ? 0 : (t < 320 ? 1 : 2), k = t - (ci == 0 ? 0 : (ci == 1 ? 256 : 320)), m = ci ? 8 : 16, sh = ci ? 1 : 0;
            const size_t o = (size_t)((y0 >> sh) + k / m) * (W >> sh) + (x0 >> sh) + k % m;
            rec.p[ci][o] = sm.save_rec[t]; lv.p[ci][o] = sm.save_lev[t];
        }
    }
    if (tid == 0) {
        ks_cell c; c.mvx = 0; c.mvy = 0; c.cu_log2 = 4; c.flags = (uint8_t)(KS_F_INTRA | sm.cbf16); c.intra_mode = md.m16; c.rsv = 0;
        cells[(y0 >> 4) * pp.cw + (x0 >> 4)] = c;
    }
    __syncthreads();          /* the caller publishes the cell as finished right after this function */
}

/* flag hand-off between CTAs: release store / acquire load at gpu scope (polling with atomics + __nanosleep cost ~9 us per hand-off) */
__device__ __forceinline__ int ks_ld_acquire(const int *p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void ks_st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

/* The dependent pass of the intra CUs (I pictures: every cell; P pictures: the cells the CU decision flagged KS_F_INTRA, after the inter
 * reconstruction of everything else).  Which neighbours a block may read is fixed by the normative z-scan order, but the order in which blocks
 * are PROCESSED is free as long as those neighbours are finished.  So: one CTA per 16-sample cell row (rows handed out by a ticket counter,
 * top first), walking its row left to right; before a cell it waits -- per-cell done flags in HBM -- for exactly the cells it reads:
 * above-left, above, above-right and below-left, each only if the z-scan order makes it available (and, in P pictures, only if it is an intra
 * cell; inter cells were finished by the previous kernel).  Every wait is on a cell that precedes this one in decoding order, and a row only
 * ever waits on cells of rows that are running (all rows are resident: <= 270 CTAs at 8K, 3 per SM): no deadlock.  The critical path is
 * (cells per row + 2 x rows) cell steps instead of the (CTUs per row + 2 x CTU rows) x 16 of a CTU wavefront: 510 vs 2048 steps at 4K.
 * (The CTU wavefront kernel this replaces spent 28 % of its issue slots polling for the left / upper CTU.) */
__global__ void __launch_bounds__(KS_INTRA_WARPS * KS_WARP, 4)       /* <= 128 registers: a waiting row must not lock other shards' kernels out of the SM */
ks_recon_intra_rows_kernel(KsPicParams pp, KsPlanes src, KsPlanes rec, KsLevels lv, ks_cell *__restrict__ cells, int *sync_ws, const int *__restrict__ n_intra,
                           const KsIntraModes *__restrict__ modes)
{
    __shared__ __align__(16) KsIntraSmem sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool masked = n_intra != nullptr;
    if (masked && *n_intra == 0) return;
    const int W = pp.W, H = pp.H, cw = pp.cw;
    int *ticket = sync_ws, *done = sync_ws + 1;                  /* done[cell] = 1 once an intra cell's reconstruction is in HBM */
    ks_load_scans(sm.scan, tid, blockDim.x);
    ks_intra_load_small_scans(sm, tid);
    if (tid == 0) sm.ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int gy = sm.ticket;
    if (gy >= pp.ch) return;
    const int y0 = gy << 4;
#ifdef KS_INTRA_TIMING
    unsigned long long sm_row_start = 0;
    if (tid == 0) { for (int k = 0; k < 8; k++) sm.acc[k] = 0; sm.t0 = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(sm_row_start)); }
#endif
    __shared__ uint8_t rowflag[512];                             /* this row's intra flags (cw <= 512 covers 8K) */
    for (int i = tid; i < cw; i += KS_INTRA_WARPS * KS_WARP) rowflag[i] = masked ? (cells[gy * cw + i].flags & KS_F_INTRA) : 1;
    __syncthreads();
    /* the source block and the modes of a cell do not depend on anything: every thread carries one word of the NEXT cell's
     * (64 luma words, 2 x 16 chroma words, 2 words of modes) in a register while the current cell is coded */
    auto next_cell = [&](int from) { while (from < cw && !rowflag[from]) from++; return from; };
    auto prefetch = [&](int cx) -> uint32_t {
        if (cx >= cw) return 0u;
        if (tid < 64) return __ldg(reinterpret_cast<const uint32_t *>(src.p[0] + (size_t)(y0 + (tid >> 2)) * W + (cx << 4) + 4 * (tid & 3)));
        if (tid < 96) { const int ci = (tid - 64) >> 4, k = (tid - 64) & 15; return __ldg(reinterpret_cast<const uint32_t *>(src.p[1 + ci] + (size_t)((y0 >> 1) + (k >> 1)) * (W >> 1) + (cx << 3) + 4 * (k & 1))); }
        if (tid < 98) return __ldg(reinterpret_cast<const uint32_t *>(&modes[gy * cw + cx]) + (tid - 96));
        return 0u;
    };
    int gx = next_cell(0);
    uint32_t pre = prefetch(gx);
#pragma unroll 1
    while (gx < cw) {
        const int x0 = gx << 4, gx_next = next_cell(gx + 1);
        if (tid < 64) reinterpret_cast<uint32_t *>(sm.srcY)[tid] = pre;
        else if (tid < 96) reinterpret_cast<uint32_t *>(&sm.srcC[0][0])[tid - 64] = pre;
        else if (tid < 98) reinterpret_cast<uint32_t *>(&sm.modes)[tid - 96] = pre;
        pre = prefetch(gx_next);
        if (tid < 4) {
            const int ox[4] = {-1, 0, 1, -1}, oy[4] = {-1, -1, -1, 1};
            const int nx = gx + ox[tid], ny = gy + oy[tid];
            if (ks_avail(W, H, pp.ctw, x0, y0, nx << 4, ny << 4) && (!masked || (cells[ny * cw + nx].flags & KS_F_INTRA))) {
#ifdef KS_INTRA_TIMING
                int polls = 0, v;
                while ((v = ks_ld_acquire(&done[ny * cw + nx])) == 0) polls++;
                if (polls) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); atomicAdd((unsigned long long *)&g_it[7], (unsigned long long)((unsigned)gt - (unsigned)v)); atomicAdd((unsigned long long *)&g_it[6], 1ull); }
#else
                while (ks_ld_acquire(&done[ny * cw + nx]) == 0) { }
#endif
            }
        }
        __syncthreads();
        KS_T(0);
#ifdef KS_INTRA_TIMING
        unsigned long long t_dep = 0; if (tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_dep));
#endif
        ks_intra_code_cell(sm, pp, src, rec, lv, cells, modes, x0, y0, masked ? 0 : 1, tid, warp, lane);
        KS_T(5);
#ifdef KS_INTRA_TIMING
        if (tid == 0) { __threadfence(); unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); ks_st_release(&done[gy * cw + gx], (int)((unsigned)gt | 1u)); }
#else
        if (tid == 0) { __threadfence(); ks_st_release(&done[gy * cw + gx], 1); }
#endif
        gx = gx_next;
    }
#ifdef KS_INTRA_TIMING
    if (tid == 0 && !masked && (gy % 8 == 0 || gy == pp.ch - 1)) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); printf("row %3d blk %3d sm-clock end %lld globaltimer end %llu first-cell-start %llu\n", gy, blockIdx.x, clock64(), gt, sm_row_start); }
    if (tid == 0 && (gy == 5 || gy == 64) && !masked) printf("intra timing row %d (cycles):", gy), printf(" wait %lld fetch %lld subst %lld pred %lld tb %lld cellrest %lld | waits that really polled: %lld, mean ns from publish to seen: %lld\n", sm.acc[0], sm.acc[1], sm.acc[2], sm.acc[3], sm.acc[4], sm.acc[5], g_it[6], g_it[6] ? g_it[7] / g_it[6] : 0),
        printf("tb<16> cumulative cycles: fwd1 %lld fwd2q %lld sbh %lld inv %lld stat %lld store %lld\n", g_tbt[0], g_tbt[1], g_tbt[2], g_tbt[3], g_tbt[4], g_tbt[5]);
#endif
}

void ks_launch_recon_intra(const KsPicParams &pp, KsPlanes src, KsPlanes rec, KsLevels lv, ks_cell *cells, int *sync_ws, const int *n_intra, void *modes_ws, cudaStream_t st)
{
    KsIntraModes *modes = reinterpret_cast<KsIntraModes *>(modes_ws);
    /* the mode search of every (flagged) cell, independent of the reconstruction; then the dependent prediction + residual pass */
    ks_intra_modes_kernel<<<pp.cw * pp.ch, KS_MODES_WARPS * KS_WARP, 0, st>>>(pp, src.p[0], cells, n_intra, modes);
    cudaMemsetAsync(sync_ws, 0, sizeof(int) * (size_t)(1 + pp.cw * pp.ch), st);
    ks_recon_intra_rows_kernel<<<pp.ch, KS_INTRA_WARPS * KS_WARP, 0, st>>>(pp, src, rec, lv, cells, sync_ws, n_intra, modes);
}
size_t ks_intra_workspace_bytes(int ncell) { return (size_t)ncell * sizeof(KsIntraModes); }
