/*
 * ks_decide.cuh -- CU quadtree + merge/skip-friendly motion decision of a P picture (reference: processTree
 * E@0x46b610, checkInterPu2Nx2N, skipFullMergeDecision E@0x47f720, GetMergeCandsForP -- closed code, so this is OUR algorithm).
 *
 * The motion search (ks_me_kernel) works per 16x16 cell against a temporal predictor, so its field is spatially noisy and every cell would pay
 * an mvd.  This kernel makes the field coherent without giving up picture-level parallelism: one CTA per CTU,
 *   stage E  (8 warps, two cells each)  one candidate list per CTU = the search results of its own cells (z-order), zero, the cells bordering
 *            it on the left / above (distinct vectors, <= KS_NCAND); every cell measures the search metric of EVERY list entry with the real
 *            8-tap interpolation (one 40x52 window serves all candidates that land inside it);
 *   stage D  (warp 0, lane = candidate)  64 -> 32 -> 16 quadtree in coding order by J = distortion + lambda * bits, bits from the 2Nx2N merge
 *            list / AMVP predictors of the vectors decided so far (cells of other CTUs count with their search results: Jacobi across
 *            CTUs, Gauss-Seidel inside one);
 *   stage F  (8 warps)  cells whose vector changed get their luma + chroma prediction rewritten.
 * Bit-exact mirror of oracle/ora_frame.c: decide_candidates / decide_ctu.
 */
#pragma once
#include "ks_me.cuh"

#define KS_NCAND 16
#define KS_DECIDE_WARPS 8            /* two cells per warp: 4 CTAs per SM instead of 2, so the serial stage D of one CTU idles 7 warps, not 15 */

struct KsDecideSmem {
    KsWarpScratch sc[KS_DECIDE_WARPS];
    uint8_t  cwin[KS_DECIDE_WARPS][144];
    int      n;
    int16_t  cmx[KS_NCAND], cmy[KS_NCAND];
    int      dist[16][KS_NCAND];          /* [j * 4 + i][k] */
    int16_t  smx[6][6], smy[6][6];        /* stage D state: vectors of the CTU's cells + a one-cell border, index [j + 1][i + 1] */
    uint8_t  sok[6][6];
    uint8_t  slog2[16];
};

/* == ora mvd_bits_est: 1, 3, then 2*floor(log2 a) + 3 */
__device__ __forceinline__ int ks_mvd_bits_est(int d) { const int a = abs(d); return a == 0 ? 1 : (a == 1 ? 3 : 2 * (31 - __clz(a)) + 3); }
__device__ __forceinline__ int ks_zcell(int i, int j) { return (i & 1) | ((j & 1) << 1) | ((i & 2) << 1) | ((j & 2) << 2); }
__device__ __forceinline__ bool ks_nb_ok(const KsDecideSmem *sm, int i, int j, int zcur)
{
    if (i < -1 || j < -1 || i > 4 || j > 3) return false;
    if (!sm->sok[j + 1][i + 1]) return false;
    if (j == -1 || i == -1) return true;
    if (i > 3) return false;
    return ks_zcell(i, j) < zcur;
}
/* == ora motion_bits */
__device__ __forceinline__ int ks_motion_bits(const KsDecideSmem *sm, int i, int j, int s, int maxc, int mx, int my)
{
    const int zc = ks_zcell(i, j);
    const int ci[5] = {i - 1, i + s - 1, i + s, i - 1, i - 1}, cj[5] = {j + s - 1, j - 1, j - 1, j + s, j - 1};   /* A1 B1 B0 A0 B2 */
    bool av[5]; int vx[5], vy[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        av[k] = ks_nb_ok(sm, ci[k], cj[k], zc);
        vx[k] = av[k] ? sm->smx[cj[k] + 1][ci[k] + 1] : 0; vy[k] = av[k] ? sm->smy[cj[k] + 1][ci[k] + 1] : 0;
    }
#define KS_SAME(a, b) (vx[a] == vx[b] && vy[a] == vy[b])
    bool use[5] = {av[0], av[1] && !(av[0] && KS_SAME(0, 1)), av[2] && !(av[1] && KS_SAME(1, 2)), av[3] && !(av[0] && KS_SAME(0, 3)),
                   av[4] && !(av[0] && KS_SAME(0, 4)) && !(av[1] && KS_SAME(1, 4))};
    if (use[0] && use[1] && use[2] && use[3]) use[4] = false;
#undef KS_SAME
    int n = 0, idx = -1;
#pragma unroll
    for (int k = 0; k < 5; k++) if (n < maxc && idx < 0 && use[k]) { if (vx[k] == mx && vy[k] == my) idx = n; n++; }
    if (idx < 0 && n < maxc && mx == 0 && my == 0) idx = n;
    if (idx >= 0) return 1 + (maxc > 1 ? (idx < maxc - 1 ? idx + 1 : maxc - 1) : 0);
    const int fa = av[3] ? 3 : (av[0] ? 0 : -1), fb = av[2] ? 2 : (av[1] ? 1 : (av[4] ? 4 : -1));
    int best = 0x7fffffff, ax = 0, ay = 0, bx = 0, by = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) { if (k == fa) { ax = vx[k]; ay = vy[k]; } if (k == fb) { bx = vx[k]; by = vy[k]; } }
    if (fa >= 0) best = min(best, ks_mvd_bits_est(mx - ax) + ks_mvd_bits_est(my - ay));
    if (fb >= 0) best = min(best, ks_mvd_bits_est(mx - bx) + ks_mvd_bits_est(my - by));
    if (fa < 0 || fb < 0 || (ax == bx && ay == by)) best = min(best, ks_mvd_bits_est(mx) + ks_mvd_bits_est(my));
    return 5 + best;
}

/* warp 0: best candidate for the block of s x s cells at CTU-local (i,j): returns (cost << 4) | k */
__device__ __forceinline__ unsigned ks_decide_eval(const KsDecideSmem *sm, int i, int j, int s, int lam, int maxc, int lane)
{
    unsigned key = 0xffffffffu;
    if (lane < sm->n) {
        int sum = 0;
        for (int b = 0; b < s; b++) for (int a = 0; a < s; a++) sum += sm->dist[(j + b) * 4 + i + a][lane];
        const int c = sum + ((lam * (ks_motion_bits(sm, i, j, s, maxc, sm->cmx[lane], sm->cmy[lane]) + (s > 1 ? 1 : 0))) >> 4);
        key = ((unsigned)c << 4) | (unsigned)lane;
    }
    return __reduce_min_sync(0xffffffffu, key);
}
__device__ __forceinline__ void ks_decide_commit(KsDecideSmem *sm, int i, int j, int s, int k, int lane)
{
    if (lane < s * s) {
        const int a = lane % s, b = lane / s;
        sm->smx[j + b + 1][i + a + 1] = sm->cmx[k]; sm->smy[j + b + 1][i + a + 1] = sm->cmy[k]; sm->sok[j + b + 1][i + a + 1] = 1;
        sm->slog2[(j + b) * 4 + i + a] = (uint8_t)(s == 4 ? 6 : (s == 2 ? 5 : 4));
    }
    __syncwarp();
}

__global__ void __launch_bounds__(KS_DECIDE_WARPS * KS_WARP, 4)
ks_decide_kernel(KsPicParams pp, const uint8_t *__restrict__ srcY, KsPlanes ref, const ks_cell *__restrict__ mv0, const int *__restrict__ dist0,
                 ks_cell *__restrict__ cells, KsPlanes pred)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KsDecideSmem *sm = reinterpret_cast<KsDecideSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int X = blockIdx.x << 2, Y = blockIdx.y << 2, cw = pp.cw, ch = pp.ch;
    const int W = pp.W, H = pp.H;
    const int maxc = 3;
    /* ---- candidate list (warp 0): 28 sources in a fixed order, first occurrences kept, at most KS_NCAND ---- */
    if (warp == 0) {
        int sx, sy; bool zero = false;
        if (lane < 16) { sx = X + ((lane & 1) | ((lane >> 1) & 2)); sy = Y + (((lane >> 1) & 1) | ((lane >> 2) & 2)); }
        else if (lane == 16) { sx = sy = 0; zero = true; }
        else if (lane < 21) { sx = X - 1; sy = Y + 3 - (lane - 17); }
        else if (lane < 25) { sx = X + lane - 21; sy = Y - 1; }
        else if (lane == 25) { sx = X + 4; sy = Y - 1; }
        else if (lane == 26) { sx = X - 1; sy = Y + 4; }
        else { sx = X - 1; sy = Y - 1; }
        const bool valid = lane < 28 && (zero || (sx >= 0 && sy >= 0 && sx < cw && sy < ch));
        uint32_t mvw = 0;
        if (valid && !zero) { const ks_cell c = mv0[sy * cw + sx]; mvw = (uint32_t)(uint16_t)c.mvx | ((uint32_t)(uint16_t)c.mvy << 16); }
        bool first = valid;
#pragma unroll 4
        for (int j = 0; j < 27; j++) {
            const uint32_t o = __shfl_sync(0xffffffffu, mvw, j); const bool ov = __shfl_sync(0xffffffffu, (int)valid, j) != 0;
            if (j < lane && ov && o == mvw) first = false;
        }
        const unsigned fb = __ballot_sync(0xffffffffu, first);
        const int rank = __popc(fb & ((1u << lane) - 1u));
        if (first && rank < KS_NCAND) { sm->cmx[rank] = (int16_t)(mvw & 0xffffu); sm->cmy[rank] = (int16_t)(mvw >> 16); }
        if (lane == 0) sm->n = min(__popc(fb), KS_NCAND);
        /* stage D state: border cells carry their search results, the CTU's own cells are undecided */
        for (int e = lane; e < 36; e += 32) {
            const int i = e % 6 - 1, j = e / 6 - 1, cx = X + i, cy = Y + j;
            const bool in = cx >= 0 && cy >= 0 && cx < cw && cy < ch && (j == -1 || (i == -1 && j <= 3));
            int16_t vx = 0, vy = 0;
            if (in) { const ks_cell c = mv0[cy * cw + cx]; vx = c.mvx; vy = c.mvy; }
            sm->smx[j + 1][i + 1] = vx; sm->smy[j + 1][i + 1] = vy; sm->sok[j + 1][i + 1] = in;
        }
    }
    __syncthreads();
    /* ---- stage E: this warp's cells against every candidate ---- */
    KsWarpScratch *sc = &sm->sc[warp];
    const bool satd = pp.satd && pp.subpel > 0;
#pragma unroll 1
    for (int cell = warp; cell < 16; cell += KS_DECIDE_WARPS) {
        const int ci = cell & 3, cj = cell >> 2, cx = X + ci, cy = Y + cj, x0 = cx << 4, y0 = cy << 4;
        if (cx >= cw || cy >= ch) continue;
        const ks_cell own = mv0[cy * cw + cx];
        const int d0 = dist0[cy * cw + cx];
        const uint2 s = *reinterpret_cast<const uint2 *>(srcY + (size_t)(y0 + (lane >> 1)) * W + x0 + 8 * (lane & 1));
        int wx0, wy0;
        ks_center_window(x0, y0, own.mvx >> 2, own.mvy >> 2, wx0, wy0);
        __syncwarp();
        ks_load_window(sc->win, ref.p[0], W, H, wx0, wy0, lane);
        const int n = sm->n;
#pragma unroll 1
        for (int k = 0; k < n; k++) {
            const int mx = sm->cmx[k], my = sm->cmy[k];
            int d;
            if (mx == own.mvx && my == own.mvy) d = d0;
            else {
                int bxw = x0 + (mx >> 2) - wx0, byw = y0 + (my >> 2) - wy0;
                if (bxw < 4 || bxw > 27 || byw < 4 || byw > 19) {
                    ks_center_window(x0, y0, mx >> 2, my >> 2, wx0, wy0);
                    __syncwarp();
                    ks_load_window(sc->win, ref.p[0], W, H, wx0, wy0, lane);
                    bxw = x0 + (mx >> 2) - wx0; byw = y0 + (my >> 2) - wy0;
                }
                uint32_t o0, o1;
                ks_interp16(sc, bxw, byw, mx & 3, my & 3, lane, o0, o1);
                d = satd ? (int)ks_satd16(o0, o1, s.x, s.y, lane) : (int)ks_warp_sum(__vsadu4(o0, s.x) + __vsadu4(o1, s.y));
            }
            if (lane == 0) sm->dist[cell][k] = d;
        }
    }
    __syncthreads();
    /* ---- stage D (warp 0, lane = candidate): children first, then the whole block; the whole block wins ties ---- */
    if (warp == 0) {
        const int ncx = min(4, cw - X), ncy = min(4, ch - Y), lam = pp.lambda_dec_q4;
        const bool in64 = ncx == 4 && ncy == 4;
        int j64 = in64 ? (lam >> 4) : 0;
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
            const int qi = (q & 1) * 2, qj = (q >> 1) * 2;
            if (qi >= ncx || qj >= ncy) continue;
            const bool in32 = qi + 2 <= ncx && qj + 2 <= ncy;
            int j32 = in32 ? (lam >> 4) : 0;
#pragma unroll 1
            for (int c = 0; c < 4; c++) {
                const int i = qi + (c & 1), j = qj + (c >> 1);
                if (i >= ncx || j >= ncy) continue;
                const unsigned key = ks_decide_eval(sm, i, j, 1, lam, maxc, lane);
                ks_decide_commit(sm, i, j, 1, (int)(key & 15u), lane);
                j32 += (int)(key >> 4);
            }
            if (in32) {
                const unsigned key = ks_decide_eval(sm, qi, qj, 2, lam, maxc, lane);
                if ((int)(key >> 4) <= j32) { ks_decide_commit(sm, qi, qj, 2, (int)(key & 15u), lane); j32 = (int)(key >> 4); }
            }
            j64 += j32;
        }
        if (in64) {
            const unsigned key = ks_decide_eval(sm, 0, 0, 4, lam, maxc, lane);
            if ((int)(key >> 4) <= j64) ks_decide_commit(sm, 0, 0, 4, (int)(key & 15u), lane);
        }
    }
    __syncthreads();
    /* ---- final cells + stage F: re-predict the cells whose vector changed ---- */
#pragma unroll 1
    for (int cell = warp; cell < 16; cell += KS_DECIDE_WARPS) {
        const int ci = cell & 3, cj = cell >> 2, cx = X + ci, cy = Y + cj, x0 = cx << 4, y0 = cy << 4;
        if (cx >= cw || cy >= ch) continue;
        const ks_cell own = mv0[cy * cw + cx];
        const int fmx = sm->smx[cj + 1][ci + 1], fmy = sm->smy[cj + 1][ci + 1];
        if (lane == 0) {
            ks_cell c; c.mvx = (int16_t)fmx; c.mvy = (int16_t)fmy; c.cu_log2 = sm->slog2[cell]; c.flags = 0; c.intra_mode = 0; c.rsv = 0;
            cells[cy * cw + cx] = c;
        }
        if (fmx == own.mvx && fmy == own.mvy) continue;
        int wx0, wy0;
        ks_center_window(x0, y0, fmx >> 2, fmy >> 2, wx0, wy0);
        __syncwarp();
        ks_load_window(sc->win, ref.p[0], W, H, wx0, wy0, lane);
        uint32_t o0, o1;
        ks_interp16(sc, x0 + (fmx >> 2) - wx0, y0 + (fmy >> 2) - wy0, fmx & 3, fmy & 3, lane, o0, o1);
        *reinterpret_cast<uint2 *>(pred.p[0] + (size_t)(y0 + (lane >> 1)) * W + x0 + 8 * (lane & 1)) = make_uint2(o0, o1);
        const int CW = W >> 1, CH = H >> 1;
#pragma unroll 1
        for (int c = 0; c < 2; c++)
            ks_mc_chroma8(sm->cwin[warp], &sc->tmp[0][0], ref.p[1 + c], CW, CH, x0 >> 1, y0 >> 1, fmx, fmy,
                          pred.p[1 + c] + (size_t)(y0 >> 1) * CW + (x0 >> 1), CW, lane);
    }
}

void ks_launch_decide(const KsPicParams &pp, const uint8_t *srcY, KsPlanes ref, const ks_cell *mv0, const int *dist0, ks_cell *cells, KsPlanes pred, cudaStream_t st)
{
    dim3 grid(pp.ctw, pp.cth);
    ks_decide_kernel<<<grid, KS_DECIDE_WARPS * KS_WARP, sizeof(KsDecideSmem), st>>>(pp, srcY, ref, mv0, dist0, cells, pred);
}
