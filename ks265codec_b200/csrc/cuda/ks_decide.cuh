/*
 * ks_decide.cuh -- CU quadtree + merge/skip-friendly motion decision of a P picture (reference: processTree
 * E@0x46b610, checkInterPu2Nx2N, skipFullMergeDecision E@0x47f720, GetMergeCandsForP -- closed code, so this is OUR algorithm).
 *
 * The motion search (ks_me_kernel) works per 16x16 cell against a temporal predictor, so its field is spatially noisy and every cell would pay
 * an mvd.  These kernels make the field coherent without giving up picture-level parallelism:
 *   stage E  (ks_decide_cand_kernel: one CTA per CTU, 8 warps, two cells each)  one candidate list per CTU = the search results of its own cells (z-order), zero, the cells bordering
 *            it on the left / above (distinct vectors, <= KS_NCAND); every cell measures the search metric of EVERY list entry with the real
 *            8-tap interpolation (one 40x52 window serves all candidates that land inside it);
 *   stage D  (ks_decide_tree_kernel: one WARP per CTU, lane = candidate -- fused into stage E it idled 7 of 8 warps at a barrier: 13 stall cycles
 *            per issue in ncu, profiles/)  64 -> 32 -> 16 quadtree in coding order by J = distortion + lambda * bits, bits from the 2Nx2N merge
 *            list / AMVP predictors of the vectors decided so far (cells of other CTUs count with their search results: Jacobi across
 *            CTUs, Gauss-Seidel inside one);
 *   stage F  (ks_decide_pred_kernel: one warp per cell)  cells whose vector changed get their luma + chroma prediction rewritten.
 * Bit-exact mirror of oracle/ora_frame.c: decide_candidates / decide_ctu.
 */
#pragma once
#include "ks_me.cuh"

#define KS_NCAND 16
#define KS_INTRA_HDR_BITS 10
#define KS_INTRA_FLOOR 64
#define KS_DECIDE_WARPS 8            /* two cells per warp: 4 CTAs per SM instead of 2, so the serial stage D of one CTU idles 7 warps, not 15 */

/* per-CTU result of stage E (global memory, 1.1 KB): the candidate list and every cell's distortion for every entry */
struct KsCtuCands {
    int      n;
    int16_t  cmx[KS_NCAND], cmy[KS_NCAND];
    int      dist[16][KS_NCAND];          /* [j * 4 + i][k] */
    int      intra[16];                   /* intra estimate of each cell (ks_intra_estimate) */
};
/* stage D working set of one CTU (one warp) */
struct KsDecideSmem {
    int      n;
    int16_t  cmx[KS_NCAND], cmy[KS_NCAND];
    int      dist[16][KS_NCAND];
    uint32_t smv[36];                     /* vectors (x | y << 16) of the CTU's cells + a one-cell border, index (j + 1) * 6 + i + 1 */
    uint8_t  sok[36];
    uint8_t  slog2[16], sintra[16];
    int      intra[16];
};
struct KsCandSmem {
    KsWarpScratch sc[KS_DECIDE_WARPS];
    int      n;
    int16_t  cmx[KS_NCAND], cmy[KS_NCAND];
};

/* == ora mvd_bits_est: 1, 3, then 2*floor(log2 a) + 3 */
__device__ __forceinline__ int ks_mvd_bits_est(int d) { const int a = abs(d); return a == 0 ? 1 : (a == 1 ? 3 : 2 * (31 - __clz(a)) + 3); }
__device__ __forceinline__ int ks_zcell(int i, int j) { return (i & 1) | ((j & 1) << 1) | ((i & 2) << 1) | ((j & 2) << 2); }
__device__ __forceinline__ bool ks_nb_ok(const KsDecideSmem *sm, int i, int j, int zcur)
{
    if (i < -1 || j < -1 || i > 4 || j > 3) return false;
    if (!sm->sok[(j + 1) * 6 + i + 1]) return false;
    if (j == -1 || i == -1) return true;
    if (i > 3) return false;
    return ks_zcell(i, j) < zcur;
}
/* == ora motion_bits; vectors travel as one word (x | y << 16) */
__device__ __forceinline__ int ks_mvw_bits(uint32_t m, uint32_t p)
{
    return ks_mvd_bits_est((int)(short)(m & 0xffffu) - (int)(short)(p & 0xffffu)) + ks_mvd_bits_est((int)(short)(m >> 16) - (int)(short)(p >> 16));
}
__device__ __forceinline__ int ks_motion_bits(const KsDecideSmem *sm, int i, int j, int s, int maxc, uint32_t m)
{
    const int zc = ks_zcell(i, j);
    /* A1 B1 B0 A0 B2 */
    const bool a1 = ks_nb_ok(sm, i - 1, j + s - 1, zc), b1 = ks_nb_ok(sm, i + s - 1, j - 1, zc), b0 = ks_nb_ok(sm, i + s, j - 1, zc),
               a0 = ks_nb_ok(sm, i - 1, j + s, zc), b2 = ks_nb_ok(sm, i - 1, j - 1, zc);
    const uint32_t va1 = a1 ? sm->smv[(j + s) * 6 + i] : 0u, vb1 = b1 ? sm->smv[j * 6 + i + s] : 0u, vb0 = b0 ? sm->smv[j * 6 + i + s + 1] : 0u,
                   va0 = a0 ? sm->smv[(j + s + 1) * 6 + i] : 0u, vb2 = b2 ? sm->smv[j * 6 + i] : 0u;
    const bool u0 = a1, u1 = b1 && !(a1 && va1 == vb1), u2 = b0 && !(b1 && vb1 == vb0), u3 = a0 && !(a1 && va1 == va0);
    const bool u4 = b2 && !(a1 && va1 == vb2) && !(b1 && vb1 == vb2) && !(u0 && u1 && u2 && u3);
    int n = 0, idx = -1;
    if (u0) { if (va1 == m) idx = n; n++; }
    if (u1 && n < maxc && idx < 0) { if (vb1 == m) idx = n; n++; }
    if (u2 && n < maxc && idx < 0) { if (vb0 == m) idx = n; n++; }
    if (u3 && n < maxc && idx < 0) { if (va0 == m) idx = n; n++; }
    if (u4 && n < maxc && idx < 0) { if (vb2 == m) idx = n; n++; }
    if (idx < 0 && n < maxc && m == 0u) idx = n;
    if (idx >= 0) return 1 + (maxc > 1 ? (idx < maxc - 1 ? idx + 1 : maxc - 1) : 0);
    /* AMVP: a = first of (A0, A1), b = first of (B0, B1, B2) */
    const bool fa = a0 || a1, fb = b0 || b1 || b2;
    const uint32_t pa = a0 ? va0 : va1, pb = b0 ? vb0 : (b1 ? vb1 : vb2);
    int best = 0x7fffffff;
    if (fa) best = min(best, ks_mvw_bits(m, pa));
    if (fb) best = min(best, ks_mvw_bits(m, pb));
    if (!fa || !fb || pa == pb) best = min(best, ks_mvw_bits(m, 0u));
    return 5 + best;
}

/* warp 0: best candidate for the block of s x s cells at CTU-local (i,j): returns (cost << 4) | k */
__device__ __forceinline__ unsigned ks_decide_eval(const KsDecideSmem *sm, int i, int j, int s, int lam, int maxc, int lane)
{
    unsigned key = 0xffffffffu;
    if (lane < sm->n) {
        int sum = 0;
        for (int b = 0; b < s; b++) for (int a = 0; a < s; a++) sum += sm->dist[(j + b) * 4 + i + a][lane];
        const uint32_t m = (uint32_t)(uint16_t)sm->cmx[lane] | ((uint32_t)(uint16_t)sm->cmy[lane] << 16);
        const int c = sum + ((lam * (ks_motion_bits(sm, i, j, s, maxc, m) + (s > 1 ? 1 : 0))) >> 4);
        key = ((unsigned)c << 4) | (unsigned)lane;
    }
    return __reduce_min_sync(0xffffffffu, key);
}
__device__ __forceinline__ void ks_decide_commit(KsDecideSmem *sm, int i, int j, int s, int k, int lane)
{
    if (lane < s * s) {
        const int a = lane % s, b = lane / s;
        sm->smv[(j + b + 1) * 6 + i + a + 1] = (uint32_t)(uint16_t)sm->cmx[k] | ((uint32_t)(uint16_t)sm->cmy[k] << 16); sm->sok[(j + b + 1) * 6 + i + a + 1] = 1;
        sm->slog2[(j + b) * 4 + i + a] = (uint8_t)(s == 4 ? 6 : (s == 2 ? 5 : 4)); sm->sintra[(j + b) * 4 + i + a] = 0;
    }
    __syncwarp();
}

/* intra ESTIMATE of a 16x16 cell in the search metric (== ora intra_estimate): best of DC / horizontal / vertical / planar predicted from the
 * SOURCE picture's neighbours (128 where the picture ends, no boundary smoothing).  It only decides inter vs intra in stage D; the real 35-mode
 * search runs on reconstructed neighbours (ks_recon_intra_kernel, masked mode).  Warp-collective; lane = row lane>>1, columns 8*(lane&1)..+7. */
__device__ __forceinline__ int ks_intra_estimate(const uint8_t *__restrict__ srcY, int W, int H, int x0, int y0, uint2 s, bool satd, int lane)
{
    const int row = lane >> 1, half = lane & 1;
    const bool hl = x0 > 0, ht = y0 > 0;
    const int l = hl ? (int)srcY[(size_t)(y0 + row) * W + x0 - 1] : 128;
    uint32_t t0 = 0x80808080u, t1 = 0x80808080u;
    if (ht) { const uint2 t = *reinterpret_cast<const uint2 *>(srcY + (size_t)(y0 - 1) * W + x0 + 8 * half); t0 = t.x; t1 = t.y; }
    const int tsum = (int)(__vsadu4(t0, 0u) + __vsadu4(t1, 0u));
    const int sl = (int)ks_warp_sum(half == 0 ? (unsigned)l : 0u), stp = (int)ks_warp_sum(row == 0 ? (unsigned)tsum : 0u);
    const int tr = ht ? (int)srcY[(size_t)(y0 - 1) * W + min(x0 + 16, W - 1)] : 128, bl = hl ? (int)srcY[(size_t)min(y0 + 16, H - 1) * W + x0 - 1] : 128;
    const int dc = (hl && ht) ? (sl + stp + 16) >> 5 : (hl ? (sl + 8) >> 4 : (ht ? (stp + 8) >> 4 : 128));
    int best = 0x7fffffff;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        uint32_t o0, o1;
        if (m == 0) o0 = o1 = (uint32_t)dc * 0x01010101u;
        else if (m == 1) o0 = o1 = (uint32_t)l * 0x01010101u;
        else if (m == 2) { o0 = t0; o1 = t1; }
        else {
            o0 = o1 = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int x = 8 * half + k, t = (int)(((k < 4 ? t0 : t1) >> (8 * (k & 3))) & 255u);
                const uint32_t v = (uint32_t)(((15 - x) * l + (x + 1) * tr + (15 - row) * t + (row + 1) * bl + 16) >> 5);
                if (k < 4) o0 |= v << (8 * k); else o1 |= v << (8 * (k - 4));
            }
        }
        const int c = satd ? (int)ks_satd16(o0, o1, s.x, s.y, lane) : (int)ks_warp_sum(__vsadu4(o0, s.x) + __vsadu4(o1, s.y));
        best = min(best, c);
    }
    return best;
}

/* ---- stage E: one CTA per CTU, 8 warps x 2 cells ---- */
__global__ void __launch_bounds__(KS_DECIDE_WARPS * KS_WARP, 4)
ks_decide_cand_kernel(KsPicParams pp, const uint8_t *__restrict__ srcY, KsPlanes ref, const ks_cell *__restrict__ mv0, const int *__restrict__ dist0,
                      KsCtuCands *__restrict__ cands)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KsCandSmem *sm = reinterpret_cast<KsCandSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int X = blockIdx.x << 2, Y = blockIdx.y << 2, cw = pp.cw, ch = pp.ch;
    const int W = pp.W, H = pp.H;
    KsCtuCands *out = &cands[blockIdx.y * pp.ctw + blockIdx.x];
    /* candidate list (warp 0): 28 sources in a fixed order, first occurrences kept, at most KS_NCAND */
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
        if (first && rank < KS_NCAND) {
            sm->cmx[rank] = (int16_t)(mvw & 0xffffu); sm->cmy[rank] = (int16_t)(mvw >> 16);
            out->cmx[rank] = (int16_t)(mvw & 0xffffu); out->cmy[rank] = (int16_t)(mvw >> 16);
        }
        if (lane == 0) { sm->n = min(__popc(fb), KS_NCAND); out->n = sm->n; }
    }
    __syncthreads();
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
        int mine = 0;                          /* lane k keeps the distortion of candidate k: one coalesced store per cell */
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
            if (lane == k) mine = d;
        }
        if (lane < n) out->dist[cell][lane] = mine;
        const int ie = ks_intra_estimate(srcY, W, H, x0, y0, s, satd, lane);
        if (lane == 0) out->intra[cell] = ie;
    }
}

/* ---- stage D: one WARP per CTU (lane = candidate); children first, then the whole block; the whole block wins ties ---- */
#define KS_TREE_WARPS 4
__global__ void __launch_bounds__(KS_TREE_WARPS * KS_WARP)
ks_decide_tree_kernel(KsPicParams pp, const ks_cell *__restrict__ mv0, const KsCtuCands *__restrict__ cands, ks_cell *__restrict__ cells, int *__restrict__ n_intra)
{
    __shared__ __align__(16) KsDecideSmem smem[KS_TREE_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctu = blockIdx.x * KS_TREE_WARPS + warp;
    if (ctu >= pp.ctw * pp.cth) return;
    KsDecideSmem *sm = &smem[warp];
    const int X = (ctu % pp.ctw) << 2, Y = (ctu / pp.ctw) << 2, cw = pp.cw, ch = pp.ch;
    const int maxc = 3;
    {
        const KsCtuCands *in = &cands[ctu];
        const int n = in->n;
        if (lane == 0) sm->n = n;
        if (lane < n) { sm->cmx[lane] = in->cmx[lane]; sm->cmy[lane] = in->cmy[lane]; }
        if (lane < 16) { sm->intra[lane] = in->intra[lane]; sm->sintra[lane] = 0; }
        for (int e = lane; e < 16 * KS_NCAND; e += 32) sm->dist[e / KS_NCAND][e % KS_NCAND] = (e % KS_NCAND) < n ? in->dist[e / KS_NCAND][e % KS_NCAND] : 0;
        /* border cells carry their search results, the CTU's own cells are undecided */
        for (int e = lane; e < 36; e += 32) {
            const int i = e % 6 - 1, j = e / 6 - 1, cx = X + i, cy = Y + j;
            const bool in_ = cx >= 0 && cy >= 0 && cx < cw && cy < ch && (j == -1 || (i == -1 && j <= 3));
            uint32_t v = 0;
            if (in_) { const ks_cell c = mv0[cy * cw + cx]; v = (uint32_t)(uint16_t)c.mvx | ((uint32_t)(uint16_t)c.mvy << 16); }
            sm->smv[e] = v; sm->sok[e] = in_;
        }
    }
    __syncwarp();
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
            /* intra 16x16 CU (about KS_INTRA_HDR_BITS of header) against the best vector */
            const int ji = sm->intra[j * 4 + i] + ((lam * KS_INTRA_HDR_BITS) >> 4);
            /* ...and only above the quantisation-noise floor (see ora decide_block) */
            if (ji < (int)(key >> 4) && sm->dist[j * 4 + i][key & 15u] > ((KS_INTRA_FLOOR * lam) >> 4)) {
                if (lane == 0) { sm->smv[(j + 1) * 6 + i + 1] = 0u; sm->sok[(j + 1) * 6 + i + 1] = 0; sm->slog2[j * 4 + i] = 4; sm->sintra[j * 4 + i] = 1; }
                __syncwarp();
                j32 += ji;
            } else {
                ks_decide_commit(sm, i, j, 1, (int)(key & 15u), lane);
                j32 += (int)(key >> 4);
            }
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
    __syncwarp();
    if (lane < 16) {
        const int ci = lane & 3, cj = lane >> 2;
        if (ci < ncx && cj < ncy) {
            const uint32_t v = sm->smv[(cj + 1) * 6 + ci + 1];
            ks_cell c; c.mvx = (int16_t)(v & 0xffffu); c.mvy = (int16_t)(v >> 16); c.cu_log2 = sm->slog2[lane]; c.flags = sm->sintra[lane] ? KS_F_INTRA : 0; c.intra_mode = 0; c.rsv = 0;
            cells[(Y + cj) * cw + X + ci] = c;
        }
    }
    const unsigned ib = __ballot_sync(0xffffffffu, lane < 16 && (lane & 3) < ncx && (lane >> 2) < ncy && sm->sintra[lane]);
    if (ib && lane == 0) atomicAdd(n_intra, __popc(ib));
}

/* ---- stage F: one warp per cell; cells whose vector changed get their luma + chroma prediction rewritten ---- */
__global__ void __launch_bounds__(KS_ME_WARPS * KS_WARP, 4)
ks_decide_pred_kernel(KsPicParams pp, KsPlanes ref, const ks_cell *__restrict__ mv0, const ks_cell *__restrict__ cells, KsPlanes pred)
{
    __shared__ __align__(16) KsWarpScratch scratch[KS_ME_WARPS];
    __shared__ uint8_t cwins[KS_ME_WARPS][144];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * KS_ME_WARPS + warp;
    if (cell >= pp.cw * pp.ch) return;
    const ks_cell own = mv0[cell], fin = cells[cell];
    if ((fin.flags & KS_F_INTRA) || (fin.mvx == own.mvx && fin.mvy == own.mvy)) return;
    KsWarpScratch *sc = &scratch[warp];
    const int cyc = cell / pp.cw, cxc = cell - cyc * pp.cw, x0 = cxc << 4, y0 = cyc << 4, W = pp.W, H = pp.H;
    const int fmx = fin.mvx, fmy = fin.mvy;
    int wx0, wy0;
    ks_center_window(x0, y0, fmx >> 2, fmy >> 2, wx0, wy0);
    ks_load_window(sc->win, ref.p[0], W, H, wx0, wy0, lane);
    uint32_t o0, o1;
    ks_interp16(sc, x0 + (fmx >> 2) - wx0, y0 + (fmy >> 2) - wy0, fmx & 3, fmy & 3, lane, o0, o1);
    *reinterpret_cast<uint2 *>(pred.p[0] + (size_t)(y0 + (lane >> 1)) * W + x0 + 8 * (lane & 1)) = make_uint2(o0, o1);
    const int CW = W >> 1, CH = H >> 1;
#pragma unroll 1
    for (int c = 0; c < 2; c++)
        ks_mc_chroma8(cwins[warp], &sc->tmp[0][0], ref.p[1 + c], CW, CH, x0 >> 1, y0 >> 1, fmx, fmy,
                      pred.p[1 + c] + (size_t)(y0 >> 1) * CW + (x0 >> 1), CW, lane);
}

void ks_launch_decide(const KsPicParams &pp, const uint8_t *srcY, KsPlanes ref, const ks_cell *mv0, const int *dist0, void *cands_ws, ks_cell *cells, KsPlanes pred, int *n_intra, cudaStream_t st)
{
    KsCtuCands *cands = reinterpret_cast<KsCtuCands *>(cands_ws);
    const int nctu = pp.ctw * pp.cth, ncell = pp.cw * pp.ch;
    cudaMemsetAsync(n_intra, 0, sizeof(int), st);
    ks_decide_cand_kernel<<<dim3(pp.ctw, pp.cth), KS_DECIDE_WARPS * KS_WARP, sizeof(KsCandSmem), st>>>(pp, srcY, ref, mv0, dist0, cands);
    ks_decide_tree_kernel<<<(nctu + KS_TREE_WARPS - 1) / KS_TREE_WARPS, KS_TREE_WARPS * KS_WARP, 0, st>>>(pp, mv0, cands, cells, n_intra);
    ks_decide_pred_kernel<<<(ncell + KS_ME_WARPS - 1) / KS_ME_WARPS, KS_ME_WARPS * KS_WARP, 0, st>>>(pp, ref, mv0, cells, pred);
}
size_t ks_decide_workspace_bytes(int nctu) { return (size_t)nctu * sizeof(KsCtuCands); }
