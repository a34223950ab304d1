/*
 * ks_loopfilter.cuh -- in-loop filters of the ks265 B200 hot path.
 *
 *  deblocking (spec 8.7.2; reference ctuDeblockFilterVer E@0x4144a0, CtuDeblockFilterHorT<> E@0x471d40, leaf
 *  EdgeFilterLuma{Ver,Hor}_c E@0x413100/0x4133f0, PixelFilterChroma*_c E@0x4137d0, Bs from CalcBsInterP E@0x413990):
 *  HEVC deblocking is picture-parallel by design -- all vertical edges first, then all horizontal edges on the
 *  result -- so it is two flat launches, one thread per 4-sample edge segment, in place.  Edges lie on the 16-sample grid,
 *  plus the 8-sample lines inside the cells that hold four 8x8 intra CUs (cu_log2 == 3).
 *
 *  SAO (spec 8.7.3; reference statSao*_c E@0x4a6370.., CEncSao::modeDecisionCtu E@0x4a9870, SaoApplyOffset*_c
 *  E@0x43dba0..): ONE launch, one CTA per CTU: the deblocked 64x64(+1 halo) tile is staged in shared memory once,
 *  statistics (per-warp packed histograms, the reference's (d<<12)|1 accumulator trick) -> offset RD decision ->
 *  apply -> write the final picture, plus the per-plane SSE for the PSNR line.  SAO of a CTU needs only its own
 *  parameters and deblocked neighbours, so stats/decide/apply need no global pass in between (the reference's
 *  1-CTU lag in CLoopFilterCtu::Execute E@0x492b40 disappears).
 * Bit-exact mirror of oracle/ora_frame.c (ora_deblock_picture, ora_sao_picture).
 */
#pragma once
#include "ks_common.cuh"

/* ------------------------------------------------------------------ deblocking ------------------- */
__device__ __forceinline__ bool ks_is_tu_edge(ks_cell p, ks_cell q, int xp, int yp, int xq, int yq, int pos)
{
    int sp = 1 << p.cu_log2, sq = 1 << q.cu_log2;
    bool same = p.cu_log2 == q.cu_log2 && (xp & ~(sp - 1)) == (xq & ~(sq - 1)) && (yp & ~(sp - 1)) == (yq & ~(sq - 1));
    if (!same) return true;
    return p.cu_log2 == 6 && (pos & 31) == 0;
}
__device__ __forceinline__ int ks_edge_bs(ks_cell p, ks_cell q, ks_cell_b pb, ks_cell_b qb)
{   /* spec 8.7.2.4; in B pictures the two lists always name different pictures, so motion compares list by list */
    if ((p.flags | q.flags) & KS_F_INTRA) return 2;
    if ((p.flags | q.flags) & KS_F_CBF_Y) return 1;
    if (pb.dir != qb.dir) return 1;
    if ((pb.dir & 1) && (abs(p.mvx - q.mvx) >= 4 || abs(p.mvy - q.mvy) >= 4)) return 1;
    if ((pb.dir & 2) && (abs(pb.mvx1 - qb.mvx1) >= 4 || abs(pb.mvy1 - qb.mvy1) >= 4)) return 1;
    return 0;
}
/* filter one 4-line segment held in registers: px[l][0..3] = p3..p0, px[l][4..7] = q0..q3 */
__device__ __forceinline__ bool ks_deblock_luma_regs(int (&px)[4][8], int beta, int tc)
{
#define P_(i, l) px[l][3 - (i)]
#define Q_(i, l) px[l][4 + (i)]
    int dp0 = abs(P_(2,0) - 2 * P_(1,0) + P_(0,0)), dp3 = abs(P_(2,3) - 2 * P_(1,3) + P_(0,3));
    int dq0 = abs(Q_(2,0) - 2 * Q_(1,0) + Q_(0,0)), dq3 = abs(Q_(2,3) - 2 * Q_(1,3) + Q_(0,3));
    int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = dpq0 + dpq3;
    if (d >= beta) return false;
    bool s0 = 2 * dpq0 < (beta >> 2) && abs(P_(3,0) - P_(0,0)) + abs(Q_(0,0) - Q_(3,0)) < (beta >> 3) && abs(P_(0,0) - Q_(0,0)) < ((5 * tc + 1) >> 1);
    bool s3 = 2 * dpq3 < (beta >> 2) && abs(P_(3,3) - P_(0,3)) + abs(Q_(0,3) - Q_(3,3)) < (beta >> 3) && abs(P_(0,3) - Q_(0,3)) < ((5 * tc + 1) >> 1);
    bool strong = s0 && s3;
    bool dep = dp < ((beta + (beta >> 1)) >> 3), deq = dq < ((beta + (beta >> 1)) >> 3);
#pragma unroll
    for (int l = 0; l < 4; l++) {
        int p0 = P_(0,l), p1 = P_(1,l), p2 = P_(2,l), p3 = P_(3,l), q0 = Q_(0,l), q1 = Q_(1,l), q2 = Q_(2,l), q3 = Q_(3,l);
        if (strong) {
            P_(0,l) = ks_clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            P_(1,l) = ks_clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
            P_(2,l) = ks_clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
            Q_(0,l) = ks_clip3(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            Q_(1,l) = ks_clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
            Q_(2,l) = ks_clip3(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
        } else {
            int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (abs(delta) < tc * 10) {
                delta = ks_clip3(-tc, tc, delta);
                P_(0,l) = ks_clip8(p0 + delta); Q_(0,l) = ks_clip8(q0 - delta);
                if (dep) P_(1,l) = ks_clip8(p1 + ks_clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
                if (deq) Q_(1,l) = ks_clip8(q1 + ks_clip3(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
            }
        }
    }
#undef P_
#undef Q_
    return true;
}

/* dir 0: vertical edges (x = 16,32,..), thread per (edge, 4-row segment); dir 1: horizontal edges */
template <int DIR>
__global__ void __launch_bounds__(256)
ks_deblock_kernel(KsPicParams pp, KsPlanes rec, const ks_cell *__restrict__ cells, const ks_cell_b *__restrict__ cells_b)
{
    const int W = pp.W, H = pp.H;
    const int nedge = ((DIR ? H : W) >> 3) - 1, nseg = (DIR ? W : H) >> 2;
    int ei, si;
    if (DIR == 0) { ei = blockIdx.x * 32 + (threadIdx.x & 31); si = blockIdx.y * 8 + (threadIdx.x >> 5); }   /* lanes across edges of one row band */
    else { si = blockIdx.x * 32 + (threadIdx.x & 31); ei = blockIdx.y * 8 + (threadIdx.x >> 5); }           /* lanes across columns (coalesced) */
    if (ei >= nedge || si >= nseg) return;
    const int e = (ei + 1) << 3, t = si << 2;
    const int xq = DIR ? t : e, yq = DIR ? e : t, xp = DIR ? t : e - 1, yp = DIR ? e - 1 : t;
    const ks_cell cp = cells[(yp >> 4) * pp.cw + (xp >> 4)];
    int bs;
    if (e & 8) {                /* inside a cell: only the boundaries between the four 8x8 intra CUs of a split cell (Bs 2, luma only) */
        if (cp.cu_log2 != 3) return;
        bs = 2;
    } else {
        const ks_cell cq = cells[(yq >> 4) * pp.cw + (xq >> 4)];
        if (!ks_is_tu_edge(cp, cq, xp, yp, xq, yq, e)) return;
        ks_cell_b pb, qb; pb.mvx1 = pb.mvy1 = qb.mvx1 = qb.mvy1 = 0; pb.dir = qb.dir = 1;
        if (cells_b) { pb = cells_b[(yp >> 4) * pp.cw + (xp >> 4)]; qb = cells_b[(yq >> 4) * pp.cw + (xq >> 4)]; }
        bs = ks_edge_bs(cp, cq, pb, qb);
        if (!bs) return;
    }
    const int beta = c_beta_table[ks_clip3(0, 51, pp.qp + (pp.beta_offset_div2 << 1))];
    const int tc = c_tc_table[ks_clip3(0, 53, pp.qp + 2 * (bs - 1) + (pp.tc_offset_div2 << 1))];
    uint8_t *Y = rec.p[0];
    int px[4][8];
    if (DIR == 0) {
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const uint32_t *r = reinterpret_cast<const uint32_t *>(Y + (size_t)(yq + l) * W + xq - 4);
            uint32_t a = r[0], b = r[1];
#pragma unroll
            for (int i = 0; i < 4; i++) { px[l][i] = (a >> (8 * i)) & 255; px[l][4 + i] = (b >> (8 * i)) & 255; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t a = *reinterpret_cast<const uint32_t *>(Y + (size_t)(yq - 4 + i) * W + xq);
#pragma unroll
            for (int l = 0; l < 4; l++) px[l][i] = (a >> (8 * l)) & 255;
        }
    }
    if (ks_deblock_luma_regs(px, beta, tc)) {
        if (DIR == 0) {
#pragma unroll
            for (int l = 0; l < 4; l++) {
                uint32_t *r = reinterpret_cast<uint32_t *>(Y + (size_t)(yq + l) * W + xq - 4);
                r[0] = (uint32_t)px[l][0] | ((uint32_t)px[l][1] << 8) | ((uint32_t)px[l][2] << 16) | ((uint32_t)px[l][3] << 24);
                r[1] = (uint32_t)px[l][4] | ((uint32_t)px[l][5] << 8) | ((uint32_t)px[l][6] << 16) | ((uint32_t)px[l][7] << 24);
            }
        } else {
#pragma unroll
            for (int i = 1; i < 7; i++)
                *reinterpret_cast<uint32_t *>(Y + (size_t)(yq - 4 + i) * W + xq) =
                    (uint32_t)px[0][i] | ((uint32_t)px[1][i] << 8) | ((uint32_t)px[2][i] << 16) | ((uint32_t)px[3][i] << 24);
        }
    }
    if (bs == 2 && !(e & 8)) {      /* chroma: only intra edges on the 8-sample chroma grid, 2 chroma lines per 4 luma lines (spec 8.7.2.5.5/.8) */
        const int tcc = c_tc_table[ks_clip3(0, 53, pp.qpc + 2 + (pp.tc_offset_div2 << 1))];
        const int CW = W >> 1, xs = DIR ? CW : 1, ys = DIR ? 1 : CW;
#pragma unroll
        for (int ci = 1; ci < 3; ci++) {
            uint8_t *q = rec.p[ci] + (size_t)(yq >> 1) * CW + (xq >> 1);
#pragma unroll
            for (int l = 0; l < 2; l++) {
                uint8_t *c = q + l * ys;
                int p0 = c[-xs], p1 = c[-2 * xs], q0 = c[0], q1 = c[xs];
                int delta = ks_clip3(-tcc, tcc, ((((q0 - p0) << 2) + p1 - q1 + 4) >> 3));
                c[-xs] = (uint8_t)ks_clip8(p0 + delta); c[0] = (uint8_t)ks_clip8(q0 - delta);
            }
        }
    }
}

void ks_launch_deblock(const KsPicParams &pp, KsPlanes rec, const ks_cell *cells, const ks_cell_b *cells_b, cudaStream_t st)
{
    int nev = (pp.W >> 3) - 1, nsv = pp.H >> 2, neh = (pp.H >> 3) - 1, nsh = pp.W >> 2;
    if (nev > 0) ks_deblock_kernel<0><<<dim3((nev + 31) / 32, (nsv + 7) / 8), 256, 0, st>>>(pp, rec, cells, cells_b);
    if (neh > 0) ks_deblock_kernel<1><<<dim3((nsh + 31) / 32, (neh + 7) / 8), 256, 0, st>>>(pp, rec, cells, cells_b);
}

/* ------------------------------------------------------------------ SAO -------------------------- */
#define KS_SAO_WARPS 8
#define KS_SAO_XOFF 16         /* tile row: [15] left halo, [16..] samples, then the right halo.  TMA needs the box's first byte on a 16-byte
                                  boundary of the plane (probed: x = -4 or 60 raise an illegal-instruction fault, x = -16 works), so the box
                                  starts 16 samples left of the CTU and its width is the row pitch */
#define KS_SAO_PITCH_Y 96
#define KS_SAO_PITCH_C 64
#define KS_SAO_TILE_Y (66 * KS_SAO_PITCH_Y)
#define KS_SAO_TILE_C (34 * KS_SAO_PITCH_C)
struct KsSaoSmem {
    /* deblocked samples incl. 1-sample halo, one TMA box per component (cp.async.bulk.tensor.2d, out-of-picture bytes arrive as
     * zeros and are masked by the passes below); 128-byte aligned destinations */
    __align__(128) uint8_t tileY[(KS_SAO_TILE_Y + 127) / 128 * 128];
    __align__(128) uint8_t tileC[2][(KS_SAO_TILE_C + 127) / 128 * 128];
    __align__(8) unsigned long long mbar;
    int hist[KS_SAO_WARPS][3][52];         /* packed (sum<<12)|count per warp: [20..51] BO band ([0..19] unused) */
    int ph[20][KS_SAO_WARPS * KS_WARP];    /* per-THREAD EO histograms (bin-major: conflict-free, no atomics), packed the same way */
    int sum[3][52], cnt[3][52];
    int cost[3][48], off[3][48];           /* [0..15] EO class*4+(cat-1), [16..47] BO band */
    ks_sao_param par[3];
    unsigned long long sse[3];
};
__device__ __forceinline__ int ks_sgn(int v) { return (v > 0) - (v < 0); }
__device__ __forceinline__ int ks_sao_offset_rd(int sum, int cnt, int signc, int lam, int is_bo, int *best_o)
{   /* == oracle sao_offset_rd */
    *best_o = 0;
    if (!cnt) return 0;
    int o = sum >= 0 ? (sum + cnt / 2) / cnt : -((-sum + cnt / 2) / cnt);
    o = ks_clip3(-7, 7, o);
    if ((signc > 0 && o < 0) || (signc < 0 && o > 0)) o = 0;
    int best = lam >> 4;
    int step = o > 0 ? -1 : 1;
    for (int v = o; v != 0; v += step) {
        int bits = abs(v) + 1 + (is_bo ? 1 : 0);
        int c = cnt * v * v - 2 * v * sum + ((lam * bits) >> 4);
        if (c < best || (c == best && abs(v) < abs(*best_o))) { best = c; *best_o = v; }
    }
    return best;
}
/* the 3x6 neighbourhood of a run of 4 samples starting at tile pointer t (row pitch `pitch`) */
struct KsSaoNb { uint32_t up, ce, dn; int ul, ur, cl, cr, dl, dr; };
__device__ __forceinline__ KsSaoNb ks_sao_load_nb(const uint8_t *t, int pitch)
{
    KsSaoNb n;
    const uint32_t *u = reinterpret_cast<const uint32_t *>(t - pitch), *c = reinterpret_cast<const uint32_t *>(t), *d = reinterpret_cast<const uint32_t *>(t + pitch);
    n.up = u[0]; n.ce = c[0]; n.dn = d[0];
    n.ul = u[-1] >> 24; n.cl = c[-1] >> 24; n.dl = d[-1] >> 24;
    n.ur = u[1] & 255; n.cr = c[1] & 255; n.dr = d[1] & 255;
    return n;
}
/* edge categories (0 none, 1..4) of sample j of the run for the four EO classes; out-of-picture neighbours -> 0 */
template <bool DIAG>
__device__ __forceinline__ void ks_sao_cats(const KsSaoNb &n, int j, bool lft, bool rgt, bool top, bool bot, int cat[4])
{
    const int lut = 0x43021;                       /* cat_of[0..4] = {1,2,0,3,4} as nibbles */
    int c = (n.ce >> (8 * j)) & 255;
    int l = j ? (n.ce >> (8 * j - 8)) & 255 : n.cl, r = j < 3 ? (n.ce >> (8 * j + 8)) & 255 : n.cr;
    int u = (n.up >> (8 * j)) & 255, d = (n.dn >> (8 * j)) & 255;
    int ul = j ? (n.up >> (8 * j - 8)) & 255 : n.ul, ur = j < 3 ? (n.up >> (8 * j + 8)) & 255 : n.ur;
    int dl = j ? (n.dn >> (8 * j - 8)) & 255 : n.dl, dr = j < 3 ? (n.dn >> (8 * j + 8)) & 255 : n.dr;
    bool h = !(lft || rgt), v = !(top || bot), hv = h && v;
    cat[0] = h ? (lut >> (4 * (2 + ks_sgn(c - l) + ks_sgn(c - r)))) & 15 : 0;
    cat[1] = v ? (lut >> (4 * (2 + ks_sgn(c - u) + ks_sgn(c - d)))) & 15 : 0;
    if (DIAG) {        /* the 135/45 degree classes are only gathered at -sao 4 (reference: SaoApplyOffsetEo2/3 only then) */
        cat[2] = hv ? (lut >> (4 * (2 + ks_sgn(c - ul) + ks_sgn(c - dr)))) & 15 : 0;
        cat[3] = hv ? (lut >> (4 * (2 + ks_sgn(c - ur) + ks_sgn(c - dl)))) & 15 : 0;
    } else cat[2] = cat[3] = 0;
}

/* edge category of sample j for EO class k only (the apply pass needs just the chosen class; k is uniform per component) */
__device__ __forceinline__ int ks_sao_cat_of(const KsSaoNb &n, int j, int k, bool lft, bool rgt, bool top, bool bot)
{
    const int lut = 0x43021;
    const int c = (n.ce >> (8 * j)) & 255;
    int a, b; bool ok;
    if (k == 0) { a = j ? (n.ce >> (8 * j - 8)) & 255 : n.cl; b = j < 3 ? (n.ce >> (8 * j + 8)) & 255 : n.cr; ok = !(lft || rgt); }
    else if (k == 1) { a = (n.up >> (8 * j)) & 255; b = (n.dn >> (8 * j)) & 255; ok = !(top || bot); }
    else if (k == 2) { a = j ? (n.up >> (8 * j - 8)) & 255 : n.ul; b = j < 3 ? (n.dn >> (8 * j + 8)) & 255 : n.dr; ok = !(lft || rgt || top || bot); }
    else { a = j < 3 ? (n.up >> (8 * j + 8)) & 255 : n.ur; b = j ? (n.dn >> (8 * j - 8)) & 255 : n.dl; ok = !(lft || rgt || top || bot); }
    return ok ? (lut >> (4 * (2 + ks_sgn(c - a) + ks_sgn(c - b)))) & 15 : 0;
}

/* ---- TMA helpers (sm_90+ PTX): one thread arms an mbarrier with the byte count and issues the tile copies; everyone polls ---- */
__device__ __forceinline__ uint32_t ks_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ks_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ks_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ks_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ks_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ks_tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(ks_smem_addr(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(ks_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void ks_mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok = 0;
    for (int spin = 0; !ok; spin++) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(ks_smem_addr(bar)), "r"(parity) : "memory");
        if (spin > (1 << 22)) __trap();       /* a copy that never lands must fail the launch, not hang the stream */
    }
}

__global__ void __launch_bounds__(KS_SAO_WARPS * KS_WARP)
ks_sao_kernel(KsPicParams pp, KsPlanes src, KsPlanes deb, KsPlanes out, ks_ctu_syn *__restrict__ ctus, uint32_t *__restrict__ sse_out,
              const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmCb, const __grid_constant__ CUtensorMap tmCr, int tma_mask)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    KsSaoSmem *sm = reinterpret_cast<KsSaoSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rx = blockIdx.x, ry = blockIdx.y;
    /* ---- stage the deblocked tile (+1 sample halo).  Components whose plane pitch is a multiple of 16 bytes come in as ONE TMA
     *      box each (96x66 / 64x34 bytes at (x0-16, y0-1)); the others (odd chroma pitch) with plain word loads.  Neighbours outside
     *      the picture are masked out in the statistics/apply passes, so the zero fill of out-of-bounds box parts is harmless. ---- */
    if (tid == 0) {
        ks_mbar_init(&sm->mbar, 1);
        unsigned bytes = 0;
        if (tma_mask & 1) bytes += KS_SAO_TILE_Y;
        if (tma_mask & 2) bytes += KS_SAO_TILE_C;
        if (tma_mask & 4) bytes += KS_SAO_TILE_C;
        if (bytes) {
            ks_mbar_expect_tx(&sm->mbar, bytes);
            if (tma_mask & 1) ks_tma_load_2d(sm->tileY, &tmY, (rx << 6) - KS_SAO_XOFF, (ry << 6) - 1, &sm->mbar);
            if (tma_mask & 2) ks_tma_load_2d(sm->tileC[0], &tmCb, (rx << 5) - KS_SAO_XOFF, (ry << 5) - 1, &sm->mbar);
            if (tma_mask & 4) ks_tma_load_2d(sm->tileC[1], &tmCr, (rx << 5) - KS_SAO_XOFF, (ry << 5) - 1, &sm->mbar);
        }
    }
    for (int i = tid; i < KS_SAO_WARPS * 3 * 52; i += blockDim.x) (&sm->hist[0][0][0])[i] = 0;
    if (tid < 3) sm->sse[tid] = 0;
    for (int ci = 0; ci < 3; ci++) {
        if ((tma_mask >> ci) & 1) continue;
        const int sh = ci ? 1 : 0, PW = pp.W >> sh, PH = pp.H >> sh, x0 = (rx << 6) >> sh, y0 = (ry << 6) >> sh;
        const int tw = 64 >> sh, pitch = ci ? KS_SAO_PITCH_C : KS_SAO_PITCH_Y, wpr = tw >> 2, bw = min(tw, PW - x0);
        const uint8_t *plane = deb.p[ci];
        uint8_t *tile = ci ? sm->tileC[ci - 1] : sm->tileY;
        for (int i = tid; i < (tw + 2) * wpr; i += blockDim.x) {
            int r = i / wpr, c = i - r * wpr;
            int gy = min(max(y0 - 1 + r, 0), PH - 1);
            uint32_t w = (4 * c < bw) ? __ldg(reinterpret_cast<const uint32_t *>(plane + (size_t)gy * PW + x0 + 4 * c)) : 0u;
            *reinterpret_cast<uint32_t *>(&tile[r * pitch + KS_SAO_XOFF + 4 * c]) = w;
        }
        for (int i = tid; i < (tw + 2) * 2; i += blockDim.x) {
            int r = i >> 1, side = i & 1;
            int gy = min(max(y0 - 1 + r, 0), PH - 1), gx = side ? min(x0 + bw, PW - 1) : max(x0 - 1, 0);
            tile[r * pitch + (side ? KS_SAO_XOFF + bw : KS_SAO_XOFF - 1)] = __ldg(plane + (size_t)gy * PW + gx);
        }
    }
    __syncthreads();                           /* also publishes the mbarrier initialisation to the pollers */
    if (tma_mask) ks_mbar_wait(&sm->mbar, 0);
    /* ---- statistics: runs of 4 samples per thread.  EO: every thread owns a private column of a bin-major shared
     *      histogram (plain read-modify-write, conflict-free, no atomics; the reference's packed (d<<12)|1 accumulator);
     *      BO: per-warp shared histograms with run aggregation. ---- */
    if (pp.sao) {
        for (int ci = 0; ci < 3; ci++) {
            const int sh = ci ? 1 : 0, PW = pp.W >> sh, PH = pp.H >> sh, x0 = (rx << 6) >> sh, y0 = (ry << 6) >> sh;
            const int bw = min(64 >> sh, PW - x0), bh = min(64 >> sh, PH - y0), pitch = ci ? KS_SAO_PITCH_C : KS_SAO_PITCH_Y, rl = 4 - sh;
            int *h = sm->hist[warp][ci];
#pragma unroll
            for (int b = 0; b < 20; b++) sm->ph[b][tid] = 0;
            const int step = pp.sao >= 4 ? 1 : 2, ncls = pp.sao >= 4 ? 4 : 2;     /* `_fast` statistics: every 2nd row, EO classes 0,1 */
            for (int i = tid; i < (((bh + step - 1) / step) << rl); i += blockDim.x) {
                const int y = (i >> rl) * step, x = (i & ((1 << rl) - 1)) << 2;
                if (x >= bw) continue;
                const KsSaoNb n = ks_sao_load_nb(&(ci ? sm->tileC[ci - 1] : sm->tileY)[(y + 1) * pitch + KS_SAO_XOFF + x], pitch);
                const uint32_t s4 = __ldg(reinterpret_cast<const uint32_t *>(src.p[ci] + (size_t)(y0 + y) * PW + x0 + x));
                const bool top = y0 + y == 0, bot = y0 + y == PH - 1;
                int cur_band = -1, band_acc = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = (n.ce >> (8 * j)) & 255, d = (int)((s4 >> (8 * j)) & 255) - c, v = d * 4096 + 1;
                    const int band = c >> 3;
                    if (band != cur_band) { if (cur_band >= 0) atomicAdd(&h[20 + cur_band], band_acc); cur_band = band; band_acc = 0; }
                    band_acc += v;
                    int cat[4];
                    if (ncls == 4) {
                        ks_sao_cats<true>(n, j, x0 + x + j == 0, x0 + x + j == PW - 1, top, bot, cat);
#pragma unroll
                        for (int k = 0; k < 4; k++) sm->ph[k * 5 + cat[k]][tid] += v;      /* bin cat 0 collects the rest and is ignored */
                    } else {
                        ks_sao_cats<false>(n, j, x0 + x + j == 0, x0 + x + j == PW - 1, top, bot, cat);
                        sm->ph[cat[0]][tid] += v; sm->ph[5 + cat[1]][tid] += v;
                    }
                }
                atomicAdd(&h[20 + cur_band], band_acc);
            }
            __syncthreads();
            /* reduce the 256 private columns: warp w owns bins w, w+8, w+16; decode before the cross-lane sum (count <= 4096) */
            for (int b = warp; b < 20; b += KS_SAO_WARPS) {
                int s_ = 0, n_ = 0;
#pragma unroll
                for (int i = 0; i < KS_SAO_WARPS; i++) { int v = sm->ph[b][lane + 32 * i]; int c = v & 4095; n_ += c; s_ += (v - c) >> 12; }
                s_ = __reduce_add_sync(0xffffffffu, s_); n_ = __reduce_add_sync(0xffffffffu, n_);
                if (lane == 0) { sm->sum[ci][b] = s_; sm->cnt[ci][b] = n_; }
            }
            __syncthreads();
        }
        if (tid < 3 * 32) {
            int ci = tid >> 5, b = 20 + (tid & 31), s_ = 0, n_ = 0;
            for (int w = 0; w < KS_SAO_WARPS; w++) { int v = sm->hist[w][ci][b]; int c = v & 4095; n_ += c; s_ += (v - c) >> 12; }
            sm->sum[ci][b] = s_; sm->cnt[ci][b] = n_;
        }
        __syncthreads();
        /* ---- per-bin offset RD (144 independent little problems) ---- */
        if (tid < 3 * 48) {
            int ci = tid / 48, j = tid - ci * 48, o, c;
            if (j < 16) { int k = j >> 2, cat = (j & 3) + 1; c = ks_sao_offset_rd(sm->sum[ci][k * 5 + cat], sm->cnt[ci][k * 5 + cat], cat <= 2 ? 1 : -1, pp.lambda_sse_q4, 0, &o); }
            else c = ks_sao_offset_rd(sm->sum[ci][20 + j - 16], sm->cnt[ci][20 + j - 16], 0, pp.lambda_sse_q4, 1, &o);
            sm->cost[ci][j] = c; sm->off[ci][j] = o;
        }
    }
    __syncthreads();
    /* ---- type decision: thread 0 luma, thread 32 chroma (Cb and Cr share type / class) ---- */
    if (tid == 0 || tid == 32) {
        const int c0 = tid ? 1 : 0, c1 = tid ? 2 : 0, lam = pp.lambda_sse_q4;
        int best_cost = 0, best_type = 0, best_class = 0, best_band[3] = {0, 0, 0};
        if (pp.sao) {
            for (int k = 0; k < (pp.sao >= 4 ? 4 : 2); k++) {
                int total = (lam * 4) >> 4;
                for (int ci = c0; ci <= c1; ci++) for (int j = 0; j < 4; j++) total += sm->cost[ci][k * 4 + j];
                if (total < best_cost) { best_cost = total; best_type = 2; best_class = k; }
            }
            int total = (lam * 7) >> 4, band[3] = {0, 0, 0};
            for (int ci = c0; ci <= c1; ci++) {
                int bestc = 0x7fffffff, bs = 0;
                for (int s = 0; s <= 28; s++) { int c = sm->cost[ci][16 + s] + sm->cost[ci][17 + s] + sm->cost[ci][18 + s] + sm->cost[ci][19 + s]; if (c < bestc) { bestc = c; bs = s; } }
                total += bestc; band[ci] = bs;
            }
            if (total < best_cost) { best_cost = total; best_type = 1; for (int ci = c0; ci <= c1; ci++) best_band[ci] = band[ci]; }
            int nz = 0;
            for (int ci = c0; ci <= c1; ci++) for (int j = 0; j < 4; j++)
                nz |= best_type == 2 ? sm->off[ci][best_class * 4 + j] : (best_type == 1 ? sm->off[ci][16 + best_band[ci] + j] : 0);
            if (!nz) best_type = 0;
        }
        for (int ci = c0; ci <= c1; ci++) {
            ks_sao_param p; p.type = (uint8_t)best_type; p.band_or_class = 0; p.off[0] = p.off[1] = p.off[2] = p.off[3] = 0;
            if (best_type) {
                p.band_or_class = (uint8_t)(best_type == 2 ? best_class : best_band[ci]);
                for (int j = 0; j < 4; j++) p.off[j] = (int8_t)(best_type == 2 ? sm->off[ci][best_class * 4 + j] : sm->off[ci][16 + best_band[ci] + j]);
            }
            sm->par[ci] = p;
            ctus[ry * pp.ctw + rx].sao[ci] = p;
        }
        if (tid == 0) { ctus[ry * pp.ctw + rx].rsv[0] = 0; ctus[ry * pp.ctw + rx].rsv[1] = 0; }
    }
    __syncthreads();
    /* ---- apply + write the final picture + SSE against the source ---- */
    for (int ci = 0; ci < 3; ci++) {
        const int sh = ci ? 1 : 0, PW = pp.W >> sh, PH = pp.H >> sh, x0 = (rx << 6) >> sh, y0 = (ry << 6) >> sh;
        const int bw = min(64 >> sh, PW - x0), bh = min(64 >> sh, PH - y0), pitch = ci ? KS_SAO_PITCH_C : KS_SAO_PITCH_Y, rl = 4 - sh;
        const ks_sao_param p = sm->par[ci];
        const int k = p.band_or_class;
        const int o0 = p.off[0], o1 = p.off[1], o2 = p.off[2], o3 = p.off[3];
        unsigned sse = 0;
        for (int i = tid; i < (bh << rl); i += blockDim.x) {
            const int y = i >> rl, x = (i & ((1 << rl) - 1)) << 2;
            if (x >= bw) continue;
            const KsSaoNb n = ks_sao_load_nb(&(ci ? sm->tileC[ci - 1] : sm->tileY)[(y + 1) * pitch + KS_SAO_XOFF + x], pitch);
            const uint32_t s4 = __ldg(reinterpret_cast<const uint32_t *>(src.p[ci] + (size_t)(y0 + y) * PW + x0 + x));
            const bool top = y0 + y == 0, bot = y0 + y == PH - 1;
            const bool in_disp_y = y0 + y < (pp.dH >> sh);
            const int dispw = pp.dW >> sh;
            uint32_t o4 = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int c = (n.ce >> (8 * j)) & 255, v = c;
                if (p.type == 1) { int b = ((c >> 3) - k) & 31; if (b < 4) v = c + (b == 0 ? o0 : b == 1 ? o1 : b == 2 ? o2 : o3); }
                else if (p.type == 2) {
                    const int ct = ks_sao_cat_of(n, j, k, x0 + x + j == 0, x0 + x + j == PW - 1, top, bot);
                    if (ct) v = c + (ct == 1 ? o0 : ct == 2 ? o1 : ct == 3 ? o2 : o3);
                }
                v = ks_clip8(v);
                o4 |= (uint32_t)v << (8 * j);
                int e = (int)((s4 >> (8 * j)) & 255) - v;
                if (in_disp_y && x0 + x + j < dispw) sse += (unsigned)(e * e);
            }
            *reinterpret_cast<uint32_t *>(out.p[ci] + (size_t)(y0 + y) * PW + x0 + x) = o4;
        }
        if (sse_out) {
            unsigned lo = ks_warp_sum(sse);                  /* per-thread SSE <= 16 samples * 65025 */
            if (lane == 0) atomicAdd(&sm->sse[ci], (unsigned long long)lo);
        }
    }
    if (sse_out) {       /* per-CTU partial sums (<= 64*64*255^2 fits 32 bits); ks_pack_scan_kernel adds them up: no contended global atomics */
        __syncthreads();
        if (tid < 3) sse_out[(ry * pp.ctw + rx) * 3 + tid] = (uint32_t)sm->sse[tid];
    }
}

void ks_launch_sao(const KsPicParams &pp, KsPlanes src, KsPlanes deb, KsPlanes out, ks_ctu_syn *ctus, uint32_t *sse_out,
                   const CUtensorMap *tm, int tma_mask, cudaStream_t st)
{
    alignas(64) CUtensorMap m[3];                   /* the caller's copy may sit at any alignment */
    memcpy(m, tm, sizeof(m));
    ks_sao_kernel<<<dim3(pp.ctw, pp.cth), KS_SAO_WARPS * KS_WARP, sizeof(KsSaoSmem) + 128, st>>>(pp, src, deb, out, ctus, sse_out, m[0], m[1], m[2], tma_mask);
}
