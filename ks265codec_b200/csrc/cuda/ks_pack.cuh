/*
 * ks_pack.cuh -- compaction of the dense int16 level planes into the boundary format of include/ks265_syntax.h:
 * per-CTU bitmaps of non-zero 4x4 coefficient groups + a pool holding only those groups, in a canonical order
 * (CTU raster; inside a CTU: Y rows, Cb rows, Cr rows, left to right) so the host CABAC stage can index it and
 * two runs produce byte-identical buffers (count -> exclusive scan -> write; no atomics).
 * This is what crosses PCIe instead of 2 bytes per sample.  Mirror of oracle ora_pack_levels.
 */
#pragma once
#include "ks_common.cuh"

struct KsPackSmem { uint32_t m[12]; uint32_t pre[13]; };

__device__ __forceinline__ bool ks_cg_nonzero(const int16_t *__restrict__ plane, int PW, int PH, int x, int y)
{
    if (x >= PW || y >= PH) return false;
    const int16_t *s = plane + (size_t)y * PW + x;
    unsigned long long a = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) a |= *reinterpret_cast<const unsigned long long *>(s + (size_t)j * PW);
    return a != 0;
}
/* 384 CG slots per CTU: 0..255 luma (16x16), 256..319 Cb (8x8), 320..383 Cr; 12 warps of 32 slots */
__device__ __forceinline__ void ks_cg_slot(int slot, int rx, int ry, int &ci, int &x, int &y)
{
    if (slot < 256) { ci = 0; x = (rx << 6) + ((slot & 15) << 2); y = (ry << 6) + ((slot >> 4) << 2); }
    else { int s = slot - 256; ci = 1 + (s >> 6); s &= 63; x = (rx << 5) + ((s & 7) << 2); y = (ry << 5) + ((s >> 3) << 2); }
}

__global__ void __launch_bounds__(384)
ks_pack_count_kernel(KsPicParams pp, KsLevels lv, ks_ctu_syn *__restrict__ ctus, uint32_t *__restrict__ counts)
{
    __shared__ uint32_t m[12];
    const int rx = blockIdx.x, ry = blockIdx.y, slot = threadIdx.x;
    int ci, x, y; ks_cg_slot(slot, rx, ry, ci, x, y);
    bool nz = ks_cg_nonzero(lv.p[ci], pp.W >> (ci ? 1 : 0), pp.H >> (ci ? 1 : 0), x, y);
    unsigned b = __ballot_sync(0xffffffffu, nz);
    if ((slot & 31) == 0) m[slot >> 5] = b;
    __syncthreads();
    if (slot == 0) {
        ks_ctu_syn *ct = &ctus[ry * pp.ctw + rx];
        uint32_t n = 0;
        for (int w = 0; w < 8; w++) { ct->cg_y[2 * w] = (uint16_t)(m[w] & 0xffffu); ct->cg_y[2 * w + 1] = (uint16_t)(m[w] >> 16); n += __popc(m[w]); }
        for (int w = 0; w < 2; w++) for (int k = 0; k < 4; k++) { ct->cg_cb[4 * w + k] = (uint8_t)(m[8 + w] >> (8 * k)); ct->cg_cr[4 * w + k] = (uint8_t)(m[10 + w] >> (8 * k)); }
        n += __popc(m[8]) + __popc(m[9]) + __popc(m[10]) + __popc(m[11]);
        counts[ry * pp.ctw + rx] = n;
    }
}

/* exclusive scan of the per-CTU counts (<= 8160 CTUs at 8K): one block */
__global__ void __launch_bounds__(1024)
ks_pack_scan_kernel(int nctu, const uint32_t *__restrict__ counts, ks_ctu_syn *__restrict__ ctus, uint32_t *__restrict__ n_cg,
                    const uint32_t *__restrict__ sse_ctu, unsigned long long *__restrict__ sse_out)
{
    __shared__ uint32_t part[1024];
    if (sse_ctu) {       /* picture SSE = sum of the SAO kernel's per-CTU partial sums (three planes) */
        __shared__ unsigned long long acc[3];
        if (threadIdx.x < 3) acc[threadIdx.x] = 0;
        __syncthreads();
        unsigned long long a0 = 0, a1 = 0, a2 = 0;
        for (int i = threadIdx.x; i < nctu; i += 1024) { a0 += sse_ctu[3 * i]; a1 += sse_ctu[3 * i + 1]; a2 += sse_ctu[3 * i + 2]; }
        for (int o = 16; o; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&acc[0], a0); atomicAdd(&acc[1], a1); atomicAdd(&acc[2], a2); }
        __syncthreads();
        if (threadIdx.x < 3) sse_out[threadIdx.x] = acc[threadIdx.x];
    }
    const int tid = threadIdx.x, per = (nctu + 1023) / 1024, b = tid * per, e = min(b + per, nctu);
    uint32_t s = 0;
    for (int i = b; i < e; i++) s += counts[i];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { uint32_t v = tid >= o ? part[tid - o] : 0; __syncthreads(); part[tid] += v; __syncthreads(); }
    uint32_t base = part[tid] - s;
    for (int i = b; i < e; i++) { ctus[i].cg_base = base; base += counts[i]; }
    if (tid == 1023) *n_cg = part[1023];
}

__global__ void __launch_bounds__(384)
ks_pack_write_kernel(KsPicParams pp, KsLevels lv, const ks_ctu_syn *__restrict__ ctus, int16_t *__restrict__ pool)
{
    __shared__ uint32_t m[12], pre[12];
    const int rx = blockIdx.x, ry = blockIdx.y, slot = threadIdx.x;
    const ks_ctu_syn *ct = &ctus[ry * pp.ctw + rx];
    if (slot < 8) m[slot] = (uint32_t)ct->cg_y[2 * slot] | ((uint32_t)ct->cg_y[2 * slot + 1] << 16);
    else if (slot < 10) { int w = slot - 8; m[slot] = (uint32_t)ct->cg_cb[4 * w] | ((uint32_t)ct->cg_cb[4 * w + 1] << 8) | ((uint32_t)ct->cg_cb[4 * w + 2] << 16) | ((uint32_t)ct->cg_cb[4 * w + 3] << 24); }
    else if (slot < 12) { int w = slot - 10; m[slot] = (uint32_t)ct->cg_cr[4 * w] | ((uint32_t)ct->cg_cr[4 * w + 1] << 8) | ((uint32_t)ct->cg_cr[4 * w + 2] << 16) | ((uint32_t)ct->cg_cr[4 * w + 3] << 24); }
    __syncthreads();
    if (slot == 0) { uint32_t a = 0; for (int w = 0; w < 12; w++) { pre[w] = a; a += __popc(m[w]); } }
    __syncthreads();
    const uint32_t bits = m[slot >> 5];
    if (!((bits >> (slot & 31)) & 1)) return;
    const uint32_t idx = ct->cg_base + pre[slot >> 5] + __popc(bits & ((1u << (slot & 31)) - 1u));
    int ci, x, y; ks_cg_slot(slot, rx, ry, ci, x, y);
    const int PW = pp.W >> (ci ? 1 : 0);
    const int16_t *s = lv.p[ci] + (size_t)y * PW + x;
    unsigned long long *d = reinterpret_cast<unsigned long long *>(pool + (size_t)idx * 16);
#pragma unroll
    for (int j = 0; j < 4; j++) d[j] = *reinterpret_cast<const unsigned long long *>(s + (size_t)j * PW);
}

void ks_launch_pack(const KsPicParams &pp, KsLevels lv, ks_ctu_syn *ctus, int16_t *pool, uint32_t *n_cg, uint32_t *scan_ws, const uint32_t *sse_ctu, unsigned long long *sse_out, cudaStream_t st)
{
    dim3 grid(pp.ctw, pp.cth);
    ks_pack_count_kernel<<<grid, 384, 0, st>>>(pp, lv, ctus, scan_ws);
    ks_pack_scan_kernel<<<1, 1024, 0, st>>>(pp.ctw * pp.cth, scan_ws, ctus, n_cg, sse_ctu, sse_out);
    ks_pack_write_kernel<<<grid, 384, 0, st>>>(pp, lv, ctus, pool);
}
