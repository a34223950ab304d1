/*
 * ks_kernels.cu -- the single CUDA translation unit of libks265gpu.so (sm_100a): all hot-path kernels
 * (ks_me.cuh, ks_recon.cuh, ks_loopfilter.cuh, ks_pack.cuh) + the small known-answer kernels behind ks_gpu_kat_*.
 */
#include "ks_common.cuh"
#include "ks_me.cuh"
#include "ks_decide.cuh"
#include "ks_recon.cuh"
#include "ks_loopfilter.cuh"
#include "ks_pack.cuh"
#include "ks_kat.h"

#include <cstdio>
#include <mutex>
int ks_init_device(int device)
{
    static std::mutex mu;
    static bool done[64];
    std::lock_guard<std::mutex> lock(mu);
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return -1;
    if (device >= 64) return -1;
    if (done[device]) return 0;
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    ks_upload_tables_impl();
    bool ok = true;
    ok = ok && cudaFuncSetAttribute(ks_recon_inter_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KsReconSmem)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(ks_recon_inter_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KsReconSmem)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(ks_decide_cand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KsCandSmem)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(ks_sao_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KsSaoSmem) + 128) == cudaSuccess;
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    if (!ok) { fprintf(stderr, "ks265gpu: device %d initialisation failed: %s\n", device, cudaGetErrorString(cudaGetLastError())); return -1; }
    done[device] = true;
    return 0;
}

/* ------------------------------------------------------------------ KAT kernels ------------------ */
__global__ void ks_kat_sad16_kernel(const uint8_t *a, const uint8_t *b16, uint32_t *out)
{
    __shared__ __align__(16) KsWarpScratch sc;
    const int lane = threadIdx.x;
    const uint2 s = *reinterpret_cast<const uint2 *>(a + (lane >> 1) * 16 + 8 * (lane & 1));
    int wx0, wy0;
    ks_center_window(0, 0, 0, 0, wx0, wy0);
    ks_load_window(sc.win, b16, 16, 16, wx0, wy0, lane);
    unsigned v = ks_warp_sum(ks_sad_partial(sc.win, -wx0, -wy0, lane, s.x, s.y));
    if (lane == 0) *out = v;
}
__global__ void ks_kat_satd16_kernel(const uint8_t *a, const uint8_t *b16, uint32_t *out)
{
    const int lane = threadIdx.x;
    const uint2 s = *reinterpret_cast<const uint2 *>(a + (lane >> 1) * 16 + 8 * (lane & 1));
    const uint2 r = *reinterpret_cast<const uint2 *>(b16 + (lane >> 1) * 16 + 8 * (lane & 1));
    unsigned v = ks_satd16(r.x, r.y, s.x, s.y, lane);
    if (lane == 0) *out = v;
}
__global__ void ks_kat_interp_kernel(const uint8_t *plane, int w, int h, int x, int y, int mvx, int mvy, uint8_t *dst)
{
    __shared__ __align__(16) KsWarpScratch sc;
    const int lane = threadIdx.x;
    int wx0, wy0;
    ks_center_window(x, y, mvx >> 2, mvy >> 2, wx0, wy0);
    ks_load_window(sc.win, plane, w, h, wx0, wy0, lane);
    uint32_t o0, o1;
    ks_interp16(&sc, x + (mvx >> 2) - wx0, y + (mvy >> 2) - wy0, mvx & 3, mvy & 3, lane, o0, o1);
    *reinterpret_cast<uint2 *>(dst + (lane >> 1) * 16 + 8 * (lane & 1)) = make_uint2(o0, o1);
}
template <int N>
__global__ void ks_kat_tb_kernel(const uint8_t *src, const uint8_t *pred, int qp, int intra_slice, int sign_hiding,
                                 int16_t *levels, uint8_t *recon, int *cbf)
{
    __shared__ __align__(16) KsTbScratch ts;
    __shared__ __align__(16) uint8_t sp[32 * 32];
    __shared__ uint16_t scan[1024];
    __shared__ __align__(16) int t0[256];
    const int lane = threadIdx.x;
    ks_load_t0(t0, lane, 32);
    for (int i = lane; i < N * N; i += 32) { sp[i] = pred[i]; scan[i] = c_scan_tb[KsLog2<N>::v - 2][i]; }
    __syncwarp();
    const int g = lane / N, r = lane % N;
    bool c = ks_tb_code<N>(&ts, scan, t0, g == 0, src + r * N, sp + r * N, recon + r * N, levels + r * N, qp, intra_slice, sign_hiding, lane);
    if (lane == 0) *cbf = c;
}

int ks_kat_sad16_dev(const uint8_t *a, const uint8_t *b16, uint32_t *out) { ks_kat_sad16_kernel<<<1, 32>>>(a, b16, out); return cudaGetLastError() == cudaSuccess ? 0 : -1; }
int ks_kat_satd16_dev(const uint8_t *a, const uint8_t *b16, uint32_t *out) { ks_kat_satd16_kernel<<<1, 32>>>(a, b16, out); return cudaGetLastError() == cudaSuccess ? 0 : -1; }
int ks_kat_interp_dev(const uint8_t *plane, int w, int h, int x, int y, int mvx, int mvy, uint8_t *dst)
{ ks_kat_interp_kernel<<<1, 32>>>(plane, w, h, x, y, mvx, mvy, dst); return cudaGetLastError() == cudaSuccess ? 0 : -1; }
int ks_kat_tb_dev(int log2n, const uint8_t *src, const uint8_t *pred, int qp, int intra_slice, int sign_hiding, int16_t *levels, uint8_t *recon, int *cbf)
{
    if (log2n == 2) ks_kat_tb_kernel<4><<<1, 32>>>(src, pred, qp, intra_slice, sign_hiding, levels, recon, cbf);
    else if (log2n == 3) ks_kat_tb_kernel<8><<<1, 32>>>(src, pred, qp, intra_slice, sign_hiding, levels, recon, cbf);
    else if (log2n == 4) ks_kat_tb_kernel<16><<<1, 32>>>(src, pred, qp, intra_slice, sign_hiding, levels, recon, cbf);
    else if (log2n == 5) ks_kat_tb_kernel<32><<<1, 32>>>(src, pred, qp, intra_slice, sign_hiding, levels, recon, cbf);
    else return -1;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
