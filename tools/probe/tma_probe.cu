// tma_probe.cu -- development probe: does a 2-D u8 TMA box load with the given geometry work on this GPU?
// usage: tma_probe W H BW BH X Y [dyn]     (prints OK / MISMATCH / CUDA error; one configuration per process)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int bytes, uint8_t *out, int use_dyn)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(128) unsigned char stat[16384];
    __shared__ __align__(8) unsigned long long bar;
    unsigned char *buf = use_dyn ? dyn : stat;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(saddr(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(saddr(buf)), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(x), "r"(y), "r"(saddr(&bar)) : "memory");
    }
    __syncthreads();
    unsigned ok = 0;
    for (int spin = 0; !ok && spin < (1 << 20); spin++)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(saddr(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = ok ? buf[i] : 0xEE;
}
int main(int argc, char **argv)
{
    int W = atoi(argv[1]), H = atoi(argv[2]), BW = atoi(argv[3]), BH = atoi(argv[4]), X = atoi(argv[5]), Y = atoi(argv[6]), dyn = argc > 7 ? atoi(argv[7]) : 0;
    uint8_t *h = (uint8_t *)malloc((size_t)W * H), *d, *o, *ho = (uint8_t *)malloc((size_t)BW * BH);
    for (int i = 0; i < W * H; i++) h[i] = (uint8_t)(1 + (i * 7 + i / W * 13) % 251);
    cudaMalloc(&d, (size_t)W * H); cudaMalloc(&o, (size_t)BW * BH); cudaMemcpy(d, h, (size_t)W * H, cudaMemcpyHostToDevice);
    void *p = NULL; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 2; }
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
    CUresult r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
    if (dyn) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    k<<<1, 256, dyn ? 32768 : 0>>>(tm, X, Y, BW * BH, o, dyn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 4; }
    cudaMemcpy(ho, o, (size_t)BW * BH, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < BH; yy++) for (int xx = 0; xx < BW; xx++) {
        int gx = X + xx, gy = Y + yy; uint8_t want = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0 : h[(size_t)gy * W + gx];
        if (ho[yy * BW + xx] != want) bad++;
    }
    printf(bad ? "MISMATCH %d (first byte %02x)\n" : "OK\n", bad, ho[0]);
    return bad ? 1 : 0;
}
