// experiment: which small TMA boxes are legal on sm_100a?  usage: tma_box <boxw> <boxh> <rank> <x> <y>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int rank, uint32_t bytes, uint8_t *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        uint32_t d = (uint32_t)__cvta_generic_to_shared(sm);
        if (rank == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(d), "l"(&tm), "r"(x), "r"(y), "r"(0), "r"(0), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(d), "l"(&tm), "r"(x), "r"(y), "r"(b) : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0u) : "memory");
    for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char **argv)
{
    int bw = atoi(argv[1]), bh = atoi(argv[2]), rank = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]);
    const int W = 256, H = 64;
    uint8_t *d, *o; cudaMalloc(&d, W * H * 2); cudaMalloc(&o, 65536);
    uint8_t h[W * H]; for (int i = 0; i < W * H; i++) h[i] = (uint8_t)(i * 7 + i / W);
    cudaMemcpy(d, h, W * H, cudaMemcpyHostToDevice);
    void *fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*F)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap tm;
    cuuint64_t dims[4] = {W, H, 1, 2}, strides[3] = {W, (cuuint64_t)W * H, (cuuint64_t)W * H};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = ((F)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d rank %d at (%d,%d): encode=%d ", bw, bh, rank, x, y, (int)r);
    k<<<1, 128, bw * bh + 128>>>(tm, x, y, rank, bw * bh, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        uint8_t *g = (uint8_t *)malloc(bw * bh); cudaMemcpy(g, o, bw * bh, cudaMemcpyDeviceToHost);
        int bad = 0; for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++) if (g[j * bw + i] != h[(y + j) * W + x + i]) bad++;
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
