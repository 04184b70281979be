// dev probe: which TMA tensor-map / coordinate combinations are legal on this part
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int rank, int x, int y, int z, int bytes, float* out, int n) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(sm)), "l"(&tmap), "r"(s32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(sm)), "l"(&tmap), "r"(s32(&bar)), "r"(x), "r"(y) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s32(&bar)), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int rank = atoi(argv[1]), W = atoi(argv[2]), H = atoi(argv[3]), B = atoi(argv[4]);
    int bw = atoi(argv[5]), bh = atoi(argv[6]), x = atoi(argv[7]), y = atoi(argv[8]), z = atoi(argv[9]);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Enc enc = (Enc)p;
    size_t n = (size_t)W * H * B;
    float* h = (float*)malloc(n * 4);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    int nb = bw * bh;
    cudaMalloc(&o, nb * 4);
    CUtensorMap m; memset(&m, 0, sizeof(m));
    cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gs[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t bx[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d  ", (int)r);
    if (r) return 1;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, nb * 4 + 128);
    k<<<1, 128, nb * 4, 0>>>(m, rank, x, y, z, nb * 4, o, nb);
    cudaError_t e = cudaDeviceSynchronize();
    float* ho = (float*)malloc(nb * 4);
    cudaMemcpy(ho, o, nb * 4, cudaMemcpyDeviceToHost);
    printf("run: %s  first=%g mid=%g last=%g\n", cudaGetErrorString(e), ho[0], ho[nb / 2], ho[nb - 1]);
    return 0;
}
