// Probe: DRAM read bandwidth of warp-private TMA tile loads as a function of the box shape, the number of boxes in
// flight per warp (DEPTH) and the resident warps per SM (limited through dynamic shared memory padding).
// Each warp walks a 64-column strip of a [B, H, W] fp32 array like the splat backward does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_tile_bw tma_tile_bw.cu -lcuda && ./tma_tile_bw
#include <cstdio>
#include <cstdlib>
#include "../../fireflies_b200/csrc/ffb_tma.cuh"
using namespace ffb;

template <int BX, int BY, int DEPTH>
__global__ void __launch_bounds__(32) walk(const __grid_constant__ CUtensorMap tm, int strips_x, int tiles_per_warp, float* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    constexpr int BOX = BX * BY * 4;
    constexpr int PER_ROW = 64 / BX > 0 ? 64 / BX : 1;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + DEPTH * BOX);
    const int lane = threadIdx.x;
    if (lane == 0) { for (int d = 0; d < DEPTH; ++d) tma::mbar_init(bar + d, 1); tma::fence_mbar_init(); }
    __syncwarp();
    const int first = blockIdx.x * tiles_per_warp, b = blockIdx.y;
    auto issue = [&](int t, int d) {
        const int strip = t / PER_ROW, j = t % PER_ROW;
        const int sx = strip % strips_x, sy = strip / strips_x;
        tma::mbar_expect_tx(bar + d, BOX);
        tma::load_3d(sm + d * BOX, &tm, bar + d, (sx * PER_ROW + j) * BX, sy * BY, b);
    };
    float acc = 0.f;
    unsigned phases = 0;
    if (lane == 0)
        for (int d = 0; d < DEPTH && d < tiles_per_warp; ++d) issue(first + d, d);
    for (int i0 = 0; i0 < tiles_per_warp; i0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const int i = i0 + d;
            if (i >= tiles_per_warp) break;
            tma::mbar_wait(bar + d, (phases >> d) & 1u); phases ^= 1u << d;
            acc += reinterpret_cast<const float*>(sm + d * BOX)[lane];
            __syncwarp();
            if (lane == 0 && i + DEPTH < tiles_per_warp) issue(first + i + DEPTH, d);
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

template <int BX, int BY, int DEPTH>
void run(const float* a, int B, int H, int W, float* sink, int warps_per_sm) {
    CUtensorMap tm;
    if (!tma::encode_f32_3d(&tm, a, W, H, B, BX, BY, CU_TENSOR_MAP_SWIZZLE_NONE)) { printf("encode failed\n"); return; }
    constexpr int PER_ROW = 64 / BX > 0 ? 64 / BX : 1;
    const int tiles = (W / BX) * (H / BY);
    const int tpw = 16 * 256 / (BX * BY) > 0 ? 16 * 256 / (BX * BY) : 1;     // same bytes per warp as 16 tiles of 16x16
    int smem = DEPTH * BX * BY * 4 + 64;
    const int want = 227 * 1024 / warps_per_sm - 1024;                       // pad shared memory to cap the resident warps
    if (want > smem) smem = want;
    cudaFuncSetAttribute(walk<BX, BY, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dim3 grid((unsigned)(tiles / tpw), B);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    walk<BX, BY, DEPTH><<<grid, 32, smem>>>(tm, (W / BX) / PER_ROW, tpw, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) walk<BX, BY, DEPTH><<<grid, 32, smem>>>(tm, (W / BX) / PER_ROW, tpw, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    printf("box %3d x %2d (row %4d B, %4d B) depth %d warps/SM %2d: %.3f ms  %.0f GB/s  [%s]\n", BX, BY, BX * 4, BX * BY * 4, DEPTH, warps_per_sm, ms,
           (double)B * H * W * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int B = 64, H = 2048, W = 2048;
    float *a, *sink;
    cudaMalloc(&a, (size_t)B * H * W * 4); cudaMalloc(&sink, 4);
    cudaMemset(a, 0, (size_t)B * H * W * 4);
    for (int w : {8, 16, 24, 32}) run<16, 16, 1>(a, B, H, W, sink, w);
    for (int w : {8, 16, 24, 32}) run<16, 16, 2>(a, B, H, W, sink, w);
    for (int w : {8, 16, 32}) run<16, 16, 4>(a, B, H, W, sink, w);
    for (int w : {8, 16, 24, 32}) run<32, 16, 1>(a, B, H, W, sink, w);
    for (int w : {8, 16, 24}) run<32, 16, 2>(a, B, H, W, sink, w);
    for (int w : {8, 16, 24}) run<64, 16, 1>(a, B, H, W, sink, w);
    for (int w : {8, 16, 24}) run<64, 16, 2>(a, B, H, W, sink, w);
    return 0;
}
