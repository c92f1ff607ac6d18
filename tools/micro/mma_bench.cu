// Micro-benchmark: cycles per tcgen05.mma for the shapes the attention kernel issues (one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I image_restoration_and_enhancement_b200/csrc tools/micro/mma_bench.cu -o gpurun_out/mma_bench
#include <cstdio>
#include "common.cuh"
using namespace rg;

// mode: 0 = SS (A smem K-major, B smem K-major), 1 = SS with B MN-major, 2 = TS (A tmem, B MN-major), 3 = TS B K-major
template <int N, int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_smem;
    if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, (MODE == 1 || MODE == 2) ? 1 : 0);
        const uint32_t sa = smem_u32(smem), sb = sa + 32 * 1024;
        // warm-up
        for (int i = 0; i < 8; ++i) umma_bf16(tb + 256, umma_desc_kmajor_sw128(sa), umma_desc_kmajor_sw128(sb), umma_idesc_bf16(128, 64, 0, 0), i);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        tc_fence_after();
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t d = tb + 256 + (k % NACC) * (N <= 128 ? 128 : 0);
                if (MODE == 0) umma_bf16(d, umma_desc_kmajor_sw128(sa + (k % 4) * 32), umma_desc_kmajor_sw128(sb + (k % 4) * 32), idesc, 1);
                else if (MODE == 1) umma_bf16(d, umma_desc_kmajor_sw128(sa + (k % 4) * 32), umma_desc_mnmajor_sw128(sb + k * 2048, 16384, 1024), idesc, 1);
                else if (MODE == 2) umma_bf16_ts(d, tb + k * 8, umma_desc_mnmajor_sw128(sb + k * 2048, 16384, 1024), idesc, 1);
                else umma_bf16_ts(d, tb + k * 8, umma_desc_kmajor_sw128(sb + (k % 4) * 32), idesc, 1);
            }
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 1);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int MODE, int NACC>
void run(const char* name, long long* dout, int grid) {
    const int reps = 64;
    cudaFuncSetAttribute(bench<N, MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 1024);
    bench<N, MODE, NACC><<<grid, 128, 64 * 1024 + 1024>>>(dout, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s grid %3d: issue %7.1f cyc/mma, complete %7.1f cyc/mma  (%s)\n", name, grid, h[0] / (reps * 8.0), h[1] / (reps * 8.0),
           cudaGetErrorString(e));
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 16);
    for (int grid : {1, 148}) {
        run<128, 0, 1>("SS  N=128 Kmaj/Kmaj  1 acc", dout, grid);
        run<64, 0, 1>("SS  N=64  Kmaj/Kmaj  1 acc", dout, grid);
        run<256, 0, 1>("SS  N=256 Kmaj/Kmaj  1 acc", dout, grid);
        run<48, 1, 1>("SS  N=48  Kmaj/MNmaj 1 acc", dout, grid);
        run<48, 2, 1>("TS  N=48  tmem/MNmaj 1 acc", dout, grid);
        run<48, 2, 2>("TS  N=48  tmem/MNmaj 2 acc", dout, grid);
        run<64, 2, 1>("TS  N=64  tmem/MNmaj 1 acc", dout, grid);
        run<128, 2, 1>("TS  N=128 tmem/MNmaj 1 acc", dout, grid);
        run<48, 3, 1>("TS  N=48  tmem/Kmaj  1 acc", dout, grid);
        run<128, 0, 2>("SS  N=128 Kmaj/Kmaj  2 acc", dout, grid);
    }
    return 0;
}
