// Micro-benchmark: the attention kernel's per-key-tile tcgen05 sequence on an otherwise idle SM:
//   S0 = Q0 K^T (3 x M128 N128 K16, SS), S1, O0 += P0 V (8 x M128 N48 K16, TS, V MN-major), O1 += P1 V, with the commits
// the kernel issues.  Variants switch the commits off / keep one shape only, to see what the sequence itself costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I image_restoration_and_enhancement_b200/csrc tools/micro/mma_seq_bench.cu -o tools/micro/mma_seq_bench.bin
#include <cstdio>
#include "common.cuh"
using namespace rg;

// VAR bit 0: commits after each group of MMAs; bit 1: skip S; bit 2: skip PV; bit 3: PV as SS (A from smem); bit 4: PV N=64
template <int VAR>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_smem;
    if (warp == 1 && lane == 0) {
        constexpr int NPV = (VAR & 16) ? 64 : 48;
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, NPV, 0, 1);
        const uint32_t sq = smem_u32(smem), sk = sq + 32 * 1024, sv = sk + 16 * 1024;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (!(VAR & 2)) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks)
                        umma_bf16(tb + t * 128, umma_desc_kmajor_sw128(sq + t * 16384 + ks * 32), umma_desc_kmajor_sw128(sk + ks * 32), idesc_s, ks != 0);
                    if (VAR & 1) umma_commit(&bars[t]);
                }
            }
            if (!(VAR & 4)) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint64_t bdesc = umma_desc_mnmajor_sw128(sv + ks * 2048, 16384, 1024);
                        if (VAR & 8) umma_bf16(tb + 384 + t * 64, umma_desc_kmajor_sw128(sq + t * 16384 + (ks % 4) * 32), bdesc, idesc_o, 1);
                        else umma_bf16_ts(tb + 384 + t * 64, tb + 256 + t * 64 + ks * 8, bdesc, idesc_o, 1);
                    }
                    if (VAR & 1) umma_commit(&bars[2 + t]);
                }
                if (VAR & 1) umma_commit(&bars[4]);
            }
        }
        const long long t1 = clock64();
        umma_commit(&bars[7]);
        mbar_wait(&bars[7], 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int VAR>
void run(const char* name, long long* dout, int grid) {
    const int reps = 64;
    cudaFuncSetAttribute(bench<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 + 1024);
    bench<VAR><<<grid, 128, 96 * 1024 + 1024>>>(dout, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-58s grid %3d: issue %7.0f, complete %7.0f cycles per key tile  (%s)\n", name, grid, h[0] / (double)reps, h[1] / (double)reps,
           cudaGetErrorString(e));
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 16);
    for (int grid : {1, 148}) {
        run<1>("S0 S1 PV0 PV1 with the kernel's commits", dout, grid);
        run<0>("S0 S1 PV0 PV1 without commits", dout, grid);
        run<1 | 4>("S0 S1 only (6 MMAs N=128 SS) + commits", dout, grid);
        run<1 | 2>("PV0 PV1 only (16 MMAs N=48 TS MN-major) + commits", dout, grid);
        run<2>("PV0 PV1 only, no commits", dout, grid);
        run<1 | 2 | 8>("PV0 PV1 only as SS (A smem) + commits", dout, grid);
        run<1 | 2 | 16>("PV0 PV1 only N=64 TS + commits", dout, grid);
        run<1 | 16>("S0 S1 PV0 PV1 with PV N=64 + commits", dout, grid);
    }
    return 0;
}
