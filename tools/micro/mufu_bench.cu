// MUFU.EX2 throughput on one SM as a function of resident warps (f32 and packed f16x2).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0xB800B800u + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            else if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
            else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[i] = fmaf(a[i], 0.5f, -1.0f); a[i] += 0.25f; }   // MUFU + FFMA + FADD
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {1, 4, 8, 16, 32}) {
            if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
            else if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
            else k<2><<<148, warps * 32>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double ops = (double)iters * 8 * warps * 32 * (mode == 1 ? 2 : 1);
            printf("mode %d (%s) warps/SM %2d: %.2f exp/clk/SM  (%.1f cycles per warp-instruction per SMSP-warp)\n", mode,
                   mode == 0 ? "ex2.f32" : mode == 1 ? "ex2.f16x2" : "ex2.f32+ffma+fadd", warps, ops / c, (double)c / (iters * 8.0));
        }
    return 0;
}
