// Does the fp32 -> f16x2 / bf16x2 pack conversion share the MUFU (XU) pipe with ex2?  Cycles per warp for
// 8 independent chains of: cvt only, ex2.f16x2 only, the softmax inner step (2 FFMA + cvt + ex2.f16x2), and variants.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8], b[8];
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); b[i] = a[i] * 0.5f; h[i] = 0xB800B800u + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {                      // cvt only
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
                a[i] += __uint_as_float(h[i]) * 1e-30f;
            } else if (MODE == 1) {               // ex2.f16x2 only
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
            } else if (MODE == 2) {               // softmax step: 2 FFMA + cvt + ex2.f16x2
                const float x0 = fmaf(a[i], 1.0001f, -0.001f), x1 = fmaf(b[i], 1.0001f, -0.001f);
                uint32_t p;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x1), "f"(x0));
                asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(h[i]) : "r"(p));
                a[i] = x0; b[i] = x1;
            } else if (MODE == 3) {               // bf16 path: 2 x ex2.f32 + cvt.bf16x2
                float e0, e1;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a[i]));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(b[i]));
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(e1), "f"(e0));
                a[i] += 1e-30f * e0; b[i] += 1e-30f * e1;
            } else if (MODE == 4) {               // fp16 arithmetic all the way: HFMA2 + ex2.f16x2 (no conversion)
                uint32_t p;
                asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(h[i]), "r"(0x3C003C00u), "r"(0x80008000u));
                asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(h[i]) : "r"(p));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + b[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char* names[] = {"cvt.rn.f16x2.f32 only", "ex2.approx.f16x2 only", "2 FFMA + cvt.f16x2 + ex2.f16x2", "2 ex2.f32 + cvt.bf16x2", "HFMA2 + ex2.f16x2 (no cvt)"};
    for (int mode = 0; mode < 5; ++mode)
        for (int warps : {4, 8}) {
            if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
            else if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
            else if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
            else if (mode == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
            else k<4><<<148, warps * 32>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-34s warps/SM %2d (%d per SMSP): %6.1f cycles per step per warp, %6.1f per SMSP\n", names[mode], warps, warps / 4,
                   (double)c / (iters * 8.0), (double)c / (iters * 8.0) / (warps / 4));
        }
    return 0;
}
