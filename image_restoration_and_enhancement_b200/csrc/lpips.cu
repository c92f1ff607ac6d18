// LPIPS (AlexNet) glue kernels: the last metric of the reference's evaluator (src/metrics.py:97-111, lpips.LPIPS(net='alex'))
// on the GPU.  The five feature convolutions run on the tcgen05 implicit-GEMM kernel (gemm.cu, RG_ACT_RELU); this file
// holds what sits between them:
//   * maxpool3x3s2_kernel     AlexNet's MaxPool2d(3, stride 2) on channels-last bf16 features (16-B vectors);
//   * lpips_layer_kernel      per feature level: unit-normalise both feature vectors of a pixel along C
//                             (x / (||x||_2 + 1e-10)), squared difference, weight by the non-negative 1x1 "lin" head,
//                             sum over channels and pixels.  One warp per pixel pair, fp32 arithmetic; every block writes
//                             ONE partial per image slice in a fixed order, the host adds the partials in block order in
//                             float64 -- the value of an image does not depend on the batch it was scored in.
#include "common.cuh"
#include "internal.h"

namespace rg {

// x bf16 [N,H,W,C] -> y bf16 [N,OH,OW,C], OH = (H-3)/2+1; C % 8 == 0
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const uint4* __restrict__ x, int N, int H, int W, int C8, int OH,
                                                           int OW, uint4* __restrict__ y) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)N * OH * OW * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8); long long r = i / C8;
        const int ow = (int)(r % OW); r /= OW;
        const int oh = (int)(r % OH); const int n = (int)(r / OH);
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const uint4 u = x[(((long long)n * H + (2 * oh + dy)) * W + (2 * ow + dx)) * C8 + c];
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_bf16x2(w4[e]);
                    m[2 * e] = fmaxf(m[2 * e], f.x); m[2 * e + 1] = fmaxf(m[2 * e + 1], f.y);
                }
            }
        y[i] = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
    }
}

// f0, f1: bf16 [N][HW][C] (post-ReLU features of the two images); lin: fp32 [C] (>= 0)
// partial[n][blockIdx.x] = sum over the block's pixels of sum_c lin[c] * (f0/(|f0|+eps) - f1/(|f1|+eps))^2
// grid = (blocks_per_image, N), 256 threads = 8 warps, a warp walks pixels blockIdx.x*8 + warp, + 8*gridDim.x, ...
__global__ void __launch_bounds__(256) lpips_layer_kernel(const __nv_bfloat16* __restrict__ f0, const __nv_bfloat16* __restrict__ f1,
                                                          const float* __restrict__ lin, int HW, int C,
                                                          float* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_warp[8];
    const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = C >> 3;                                  // 16-byte vectors per pixel (<= 48 for C <= 384)
    float acc = 0.f;                                         // this warp's pixels, added in pixel order
    for (int pix = blockIdx.x * 8 + warp; pix < HW; pix += 8 * (int)gridDim.x) {
        const uint4* a = reinterpret_cast<const uint4*>(f0 + ((long long)n * HW + pix) * C);
        const uint4* b = reinterpret_cast<const uint4*>(f1 + ((long long)n * HW + pix) * C);
        float va[2][8], vb[2][8];                            // up to two vectors per lane (C <= 512)
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int v = lane + 32 * t;
            uint4 ua = make_uint4(0u, 0u, 0u, 0u), ub = ua;
            if (v < nv) { ua = a[v]; ub = b[v]; }
            const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 fa = unpack_bf16x2(wa[e]), fb = unpack_bf16x2(wb[e]);
                va[t][2 * e] = fa.x; va[t][2 * e + 1] = fa.y; vb[t][2 * e] = fb.x; vb[t][2 * e + 1] = fb.y;
                sa += fa.x * fa.x + fa.y * fa.y; sb += fb.x * fb.x + fb.y * fb.y;
            }
        }
        const float ia = 1.0f / (sqrtf(warp_sum(sa)) + 1e-10f), ib = 1.0f / (sqrtf(warp_sum(sb)) + 1e-10f);
        float d = 0.f;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int v = lane + 32 * t;
            if (v < nv) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float df = va[t][e] * ia - vb[t][e] * ib;
                    d += __ldg(lin + v * 8 + e) * df * df;
                }
            }
        }
        acc += warp_sum(d);
    }
    if (lane == 0) s_warp[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_warp[w];          // fixed warp order
        partial[(long long)n * gridDim.x + blockIdx.x] = t;
    }
}

}  // namespace rg

using namespace rg;

extern "C" int rg_maxpool3x3s2(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, void* y, rg_stream_t stream) {
    if (!x || !y || C % 8 || H < 3 || W < 3) return set_error(RG_ERR_ARG, "maxpool3x3s2: bad argument");
    const int OH = (H - 3) / 2 + 1, OW = (W - 3) / 2 + 1;
    const long long total = (long long)N * OH * OW * (C / 8);
    long long g = (total + 255) / 256;
    if (g > 148LL * 16) g = 148LL * 16;
    launch_kernel(maxpool3x3s2_kernel, dim3((unsigned)g), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                  reinterpret_cast<const uint4*>(x), N, H, W, C / 8, OH, OW, reinterpret_cast<uint4*>(y));
    count_launch();
    return check_launch("maxpool3x3s2_kernel");
}

extern "C" int rg_lpips_layer_blocks(int32_t HW) {
    // per-image geometry only (never the batch size): 8 pixels per block pass, at most RG_LPIPS_MAX_BLOCKS blocks
    int b = (HW + 63) / 64;
    return b < 1 ? 1 : (b > RG_LPIPS_MAX_BLOCKS ? RG_LPIPS_MAX_BLOCKS : b);
}

extern "C" int rg_lpips_layer(const void* f0, const void* f1, const float* lin, int32_t N, int32_t HW, int32_t C,
                              float* partial, rg_stream_t stream) {
    if (!f0 || !f1 || !lin || !partial || C % 8 || C > 512 || N < 1 || HW < 1)
        return set_error(RG_ERR_ARG, "lpips_layer: bad argument (C must be a multiple of 8, <= 512)");
    launch_kernel(lpips_layer_kernel, dim3((unsigned)rg_lpips_layer_blocks(HW), (unsigned)N), dim3(256), 0,
                  reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(f0),
                  reinterpret_cast<const __nv_bfloat16*>(f1), lin, HW, C, partial);
    count_launch();
    return check_launch("lpips_layer_kernel");
}
