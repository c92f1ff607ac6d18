// K15: per-image PSNR / SSIM on u8 images resident in HBM, bit-exact against the reference's float64 CPU bookkeeping.
//
// The reference scores every prediction with scikit-image (/root/reference/src/metrics.py:87 psnr(gt, pred,
// data_range=255.0); :95 ssim(gt, pred, data_range=255.0, channel_axis=2)).  Those are float64 computations whose
// results depend on the ORDER of the additions, so these kernels reproduce the order, not just the formula:
//   * PSNR: the squared error sum is an integer (< 2^53), exact in any order -> uint64 atomics; the host divides and
//     takes the log.
//   * SSIM: scipy.ndimage.uniform_filter (7x7, mode="reflect") is two 1-D passes, each a RUNNING SUM along the line
//     (t += entering - leaving; out = t / 7).  Pass 1 (along H) runs over integer-valued data, so its sums are exact
//     and order-free: out = double(sum of 7 u8 / u8*u8 values) / 7.  Pass 2 (along W) accumulates rounding errors
//     sequentially, so one thread owns one image row and walks it left to right with the five running sums
//     (x, y, xx, yy, xy) in registers, evaluating the SSIM expression per pixel with explicitly rounded operations
//     (no FMA contraction) in the order of the numpy expression.
//   * the mean over the cropped map: numpy reduces a non-contiguous 2-D view in buffer-sized chunks of whole rows
//     (floor(8192 / width) rows), each chunk with its pairwise summation (8 strided accumulators per <=128-element
//     leaf, leaves combined by the recursive halving rule), chunk sums added sequentially.  ssim_chunk_sum_kernel
//     reproduces the leaf / tree structure; the host adds the (few) chunk sums in order and divides.
// tests/test_kernels_gpu.py checks bit equality against image_restoration_and_enhancement_b200.metrics (numpy/scipy).
#include "common.cuh"
#include "internal.h"

namespace rg {

constexpr int SSIM_WIN = 7;
constexpr int SSIM_PAD = 3;
constexpr int NPY_BUFSIZE = 8192;      // numpy's default iterator buffer, in elements
constexpr int PW_BLOCK = 128;          // numpy pairwise-summation leaf size
constexpr int MAX_LEAVES = 256;

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sse_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                     unsigned long long* __restrict__ sse, long long elems) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y;
    const uint8_t* pa = a + (long long)n * elems;
    const uint8_t* pb = b + (long long)n * elems;
    unsigned long long acc = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb)) & 15) == 0;
    const long long nvec = vec ? elems / 16 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const uint4 va = reinterpret_cast<const uint4*>(pa)[i];
        const uint4 vb = reinterpret_cast<const uint4*>(pb)[i];
        const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int d = (int)((wa[w] >> (8 * k)) & 255u) - (int)((wb[w] >> (8 * k)) & 255u);
                s += (uint32_t)(d * d);
            }
        acc += s;
    }
    for (long long i = nvec * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems;
         i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)pa[i] - (int)pb[i];
        acc += (unsigned long long)(d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        atomicAdd(sse + n, t);
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct Five {
    double v[5];
};

// pass 1 of uniform_filter at (rows h-3..h+3, column col): exact integer sums / 7
__device__ __forceinline__ Five vertical7(const uint8_t* __restrict__ pa, const uint8_t* __restrict__ pb, int h, int col,
                                          int W, int C) {
    int sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
    for (int k = -SSIM_PAD; k <= SSIM_PAD; ++k) {
        const long long o = ((long long)(h + k) * W + col) * C;
        const int x = pa[o], y = pb[o];
        sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;
    }
    Five f;
    f.v[0] = __ddiv_rn((double)sx, 7.0);
    f.v[1] = __ddiv_rn((double)sy, 7.0);
    f.v[2] = __ddiv_rn((double)sxx, 7.0);
    f.v[3] = __ddiv_rn((double)syy, 7.0);
    f.v[4] = __ddiv_rn((double)sxy, 7.0);
    return f;
}

// One thread per (image, channel, cropped row).  smap: [N][C][H-6][W-6] float64.
__global__ void __launch_bounds__(64) ssim_map_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                      double* __restrict__ smap, int H, int W, int C, double c1,
                                                      double c2, double cov_norm) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int HC = H - 2 * SSIM_PAD, WC = W - 2 * SSIM_PAD;
    if (r >= HC) return;
    const int c = blockIdx.y, n = blockIdx.z;
    const int h = r + SSIM_PAD;
    const uint8_t* pa = a + (long long)n * H * W * C + c;
    const uint8_t* pb = b + (long long)n * H * W * C + c;
    double* out = smap + (((long long)n * C + c) * HC + r) * WC;

    // the line buffer scipy builds: ext[e] = v[e-3], mirrored at the left edge: v2 v1 v0 | v0 v1 v2 v3
    Five ring[SSIM_WIN];
    {
        const Five v0 = vertical7(pa, pb, h, 0, W, C), v1 = vertical7(pa, pb, h, 1, W, C);
        const Five v2 = vertical7(pa, pb, h, 2, W, C), v3 = vertical7(pa, pb, h, 3, W, C);
        ring[0] = v2; ring[1] = v1; ring[2] = v0; ring[3] = v0; ring[4] = v1; ring[5] = v2; ring[6] = v3;
    }
    double t[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < SSIM_WIN; ++e) s = __dadd_rn(s, ring[e].v[m]);
        t[m] = s;
    }
    const int lmax = W - 1 - SSIM_PAD;                 // last column of the crop
    for (int base = 1; base <= lmax; base += SSIM_WIN) {
#pragma unroll
        for (int j = 0; j < SSIM_WIN; ++j) {
            const int l = base + j;
            if (l <= lmax) {
                const Five e = vertical7(pa, pb, h, l + SSIM_PAD, W, C);        // entering column l+3 (< W)
                double u[5];
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    t[m] = __dadd_rn(t[m], __dsub_rn(e.v[m], ring[j].v[m]));    // leaving = ext[l-1]
                    ring[j].v[m] = e.v[m];
                    u[m] = __ddiv_rn(t[m], 7.0);
                }
                if (l >= SSIM_PAD) {
                    const double ux = u[0], uy = u[1];
                    const double vx = __dmul_rn(cov_norm, __dsub_rn(u[2], __dmul_rn(ux, ux)));
                    const double vy = __dmul_rn(cov_norm, __dsub_rn(u[3], __dmul_rn(uy, uy)));
                    const double vxy = __dmul_rn(cov_norm, __dsub_rn(u[4], __dmul_rn(ux, uy)));
                    const double A1 = __dadd_rn(__dmul_rn(__dmul_rn(2.0, ux), uy), c1);
                    const double A2 = __dadd_rn(__dmul_rn(2.0, vxy), c2);
                    const double B1 = __dadd_rn(__dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)), c1);
                    const double B2 = __dadd_rn(__dadd_rn(vx, vy), c2);
                    out[l - SSIM_PAD] = __ddiv_rn(__dmul_rn(A1, A2), __dmul_rn(B1, B2));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// numpy's pairwise summation of one buffer chunk.  Thread 0 lists the leaves of the recursion (left to right), eight
// lanes per leaf run the eight strided accumulators, thread 0 folds the leaf sums back up the same recursion.
__device__ void pw_enumerate(int off, int n, int* leaf_off, int* leaf_n, int& count) {
    if (n <= PW_BLOCK) {
        if (count < MAX_LEAVES) { leaf_off[count] = off; leaf_n[count] = n; }
        ++count;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    pw_enumerate(off, n2, leaf_off, leaf_n, count);
    pw_enumerate(off + n2, n - n2, leaf_off, leaf_n, count);
}
__device__ double pw_combine(int n, const double* leaf_sum, int& idx) {
    if (n <= PW_BLOCK) return leaf_sum[idx++];
    int n2 = n / 2;
    n2 -= n2 % 8;
    const double l = pw_combine(n2, leaf_sum, idx);
    const double r = pw_combine(n - n2, leaf_sum, idx);
    return __dadd_rn(l, r);
}

// grid (chunks, C, N), 256 threads.  chunk ck covers cropped rows [ck*rows_per_chunk, ...), contiguous in smap.
__global__ void __launch_bounds__(256) ssim_chunk_sum_kernel(const double* __restrict__ smap,
                                                             double* __restrict__ chunk_sums, int HC, int WC,
                                                             int rows_per_chunk) {
    pdl_trigger();
    pdl_wait();
    __shared__ int leaf_off[MAX_LEAVES], leaf_n[MAX_LEAVES];
    __shared__ double leaf_sum[MAX_LEAVES];
    __shared__ int n_leaves;
    const int ck = blockIdx.x, c = blockIdx.y, n = blockIdx.z, C = gridDim.y;
    const int row0 = ck * rows_per_chunk;
    const int rows = min(rows_per_chunk, HC - row0);
    const int total = rows * WC;
    const double* src = smap + (((long long)n * C + c) * HC + row0) * WC;
    if (threadIdx.x == 0) {
        int count = 0;
        pw_enumerate(0, total, leaf_off, leaf_n, count);
        n_leaves = count;
    }
    __syncthreads();
    const int lane8 = threadIdx.x & 7, group = threadIdx.x >> 3, groups = blockDim.x >> 3;
    const int nl = min(n_leaves, MAX_LEAVES);
    for (int base = 0; base < nl; base += groups) {        // uniform trip count: shuffles need the full warp
        const int lf = base + group;
        const bool live = lf < nl;
        const double* p = src + (live ? leaf_off[lf] : 0);
        const int ln = live ? leaf_n[lf] : 0;
        int full = 0;
        double r = 0.0;
        if (ln >= 8) {
            full = ln - (ln % 8);
            r = p[lane8];
            for (int i = 8; i < full; i += 8) r = __dadd_rn(r, p[i + lane8]);
        }
        __syncwarp();
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));          // (r0+r1) (r2+r3) (r4+r5) (r6+r7)
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));          // ((r0+r1)+(r2+r3)) ((r4+r5)+(r6+r7))
        r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        double res = r;                                                // 0.0 for a leaf shorter than 8
        if (lane8 == 0)
            for (int i = full; i < ln; ++i) res = __dadd_rn(res, p[i]);
        if (live && lane8 == 0) leaf_sum[lf] = res;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int idx = 0;
        chunk_sums[((long long)n * C + c) * gridDim.x + ck] = pw_combine(total, leaf_sum, idx);
    }
}

}  // namespace rg

using namespace rg;

extern "C" int rg_metrics_sse_u8(const uint8_t* pred, const uint8_t* gt, int32_t N, int64_t elems_per_image,
                                 uint64_t* sse, rg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!pred || !gt || !sse || N <= 0 || N > 65535 || elems_per_image <= 0)
        return set_error(RG_ERR_ARG, "rg_metrics_sse_u8: bad arguments");
    cudaError_t e = cudaMemsetAsync(sse, 0, sizeof(uint64_t) * N, stream);
    if (e != cudaSuccess) return set_cuda_error(e, "rg_metrics_sse_u8 memset");
    const long long per_block = 256LL * 16 * 4;
    int bx = (int)((elems_per_image + per_block - 1) / per_block);
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    launch_kernel(sse_u8_kernel, dim3(dim3(bx, N)), dim3(256), 0, stream, pred, gt, reinterpret_cast<unsigned long long*>(sse),
                                                   (long long)elems_per_image);
    count_launch();
    return check_launch("sse_u8_kernel");
}

extern "C" int rg_metrics_ssim_chunks(int32_t H, int32_t W) {
    const int HC = H - 2 * SSIM_PAD, WC = W - 2 * SSIM_PAD;
    if (HC < 1 || WC < 1 || WC > NPY_BUFSIZE) return -1;
    const int rows = NPY_BUFSIZE / WC;
    return (HC + rows - 1) / rows;
}

extern "C" int rg_metrics_ssim_u8(const uint8_t* pred, const uint8_t* gt, int32_t N, int32_t H, int32_t W, int32_t C,
                                  double c1, double c2, double cov_norm, double* smap_ws, double* chunk_sums,
                                  rg_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!pred || !gt || !smap_ws || !chunk_sums || N <= 0 || N > 65535 || C <= 0 || C > 65535)
        return set_error(RG_ERR_ARG, "rg_metrics_ssim_u8: bad arguments");
    if (H < SSIM_WIN || W < SSIM_WIN) return set_error(RG_ERR_ARG, "rg_metrics_ssim_u8: image smaller than the 7x7 window");
    const int chunks = rg_metrics_ssim_chunks(H, W);
    if (chunks < 1) return set_error(RG_ERR_ARG, "rg_metrics_ssim_u8: width - 6 exceeds numpy's 8192-element buffer");
    const int HC = H - 2 * SSIM_PAD, WC = W - 2 * SSIM_PAD;
    // gt is im1 (x), pred is im2 (y): the expression is symmetric operation by operation, the order is kept anyway
    launch_kernel(ssim_map_kernel, dim3(dim3((HC + 63) / 64, C, N)), dim3(64), 0, stream, gt, pred, smap_ws, H, W, C, c1, c2, cov_norm);
    count_launch();
    int rc = check_launch("ssim_map_kernel");
    if (rc != RG_OK) return rc;
    launch_kernel(ssim_chunk_sum_kernel, dim3(dim3(chunks, C, N)), dim3(256), 0, stream, smap_ws, chunk_sums, HC, WC, NPY_BUFSIZE / WC);
    count_launch();
    return check_launch("ssim_chunk_sum_kernel");
}
