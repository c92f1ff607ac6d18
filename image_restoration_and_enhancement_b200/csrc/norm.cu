// K7/K8: GroupNorm(+SiLU), LayerNorm and row softmax for channels-last activations (HBM-bound kernels).
//
// Layout: x[n][pixel][C], C contiguous.  Every thread owns 8 consecutive channels (one 16-byte bf16 vector,
// two float4 for fp32 input) for all pixels it visits, so the per-channel affine terms live in registers and
// consecutive threads touch consecutive 16/32-byte pieces of a pixel row (fully coalesced).
#include "common.cuh"
#include "internal.h"

namespace rg {

struct GnParams {
    const void* x1; const void* x2;
    int C1, C2, C;             // C = C1 + C2
    int in_f32;
    int N; long long HW;
    int groups, cpg; float eps;
    const float* gamma; const float* beta;
    float* sums;
    __nv_bfloat16* y; __nv_bfloat16* raw;
    int silu;
    int tpr;                   // threads per pixel row = C / 8
    int rpb;                   // pixel rows handled concurrently by a block
    long long pix_per_block;
    int n_blocks;              // statistics blocks per image (<= RG_GN_MAX_BLOCKS)
};

// Compile-time input dtype: the loads of an unrolled pixel loop stay independent (no branch between them).
// src / Cs / cs: the thread's source tensor, its channel count and the thread's first channel inside it
// (c0 is a multiple of 8 and C1 is a multiple of 8, so a vector never straddles the two sources).
template <bool IN_F32>
__device__ __forceinline__ void load8(const void* src, int Cs, int cs, const GnParams& p, int n, long long pix, float (&v)[8]) {
    const long long idx = ((long long)n * p.HW + pix) * Cs + cs;
    if (IN_F32) {
        const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + idx);
        const float4 a = s[0], b = s[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + idx);
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
    }
}

// workspace layout (floats): [RG_GN_MAX_IMAGES counters (int), zero between launches] [N][G][2] (mean, rstd) [N][blocks][G][2] partial (sum, sumsq)
__device__ __forceinline__ float* gn_final(const GnParams& p) { return p.sums + RG_GN_MAX_IMAGES; }
__device__ __forceinline__ float* gn_partials(const GnParams& p) { return gn_final(p) + (long long)p.N * p.groups * 2; }

// Deterministic, batch-invariant statistics: every block reduces a fixed slice of one image in a fixed order and
// writes one (sum, sumsq) pair per group to partials[n][block][group]; the apply kernel combines the partials in
// block order in fp64.  No atomics anywhere, so a run is bitwise reproducible and an image's result does not depend
// on how many other images share the launch.
// grid = (blocks_per_image, N); block = tpr * rpb threads
template <bool IN_F32>
__global__ void __launch_bounds__(512) gn_stats_kernel(const GnParams p) {
    extern __shared__ float s_red[];          // [rpb][tpr][16] thread partials, then [C][2] channel sums
    const int n = blockIdx.y;
    const int tc = threadIdx.x % p.tpr, tr = threadIdx.x / p.tpr;
    const int c0 = tc * 8;
    float s[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; ss[i] = 0.f; }
    const long long p0 = (long long)blockIdx.x * p.pix_per_block;
    long long p1 = p0 + p.pix_per_block;
    if (p1 > p.HW) p1 = p.HW;
    const bool first = c0 < p.C1;
    const void* const src = first ? p.x1 : p.x2;
    const int Cs = first ? p.C1 : p.C2, cs = first ? c0 : c0 - p.C1;
    long long pix = p0 + tr;
    for (; pix + 3LL * p.rpb < p1; pix += 4LL * p.rpb) {      // four independent loads in flight per thread
        float v0[8], v1[8], v2[8], v3[8];
        load8<IN_F32>(src, Cs, cs, p, n, pix, v0);
        load8<IN_F32>(src, Cs, cs, p, n, pix + p.rpb, v1);
        load8<IN_F32>(src, Cs, cs, p, n, pix + 2LL * p.rpb, v2);
        load8<IN_F32>(src, Cs, cs, p, n, pix + 3LL * p.rpb, v3);
#pragma unroll
        for (int i = 0; i < 8; ++i) {                          // same accumulation order as the scalar loop
            s[i] += v0[i]; ss[i] += v0[i] * v0[i];
            s[i] += v1[i]; ss[i] += v1[i] * v1[i];
            s[i] += v2[i]; ss[i] += v2[i] * v2[i];
            s[i] += v3[i]; ss[i] += v3[i] * v3[i];
        }
    }
    for (; pix < p1; pix += p.rpb) {
        float v[8];
        load8<IN_F32>(src, Cs, cs, p, n, pix, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += v[i]; ss[i] += v[i] * v[i]; }
    }
    float* mine = s_red + ((long long)tr * p.tpr + tc) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = ss[i]; }
    __syncthreads();
    float* chan = s_red + (long long)p.rpb * p.tpr * 16;      // [C][2]
    if (tr == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float a = 0.f, b = 0.f;
            for (int r = 0; r < p.rpb; ++r) {                 // fixed order over the row slots
                const float* o = s_red + ((long long)r * p.tpr + tc) * 16;
                a += o[i]; b += o[8 + i];
            }
            chan[(c0 + i) * 2] = a; chan[(c0 + i) * 2 + 1] = b;
        }
    }
    __syncthreads();
    float* partials = gn_partials(p);
    if ((int)threadIdx.x < p.groups) {
        const int g = threadIdx.x;
        float a = 0.f, b = 0.f;
        for (int c = g * p.cpg; c < (g + 1) * p.cpg; ++c) { a += chan[c * 2]; b += chan[c * 2 + 1]; }
        float* dst = partials + (((long long)n * gridDim.x + blockIdx.x) * p.groups + g) * 2;
        dst[0] = a; dst[1] = b;
    }
    // the block that finishes last for this image combines the partials in block order (fp64): the counter only
    // elects WHO does it, so the result does not depend on arrival order
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int* counter = reinterpret_cast<int*>(p.sums) + n;
        const int done = atomicAdd(counter, 1);
        is_last = (done == (int)gridDim.x - 1);
        if (is_last) *counter = 0;                            // self-resetting
    }
    __syncthreads();
    if (is_last && (int)threadIdx.x < p.groups) {
        __threadfence();
        const int g = threadIdx.x;
        const volatile float* part = partials + ((long long)n * gridDim.x * p.groups + g) * 2;
        double sum = 0.0, sq = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) { sum += part[(long long)b * p.groups * 2]; sq += part[(long long)b * p.groups * 2 + 1]; }
        const double inv_cnt = 1.0 / ((double)p.cpg * (double)p.HW);
        const double m = sum * inv_cnt;
        const double var = fmax(sq * inv_cnt - m * m, 0.0);
        float* fin = gn_final(p) + ((long long)n * p.groups + g) * 2;
        fin[0] = (float)m;
        fin[1] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
}

template <bool IN_F32>
__global__ void __launch_bounds__(512) gn_apply_kernel(const GnParams p) {
    const int n = blockIdx.y;
    const int tc = threadIdx.x % p.tpr, tr = threadIdx.x / p.tpr;
    if (tr >= p.rpb) return;
    const int c0 = tc * 8;
    float sc[8], sh[8];
    const float* fin = gn_final(p) + (long long)n * p.groups * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + i, g = c / p.cpg;
        const float mean = fin[g * 2], rstd = fin[g * 2 + 1];
        sc[i] = rstd * p.gamma[c];
        sh[i] = p.beta[c] - mean * sc[i];
    }
    const long long p0 = (long long)blockIdx.x * p.pix_per_block;
    long long p1 = p0 + p.pix_per_block;
    if (p1 > p.HW) p1 = p.HW;
    auto emit = [&](long long pix, float (&v)[8]) {
        const long long o = ((long long)n * p.HW + pix) * p.C + c0;
        if (p.raw) {
            *reinterpret_cast<uint4*>(p.raw + o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                               pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float y = v[i] * sc[i] + sh[i];
            v[i] = p.silu ? silu_f(y) : y;
        }
        *reinterpret_cast<uint4*>(p.y + o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                         pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    };
    const bool first = c0 < p.C1;
    const void* const src = first ? p.x1 : p.x2;
    const int Cs = first ? p.C1 : p.C2, cs = first ? c0 : c0 - p.C1;
    long long pix = p0 + tr;
    for (; pix + 3LL * p.rpb < p1; pix += 4LL * p.rpb) {      // four independent loads in flight per thread
        float v0[8], v1[8], v2[8], v3[8];
        load8<IN_F32>(src, Cs, cs, p, n, pix, v0);
        load8<IN_F32>(src, Cs, cs, p, n, pix + p.rpb, v1);
        load8<IN_F32>(src, Cs, cs, p, n, pix + 2LL * p.rpb, v2);
        load8<IN_F32>(src, Cs, cs, p, n, pix + 3LL * p.rpb, v3);
        emit(pix, v0); emit(pix + p.rpb, v1); emit(pix + 2LL * p.rpb, v2); emit(pix + 3LL * p.rpb, v3);
    }
    for (; pix < p1; pix += p.rpb) {
        float v[8];
        load8<IN_F32>(src, Cs, cs, p, n, pix, v);
        emit(pix, v);
    }
}

static int fill_gn(const rg_gn_t* g, GnParams& p, dim3& grid, int& threads) {
    if (!g || !g->x1 || !g->gamma || !g->beta || !g->sums) return set_error(RG_ERR_ARG, "groupnorm: null pointer");
    const int C = g->C1 + g->C2;
    if (g->C1 % 8 || g->C2 % 8 || C % g->groups || g->groups > 64 || C / 8 > 512)
        return set_error(RG_ERR_ARG, "groupnorm: channels must be multiples of 8 (<= 4096) and divisible by groups");
    if (g->C2 && !g->x2) return set_error(RG_ERR_ARG, "groupnorm: C2 without x2");
    if (g->N < 1 || g->N > RG_GN_MAX_IMAGES) return set_error(RG_ERR_ARG, "groupnorm: batch size out of range");
    p.x1 = g->x1; p.x2 = g->x2; p.C1 = g->C1; p.C2 = g->C2; p.C = C;
    p.in_f32 = g->in_dtype == RG_DT_F32;
    p.N = g->N; p.HW = g->HW; p.groups = g->groups; p.cpg = C / g->groups; p.eps = g->eps;
    p.gamma = g->gamma; p.beta = g->beta; p.sums = g->sums;
    p.y = reinterpret_cast<__nv_bfloat16*>(g->y); p.raw = reinterpret_cast<__nv_bfloat16*>(g->raw);
    p.silu = g->silu;
    p.tpr = C / 8;
    p.rpb = 256 / p.tpr; if (p.rpb < 1) p.rpb = 1;
    threads = p.tpr * p.rpb;
    // The split depends on (HW, C) only -- never on N -- so an image's statistics are reduced in the same order
    // whatever the batch size: at most RG_GN_MAX_BLOCKS blocks per image, at least 8 pixels per row slot.
    long long ppb = (g->HW + RG_GN_MAX_BLOCKS - 1) / RG_GN_MAX_BLOCKS;
    const long long min_ppb = 8LL * p.rpb;
    if (ppb < min_ppb) ppb = min_ppb;
    ppb = (ppb + p.rpb - 1) / p.rpb * p.rpb;
    p.pix_per_block = ppb;
    p.n_blocks = (int)((g->HW + ppb - 1) / ppb);
    grid = dim3((unsigned)p.n_blocks, (unsigned)g->N);
    return RG_OK;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per row, the row lives in registers between the two reduction passes, and the NEXT row of the warp is
// already in flight while the current one is reduced (the kernel is a pure HBM stream: fp32 in, bf16 out).
// EPL = elements per lane = C / 32 (10 / 20 / 40 for SD-1.5), loaded as EPL/VEC vectors of VEC floats.
template <bool IN_F32, int EPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ x_, long long rows,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __nv_bfloat16* __restrict__ y) {
    constexpr int C = EPL * 32;
    constexpr int VEC = (EPL % 4 == 0) ? 4 : 2;          // floats per vector access
    constexpr int NV = EPL / VEC;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    constexpr bool AFFINE_IN_REGS = EPL <= 20;           // wider rows re-read gamma / beta from L1 per row
    float g[AFFINE_IN_REGS ? EPL : 1], bt[AFFINE_IN_REGS ? EPL : 1];
    if (AFFINE_IN_REGS) {
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                g[j * VEC + e] = __ldg(gamma + (j * 32 + lane) * VEC + e);
                bt[j * VEC + e] = __ldg(beta + (j * 32 + lane) * VEC + e);
            }
    }
    auto load_row = [&](long long row, float (&v)[EPL]) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const long long idx = row * C + (j * 32 + lane) * VEC;
            if (IN_F32) {
                const float* src = reinterpret_cast<const float*>(x_) + idx;
                if (VEC == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(src);
                    v[j * VEC] = t.x; v[j * VEC + 1] = t.y; v[j * VEC + 2] = t.z; v[j * VEC + 3] = t.w;
                } else {
                    const float2 t = *reinterpret_cast<const float2*>(src);
                    v[j * VEC] = t.x; v[j * VEC + 1] = t.y;
                }
            } else {
                const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(x_) + idx;
                if (VEC == 4) {
                    const uint2 u = *reinterpret_cast<const uint2*>(src);
                    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
                    v[j * VEC] = a.x; v[j * VEC + 1] = a.y; v[j * VEC + 2] = b.x; v[j * VEC + 3] = b.y;
                } else {
                    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src));
                    v[j * VEC] = a.x; v[j * VEC + 1] = a.y;
                }
            }
        }
    };
    float cur[EPL], nxt[EPL];
    if (warp0 < rows) load_row(warp0, cur);
    for (long long row = warp0; row < rows; row += nwarps) {
        const bool more = row + nwarps < rows;
        if (more) load_row(row + nwarps, nxt);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) sum += cur[i];
        const float mean = warp_sum(sum) * (1.0f / (float)C);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) { const float d = cur[i] - mean; sq += d * d; }
        const float rstd = rsqrtf(warp_sum(sq) * (1.0f / (float)C) + eps);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            float o[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float ga = AFFINE_IN_REGS ? g[j * VEC + e] : __ldg(gamma + (j * 32 + lane) * VEC + e);
                const float be = AFFINE_IN_REGS ? bt[j * VEC + e] : __ldg(beta + (j * 32 + lane) * VEC + e);
                o[e] = (cur[j * VEC + e] - mean) * rstd * ga + be;
            }
            __nv_bfloat16* dst = y + row * C + (j * 32 + lane) * VEC;
            if (VEC == 4) *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
            else *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(o[0], o[1]);
        }
        if (more) {
#pragma unroll
            for (int i = 0; i < EPL; ++i) cur[i] = nxt[i];
        }
    }
}

// generic fallback (any C that is a multiple of 4, up to 1536): one row per warp, no prefetch
template <bool IN_F32>
__global__ void __launch_bounds__(256) layernorm_generic_kernel(const void* __restrict__ x_, long long rows, int C,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float eps, __nv_bfloat16* __restrict__ y) {
    constexpr int MAXV = 12;                     // 12 * 32 lanes * 4 = 1536 channels max
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nv = C >> 2;
    float4 v[MAXV];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int i = lane + j * 32;
        if (i < nv) {
            if (IN_F32) {
                v[j] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x_) + row * C)[i];
            } else {
                const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x_) + row * C)[i];
                const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
                v[j] = make_float4(a.x, a.y, b.x, b.y);
            }
            sum += v[j].x + v[j].y + v[j].z + v[j].w;
        }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int i = lane + j * 32;
        if (i < nv) {
            const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
            sq += a * a + b * b + c * c + d * d;
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int i = lane + j * 32;
        if (i < nv) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i);
            const float o0 = (v[j].x - mean) * rstd * g.x + b.x, o1 = (v[j].y - mean) * rstd * g.y + b.y;
            const float o2 = (v[j].z - mean) * rstd * g.z + b.z, o3 = (v[j].w - mean) * rstd * g.w + b.w;
            reinterpret_cast<uint2*>(y + row * C)[i] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        }
    }
}

// ---------------------------------------------------------------------------------------------- row softmax (bf16, in place)
__global__ void __launch_bounds__(256) softmax_rows_kernel(__nv_bfloat16* x, int cols, long long ld) {
    __shared__ float red[8];
    __nv_bfloat16* row = x + (long long)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float m = -INFINITY;
    for (int i = tid; i < cols; i += 256) m = fmaxf(m, __bfloat162float(row[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int i = tid; i < cols; i += 256) s += __expf(__bfloat162float(row[i]) - m);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    const float inv = 1.0f / s;
    for (int i = tid; i < cols; i += 256) row[i] = __float2bfloat16(__expf(__bfloat162float(row[i]) - m) * inv);
}

}  // namespace rg

using namespace rg;

extern "C" int rg_groupnorm_stats(const rg_gn_t* g, rg_stream_t stream) {
    GnParams p; dim3 grid; int threads;
    memset(&p, 0, sizeof(p));
    int rc = fill_gn(g, p, grid, threads);
    if (rc) return rc;
    const size_t smem = ((size_t)threads * 16 + (size_t)p.C * 2) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(gn_stats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(gn_stats_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_done = true;
    }
    if (p.in_f32) gn_stats_kernel<true><<<grid, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    else gn_stats_kernel<false><<<grid, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    count_launch();
    return check_launch("gn_stats_kernel");
}

extern "C" int rg_groupnorm_apply(const rg_gn_t* g, rg_stream_t stream) {
    GnParams p; dim3 grid; int threads;
    memset(&p, 0, sizeof(p));
    int rc = fill_gn(g, p, grid, threads);
    if (rc) return rc;
    if (!g->y) return set_error(RG_ERR_ARG, "groupnorm_apply: null output");
    if (p.in_f32) gn_apply_kernel<true><<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    else gn_apply_kernel<false><<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    count_launch();
    return check_launch("gn_apply_kernel");
}

template <bool IN_F32, int EPL>
static void launch_ln(const void* x, int64_t rows, const float* gamma, const float* beta, float eps, void* y, cudaStream_t s) {
    const int wpb = 8;
    long long blocks = (rows + wpb - 1) / wpb;
    const long long cap = (long long)sm_count() * 8;       // persistent: each warp walks rows with a one-row prefetch
    if (blocks > cap) blocks = cap;
    layernorm_kernel<IN_F32, EPL><<<(unsigned)blocks, wpb * 32, 0, s>>>(x, rows, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y));
}

extern "C" int rg_layernorm(const void* x, int32_t in_dtype, int64_t rows, int32_t C, const float* gamma,
                            const float* beta, float eps, void* y, rg_stream_t stream) {
    if (!x || !y || !gamma || !beta) return set_error(RG_ERR_ARG, "layernorm: null pointer");
    if (C % 4 || C > 1536) return set_error(RG_ERR_ARG, "layernorm: C must be a multiple of 4 and <= 1536");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const bool f32 = in_dtype == RG_DT_F32;
    if (C == 320) { if (f32) launch_ln<true, 10>(x, rows, gamma, beta, eps, y, s); else launch_ln<false, 10>(x, rows, gamma, beta, eps, y, s); }
    else if (C == 640) { if (f32) launch_ln<true, 20>(x, rows, gamma, beta, eps, y, s); else launch_ln<false, 20>(x, rows, gamma, beta, eps, y, s); }
    else if (C == 1280) { if (f32) launch_ln<true, 40>(x, rows, gamma, beta, eps, y, s); else launch_ln<false, 40>(x, rows, gamma, beta, eps, y, s); }
    else {
        const int wpb = 8;
        const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
        if (f32) layernorm_generic_kernel<true><<<grid, wpb * 32, 0, s>>>(x, rows, C, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y));
        else layernorm_generic_kernel<false><<<grid, wpb * 32, 0, s>>>(x, rows, C, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y));
    }
    count_launch();
    return check_launch("layernorm_kernel");
}

extern "C" int rg_softmax_rows(void* x, int64_t rows, int32_t cols, int64_t ld, rg_stream_t stream) {
    if (!x) return set_error(RG_ERR_ARG, "softmax_rows: null pointer");
    softmax_rows_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__nv_bfloat16*>(x), cols, ld);
    count_launch();
    return check_launch("softmax_rows_kernel");
}
