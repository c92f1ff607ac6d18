// K7/K8: GroupNorm(+SiLU), LayerNorm and row softmax for channels-last activations (HBM-bound kernels).
//
// Layout: x[n][pixel][C], C contiguous.  Every thread owns 8 consecutive channels (one 16-byte bf16 vector,
// two float4 for fp32 input) for all pixels it visits, so the per-channel affine terms live in registers and
// consecutive threads touch consecutive 16/32-byte pieces of a pixel row (fully coalesced).
#include <type_traits>
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

// the same load split in two: raw 16/32-byte vectors first (so that an unrolled loop keeps only the packed words of
// its U in-flight loads live), conversion to fp32 when they are consumed
template <bool IN_F32> struct Raw8 { uint4 a; uint4 b; };
template <> struct Raw8<false> { uint4 a; };
template <bool IN_F32>
__device__ __forceinline__ Raw8<IN_F32> load_raw8(const unsigned char* ptr, bool ok) {
    Raw8<IN_F32> r;
    r.a = make_uint4(0u, 0u, 0u, 0u);
    if constexpr (IN_F32) {
        r.b = make_uint4(0u, 0u, 0u, 0u);
        if (ok) {
            const uint4* s = reinterpret_cast<const uint4*>(ptr);
            r.a = s[0]; r.b = s[1];
        }
    } else {
        if (ok) r.a = *reinterpret_cast<const uint4*>(ptr);
    }
    return r;
}
// byte address of the thread's 8 channels at pixel `pix` of image n, and the byte stride between its pixel visits
template <bool IN_F32>
__device__ __forceinline__ const unsigned char* pix_ptr(const void* src, int Cs, int cs, const GnParams& p, int n, long long pix) {
    return reinterpret_cast<const unsigned char*>(src) + (((long long)n * p.HW + pix) * Cs + cs) * (IN_F32 ? 4 : 2);
}
template <bool IN_F32>
__device__ __forceinline__ void unpack8(const Raw8<IN_F32>& r, float (&v)[8]) {
    if constexpr (IN_F32) {
        v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y); v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
        v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y); v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
    } else {
        float2 f;
        f = unpack_bf16x2(r.a.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(r.a.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(r.a.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(r.a.w); v[6] = f.x; v[7] = f.y;
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
    pdl_trigger();
    pdl_wait();
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
    // U independent loads in flight per thread (8 x 16 B for bf16, 4 x 32 B for fp32); the accumulation below is in
    // pixel order, exactly the order of the scalar tail loop, so the unroll factor does not change a single bit
    constexpr int U = 8;
    for (; pix < p1; pix += (long long)U * p.rpb) {
        // ragged end: slots past the slice load nothing and contribute +0 (no serial tail loop -- its dependent
        // load -> add chain cost one full memory latency per pixel)
        Raw8<IN_F32> raw[U];
        const unsigned char* ptr = pix_ptr<IN_F32>(src, Cs, cs, p, n, pix);
        const long long step = (long long)p.rpb * Cs * (IN_F32 ? 4 : 2);
        const int left = (int)((p1 - pix + p.rpb - 1) / p.rpb);   // visits of this thread that are still inside the slice
#pragma unroll
        for (int u = 0; u < U; ++u) {
            raw[u] = load_raw8<IN_F32>(ptr, u < left);
            ptr += step;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {                          // pixel order, channel by channel: the scalar order
            float v[8];
            unpack8<IN_F32>(raw[u], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += v[i]; ss[i] += v[i] * v[i]; }
        }
    }
    float* mine = s_red + ((long long)tr * p.tpr + tc) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mine[i] = s[i]; mine[8 + i] = ss[i]; }
    __syncthreads();
    // channel sums: output o = tc * 16 + k (k < 8: sum of channel tc*8+k, k >= 8: its sum of squares), rows added in
    // slot order; all threads take part, consecutive threads read consecutive words
    float* chan = s_red + (long long)p.rpb * p.tpr * 16;      // [C][2]
    const int row_words = p.tpr * 16;
    for (int o = threadIdx.x; o < row_words; o += blockDim.x) {
        float a = 0.f;
        for (int r = 0; r < p.rpb; ++r) a += s_red[(long long)r * row_words + o];
        const int k = o & 15;
        chan[((o >> 4) * 8 + (k & 7)) * 2 + (k >> 3)] = a;
    }
    __syncthreads();
    // group sums: four lanes per group, each adds a contiguous quarter of the group's channels in order, then
    // ((q0 + q1) + (q2 + q3)) by two butterfly steps -- a fixed tree
    float* partials = gn_partials(p);
    {
        const int g = threadIdx.x >> 2, q = threadIdx.x & 3;
        const bool live = g < p.groups;                       // blockDim >= 4 * groups is checked on the host
        float a = 0.f, b = 0.f;
        if (live) {
            const int lo = g * p.cpg + q * p.cpg / 4, hi = g * p.cpg + (q + 1) * p.cpg / 4;
            for (int c = lo; c < hi; ++c) { a += chan[c * 2]; b += chan[c * 2 + 1]; }
        }
        if ((int)(threadIdx.x & ~31u) < p.groups * 4) {       // warp-uniform: whole warps run the shuffles
            a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2);
        }
        if (live && q == 0) {
            float* dst = partials + (((long long)n * gridDim.x + blockIdx.x) * p.groups + g) * 2;
            dst[0] = a; dst[1] = b;
        }
    }
    // the block that finishes last for this image combines the partials in a fixed order (fp64): the counter only
    // elects WHO does it, so the result does not depend on arrival order
    __shared__ int is_last;
    __syncthreads();                                          // the block's partial writes happen-before thread 0 ...
    if (threadIdx.x == 0) {
        __threadfence();                                      // ... whose fence is cumulative: ONE gpu-scope fence per block
        int* counter = reinterpret_cast<int*>(p.sums) + n;    // (a fence in every thread was 7 % of the kernel's stalls)
        const int done = atomicAdd(counter, 1);
        is_last = (done == (int)gridDim.x - 1);
        if (is_last) *counter = 0;                            // self-resetting
    }
    __syncthreads();
    // LPG lanes per group (8 if the block has 8 * groups threads in full warps, else 4 -- a function of C and groups
    // only): lane q adds blocks [q*nb/LPG, (q+1)*nb/LPG) in block order -- independent L2 loads, eight in flight --
    // then a butterfly over the LPG lanes: a fixed tree
    const int lpg_shift = ((int)(blockDim.x & ~31u) >= p.groups * 8) ? 3 : 2;
    if (is_last && (int)(threadIdx.x & ~31u) < (p.groups << lpg_shift)) {
        __threadfence();
        const int lpg = 1 << lpg_shift;
        const int g = threadIdx.x >> lpg_shift, q = threadIdx.x & (lpg - 1), nb = (int)gridDim.x;
        double sum = 0.0, sq = 0.0;
        if (g < p.groups) {
            const float* part = partials + ((long long)n * nb * p.groups + g) * 2;
            const int b1 = (q + 1) * nb / lpg;
#pragma unroll 8
            for (int b = q * nb / lpg; b < b1; ++b) {
                const float2 v = __ldcg(reinterpret_cast<const float2*>(part + (long long)b * p.groups * 2));
                sum += v.x; sq += v.y;
            }
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1); sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2); sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        if (lpg_shift == 3) { sum += __shfl_xor_sync(0xffffffffu, sum, 4); sq += __shfl_xor_sync(0xffffffffu, sq, 4); }
        if (g < p.groups && q == 0) {
            const double inv_cnt = 1.0 / ((double)p.cpg * (double)p.HW);
            const double m = sum * inv_cnt;
            const double var = fmax(sq * inv_cnt - m * m, 0.0);
            float* fin = gn_final(p) + ((long long)n * p.groups + g) * 2;
            fin[0] = (float)m;
            fin[1] = (float)(1.0 / sqrt(var + (double)p.eps));
        }
    }
}

template <bool IN_F32>
__global__ void __launch_bounds__(512, IN_F32 ? 2 : 1) gn_apply_kernel(const GnParams p) {
    pdl_trigger();
    pdl_wait();
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
    constexpr int U = IN_F32 ? 4 : 8;                          // independent loads in flight per thread
    for (; pix < p1; pix += (long long)U * p.rpb) {
        Raw8<IN_F32> raw[U];
        const unsigned char* ptr = pix_ptr<IN_F32>(src, Cs, cs, p, n, pix);
        const long long step = (long long)p.rpb * Cs * (IN_F32 ? 4 : 2);
        const int left = (int)((p1 - pix + p.rpb - 1) / p.rpb);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            raw[u] = load_raw8<IN_F32>(ptr, u < left);
            ptr += step;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long px = pix + (long long)u * p.rpb;
            if (u < left) {
                float v[8];
                unpack8<IN_F32>(raw[u], v);
                emit(px, v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- one-pass GroupNorm
// Few-pixel levels (32x32 and below in the UNet): the slice of one (image, group) -- HW pixels x cpg channels -- fits in
// shared memory, so one block reads it ONCE, reduces it in a fixed order, normalises it from shared memory and writes
// the bf16 result: one launch and one HBM read instead of two launches and two reads.  No inter-block communication at
// all, so the result is deterministic and independent of the batch size by construction.  An item is 4 consecutive
// channels of one pixel (16 B fp32 / 8 B bf16); C1 is a multiple of 8, so an item never straddles the two sources.
// grid = (groups, N), NT = 256 threads, dynamic smem = HW * cpg * sizeof(input element).  (Tried and NOT kept: slices up to
// 200 KB with 512 threads, one block per SM -- the VAE's 64x64 x 512 bf16 level took 65.6 us against 33.2 us for the
// two-pass kernels at 8 images, profiles/r02_gn_onepass_experiment.txt; the UNet's 64x64 x 320 level has 10 channels per
// group, which the 4-channel items of this kernel do not tile, so it stays two-pass either way; a cluster-of-8 one-pass
// kernel for that level was built and measured too -- 70 us against 53.6 us at batch 16, profiles/r02_gn_cluster_experiment.txt.)
template <bool IN_F32, int NT>
__global__ void __launch_bounds__(NT) gn_fused_small_kernel(const GnParams p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char s_slice[];
    __shared__ double s_part[2][NT / 32];
    __shared__ float s_stat[2];
    __shared__ float s_scale[128], s_shift[128];
    using Vec = typename std::conditional<IN_F32, float4, uint2>::type;
    Vec* slice = reinterpret_cast<Vec*>(s_slice);
    const int g = blockIdx.x, n = blockIdx.y;
    const int ipp = p.cpg >> 2;                       // items per pixel
    const int items = (int)p.HW * ipp;
    const int cg0 = g * p.cpg;
    auto unpack = [](const Vec& v, float (&f)[4]) {
        if constexpr (IN_F32) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
        else { const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y); f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; }
    };
    auto src_of = [&](int it) -> const Vec* {
        const int pix = it / ipp, c = cg0 + (it - pix * ipp) * 4;
        const long long row = (long long)n * p.HW + pix;
        using Elem = typename std::conditional<IN_F32, float, __nv_bfloat16>::type;
        const Elem* e = c < p.C1 ? reinterpret_cast<const Elem*>(p.x1) + row * p.C1 + c
                                 : reinterpret_cast<const Elem*>(p.x2) + row * p.C2 + (c - p.C1);
        return reinterpret_cast<const Vec*>(e);
    };
    float s = 0.f, ss = 0.f;
    int it = threadIdx.x;
    constexpr int UF = 8;                             // loads in flight per thread; accumulation stays in item order
    for (; it < items; it += UF * NT) {
        Vec v[UF];
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            if constexpr (IN_F32) v[u] = make_float4(0.f, 0.f, 0.f, 0.f); else v[u] = make_uint2(0u, 0u);
            if (it + u * NT < items) v[u] = *src_of(it + u * NT);
        }
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            if (it + u * NT < items) slice[it + u * NT] = v[u];
            float f[4];
            unpack(v[u], f);                          // slots past the end hold +0
#pragma unroll
            for (int i = 0; i < 4; ++i) { s += f[i]; ss += f[i] * f[i]; }
        }
    }
    // fixed-order block reduction in fp64: butterfly inside each warp, the NT / 32 warp sums added in warp order
    double ds = s, dq = ss;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, o); dq += __shfl_xor_sync(0xffffffffu, dq, o); }
    if ((threadIdx.x & 31) == 0) { s_part[0][threadIdx.x >> 5] = ds; s_part[1][threadIdx.x >> 5] = dq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0, sq = 0.0;
        for (int w = 0; w < NT / 32; ++w) { sum += s_part[0][w]; sq += s_part[1][w]; }
        const double inv_cnt = 1.0 / ((double)p.cpg * (double)p.HW);
        const double m = sum * inv_cnt;
        const double var = fmax(sq * inv_cnt - m * m, 0.0);
        s_stat[0] = (float)m;
        s_stat[1] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
    __syncthreads();
    if ((int)threadIdx.x < p.cpg) {
        const float sc = s_stat[1] * p.gamma[cg0 + threadIdx.x];
        s_scale[threadIdx.x] = sc;
        s_shift[threadIdx.x] = p.beta[cg0 + threadIdx.x] - s_stat[0] * sc;
    }
    __syncthreads();
    for (it = threadIdx.x; it < items; it += NT) {
        const int pix = it / ipp, j = (it - pix * ipp) * 4;
        float f[4];
        unpack(slice[it], f);
        const long long o = ((long long)n * p.HW + pix) * p.C + cg0 + j;
        if (p.raw) *reinterpret_cast<uint2*>(p.raw + o) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float y = f[i] * s_scale[j + i] + s_shift[j + i];
            f[i] = p.silu ? silu_f(y) : y;
        }
        *reinterpret_cast<uint2*>(p.y + o) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
    }
}

constexpr size_t kGnFusedMaxSmem = 96 * 1024;        // two blocks per SM
// the choice depends on (HW, C, groups, dtype) only -- never on N -- so batch invariance is kept
static bool gn_use_fused(const rg_gn_t* g) {
    const int C = g->C1 + g->C2;
    if (C % g->groups) return false;
    const int cpg = C / g->groups;
    const size_t isz = g->in_dtype == RG_DT_F32 ? 4 : 2;
    return g->HW <= 1024 && cpg % 4 == 0 && cpg <= 128 && (size_t)g->HW * cpg * isz <= kGnFusedMaxSmem;
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
    if ((threads & ~31) < 4 * g->groups)
        return set_error(RG_ERR_ARG, "groupnorm: needs 4 * groups <= the block's full warps (groups <= 32 for C = 320)");
    // The split depends on (HW, C) only -- never on N -- so an image's statistics are reduced in the same order
    // whatever the batch size: 64 blocks per image (256 = RG_GN_MAX_BLOCKS for the VAE's >= 256x256-pixel levels, where 8
    // images x 64 blocks would leave half of the SMs' thread slots empty), at least 8 pixels per row slot.
    const long long max_blocks = g->HW >= 32768 ? RG_GN_MAX_BLOCKS : 64;
    long long ppb = (g->HW + max_blocks - 1) / max_blocks;
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
__global__ void __launch_bounds__(256, EPL <= 10 ? 4 : 0) layernorm_kernel(const void* __restrict__ x_, long long rows,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __nv_bfloat16* __restrict__ y) {
    pdl_trigger();
    pdl_wait();
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
    pdl_trigger();
    pdl_wait();
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
    pdl_trigger();
    pdl_wait();
    __shared__ float red[8];
    __nv_bfloat16* row = x + (long long)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float m = -INFINITY;
    for (int i = tid; i < cols; i += 256) m = fmaxf(m, op2f(row[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int i = tid; i < cols; i += 256) s += __expf(op2f(row[i]) - m);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    const float inv = 1.0f / s;
    for (int i = tid; i < cols; i += 256) row[i] = f2op(__expf(op2f(row[i]) - m) * inv);
}

}  // namespace rg

using namespace rg;

extern "C" int rg_groupnorm_stats(const rg_gn_t* g, rg_stream_t stream) {
    GnParams p; dim3 grid; int threads;
    memset(&p, 0, sizeof(p));
    int rc = fill_gn(g, p, grid, threads);
    if (rc) return rc;
    const size_t smem = ((size_t)threads * 16 + (size_t)p.C * 2) * sizeof(float);
    static std::atomic<bool> done_t[kMaxDevices], done_f[kMaxDevices];
    if ((rc = ensure_smem_attr(reinterpret_cast<const void*>(&gn_stats_kernel<true>), 96 * 1024, done_t, "cudaFuncSetAttribute(gn_stats_kernel)"))) return rc;
    if ((rc = ensure_smem_attr(reinterpret_cast<const void*>(&gn_stats_kernel<false>), 96 * 1024, done_f, "cudaFuncSetAttribute(gn_stats_kernel)"))) return rc;
    if (p.in_f32) launch_kernel<1>(gn_stats_kernel<true>, dim3(grid), dim3(threads), smem, reinterpret_cast<cudaStream_t>(stream), p);
    else launch_kernel<1>(gn_stats_kernel<false>, dim3(grid), dim3(threads), smem, reinterpret_cast<cudaStream_t>(stream), p);
    count_launch();
    return check_launch("gn_stats_kernel");
}

extern "C" int rg_groupnorm_apply(const rg_gn_t* g, rg_stream_t stream) {
    GnParams p; dim3 grid; int threads;
    memset(&p, 0, sizeof(p));
    int rc = fill_gn(g, p, grid, threads);
    if (rc) return rc;
    if (!g->y) return set_error(RG_ERR_ARG, "groupnorm_apply: null output");
    if (p.in_f32) launch_kernel<1>(gn_apply_kernel<true>, dim3(grid), dim3(threads), 0, reinterpret_cast<cudaStream_t>(stream), p);
    else launch_kernel<1>(gn_apply_kernel<false>, dim3(grid), dim3(threads), 0, reinterpret_cast<cudaStream_t>(stream), p);
    count_launch();
    return check_launch("gn_apply_kernel");
}

// GroupNorm in one call: the one-pass kernel where the (image, group) slice fits in shared memory, else stats + apply.
extern "C" int rg_groupnorm(const rg_gn_t* g, rg_stream_t stream) {
    if (g && g->groups > 0 && g->HW > 0 && gn_use_fused(g)) {
        GnParams p; dim3 grid; int threads;
        memset(&p, 0, sizeof(p));
        int rc = fill_gn(g, p, grid, threads);
        if (rc) return rc;
        if (!g->y) return set_error(RG_ERR_ARG, "groupnorm: null output");
        static std::atomic<bool> done_t[kMaxDevices], done_f[kMaxDevices];
        if ((rc = ensure_smem_attr(reinterpret_cast<const void*>(&gn_fused_small_kernel<true, 256>), (int)kGnFusedMaxSmem, done_t, "cudaFuncSetAttribute(gn_fused_small_kernel)"))) return rc;
        if ((rc = ensure_smem_attr(reinterpret_cast<const void*>(&gn_fused_small_kernel<false, 256>), (int)kGnFusedMaxSmem, done_f, "cudaFuncSetAttribute(gn_fused_small_kernel)"))) return rc;
        const size_t smem = (size_t)p.HW * p.cpg * (p.in_f32 ? 4 : 2);
        const dim3 fgrid((unsigned)p.groups, (unsigned)p.N);
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        if (p.in_f32) launch_kernel<1>(gn_fused_small_kernel<true, 256>, dim3(fgrid), dim3(256), smem, st, p);
        else launch_kernel<1>(gn_fused_small_kernel<false, 256>, dim3(fgrid), dim3(256), smem, st, p);
        count_launch();
        return check_launch("gn_fused_small_kernel");
    }
    int rc = rg_groupnorm_stats(g, stream);
    if (rc) return rc;
    return rg_groupnorm_apply(g, stream);
}

template <bool IN_F32, int EPL>
static void launch_ln(const void* x, int64_t rows, const float* gamma, const float* beta, float eps, void* y, cudaStream_t s) {
    const int wpb = 8;
    long long blocks = (rows + wpb - 1) / wpb;
    const long long cap = (long long)sm_count() * 8;       // persistent: each warp walks rows with a one-row prefetch
    if (blocks > cap) blocks = cap;
    launch_kernel(layernorm_kernel<IN_F32, EPL>, dim3((unsigned)blocks), dim3(wpb * 32), 0, s, x, rows, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y));
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
        if (f32) launch_kernel(layernorm_generic_kernel<true>, dim3(grid), dim3(wpb * 32), 0, s, x, rows, C, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y));
        else launch_kernel(layernorm_generic_kernel<false>, dim3(grid), dim3(wpb * 32), 0, s, x, rows, C, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y));
    }
    count_launch();
    return check_launch("layernorm_kernel");
}

extern "C" int rg_softmax_rows(void* x, int64_t rows, int32_t cols, int64_t ld, rg_stream_t stream) {
    if (!x) return set_error(RG_ERR_ARG, "softmax_rows: null pointer");
    launch_kernel(softmax_rows_kernel, dim3((unsigned)rows), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
        reinterpret_cast<__nv_bfloat16*>(x), cols, ld);
    count_launch();
    return check_launch("softmax_rows_kernel");
}
