// K5/K6: fused flash-style attention for sm_100a.  O = softmax(scale * Q K^T) V per (batch, head).
//
// One CTA owns 128 query rows of one (batch, head) and walks the keys in tiles of BKV:
//   warp 0      TMA producer: Q once, then K/V tiles through a 2-stage ring
//   warp 1      tcgen05.mma issuer:  S = Q K^T  (M=128, N=BKV, K=d)  into TMEM columns [0, BKV)
//                                    Opart = P V (M=128, N=DV,  K=BKV) into TMEM columns [128, 128+DV)
//   warps 2..5  softmax: thread r owns query row r (TMEM lane r), so row max / row sum need no shuffles;
//               P is written to shared memory as the bf16 K-major SWIZZLE_128B A operand of the second MMA;
//               the running output lives in fp32 registers and is rescaled online.
// Head dims that are not a multiple of 64 (40, 80, 160 in SD-1.5) are handled by the TMA engine: the tensor
// map's innermost extent is d, so the rest of each 64-wide box is zero-filled in shared memory.
// V is consumed directly as an MN-major B operand -- no transpose anywhere.
#include "common.cuh"
#include "internal.h"

namespace rg {

struct AttnParams {
    CUtensorMap qmap, kmap, vmap;
    __nv_bfloat16* out;
    long long osb, ost, osh;
    int Nq, Nk, d;
    float scale_log2;     // scale * log2(e)
};

constexpr int TMEM_COLS_FOR(int dv) { return dv <= 128 ? 256 : 512; }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float y;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
    return y;
}

template <int DKA, int DV, int BKV>
struct AttnCfg {
    static constexpr int Q_BYTES = DKA * 128 * 128;
    static constexpr int KV_ATOM_BYTES = BKV * 128;
    static constexpr int K_BYTES = DKA * KV_ATOM_BYTES;
    static constexpr int V_BYTES = DKA * KV_ATOM_BYTES;
    static constexpr int P_BYTES = (BKV / 64) * 128 * 128;
    static constexpr int STAGES = 2;
    static constexpr int TILE_BYTES = Q_BYTES + STAGES * (K_BYTES + V_BYTES) + P_BYTES;
    // barriers live behind the tiles; there is no static shared memory, so the dynamic window starts 1024-aligned
    // (checked at run time) and no alignment slack is needed: d=40 uses 112.1 KB and two CTAs share an SM
    static constexpr int SMEM_BYTES = TILE_BYTES + 128;
    static constexpr int MIN_CTAS = (2 * (SMEM_BYTES + 1024) <= 228 * 1024 && TMEM_COLS_FOR(DV) <= 256) ? 2 : 1;
    static constexpr int TMEM_COLS = TMEM_COLS_FOR(DV);
    static constexpr int O_COL = 128;
};

constexpr int kAttnThreads = 192;

template <int DKA, int DV, int BKV>
__global__ void __launch_bounds__(kAttnThreads, AttnCfg<DKA, DV, BKV>::MIN_CTAS)
attention_kernel(const __grid_constant__ AttnParams p) {
    using Cfg = AttnCfg<DKA, DV, BKV>;
    extern __shared__ __align__(16) uint8_t smem[];                    // window-relative base 0: no static smem
    if ((smem_u32(smem) & 1023u) != 0) __trap();                       // SWIZZLE_128B tiles need 1024-B alignment
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + Cfg::Q_BYTES;                                  // [stage][K | V]
    uint8_t* sP = sKV + Cfg::STAGES * (Cfg::K_BYTES + Cfg::V_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::TILE_BYTES);
    uint64_t& q_full = bars[0]; uint64_t& s_full = bars[1]; uint64_t& p_full = bars[2]; uint64_t& o_full = bars[3];
    uint64_t* kv_full = bars + 4; uint64_t* kv_empty = bars + 6;
    uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_kv = (p.Nk + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap); }
    if (warp == 1 && lane == 0) {
        mbar_init(&q_full, 1); mbar_init(&s_full, 1); mbar_init(&p_full, 128); mbar_init(&o_full, 1);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&q_full, Cfg::Q_BYTES);
#pragma unroll
            for (int a = 0; a < DKA; ++a) tma_load_4d(sQ + a * 128 * 128, &p.qmap, &q_full, a * 64, h, q0, b);
            int stage = 0; uint32_t phase = 0;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                uint8_t* sk = sKV + stage * (Cfg::K_BYTES + Cfg::V_BYTES);
                uint8_t* sv = sk + Cfg::K_BYTES;
                mbar_expect_tx(&kv_full[stage], Cfg::K_BYTES + Cfg::V_BYTES);
#pragma unroll
                for (int a = 0; a < DKA; ++a) {
                    tma_load_4d(sk + a * Cfg::KV_ATOM_BYTES, &p.kmap, &kv_full[stage], a * 64, h, j * BKV, b);
                    tma_load_4d(sv + a * Cfg::KV_ATOM_BYTES, &p.vmap, &kv_full[stage], a * 64, h, j * BKV, b);
                }
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, BKV, 0, 0);    // Q (K-major) x K (K-major)
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV, 0, 1);     // P (K-major) x V (MN-major)
            const uint32_t s_tmem = tmem_base, o_tmem = tmem_base + Cfg::O_COL;
            mbar_wait(&q_full, 0);
            int stage = 0; uint32_t phase = 0;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&kv_full[stage], phase);
                tc_fence_after();
                const uint32_t sk = smem_u32(sKV + stage * (Cfg::K_BYTES + Cfg::V_BYTES));
                const uint32_t sv = sk + Cfg::K_BYTES;
                // ---- S = Q K^T over d (DV/16 k-steps; columns >= d are zero in both operands)
#pragma unroll
                for (int ks = 0; ks < DV / 16; ++ks) {
                    const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sQ) + (ks / 4) * 128 * 128 + (ks % 4) * 32);
                    const uint64_t bdesc = umma_desc_kmajor_sw128(sk + (ks / 4) * Cfg::KV_ATOM_BYTES + (ks % 4) * 32);
                    umma_bf16(s_tmem, adesc, bdesc, idesc_s, ks != 0 ? 1u : 0u);
                }
                umma_commit(&s_full);
                // ---- Opart = P V once the softmax warps have published P
                mbar_wait(&p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sP) + (ks / 4) * 128 * 128 + (ks % 4) * 32);
                    // 16 kv rows per k-step = 2 groups of 8 rows (1024 B each); 64-wide d blocks are KV_ATOM_BYTES apart
                    const uint64_t bdesc = umma_desc_mnmajor_sw128(sv + ks * 2048, Cfg::KV_ATOM_BYTES, 1024);
                    umma_bf16(o_tmem, adesc, bdesc, idesc_o, ks != 0 ? 1u : 0u);
                }
                umma_commit(&o_full);
                umma_commit(&kv_empty[stage]);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const uint32_t s_tmem = tmem_base + lane_addr, o_tmem = tmem_base + lane_addr + Cfg::O_COL;
        float o_acc[DV];
#pragma unroll
        for (int i = 0; i < DV; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        uint8_t* p_row = sP + row * 128;
        const int sw = row & 7;

        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full, j & 1);
            tc_fence_after();
            const int kv_left = p.Nk - j * BKV;              // columns >= kv_left are padding
            const bool full = kv_left >= BKV;                // warp-uniform: only the last tile can be ragged
            // ---- pass 1: row max of the raw logits
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < BKV; c += 32) {
                uint32_t v[32];
                tmem_ld32(s_tmem + c, v);
                tmem_ld_wait();
                if (full) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c + i < kv_left) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = ex2_approx(m_run - m_new);
            // ---- pass 2: p = 2^(s*scale*log2e - m), row sum, bf16 P -> swizzled smem
            float l_tile = 0.f;
#pragma unroll 1
            for (int c = 0; c < BKV; c += 32) {
                uint32_t v[32];
                tmem_ld32(s_tmem + c, v);
                tmem_ld_wait();
                float pv[32];
                if (full) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        pv[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
                        l_tile += pv[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float e = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
                        pv[i] = (c + i < kv_left) ? e : 0.f;
                        l_tile += pv[i];
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int kc = (c >> 3) + g;             // 16-byte chunk index along kv
                    uint8_t* dst = p_row + (kc >> 3) * (128 * 128) + (((kc & 7) ^ sw) << 4);
                    *reinterpret_cast<uint4*>(dst) =
                        make_uint4(pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]), pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]),
                                   pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]), pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]));
                }
            }
            l_run = l_run * alpha + l_tile;
            m_run = m_new;
            fence_proxy_async_smem();        // generic-proxy writes of P -> visible to the tensor core (async proxy)
            tc_fence_before();
            mbar_arrive(&p_full);
            // ---- fold Opart into the running output
            mbar_wait(&o_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < DV; c += 16) {
                uint32_t v[16];
                tmem_ld16(o_tmem + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o_acc[c + i] = o_acc[c + i] * alpha + __uint_as_float(v[i]);
            }
            tc_fence_before();
        }
        const int t = q0 + row;
        if (t < p.Nq) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = p.out + (long long)b * p.osb + (long long)t * p.ost + (long long)h * p.osh;
#pragma unroll
            for (int c = 0; c < DV; c += 8) {
                if (c < p.d) {
                    *reinterpret_cast<uint4*>(dst + c) =
                        make_uint4(pack_bf16x2(o_acc[c] * inv, o_acc[c + 1] * inv), pack_bf16x2(o_acc[c + 2] * inv, o_acc[c + 3] * inv),
                                   pack_bf16x2(o_acc[c + 4] * inv, o_acc[c + 5] * inv), pack_bf16x2(o_acc[c + 6] * inv, o_acc[c + 7] * inv));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

static int encode_qkv_map(CUtensorMap* m, const void* base, int d, int heads, long long tokens, int B, long long sh,
                          long long st, long long sb, int rows) {
    cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)st * 2, (cuuint64_t)sb * 2};
    cuuint32_t box[4] = {64, 1, (cuuint32_t)rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                             CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int DKA, int DV, int BKV>
static int launch_attn(const rg_attn_t* a, cudaStream_t stream) {
    using Cfg = AttnCfg<DKA, DV, BKV>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel<DKA, DV, BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attention_kernel)");
        attr_done = true;
    }
    AttnParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    if ((rc = encode_qkv_map(&p.qmap, a->q, a->d, a->heads, a->Nq, a->B, a->q_stride_h, a->q_stride_t, a->q_stride_b, 128))) return rc;
    if ((rc = encode_qkv_map(&p.kmap, a->k, a->d, a->heads, a->Nk, a->B, a->k_stride_h, a->k_stride_t, a->k_stride_b, BKV))) return rc;
    if ((rc = encode_qkv_map(&p.vmap, a->v, a->d, a->heads, a->Nk, a->B, a->v_stride_h, a->v_stride_t, a->v_stride_b, BKV))) return rc;
    p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
    p.osb = a->o_stride_b; p.ost = a->o_stride_t; p.osh = a->o_stride_h;
    p.Nq = a->Nq; p.Nk = a->Nk; p.d = a->d;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    dim3 grid((a->Nq + 127) / 128, a->heads, a->B);
    attention_kernel<DKA, DV, BKV><<<grid, kAttnThreads, Cfg::SMEM_BYTES, stream>>>(p);
    count_launch();
    return check_launch("attention_kernel");
}

}  // namespace rg

using namespace rg;

extern "C" int rg_attention(const rg_attn_t* a, rg_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->q || !a->k || !a->v || !a->out) return set_error(RG_ERR_ARG, "attention: null pointer");
    if (a->d % 8 || a->d < 8 || a->d > 160) return set_error(RG_ERR_ARG, "attention: head dim must be a multiple of 8 in [8,160]");
    if (a->Nq < 1 || a->Nk < 1) return set_error(RG_ERR_ARG, "attention: empty sequence");
    const int64_t st[] = {a->q_stride_b, a->q_stride_t, a->q_stride_h, a->k_stride_b, a->k_stride_t, a->k_stride_h,
                          a->v_stride_b, a->v_stride_t, a->v_stride_h, a->o_stride_b, a->o_stride_t, a->o_stride_h};
    for (int64_t s : st)
        if (s % 8) return set_error(RG_ERR_ARG, "attention: strides must be multiples of 8 elements");
    if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
         reinterpret_cast<uintptr_t>(a->out)) & 15)
        return set_error(RG_ERR_ARG, "attention: pointers must be 16-byte aligned");
    const int dv = (a->d + 15) / 16 * 16;
    if (dv <= 48) return launch_attn<1, 48, 128>(a, stream);
    if (dv <= 64) return launch_attn<1, 64, 128>(a, stream);
    if (dv <= 80) return launch_attn<2, 80, 128>(a, stream);
    if (dv <= 128) return launch_attn<2, 128, 64>(a, stream);
    return launch_attn<3, 160, 64>(a, stream);
}
