// K1-K4: implicit-GEMM convolution / linear layer for sm_100a.
//
//   D[128 x BN] (fp32, TMEM)  +=  A[128 x 64] (bf16, smem via TMA)  x  B[BN x 64]^T (bf16, smem via TMA)
//
// * A is never materialised: for every filter tap the TMA engine fetches a {64 ch, TW, TH, TN} box of the
//   channels-last activation at the tap's spatial offset; out-of-bounds coordinates are zero-filled by the
//   hardware, which is exactly the convolution's zero padding.  Stride-2 convolutions read four "parity"
//   views of the input (one tensor map each), so they are plain shifted boxes as well.
// * B is the packed weight matrix [Cout][taps*Cin (+C2)], K-major.
// * Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..9 =
//   epilogue (TMEM -> registers -> global; two warps per TMEM lane quarter, each draining half of the tile's
//   columns).  The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of
//   tile i+1.
// * Epilogue fuses: scale, bias, per-image bias (time embedding), residual add (bf16 or fp32), SiLU,
//   GEGLU (a * gelu(g)), and writes bf16 and/or fp32, with arbitrary output pixel strides.
#include "common.cuh"
#include "internal.h"

namespace rg {

struct GemmItem { int map, dw, dh, nblk; };

struct GemmParams {
    CUtensorMap amap[5];
    CUtensorMap bmap;
    GemmItem items[10];
    int n_items;
    int total_kblk;
    int lw, lh;                       // log2 of the spatial tile (TW, TH); TN = 128 >> (lw+lh)
    int tiles_w, tiles_h, tiles_n;
    int n_tiles_m, n_tiles_n;
    int N, OH, OW, Cout;
    const float* bias;
    const float* bias_n;
    long long bias_n_ld;
    const void* res;
    int res_f32;
    __nv_bfloat16* out_bf16;
    float* out_f32;
    long long osn, osh, osw;
    int act;
    float scale;
    int vec_ok;
};

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int MAX_STAGES = (227 * 1024 - 1024) / STAGE_BYTES;
    static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
    static constexpr int ACC_STRIDE = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

constexpr int kGemmThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)

__device__ __forceinline__ float apply_act(float x, int act) { return act == 1 ? silu_f(x) : x; }

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1) conv_gemm_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES];
    __shared__ __align__(8) uint64_t empty_bar[Cfg::STAGES];
    __shared__ __align__(8) uint64_t acc_full_bar[2];
    __shared__ __align__(8) uint64_t acc_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 5; ++i) tma_prefetch_desc(&p.amap[i]);
        tma_prefetch_desc(&p.bmap);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full_bar[s], 1); mbar_init(&acc_empty_bar[s], 256); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int total_tiles = p.n_tiles_m * p.n_tiles_n;
    const int TW = 1 << p.lw, TH = 1 << p.lh;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.n_tiles_n, n_tile = tile - m_tile * p.n_tiles_n;
                const int twi = m_tile % p.tiles_w;
                const int rest = m_tile / p.tiles_w;
                const int thi = rest % p.tiles_h, tni = rest / p.tiles_h;
                const int w0 = twi * TW, h0 = thi * TH, n0 = tni * (128 >> (p.lw + p.lh));
                int kblk = 0;
                for (int it = 0; it < p.n_items; ++it) {
                    const GemmItem item = p.items[it];
                    for (int cb = 0; cb < item.nblk; ++cb, ++kblk) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sb = sa + Cfg::A_BYTES;
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_4d(sa, &p.amap[item.map], &full_bar[stage], cb * 64, w0 + item.dw, h0 + item.dh, n0);
                        tma_load_2d(sb, &p.bmap, &full_bar[stage], kblk * 64, n_tile * BN);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
                for (int kb = 0; kb < p.total_kblk; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t adesc = umma_desc_kmajor_sw128(sa);
                    const uint64_t bdesc = umma_desc_kmajor_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // advance 16 bf16 = 32 B along K inside the 128-B swizzle atom: +2 in the (addr >> 4) field
                        umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);        // frees the smem stage when these MMAs retire
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full_bar[acc]);            // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================================== epilogue (8 warps: 128 rows x 2 column halves)
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;          // which half of the tile's columns this warp drains
        const int row = q * 32 + lane;
        const int tw = row & (TW - 1);
        const int th = (row >> p.lw) & (TH - 1);
        const int tn = row >> (p.lw + p.lh);
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.n_tiles_n, n_tile = tile - m_tile * p.n_tiles_n;
            const int twi = m_tile % p.tiles_w;
            const int rest = m_tile / p.tiles_w;
            const int thi = rest % p.tiles_h, tni = rest / p.tiles_h;
            const int ow = twi * TW + tw, oh = thi * TH + th, n = tni * (128 >> (p.lw + p.lh)) + tn;
            const bool valid = (ow < p.OW) && (oh < p.OH) && (n < p.N);
            const long long off = valid ? ((long long)n * p.osn + (long long)oh * p.osh + (long long)ow * p.osw) : 0;

            mbar_wait(&acc_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_STRIDE;

            if (p.act == 2) {
                // GEGLU: tile columns [0,BN/2) = a, [BN/2,BN) = gate
                if constexpr (BN == 160) {
                    constexpr int HALF = BN / 2;
#pragma unroll 1
                    for (int j = half ? 48 : 0; j < (half ? HALF : 48); j += 16) {
                        uint32_t va[16], vg[16];
                        tmem_ld16(taddr + j, va);
                        tmem_ld16(taddr + HALF + j, vg);
                        tmem_ld_wait();
                        if (valid) {
                            const int ca = n_tile * BN + j, cg = ca + HALF, co = n_tile * HALF + j;
                            uint32_t packed[8];
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                float a0 = __uint_as_float(va[i]) * p.scale + (p.bias ? __ldg(p.bias + ca + i) : 0.f);
                                float a1 = __uint_as_float(va[i + 1]) * p.scale + (p.bias ? __ldg(p.bias + ca + i + 1) : 0.f);
                                float g0 = __uint_as_float(vg[i]) * p.scale + (p.bias ? __ldg(p.bias + cg + i) : 0.f);
                                float g1 = __uint_as_float(vg[i + 1]) * p.scale + (p.bias ? __ldg(p.bias + cg + i + 1) : 0.f);
                                packed[i / 2] = pack_bf16x2(a0 * gelu_erf_f(g0), a1 * gelu_erf_f(g1));
                            }
                            uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + off + co);
                            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                        }
                    }
                }
            } else {
                constexpr int HALFN = BN >= 32 ? BN / 2 : BN;
                constexpr int CH = (HALFN % 32 == 0) ? 32 : 16;
                const int c_end = half * HALFN + HALFN < BN ? half * HALFN + HALFN : BN;
#pragma unroll 1
                for (int c = half * HALFN; c < c_end; c += CH) {
                    uint32_t v[32];
                    if constexpr (CH == 32) {
                        tmem_ld32(taddr + c, v);
                    } else {
                        uint32_t v16[16];
                        tmem_ld16(taddr + c, v16);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = v16[i];
                    }
                    tmem_ld_wait();
                    const int col0 = n_tile * BN + c;
                    if (valid && col0 < p.Cout) {
                        const float* bn_row = p.bias_n ? p.bias_n + (long long)n * p.bias_n_ld : nullptr;
                        if (p.vec_ok && col0 + CH <= p.Cout) {
#pragma unroll
                            for (int g8 = 0; g8 < CH; g8 += 8) {
                                float x[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[g8 + i]) * p.scale;
                                const int col = col0 + g8;
                                if (p.bias) {
                                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                                    x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                                    x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
                                }
                                if (bn_row) {
                                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bn_row + col));
                                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bn_row + col + 4));
                                    x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                                    x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
                                }
                                if (p.res) {
                                    if (p.res_f32) {
                                        const float* r = reinterpret_cast<const float*>(p.res) + off + col;
                                        const float4 r0 = *reinterpret_cast<const float4*>(r);
                                        const float4 r1 = *reinterpret_cast<const float4*>(r + 4);
                                        x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
                                        x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
                                    } else {
                                        const uint4 r = *reinterpret_cast<const uint4*>(
                                            reinterpret_cast<const __nv_bfloat16*>(p.res) + off + col);
                                        float2 f;
                                        f = unpack_bf16x2(r.x); x[0] += f.x; x[1] += f.y;
                                        f = unpack_bf16x2(r.y); x[2] += f.x; x[3] += f.y;
                                        f = unpack_bf16x2(r.z); x[4] += f.x; x[5] += f.y;
                                        f = unpack_bf16x2(r.w); x[6] += f.x; x[7] += f.y;
                                    }
                                }
                                if (p.act == 1) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) x[i] = silu_f(x[i]);
                                }
                                if (p.out_f32) {
                                    float* o = p.out_f32 + off + col;
                                    *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
                                    *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
                                }
                                if (p.out_bf16) {
                                    *reinterpret_cast<uint4*>(p.out_bf16 + off + col) =
                                        make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]),
                                                   pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                                }
                            }
                        } else {
                            // ragged tail / unaligned output: scalar path (fully unrolled so v[] stays in registers)
#pragma unroll
                            for (int i = 0; i < CH; ++i) {
                                const int col = col0 + i;
                                if (col < p.Cout) {
                                    float x = __uint_as_float(v[i]) * p.scale;
                                    if (p.bias) x += __ldg(p.bias + col);
                                    if (bn_row) x += __ldg(bn_row + col);
                                    if (p.res) {
                                        x += p.res_f32 ? reinterpret_cast<const float*>(p.res)[off + col]
                                                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.res)[off + col]);
                                    }
                                    x = apply_act(x, p.act);
                                    if (p.out_f32) p.out_f32[off + col] = x;
                                    if (p.out_bf16) p.out_bf16[off + col] = __float2bfloat16(x);
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ============================================================================================ host side
static int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

static int encode_act_map(CUtensorMap* m, const void* base, int C, long long W, long long H, long long N,
                          long long sw, long long sh, long long sn, int TW, int TH, int TN) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                             CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN>
static int launch_gemm(const GemmParams& gp, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_gemm_kernel)");
        attr_done = true;
    }
    const int total = gp.n_tiles_m * gp.n_tiles_n;
    const int grid = total < sm_count() ? total : sm_count();
    conv_gemm_kernel<BN><<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(gp);
    count_launch();
    return check_launch("conv_gemm_kernel");
}

}  // namespace rg

using namespace rg;

extern "C" int rg_conv2d(const rg_conv_t* c, rg_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!c || !c->x.data || !c->w) return set_error(RG_ERR_ARG, "rg_conv2d: null pointer");
    const bool plain = c->kh == 1 && c->kw == 1 && !c->has_x2;      // single K segment: a ragged last K block is
    if (!plain && (c->x.C % 64 != 0 || (c->has_x2 && c->x2.C % 64 != 0)))   // zero-filled by TMA in both operands
        return set_error(RG_ERR_ARG, "rg_conv2d: channel counts must be multiples of 64");
    if (c->kh < 1 || c->kh > 3 || c->kw < 1 || c->kw > 3 || (c->stride != 1 && c->stride != 2))
        return set_error(RG_ERR_ARG, "rg_conv2d: unsupported kernel size / stride");
    if (c->kh * c->kw + (c->has_x2 ? 1 : 0) > 10) return set_error(RG_ERR_ARG, "rg_conv2d: too many taps");
    if (!c->out_bf16 && !c->out_f32) return set_error(RG_ERR_ARG, "rg_conv2d: no output");
    if (c->x.stride_w % 8 || c->x.stride_h % 8 || c->x.stride_n % 8 || (reinterpret_cast<uintptr_t>(c->x.data) & 15))
        return set_error(RG_ERR_ARG, "rg_conv2d: x strides must be multiples of 8 elements, base 16-B aligned");

    GemmParams gp;
    memset(&gp, 0, sizeof(gp));
    const int OW = c->OW, OH = c->OH, N = c->x.N;
    // spatial tile: TW x TH x TN = 128 output pixels
    int lw = ilog2_ceil(OW < 128 ? OW : 128);
    if (lw > 7) lw = 7;
    int lh = ilog2_ceil(OH);
    if (lh > 7 - lw) lh = 7 - lw;
    const int TW = 1 << lw, TH = 1 << lh, TN = 128 >> (lw + lh);
    gp.lw = lw; gp.lh = lh;
    gp.tiles_w = (OW + TW - 1) / TW;
    gp.tiles_h = (OH + TH - 1) / TH;
    gp.tiles_n = (N + TN - 1) / TN;
    gp.n_tiles_m = gp.tiles_w * gp.tiles_h * gp.tiles_n;
    gp.N = N; gp.OH = OH; gp.OW = OW; gp.Cout = c->Cout;

    const int cblk = (c->x.C + 63) / 64;
    int n_items = 0, n_maps = 0, rc;
    const char* xb = reinterpret_cast<const char*>(c->x.data);
    if (c->stride == 1) {
        rc = encode_act_map(&gp.amap[0], xb, c->x.C, c->x.W, c->x.H, N, c->x.stride_w, c->x.stride_h, c->x.stride_n,
                            TW, TH, TN);
        if (rc) return rc;
        n_maps = 1;
        for (int kh = 0; kh < c->kh; ++kh)
            for (int kw = 0; kw < c->kw; ++kw) gp.items[n_items++] = GemmItem{0, kw - c->pad_l, kh - c->pad_t, cblk};
    } else {
        // four parity views: view (ph,pw) holds input pixels (2i+ph, 2j+pw)
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const long long Wp = (c->x.W - pw + 1) / 2, Hp = (c->x.H - ph + 1) / 2;
                const char* base = xb + ((long long)ph * c->x.stride_h + (long long)pw * c->x.stride_w) * 2;
                if (Wp <= 0 || Hp <= 0) return set_error(RG_ERR_ARG, "rg_conv2d: stride-2 input too small");
                rc = encode_act_map(&gp.amap[ph * 2 + pw], base, c->x.C, Wp, Hp, N, 2 * c->x.stride_w,
                                    2 * c->x.stride_h, c->x.stride_n, TW, TH, TN);
                if (rc) return rc;
            }
        n_maps = 4;
        for (int kh = 0; kh < c->kh; ++kh)
            for (int kw = 0; kw < c->kw; ++kw) {
                const int uh = kh - c->pad_t, uw = kw - c->pad_l;
                const int ph = ((uh % 2) + 2) % 2, pw = ((uw % 2) + 2) % 2;
                gp.items[n_items++] = GemmItem{ph * 2 + pw, (uw - pw) / 2, (uh - ph) / 2, cblk};
            }
    }
    int ktot = c->kh * c->kw * c->x.C;
    if (c->has_x2) {
        if (!c->x2.data || (reinterpret_cast<uintptr_t>(c->x2.data) & 15) || c->x2.stride_w % 8 || c->x2.stride_h % 8 ||
            c->x2.stride_n % 8)
            return set_error(RG_ERR_ARG, "rg_conv2d: bad x2");
        rc = encode_act_map(&gp.amap[n_maps], c->x2.data, c->x2.C, c->x2.W, c->x2.H, N, c->x2.stride_w, c->x2.stride_h,
                            c->x2.stride_n, TW, TH, TN);
        if (rc) return rc;
        gp.items[n_items++] = GemmItem{n_maps, 0, 0, c->x2.C / 64};
        ++n_maps;
        ktot += c->x2.C;
    }
    for (int i = n_maps; i < 5; ++i) gp.amap[i] = gp.amap[0];
    gp.n_items = n_items;
    gp.total_kblk = (ktot + 63) / 64;

    // tile width in N
    int BN;
    if (c->act == RG_ACT_GEGLU) {
        if (c->Cout % 160 != 0 || !c->out_bf16 || c->out_f32 || c->res || c->bias_n)
            return set_error(RG_ERR_ARG, "rg_conv2d: GEGLU needs Cout % 160 == 0 and a bf16 output only");
        BN = 160;
    } else if (c->Cout <= 16) BN = 16;
    else if (c->Cout <= 32) BN = 32;
    else if (c->Cout <= 64) BN = 64;
    else if (c->Cout % 160 == 0) BN = 160;
    else BN = 128;
    gp.n_tiles_n = (c->Cout + BN - 1) / BN;

    {
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)c->Cout};
        const long long w_ld = c->w_ld ? c->w_ld : ktot;
        if (w_ld % 8 || w_ld < ktot) return set_error(RG_ERR_ARG, "rg_conv2d: w_ld must be a multiple of 8 and >= Ktot");
        cuuint64_t strides[1] = {(cuuint64_t)w_ld * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        if (reinterpret_cast<uintptr_t>(c->w) & 15) return set_error(RG_ERR_ARG, "rg_conv2d: weights must be 16-B aligned");
        rc = encode_tensor_map(&gp.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c->w, dims, strides, box, estr,
                               CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }

    gp.bias = c->bias; gp.bias_n = c->bias_n; gp.bias_n_ld = c->bias_n_ld;
    gp.res = c->res; gp.res_f32 = c->res_dtype == RG_DT_F32;
    gp.out_bf16 = reinterpret_cast<__nv_bfloat16*>(c->out_bf16);
    gp.out_f32 = c->out_f32;
    gp.osn = c->out_stride_n; gp.osh = c->out_stride_h; gp.osw = c->out_stride_w;
    gp.act = c->act; gp.scale = c->scale;
    const bool aligned = (c->out_stride_n % 8 == 0) && (c->out_stride_h % 8 == 0) && (c->out_stride_w % 8 == 0) &&
                         !(reinterpret_cast<uintptr_t>(c->out_bf16) & 15) && !(reinterpret_cast<uintptr_t>(c->out_f32) & 15) &&
                         !(reinterpret_cast<uintptr_t>(c->res) & 15) && !(reinterpret_cast<uintptr_t>(c->bias) & 15) &&
                         !(reinterpret_cast<uintptr_t>(c->bias_n) & 15) && (c->bias_n_ld % 4 == 0) && (c->Cout % 4 == 0);
    gp.vec_ok = aligned ? 1 : 0;
    if (c->act == RG_ACT_GEGLU && !aligned) return set_error(RG_ERR_ARG, "rg_conv2d: GEGLU output must be 16-B aligned");

    switch (BN) {
        case 16: return launch_gemm<16>(gp, stream);
        case 32: return launch_gemm<32>(gp, stream);
        case 64: return launch_gemm<64>(gp, stream);
        case 128: return launch_gemm<128>(gp, stream);
        default: return launch_gemm<160>(gp, stream);
    }
}
