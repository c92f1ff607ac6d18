// K5/K6: fused flash-style attention for sm_100a.  O = softmax(scale * Q K^T) V per (batch, head).
//
// Persistent kernel, one CTA per SM.  A work item is 256 query rows (two 128-row tiles) of one (batch, head); the
// CTA walks the keys in tiles of BKV for both query tiles at once, so the two softmax warpgroups ping-pong on
// the tensor core and every K/V tile is fetched once per 256 queries.
//
//   warps 0..3   softmax warpgroup of query tile 0   (thread r owns query row r = TMEM lane r: no shuffles)
//   warps 4..7   softmax warpgroup of query tile 1
//   warp  8      TMA producer: Q tiles once per item, K/V tiles through a STAGES-deep ring
//   warp  9, 11  tcgen05.mma issuers, ONE THREAD each, one per query tile t (a single thread serves both tiles for the short launches,
//                AttnParams::n_issuers):
//                                     S_t = Q_t K^T  (M=128, N=BKV, K=d)   -> TMEM columns of S_t
//                                     O_t += P_t V   (M=128, N=DV,  K=BKV) -> TMEM columns of O_t
//   warp  10     fp16 only: writes the ones column into every freshly landed V tile (see below) and only then
//                publishes the stage to the issuers
//
// Why two issuers and a patch warp (round 2): with one issuing warp the period of a 256-query x 128-key step was bound
// by that warp alone -- tcgen05.mma issue blocks while the pipe's short queue is full (1218 cycles per step), and in
// between the same thread did five barrier round trips, commits, warp syncs and the 400-cycle ones-column patch, during
// which the tensor pipe idled (3285 cycles per step with the exponentials compiled out).  Both
// warpgroups then waited for the same thread at the same point, ran their exponentials in lock-step and left the MUFU
// pipe idle during their bookkeeping.  Now each warpgroup has its own issuing thread that never syncs a warp, and the
// patch runs ahead of both in a warp of its own, inside the K/V ring's slack.
//
// S is double-buffered per query tile (SBUF = 2): S_t(j+1) is computed while the warpgroup is still busy with
// S_t(j), so the softmax threads never wait for the tensor core and the MUFU (exp) pipe is the only limiter.
// Everything between the two GEMMs stays in tensor memory: the softmax threads read S with tcgen05.ld, write the
// bf16 probabilities P back over the same columns with tcgen05.st, and the second GEMM takes P as its A operand
// straight from TMEM (no shared-memory round trip, no swizzled stores).  O accumulates in TMEM across the whole
// key loop; the running maximum is only raised when a tile exceeds it by more than 2^8 (lazy rescale), in which
// case the owning warp waits for the previous P V to retire and rescales its O rows in TMEM before publishing P.  exp is ex2.approx on pre-scaled logits.
// fp16 operands (F16; what the UNet uses, like the reference's CUDA dtype): ex2.approx.f16x2 on the packed, pre-scaled
// logits writes P directly as fp16 pairs (same MUFU rate as fp32 -- 16 ex2 / clk / SM measured either way -- but no
// separate pack or row-sum instructions), and the softmax denominator is not summed by the threads at all -- the MMA
// warp writes a column of ones into the zero padding of every V tile (column d of DV), so O_t[:, d] = sum_k P and it is
// rescaled together with the rest of O_t.  bf16 operands keep the one-exponential-per-instruction fp32 path.
// Head dims that are not a multiple of 64 (40, 80, 160 in SD-1.5) are handled by the TMA engine: the tensor
// map's innermost extent is d, so the rest of each 64-wide box is zero-filled in shared memory.
// V is consumed directly as an MN-major B operand -- no transpose anywhere.
#include <stdlib.h>
#include "common.cuh"
#include "internal.h"

namespace rg {

struct AttnParams {
    CUtensorMap qmap, kmap, vmap;
    __nv_bfloat16* out;
    long long osb, ost, osh;
    int Nq, Nk, d;
    int heads, n_qpairs, n_items;
    float scale_log2;     // scale * log2(e)
    long long* trace;     // debug only (rg_debug_attn_trace): per-tile clock64 stamps of CTA 0, else nullptr
    int turns;            // the two softmax warps of every SM sub-partition take turns for their exponential phases
    int trace_item;       // which of CTA 0's work items is traced (rg_debug_attn_trace_item; default the first)
    int n_issuers;        // 1: warp 9 issues for both query tiles; 2: warp 9 -> tile 0, warp 11 -> tile 1
    int skew;             // cycles by which tile 1 starts behind tile 0 at a CTA's first work item (2 issuers only)
};

constexpr int kTraceTiles = 64, kTraceStamps = 8;
#define RG_STAMP(k) do { if (tr && j < kTraceTiles) tr[((warp * kTraceTiles) + j) * kTraceStamps + (k)] = clock64(); } while (0)

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

template <int DKA, int DQK, int DV, int BKV, int SBUF, bool PSEP, int STAGES, bool F16>
struct AttnCfg {
    static constexpr int Q_TILE_BYTES = DKA * 128 * 128;           // one 128-query tile: DKA atoms of [128][64] bf16
    static constexpr int KV_ATOM_BYTES = BKV * 128;
    static constexpr int K_BYTES = DKA * KV_ATOM_BYTES;
    static constexpr int V_BYTES = DKA * KV_ATOM_BYTES;
    static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;
    static constexpr int TILE_BYTES = 2 * Q_TILE_BYTES + STAGES * STAGE_BYTES;
    // barriers live behind the tiles; there is no static shared memory, so the dynamic window starts 1024-aligned
    static constexpr int SMEM_BYTES = TILE_BYTES + 512;
    static constexpr int O_STRIDE = (DV + 31) / 32 * 32;
    // TMEM columns: S_t[buf] at (t*SBUF+buf)*BKV.  PSEP: P_t has its own BKV/2 columns (two bf16 per column), so
    // S_t(G+SBUF) can be issued as soon as S_t(G) is in registers and runs under the softmax of tile G; otherwise P_t
    // is written over S_t and the next score GEMM is issued in order behind O_t += P_t V.
    static constexpr int P_COL = 2 * SBUF * BKV;
    static constexpr int P_STRIDE = PSEP ? BKV / 2 : 0;
    static constexpr int O_COL = P_COL + 2 * P_STRIDE;
    static constexpr int TMEM_COLS = 512;
    static constexpr int NBAR = 2 + 4 * SBUF + 4 + 3 * STAGES + 8;
    static_assert(O_COL + 2 * O_STRIDE <= 512, "TMEM budget");
    static_assert(NBAR * 8 + 8 + 16 <= 512, "barrier area");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(BKV == 64 || BKV == 128, "BKV");
    static_assert(SBUF == 1 || SBUF == 2, "SBUF");
    static_assert(PSEP || SBUF == 1, "aliased P implies a single score buffer");
    static_assert(!PSEP || STAGES >= SBUF + 1, "SBUF tiles of score look-ahead need SBUF+1 K/V stages");
};

#ifndef RG_ATTN_SETMAXNREG
#define RG_ATTN_SETMAXNREG 1
#endif
#ifndef RG_ATTN_EXP_F32
#define RG_ATTN_EXP_F32 0
#endif
// Register budget by role (setmaxnreg, per warpgroup): the helper warpgroup (warps 8-11: one or zero active threads each)
// gives registers back so that the softmax warpgroups keep S (128 fp32), P (64 packed pairs) and the exponentials in
// flight without spills: 2 x 128 x kSoftmaxRegs + 128 x kHelperRegs <= 64 K.  The instruction must dominate the role's
// code (ptxas budgets registers per region), hence the two-level role dispatch in the kernel.
constexpr int kSoftmaxRegs = 232, kHelperRegs = 40;
static_assert(256 * kSoftmaxRegs + 128 * kHelperRegs <= 65536, "register file");
constexpr int kAttnThreads = 384;
// Stagger (cycles) of query tile 1 behind tile 0 for the long key loops: the two warps that share an SM sub-partition
// then run their exponential phase (MUFU-bound) and their bookkeeping (barrier round trips, tcgen05.ld / st, row max:
// issue- and latency-bound, MUFU idle) in anti-phase instead of in lock-step.  About half a step period.
constexpr int kAttnSkewCycles = 700;      // flat optimum 300..1100 cycles, lock-step again from ~1500 (profiles/r02_attn_variants.txt)
#ifndef RG_ATTN_PSTREAM
#define RG_ATTN_PSTREAM 1
#endif
// P_t in its own TMEM columns (PSEP): wait for the previous P V before the exponentials and stream every finished 32-key
// chunk of P out during them, instead of waiting and storing all of P after the last exponential
constexpr bool kAttnPStream = RG_ATTN_PSTREAM != 0;
constexpr float kLazyRescale = 8.0f;       // raise the running max only when a tile exceeds it by > 2^8

// All barrier phases are indexed by the CTA-global key-tile counter G = (items done) * n_kv + j, which is also the
// K/V ring position; score buffer G % SBUF is used for the (G / SBUF)-th time.
template <int DKA, int DQK, int DV, int BKV, int SBUF, bool PSEP, int STAGES, bool F16, bool CAUSAL = false>
__global__ void __launch_bounds__(kAttnThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
    using Cfg = AttnCfg<DKA, DQK, DV, BKV, SBUF, PSEP, STAGES, F16>;
    extern __shared__ __align__(16) uint8_t smem[];                    // window-relative base 0: no static smem
    if ((smem_u32(smem) & 1023u) != 0) __trap();                       // SWIZZLE_128B tiles need 1024-B alignment
    uint8_t* sQ = smem;                                                // [tile][atom][128][64]
    uint8_t* sKV = sQ + 2 * Cfg::Q_TILE_BYTES;                         // [stage][K | V]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::TILE_BYTES);
    uint64_t& q_full = bars[0]; uint64_t& q_empty = bars[1];
    uint64_t* s_full = bars + 2;                       // [t][buf]  S_t(G) accumulated              (tensor core -> softmax)
    uint64_t* s_free = s_full + 2 * SBUF;              // [t][buf]  S_t(G) copied to registers      (softmax -> tensor core)
    uint64_t* p_full = s_free + 2 * SBUF;              // [t]       P_t(G) stored, O_t rescaled     (softmax -> tensor core)
    uint64_t* pv_done = p_full + 2;                    // [t]       O_t += P_t(G) V retired         (tensor core -> softmax)
    uint64_t* kv_full = pv_done + 2; uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* kv_land = kv_empty + STAGES;             // [stage]   fp16: TMA bytes landed (-> patch warp -> kv_full)
    uint64_t* exp_done = kv_land + STAGES;             // [sub-partition q][t] warp q of warpgroup t has issued its exponentials of tile G
    uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(exp_done + 8);
    volatile long long* t0_stamp = reinterpret_cast<volatile long long*>(exp_done + 8 + 1);   // [item parity] clock64 of warpgroup 0's start

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kv = (p.Nk + BKV - 1) / BKV;

    pdl_trigger();
    if (warp == 8 && lane == 0) { tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap); }
    if (warp == 9 && lane == 0) {
        mbar_init(&q_full, 1); mbar_init(&q_empty, p.n_issuers);
        for (int i = 0; i < 2 * SBUF; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128); }
        for (int t = 0; t < 2; ++t) { mbar_init(&p_full[t], 128); mbar_init(&pv_done[t], 1); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], p.n_issuers); mbar_init(&kv_land[s], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&exp_done[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();                         // prologue above overlapped the previous kernel

    if (warp >= 8) {
#if RG_ATTN_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kHelperRegs));
#endif
    if (warp == 8) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            uint32_t g = 0, it = 0;                                     // g: K/V tiles loaded so far (ring position)
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
                const int qp = item % p.n_qpairs, bh = item / p.n_qpairs;
                const int h = bh % p.heads, b = bh / p.heads;
                const int q0 = qp * 256;
                if (it > 0) mbar_wait(&q_empty, (it - 1) & 1);          // previous item's last S GEMM has read Q
                mbar_expect_tx(&q_full, 2 * Cfg::Q_TILE_BYTES);
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int a = 0; a < DKA; ++a)
                        tma_load_4d(sQ + t * Cfg::Q_TILE_BYTES + a * 128 * 128, &p.qmap, &q_full, a * 64, h,
                                    q0 + t * 128, b);
                for (int j = 0; j < n_kv; ++j, ++g) {
                    const uint32_t stage = g % STAGES, phase = (g / STAGES) & 1;
                    mbar_wait(&kv_empty[stage], phase ^ 1);
                    uint8_t* sk = sKV + stage * Cfg::STAGE_BYTES;
                    uint8_t* sv = sk + Cfg::K_BYTES;
                    uint64_t* const land = F16 ? &kv_land[stage] : &kv_full[stage];     // fp16: the patch warp publishes kv_full
                    mbar_expect_tx(land, Cfg::STAGE_BYTES);
#pragma unroll
                    for (int a = 0; a < DKA; ++a) {
                        tma_load_4d(sk + a * Cfg::KV_ATOM_BYTES, &p.kmap, land, a * 64, h, j * BKV, b);
                        tma_load_4d(sv + a * Cfg::KV_ATOM_BYTES, &p.vmap, land, a * 64, h, j * BKV, b);
                    }
                }
            }
        }
    } else if (warp == 10) {
        // ===================================================================== fp16: ones-column patch warp
        // V[k][d] = 1 for every key row k, so that O_t[:, d] = sum_k P (the softmax denominator) comes out of the P V GEMM.
        // Column d sits in the TMA zero padding (atom d / 64, 16-byte chunk (d % 64) / 8 of the 128-byte row, SWIZZLE_128B:
        // chunk ^= row & 7).  TMA rewrites the padding with zeros on every load, so every landed tile is patched, then
        // handed to the issuers: generic-proxy stores -> fence.proxy.async -> mbarrier arrive -> their tcgen05.mma reads.
        if constexpr (F16) {
            uint32_t g = 0;
            const int chunk = (p.d & 63) >> 3;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                for (int j = 0; j < n_kv; ++j, ++g) {
                    const uint32_t stage = g % STAGES;
                    mbar_wait(&kv_land[stage], (g / STAGES) & 1);
                    uint8_t* sv = sKV + stage * Cfg::STAGE_BYTES + Cfg::K_BYTES + (p.d >> 6) * Cfg::KV_ATOM_BYTES;
                    for (int k = lane; k < BKV; k += 32)
                        *reinterpret_cast<uint16_t*>(sv + k * 128 + ((chunk ^ (k & 7)) << 4)) = 0x3C00;   // fp16 1.0
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&kv_full[stage]);
                }
            }
        }
    } else if (warp == 9 || warp == 11) {
        // ===================================================================== MMA issuers: one thread per query tile
        // (warp 9 -> t = 0, warp 11 -> t = 1; with p.n_issuers == 1 warp 9 serves both tiles in turn and warp 11 idles).
        // The two threads never talk to each other: each waits for the barriers of its own tile, writes its own TMEM
        // columns, and the barriers they share (q_empty, kv_empty) simply expect one commit from each.
        const bool two = p.n_issuers == 2;
        const int t_lo = two ? (warp == 9 ? 0 : 1) : 0;
        const int t_hi = two ? t_lo + 1 : 2;
        if (lane == 0 && (two || warp == 9)) {
            // kind::f16 operand format: bits [7,10) A, [10,13) B: 0 = fp16, 1 = bf16
            constexpr uint32_t fmt_clear = F16 ? ~((7u << 7) | (7u << 10)) : ~0u;
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, BKV, 0, 0) & fmt_clear;    // Q (K-major smem) x K (K-major smem)
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV, 0, 1) & fmt_clear;     // P (TMEM)         x V (MN-major smem)
            const uint32_t sq = smem_u32(sQ);
            const uint32_t skv = smem_u32(sKV);
            auto wait_kv = [&](uint32_t G) {
                mbar_wait(&kv_full[G % STAGES], (G / STAGES) & 1);
                tc_fence_after();
            };
            // S_t(G) = Q_t K_G^T
            auto issue_s = [&](int t, uint32_t G) {
                const int buf = (int)(G % SBUF);
                if (PSEP && G >= SBUF) {                         // the warpgroup has copied S_t(G-SBUF) out of this buffer
                    mbar_wait(&s_free[t * SBUF + buf], (G / SBUF - 1) & 1);
                    tc_fence_after();
                }
                const uint32_t sk = skv + (G % STAGES) * Cfg::STAGE_BYTES;
#pragma unroll
                for (int ks = 0; ks < DQK / 16; ++ks) {          // columns >= d are zero in both operands
                    const uint64_t adesc = umma_desc_kmajor_sw128(sq + t * Cfg::Q_TILE_BYTES + (ks / 4) * 128 * 128 + (ks % 4) * 32);
                    const uint64_t bdesc = umma_desc_kmajor_sw128(sk + (ks / 4) * Cfg::KV_ATOM_BYTES + (ks % 4) * 32);
                    umma_bf16(tmem_base + (t * SBUF + buf) * BKV, adesc, bdesc, idesc_s, ks != 0 ? 1u : 0u);
                }
                umma_commit(&s_full[t * SBUF + buf]);
            };
            // O_t (+)= P_t(G) V_G
            auto issue_pv = [&](int t, uint32_t G, uint32_t acc) {
                mbar_wait(&p_full[t], G & 1);
                tc_fence_after();
                const uint32_t sv = skv + (G % STAGES) * Cfg::STAGE_BYTES + Cfg::K_BYTES;
                const uint32_t p_tmem = PSEP ? tmem_base + Cfg::P_COL + t * Cfg::P_STRIDE : tmem_base + t * BKV;
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    // 16 kv rows per k-step = 2 groups of 8 rows (1024 B each); 64-wide d blocks are KV_ATOM_BYTES apart
                    const uint64_t bdesc = umma_desc_mnmajor_sw128(sv + ks * 2048, Cfg::KV_ATOM_BYTES, 1024);
                    // P: two 16-bit values per 32-bit TMEM column -> 16 keys = 8 columns
                    umma_bf16_ts(tmem_base + Cfg::O_COL + t * Cfg::O_STRIDE, p_tmem + ks * 8, bdesc, idesc_o,
                                 (acc | (uint32_t)ks) != 0 ? 1u : 0u);
                }
                umma_commit(&pv_done[t]);
            };
            // tile 1's issuer holds the first score GEMM of every work item back until p.skew cycles after warpgroup 0 has
            // picked up ITS first score tile (warp 0 stamps clock64 before its s_free arrival; the barrier is only polled
            // here, warpgroup 0's own issuer consumes it): warpgroup 1 then runs that far behind warpgroup 0 through the
            // item's key loop, whatever happened at the item boundary
            // ... and tile 0's issuer starts an item only when warpgroup 1 has published the last P of the previous one
            // (p_full[1] is only polled here; tile 1's issuer consumes it).  Without this the offset is free to grow item
            // by item until warpgroup 1 runs exactly one step behind -- lock-step again, measured with the fp32
            // exponentials: skew 1200 cycles in a CTA's first item, 3300 = one full step in its seventh.
            auto rephase = [&](uint32_t g0_, uint32_t it_) {
                if (PSEP && two && t_lo == 0 && p.skew > 0 && it_ > 0) mbar_wait(&p_full[1], (g0_ - 1) & 1);
            };
            auto stagger = [&](uint32_t g0_, uint32_t it_) {
                if (PSEP && two && t_lo == 1 && p.skew > 0) {          // (s_free / t0_stamp exist for separate P columns only)
                    mbar_wait(&s_free[g0_ % SBUF], (g0_ / SBUF) & 1);      // warpgroup 0 has started the item's first tile ...
                    const long long target = t0_stamp[it_ & 1] + (long long)p.skew;   // ... at this time (same SM, same counter)
                    while (clock64() < target) { }
                }
            };
            uint32_t g0 = 0, it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it, g0 += n_kv) {
                mbar_wait(&q_full, it & 1);
                tc_fence_after();
                rephase(g0, it);
                if constexpr (!PSEP) {
                    wait_kv(g0);
                    stagger(g0, it);
                    for (int t = t_lo; t < t_hi; ++t) issue_s(t, g0);
                    if (n_kv == 1) umma_commit(&q_empty);
                    for (int j = 0; j < n_kv; ++j) {
                        const uint32_t G = g0 + j;
                        for (int t = t_lo; t < t_hi; ++t) {
                            issue_pv(t, G, j > 0 ? 1u : 0u);
                            if (t == t_hi - 1) umma_commit(&kv_empty[G % STAGES]);    // K_G / V_G consumed once these retire
                            if (j + 1 < n_kv) {                                        // in order behind P_t(G) V: P aliases S_t
                                if (t == t_lo) wait_kv(G + 1);
                                issue_s(t, G + 1);
                                if (t == t_hi - 1 && j + 2 == n_kv) umma_commit(&q_empty);
                            }
                        }
                    }
                } else {
                    for (int j = 0; j < SBUF && j < n_kv; ++j) {              // SBUF score tiles of look-ahead
                        wait_kv(g0 + j);
                        if (j == 0) stagger(g0, it);
                        for (int t = t_lo; t < t_hi; ++t) issue_s(t, g0 + j);
                        if (j + 1 == n_kv) umma_commit(&q_empty);
                    }
                    for (int j = 0; j < n_kv; ++j) {
                        const uint32_t G = g0 + j;
                        if (j + SBUF < n_kv) {
                            wait_kv(G + SBUF);
                            for (int t = t_lo; t < t_hi; ++t) issue_s(t, G + SBUF);
                            if (j + SBUF + 1 == n_kv) umma_commit(&q_empty);
                        }
                        for (int t = t_lo; t < t_hi; ++t) issue_pv(t, G, j > 0 ? 1u : 0u);
                        umma_commit(&kv_empty[G % STAGES]);
                    }
                }
            }
        }
    }
    } else {
        // ===================================================================== softmax warpgroups
#if RG_ATTN_SETMAXNREG
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSoftmaxRegs));
#endif
        const int t = warp >> 2, qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const uint32_t s_tmem0 = tmem_base + lane_addr + t * SBUF * BKV;
        const uint32_t p_tmem0 = tmem_base + lane_addr + Cfg::P_COL + t * Cfg::P_STRIDE;
        const uint32_t o_tmem = tmem_base + lane_addr + Cfg::O_COL + t * Cfg::O_STRIDE;
        const float sl = p.scale_log2;
        uint32_t g0 = 0, it_s = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, g0 += n_kv, ++it_s) {
            const int qp = item % p.n_qpairs, bh = item / p.n_qpairs;
            const int h = bh % p.heads, b = bh / p.heads;
            float m_run = -INFINITY, l_run = 0.f;
            long long* tr = (p.trace && blockIdx.x == 0 && lane == 0 && (int)it_s == p.trace_item) ? p.trace : nullptr;
            for (int j = 0; j < n_kv; ++j) {
                const uint32_t G = g0 + j;
                RG_STAMP(0);
                const int buf = (int)(G % SBUF);
                const uint32_t s_tmem = s_tmem0 + buf * BKV;
                const uint32_t p_tmem = PSEP ? p_tmem0 : s_tmem;
                mbar_wait(&s_full[t * SBUF + buf], (G / SBUF) & 1);
                tc_fence_after();
                RG_STAMP(1);
                float s[BKV];
                {
                    uint32_t (&su)[BKV] = reinterpret_cast<uint32_t (&)[BKV]>(s);
#pragma unroll
                    for (int c = 0; c < BKV; c += 32) tmem_ld32(s_tmem + c, reinterpret_cast<uint32_t (&)[32]>(su[c]));
                    tmem_ld_wait();
                }
                RG_STAMP(2);
                if constexpr (PSEP) {                            // the buffer can take S_t(G+SBUF) now
                    if (j == 0 && warp == 0 && lane == 0) t0_stamp[it_s & 1] = clock64();      // see stagger() in the issuer
                    tc_fence_before();
                    mbar_arrive(&s_free[t * SBUF + buf]);
                }
                const int kv_left = p.Nk - j * BKV;              // columns >= kv_left are padding
                if (kv_left < BKV) {                             // warp-uniform: only the last tile can be ragged
#pragma unroll
                    for (int i = 0; i < BKV; ++i)
                        if (i >= kv_left) s[i] = -INFINITY;
                }
                if constexpr (CAUSAL) {                          // key index <= query index (CLIP text encoder)
                    const int lim = qp * 256 + t * 128 + row - j * BKV;
#pragma unroll
                    for (int i = 0; i < BKV; ++i)
                        if (i > lim) s[i] = -INFINITY;
                }
                float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
                for (int i = 4; i < BKV; i += 8) {
                    mx0 = max3(mx0, s[i], s[i + 1]);
                    mx1 = max3(mx1, s[i + 2], s[i + 3]);
                    if (i + 4 < BKV) {
                        mx2 = max3(mx2, s[i + 4], s[i + 5]);
                        mx3 = max3(mx3, s[i + 6], s[i + 7]);
                    }
                }
                const float cand = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl;
                // O_t and the P_t buffer must be quiescent before they are touched.  Aliased P: S_t(G) was issued
                // behind O_t += P_t(G-1) V, so its arrival already implies that.  PSEP: wait for phase G-1 of pv_done
                // (phase G-2 was waited for at the previous tile, so the parity cannot alias) -- as late as possible:
                // right before the P stores, or before a (rare) rescale of O_t.
                bool quiescent = !PSEP || G == 0;
                if (j == 0) {
                    m_run = cand;
                } else {
                    const bool need = cand > m_run + kLazyRescale;
                    if (__any_sync(0xffffffffu, need)) {
                        if (!quiescent) { mbar_wait(&pv_done[t], (G - 1) & 1); tc_fence_after(); quiescent = true; }
                        const float alpha = need ? ex2_approx(m_run - cand) : 1.0f;
                        if (need) { m_run = cand; l_run *= alpha; }
#pragma unroll
                        for (int c = 0; c < DV; c += 16) {
                            uint32_t v[16];
                            tmem_ld16(o_tmem + c, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                            tmem_st16(o_tmem + c, v);
                        }
                    }
                }
                RG_STAMP(3);
                if constexpr (PSEP && kAttnPStream) {            // P_t's own columns: free once O_t += P_t(G-1) V has retired,
                    if (!quiescent) { mbar_wait(&pv_done[t], (G - 1) & 1); tc_fence_after(); quiescent = true; }   // long ago by now
                }
                // Turn-taking (p.turns, off by default -- measured slower than the stagger, see launch_attn): this warp and the
                // other softmax warp of its SM sub-partition (same quarter qd, other warpgroup) never run their exponentials
                // together -- warpgroup 0's tile G, then warpgroup 1's tile G, then warpgroup 0's tile G + 1 ...
                if (p.turns) {
                    if (t == 0) { if (G > 0) mbar_wait(&exp_done[qd * 2 + 1], (G - 1) & 1); }
                    else mbar_wait(&exp_done[qd * 2 + 0], G & 1);
                }
                const float neg_m = -m_run;
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                uint32_t pk[BKV / 2];
                if constexpr (F16) {
                    // two exponentials per MUFU instruction on the packed logits; the row sum comes out of the P V GEMM
#pragma unroll
                    for (int i = 0; i < BKV; i += 2) {
#if RG_ATTN_EXP_F32
                        // (-DRG_ATTN_EXP_F32=1, not the default.)  ex2.approx.f16x2 is two MUFU.EX2.F16 plus a PRMT, fed by a
                        // conversion: 6 instructions per pair in a chain.  Two fp32 exponentials and ONE packing conversion are
                        // 5 and the SASS is a clean MUFU, MUFU, F2FP stream: 2650 cycles per step instead of 2890 while the
                        // warpgroups stay half a step apart -- but with this form they drift into lock-step (3350 cycles).
                        const float e0 = ex2_approx(fmaf(s[i], sl, neg_m)), e1 = ex2_approx(fmaf(s[i + 1], sl, neg_m));
                        pk[i / 2] = pack_f16x2(e0, e1);
#else
                        const float x0 = fmaf(s[i], sl, neg_m), x1 = fmaf(s[i + 1], sl, neg_m);
                        uint32_t h;
                        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
                        asm("ex2.approx.f16x2 %0, %1;" : "=r"(pk[i / 2]) : "r"(h));
#endif
                        if ((!PSEP || kAttnPStream) && (i & 31) == 30)             // stream each finished 32-key chunk out
                            tmem_st16(p_tmem + (i - 30) / 2, reinterpret_cast<uint32_t (&)[16]>(pk[(i - 30) / 2]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < BKV; i += 4) {
                        const float e0 = ex2_approx(fmaf(s[i], sl, neg_m));
                        const float e1 = ex2_approx(fmaf(s[i + 1], sl, neg_m));
                        const float e2 = ex2_approx(fmaf(s[i + 2], sl, neg_m));
                        const float e3 = ex2_approx(fmaf(s[i + 3], sl, neg_m));
                        l0 += e0; l1 += e1; l2 += e2; l3 += e3;
                        pk[i / 2] = pack_bf16x2(e0, e1);
                        pk[i / 2 + 1] = pack_bf16x2(e2, e3);
                        if ((!PSEP || kAttnPStream) && (i & 31) == 28)             // stream each finished 32-key chunk out
                            tmem_st16(p_tmem + (i - 28) / 2, reinterpret_cast<uint32_t (&)[16]>(pk[(i - 28) / 2]));
                    }
                }
                if (p.turns) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&exp_done[qd * 2 + t]);
                }
                RG_STAMP(4);
                if constexpr (PSEP && !kAttnPStream) {
                    if (!quiescent) { mbar_wait(&pv_done[t], (G - 1) & 1); tc_fence_after(); }
                    RG_STAMP(5);
#pragma unroll
                    for (int c = 0; c < BKV / 2; c += 16) tmem_st16(p_tmem + c, reinterpret_cast<uint32_t (&)[16]>(pk[c]));
                } else {
                    RG_STAMP(5);
                }
                l_run += (l0 + l1) + (l2 + l3);
                tmem_st_wait();
                RG_STAMP(6);
                tc_fence_before();
                mbar_arrive(&p_full[t]);
                RG_STAMP(7);
            }
            // ---- item epilogue: O_t / l -> bf16 -> global
            mbar_wait(&pv_done[t], (g0 + n_kv - 1) & 1);
            tc_fence_after();
            const int tq = qp * 256 + t * 128 + row;
            float l_tot = l_run;
            if constexpr (F16) {                                 // the ones column of V: O_t[:, d] = sum_k P
                uint32_t v[16];
                tmem_ld16(o_tmem + (p.d & ~15), v);
                tmem_ld_wait();
                l_tot = __uint_as_float(v[0]);
#pragma unroll
                for (int i = 1; i < 16; ++i) if ((p.d & 15) == i) l_tot = __uint_as_float(v[i]);
            }
            const float inv = 1.0f / l_tot;
            __nv_bfloat16* dst = p.out + (long long)b * p.osb + (long long)tq * p.ost + (long long)h * p.osh;
#pragma unroll
            for (int c = 0; c < DV; c += 16) {
                uint32_t v[16];
                tmem_ld16(o_tmem + c, v);
                tmem_ld_wait();
                if (tq < p.Nq) {
#pragma unroll
                    for (int g = 0; g < 16; g += 8) {
                        if (c + g < p.d) {
                            *reinterpret_cast<uint4*>(dst + c + g) = make_uint4(
                                pack_bf16x2(__uint_as_float(v[g]) * inv, __uint_as_float(v[g + 1]) * inv),
                                pack_bf16x2(__uint_as_float(v[g + 2]) * inv, __uint_as_float(v[g + 3]) * inv),
                                pack_bf16x2(__uint_as_float(v[g + 4]) * inv, __uint_as_float(v[g + 5]) * inv),
                                pack_bf16x2(__uint_as_float(v[g + 6]) * inv, __uint_as_float(v[g + 7]) * inv));
                        }
                    }
                }
            }
            tc_fence_before();       // O_t reads are ordered before the next item's p_full arrival (-> next PV with acc=0)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

static long long* g_attn_trace = nullptr;
static int g_attn_trace_item = 0;

static int encode_qkv_map(CUtensorMap* m, const void* base, int d, int heads, long long tokens, int B, long long sh,
                          long long st, long long sb, int rows) {
    cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)st * 2, (cuuint64_t)sb * 2};
    cuuint32_t box[4] = {64, 1, (cuuint32_t)rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                             CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int DKA, int DQK, int DV, int BKV, int SBUF, bool PSEP, int STAGES, bool F16, bool CAUSAL = false>
static int launch_attn(const rg_attn_t* a, cudaStream_t stream) {
    using Cfg = AttnCfg<DKA, DQK, DV, BKV, SBUF, PSEP, STAGES, F16>;
    static std::atomic<bool> attr_done[kMaxDevices];
    if (int rc0 = ensure_smem_attr(reinterpret_cast<const void*>(&attention_kernel<DKA, DQK, DV, BKV, SBUF, PSEP, STAGES, F16, CAUSAL>),
                                   Cfg::SMEM_BYTES, attr_done, "cudaFuncSetAttribute(attention_kernel)")) return rc0;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    if ((rc = encode_qkv_map(&p.qmap, a->q, a->d, a->heads, a->Nq, a->B, a->q_stride_h, a->q_stride_t, a->q_stride_b, 128))) return rc;
    if ((rc = encode_qkv_map(&p.kmap, a->k, a->d, a->heads, a->Nk, a->B, a->k_stride_h, a->k_stride_t, a->k_stride_b, BKV))) return rc;
    if ((rc = encode_qkv_map(&p.vmap, a->v, a->d, a->heads, a->Nk, a->B, a->v_stride_h, a->v_stride_t, a->v_stride_b, BKV))) return rc;
    p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
    p.osb = a->o_stride_b; p.ost = a->o_stride_t; p.osh = a->o_stride_h;
    p.Nq = a->Nq; p.Nk = a->Nk; p.d = a->d;
    p.heads = a->heads;
    p.n_qpairs = (a->Nq + 255) / 256;
    const long long items = (long long)p.n_qpairs * a->heads * a->B;
    if (items > 0x7fffffffLL) return set_error(RG_ERR_ARG, "attention: too many work items");
    p.n_items = (int)items;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    p.trace = g_attn_trace;
    p.trace_item = g_attn_trace_item;
    // long key loops: one issuer per warpgroup and the warpgroups half a period apart; short ones (cross attention, the
    // few-token levels) are latency-bound chains where a second issuer only adds contention (profiles/r02_attn_variants.txt)
    p.n_issuers = (PSEP && (a->Nk + BKV - 1) / BKV >= 8) ? 2 : 1;
    // Measured on the 4096-token launch (profiles/r02_attn_variants.txt): stagger 734 us; strict turn-taking of the
    // exponential phases 770 us with the fp32 exponentials (a lone warp needs ~1400 cycles per tile, so 2 x 1500 per step)
    // and 883 us with the f16x2 form; fp32 exponentials + stagger 840 us (2650 cycles per step while the offset holds,
    // but the warpgroups drift into lock-step within an item).
    p.turns = 0;
    p.skew = p.n_issuers == 2 ? kAttnSkewCycles : 0;
#ifdef RG_ATTN_TUNING            /* perf experiments only: never in the product build */
    { const char* e = getenv("RG_ATTN_ISSUERS"); if (e) p.n_issuers = atoi(e) == 2 ? 2 : 1; }
    { const char* e = getenv("RG_ATTN_SKEW"); if (e) p.skew = p.n_issuers == 2 ? atoi(e) : 0; }
    { const char* e = getenv("RG_ATTN_TURNS"); if (e) p.turns = (p.n_issuers == 2 && atoi(e)) ? 1 : 0; }
#endif
    const int grid = p.n_items < sm_count() ? p.n_items : sm_count();
    launch_kernel(attention_kernel<DKA, DQK, DV, BKV, SBUF, PSEP, STAGES, F16, CAUSAL>, dim3(grid), dim3(kAttnThreads), Cfg::SMEM_BYTES, stream, p);
    count_launch();
    return check_launch("attention_kernel");
}

}  // namespace rg

using namespace rg;

// Debug hook (not part of the public header): device buffer of 8 warps x 64 tiles x 8 clock64 stamps written by
// CTA 0 for its first work item; nullptr switches tracing off.
extern "C" void rg_debug_attn_trace(void* dev_buf) { g_attn_trace = reinterpret_cast<long long*>(dev_buf); }
extern "C" void rg_debug_attn_trace_item(int item) { g_attn_trace_item = item; }

extern "C" int rg_attention(const rg_attn_t* a, rg_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->q || !a->k || !a->v || !a->out) return set_error(RG_ERR_ARG, "attention: null pointer");
    if (a->d % 8 || a->d < 8 || a->d > 160) return set_error(RG_ERR_ARG, "attention: head dim must be a multiple of 8 in [8,160]");
    if (a->Nq < 1 || a->Nk < 1) return set_error(RG_ERR_ARG, "attention: empty sequence");
    if (!(a->scale > 0.f)) return set_error(RG_ERR_ARG, "attention: scale must be positive");
    const int64_t st[] = {a->q_stride_b, a->q_stride_t, a->q_stride_h, a->k_stride_b, a->k_stride_t, a->k_stride_h,
                          a->v_stride_b, a->v_stride_t, a->v_stride_h, a->o_stride_b, a->o_stride_t, a->o_stride_h};
    for (int64_t s : st)
        if (s % 8) return set_error(RG_ERR_ARG, "attention: strides must be multiples of 8 elements");
    if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
         reinterpret_cast<uintptr_t>(a->out)) & 15)
        return set_error(RG_ERR_ARG, "attention: pointers must be 16-byte aligned");
    const int dv = (a->d + 15) / 16 * 16;
    // template arguments: DKA (64-wide atoms of d), DQK (k extent of Q K^T), DV (n extent of P V), BKV, SBUF, PSEP, STAGES, F16
    if (a->dtype == RG_DT_F16) {
        if (a->causal) return set_error(RG_ERR_ARG, "attention: the causal path is bf16 only");
        // fp16 operands: DV includes the ones column at index d (inside the padding for d = 40, one more 16-column step else)
        if (a->d == 40) return launch_attn<1, 48, 48, 128, 1, true, 4, true>(a, stream);
        if (a->d == 80) return launch_attn<2, 80, 96, 128, 1, false, 2, true>(a, stream);
        if (a->d == 160) return launch_attn<3, 160, 176, 64, 1, false, 2, true>(a, stream);
        return set_error(RG_ERR_ARG, "attention: the fp16 path supports head dims 40, 80 and 160");
    }
    if (a->dtype != RG_DT_BF16) return set_error(RG_ERR_ARG, "attention: dtype must be RG_DT_BF16 or RG_DT_F16");
    if (a->causal) {
        // the CLIP text encoder's attention (12 heads of 64, 77 tokens); query i attends keys 0..i
        if (a->Nq != a->Nk) return set_error(RG_ERR_ARG, "attention: causal needs Nq == Nk");
        if (dv <= 64) return launch_attn<1, 64, 64, 128, 1, true, 4, false, true>(a, stream);
        return set_error(RG_ERR_ARG, "attention: the causal path supports head dims up to 64 (bf16)");
    }
    if (dv <= 48) return launch_attn<1, 48, 48, 128, 1, true, 4, false>(a, stream);
    if (dv <= 64) return launch_attn<1, 64, 64, 128, 1, true, 4, false>(a, stream);
    if (dv <= 80) return launch_attn<2, 80, 80, 128, 1, false, 2, false>(a, stream);
    if (dv <= 128) return launch_attn<2, 128, 128, 128, 1, false, 2, false>(a, stream);
    return launch_attn<3, 160, 160, 64, 1, false, 2, false>(a, stream);
}
