// K1-K4: implicit-GEMM convolution / linear layer for sm_100a, CTA-pair (cta_group::2) edition.
//
//   D[256 x N_TILE] (fp32, TMEM of two SMs)  +=  A[256 x 64] (bf16)  x  B[N_TILE x 64]^T (bf16)
//
// Why pairs: on B200 a dense GEMM is bound by the L2 -> SM fill rate (about 43 B/clk/SM chip-wide) long before the
// tensor pipe saturates.  A 128 x 160 single-CTA tile needs 115 B/clk/SM at full MMA rate; a 256 x 320 tile shared
// by the two SMs of a TPC needs 58 B/clk/SM, because each CTA fetches only its own 128 rows of A and HALF of the B
// tile, and tcgen05.mma.cta_group::2 reads both halves.
//
// * A is never materialised: for every filter tap the TMA engine fetches a {64 ch, TW, TH, TN} box of the
//   channels-last activation at the tap's spatial offset; out-of-bounds coordinates are zero-filled by the
//   hardware, which is exactly the convolution's zero padding.  Stride-2 convolutions read four "parity"
//   views of the input (one tensor map each), so they are plain shifted boxes as well.
// * B is the packed weight matrix [Cout][taps*Cin (+C2)], K-major.  A tile is NC chunks of BNC columns (one MMA
//   per chunk and k-step); CTA r of the pair holds rows [r*BNC/2, (r+1)*BNC/2) of every chunk.
// * Persistent, warp-specialised, one cluster of two CTAs per TPC: warp 0 = TMA producer (both CTAs; all
//   completion bytes are signalled on the leader's barrier), warp 1 = tcgen05.mma issuer (leader CTA only;
//   commits are multicast to both CTAs), warps 2..9 = epilogue of the CTA's own 128 accumulator rows.
// * The accumulator chunks live in a ring of TMEM slots, so the epilogue of tile i overlaps the main loop of
//   tile i+1 (fully for NC = 1, from the second chunk on for the 2 x 160 tile that uses 3 slots).
// * Epilogue (epilogue_tma below): every epilogue warp owns 32 accumulator rows and walks 16-column units --
//   residual box prefetched by TMA, tcgen05.ld, combine in shared memory in place (scale, bias, per-image bias,
//   residual, SiLU / GEGLU), TMA store of the fp32 and / or bf16 box; unit code specialised per epilogue mode at
//   compile time.  Outputs TMA cannot address (Cout = 3 fp32, unaligned pitches) take the direct epilogue: TMEM ->
//   registers -> per-warp shared-memory transpose -> row-contiguous global accesses.
// * Few-pixel levels with a long K (32x32 and below) run deterministic split-K (x2 / x4 / x8 by per-image geometry, so
//   results stay batch-invariant): every K slice dumps its fp32 partial tile into a workspace, the last slice to arrive
//   (per-warp arrival counters) adds the partials in slice order and runs the normal epilogue (epilogue_splitk).
// What bounds it (measured): all traffic through L2 -- operand fetches AND epilogue boxes --
// shares about 6300 B/clk chip-wide; the short-K GEMMs of the transformer blocks sit on that limit, not on the MMA.
#include <stdlib.h>
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
    int n_pairs_m, n_tiles_n;         // cluster tiles: pairs of 128-pixel tiles x N_TILE-wide column tiles
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
    // TMA epilogue (epi_tma != 0): per-warp boxes of {16 columns, 32 pixels}
    int epi_tma;
    int prim_f32;                     // dtype of the primary staging buffer (residual in / same-dtype output in place)
    int prim_store;                   // an output of the primary dtype exists
    int sec_store;                    // bf16 output next to an fp32 primary buffer
    int ksplit, kper;                 // split-K: work item (tile, ks) covers k-blocks [ks*kper, (ks+1)*kper)
    float* ws_part;                   // split-K partial tiles: [out tile][ks][rank][epilogue warp][unit][4][32 lanes] float4
    int* ws_cnt;                      // split-K arrival counters: [out tile][rank][epilogue warp], zero between launches
    int epi_mode;                     // EPI_* bit set when it matches a specialised epilogue, else EPI_GENERIC
    int dbg;                          // RG_GEMM_DEBUG bits (perf experiments only): 1 skip units, 2 skip stores, 4 skip residual
    int out_f16;                      // the 16-bit output is fp16 (attention operands), not bf16
    int out_cols;                     // Cout, or Cout / 2 for GEGLU
    CUtensorMap resmap, pmap, hmap;   // residual load, primary store, secondary (bf16) store
    // parity-split upsample in one launch (rg_conv_t::parities == 4): column tiles [par * tiles_per_par, ..) belong to output
    // parity par = 2 py + px: its taps are shifted by (px, py) and its boxes go through its own strided view of the output
    int n_par, tiles_per_par;
    CUtensorMap pmap_par[3];          // primary store maps of parities 1..3 (parity 0: pmap)
};

// warp 0 TMA, warp 1 MMA, warps 2..EW+1 epilogue (EW / 4 per TMEM lane quarter).  EW = 8 for main-loop-bound shapes
// (deep operand pipeline), EW = 16 for the short-K GEMMs of the transformer blocks whose epilogue is the bottleneck.

template <int BNC, int NC, int EW>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;                   // per CTA; the pair computes 256 rows
    static constexpr int N_TILE = BNC * NC;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_CHUNK_BYTES = (BNC / 2) * BK * 2;  // this CTA's half of one chunk
    static constexpr int STAGE_BYTES = A_BYTES + NC * B_CHUNK_BYTES;
    // per epilogue warp: NBUF primary buffers [32 rows][16 fp32] (residual in, result in place, TMA store out) and
    // HBUF secondary buffers [32 rows][16 bf16]; the direct (non-TMA) path uses the first two primaries to transpose
    static constexpr int NBUF = (NC == 1 && EW == 8) ? 3 : 2;
    static constexpr int HBUF = (NC == 1 && EW == 8) ? 2 : 1;
    static constexpr int WARP_STAGING = NBUF * 2048 + HBUF * 1024;
    static constexpr int STAGING_BYTES = EW * WARP_STAGING;
    static constexpr int THREADS = (EW + 2) * 32;
    static constexpr int PARTS = EW / 4;                        // epilogue warps per TMEM lane quarter
    static_assert(EW == 8 || EW == 16, "EW");
    static constexpr int BAR_BYTES = 512;
    static constexpr int MAX_STAGES = (227 * 1024 - 1024 - STAGING_BYTES - BAR_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
    static_assert(STAGES >= 4, "pipeline too shallow");
    static constexpr int SLOTS = 512 / BNC > 4 ? 4 : 512 / BNC;
    static constexpr int TMEM_COLS = SLOTS * BNC <= 32 ? 32 : SLOTS * BNC <= 64 ? 64 : SLOTS * BNC <= 128 ? 128
                                   : SLOTS * BNC <= 256 ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;
    static_assert(BNC % 32 == 0 && BNC >= 32 && BNC <= 256, "BNC");
    static_assert(SLOTS >= NC + (NC > 1 ? 1 : 1), "TMEM ring too small");
    static_assert((2 * STAGES + 2 * SLOTS + EW * NBUF) * 8 + 8 <= BAR_BYTES, "barrier area");
    static_assert(B_CHUNK_BYTES % 1024 == 0, "operand tiles must stay 1024-B aligned");
};

__device__ __forceinline__ float apply_act(float x, int act) { return act == 1 ? silu_f(x) : (act == 3 ? fmaxf(x, 0.f) : x); }

// geometry of the 128 accumulator rows a CTA owns for one tile
struct RowGeom {
    long long off[4];      // element offset of the output pixel of read-phase row i (rows rsub + 8 i)
    int n[4];
    unsigned valid;        // bit i: read-phase row i is a real output pixel
};

// ---------------------------------------------------------------------------------------------------------------
// TMA epilogue.  Every epilogue warp owns 32 accumulator rows (its TMEM lane quarter) and walks its share of the
// tile's 16-column units.  Per unit: the residual box {16 cols, 32 pixels} was prefetched by TMA into the warp's
// primary buffer one unit (or more) ahead; the accumulator comes from TMEM with one tcgen05.ld; thread r combines
// row r in shared memory IN PLACE (scale, bias, per-image bias, residual, activation) and the warp's lane 0 hands the
// finished box to a TMA store.  No global load or store is issued by the math threads, nothing waits for a memory
// round trip, every byte moves in full 32/64-byte row segments, and ragged tile edges are clipped by the tensor map.
// Requires Cout % N_TILE == 0 (every unit of every chunk is real), which holds for all UNet / VAE body layers.
// MODE: bit set of EPI_* known at compile time (straight-line unit code, four independent column groups in flight),
// or EPI_GENERIC to read every flag from the parameters at run time.
enum : int { EPI_GEGLU = 1, EPI_PRIM_F32 = 2, EPI_RES = 4, EPI_PRIM_STORE = 8, EPI_SEC_STORE = 16, EPI_BIAS = 32,
             EPI_BIASN = 64, EPI_SILU = 128, EPI_SCALE = 256, EPI_F16 = 512, EPI_GENERIC = 1 << 20 };

template <int BNC, int NC, int EW, int MODE>
__device__ __noinline__ void epilogue_tma(const GemmParams& p, uint8_t* staging, uint64_t* acc_full, uint64_t* acc_empty,
                                             uint64_t* res_bar_all, uint32_t tmem_base, uint32_t rank, int cluster_id,
                                             int n_clusters, int warp, int lane) {
    using Cfg = GemmCfg<BNC, NC, EW>;
    constexpr int NBUF = Cfg::NBUF, HBUF = Cfg::HBUF;
    const int ew = warp - 2;
    constexpr int PARTS = Cfg::PARTS;
    const int q = warp & 3, part = ew >> 2;               // units part, part + PARTS, .. of every chunk are this warp's
    // everything the unit loop needs, in registers (compile-time constants unless MODE == EPI_GENERIC)
    constexpr bool G = MODE == EPI_GENERIC;
    const bool geglu = G ? p.act == 2 : (MODE & EPI_GEGLU) != 0;
    const bool silu = G ? p.act == 1 : (MODE & EPI_SILU) != 0;
    const bool relu = G && p.act == 3;                    // LPIPS feature convs: run-time-flag epilogue only
    const bool prim_f32 = G ? p.prim_f32 != 0 : (MODE & EPI_PRIM_F32) != 0;
    const bool prim_store = G ? p.prim_store != 0 : (MODE & EPI_PRIM_STORE) != 0;
    const bool sec_store = G ? p.sec_store != 0 : (MODE & EPI_SEC_STORE) != 0;
    const bool has_res = G ? (p.res != nullptr && !(p.dbg & 4)) : (MODE & EPI_RES) != 0;
    const bool skip_units = G && (p.dbg & 1) != 0, skip_store = G && (p.dbg & 2) != 0;
    const float scale = (G || (MODE & EPI_SCALE)) ? p.scale : 1.0f;
    const float* const bias = (G || (MODE & EPI_BIAS)) ? p.bias : nullptr;
    const bool use_bn = G ? p.bias_n != nullptr : (MODE & EPI_BIASN) != 0;
    const bool out_f16 = G ? p.out_f16 != 0 : (MODE & EPI_F16) != 0;
    auto pack16 = [&](float lo, float hi) { return out_f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); };
    const int gsh = geglu ? 1 : 0;                        // output column = GEMM column >> gsh
    const int UW = geglu ? 32 : 16;                       // accumulator columns per unit
    const int UPC = BNC / UW;                             // units per chunk
    const int CNT = (UPC - part + PARTS - 1) / PARTS;     // units per chunk handled by this warp
    uint8_t* const pbuf0 = staging + ew * Cfg::WARP_STAGING;
    uint8_t* const hbuf0 = pbuf0 + NBUF * 2048;
    uint64_t* const res_bar = res_bar_all + ew * NBUF;
    const uint32_t acc_empty_leader = mapa_shared(smem_u32(acc_empty), 0);
    const int total_tiles = p.n_pairs_m * p.n_tiles_n;   // (split-K launches take epilogue_splitk instead)
    const int TW = 1 << p.lw, TH = 1 << p.lh, lwh = p.lw + p.lh, TN = 128 >> lwh;
    const int row0 = q * 32;
    const int res_bytes = prim_f32 ? 2048 : 1024;
    const int sw = (lane >> 1) & 3;                       // SWIZZLE_64B phase of this lane's 64-byte fp32 row
    const int fo = lane * 64, ho = lane * 32;
    const int hsw = (lane >> 2) & 1;                      // SWIZZLE_32B phase of this lane's 32-byte bf16 row

    // box origin (w, h, n) of this warp's 32 rows in tile `tile`
    auto box_origin = [&](int tile, int& cw, int& ch, int& cn, int& n_tile, int& tni) {
        const int mp = tile / p.n_tiles_n;
        n_tile = tile - mp * p.n_tiles_n;
        const int m_tile = 2 * mp + (int)rank;
        const int twi = m_tile % p.tiles_w;
        const int rest = m_tile / p.tiles_w;
        const int thi = rest % p.tiles_h;
        tni = rest / p.tiles_h;
        cw = twi * TW + (row0 & (TW - 1));
        ch = thi * TH + ((row0 >> p.lw) & (TH - 1));
        cn = tni * TN + (row0 >> lwh);
    };

    uint32_t A = 0;                                       // units processed so far by this warp
    uint32_t cc = 0;
    int cw = 0, ch = 0, cn = 0, n_tile = 0, tni = 0;
    if (cluster_id < total_tiles) box_origin(cluster_id, cw, ch, cn, n_tile, tni);
    pdl_wait();                                           // first global access below (residual prefetch, bias, stores)
    if (has_res && lane == 0 && CNT > 0 && cluster_id < total_tiles && !skip_units) {      // very first residual box
        mbar_expect_tx(&res_bar[0], res_bytes);
        tma_load_4d(pbuf0, &p.resmap, &res_bar[0], (n_tile * Cfg::N_TILE + part * UW) >> gsh, cw, ch, cn);
    }
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, cc += NC) {
        // next tile's origin: target of the residual prefetch issued from this tile's last unit
        const bool has_next = tile + n_clusters < total_tiles;
        int cw2 = 0, ch2 = 0, cn2 = 0, n_tile2 = 0, tni2 = 0;
        if (has_next) box_origin(tile + n_clusters, cw2, ch2, cn2, n_tile2, tni2);
        int n_row = tni * TN + ((row0 + lane) >> lwh);
        if (n_row >= p.N) n_row = p.N - 1;                // padding rows: any valid image (their result is clipped)
        const float* const bn_row = use_bn ? p.bias_n + (long long)n_row * p.bias_n_ld : nullptr;
        const float* const bias_t = bias;
        const int par = p.n_par > 1 ? n_tile / p.tiles_per_par : 0;          // parity-split upsample: this tile's output view
        const int n_loc = n_tile - par * p.tiles_per_par;
        const CUtensorMap* const pm = par ? &p.pmap_par[par - 1] : &p.pmap;
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
            const uint32_t slot = (cc + c) % Cfg::SLOTS, use = (cc + c) / Cfg::SLOTS;
            mbar_wait(&acc_full[slot], use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)row0 << 16) + slot * BNC;
            const int gcol0 = n_loc * Cfg::N_TILE + c * BNC;
#pragma unroll 1
            for (int k = 0; k < CNT && !skip_units; ++k, ++A) {
                const int u = part + PARTS * k;
                const int gcol = gcol0 + u * UW;                               // GEMM column (bias index)
                const int ocol = gcol >> gsh;                                  // output column
                const int buf = A % NBUF;
                uint8_t* const pb = pbuf0 + buf * 2048;
                uint8_t* const hb = hbuf0 + (A % HBUF) * 1024;
                // ---- lane 0: the buffers about to be refilled are no longer read by an earlier store; prefetch the
                //      next unit's residual
                if (lane == 0) {
                    bulk_wait_read<NBUF - 2>();
                    if (has_res) {
                        const int nb = (A + 1) % NBUF;
                        if (k + 1 < CNT) {
                            mbar_expect_tx(&res_bar[nb], res_bytes);
                            tma_load_4d(pbuf0 + nb * 2048, &p.resmap, &res_bar[nb], (gcol + PARTS * UW) >> gsh, cw, ch, cn);
                        } else if (c + 1 < NC) {
                            mbar_expect_tx(&res_bar[nb], res_bytes);
                            tma_load_4d(pbuf0 + nb * 2048, &p.resmap, &res_bar[nb], (gcol0 + BNC + part * UW) >> gsh, cw, ch, cn);
                        } else if (has_next) {
                            mbar_expect_tx(&res_bar[nb], res_bytes);
                            tma_load_4d(pbuf0 + nb * 2048, &p.resmap, &res_bar[nb], (n_tile2 * Cfg::N_TILE + part * UW) >> gsh,
                                        cw2, ch2, cn2);
                        }
                    }
                }
                __syncwarp();
                // ---- accumulator
                uint32_t va[16];
                tmem_ld16(taddr + u * UW, va);
                if (geglu) {
                    uint32_t vg[16];
                    tmem_ld16(taddr + u * UW + 16, vg);
                    tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
                        if (bias_t) {
                            ba = __ldg(reinterpret_cast<const float4*>(bias_t + gcol + 4 * j));
                            bg = __ldg(reinterpret_cast<const float4*>(bias_t + gcol + 16 + 4 * j));
                        }
                        const float y0 = fmaf(__uint_as_float(va[4 * j]), scale, ba.x) * gelu_fast_f(fmaf(__uint_as_float(vg[4 * j]), scale, bg.x));
                        const float y1 = fmaf(__uint_as_float(va[4 * j + 1]), scale, ba.y) * gelu_fast_f(fmaf(__uint_as_float(vg[4 * j + 1]), scale, bg.y));
                        const float y2 = fmaf(__uint_as_float(va[4 * j + 2]), scale, ba.z) * gelu_fast_f(fmaf(__uint_as_float(vg[4 * j + 2]), scale, bg.z));
                        const float y3 = fmaf(__uint_as_float(va[4 * j + 3]), scale, ba.w) * gelu_fast_f(fmaf(__uint_as_float(vg[4 * j + 3]), scale, bg.w));
                        pk[2 * j] = pack16(y0, y1); pk[2 * j + 1] = pack16(y2, y3);
                    }
                    *reinterpret_cast<uint4*>(pb + ho + (hsw << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(pb + ho + ((hsw ^ 1) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                } else {
                    tmem_ld_wait();
                    if (has_res) mbar_wait(&res_bar[buf], (A / NBUF) & 1);
                    // bf16 boxes are [32 rows][32 B] with the SWIZZLE_32B pattern (16-byte halves of rows 4..7, 12..15, ..
                    // swapped): two conflict-free 16-byte accesses per row
                    uint32_t pk[8], rb[8];
                    if (!prim_f32 && has_res) {
                        const uint4 r0 = *reinterpret_cast<const uint4*>(pb + ho + (hsw << 4));
                        const uint4 r1 = *reinterpret_cast<const uint4*>(pb + ho + ((hsw ^ 1) << 4));
                        rb[0] = r0.x; rb[1] = r0.y; rb[2] = r0.z; rb[3] = r0.w; rb[4] = r1.x; rb[5] = r1.y; rb[6] = r1.z; rb[7] = r1.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float y0 = __uint_as_float(va[4 * j]) * scale, y1 = __uint_as_float(va[4 * j + 1]) * scale;
                        float y2 = __uint_as_float(va[4 * j + 2]) * scale, y3 = __uint_as_float(va[4 * j + 3]) * scale;
                        if (bias_t) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias_t + gcol + 4 * j));
                            y0 += b.x; y1 += b.y; y2 += b.z; y3 += b.w;
                        }
                        if (bn_row) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bn_row + gcol + 4 * j));
                            y0 += b.x; y1 += b.y; y2 += b.z; y3 += b.w;
                        }
                        if (prim_f32) {
                            float4* const slot4 = reinterpret_cast<float4*>(pb + fo + ((j ^ sw) << 4));
                            if (has_res) { const float4 r = *slot4; y0 += r.x; y1 += r.y; y2 += r.z; y3 += r.w; }
                            if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
                            if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
                            if (prim_store) *slot4 = make_float4(y0, y1, y2, y3);
                            if (sec_store) { pk[2 * j] = pack16(y0, y1); pk[2 * j + 1] = pack16(y2, y3); }
                        } else {
                            if (has_res) {
                                const float2 f0 = unpack_bf16x2(rb[2 * j]), f1 = unpack_bf16x2(rb[2 * j + 1]);
                                y0 += f0.x; y1 += f0.y; y2 += f1.x; y3 += f1.y;
                            }
                            if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
                            if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
                            pk[2 * j] = pack16(y0, y1); pk[2 * j + 1] = pack16(y2, y3);
                        }
                    }
                    if (!prim_f32 || sec_store) {
                        uint8_t* const bb = prim_f32 ? hb : pb;
                        *reinterpret_cast<uint4*>(bb + ho + (hsw << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(bb + ho + ((hsw ^ 1) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                // ---- hand the finished box(es) to the TMA engine
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && !skip_store) {
                    if (prim_store) tma_store_4d(pm, pb, ocol, cw, ch, cn);
                    if (sec_store) tma_store_4d(&p.hmap, hb, ocol, cw, ch, cn);
                    bulk_commit();
                }
            }
            // this warp is done with the slot: one arrival per warp on the LEADER's barrier
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_leader + slot * 8);
        }
        cw = cw2; ch = ch2; cn = cn2; n_tile = n_tile2; tni = tni2;
    }
    if (lane == 0) bulk_wait_read<0>();                   // the stores have read their shared-memory source: the CTA may exit
    __syncwarp();                                         // (the writes themselves are flushed by the end of the grid)
}


// sum[j] = partial[0][j] + partial[1][j] + ... in slice order; all KS * 4 loads (L2, bypassing L1) are in flight together
template <int KS>
__device__ __forceinline__ void splitk_sum(const float4* src, size_t pitch4, float4 (&sum)[4]) {
    float4 v[KS][4];
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[s][j] = __ldcg(src + s * pitch4 + j * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 a = v[0][j];
#pragma unroll
        for (int s = 1; s < KS; ++s) { a.x += v[s][j].x; a.y += v[s][j].y; a.z += v[s][j].z; a.w += v[s][j].w; }
        sum[j] = a;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Deterministic split-K epilogue (NC = 1, EW = 8), used when a layer has too few output tiles to fill the SMs (UNet
// batch 2: everything below the 64x64 level).  Work item = (output tile t2, K slice ks), ks fastest, dealt round-robin,
// so the slices of one tile run on neighbouring clusters at the same time.  Every epilogue warp owns 32 accumulator rows
// and its share of the tile's 16-column units, exactly as in epilogue_tma:
//   phase 1  dump the raw fp32 accumulator units into the workspace region of (t2, ks, rank, warp) -- lane-interleaved
//            float4s, every store instruction writes 512 contiguous bytes -- release the TMEM slot, fence, bump the
//            arrival counter of (t2, rank, warp);
//   phase 2  only the warp whose arrival was the last of the ksplit slices: read the partials of ALL slices back (L2
//            hits) and add them in slice order 0..ksplit-1 -- the order never depends on who arrived when, so a launch is
//            bitwise reproducible -- then scale / bias / per-image bias / residual / activation and the TMA store(s).
//            The counter is reset by the warp that consumed it.
// No CTA ever waits for another one (the last arriver does the work), so there is no forward-progress hazard.
// (Tried and dropped: a "local" mode that ran all slices of a tile in one cluster with register sums, to make the result
// independent of the batch size bit for bit -- it forces the 160-column tile on large batches, where the 2 x 160 tile's
// halved operand traffic is worth 15 % of a UNet evaluation; profiles/r02_layers_splitk_b8.txt.)
template <int BNC, int EW>
__device__ __noinline__ void epilogue_splitk(const GemmParams& p, uint8_t* staging, uint64_t* acc_full, uint64_t* acc_empty,
                                             uint32_t tmem_base, uint32_t rank, int cluster_id, int n_clusters, int warp,
                                             int lane) {
    using Cfg = GemmCfg<BNC, 1, EW>;
    constexpr int NBUF = Cfg::NBUF, HBUF = Cfg::HBUF, PARTS = Cfg::PARTS;
    static_assert(NBUF >= 2, "split-K epilogue rotates two primary store buffers");
    constexpr int UPC = BNC / 16;                         // 16-column units per tile
    constexpr int CNT_MAX = (UPC + PARTS - 1) / PARTS;    // units per warp (region pitch)
    constexpr int REGION = CNT_MAX * 512;                 // floats per (tile, slice, rank, warp)
    const int ew = warp - 2, q = warp & 3, part = ew >> 2;
    const int CNT = (UPC - part + PARTS - 1) / PARTS;
    uint8_t* const pbuf0 = staging + ew * Cfg::WARP_STAGING;
    uint8_t* const hbuf0 = pbuf0 + NBUF * 2048;
    const uint32_t acc_empty_leader = mapa_shared(smem_u32(acc_empty), 0);
    const int total_tiles = p.n_pairs_m * p.n_tiles_n * p.ksplit;
    const int TW = 1 << p.lw, TH = 1 << p.lh, lwh = p.lw + p.lh, TN = 128 >> lwh;
    const int row0 = q * 32, row = row0 + lane;
    const int sw = (lane >> 1) & 3, fo = lane * 64, ho = lane * 32, hsw = (lane >> 2) & 1;
    const bool prim_f32 = p.prim_f32 != 0, prim_store = p.prim_store != 0, sec_store = p.sec_store != 0;
    const bool silu = p.act == 1, relu = p.act == 3, out_f16 = p.out_f16 != 0, res_f32 = p.res_f32 != 0;
    const float scale = p.scale;
    // `p` lives in the kernel parameter space and is reached through a generic pointer here: every p.xxx is a full
    // memory round trip (~1 us).  Everything the loops need is read ONCE into registers -- four dependent parameter
    // loads per unit made the fix-up 25 us long (profiles/r02_splitk_experiments.txt).
    const int ksplit = p.ksplit, n_tiles_n = p.n_tiles_n, tiles_w = p.tiles_w, tiles_h = p.tiles_h, lw = p.lw;
    const int pOW = p.OW, pOH = p.OH, pN = p.N, dbg = p.dbg;
    const long long osn = p.osn, osh = p.osh, osw = p.osw, bias_n_ld = p.bias_n_ld;
    const float* const bias = p.bias;
    const float* const bias_n = p.bias_n;
    const void* const res = p.res;
    float* const ws_part = p.ws_part;
    int* const ws_cnt = p.ws_cnt;
    const CUtensorMap* const pmap = &p.pmap;
    const CUtensorMap* const hmap = &p.hmap;
    auto pack16 = [&](float lo, float hi) { return out_f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); };
    const size_t slice_pitch = (size_t)2 * EW * REGION;   // floats between slice s and s + 1 of one output tile
    uint32_t cc = 0, A = 0;
    pdl_wait();                                           // parameters are in registers; global memory from here on
#pragma unroll 1
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++cc) {
        const int ks = tile % ksplit, t2 = tile / ksplit;
        const int mp = t2 / n_tiles_n, n_tile = t2 - mp * n_tiles_n;
        const int m_tile = 2 * mp + (int)rank;
        const int twi = m_tile % tiles_w;
        const int rest = m_tile / tiles_w;
        const int thi = rest % tiles_h, tni = rest / tiles_h;
        const int cw = twi * TW + (row0 & (TW - 1)), ch = thi * TH + ((row0 >> lw) & (TH - 1)), cn = tni * TN + (row0 >> lwh);
        const int ow = twi * TW + (row & (TW - 1)), oh = thi * TH + ((row >> lw) & (TH - 1)), n = tni * TN + (row >> lwh);
        const bool valid = ow < pOW && oh < pOH && n < pN;
        const bool any_valid = __any_sync(0xffffffffu, valid);       // all-padding warps (odd last tile, M < 256) skip the work
        const uint32_t slot = cc % Cfg::SLOTS, use = cc / Cfg::SLOTS;
        mbar_wait(&acc_full[slot], use & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)row0 << 16) + slot * BNC;
        float* const tile_base = ws_part + ((size_t)t2 * ksplit * 2 + rank) * EW * REGION + (size_t)ew * REGION;
        if (any_valid && !(dbg & 8)) {
            // ---- phase 1: dump this slice's partial units  (dbg bits: timing experiments, RG_GEMM_TUNING builds)
            float4* const my = reinterpret_cast<float4*>(tile_base + (size_t)ks * slice_pitch);
#pragma unroll 1
            for (int k = 0; k < CNT; ++k) {
                uint32_t va[16];
                tmem_ld16(taddr + (part + PARTS * k) * 16, va);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    __stcg(my + k * 128 + j * 32 + lane, make_float4(__uint_as_float(va[4 * j]), __uint_as_float(va[4 * j + 1]),
                                                                      __uint_as_float(va[4 * j + 2]), __uint_as_float(va[4 * j + 3])));
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader + slot * 8);
        if (!any_valid || (dbg & 16)) continue;
        // ---- publish, elect the last arriver
        if (!(dbg & 64)) __threadfence();
        __syncwarp();
        int last = 0;
        if (dbg & 128) {
            last = ks == ksplit - 1;
        } else if (lane == 0) {
            int* const c = ws_cnt + ((size_t)t2 * 2 + rank) * EW + ew;
            last = atomicAdd(c, 1) == ksplit - 1;
            if (last) atomicExch(c, 0);                    // self-resetting: every slice of this launch has arrived
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) continue;
        if (!(dbg & 256)) __threadfence();
        // ---- finish the tile: ordered sums -> scale / bias / residual / activation -> TMA store(s)
        const long long off = valid ? (long long)n * osn + (long long)oh * osh + (long long)ow * osw : 0;
        const float* const bn_row = bias_n ? bias_n + (long long)(n < pN ? n : pN - 1) * bias_n_ld : nullptr;
        const float4* const part0 = reinterpret_cast<const float4*>(tile_base);
#pragma unroll 1
        for (int k = 0; k < CNT; ++k) {
            const int gcol = n_tile * BNC + (part + PARTS * k) * 16;
            uint8_t* const pb = pbuf0 + (A % 2) * 2048;
            uint8_t* const hb = hbuf0 + (A % HBUF) * 1024;
            if (lane == 0) {                               // the store that last used these buffers has drained them
                if (HBUF >= 2 || !sec_store) bulk_wait_read<1>(); else bulk_wait_read<0>();
            }
            __syncwarp();
            // every partial of the unit is requested before the first one is used, then added in slice order
            // (fixed, independent of arrival order); the slice count is a compile-time constant of splitk_sum (2 / 4 / 8)
            // -- a run-time count with predicated loads tripled the instruction count of this loop, and one warp per SM
            // sub-partition executes it alone
            float4 sum[4];
            const float4* const src = part0 + k * 128 + lane;
            if (dbg & 32) splitk_sum<1>(src, slice_pitch / 4, sum);
            else if (ksplit == 8) splitk_sum<8>(src, slice_pitch / 4, sum);
            else if (ksplit == 4) splitk_sum<4>(src, slice_pitch / 4, sum);
            else splitk_sum<2>(src, slice_pitch / 4, sum);
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 a = sum[j];
                float y0 = a.x * scale, y1 = a.y * scale, y2 = a.z * scale, y3 = a.w * scale;
                if (bias) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + gcol + 4 * j));
                    y0 += bv.x; y1 += bv.y; y2 += bv.z; y3 += bv.w;
                }
                if (bn_row) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bn_row + gcol + 4 * j));
                    y0 += bv.x; y1 += bv.y; y2 += bv.z; y3 += bv.w;
                }
                if (res && valid) {
                    if (res_f32) {
                        const float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(res) + off + gcol + 4 * j);
                        y0 += r.x; y1 += r.y; y2 += r.z; y3 += r.w;
                    } else {
                        const uint2 r2 = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(res) + off + gcol + 4 * j);
                        const float2 f0 = unpack_bf16x2(r2.x), f1 = unpack_bf16x2(r2.y);
                        y0 += f0.x; y1 += f0.y; y2 += f1.x; y3 += f1.y;
                    }
                }
                if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
                if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
                if (prim_f32 && prim_store) *reinterpret_cast<float4*>(pb + fo + ((j ^ sw) << 4)) = make_float4(y0, y1, y2, y3);
                pk[2 * j] = pack16(y0, y1); pk[2 * j + 1] = pack16(y2, y3);
            }
            if (!prim_f32 || sec_store) {
                uint8_t* const bb = prim_f32 ? hb : pb;
                *reinterpret_cast<uint4*>(bb + ho + (hsw << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(bb + ho + ((hsw ^ 1) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(dbg & 512)) {
                if (prim_store) tma_store_4d(pmap, pb, gcol, cw, ch, cn);
                if (sec_store) tma_store_4d(hmap, hb, gcol, cw, ch, cn);
                bulk_commit();
            }
            ++A;
        }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// In-cluster split-K (CKS = 2 or 4 K slices; NC = 1, EW = 16): the few-tile, long-K layers of the small UNet batches.
// A cluster of CKS CTA pairs computes ONE output tile, pair s the K slice s, and the partial tiles meet in shared
// memory instead of in an L2 workspace:
//   send   after a cluster barrier (every pair's MMAs have retired, so the operand rings are free), every epilogue warp
//          reads its accumulator units from TMEM and stores them into the ring of the CTA that OWNS its 32-row quarter
//          (quarter q belongs to pair q % CKS, same rank) -- st.shared::cluster, lane-interleaved float4s;
//   final  after a second cluster barrier, every CTA holds the CKS partials of the row quarters it owns: its 16 warps take
//          one 16-column unit each, add the partials in slice order 0..CKS-1 (fixed: bitwise reproducible), apply scale /
//          bias / per-image bias / residual / activation and hand the box to a TMA store.
// Against the workspace version (epilogue_splitk) this removes the dump to L2, two gpu-scope fences, the arrival atomics
// and the serial chain of L2 round trips in the last-arriving warp: the fix-up costs two cluster barriers and one unit
// per warp.  One tile per cluster and no persistence (grid = tiles x 2 CKS CTAs), so both barriers are executed exactly
// once by every thread of the cluster.
template <int BNC, int EW, int CKS>
__device__ __forceinline__ void cluster_splitk_send(uint8_t* ring, uint32_t tmem_base, uint32_t rank, uint32_t pair, int warp, int lane) {
    constexpr int UPC = BNC / 16, PARTS = EW / 4, QPP = 4 / CKS;      // units per tile, warps per quarter, quarters per pair
    const int ew = warp - 2, q = warp & 3, part = ew >> 2;
    const uint32_t owner = (uint32_t)(q % CKS), q_local = (uint32_t)(q / CKS);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t dst0 = mapa_shared(smem_u32(ring), owner * 2 + rank) + ((pair * QPP + q_local) * UPC) * 2048u + lane * 16u;
#pragma unroll 1
    for (int u = part; u < UPC; u += PARTS) {
        uint32_t va[16];
        tmem_ld16(taddr + u * 16, va);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
            st_cluster_f4(dst0 + u * 2048u + j * 512u, make_float4(__uint_as_float(va[4 * j]), __uint_as_float(va[4 * j + 1]),
                                                                   __uint_as_float(va[4 * j + 2]), __uint_as_float(va[4 * j + 3])));
    }
}

template <int BNC, int EW, int CKS>
__device__ __noinline__ void cluster_splitk_final(const GemmParams& p, const uint8_t* ring, uint8_t* staging, uint32_t rank,
                                                  uint32_t pair, int t2, int warp, int lane) {
    using Cfg = GemmCfg<BNC, 1, EW>;
    constexpr int UPC = BNC / 16, QPP = 4 / CKS, HBUF = Cfg::HBUF;
    const int ew = warp - 2;
    uint8_t* const pbuf0 = staging + ew * Cfg::WARP_STAGING;
    uint8_t* const hbuf0 = pbuf0 + Cfg::NBUF * 2048;
    const int TW = 1 << p.lw, TH = 1 << p.lh, lwh = p.lw + p.lh, TN = 128 >> lwh;
    const int sw = (lane >> 1) & 3, fo = lane * 64, ho = lane * 32, hsw = (lane >> 2) & 1;
    const bool prim_f32 = p.prim_f32 != 0, prim_store = p.prim_store != 0, sec_store = p.sec_store != 0;
    const bool silu = p.act == 1, relu = p.act == 3, out_f16 = p.out_f16 != 0, res_f32 = p.res_f32 != 0;
    const float scale = p.scale;
    const int n_tiles_n = p.n_tiles_n, tiles_w = p.tiles_w, tiles_h = p.tiles_h, lw = p.lw;
    const int pOW = p.OW, pOH = p.OH, pN = p.N;
    const long long osn = p.osn, osh = p.osh, osw = p.osw, bias_n_ld = p.bias_n_ld;
    const float* const bias = p.bias;
    const float* const bias_n = p.bias_n;
    const void* const res = p.res;
        const CUtensorMap* const hmap = &p.hmap;
    auto pack16 = [&](float lo, float hi) { return out_f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); };
    const int mp = t2 / n_tiles_n, n_tile_g = t2 - mp * n_tiles_n;
    const int par = p.n_par > 1 ? n_tile_g / p.tiles_per_par : 0;          // parity-split upsample: this tile's output view
    const int n_tile = n_tile_g - par * p.tiles_per_par;
    const CUtensorMap* const pmap = par ? &p.pmap_par[par - 1] : &p.pmap;
    const int m_tile = 2 * mp + (int)rank;
    const int twi = m_tile % tiles_w;
    const int rest = m_tile / tiles_w;
    const int thi = rest % tiles_h, tni = rest / tiles_h;
    pdl_wait();                                           // parameters are in registers; global memory from here on
    uint32_t A = 0;
#pragma unroll 1
    for (int item = ew; item < QPP * UPC; item += EW) {
        const int q_local = item / UPC, u = item - q_local * UPC;
        const int q = q_local * CKS + (int)pair;          // the row quarter (TMEM lane quarter of the senders) this CTA owns
        const int row0 = q * 32, row = row0 + lane;
        const int cw = twi * TW + (row0 & (TW - 1)), ch = thi * TH + ((row0 >> lw) & (TH - 1)), cn = tni * TN + (row0 >> lwh);
        const int ow = twi * TW + (row & (TW - 1)), oh = thi * TH + ((row >> lw) & (TH - 1)), n = tni * TN + (row >> lwh);
        const bool valid = ow < pOW && oh < pOH && n < pN;
        if (!__any_sync(0xffffffffu, valid)) continue;    // all-padding quarter (odd last tile, M < 256)
        const long long off = valid ? (long long)n * osn + (long long)oh * osh + (long long)ow * osw : 0;
        const float* const bn_row = bias_n ? bias_n + (long long)(n < pN ? n : pN - 1) * bias_n_ld : nullptr;
        const int gcol = n_tile * BNC + u * 16;
        uint8_t* const pb = pbuf0 + (A % 2) * 2048;
        uint8_t* const hb = hbuf0 + (A % HBUF) * 1024;
        if (lane == 0) {                                   // the store that last used these buffers has drained them
            if (HBUF >= 2 || !sec_store) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        __syncwarp();
        // partials of slices 0..CKS-1, added in slice order
        float4 sum[4];
        const float4* const src = reinterpret_cast<const float4*>(ring + (size_t)(q_local * UPC + u) * 2048) + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) sum[j] = src[j * 32];
#pragma unroll
        for (int s = 1; s < CKS; ++s) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = src[(size_t)s * QPP * UPC * 128 + j * 32];
                sum[j].x += v.x; sum[j].y += v.y; sum[j].z += v.z; sum[j].w += v.w;
            }
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 a = sum[j];
            float y0 = a.x * scale, y1 = a.y * scale, y2 = a.z * scale, y3 = a.w * scale;
            if (bias) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + gcol + 4 * j));
                y0 += bv.x; y1 += bv.y; y2 += bv.z; y3 += bv.w;
            }
            if (bn_row) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bn_row + gcol + 4 * j));
                y0 += bv.x; y1 += bv.y; y2 += bv.z; y3 += bv.w;
            }
            if (res && valid) {
                if (res_f32) {
                    const float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(res) + off + gcol + 4 * j);
                    y0 += r.x; y1 += r.y; y2 += r.z; y3 += r.w;
                } else {
                    const uint2 r2 = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(res) + off + gcol + 4 * j);
                    const float2 f0 = unpack_bf16x2(r2.x), f1 = unpack_bf16x2(r2.y);
                    y0 += f0.x; y1 += f0.y; y2 += f1.x; y3 += f1.y;
                }
            }
            if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
            if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
            if (prim_f32 && prim_store) *reinterpret_cast<float4*>(pb + fo + ((j ^ sw) << 4)) = make_float4(y0, y1, y2, y3);
            pk[2 * j] = pack16(y0, y1); pk[2 * j + 1] = pack16(y2, y3);
        }
        if (!prim_f32 || sec_store) {
            uint8_t* const bb = prim_f32 ? hb : pb;
            *reinterpret_cast<uint4*>(bb + ho + (hsw << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(bb + ho + ((hsw ^ 1) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (prim_store) tma_store_4d(pmap, pb, gcol, cw, ch, cn);
            if (sec_store) tma_store_4d(hmap, hb, gcol, cw, ch, cn);
            bulk_commit();
        }
        ++A;
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
}

template <int BNC, int NC, int EW, int CKS = 0>
__global__ void __cluster_dims__(CKS > 0 ? 2 * CKS : 2, 1, 1) __launch_bounds__((EW + 2) * 32, 1)
conv_gemm_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BNC, NC, EW>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment; both CTAs of the pair compute the same offset
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);   // leader's are the live ones
    uint64_t* empty_bar = full_bar + Cfg::STAGES;
    uint64_t* acc_full = empty_bar + Cfg::STAGES;
    uint64_t* acc_empty = acc_full + Cfg::SLOTS;                                       // leader's are the live ones
    uint64_t* res_bar = acc_empty + Cfg::SLOTS;                                        // [epilogue warp][NBUF]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(res_bar + EW * Cfg::NBUF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // the cluster is one CTA pair, or (CKS > 0, in-cluster split-K) CKS pairs: pair k = cluster ranks 2k, 2k + 1
    static_assert(CKS == 0 || ((CKS == 2 || CKS == 4) && NC == 1 && EW == 16), "in-cluster split-K: 2 or 4 slices of the 1 x BNC tile");
    constexpr int CSZ = CKS > 0 ? 2 * CKS : 2;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1, pair = crank >> 1, leader = crank & ~1u;
    const uint16_t pair_mask = (uint16_t)(3u << leader);
    const int cluster_id = blockIdx.x / CSZ, n_clusters = gridDim.x / CSZ;

    pdl_trigger();                      // the next kernel may be scheduled now (it waits for this one before touching memory)
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 5; ++i) tma_prefetch_desc(&p.amap[i]);
        tma_prefetch_desc(&p.bmap);
        if (p.epi_tma) { tma_prefetch_desc(&p.resmap); tma_prefetch_desc(&p.pmap); tma_prefetch_desc(&p.hmap); }
        if (p.n_par > 1) for (int i = 0; i < 3; ++i) tma_prefetch_desc(&p.pmap_par[i]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < Cfg::SLOTS; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * EW); }
        for (int s = 0; s < EW * Cfg::NBUF; ++s) mbar_init(&res_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_base_smem, Cfg::TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();                 // barriers of BOTH CTAs are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    // Everything above overlapped the previous kernel (programmatic dependent launch).  Each role calls pdl_wait() itself,
    // right before ITS first access to global memory: the producer before its first TMA load, the epilogue warps after
    // their parameter reads and geometry set-up; the MMA warp only touches shared and tensor memory and never waits.

    // work items: (pair of 128-pixel tiles, column tile, K split); the K split is the fastest index
    const int total_tiles = p.n_pairs_m * p.n_tiles_n * p.ksplit;
    const int TW = 1 << p.lw, TH = 1 << p.lh;
    // in-cluster split-K: exactly one work item per pair -- output tile = cluster, K slice = pair (p.ksplit == CKS)
    const int tile_first = CKS > 0 ? cluster_id * CKS + (int)pair : cluster_id;
    const int tile_step = CKS > 0 ? total_tiles : n_clusters;

    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs)
        if (lane == 0) {
            pdl_wait();
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
                const int ks = tile % p.ksplit, t2 = tile / p.ksplit;
                const int mp = t2 / p.n_tiles_n, n_tile = t2 - mp * p.n_tiles_n;
                const int m_tile = 2 * mp + (int)rank;
                const int twi = m_tile % p.tiles_w;
                const int rest = m_tile / p.tiles_w;
                const int thi = rest % p.tiles_h, tni = rest / p.tiles_h;
                const int par = p.n_par > 1 ? n_tile / p.tiles_per_par : 0;                  // output parity of this column tile
                const int w0 = twi * TW + (par & 1), h0 = thi * TH + (par >> 1), n0 = tni * (128 >> (p.lw + p.lh));   // n0 >= N: zero fill
                const int brow = n_tile * Cfg::N_TILE + (int)rank * (BNC / 2);
                const int kb0 = ks * p.kper, kb1 = kb0 + p.kper < p.total_kblk ? kb0 + p.kper : p.total_kblk;
                int kblk = 0;
                for (int it = 0; it < p.n_items; ++it) {
                    const GemmItem item = p.items[it];
                    if (kblk + item.nblk <= kb0) { kblk += item.nblk; continue; }        // before this split's range
                    for (int cb = kblk < kb0 ? kb0 - kblk : 0; cb < item.nblk && kblk + cb < kb1; ++cb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), leader);
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        tma_load_4d_pair(sa, &p.amap[item.map], fb, cb * 64, w0 + item.dw, h0 + item.dh, n0);
#pragma unroll
                        for (int c = 0; c < NC; ++c)
                            tma_load_2d_pair(sa + Cfg::A_BYTES + c * Cfg::B_CHUNK_BYTES, &p.bmap, fb, (kblk + cb) * 64,
                                             brow + c * BNC);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                    }
                    kblk += item.nblk;
                    if (kblk >= kb1) break;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only)
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, BNC, 0, 0);
            int stage = 0; uint32_t phase = 0;
            uint32_t cc = 0;                                   // accumulator chunks started so far (ring position)
            for (int tile = tile_first; tile < total_tiles; tile += tile_step, cc += NC) {
                uint32_t d_tmem[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const uint32_t slot = (cc + c) % Cfg::SLOTS, use = (cc + c) / Cfg::SLOTS;
                    mbar_wait(&acc_empty[slot], (use & 1) ^ 1);            // both CTAs have drained this slot
                    d_tmem[c] = tmem_base + slot * BNC;
                }
                tc_fence_after();
                const int ks = tile % p.ksplit;
                const int kb0 = ks * p.kper, kb1 = kb0 + p.kper < p.total_kblk ? kb0 + p.kper : p.total_kblk;
                for (int kb = 0; kb < kb1 - kb0; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t adesc = umma_desc_kmajor_sw128(sa);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const uint64_t bdesc = umma_desc_kmajor_sw128(sa + Cfg::A_BYTES + c * Cfg::B_CHUNK_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // advance 16 bf16 = 32 B along K inside the 128-B swizzle atom: +2 in the (addr >> 4) field
                            umma_bf16_pair(d_tmem[c], adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit_pair(&empty_bar[stage], pair_mask);   // frees the stage in both CTAs when these MMAs retire
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) umma_commit_pair(&acc_full[(cc + c) % Cfg::SLOTS], pair_mask);
            }
        }
    } else {
        // ===================================================================== epilogue (both CTAs, own 128 rows)
        if constexpr (CKS > 0) {
            mbar_wait(&acc_full[0], 0);                 // this pair's K slice is accumulated (send / final follow below)
            tc_fence_after();
        } else if (p.ksplit > 1) {
            if constexpr (NC == 1)
                epilogue_splitk<BNC, EW>(p, staging, acc_full, acc_empty, tmem_base, rank, cluster_id, n_clusters, warp, lane);
        } else if (p.epi_tma) {
#define RG_EPI_CASE(M) case (M): epilogue_tma<BNC, NC, EW, (M)>(p, staging, acc_full, acc_empty, res_bar, tmem_base, rank, cluster_id, n_clusters, warp, lane); break;
            switch (p.epi_mode) {
                RG_EPI_CASE(EPI_PRIM_STORE)                                                   // bf16 out
                RG_EPI_CASE(EPI_PRIM_STORE | EPI_F16)                                         // fp16 out (q|k|v, q)
                RG_EPI_CASE(EPI_PRIM_STORE | EPI_BIAS)                                        // bf16 out + bias
                RG_EPI_CASE(EPI_PRIM_STORE | EPI_BIAS | EPI_BIASN)                            // resnet conv1 (+ time embedding)
                RG_EPI_CASE(EPI_PRIM_STORE | EPI_BIAS | EPI_RES)                              // VAE conv2: bf16 residual
                RG_EPI_CASE(EPI_PRIM_F32 | EPI_PRIM_STORE | EPI_BIAS)                         // fp32 out (proj_in, up/down-sample)
                RG_EPI_CASE(EPI_PRIM_F32 | EPI_PRIM_STORE | EPI_SEC_STORE | EPI_BIAS)         // fp32 + bf16 out
                RG_EPI_CASE(EPI_PRIM_F32 | EPI_PRIM_STORE | EPI_BIAS | EPI_RES)               // fp32 stream += (attention out)
                RG_EPI_CASE(EPI_PRIM_F32 | EPI_PRIM_STORE | EPI_SEC_STORE | EPI_BIAS | EPI_RES)  // conv2 / proj_out with a bf16 copy
                RG_EPI_CASE(EPI_PRIM_F32 | EPI_SEC_STORE | EPI_BIAS | EPI_RES)                // feed-forward out: fp32 residual, bf16 out
                RG_EPI_CASE(EPI_GEGLU | EPI_PRIM_STORE | EPI_BIAS)                            // GEGLU
                default: epilogue_tma<BNC, NC, EW, EPI_GENERIC>(p, staging, acc_full, acc_empty, res_bar, tmem_base, rank, cluster_id, n_clusters, warp, lane); break;
            }
#undef RG_EPI_CASE
        } else {
            const int q = warp & 3;                    // TMEM lane quarter this warp may access
            const int half = (warp - 2) >> 2;          // units are dealt round-robin to the warps of a quarter
            constexpr int PARTS = Cfg::PARTS;
            const int rsub = lane >> 2, quad = lane & 3;
            float4* st0 = reinterpret_cast<float4*>(staging + (warp - 2) * Cfg::WARP_STAGING);
            float4* st1 = st0 + 128;
            const int wsw = (lane >> 1) & 3;           // write-phase swizzle of this lane's row
            const bool geglu = p.act == 2;
            const uint32_t acc_empty_leader = mapa_shared(smem_u32(acc_empty), 0);
            uint32_t cc = 0;
            pdl_wait();
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, cc += NC) {
                const int mp = tile / p.n_tiles_n, n_tile = tile - mp * p.n_tiles_n;
                const int m_tile = 2 * mp + (int)rank;
                const int twi = m_tile % p.tiles_w;
                const int rest = m_tile / p.tiles_w;
                const int thi = rest % p.tiles_h, tni = rest / p.tiles_h;
                // geometry of this lane's own row (TMEM lane q*32 + lane), then of the 4 rows it handles after the transpose
                RowGeom g;
                {
                    const int row = q * 32 + lane;
                    const int ow = twi * TW + (row & (TW - 1));
                    const int oh = thi * TH + ((row >> p.lw) & (TH - 1));
                    const int n = tni * (128 >> (p.lw + p.lh)) + (row >> (p.lw + p.lh));
                    const bool valid = (ow < p.OW) && (oh < p.OH) && (n < p.N);
                    const long long off = valid ? ((long long)n * p.osn + (long long)oh * p.osh + (long long)ow * p.osw) : 0;
                    g.valid = 0;
    #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int src = rsub + 8 * i;
                        g.off[i] = __shfl_sync(0xffffffffu, off, src);
                        g.n[i] = __shfl_sync(0xffffffffu, n, src);
                        g.valid |= (__shfl_sync(0xffffffffu, (int)valid, src) ? 1u : 0u) << i;
                    }
                }
    #pragma unroll 1
                for (int c = 0; c < NC; ++c) {
                    const uint32_t slot = (cc + c) % Cfg::SLOTS, use = (cc + c) / Cfg::SLOTS;
                    mbar_wait(&acc_full[slot], use & 1);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * BNC;
                    const int ncol0 = n_tile * Cfg::N_TILE + c * BNC;          // first weight row / GEMM column of this chunk
                    if (geglu) {
                        // unit = 16 value columns followed by their 16 gate columns (host interleave, weights.interleave_geglu)
    #pragma unroll 1
                        for (int u = half; u < BNC / 32; u += PARTS) {
                            if (ncol0 + u * 32 >= p.Cout) break;             // warp-uniform: the rest of the chunk is padding
                            uint32_t va[16], vg[16];
                            tmem_ld16(taddr + u * 32, va);
                            tmem_ld16(taddr + u * 32 + 16, vg);
                            tmem_ld_wait();
    #pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                st0[lane * 4 + (j ^ wsw)] = make_float4(__uint_as_float(va[4 * j]), __uint_as_float(va[4 * j + 1]),
                                                                        __uint_as_float(va[4 * j + 2]), __uint_as_float(va[4 * j + 3]));
                                st1[lane * 4 + (j ^ wsw)] = make_float4(__uint_as_float(vg[4 * j]), __uint_as_float(vg[4 * j + 1]),
                                                                        __uint_as_float(vg[4 * j + 2]), __uint_as_float(vg[4 * j + 3]));
                            }
                            __syncwarp();
                            const int ca = ncol0 + u * 32 + quad * 4, cg = ca + 16;
                            const int co = (ncol0 >> 1) + u * 16 + quad * 4;
                            float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
                            if (p.bias) {
                                ba = __ldg(reinterpret_cast<const float4*>(p.bias + ca));
                                bg = __ldg(reinterpret_cast<const float4*>(p.bias + cg));
                            }
    #pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int r = rsub + 8 * i;
                                const float4 a = st0[r * 4 + (quad ^ ((r >> 1) & 3))];
                                const float4 gt = st1[r * 4 + (quad ^ ((r >> 1) & 3))];
                                if (g.valid >> i & 1) {
                                    const float y0 = (a.x * p.scale + ba.x) * gelu_fast_f(gt.x * p.scale + bg.x);
                                    const float y1 = (a.y * p.scale + ba.y) * gelu_fast_f(gt.y * p.scale + bg.y);
                                    const float y2 = (a.z * p.scale + ba.z) * gelu_fast_f(gt.z * p.scale + bg.z);
                                    const float y3 = (a.w * p.scale + ba.w) * gelu_fast_f(gt.w * p.scale + bg.w);
                                    *reinterpret_cast<uint2*>(p.out_bf16 + g.off[i] + co) =
                                        (p.out_f16 ? make_uint2(pack_f16x2(y0, y1), pack_f16x2(y2, y3)) : make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3)));
                                }
                            }
                            __syncwarp();
                        }
                    } else {
                        int ub = 0;                              // staging buffer toggle: one __syncwarp per unit suffices
    #pragma unroll 1
                        for (int u = half; u < BNC / 16; u += PARTS, ub ^= 1) {
                            const int col = ncol0 + u * 16 + quad * 4;
                            if (ncol0 + u * 16 >= p.Cout) break;             // warp-uniform: the rest of the chunk is padding
                            uint32_t v[16];
                            tmem_ld16(taddr + u * 16, v);
                            tmem_ld_wait();
                            float4* st = ub ? st1 : st0;
    #pragma unroll
                            for (int j = 0; j < 4; ++j)
                                st[lane * 4 + (j ^ wsw)] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                            __syncwarp();
                            if (p.vec_ok && col + 4 <= p.Cout) {
                                float4 x[4], rr[4];
                                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                                // issue every global read of the unit before the first use
    #pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (g.valid >> i & 1) {
                                        if (p.res) {
                                            if (p.res_f32) {
                                                rr[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + g.off[i] + col);
                                            } else {
                                                const uint2 r2 = *reinterpret_cast<const uint2*>(
                                                    reinterpret_cast<const __nv_bfloat16*>(p.res) + g.off[i] + col);
                                                const float2 f0 = unpack_bf16x2(r2.x), f1 = unpack_bf16x2(r2.y);
                                                rr[i] = make_float4(f0.x, f0.y, f1.x, f1.y);
                                            }
                                        }
                                        if (p.bias_n) {
                                            const float4 bn = __ldg(reinterpret_cast<const float4*>(p.bias_n + (long long)g.n[i] * p.bias_n_ld + col));
                                            rr[i].x += bn.x; rr[i].y += bn.y; rr[i].z += bn.z; rr[i].w += bn.w;
                                        }
                                    }
                                }
    #pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int r = rsub + 8 * i;
                                    x[i] = st[r * 4 + (quad ^ ((r >> 1) & 3))];
                                }
    #pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    if (g.valid >> i & 1) {
                                        float y0 = x[i].x * p.scale + b.x + rr[i].x, y1 = x[i].y * p.scale + b.y + rr[i].y;
                                        float y2 = x[i].z * p.scale + b.z + rr[i].z, y3 = x[i].w * p.scale + b.w + rr[i].w;
                                        if (p.act == 1) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
                                        if (p.act == 3) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
                                        if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + g.off[i] + col) = make_float4(y0, y1, y2, y3);
                                        if (p.out_bf16)
                                            *reinterpret_cast<uint2*>(p.out_bf16 + g.off[i] + col) =
                                                (p.out_f16 ? make_uint2(pack_f16x2(y0, y1), pack_f16x2(y2, y3)) : make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3)));
                                    }
                                }
                            } else {
                                // ragged tail / unaligned output: scalar path
    #pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int r = rsub + 8 * i;
                                    const float4 xv = st[r * 4 + (quad ^ ((r >> 1) & 3))];
                                    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
                                    if (g.valid >> i & 1) {
    #pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            const int cidx = col + e;
                                            if (cidx < p.Cout) {
                                                float y = xs[e] * p.scale;
                                                if (p.bias) y += __ldg(p.bias + cidx);
                                                if (p.bias_n) y += __ldg(p.bias_n + (long long)g.n[i] * p.bias_n_ld + cidx);
                                                if (p.res) {
                                                    y += p.res_f32 ? reinterpret_cast<const float*>(p.res)[g.off[i] + cidx]
                                                                   : op2f(reinterpret_cast<const __nv_bfloat16*>(p.res)[g.off[i] + cidx]);
                                                }
                                                y = apply_act(y, p.act);
                                                if (p.out_f32) p.out_f32[g.off[i] + cidx] = y;
                                                if (p.out_bf16) { if (p.out_f16) reinterpret_cast<__half*>(p.out_bf16)[g.off[i] + cidx] = __float2half_rn(y); else p.out_bf16[g.off[i] + cidx] = f2op(y); }
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                    // this warp is done with the slot: one arrival per warp on the LEADER's barrier
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc_empty_leader + slot * 8);
                }
            }
        }
    }

    if constexpr (CKS > 0) {
        __syncwarp();
        cluster_sync_all();             // every pair's MMAs have retired: all operand rings of the cluster are free
        if (warp >= 2) cluster_splitk_send<BNC, EW, CKS>(smem, tmem_base, rank, pair, warp, lane);
        cluster_sync_all();             // the partials have landed in their owners' rings
        if (warp >= 2) cluster_splitk_final<BNC, EW, CKS>(p, smem, staging, rank, pair, cluster_id, warp, lane);
    }
    tc_fence_before();
    cluster_sync_all();                 // the peer may still be reading this CTA's shared memory / signalling its barriers
    if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
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

template <int BNC, int NC, int EW = 8, int CKS = 0>
static int launch_gemm(GemmParams& gp, const void* w, long long ktot, long long w_ld, cudaStream_t stream) {
    using Cfg = GemmCfg<BNC, NC, EW>;
    static_assert(CKS == 0 || 4 * (BNC / 16) * 2048 <= Cfg::STAGES * Cfg::STAGE_BYTES, "CKS partials of the owned row quarters must fit in the operand ring");
    static std::atomic<bool> attr_done[kMaxDevices];
    if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&conv_gemm_kernel<BNC, NC, EW, CKS>), Cfg::SMEM_BYTES, attr_done,
                                  "cudaFuncSetAttribute(conv_gemm_kernel)")) return rc;
    {
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)gp.Cout};
        cuuint64_t strides[1] = {(cuuint64_t)w_ld * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(BNC / 2)};
        cuuint32_t estr[2] = {1, 1};
        int rc = encode_tensor_map(&gp.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr,
                                   CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    gp.n_tiles_n = (gp.Cout + Cfg::N_TILE - 1) / Cfg::N_TILE;
    if (gp.Cout % Cfg::N_TILE != 0) gp.epi_tma = 0;        // the TMA epilogue assumes every 16-column unit is real
    gp.tiles_per_par = gp.n_tiles_n / (gp.n_par > 1 ? gp.n_par : 1);
    if (gp.n_par > 1 && (!gp.epi_tma || gp.n_tiles_n % gp.n_par != 0))
        return set_error(RG_ERR_ARG, "rg_conv2d: parities = 4 needs Cout to be a multiple of the column tile");
    const int total = gp.n_pairs_m * gp.n_tiles_n * gp.ksplit;
    const int max_clusters = sm_count() / 2;
    const int clusters = total < max_clusters ? total : max_clusters;
    // in-cluster split-K: one cluster of CKS pairs per output tile, not persistent (the hardware queues clusters that do
    // not fit at once; they are independent)
    const int ctas = CKS > 0 ? 2 * total : 2 * clusters;
    launch_kernel(conv_gemm_kernel<BNC, NC, EW, CKS>, dim3(ctas), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, gp);
    count_launch();
    return check_launch("conv_gemm_kernel");
}

// How many clusters of CKS pairs the device can hold at once (one CTA per SM at this shared-memory footprint; a cluster
// needs CKS whole TPCs inside one GPC).  The in-cluster split-K is only chosen when every output tile's cluster is
// resident in the first wave.
template <int CKS>
static int max_active_clusters() {
    using Cfg = GemmCfg<160, 1, 16>;
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    int n = cache[dev].load();
    if (n > 0) return n;
    static std::atomic<bool> attr_done[kMaxDevices];
    if (ensure_smem_attr(reinterpret_cast<const void*>(&conv_gemm_kernel<160, 1, 16, CKS>), Cfg::SMEM_BYTES, attr_done,
                         "cudaFuncSetAttribute(conv_gemm_kernel)")) return 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * CKS * 64); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2 * CKS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<160, 1, 16, CKS>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = CKS == 4 ? 14 : 30;          // conservative: B200 has 74 TPCs in 8 GPCs
    }
    cache[dev].store(n);
    return n;
}

}  // namespace rg

using namespace rg;

// Debug hook (not part of the public header): clusters of `cks` CTA pairs the current device holds at once.
extern "C" int rg_debug_max_active_clusters(int cks) { return cks == 4 ? max_active_clusters<4>() : max_active_clusters<2>(); }

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

    const bool parity4 = c->parities == 4;
    if (c->parities != 0 && c->parities != 1 && !parity4) return set_error(RG_ERR_ARG, "rg_conv2d: parities must be 0, 1 or 4");
    if (parity4 && (c->kh != 2 || c->kw != 2 || c->stride != 1 || c->has_x2 || c->res || c->bias_n || c->out_bf16 || !c->out_f32 ||
                    c->act != RG_ACT_NONE || c->Cout % 160 != 0))
        return set_error(RG_ERR_ARG, "rg_conv2d: parities = 4 is the 2x2 parity-split upsample: fp32 output, bias only, Cout % 160 == 0");
    GemmParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.n_par = parity4 ? 4 : 1;
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
    const int n_tiles_m = gp.tiles_w * gp.tiles_h * gp.tiles_n;
    gp.n_pairs_m = (n_tiles_m + 1) / 2;          // an odd last 128-pixel tile is paired with an all-padding one
    gp.N = N; gp.OH = OH; gp.OW = OW; gp.Cout = parity4 ? 4 * c->Cout : c->Cout;      // rows of the weight matrix

    const int cblk = (c->x.C + 63) / 64;
    int n_items = 0, n_maps = 0, rc;
    const char* xb = reinterpret_cast<const char*>(c->x.data);
    if (c->stride == 1) {
        rc = encode_act_map(&gp.amap[0], xb, c->x.C, c->x.W, c->x.H, N, c->x.stride_w, c->x.stride_h, c->x.stride_n,
                            TW, TH, TN);
        if (rc) return rc;
        n_maps = 1;
        for (int kh = 0; kh < c->kh; ++kh)
            for (int kw = 0; kw < c->kw; ++kw)
                gp.items[n_items++] = GemmItem{0, kw - (parity4 ? 1 : c->pad_l), kh - (parity4 ? 1 : c->pad_t), cblk};   // parity (py, px) adds (px, py) in the kernel
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

    const long long w_ld = c->w_ld ? c->w_ld : ktot;
    if (w_ld % 8 || w_ld < ktot) return set_error(RG_ERR_ARG, "rg_conv2d: w_ld must be a multiple of 8 and >= Ktot");
    if (reinterpret_cast<uintptr_t>(c->w) & 15) return set_error(RG_ERR_ARG, "rg_conv2d: weights must be 16-B aligned");

    gp.bias = c->bias; gp.bias_n = c->bias_n; gp.bias_n_ld = c->bias_n_ld;
    gp.res = c->res; gp.res_f32 = c->res_dtype == RG_DT_F32;
    gp.out_bf16 = reinterpret_cast<__nv_bfloat16*>(c->out_bf16);
    gp.out_f32 = c->out_f32;
    gp.osn = c->out_stride_n; gp.osh = c->out_stride_h; gp.osw = c->out_stride_w;
    gp.act = c->act; gp.scale = c->scale;
    gp.out_f16 = c->out16_dtype == RG_DT_F16;
    if (gp.out_f16 && (c->act == RG_ACT_GEGLU || (c->res && c->res_dtype == RG_DT_BF16)))
        return set_error(RG_ERR_ARG, "rg_conv2d: fp16 output is for plain projections (no GEGLU, no bf16 residual)");
    const bool aligned = (c->out_stride_n % 8 == 0) && (c->out_stride_h % 8 == 0) && (c->out_stride_w % 8 == 0) &&
                         !(reinterpret_cast<uintptr_t>(c->out_bf16) & 15) && !(reinterpret_cast<uintptr_t>(c->out_f32) & 15) &&
                         !(reinterpret_cast<uintptr_t>(c->res) & 15) && !(reinterpret_cast<uintptr_t>(c->bias) & 15) &&
                         !(reinterpret_cast<uintptr_t>(c->bias_n) & 15) && (c->bias_n_ld % 4 == 0) && (c->Cout % 4 == 0);
    gp.vec_ok = aligned ? 1 : 0;

    // ---- TMA epilogue: per-warp boxes of {16 columns, 32 pixels}; needs 16-byte aligned row segments and a residual
    // of the same dtype as the buffer it is combined in (fp32 residual -> fp32 primary buffer, else the output dtype)
    gp.out_cols = c->act == RG_ACT_GEGLU ? c->Cout / 2 : c->Cout;
#ifdef RG_GEMM_TUNING            /* perf experiments only (RG_NVCC_EXTRA=-DRG_GEMM_TUNING): never in the product build */
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("RG_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
        gp.dbg = dbg;
    }
#endif
    {
        const bool res_f32 = c->res && c->res_dtype == RG_DT_F32, res_b16 = c->res && c->res_dtype == RG_DT_BF16;
        const bool prim_f32 = c->res ? res_f32 : (c->out_f32 != nullptr);
        const bool ok = aligned && !(res_b16 && c->out_f32) && gp.out_cols % 4 == 0 && (c->out_stride_w % 8 == 0);
        if (ok) {
            gp.epi_tma = 1;
            gp.prim_f32 = prim_f32;
            gp.prim_store = prim_f32 ? (c->out_f32 != nullptr) : 1;
            gp.sec_store = prim_f32 && c->out_bf16 != nullptr;
            const int bw = TW < 32 ? TW : 32;
            const int bh = TH < 32 / bw ? TH : 32 / bw;
            const int bn = 32 / (bw * bh);
            auto enc = [&](CUtensorMap* m, const void* base, bool f32, int par = 0) -> int {
                const long long esz = f32 ? 4 : 2;
                // extent-1 dimensions may come with a zero / arbitrary pitch: give them a sane one
                long long sw = c->out_stride_w, sh = c->out_stride_h, sn = c->out_stride_n;
                if (parity4) {                       // view of output parity (py, px): every second row / column of the full tensor
                    base = reinterpret_cast<const char*>(base) + ((par >> 1) * sh + (par & 1) * sw) * esz;
                    sw *= 2; sh *= 2;
                }
                if (OW == 1 && sw < gp.out_cols) sw = (gp.out_cols + 7) / 8 * 8;
                if (OH == 1 && sh < sw * OW) sh = sw * OW;
                if (N == 1 && sn < sh * OH) sn = sh * OH;
                cuuint64_t dims[4] = {(cuuint64_t)gp.out_cols, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
                cuuint64_t strides[3] = {(cuuint64_t)(sw * esz), (cuuint64_t)(sh * esz), (cuuint64_t)(sn * esz)};
                cuuint32_t box[4] = {16, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
                cuuint32_t estr[4] = {1, 1, 1, 1};
                return encode_tensor_map(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base,
                                         dims, strides, box, estr, f32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
            };
            const void* pout = prim_f32 ? (const void*)c->out_f32 : (const void*)c->out_bf16;
            if (gp.prim_store && (rc = enc(&gp.pmap, pout, prim_f32))) return rc;
            if (parity4)
                for (int par = 1; par < 4; ++par)
                    if ((rc = enc(&gp.pmap_par[par - 1], pout, prim_f32, par))) return rc;
            if (c->res && (rc = enc(&gp.resmap, c->res, prim_f32))) return rc;
            if (gp.sec_store && (rc = enc(&gp.hmap, c->out_bf16, false))) return rc;
            if (!gp.prim_store) gp.pmap = c->res ? gp.resmap : gp.hmap;
            if (!c->res) gp.resmap = gp.pmap;
            if (!gp.sec_store) gp.hmap = gp.pmap;
            int mode = 0;
            if (c->act == RG_ACT_GEGLU) mode |= 1;       // EPI_GEGLU
            if (gp.prim_f32) mode |= 2;
            if (c->res) mode |= 4;
            if (gp.prim_store) mode |= 8;
            if (gp.sec_store) mode |= 16;
            if (c->bias) mode |= 32;
            if (c->bias_n) mode |= 64;
            if (c->act == RG_ACT_SILU) mode |= 128;
            if (c->act == RG_ACT_RELU) mode |= 1 << 20;  // no specialised unit code: run-time-flag epilogue
            if (c->scale != 1.0f) mode |= 256;
            if (c->out16_dtype == RG_DT_F16) mode |= 512;
            gp.epi_mode = (gp.dbg & 7) ? (1 << 20) : mode;    // debug experiments run the generic (run-time flag) epilogue
        }
    }

    // ---- tile shape.  Columns: one or two chunks of BNC.  The 2 x 160 tile halves the L2 traffic per FLOP but
    // quantises the grid more coarsely; estimate both (cycles per k-block: max(MMA, L2 fill at ~43 B/clk/SM)).
    const int Cout = gp.Cout;                    // weight rows (4 x Cout for the parity-split upsample)
    if (c->act == RG_ACT_GEGLU) {
        if (Cout % 32 != 0 || !c->out_bf16 || c->out_f32 || c->res || c->bias_n || !aligned)
            return set_error(RG_ERR_ARG, "rg_conv2d: GEGLU needs Cout % 32 == 0 and a 16-B aligned bf16 output only");
    }
    gp.ksplit = 1;
    gp.kper = gp.total_kblk;
    const int clusters = sm_count() / 2;
    // ---- deterministic split-K (epilogue_splitk) when the layer has too few output tiles to fill the SMs and a long K:
    // as many slices (8, 4 or 2) as it takes to give every cluster a work item, at least 12 k-blocks per slice, and only
    // for main loops of >= 48 k-blocks (the dump + fix-up costs about as much as 30 k-blocks of MMA).  At UNet batch 2
    // that is every conv below the 64x64 level and the long-K feed-forward projections; at batch 16 only the 8x8 level.
    // The slice count depends on the batch size, so results are reproducible per batch size, not across batch sizes.
    if (gp.epi_tma && Cout % 160 == 0 && c->act != RG_ACT_GEGLU && c->splitk_ws && (gp.dbg & 7) == 0) {
        const long long tiles160 = (long long)gp.n_pairs_m * (Cout / 160);
        int ks = 1;
        if (gp.total_kblk >= 48 && tiles160 * 2 <= clusters) {
            ks = 8;
            while (ks > 1 && (ks * tiles160 > clusters || ks > gp.total_kblk / 12)) ks >>= 1;
        }
        // ---- first choice: the K slices of a tile inside ONE cluster (4 or 2 CTA pairs), partials exchanged through
        // shared memory (cluster_splitk_send / _final) -- when every tile's cluster fits on the device at once
        if (ks > 1) {
            // Measured (profiles/r02_cluster_splitk.txt): the main loop of these layers runs at ~0.29 us per k-block whichever
            // way the K slices meet (operand fetch, not MMA), the in-cluster fix-up saves 3.5-8 us per launch, and a launch
            // that engages half as many pairs loses that again after ~45 k-blocks.  So: in-cluster when it engages as many
            // pairs as the workspace version would, or when its slices are short anyway.
            int cks = 0;
            if (tiles160 * 4 <= clusters && tiles160 <= max_active_clusters<4>()) cks = 4;
            else if (tiles160 <= max_active_clusters<2>()) cks = 2;
            if (cks && cks < ks && (gp.total_kblk + cks - 1) / cks > 48) cks = 0;
#ifdef RG_GEMM_TUNING
            { const char* e = getenv("RG_GEMM_CLUSTER"); if (e) { const int v = atoi(e); if (v == 0) cks = 0; else if (v == 2 && cks == 4) cks = 2; } }
#endif
            const int kper = cks ? (gp.total_kblk + cks - 1) / cks : 0;
            if (cks && (long long)kper * (cks - 1) < gp.total_kblk) {
                gp.ksplit = cks;
                gp.kper = kper;
                return cks == 4 ? launch_gemm<160, 1, 16, 4>(gp, c->w, ktot, w_ld, stream)
                                : launch_gemm<160, 1, 16, 2>(gp, c->w, ktot, w_ld, stream);
            }
        }
        if (ks > 1 && gp.n_par == 1) {
            // 16 epilogue warps: 2-3 units per warp instead of 5 -- the fix-up of a tile is a serial chain of units in the
            // last-arriving warp (one L2 round trip each), so its depth is what the launch waits for
            using Cfg = GemmCfg<160, 1, 16>;
            constexpr long long kRegion = (long long)((Cfg::N_TILE / 16 + Cfg::PARTS - 1) / Cfg::PARTS) * 2048;   // bytes per (tile, slice, rank, warp)
            const long long out_tiles = (long long)gp.n_pairs_m * (Cout / 160);
            const long long cnt_bytes = out_tiles * 2 * 16 * (long long)sizeof(int);
            const long long part_bytes = out_tiles * ks * 2 * 16 * kRegion;
            if ((reinterpret_cast<uintptr_t>(c->splitk_ws) & 15) == 0 && cnt_bytes <= RG_SPLITK_COUNTER_BYTES &&
                RG_SPLITK_COUNTER_BYTES + part_bytes <= c->splitk_ws_bytes) {
                gp.ksplit = ks;
                gp.kper = (gp.total_kblk + ks - 1) / ks;
                if ((long long)gp.kper * (ks - 1) >= gp.total_kblk) {               // an empty last slice would never commit
                    gp.ksplit = 1; gp.kper = gp.total_kblk;
                } else {
                    gp.ws_cnt = reinterpret_cast<int*>(c->splitk_ws);
                    gp.ws_part = reinterpret_cast<float*>(reinterpret_cast<char*>(c->splitk_ws) + RG_SPLITK_COUNTER_BYTES);
                    return launch_gemm<160, 1, 16>(gp, c->w, ktot, w_ld, stream);
                }
            }
        }
    }
    auto waves = [&](int n_tile) { return (int)(((long long)gp.n_pairs_m * ((Cout + n_tile - 1) / n_tile) + clusters - 1) / clusters); };
    if (Cout % 160 == 0) {
        // short K: the epilogue dominates and wants the full double buffering of the 1 x 160 tile (3 TMEM slots)
        int min_kblk = 24, ew16 = 24;
#ifdef RG_GEMM_TUNING
        { const char* e = getenv("RG_GEMM_NC2_MIN_KBLK"); if (e) min_kblk = atoi(e); }
        { const char* e = getenv("RG_GEMM_EW16_MAX_KBLK"); if (e) ew16 = atoi(e); }
#endif
        if (Cout % 320 == 0 && (!parity4 || c->Cout % 320 == 0) && gp.total_kblk > min_kblk &&
            (long long)waves(320) * 865 <= (long long)waves(160) * 625)
            return launch_gemm<160, 2>(gp, c->w, ktot, w_ld, stream);
        if (gp.epi_tma && gp.total_kblk <= ew16) return launch_gemm<160, 1, 16>(gp, c->w, ktot, w_ld, stream);
        return launch_gemm<160, 1>(gp, c->w, ktot, w_ld, stream);
    }
    if (Cout > 128) return launch_gemm<256, 1>(gp, c->w, ktot, w_ld, stream);
    if (Cout > 64) return launch_gemm<128, 1>(gp, c->w, ktot, w_ld, stream);
    if (Cout > 32) return launch_gemm<64, 1>(gp, c->w, ktot, w_ld, stream);
    return launch_gemm<32, 1>(gp, c->w, ktot, w_ld, stream);
}
