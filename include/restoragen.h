/*
 * restoragen.h -- C ABI of librestoragen.so, the B200 (sm_100a) kernel library behind the
 * RestoraGen Stable-Diffusion sampling loop.
 *
 * Boundary.  The reference (qmoututu11/Image_Restoration_and_Enhancement) has no FFI layer of its
 * own: its hot path is the pair of diffusers pipeline calls made at
 *     src/inference.py:486-494   (denoise,  StableDiffusionImg2ImgPipeline.__call__)
 *     src/inference.py:566-573   (sr,       StableDiffusionImg2ImgPipeline.__call__)
 *     src/inference.py:664-672   (colorize, StableDiffusionImg2ImgPipeline.__call__)
 *     src/inference.py:758-767   (inpaint,  StableDiffusionInpaintPipeline.__call__)
 * and everything below those calls is torch -> cuDNN / cuBLASLt / SDPA library kernels.  This header
 * declares the operator set that replaces that library layer.  The
 * Python host side (image_restoration_and_enhancement_b200/pipelines.py) re-provides the two pipeline
 * classes on top of it; INTEGRATION.md shows the ctypes binding.
 *
 * Conventions.
 *   - every entry point returns 0 on success, a cudaError_t (> 0) for CUDA failures, or a negative
 *     RG_ERR_* for argument errors; nothing throws; rg_last_error() returns a message for the
 *     calling thread;
 *   - all data pointers are DEVICE pointers; activations are channels-last (N,H,W,C), C contiguous;
 *   - all work is enqueued on the caller's stream and nothing allocates, so every call can be
 *     captured into a CUDA graph;
 *   - bf16 tensors are passed as void*, fp32 as float*.
 */
#ifndef RESTORAGEN_H_
#define RESTORAGEN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rg_stream_t; /* cudaStream_t */

#define RG_OK 0
#define RG_ERR_ARG (-1)         /* invalid argument / unsupported shape */
#define RG_ERR_DRIVER (-2)      /* driver entry point (cuTensorMapEncodeTiled) unavailable */
#define RG_ERR_TENSORMAP (-3)   /* tensor-map encode failed */

#define RG_ACT_NONE 0
#define RG_ACT_SILU 1
#define RG_ACT_GEGLU 2          /* columns interleaved 16 value | 16 gate per 32-wide unit; output width = Cout/2 */
#define RG_ACT_RELU 3           /* the AlexNet feature convolutions of LPIPS (src/metrics.py:97-111) */

#define RG_DT_BF16 0
#define RG_DT_F32 1
#define RG_DT_F16 2

const char* rg_last_error(void);
int rg_version(void);
/* Encoding of every 16-bit activation / weight this build reads and writes (the buffers the structs below call "bf16"):
   RG_DT_BF16 for librestoragen.so (the product), RG_DT_F16 for librestoragen_f16.so -- the same sources compiled with
   -DRG_OPERAND_F16, the "fp16 parity mode": fp16 storage and fp16 tensor-core operands end to end, the dtype the
   reference runs on CUDA (src/inference.py:57 and :162-166, torch_dtype=torch.float16).  A process uses one of the two. */
int rg_operand_dtype(void);
/* number of kernel launches issued by this library in the calling process (all threads) */
int64_t rg_launch_count(void);
int rg_device_sm_count(void);
/* Programmatic dependent launch between consecutive kernels of this library (default on): the next kernel's launch
   latency and prologue overlap the running kernel; every kernel waits for its predecessor before it touches global
   memory, so results are unchanged.  mode: bit 0 = GEMM / attention / glue kernels, bit 1 = GroupNorm kernels; 0 = plain
   stream serialization.  Returns the previous mode.  Process-wide; set before capturing CUDA graphs. */
int rg_set_pdl(int mode);

/* ---------------------------------------------------------------------------------------------
 * K1-K4  implicit-GEMM convolution / linear layer on tcgen05 (TMEM accumulators, TMA operand loads).
 *   replaces nn.Conv2d 3x3 / 1x1 and nn.Linear inside UNet2DConditionModel / AutoencoderKL
 *   (shapes Appendix C).
 *
 *   out[m][co] = act( scale * sum_k A[m][k] * W[co][k] + bias[co] + bias_n[n(m)][co] + res[m][co] )
 *   where m = (n, oh, ow) and A is the im2col view of x (taps outer, channels inner), optionally
 *   followed in K by the channels of x2 sampled at the output pixel (a fused 1x1 shortcut).
 * ------------------------------------------------------------------------------------------- */
typedef struct rg_act {
    const void* data;                  /* bf16, element (n=0,h=0,w=0,c=0) */
    int32_t N, H, W, C;                /* C: multiple of 64 (any size for a plain 1x1 / linear layer) */
    int64_t stride_n, stride_h, stride_w; /* in elements; multiples of 8 */
} rg_act_t;

typedef struct rg_conv {
    rg_act_t x;
    int32_t kh, kw;        /* 1..3 each (2x2 is used by the parity-split upsample convolution) */
    int32_t stride;        /* 1 or 2 */
    int32_t pad_t, pad_l;  /* top/left zero padding; bottom/right padding is implied by OH/OW */
    int32_t OH, OW;
    int32_t has_x2;
    rg_act_t x2;           /* same N, spatial dims OH x OW */
    const void* w;         /* bf16 [Cout][Ktot], Ktot = kh*kw*x.C + x2.C */
    int64_t w_ld;          /* row pitch of w in elements (0 = Ktot); multiple of 8 */
    int32_t Cout;
    const float* bias;     /* [Cout] or NULL */
    const float* bias_n;   /* per-image bias [N][bias_n_ld] (time-embedding projection) or NULL */
    int64_t bias_n_ld;     /* row pitch of bias_n in elements (multiple of 4) */
    const void* res;       /* residual, same addressing as the outputs, or NULL */
    int32_t res_dtype;     /* RG_DT_* */
    void* out_bf16;        /* either or both outputs */
    float* out_f32;
    /* address of output pixel (n,oh,ow), in elements, for res and both outputs:
       n*out_stride_n + oh*out_stride_h + ow*out_stride_w  (strided writes serve the upsample split) */
    int64_t out_stride_n, out_stride_h, out_stride_w;
    int32_t act;           /* RG_ACT_* */
    float scale;
    int32_t out16_dtype;   /* element type written through out_bf16: RG_DT_BF16 (default) or RG_DT_F16 (the attention
                              operands q, k, v: fp16 like the reference's CUDA path, src/inference.py:57) */
    /* Optional split-K workspace (NULL = never split).  A layer with too few output tiles to fill the SMs and a long K
       (everything below the 64x64 level of the UNet at batch 1-2) has its K range cut into 2, 4 or 8 slices computed by
       different CTA pairs; every slice writes its fp32 partial tile here and the LAST slice to arrive (a per-tile arrival
       counter at the head of the workspace elects it) adds the partials IN SLICE ORDER and runs the normal epilogue, so a
       launch is bitwise reproducible.  The slice count follows the number of output tiles, i.e. the batch size: results
       are reproducible per batch size.  The workspace must be ZERO when first used (the counters reset themselves), at
       least RG_SPLITK_WS_MIN_BYTES, and is shared by all calls of one stream. */
    void* splitk_ws;
    int64_t splitk_ws_bytes;
    /* 0 / 1: an ordinary convolution.  4: the parity-split form of "nearest-2x upsample + 3x3 convolution" in ONE launch
       (kh = kw = 2, stride 1, fp32 output only, no residual): w = bf16 [4][Cout][4 * x.C], the four 2x2 kernels of output
       parities (py, px) = (0,0), (0,1), (1,0), (1,1) stacked along the rows; parity (py, px) reads input rows
       {j-1+py, j+py} and columns {i-1+px, i+px} (pad_t / pad_l are ignored) and writes output pixel (2j+py, 2i+px);
       OH x OW is the INPUT-resolution grid and out_f32 / out_stride_* address the full [N][2 OH][2 OW][Cout] tensor. */
    int32_t parities;
} rg_conv_t;

#define RG_SPLITK_COUNTER_BYTES (256 * 1024)
#define RG_SPLITK_WS_MIN_BYTES (RG_SPLITK_COUNTER_BYTES + (32ll << 20))

int rg_conv2d(const rg_conv_t* p, rg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K5/K6  fused flash-style attention on tcgen05: O = softmax(scale * Q K^T) V, no mask or a causal mask.
 *   replaces F.scaled_dot_product_attention in attn1 (self, N = H*W tokens) and attn2
 *   (cross, 77 CLIP tokens) of every BasicTransformerBlock.
 *   q/k/v are bf16 or fp16, out is bf16: [B][tokens][heads][d] views with arbitrary (multiple-of-8) strides.
 *   The fp16 path (d in {40, 80, 160}) evaluates two exponentials per MUFU instruction (ex2.approx.f16x2) and takes
 *   the softmax denominator from a ones column appended to V inside the P V GEMM.
 * ------------------------------------------------------------------------------------------- */
typedef struct rg_attn {
    const void *q, *k, *v;
    void* out;
    int32_t B, heads, d;           /* d in {40, 80, 160} (any multiple of 8 up to 192) */
    int32_t Nq, Nk;
    int64_t q_stride_b, q_stride_t, q_stride_h;   /* elements */
    int64_t k_stride_b, k_stride_t, k_stride_h;
    int64_t v_stride_b, v_stride_t, v_stride_h;
    int64_t o_stride_b, o_stride_t, o_stride_h;
    float scale;
    int32_t dtype;                 /* element type of q, k, v: RG_DT_BF16 (default) or RG_DT_F16; out is always bf16 */
    int32_t causal;                /* 1: query i attends keys 0..i only (CLIPTextModel's causal mask); bf16, d <= 64, Nq == Nk */
} rg_attn_t;

int rg_attention(const rg_attn_t* p, rg_stream_t stream);

/* row softmax in place on a bf16 [rows][cols] matrix (VAE mid-block attention, materialised scores) */
int rg_softmax_rows(void* x, int64_t rows, int32_t cols, int64_t ld, rg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K7  GroupNorm (+SiLU), channels-last.  x may be the channel-concatenation of two tensors
 *   (skip connections of the up blocks) -- the concat is never materialised in fp32.
 *   replaces nn.GroupNorm + F.silu.
 *   rg_groupnorm_stats writes per-block (sum, sumsq) partials per (n, group) into the workspace and the last block
 *   of each image combines them in block order in fp64 into (mean, rstd); the split depends on (HW, C) only and the
 *   reduction order is fixed, so results are bitwise reproducible and independent of the batch size;
 *   rg_groupnorm_apply writes y = silu?((x-mean)*rstd*gamma+beta) as bf16 and optionally the raw
 *   concatenated input as bf16 (feeds the fused 1x1 shortcut).
 *   The first RG_GN_MAX_IMAGES words of the workspace are block-arrival counters: they must be ZERO when
 *   rg_groupnorm_stats is enqueued and are reset by the kernel, so a workspace zeroed once can be reused
 *   by every later call on the same stream.
 * ------------------------------------------------------------------------------------------- */
#define RG_GN_MAX_BLOCKS 256
#define RG_GN_MAX_IMAGES 1024   /* the counter area has a FIXED size so that one workspace serves any batch size */
#define RG_GN_WORKSPACE_FLOATS(N, G) (RG_GN_MAX_IMAGES + (N) * (G) * 2 + (N) * RG_GN_MAX_BLOCKS * (G) * 2)

typedef struct rg_gn {
    const void* x1; int32_t C1;     /* first source  [N][HW][C1] */
    const void* x2; int32_t C2;     /* second source [N][HW][C2] or NULL/0 */
    int32_t in_dtype;               /* RG_DT_* (both sources) */
    int32_t N; int64_t HW;
    int32_t groups; float eps;
    const float* gamma; const float* beta;   /* [C1+C2] */
    float* sums;                    /* workspace of RG_GN_WORKSPACE_FLOATS(N, groups) floats (counters | mean,rstd | partials) */
    void* y;                        /* bf16 [N][HW][C1+C2] */
    void* raw;                      /* bf16 copy of the concatenated input, or NULL */
    int32_t silu;
} rg_gn_t;

int rg_groupnorm_stats(const rg_gn_t* p, rg_stream_t stream);
int rg_groupnorm_apply(const rg_gn_t* p, rg_stream_t stream);
/* both in one call.  When one (image, group) slice -- HW x C/groups elements -- fits in 96 KB of shared memory (the
   32x32-and-below levels of the UNet) a single one-pass kernel reads the input once; otherwise stats + apply.  The
   choice depends on (HW, C, groups, dtype) only, never on N, so results stay independent of the batch size. */
int rg_groupnorm(const rg_gn_t* p, rg_stream_t stream);

/* K8  LayerNorm over the last dim, affine, eps; x [rows][C] f32|bf16 -> y bf16 */
int rg_layernorm(const void* x, int32_t in_dtype, int64_t rows, int32_t C, const float* gamma,
                 const float* beta, float eps, void* y, rg_stream_t stream);

/* K10  sinusoidal timestep embedding (flip_sin_to_cos, freq_shift 0): out bf16 [B][dim] */
int rg_timestep_embedding(const float* timesteps, int32_t B, int32_t dim, void* out, rg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K11  classifier-free-guidance combine fused with the scheduler update (PLMS or DDIM eta=0).
 *   replaces `eps = u + g*(c-u)` plus PNDMScheduler.step_plms / DDIMScheduler.step
 *   (SURVEY.md Appendix A.3a/A.3b).  All state is fp32, layout-agnostic flat arrays of n elements
 *   per image batch half:  eps_uc holds [2][n] (uncond, cond) when do_cfg else [1][n].
 *     e      = do_cfg ? u + g*(c-u) : eps
 *     if store_slot >= 0: ets[store_slot] = e
 *     e_mix  = sum_i w[i] * (i == 4 ? e : ets[i])            (w[0..3] history slots, w[4] current)
 *     base   = use_cur ? cur_sample : sample
 *     if save_cur: cur_sample = sample
 *     sample_out = c_sample * base - c_eps * e_mix
 *   The host computes (w, c_sample, c_eps) in the scheduler's own float32 arithmetic.
 * ------------------------------------------------------------------------------------------- */
typedef struct rg_sched {
    const float* eps_uc;
    float* sample;          /* in/out [n] */
    float* ets;             /* [4][n] history ring (PLMS) or NULL */
    float* cur_sample;      /* [n] or NULL */
    int64_t n;
    int32_t do_cfg; float guidance;
    int32_t store_slot;     /* -1: do not store */
    float w[5];
    int32_t use_cur, save_cur;
    float c_sample, c_eps;
} rg_sched_t;

int rg_sched_step(const rg_sched_t* p, rg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K9/K14  layout / glue kernels
 * ------------------------------------------------------------------------------------------- */
/* im2col for convolutions with very few input channels (conv_in of UNet/VAE, post_quant):
   x f32|bf16 channels-last [n_mod][H][W][Cin] -> bf16 [N*OH*OW][Kpad], K = (tap, c) order, zero padded;
   output image n reads input image n % n_mod (classifier-free guidance feeds the same latents twice) */
int rg_im2col_small(const void* x, int32_t in_dtype, int32_t N, int32_t n_mod, int32_t H, int32_t W,
                    int32_t Cin, int32_t ksize, int32_t stride, int32_t pad, int32_t OH, int32_t OW,
                    int32_t Kpad, void* out, rg_stream_t stream);
/* nearest-neighbour resize (F.interpolate mode="nearest"), bf16 channels-last [N][H][W][C] -> [N][OH][OW][C];
   only used when an upsampler's target is not exactly 2x (UNet inputs not divisible by 8) -- the exact-2x case is
   folded into the convolution as four parity convolutions */
int rg_upsample_nearest(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t OH, int32_t OW,
                        void* y, rg_stream_t stream);
/* f32 NCHW <-> f32 NHWC (latents and RNG draws arrive in torch's NCHW order) */
int rg_nchw_to_nhwc(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, float* y, rg_stream_t stream);
int rg_nhwc_to_nchw(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, float* y, rg_stream_t stream);
/* VaeImageProcessor.preprocess tail: u8 HWC [N][H][W][3] -> f32 NHWC in [-1,1] (x/255*2-1),
   optionally multiplied by (mask < 0.5) with mask f32 [N][H][W] (inpaint masked_image) */
int rg_preprocess_u8(const uint8_t* img, const float* mask, int32_t N, int32_t H, int32_t W, float* out,
                     rg_stream_t stream);
/* VaeImageProcessor.postprocess: f32 NHWC [N][H][W][ldc>=3] -> u8 HWC, round(clamp(x/2+.5,0,1)*255) */
int rg_postprocess_u8(const float* x, int32_t N, int32_t H, int32_t W, int32_t ldc, uint8_t* out,
                      rg_stream_t stream);
/* DiagonalGaussianDistribution.sample * scaling fused with scheduler.add_noise:
   moments f32 NHWC [N][HW][8] (mean 0-3, logvar 4-7), eps_post / noise f32 NHWC [N][HW][4]
   z = (mean + exp(0.5*clamp(logvar,-30,20))*eps_post) * scaling
   latents = add_noise ? sqrt_ac*z + sqrt_1mac*noise : z */
int rg_vae_sample(const float* moments, int64_t moments_ld, const float* eps_post, const float* noise,
                  int64_t npix, float scaling, int32_t add_noise, float sqrt_ac, float sqrt_1mac,
                  float* out, rg_stream_t stream);
/* inpaint UNet input assembly (StableDiffusionInpaintPipeline: cat([latents, mask, masked_image_latents], 1)):
   f32 [npix][9] <- latents f32 [npix][4], mask f32 [npix], masked latents f32 [npix][4] */
int rg_pack_unet_input(const float* latents, const float* mask, const float* masked, int64_t npix, float* out,
                       rg_stream_t stream);
/* 1x1 convolution with very few channels on fp32 pixels (quant_conv 8->8, post_quant_conv 4->4):
   out[p][co] = b[co] + sum_c W[co][c] * (x[p][c] * scale_in) */
int rg_pointwise_small(const float* x, int64_t npix, int32_t Cin, int32_t Cout, const float* W, const float* b,
                       float scale_in, float* out, rg_stream_t stream);
/* nearest resize of the binarised mask f32 [N][H][W] -> [N][h][w] (F.interpolate default) */
int rg_mask_nearest(const float* mask, int32_t N, int32_t H, int32_t W, int32_t h, int32_t w, float* out,
                    rg_stream_t stream);
/* elementwise helpers */
int rg_scale_f32(const float* x, float a, int64_t n, float* y, rg_stream_t stream);
int rg_cast_f32_bf16(const float* x, int64_t n, void* y, rg_stream_t stream);
int rg_memset_zero(void* p, int64_t bytes, rg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K16  CLIP text encoder glue (SURVEY 8f "f3"; transformers CLIPTextModel behind pipe.text_encoder,
 *      /root/reference/outputs/models/denoising/best/text_encoder/config.json): the 12 layers run on rg_layernorm,
 *      rg_conv2d (as linear) and rg_attention with causal = 1; these two cover what is left.
 *   rg_embed_tokens: out f32 [B*T][C] = token_embedding[ids[b][t]] + position_embedding[t]  (CLIPTextEmbeddings)
 *   rg_quick_gelu_bf16: x <- x * sigmoid(1.702 x) in place on n bf16 values (hidden_act "quick_gelu")
 * ------------------------------------------------------------------------------------------- */
int rg_embed_tokens(const int32_t* ids, const float* token_embedding, const float* position_embedding, int32_t B,
                    int32_t T, int32_t C, int32_t vocab, float* out, rg_stream_t stream);
int rg_quick_gelu_bf16(void* x, int64_t n, rg_stream_t stream);
int rg_cast_bf16_f32(const void* x, int64_t n, float* y, rg_stream_t stream);   /* last_hidden_state handed out as fp32 */

/* ---------------------------------------------------------------------------------------------
 * K15  per-image PSNR / SSIM on u8 images in HBM (SURVEY 8f "f2"), bit-exact against the reference's
 *      float64 CPU bookkeeping.  Replaces MetricsCalculator.calculate_psnr / calculate_ssim
 *      (/root/reference/src/metrics.py:82-96 -> skimage.metrics.peak_signal_noise_ratio /
 *      structural_similarity with data_range=255, channel_axis=2).
 *   pred, gt: u8 [N][H][W][C] (channels last, contiguous).
 *   rg_metrics_sse_u8: sse[n] = sum over the image of (pred-gt)^2, exact (uint64).  PSNR = 10 log10(255^2 /
 *     (sse / elems)) is finished by the caller in float64.
 *   rg_metrics_ssim_u8: writes the cropped SSIM map (float64 [N][C][H-6][W-6]) into smap_ws and, per (n, c),
 *     rg_metrics_ssim_chunks(H, W) partial sums into chunk_sums [N][C][chunks]; the caller adds them left to right
 *     starting from 0.0 and divides by (H-6)*(W-6) to obtain exactly the value numpy's mean of the cropped view
 *     gives.  c1 = (K1*255)^2, c2 = (K2*255)^2, cov_norm = 49/48 are passed in so that they are the caller's own
 *     float64 values.  7x7 uniform window, "reflect" borders; needs H, W >= 7 and W - 6 <= 8192.
 * ------------------------------------------------------------------------------------------- */
int rg_metrics_sse_u8(const uint8_t* pred, const uint8_t* gt, int32_t N, int64_t elems_per_image, uint64_t* sse,
                      rg_stream_t stream);
int rg_metrics_ssim_chunks(int32_t H, int32_t W);   /* number of partial sums per (image, channel); -1 if unsupported */
int rg_metrics_ssim_u8(const uint8_t* pred, const uint8_t* gt, int32_t N, int32_t H, int32_t W, int32_t C, double c1,
                       double c2, double cov_norm, double* smap_ws, double* chunk_sums, rg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * LPIPS (AlexNet) glue -- the learned perceptual metric of the reference's evaluator
 *   (src/metrics.py:67 lpips.LPIPS(net='alex'), :97-111 calculate_lpips).  The five feature convolutions run on
 *   rg_conv2d with RG_ACT_RELU; these two kernels are the rest of the graph.
 *   rg_maxpool3x3s2: bf16 [N,H,W,C] -> bf16 [N,(H-3)/2+1,(W-3)/2+1,C]  (MaxPool2d(3, stride 2)), C % 8 == 0.
 *   rg_lpips_layer:  f0, f1 bf16 [N][HW][C] post-ReLU features of the two images, lin fp32 [C] (the non-negative 1x1
 *     head); partial[n][b] (b < rg_lpips_layer_blocks(HW)) = sum over block b's pixels of
 *     sum_c lin[c] * (f0/(||f0||_2 + 1e-10) - f1/(||f1||_2 + 1e-10))^2.  The caller adds the partials of an image in
 *     block order and divides by HW (spatial average); the split depends on HW only, so the value of an image does not
 *     depend on the batch size.
 * ------------------------------------------------------------------------------------------- */
#define RG_LPIPS_MAX_BLOCKS 256
int rg_maxpool3x3s2(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, void* y, rg_stream_t stream);
int rg_lpips_layer_blocks(int32_t HW);
int rg_lpips_layer(const void* f0, const void* f1, const float* lin, int32_t N, int32_t HW, int32_t C, float* partial,
                   rg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RESTORAGEN_H_ */
