// K9-K11, K14: scheduler step, timestep embedding and the small layout / glue kernels of the sampling loop.
#include "common.cuh"
#include "internal.h"

namespace rg {

static inline unsigned grid_for(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    const long long cap = 148LL * 32;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
#define RG_GRID_STRIDE(i, n) \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// ---------------------------------------------------------------------------------------------- K11 scheduler
struct SchedParams {
    const float* eps_uc; float* sample; float* ets; float* cur;
    long long n; int do_cfg; float g; int store_slot; float w0, w1, w2, w3, w4; int use_cur, save_cur;
    float c_sample, c_eps;
};
__global__ void __launch_bounds__(256) sched_step_kernel(const SchedParams p) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, p.n) {
        float e;
        if (p.do_cfg) {
            const float u = p.eps_uc[i], c = p.eps_uc[p.n + i];
            e = u + p.g * (c - u);
        } else {
            e = p.eps_uc[i];
        }
        if (p.store_slot >= 0) p.ets[(long long)p.store_slot * p.n + i] = e;
        float mix = p.w4 * e;
        // history slots other than the one just written (its weight, if any, is folded into w4 by the host)
        if (p.w0 != 0.f) mix += p.w0 * p.ets[i];
        if (p.w1 != 0.f) mix += p.w1 * p.ets[p.n + i];
        if (p.w2 != 0.f) mix += p.w2 * p.ets[2 * p.n + i];
        if (p.w3 != 0.f) mix += p.w3 * p.ets[3 * p.n + i];
        const float s = p.sample[i];
        const float base = p.use_cur ? p.cur[i] : s;
        if (p.save_cur) p.cur[i] = s;
        p.sample[i] = p.c_sample * base - p.c_eps * mix;
    }
}

// ---------------------------------------------------------------------------------------------- K10 timestep embedding
__global__ void timestep_embedding_kernel(const float* t, int B, int dim, __nv_bfloat16* out) {
    pdl_trigger();
    pdl_wait();
    const int half = dim / 2;
    RG_GRID_STRIDE(i, (long long)B * dim) {
        const int b = (int)(i / dim), j = (int)(i % dim);
        const int f = j < half ? j : j - half;
        const float freq = expf(-logf(10000.0f) * (float)f / (float)half);
        const float a = t[b] * freq;
        // flip_sin_to_cos: first half cos, second half sin
        out[i] = f2op(j < half ? cosf(a) : sinf(a));
    }
}

// ---------------------------------------------------------------------------------------------- im2col for tiny Cin
template <bool IN_F32>
__global__ void im2col_small_kernel(const void* x_, int N, int n_mod, int H, int W, int Cin, int ks, int stride,
                                    int pad, int OH, int OW, int Kpad, __nv_bfloat16* out) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)N * OH * OW * Kpad;
    const int K = ks * ks * Cin;
    RG_GRID_STRIDE(i, total) {
        const int k = (int)(i % Kpad);
        const long long m = i / Kpad;
        float v = 0.f;
        if (k < K) {
            const int c = k % Cin, tap = k / Cin, kw = tap % ks, kh = tap / ks;
            const int ow = (int)(m % OW); const long long r = m / OW;
            const int oh = (int)(r % OH); const int n = (int)(r / OH) % n_mod;
            const int ih = oh * stride + kh - pad, iw = ow * stride + kw - pad;
            if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
                const long long idx = (((long long)n * H + ih) * W + iw) * Cin + c;
                v = IN_F32 ? reinterpret_cast<const float*>(x_)[idx]
                           : op2f(reinterpret_cast<const __nv_bfloat16*>(x_)[idx]);
            }
        }
        out[i] = f2op(v);
    }
}

// ---------------------------------------------------------------------------------------------- nearest 2x upsample (bf16, 16-B vectors)
// F.interpolate(mode="nearest"): src = min(floor(dst * in / out), in - 1)
__global__ void upsample_nearest_kernel(const uint4* x, int N, int H, int W, int C8, int OH, int OW, uint4* y) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)N * OH * OW * C8;
    const float sh = (float)H / (float)OH, sw = (float)W / (float)OW;
    RG_GRID_STRIDE(i, total) {
        const int c = (int)(i % C8); long long r = i / C8;
        const int ow = (int)(r % OW); r /= OW;
        const int oh = (int)(r % OH); const int n = (int)(r / OH);
        const int ih = min((int)floorf(oh * sh), H - 1), iw = min((int)floorf(ow * sw), W - 1);
        y[i] = x[(((long long)n * H + ih) * W + iw) * C8 + c];
    }
}

// ---------------------------------------------------------------------------------------------- layout
__global__ void nchw_to_nhwc_kernel(const float* x, int N, int C, int H, int W, float* y) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)N * C * H * W;
    RG_GRID_STRIDE(i, total) {   // i indexes the NHWC output
        const int c = (int)(i % C); long long r = i / C;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H); const int n = (int)(r / H);
        y[i] = x[(((long long)n * C + c) * H + h) * W + w];
    }
}
__global__ void nhwc_to_nchw_kernel(const float* x, int N, int C, int H, int W, float* y) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)N * C * H * W;
    RG_GRID_STRIDE(i, total) {   // i indexes the NCHW output
        const int w = (int)(i % W); long long r = i / W;
        const int h = (int)(r % H); r /= H;
        const int c = (int)(r % C); const int n = (int)(r / C);
        y[i] = x[(((long long)n * H + h) * W + w) * C + c];
    }
}

__global__ void preprocess_u8_kernel(const uint8_t* img, const float* mask, long long npix, float* out) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, npix) {
        const float keep = mask ? (mask[i] < 0.5f ? 1.f : 0.f) : 1.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // same fp32 sequence as VaeImageProcessor: x/255 -> 2x-1
            const float v = (float)img[i * 3 + c] / 255.0f;
            out[i * 3 + c] = (2.0f * v - 1.0f) * keep;
        }
    }
}
__global__ void postprocess_u8_kernel(const float* x, long long npix, int ldc, uint8_t* out) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, npix) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = x[i * ldc + c] / 2.0f + 0.5f;
            v = fminf(fmaxf(v, 0.f), 1.f);
            out[i * 3 + c] = (uint8_t)rintf(v * 255.0f);      // numpy round(): half to even
        }
    }
}

__global__ void vae_sample_kernel(const float* mom, long long ld, const float* eps_post, const float* noise,
                                  long long npix, float scaling, int add_noise, float sa, float sb, float* out) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, npix * 4) {
        const long long pix = i >> 2; const int c = (int)(i & 3);
        const float mean = mom[pix * ld + c];
        float logvar = mom[pix * ld + 4 + c];
        logvar = fminf(fmaxf(logvar, -30.f), 20.f);
        const float z = (mean + expf(0.5f * logvar) * eps_post[i]) * scaling;
        out[i] = add_noise ? sa * z + sb * noise[i] : z;
    }
}

__global__ void pack_unet_input_kernel(const float* lat, const float* mask, const float* masked, long long npix,
                                       float* out) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, npix * 9) {
        const long long pix = i / 9; const int c = (int)(i % 9);
        out[i] = c < 4 ? lat[pix * 4 + c] : (c == 4 ? mask[pix] : masked[pix * 4 + c - 5]);
    }
}

__global__ void mask_nearest_kernel(const float* m, int N, int H, int W, int h, int w, float* out) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, (long long)N * h * w) {
        const int x = (int)(i % w); long long r = i / w;
        const int y = (int)(r % h); const int n = (int)(r / h);
        // F.interpolate(mode="nearest"): src = floor(dst * (in/out))
        const int sy = min((int)floorf(y * ((float)H / h)), H - 1), sx = min((int)floorf(x * ((float)W / w)), W - 1);
        out[i] = m[((long long)n * H + sy) * W + sx];
    }
}

__global__ void pointwise_small_kernel(const float* x, long long npix, int Cin, int Cout, const float* W,
                                       const float* b, float scale_in, float* out) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, npix * Cout) {
        const long long pix = i / Cout; const int co = (int)(i % Cout);
        float acc = b ? b[co] : 0.f;
        for (int c = 0; c < Cin; ++c) acc += W[co * Cin + c] * (x[pix * Cin + c] * scale_in);
        out[i] = acc;
    }
}

__global__ void scale_f32_kernel(const float* x, float a, long long n, float* y) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, n) y[i] = x[i] * a;
}
__global__ void cast_f32_bf16_kernel(const float* x, long long n, __nv_bfloat16* y) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, n) y[i] = f2op(x[i]);
}

}  // namespace rg

using namespace rg;
#define RG_STREAM(s) reinterpret_cast<cudaStream_t>(s)

namespace rg {
// CLIPTextEmbeddings: out[b*T + t][:] = token_embedding[ids[b*T + t]][:] + position_embedding[t][:]  (fp32, 4 per thread)
__global__ void embed_tokens_kernel(const int32_t* ids, const float4* tok, const float4* pos, int rows, int T, int C4,
                                    int vocab, float4* out) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, (long long)rows * C4) {
        const int r = (int)(i / C4), c = (int)(i % C4);
        int id = ids[r];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        const float4 a = tok[(long long)id * C4 + c], b = pos[(long long)(r % T) * C4 + c];
        out[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* x, long long n, float* y) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, n) y[i] = op2f(x[i]);
}
// quick_gelu(x) = x * sigmoid(1.702 x), bf16 in place (CLIP MLP activation)
__global__ void quick_gelu_bf16_kernel(__nv_bfloat162* x, long long n2) {
    pdl_trigger();
    pdl_wait();
    RG_GRID_STRIDE(i, n2) {
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(&x[i]));
        const uint32_t o = pack_bf16x2(v.x / (1.f + __expf(-1.702f * v.x)), v.y / (1.f + __expf(-1.702f * v.y)));
        x[i] = *reinterpret_cast<const __nv_bfloat162*>(&o);
    }
}

}  // namespace rg

extern "C" int rg_sched_step(const rg_sched_t* s, rg_stream_t stream) {
    if (!s || !s->eps_uc || !s->sample) return set_error(RG_ERR_ARG, "sched_step: null pointer");
    const bool hist = s->w[0] != 0.f || s->w[1] != 0.f || s->w[2] != 0.f || s->w[3] != 0.f || s->store_slot >= 0;
    if (hist && !s->ets) return set_error(RG_ERR_ARG, "sched_step: history needed but ets is NULL");
    if ((s->use_cur || s->save_cur) && !s->cur_sample) return set_error(RG_ERR_ARG, "sched_step: cur_sample is NULL");
    if (s->store_slot > 3) return set_error(RG_ERR_ARG, "sched_step: bad slot");
    SchedParams p{s->eps_uc, s->sample, s->ets, s->cur_sample, s->n, s->do_cfg, s->guidance, s->store_slot,
                  s->w[0], s->w[1], s->w[2], s->w[3], s->w[4], s->use_cur, s->save_cur, s->c_sample, s->c_eps};
    // the freshly stored slot is read back through `e`: fold its weight into w4
    if (s->store_slot >= 0) {
        float* ws[4] = {&p.w0, &p.w1, &p.w2, &p.w3};
        p.w4 += *ws[s->store_slot];
        *ws[s->store_slot] = 0.f;
    }
    launch_kernel(sched_step_kernel, dim3(grid_for(s->n, 256)), dim3(256), 0, RG_STREAM(stream), p);
    count_launch();
    return check_launch("sched_step_kernel");
}

extern "C" int rg_timestep_embedding(const float* t, int32_t B, int32_t dim, void* out, rg_stream_t stream) {
    if (!t || !out || dim % 2) return set_error(RG_ERR_ARG, "timestep_embedding: bad argument");
    launch_kernel(timestep_embedding_kernel, dim3(grid_for((long long)B * dim, 256)), dim3(256), 0, RG_STREAM(stream), 
        t, B, dim, reinterpret_cast<__nv_bfloat16*>(out));
    count_launch();
    return check_launch("timestep_embedding_kernel");
}

extern "C" int rg_im2col_small(const void* x, int32_t in_dtype, int32_t N, int32_t n_mod, int32_t H, int32_t W,
                               int32_t Cin, int32_t ksize, int32_t stride, int32_t pad, int32_t OH, int32_t OW,
                               int32_t Kpad, void* out, rg_stream_t stream) {
    if (!x || !out || ksize * ksize * Cin > Kpad || n_mod < 1) return set_error(RG_ERR_ARG, "im2col_small: bad argument");
    const long long total = (long long)N * OH * OW * Kpad;
    auto* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (in_dtype == RG_DT_F32)
        launch_kernel(im2col_small_kernel<true>, dim3(grid_for(total, 256)), dim3(256), 0, RG_STREAM(stream), x, N, n_mod, H, W, Cin, ksize,
                                                                                      stride, pad, OH, OW, Kpad, o);
    else
        launch_kernel(im2col_small_kernel<false>, dim3(grid_for(total, 256)), dim3(256), 0, RG_STREAM(stream), x, N, n_mod, H, W, Cin, ksize,
                                                                                       stride, pad, OH, OW, Kpad, o);
    count_launch();
    return check_launch("im2col_small_kernel");
}

extern "C" int rg_upsample_nearest(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t OH, int32_t OW,
                                   void* y, rg_stream_t stream) {
    if (!x || !y || C % 8 || OH < 1 || OW < 1) return set_error(RG_ERR_ARG, "upsample_nearest: bad argument");
    const long long total = (long long)N * OH * OW * (C / 8);
    launch_kernel(upsample_nearest_kernel, dim3(grid_for(total, 256)), dim3(256), 0, RG_STREAM(stream), 
        reinterpret_cast<const uint4*>(x), N, H, W, C / 8, OH, OW, reinterpret_cast<uint4*>(y));
    count_launch();
    return check_launch("upsample_nearest_kernel");
}

extern "C" int rg_nchw_to_nhwc(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, float* y, rg_stream_t stream) {
    if (!x || !y) return set_error(RG_ERR_ARG, "nchw_to_nhwc: null pointer");
    launch_kernel(nchw_to_nhwc_kernel, dim3(grid_for((long long)N * C * H * W, 256)), dim3(256), 0, RG_STREAM(stream), x, N, C, H, W, y);
    count_launch();
    return check_launch("nchw_to_nhwc_kernel");
}
extern "C" int rg_nhwc_to_nchw(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, float* y, rg_stream_t stream) {
    if (!x || !y) return set_error(RG_ERR_ARG, "nhwc_to_nchw: null pointer");
    launch_kernel(nhwc_to_nchw_kernel, dim3(grid_for((long long)N * C * H * W, 256)), dim3(256), 0, RG_STREAM(stream), x, N, C, H, W, y);
    count_launch();
    return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int rg_preprocess_u8(const uint8_t* img, const float* mask, int32_t N, int32_t H, int32_t W, float* out,
                                rg_stream_t stream) {
    if (!img || !out) return set_error(RG_ERR_ARG, "preprocess_u8: null pointer");
    const long long npix = (long long)N * H * W;
    launch_kernel(preprocess_u8_kernel, dim3(grid_for(npix, 256)), dim3(256), 0, RG_STREAM(stream), img, mask, npix, out);
    count_launch();
    return check_launch("preprocess_u8_kernel");
}
extern "C" int rg_postprocess_u8(const float* x, int32_t N, int32_t H, int32_t W, int32_t ldc, uint8_t* out,
                                 rg_stream_t stream) {
    if (!x || !out || ldc < 3) return set_error(RG_ERR_ARG, "postprocess_u8: bad argument");
    const long long npix = (long long)N * H * W;
    launch_kernel(postprocess_u8_kernel, dim3(grid_for(npix, 256)), dim3(256), 0, RG_STREAM(stream), x, npix, ldc, out);
    count_launch();
    return check_launch("postprocess_u8_kernel");
}

extern "C" int rg_vae_sample(const float* moments, int64_t moments_ld, const float* eps_post, const float* noise,
                             int64_t npix, float scaling, int32_t add_noise, float sqrt_ac, float sqrt_1mac,
                             float* out, rg_stream_t stream) {
    if (!moments || !eps_post || !out || (add_noise && !noise)) return set_error(RG_ERR_ARG, "vae_sample: null pointer");
    launch_kernel(vae_sample_kernel, dim3(grid_for(npix * 4, 256)), dim3(256), 0, RG_STREAM(stream), moments, moments_ld, eps_post, noise, npix,
                                                                              scaling, add_noise, sqrt_ac, sqrt_1mac, out);
    count_launch();
    return check_launch("vae_sample_kernel");
}

extern "C" int rg_pack_unet_input(const float* latents, const float* mask, const float* masked, int64_t npix,
                                  float* out, rg_stream_t stream) {
    if (!latents || !mask || !masked || !out) return set_error(RG_ERR_ARG, "pack_unet_input: null pointer");
    launch_kernel(pack_unet_input_kernel, dim3(grid_for(npix * 9, 256)), dim3(256), 0, RG_STREAM(stream), latents, mask, masked, npix, out);
    count_launch();
    return check_launch("pack_unet_input_kernel");
}

extern "C" int rg_mask_nearest(const float* mask, int32_t N, int32_t H, int32_t W, int32_t h, int32_t w, float* out,
                               rg_stream_t stream) {
    if (!mask || !out) return set_error(RG_ERR_ARG, "mask_nearest: null pointer");
    launch_kernel(mask_nearest_kernel, dim3(grid_for((long long)N * h * w, 256)), dim3(256), 0, RG_STREAM(stream), mask, N, H, W, h, w, out);
    count_launch();
    return check_launch("mask_nearest_kernel");
}

extern "C" int rg_pointwise_small(const float* x, int64_t npix, int32_t Cin, int32_t Cout, const float* W,
                                  const float* b, float scale_in, float* out, rg_stream_t stream) {
    if (!x || !W || !out || Cin > 64 || Cout > 64) return set_error(RG_ERR_ARG, "pointwise_small: bad argument");
    launch_kernel(pointwise_small_kernel, dim3(grid_for(npix * Cout, 256)), dim3(256), 0, RG_STREAM(stream), x, npix, Cin, Cout, W, b, scale_in, out);
    count_launch();
    return check_launch("pointwise_small_kernel");
}

extern "C" int rg_scale_f32(const float* x, float a, int64_t n, float* y, rg_stream_t stream) {
    if (!x || !y) return set_error(RG_ERR_ARG, "scale_f32: null pointer");
    launch_kernel(scale_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, RG_STREAM(stream), x, a, n, y);
    count_launch();
    return check_launch("scale_f32_kernel");
}
extern "C" int rg_cast_f32_bf16(const float* x, int64_t n, void* y, rg_stream_t stream) {
    if (!x || !y) return set_error(RG_ERR_ARG, "cast_f32_bf16: null pointer");
    launch_kernel(cast_f32_bf16_kernel, dim3(grid_for(n, 256)), dim3(256), 0, RG_STREAM(stream), x, n, reinterpret_cast<__nv_bfloat16*>(y));
    count_launch();
    return check_launch("cast_f32_bf16_kernel");
}
extern "C" int rg_memset_zero(void* p, int64_t bytes, rg_stream_t stream) {
    if (!p) return set_error(RG_ERR_ARG, "memset_zero: null pointer");
    cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, RG_STREAM(stream));
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync");
    return RG_OK;
}

extern "C" int rg_embed_tokens(const int32_t* ids, const float* token_embedding, const float* position_embedding,
                               int32_t B, int32_t T, int32_t C, int32_t vocab, float* out, rg_stream_t stream) {
    if (!ids || !token_embedding || !position_embedding || !out || B < 1 || T < 1 || C < 4 || C % 4 || vocab < 1)
        return set_error(RG_ERR_ARG, "embed_tokens: bad argument");
    launch_kernel(embed_tokens_kernel, dim3(grid_for((long long)B * T * (C / 4), 256)), dim3(256), 0, RG_STREAM(stream), 
        ids, reinterpret_cast<const float4*>(token_embedding), reinterpret_cast<const float4*>(position_embedding), B * T, T,
        C / 4, vocab, reinterpret_cast<float4*>(out));
    count_launch();
    return check_launch("embed_tokens_kernel");
}
extern "C" int rg_quick_gelu_bf16(void* x, int64_t n, rg_stream_t stream) {
    if (!x || n < 0 || n % 2) return set_error(RG_ERR_ARG, "quick_gelu_bf16: bad argument");
    if (n == 0) return RG_OK;
    launch_kernel(quick_gelu_bf16_kernel, dim3(grid_for(n / 2, 256)), dim3(256), 0, RG_STREAM(stream), reinterpret_cast<__nv_bfloat162*>(x), n / 2);
    count_launch();
    return check_launch("quick_gelu_bf16_kernel");
}
extern "C" int rg_cast_bf16_f32(const void* x, int64_t n, float* y, rg_stream_t stream) {
    if (!x || !y) return set_error(RG_ERR_ARG, "cast_bf16_f32: null pointer");
    launch_kernel(cast_bf16_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, RG_STREAM(stream), reinterpret_cast<const __nv_bfloat16*>(x), n, y);
    count_launch();
    return check_launch("cast_bf16_f32_kernel");
}
