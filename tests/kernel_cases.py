"""GPU parity cases for the individual kernels of librestoragen.so.

Each case builds seeded inputs, runs the CUDA kernel through the C ABI (via ops.py) and compares with a plain
PyTorch fp32 reference of the same op on the same (bf16-rounded) inputs.  The cases are plain functions so they
can be run one per process by tools/gpu_kernel_check.py (a faulting kernel poisons its CUDA context) and
in-process by tests/test_kernels_gpu.py.

Tolerances (relative L2 unless noted): bf16 outputs 4e-3 (one bf16 rounding of an fp32-accumulated result is
2^-9 = 2e-3 worst case), fp32 outputs 2e-5 for GEMMs (fp32 accumulation order differs), attention 8e-3
(P is rounded to bf16 before the second GEMM, as in every flash-attention kernel).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from image_restoration_and_enhancement_b200 import ops
from image_restoration_and_enhancement_b200._lib import RG_ACT_GEGLU, RG_ACT_NONE, RG_ACT_SILU

DEV = "cuda"
TOL_BF16, TOL_F32, TOL_ATTN = 4e-3, 2e-5, 8e-3


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


def pack_w(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] (taps outer, channels inner)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


# ------------------------------------------------------------------------------------------------ GEMM / conv
def case_linear(M=256, K=128, N=160, bias=True, res=None, act=RG_ACT_NONE, f32_out=False, scale=1.0, seed=0):
    _setup()
    x = _rand((M, K), seed)
    w = _rand((N, K), seed + 1, 1.0 / math.sqrt(K))
    b = _rand((N,), seed + 2, 0.5, torch.float32) if bias else None
    r = None
    if res == "f32":
        r = _rand((M, N), seed + 3, 1.0, torch.float32)
    elif res == "bf16":
        r = _rand((M, N), seed + 3, 1.0)
    ref = (x.float() @ w.float().t()) * scale
    if b is not None:
        ref = ref + b
    if r is not None:
        ref = ref + r.float()
    if act == RG_ACT_SILU:
        ref = F.silu(ref)
    ob, of = ops.linear(x, w, bias=b, res=r, act=act, scale=scale, out_bf16=not f32_out, out_f32=f32_out)
    torch.cuda.synchronize()
    out = of if f32_out else ob
    return rel_l2(out, ref), (TOL_F32 if f32_out else TOL_BF16)


def case_splitk_invariance(seed=130):
    """Deterministic split-K (epilogue_splitk): the partials are added in slice order by whichever warp arrives last, so
    two launches give the same bits -- 8 slices (8x8 level, batch 2), 4 slices (16x16, bf16 out), 2 slices (8x8 level at
    batch 16) and a token linear (4 slices).  Returns the number of mismatching comparisons (tolerance 0)."""
    _setup()
    bad = 0
    for (NB, H, W, Cin, Cout, f32_out) in ((2, 8, 8, 1280, 1280, True), (2, 16, 16, 1280, 640, False), (16, 8, 8, 1280, 1280, True)):
        x = _rand((NB, H, W, Cin), seed)
        w = pack_w(_rand((Cout, Cin, 3, 3), seed + 1, 1.0 / math.sqrt(Cin * 9)))
        b = _rand((Cout,), seed + 2, 0.5, torch.float32)
        r = _rand((NB, H, W, Cout), seed + 3, 1.0, torch.float32)
        kw = dict(kh=3, kw=3, pad_t=1, pad_l=1, bias=b, out_bf16=not f32_out, out_f32=f32_out)
        pick = (lambda o: o[1]) if f32_out else (lambda o: o[0])
        runs = [pick(ops.conv2d(x, w, res=r, **kw)).clone() for _ in range(4)]
        torch.cuda.synchronize()
        bad += sum(int(not torch.equal(runs[0], o)) for o in runs[1:])
    xt = _rand((2 * 256, 5120), seed + 5)
    wt = _rand((1280, 5120), seed + 6, 1.0 / math.sqrt(5120))
    rt = _rand((2 * 256, 1280), seed + 7, 1.0, torch.float32)
    a1, _ = ops.linear(xt, wt, images=2, res=rt, out_bf16=True)
    a2, _ = ops.linear(xt, wt, images=2, res=rt, out_bf16=True)
    ref = xt.float() @ wt.float().t() + rt
    torch.cuda.synchronize()
    bad += int(not torch.equal(a1, a2)) + int(rel_l2(a1, ref) > TOL_BF16)
    return float(bad), 0.0


def case_conv_relu(seed=140):
    """RG_ACT_RELU (the LPIPS feature convs) through the three epilogue flavours: TMA run-time-flag (Cout 256 and 64),
    direct stores (Cout 192 / 384: not a multiple of the column tile)."""
    from image_restoration_and_enhancement_b200._lib import RG_ACT_RELU
    _setup()
    worst = 0.0
    for i, (Cin, Cout, k) in enumerate(((256, 256, 3), (384, 64, 1), (192, 384, 3), (1600, 192, 1))):
        x = _rand((2, 15, 17, Cin), seed + 4 * i)
        w = _rand((Cout, Cin, k, k), seed + 4 * i + 1, 1.0 / math.sqrt(Cin * k * k))
        b = _rand((Cout,), seed + 4 * i + 2, 0.5, torch.float32)
        ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=k // 2))
        ob, _ = ops.conv2d(x, pack_w(w), kh=k, kw=k, pad_t=k // 2, pad_l=k // 2, bias=b, act=RG_ACT_RELU, out_bf16=True)
        torch.cuda.synchronize()
        worst = max(worst, rel_l2(ob.float().permute(0, 3, 1, 2), ref))
        assert float(ob.float().min()) >= 0.0
    return worst, TOL_BF16


def case_maxpool(seed=150):
    _setup()
    x = _rand((3, 31, 127, 64), seed)
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1)
    y = ops.maxpool3x3s2(x)
    torch.cuda.synchronize()
    return float((y.float() - ref).abs().max()), 0.0


def case_lpips_layer(seed=160):
    """normalise / diff / lin head / spatial sum of one LPIPS level vs the torch expression on the same bf16 features;
    and an image's partial sums do not depend on the batch it sits in."""
    _setup()
    worst = 0.0
    for (H, W, Cc) in ((31, 31, 384), (127, 127, 64), (7, 9, 256)):
        f0 = F.relu(_rand((3, H, W, Cc), seed)).contiguous()
        f1 = F.relu(_rand((3, H, W, Cc), seed + 1)).contiguous()
        lin = _rand((Cc,), seed + 2, 1.0, torch.float32).abs().contiguous()
        part = ops.lpips_layer(f0, f1, lin)
        one = ops.lpips_layer(f0[1:2].contiguous(), f1[1:2].contiguous(), lin)
        torch.cuda.synchronize()
        a, b = f0.float(), f1.float()
        na = a / (a.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
        nb = b / (b.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
        ref = ((na - nb).pow(2) * lin).sum(dim=(1, 2, 3))
        got = part.double().sum(dim=1).float()
        worst = max(worst, rel_l2(got, ref))
        assert torch.equal(part[1:2], one)
    return worst, 2e-5


def case_geglu(M=300, K=320, C4=1280, seed=5):
    """ff.net.0 (GEGLU): proj -> chunk(2) -> a * gelu(g); weight rows interleaved per 32-wide unit on the host."""
    _setup()
    from image_restoration_and_enhancement_b200.weights import interleave_geglu
    x = _rand((M, K), seed)
    w = _rand((2 * C4, K), seed + 1, 1.0 / math.sqrt(K))
    b = _rand((2 * C4,), seed + 2, 0.5, torch.float32)
    h = x.float() @ w.float().t() + b
    a, g = h.chunk(2, dim=-1)
    ref = a * F.gelu(g)
    wi, bi = interleave_geglu(w, b)
    ob, _ = ops.linear(x, wi, bias=bi, act=RG_ACT_GEGLU, out_bf16=True)
    torch.cuda.synchronize()
    return rel_l2(ob, ref), TOL_BF16


def case_conv(N=2, H=16, W=16, Cin=64, Cout=160, k=3, stride=1, pad=1, asym=False, bias=True, bias_n=False,
              res=None, x2c=0, f32_out=False, both_out=False, seed=10):
    """3x3 / 1x1 convolution vs F.conv2d.  ``asym``: VAE-encoder downsample (pad (0,1,0,1), stride 2, pad 0)."""
    _setup()
    x = _rand((N, H, W, Cin), seed)
    w = _rand((Cout, Cin, k, k), seed + 1, 1.0 / math.sqrt(Cin * k * k))
    b = _rand((Cout,), seed + 2, 0.5, torch.float32) if bias else None
    xn = x.float().permute(0, 3, 1, 2)
    if asym:
        ref = F.conv2d(F.pad(xn, (0, 1, 0, 1)), w.float(), None, stride=2, padding=0)
        pad_t = pad_l = 0
    else:
        ref = F.conv2d(xn, w.float(), None, stride=stride, padding=pad)
        pad_t = pad_l = pad
    OH, OW = ref.shape[2], ref.shape[3]
    wp = pack_w(w)
    x2 = None
    if x2c:
        x2 = _rand((N, OH, OW, x2c), seed + 4)
        w2 = _rand((Cout, x2c), seed + 5, 1.0 / math.sqrt(x2c))
        ref = ref + torch.einsum("nhwc,oc->nohw", x2.float(), w2.float())
        wp = torch.cat([wp, w2], dim=1).contiguous()
    if b is not None:
        ref = ref + b[None, :, None, None]
    bn = None
    if bias_n:
        bn = _rand((N, Cout), seed + 6, 0.5, torch.float32)
        ref = ref + bn[:, :, None, None]
    r = None
    if res is not None:
        r = _rand((N, OH, OW, Cout), seed + 3, 1.0, torch.float32 if res == "f32" else torch.bfloat16)
        ref = ref + r.float().permute(0, 3, 1, 2)
    ob, of = ops.conv2d(x, wp, kh=k, kw=k, stride=2 if asym else stride, pad_t=pad_t, pad_l=pad_l, OH=OH, OW=OW,
                        x2=x2, bias=b, bias_n=bn, res=r, out_bf16=both_out or not f32_out, out_f32=both_out or f32_out)
    torch.cuda.synchronize()
    if both_out:        # fp32 stream + bf16 copy from one epilogue (resnet conv2 / proj_out feeding a down/up-sampler)
        e32 = rel_l2(of.float().permute(0, 3, 1, 2), ref)
        e16 = rel_l2(ob.float().permute(0, 3, 1, 2), ref)
        return max(e32 / TOL_F32, e16 / TOL_BF16), 1.0
    out = (of if f32_out else ob).float().permute(0, 3, 1, 2)
    return rel_l2(out, ref), (TOL_F32 if f32_out else TOL_BF16)


def case_conv_strided_out(seed=29):
    """Parity-split write: a 2x2 conv whose fp32 output lands on every second pixel of a larger tensor (the
    nearest-2x upsample + 3x3 conv split of unet.py / vae.py)."""
    _setup()
    N, H, W, C = 2, 16, 16, 320
    x = _rand((N, H, W, C), seed)
    w = _rand((C, C, 2, 2), seed + 1, 1.0 / math.sqrt(C * 4))
    b = _rand((C,), seed + 2, 0.5, torch.float32)
    out = torch.zeros((N, 2 * H, 2 * W, C), dtype=torch.float32, device=DEV)
    sn, sh, sw = out.stride(0), out.stride(1), out.stride(2)
    errs = []
    for py in (0, 1):
        for px in (0, 1):
            ops.conv2d(x, pack_w(w), kh=2, kw=2, pad_t=1 - py, pad_l=1 - px, OH=H, OW=W, bias=b,
                       out_f32=out[:, py:, px:], out_strides=(sn, 2 * sh, 2 * sw))
            xn = F.pad(x.float().permute(0, 3, 1, 2), (1 - px, px, 1 - py, py))
            ref = F.conv2d(xn, w.float(), b)
            torch.cuda.synchronize()
            errs.append(rel_l2(out[:, py::2, px::2].permute(0, 3, 1, 2), ref))
    return max(errs), TOL_F32


def case_upsample_conv_one_launch(N=2, H=16, W=16, Cin=1280, Cout=1280, seed=150):
    """rg_conv_t::parities = 4: nearest-2x upsample + 3x3 conv as ONE launch of the four stacked 2x2 parity kernels, against
    F.interpolate + F.conv2d in fp32 (weights packed by the product's own upsample_parity_weights)."""
    _setup()
    from image_restoration_and_enhancement_b200.weights import upsample_parity_weights
    x = _rand((N, H, W, Cin), seed)
    w = _rand((Cout, Cin, 3, 3), seed + 1, 1.0 / math.sqrt(Cin * 9), torch.float32)
    w = w.to(torch.bfloat16).float()
    b = _rand((Cout,), seed + 2, 0.5, torch.float32)
    par = {(py, px): wp for py, px, wp in upsample_parity_weights(w)}
    wst = torch.cat([par[(py, px)] for py in (0, 1) for px in (0, 1)], dim=0).to(torch.bfloat16).contiguous()
    out = torch.zeros((N, 2 * H, 2 * W, Cout), dtype=torch.float32, device=DEV)
    ops.conv2d(x, wst, kh=2, kw=2, OH=H, OW=W, bias=b, out_f32=out, parities=4)
    torch.cuda.synchronize()
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w, b, padding=1)
    # the parity kernels are sums of up to four bf16-representable taps rounded to bf16 once more: operand rounding, as for
    # every conv of the library
    return rel_l2(out.permute(0, 3, 1, 2), ref), TOL_BF16


# ------------------------------------------------------------------------------------------------ attention
def case_attention(B=2, heads=8, d=40, Nq=1024, Nk=None, seed=20, fused_qkv=True, dtype=torch.bfloat16, causal=False):
    """``dtype`` fp16: the UNet's path (two exponentials per MUFU op, denominator from a ones column of V)."""
    _setup()
    Nk = Nq if Nk is None else Nk
    C = heads * d
    if fused_qkv and Nk == Nq:
        qkv = _rand((B, Nq, 3 * C), seed, dtype=dtype)
        q = qkv[:, :, 0:C].unflatten(2, (heads, d))
        k = qkv[:, :, C:2 * C].unflatten(2, (heads, d))
        v = qkv[:, :, 2 * C:].unflatten(2, (heads, d))
    else:
        q = _rand((B, Nq, heads, d), seed, dtype=dtype)
        kv = _rand((B, Nk, 2 * C), seed + 1, dtype=dtype)
        k = kv[:, :, :C].unflatten(2, (heads, d))
        v = kv[:, :, C:].unflatten(2, (heads, d))
    scale = d ** -0.5
    ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2),
                                         v.float().transpose(1, 2), is_causal=causal).transpose(1, 2)
    out = ops.attention(q, k, v, scale, causal=causal)
    torch.cuda.synchronize()
    return rel_l2(out, ref), TOL_ATTN


def case_clip_glue(seed=49):
    """CLIPTextEmbeddings gather + add, quick_gelu in place, bf16 -> fp32 cast."""
    _setup()
    g = torch.Generator(device="cpu").manual_seed(seed)
    tok = torch.randn((1000, 768), generator=g).to(DEV)
    pos = torch.randn((77, 768), generator=g).to(DEV)
    ids = torch.randint(0, 1000, (3, 77), generator=g).to(torch.int32).to(DEV)
    e = ops.embed_tokens(ids, tok, pos)
    ref = (tok[ids.long()] + pos[None]).view(-1, 768)
    x = _rand((231, 3072), seed + 1, 2.0)
    want = (x.float() * torch.sigmoid(1.702 * x.float()))
    got = ops.quick_gelu_(x.clone())
    c = ops.cast_bf16_f32(x)
    torch.cuda.synchronize()
    assert torch.equal(e, ref) and torch.equal(c, x.float())
    return rel_l2(got, want), TOL_BF16


def case_attention_large_logits(seed=47):
    """fp16 path with logits spread over +-60 (log2 units): exercises the lazy rescale of O and of the ones column."""
    _setup()
    B, heads, d, N = 1, 8, 40, 1024
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = (torch.randn((B, N, heads, d), generator=g) * 3.0).to(torch.float16).to(DEV)
    k = (torch.randn((B, N, heads, d), generator=g) * 3.0).to(torch.float16).to(DEV)
    k[:, 700:] *= 2.0                                   # later key tiles dominate: the running max must be raised
    v = torch.randn((B, N, heads, d), generator=g).to(torch.float16).to(DEV)
    ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2),
                                         v.float().transpose(1, 2)).transpose(1, 2)
    out = ops.attention(q, k, v, d ** -0.5)
    torch.cuda.synchronize()
    return rel_l2(out, ref), TOL_ATTN


def case_linear_f16(seed=48):
    """q|k|v projection written as fp16 (16-bit output dtype flag of rg_conv2d)."""
    _setup()
    x = _rand((4096, 320), seed)
    w = _rand((960, 320), seed + 1, 1.0 / math.sqrt(320))
    ob, _ = ops.linear(x, w, out_bf16=True, out_half=torch.float16)
    torch.cuda.synchronize()
    assert ob.dtype == torch.float16
    return rel_l2(ob, x.float() @ w.float().t()), 1e-3      # fp16 rounding: 2^-11


# ------------------------------------------------------------------------------------------------ norms
def case_groupnorm(N=2, H=16, W=16, C1=320, C2=0, in_f32=True, silu=True, eps=1e-5, raw=False, seed=30):
    _setup()
    dt = torch.float32 if in_f32 else torch.bfloat16
    x1 = _rand((N, H, W, C1), seed, 1.5, dt) + 0.3
    x2 = (_rand((N, H, W, C2), seed + 1, 0.7, dt) - 0.2) if C2 else None
    C = C1 + C2
    gamma = _rand((C,), seed + 2, 0.2, torch.float32) + 1.0
    beta = _rand((C,), seed + 3, 0.2, torch.float32)
    xc = x1 if x2 is None else torch.cat([x1, x2], dim=3)
    ref = F.group_norm(xc.float().permute(0, 3, 1, 2), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    y, r = ops.groupnorm(x1, gamma, beta, eps=eps, silu=silu, x2=x2, want_raw=raw)
    torch.cuda.synchronize()
    err = rel_l2(y.float().permute(0, 3, 1, 2), ref)
    if raw:
        err = max(err, rel_l2(r, xc))
    return err, TOL_BF16


def case_groupnorm_shapes_agree(seed=34, N=16, H=16, W=16, C1=640, C2=320):
    """GroupNorm is batch-invariant BITWISE: a batch of 16 against each image normalised on its own (fp32 two-source
    input with raw copy)."""
    _setup()
    x1 = _rand((N, H, W, C1), seed, 1.5, torch.float32) + 0.3
    x2 = _rand((N, H, W, C2), seed + 1, 0.7, torch.float32) - 0.2
    gamma = _rand((C1 + C2,), seed + 2, 0.2, torch.float32) + 1.0
    beta = _rand((C1 + C2,), seed + 3, 0.2, torch.float32)
    y, r = ops.groupnorm(x1, gamma, beta, silu=True, x2=x2, want_raw=True)
    ref = F.silu(F.group_norm(torch.cat([x1, x2], dim=3).permute(0, 3, 1, 2), 32, gamma, beta, 1e-5))
    err = rel_l2(y.float().permute(0, 3, 1, 2), ref)
    bad = 0
    for n in (0, N // 2, N - 1):
        y1, r1 = ops.groupnorm(x1[n:n + 1].contiguous(), gamma, beta, silu=True, x2=x2[n:n + 1].contiguous(), want_raw=True)
        torch.cuda.synchronize()
        bad += int((y1 != y[n:n + 1]).sum()) + int((r1 != r[n:n + 1]).sum())
    return (err if bad == 0 else 1.0), TOL_BF16


def case_layernorm(rows=1000, C=320, in_f32=True, seed=40):
    _setup()
    x = _rand((rows, C), seed, 2.0, torch.float32 if in_f32 else torch.bfloat16) + 0.5
    gamma = _rand((C,), seed + 1, 0.2, torch.float32) + 1.0
    beta = _rand((C,), seed + 2, 0.2, torch.float32)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    y = ops.layernorm(x, gamma, beta, 1e-5)
    torch.cuda.synchronize()
    return rel_l2(y, ref), TOL_BF16


def case_softmax_rows(rows=64, cols=4096, seed=50):
    _setup()
    x = _rand((rows, cols), seed, 3.0)
    ref = torch.softmax(x.float(), dim=-1)
    ops.softmax_rows_(x)
    torch.cuda.synchronize()
    return rel_l2(x, ref), TOL_BF16


# ------------------------------------------------------------------------------------------------ glue
def case_timestep_embedding():
    _setup()
    t = torch.tensor([1.0, 501.0, 951.0, 727.0], device=DEV)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=DEV) / half)
    a = t[:, None] * freqs[None]
    ref = torch.cat([torch.cos(a), torch.sin(a)], dim=-1)
    out = ops.timestep_embedding(t, 320)
    torch.cuda.synchronize()
    return float((out.float() - ref).abs().max()), 8e-3     # abs: bf16 rounding of values in [-1,1]


def case_sched(do_cfg=True, seed=60):
    _setup()
    n = 4 * 64 * 64
    eps = _rand((2 if do_cfg else 1, n), seed, 1.0, torch.float32)
    sample = _rand((n,), seed + 1, 1.0, torch.float32)
    ets = _rand((4, n), seed + 2, 1.0, torch.float32)
    cur = _rand((n,), seed + 3, 1.0, torch.float32)
    g = 7.5
    e = eps[0] + g * (eps[1] - eps[0]) if do_cfg else eps[0]
    w = [0.0, -59 / 24, 37 / 24, -9 / 24, 55 / 24]      # store into slot 0, history in slots 1..3
    ref_mix = w[4] * e + w[1] * ets[1] + w[2] * ets[2] + w[3] * ets[3]
    ref = 1.01 * sample - 0.07 * ref_mix
    s2, ets2 = sample.clone(), ets.clone()
    ops.sched_step(eps, s2, ets=ets2, cur=cur, do_cfg=do_cfg, guidance=g, store_slot=0, w=w, use_cur=False,
                   save_cur=False, c_sample=1.01, c_eps=0.07)
    torch.cuda.synchronize()
    err = rel_l2(s2, ref)
    err = max(err, rel_l2(ets2[0], e))
    return err, 1e-6


def case_im2col(seed=70):
    _setup()
    x = _rand((2, 16, 16, 4), seed, 1.0, torch.float32)
    w = _rand((320, 4, 3, 3), seed + 1, 0.2)
    ref = F.conv2d(x.to(torch.bfloat16).float().permute(0, 3, 1, 2), w.float(), None, padding=1)
    cols = ops.im2col_small(x, 4, 3, 1, 1, 16, 16, 64)      # N_out = 4 = 2 x CFG duplication
    wp = torch.zeros((320, 64), dtype=torch.bfloat16, device=DEV)
    wp[:, :36] = pack_w(w)
    ob, _ = ops.conv2d(cols, wp, out_bf16=True)
    torch.cuda.synchronize()
    out = ob.float().permute(0, 3, 1, 2)
    return max(rel_l2(out[:2], ref), rel_l2(out[2:], ref)), TOL_BF16


def case_upsample(seed=80):
    _setup()
    x = _rand((2, 6, 8, 64), seed)
    errs = []
    for size in ((12, 16), (11, 16), (11, 15)):
        ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=size, mode="nearest").permute(0, 2, 3, 1)
        y = ops.upsample_nearest(x, *size)
        torch.cuda.synchronize()
        errs.append(float((y.float() - ref).abs().max()))
    return max(errs), 0.0


def case_layout_and_io(seed=90):
    _setup()
    x = _rand((2, 4, 8, 6), seed, 1.0, torch.float32)
    y = ops.nchw_to_nhwc(x)
    z = ops.nhwc_to_nchw(y)
    e1 = float((y - x.permute(0, 2, 3, 1)).abs().max()) + float((z - x).abs().max())
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (2, 16, 16, 3), generator=g, dtype=torch.uint8).to(DEV)
    pre = ops.preprocess_u8(img)
    ref = 2.0 * (img.float() / 255.0) - 1.0
    e2 = float((pre - ref).abs().max())
    dec = _rand((2, 16, 16, 3), seed + 1, 1.0, torch.float32)
    post = ops.postprocess_u8(dec)
    refp = ((dec / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
    e3 = float((post.int() - refp.int()).abs().max())
    mom = _rand((2, 8, 8, 8), seed + 2, 1.0, torch.float32)
    ep = _rand((2, 8, 8, 4), seed + 3, 1.0, torch.float32)
    nz = _rand((2, 8, 8, 4), seed + 4, 1.0, torch.float32)
    lat = ops.vae_sample(mom, ep, nz, 0.18215, 0.8, 0.6)
    mean, logvar = mom[..., :4], mom[..., 4:].clamp(-30, 20)
    refl = 0.8 * ((mean + torch.exp(0.5 * logvar) * ep) * 0.18215) + 0.6 * nz
    e4 = rel_l2(lat, refl)
    torch.cuda.synchronize()
    return max(e1, e2, float(e3), e4 * 1e-0 if e4 > 1e-6 else 0.0), 1e-6


def case_metrics(N=2, H=64, W=96, C=3, kind="noisy", seed=100):
    """GPU PSNR / SSIM (csrc/metrics.cu) must equal the float64 numpy/scipy restatement of the scikit-image calls
    BIT FOR BIT (metric bookkeeping is exact).  Returns the number of values that differ."""
    import numpy as np
    from image_restoration_and_enhancement_b200 import metrics
    rng = np.random.default_rng(seed)
    if kind == "random":
        gt = rng.integers(0, 256, (N, H, W, C), dtype=np.uint8)
        pred = rng.integers(0, 256, (N, H, W, C), dtype=np.uint8)
    else:
        yy, xx = np.mgrid[0:H, 0:W]
        base = np.stack([127 + 100 * np.sin(xx / (5.0 + c) + n) * np.cos(yy / (7.0 + n)) for n in range(N)
                         for c in range(C)]).reshape(N, C, H, W).transpose(0, 2, 3, 1)
        gt = np.clip(base, 0, 255).astype(np.uint8)
        if kind == "identical":
            pred = gt.copy()
        elif kind == "constant":
            gt = np.full((N, H, W, C), 37, np.uint8)
            pred = np.full((N, H, W, C), 200, np.uint8)
        else:
            pred = np.clip(gt.astype(np.float64) + rng.normal(0, 6.5, gt.shape), 0, 255).astype(np.uint8)
    calc = metrics.MetricsCalculator(use_lpips=False)
    want_p = [calc.calculate_psnr(pred[n], gt[n]) for n in range(N)]
    want_s = [calc.calculate_ssim(pred[n], gt[n]) for n in range(N)]
    got_p, got_s = metrics.psnr_ssim_device(torch.from_numpy(pred).to(DEV), torch.from_numpy(gt).to(DEV))
    bad = sum(np.float64(a).tobytes() != np.float64(b).tobytes() for a, b in zip(want_p + want_s, got_p + got_s))
    if bad:
        print("metrics mismatch", kind, (N, H, W, C), list(zip(want_p, got_p)), list(zip(want_s, got_s)))
    return float(bad), 0.0


def case_metrics_calculator(seed=110):
    """MetricsCalculator(device="cuda").calculate_all == the CPU calculator, including the cv2.resize of a prediction
    whose shape differs from the ground truth (src/metrics.py:85-86) and 2-D (grayscale) inputs."""
    import numpy as np
    from image_restoration_and_enhancement_b200 import metrics
    rng = np.random.default_rng(seed)
    cpu, gpu = metrics.MetricsCalculator(use_lpips=False), metrics.MetricsCalculator(use_lpips=False, device="cuda")
    bad = 0
    for gshape, pshape in (((96, 80, 3), (96, 80, 3)), ((96, 80, 3), (48, 40, 3)), ((64, 64), (64, 64))):
        gt = rng.integers(0, 256, gshape, dtype=np.uint8)
        pred = rng.integers(0, 256, pshape, dtype=np.uint8)
        a, b = cpu.calculate_all(pred, gt), gpu.calculate_all(pred, gt)
        bad += sum(np.float64(a[k]).tobytes() != np.float64(b[k]).tobytes() for k in ("psnr", "ssim"))
    return float(bad), 0.0


CASES = {
    "metrics_calculator_gpu": lambda: case_metrics_calculator(),
    # ---- CLIP text encoder pieces (causal attention on the tcgen05 kernel, embedding gather, quick_gelu)
    "attn_causal_clip_77": lambda: case_attention(B=2, heads=12, d=64, Nq=77, seed=120, causal=True),
    "attn_causal_multi_tile": lambda: case_attention(B=1, heads=3, d=64, Nq=300, seed=121, causal=True),
    "attn_causal_d40": lambda: case_attention(B=1, heads=2, d=40, Nq=130, seed=122, causal=True),
    "clip_glue": lambda: case_clip_glue(),

    # ---- PSNR / SSIM on the GPU: bit equality with the float64 CPU bookkeeping
    "metrics_512_noisy": lambda: case_metrics(N=2, H=512, W=512, kind="noisy", seed=100),
    "metrics_512_random": lambda: case_metrics(N=1, H=512, W=512, kind="random", seed=101),
    "metrics_identical": lambda: case_metrics(N=1, H=64, W=64, kind="identical", seed=102),
    "metrics_constant": lambda: case_metrics(N=1, H=40, W=40, kind="constant", seed=103),
    "metrics_ragged": lambda: case_metrics(N=3, H=37, W=50, kind="noisy", seed=104),
    "metrics_min_7x7": lambda: case_metrics(N=2, H=7, W=7, kind="random", seed=105),
    "metrics_wide_multi_chunk": lambda: case_metrics(N=1, H=24, W=2000, kind="noisy", seed=106),
    "metrics_tall_gray": lambda: case_metrics(N=2, H=700, W=300, C=1, kind="noisy", seed=107),
    "metrics_8x13": lambda: case_metrics(N=1, H=8, W=13, kind="random", seed=108),

    # --- tcgen05 GEMM, plain linear layers
    "linear_basic": lambda: case_linear(),
    "linear_ragged_n320": lambda: case_linear(M=300, K=320, N=320, seed=1),
    "linear_n128": lambda: case_linear(M=512, K=256, N=128, seed=2),
    "linear_n64": lambda: case_linear(M=130, K=64, N=64, seed=3),
    "linear_n4_f32": lambda: case_linear(M=200, K=128, N=4, f32_out=True, seed=4),
    "linear_n8_bias_f32": lambda: case_linear(M=200, K=512, N=8, f32_out=True, seed=41),
    "linear_res_f32_silu": lambda: case_linear(M=256, K=192, N=320, res="f32", act=RG_ACT_SILU, seed=5),
    "linear_res_bf16_f32out": lambda: case_linear(M=256, K=1280, N=640, res="bf16", f32_out=True, seed=6),
    "linear_scale_nobias": lambda: case_linear(M=4096, K=512, N=4096, bias=False, scale=0.044, seed=7),
    "linear_tiny_m": lambda: case_linear(M=16, K=1280, N=1280, seed=8),
    "linear_long_k": lambda: case_linear(M=1024, K=5120, N=1280, seed=9),
    "geglu": lambda: case_geglu(),
    # --- implicit-GEMM convolutions
    "conv3x3_small": lambda: case_conv(),
    "conv3x3_unet64": lambda: case_conv(N=2, H=64, W=64, Cin=320, Cout=320, bias_n=True, seed=11),
    "conv3x3_unet8_n3": lambda: case_conv(N=3, H=8, W=8, Cin=1280, Cout=1280, res="f32", f32_out=True, seed=12),
    "conv3x3_ragged": lambda: case_conv(N=1, H=41, W=62, Cin=128, Cout=128, seed=13),
    "conv3x3_stride2": lambda: case_conv(N=2, H=32, W=32, Cin=320, Cout=320, stride=2, seed=14),
    "conv3x3_stride2_odd": lambda: case_conv(N=1, H=41, W=31, Cin=64, Cout=64, stride=2, seed=15),
    "conv3x3_vae_asym": lambda: case_conv(N=1, H=64, W=64, Cin=128, Cout=128, asym=True, seed=16),
    "conv3x3_shortcut": lambda: case_conv(N=2, H=16, W=16, Cin=640, Cout=640, x2c=1920, res=None, f32_out=True, seed=17),
    "conv1x1": lambda: case_conv(N=2, H=32, W=32, Cin=640, Cout=640, k=1, pad=0, res="f32", f32_out=True, seed=18),
    "conv3x3_vae512_cout3": lambda: case_conv(N=1, H=128, W=128, Cin=128, Cout=3, f32_out=True, seed=19),
    "conv3x3_wide": lambda: case_conv(N=1, H=8, W=256, Cin=64, Cout=64, seed=20),
    "conv3x3_both_out_res": lambda: case_conv(N=2, H=32, W=32, Cin=640, Cout=640, res="f32", both_out=True, seed=21),
    "conv3x3_both_out_shortcut": lambda: case_conv(N=2, H=16, W=16, Cin=320, Cout=640, x2c=320, both_out=True, seed=22),
    "conv3x3_vae_res_bf16": lambda: case_conv(N=1, H=64, W=64, Cin=256, Cout=256, res="bf16", seed=23),
    "conv3x3_vae_cout512": lambda: case_conv(N=1, H=32, W=32, Cin=512, Cout=512, seed=24),
    "conv3x3_long_k_2chunk": lambda: case_conv(N=4, H=32, W=32, Cin=1280, Cout=640, bias_n=True, seed=25),
    "conv2x2_parity_strided_out": case_conv_strided_out,
    # --- the four parity kernels in one launch (parities = 4): in-cluster split-K (8x8), one wave (16x16), two column chunks (32x32), batch 16
    "upconv_one_launch_16x16": case_upsample_conv_one_launch,
    "upconv_one_launch_8x8_splitk": lambda: case_upsample_conv_one_launch(N=2, H=8, W=8, seed=151),
    "upconv_one_launch_32x32_c640": lambda: case_upsample_conv_one_launch(N=2, H=32, W=32, Cin=640, Cout=640, seed=152),
    "upconv_one_launch_b16_ragged": lambda: case_upsample_conv_one_launch(N=16, H=7, W=9, Cin=320, Cout=320, seed=153),
    "conv3x3_unet8_splitk_b16": lambda: case_conv(N=16, H=8, W=8, Cin=1280, Cout=1280, bias_n=True, f32_out=True, seed=27),
    "conv3x3_unet8_splitk_shortcut": lambda: case_conv(N=16, H=8, W=8, Cin=1280, Cout=1280, x2c=2560, f32_out=True, seed=28),
    # --- deterministic split-K (epilogue_splitk): 8 / 4 / 3 / 2 slices, every epilogue flavour, ragged tiles
    "splitk8_8x8_res_f32": lambda: case_conv(N=1, H=8, W=8, Cin=1280, Cout=1280, res="f32", f32_out=True, seed=131),
    "splitk8_8x8_n2_biasn_bf16": lambda: case_conv(N=2, H=8, W=8, Cin=1280, Cout=1280, bias_n=True, seed=132),
    "splitk8_ragged_7x9": lambda: case_conv(N=3, H=7, W=9, Cin=1280, Cout=320, res="bf16", seed=133),
    "splitk4_16x16_both_out": lambda: case_conv(N=2, H=16, W=16, Cin=1280, Cout=1280, res="f32", both_out=True, seed=134),
    "splitk4_16x16_shortcut_k2560": lambda: case_conv(N=2, H=16, W=16, Cin=1280, Cout=1280, x2c=2560, f32_out=True, seed=135),
    "splitk2_32x32_bf16": lambda: case_conv(N=2, H=32, W=32, Cin=1280, Cout=640, bias_n=True, seed=136),
    "splitk_stride2_16x16": lambda: case_conv(N=2, H=32, W=32, Cin=1280, Cout=1280, stride=2, f32_out=True, seed=137),
    "splitk_reproducible_bitwise": case_splitk_invariance,
    "splitk2_8x8_b16_res_both": lambda: case_conv(N=16, H=8, W=8, Cin=1280, Cout=1280, res="f32", both_out=True, seed=138),
    "nosplit_16x16_b24_biasn": lambda: case_conv(N=24, H=16, W=16, Cin=1280, Cout=1280, bias_n=True, seed=139),
    "conv_relu_epilogues": case_conv_relu,
    "maxpool3x3s2": case_maxpool,
    "lpips_layer": case_lpips_layer,
    "linear_ff_out_res_f32_bf16out": lambda: case_linear(M=4096, K=1280, N=320, res="f32", seed=26),
    # --- attention
    "attn_self_d40": lambda: case_attention(B=2, heads=8, d=40, Nq=1024),
    "attn_self_d40_4096": lambda: case_attention(B=1, heads=8, d=40, Nq=4096, seed=21),
    "attn_self_d80": lambda: case_attention(B=2, heads=8, d=80, Nq=1024, seed=22),
    "attn_self_d160": lambda: case_attention(B=2, heads=8, d=160, Nq=256, seed=23),
    "attn_self_d160_64": lambda: case_attention(B=3, heads=8, d=160, Nq=64, seed=24),
    "attn_cross_d40": lambda: case_attention(B=2, heads=8, d=40, Nq=4096, Nk=77, seed=25, fused_qkv=False),
    "attn_cross_d160": lambda: case_attention(B=2, heads=8, d=160, Nq=64, Nk=77, seed=26, fused_qkv=False),
    "attn_ragged": lambda: case_attention(B=1, heads=8, d=40, Nq=2542, seed=27),
    "attn_d64": lambda: case_attention(B=1, heads=4, d=64, Nq=300, seed=28),
    "attn_f16_self_d40": lambda: case_attention(B=2, heads=8, d=40, Nq=4096, seed=41, dtype=torch.float16),
    "attn_f16_self_d80": lambda: case_attention(B=2, heads=8, d=80, Nq=1024, seed=42, dtype=torch.float16),
    "attn_f16_self_d160": lambda: case_attention(B=3, heads=8, d=160, Nq=256, seed=43, dtype=torch.float16),
    "attn_f16_cross_d40": lambda: case_attention(B=2, heads=8, d=40, Nq=4096, Nk=77, seed=44, fused_qkv=False, dtype=torch.float16),
    "attn_f16_cross_d160": lambda: case_attention(B=2, heads=8, d=160, Nq=64, Nk=77, seed=45, fused_qkv=False, dtype=torch.float16),
    "attn_f16_ragged": lambda: case_attention(B=1, heads=8, d=80, Nq=651, seed=46, dtype=torch.float16),
    "attn_f16_large_logits": lambda: case_attention_large_logits(),
    "linear_f16_out": lambda: case_linear_f16(),
    "attn_d128": lambda: case_attention(B=1, heads=2, d=128, Nq=700, seed=29),
    "attn_many_items": lambda: case_attention(B=4, heads=8, d=40, Nq=2048, seed=30),       # > 148 items: persistent loop
    # --- norms
    "gn_f32_silu": lambda: case_groupnorm(),
    "gn_concat_raw": lambda: case_groupnorm(N=2, H=8, W=8, C1=1280, C2=640, raw=True, seed=31),
    "gn_bf16_nosilu_eps6": lambda: case_groupnorm(N=1, H=64, W=64, C1=128, in_f32=False, silu=False, eps=1e-6, seed=32),
    "gn_big": lambda: case_groupnorm(N=1, H=256, W=256, C1=128, in_f32=False, seed=33),
    "gn_b16_f32": lambda: case_groupnorm(N=16, H=32, W=32, C1=640, seed=35),
    "gn_b16_bf16_tiny": lambda: case_groupnorm(N=16, H=8, W=8, C1=1280, in_f32=False, seed=36),
    "gn_b16_c2560": lambda: case_groupnorm(N=16, H=8, W=8, C1=1280, C2=1280, raw=True, seed=37),
    "gn_batch_invariant_bitwise": case_groupnorm_shapes_agree,
    # 64x64-level and concatenated shapes at small batch (two-pass kernels; the one-pass kernel covers slices <= 96 KB)
    "gn_64x64_c320_f32_n2": lambda: case_groupnorm(N=2, H=64, W=64, C1=320, seed=38),
    "gn_64x64_concat640_bf16_n3": lambda: case_groupnorm(N=3, H=64, W=64, C1=320, C2=320, in_f32=False, raw=True, seed=39),
    "gn_32x32_concat1280_f32_n2": lambda: case_groupnorm(N=2, H=32, W=32, C1=640, C2=640, raw=True, seed=40),
    # --- more 64x64-level shapes (written for the cluster one-pass experiment, kept for the two-pass kernels)
    "gn_cluster_concat640_f32_raw": lambda: case_groupnorm(N=2, H=64, W=64, C1=320, C2=320, raw=True, seed=141),
    "gn_cluster_ragged_62x41": lambda: case_groupnorm(N=3, H=62, W=41, C1=320, silu=False, seed=142),
    "gn_cluster_c320_bf16_n16": lambda: case_groupnorm(N=16, H=64, W=64, C1=320, in_f32=False, seed=143),
    "gn_cluster_48x48_c1280": lambda: case_groupnorm(N=1, H=48, W=48, C1=1280, seed=144),
    "gn_cluster_batch_invariant": lambda: case_groupnorm_shapes_agree(seed=145, N=5, H=64, W=64, C1=320, C2=320),
    "ln_f32_320": lambda: case_layernorm(),
    "ln_bf16_1280": lambda: case_layernorm(rows=333, C=1280, in_f32=False, seed=42),
    "softmax_rows": lambda: case_softmax_rows(),
    # --- glue
    "timestep_embedding": case_timestep_embedding,
    "sched_cfg": lambda: case_sched(True),
    "sched_nocfg": lambda: case_sched(False, seed=61),
    "im2col_conv_in": case_im2col,
    "upsample_nearest": case_upsample,
    "layout_io": case_layout_and_io,
}
