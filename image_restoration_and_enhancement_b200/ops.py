"""Tensor-level wrappers over the C ABI (include/restoragen.h).

torch is used for device memory and streams only: every function here enqueues hand-written
sm_100a kernels from librestoragen.so on torch's current CUDA stream (so the calls can be captured
by ``torch.cuda.graph``) and returns torch tensors that own the output memory.

Activations are channels-last: ``[N, H, W, C]`` (or ``[rows, C]`` for token matrices).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (RG_ACT_GEGLU, RG_ACT_NONE, RG_ACT_RELU, RG_ACT_SILU, RG_DT_BF16, RG_DT_F16, RG_DT_F32, RgAct, RgAttn, RgConv, RgGn,
                   RgSched, check)

# `bf16` is the build's 16-bit OPERAND dtype: torch.bfloat16 with librestoragen.so, torch.float16 in the fp16 parity mode
# (RESTORAGEN_OPERAND_DTYPE=fp16 -> librestoragen_f16.so, see _lib.py); the C ABI calls it RG_DT_BF16 either way
bf16, f32, f16 = getattr(torch, _lib.OPERAND_DTYPE_NAME), torch.float32, torch.float16
OPERAND_DTYPE = bf16
GN_MAX_IMAGES = 1024      # RG_GN_MAX_IMAGES: fixed-size counter area in front of the GroupNorm workspace
GN_MAX_BLOCKS = 256       # RG_GN_MAX_BLOCKS in include/restoragen.h

# When set to a list, conv2d / attention append (start_event, end_event, algorithmic_flops, kind) per launch
# (bench.py's roofline leg); None in normal operation.
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(e0, work, kind, desc=""):
    """``work``: algorithmic FLOPs for gemm / attention, algorithmic bytes for the HBM-bound kernels."""
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((e0, e1, work, kind, desc))


_GN_WS: dict = {}
GN_WS_MAX_IMAGES = 64     # images per GroupNorm call served by the shared workspace (UNet batch 64 = 32 images under CFG)


def _gn_workspace(device, N: int, groups: int) -> torch.Tensor:
    """One zero-initialised workspace per device, allocated ONCE at its maximum size and never replaced: captured CUDA
    graphs keep its address, and the self-resetting arrival counters at its head stay valid.  Larger batches get a
    private zeroed workspace per call (not graph-persistent, so they are refused while a graph is being captured)."""
    need = GN_MAX_IMAGES + N * groups * 2 + N * GN_MAX_BLOCKS * groups * 2
    if N > GN_WS_MAX_IMAGES or groups > 32:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError(f"GroupNorm batch {N} exceeds the graph-persistent workspace ({GN_WS_MAX_IMAGES} images)")
        return torch.zeros((need,), dtype=f32, device=device)
    key = (device.type, device.index)
    ws = _GN_WS.get(key)
    if ws is None:
        full = GN_MAX_IMAGES + GN_WS_MAX_IMAGES * 32 * 2 + GN_WS_MAX_IMAGES * GN_MAX_BLOCKS * 32 * 2
        ws = _GN_WS[key] = torch.zeros((full,), dtype=f32, device=device)
    return ws


# Deterministic split-K workspace of rg_conv2d (include/restoragen.h: rg_conv_t.splitk_ws): arrival counters + fp32 partial
# tiles (at most one 160 KB partial per CTA pair: 12 MB).  One zero-initialised buffer per device, allocated once (captured
# graphs keep its address).  SPLITK = False switches the split off (A/B measurements only).
SPLITK = True
SPLITK_WS_BYTES = (256 << 10) + (32 << 20)
_SPLITK_WS: dict = {}


class splitk:
    """``with ops.splitk(False):`` -- run (and capture) GEMMs without the K split.  The split follows the number of output
    tiles, i.e. the batch size, so results are reproducible per batch size; with it off every kernel of the library is
    batch-invariant bit for bit (what the sharded sweep's bookkeeping contract needs)."""

    def __init__(self, enabled: bool):
        self.enabled = enabled

    def __enter__(self):
        global SPLITK
        self.prev, SPLITK = SPLITK, self.enabled
        return self

    def __exit__(self, *exc):
        global SPLITK
        SPLITK = self.prev
        return False


def _splitk_workspace(device) -> torch.Tensor:
    key = (device.type, device.index)
    ws = _SPLITK_WS.get(key)
    if ws is None:
        ws = _SPLITK_WS[key] = torch.zeros((SPLITK_WS_BYTES,), dtype=torch.uint8, device=device)
    return ws


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _dt(t: torch.Tensor) -> int:
    if t.dtype == bf16:
        return RG_DT_BF16
    if t.dtype == f32:
        return RG_DT_F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _act_desc(x: torch.Tensor) -> RgAct:
    assert x.dtype == bf16 and x.dim() == 4 and x.stride(3) == 1, "expect bf16 [N,H,W,C] with contiguous C"
    N, H, W, Cc = x.shape
    sw = x.stride(2) if W > 1 else max(x.stride(2), (Cc + 7) // 8 * 8)
    sh = x.stride(1) if H > 1 else max(x.stride(1), sw * W)       # size-1 dims: any sane non-zero pitch
    sn = x.stride(0) if N > 1 else max(x.stride(0), sh * H)
    return RgAct(x.data_ptr(), N, H, W, Cc, sn, sh, sw)


def launch_count() -> int:
    return int(_lib.load().rg_launch_count())


def conv2d(x: torch.Tensor, w: torch.Tensor, *, kh: int = 1, kw: int = 1, stride: int = 1, pad_t: int = 0,
           pad_l: int = 0, OH: int | None = None, OW: int | None = None, x2: torch.Tensor | None = None,
           bias: torch.Tensor | None = None, bias_n: torch.Tensor | None = None, res: torch.Tensor | None = None,
           out_bf16: torch.Tensor | bool | None = None, out_f32: torch.Tensor | bool | None = None,
           act: int = RG_ACT_NONE, scale: float = 1.0, out_strides: tuple | None = None, w_ld: int = 0,
           out_half: torch.dtype = bf16, parities: int = 0):
    """Implicit-GEMM convolution / linear (rg_conv2d).  ``x``: bf16 [N,H,W,C]; ``w``: bf16 [Cout, kh*kw*C (+C2)].

    ``out_bf16`` / ``out_f32``: True to allocate a contiguous [N,OH,OW,Cout'] output, or a tensor to write into
    (with ``out_strides`` = element strides (n, h, w) when it is not contiguous).  ``out_half``: element type of the
    16-bit output (bf16, or fp16 for the attention operands).  Returns (out_bf16, out_f32).
    ``parities=4``: the parity-split "nearest-2x upsample + 3x3 conv" in one launch -- ``w`` = [4 * Cout, 4 * C] (the four 2x2
    kernels stacked in (py, px) order), OH x OW = the input grid, ``out_f32`` = the full [N, 2 OH, 2 OW, Cout] tensor.
    """
    lib = _lib.load()
    N, H, W, Cin = x.shape
    OH = H if OH is None else OH
    OW = W if OW is None else OW
    Cout = w.shape[0] // 4 if parities == 4 else w.shape[0]
    ktot = kh * kw * Cin + (x2.shape[3] if x2 is not None else 0)
    assert w.dtype == bf16 and w.stride(1) == 1 and w.shape[1] == ktot, (w.shape, ktot)
    if w_ld == 0 and w.stride(0) != ktot:
        w_ld = w.stride(0)
    Cw = Cout // 2 if act == RG_ACT_GEGLU else Cout
    if out_bf16 is True:
        out_bf16 = torch.empty((N, OH, OW, Cw), dtype=out_half, device=x.device)
    if out_f32 is True:
        out_f32 = torch.empty((N, OH, OW, Cw), dtype=f32, device=x.device)
    if out_bf16 is False:
        out_bf16 = None
    if out_f32 is False:
        out_f32 = None
    if parities == 4:
        assert isinstance(out_f32, torch.Tensor) and out_f32.shape == (N, 2 * OH, 2 * OW, Cout) and out_bf16 is None
        if out_strides is None:
            out_strides = (out_f32.stride(0), out_f32.stride(1), out_f32.stride(2))
    if out_strides is None:
        out_strides = (OH * OW * Cw, OW * Cw, Cw)
    p = RgConv()
    p.x = _act_desc(x)
    p.kh, p.kw, p.stride, p.pad_t, p.pad_l, p.OH, p.OW = kh, kw, stride, pad_t, pad_l, OH, OW
    if x2 is not None:
        p.has_x2 = 1
        p.x2 = _act_desc(x2)
    p.w, p.w_ld, p.Cout = w.data_ptr(), w_ld, Cout
    if bias is not None:
        assert bias.dtype == f32 and bias.is_contiguous() and bias.numel() == Cout
        p.bias = bias.data_ptr()
    if bias_n is not None:
        assert bias_n.dtype == f32 and bias_n.dim() == 2 and bias_n.stride(1) == 1 and bias_n.shape == (N, Cout)
        p.bias_n, p.bias_n_ld = bias_n.data_ptr(), bias_n.stride(0)
    if res is not None:
        p.res, p.res_dtype = res.data_ptr(), _dt(res)
    if out_bf16 is not None:
        assert out_bf16.dtype in (bf16, f16)
        p.out_bf16 = out_bf16.data_ptr()
        p.out16_dtype = RG_DT_F16 if (out_bf16.dtype == f16 and bf16 is not f16) else RG_DT_BF16
    if out_f32 is not None:
        assert out_f32.dtype == f32
        p.out_f32 = out_f32.data_ptr()
    p.out_stride_n, p.out_stride_h, p.out_stride_w = out_strides
    p.act, p.scale = act, scale
    p.parities = parities
    if SPLITK:
        ws = _splitk_workspace(x.device)
        p.splitk_ws, p.splitk_ws_bytes = ws.data_ptr(), ws.numel()
    e0 = _prof_begin()
    check(lib.rg_conv2d(C.byref(p), _stream()), "rg_conv2d")
    _prof_end(e0, 2.0 * N * OH * OW * w.shape[0] * ktot, "gemm",
              f"conv{kh}x{kw}s{stride} M={N * OH * OW} ({N}x{OH}x{OW}) N={Cout} K={ktot} act={act}"
              f"{' res' if res is not None else ''}{' f32out' if out_f32 is not None else ''}")
    return out_bf16, out_f32


def linear(x: torch.Tensor, w: torch.Tensor, images: int = 1, **kw):
    """x bf16 [M, K] (row stride arbitrary multiple of 8) -> [M, N]; same epilogue options as conv2d.
    ``images``: the M rows are ``images`` equal groups of tokens (one per image): the kernel then tiles the rows per
    image ([images, 1, rows, K] view), like the convolutions do."""
    M, K = x.shape
    assert M % images == 0
    rows = M // images
    x4 = x.as_strided((images, 1, rows, K), (rows * x.stride(0), 0, x.stride(0), 1))
    ob, of = conv2d(x4, w, **kw)
    return (None if ob is None else ob.view(M, -1)), (None if of is None else of.view(M, -1))


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, out: torch.Tensor | None = None,
              causal: bool = False):
    """q [B,Nq,H,d], k/v [B,Nk,H,d] bf16 or fp16 views (d contiguous) -> out bf16 [B,Nq,H,d] contiguous.
    ``causal``: query i attends keys 0..i (the CLIP text encoder; bf16, d <= 64)."""
    lib = _lib.load()
    B, Nq, Hh, d = q.shape
    Nk = k.shape[1]
    for t in (q, k, v):
        assert t.dtype == q.dtype and t.dtype in (bf16, f16) and t.stride(3) == 1
    if out is None:
        out = torch.empty((B, Nq, Hh, d), dtype=bf16, device=q.device)
    p = RgAttn()
    p.q, p.k, p.v, p.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    p.B, p.heads, p.d, p.Nq, p.Nk = B, Hh, d, Nq, Nk
    p.q_stride_b, p.q_stride_t, p.q_stride_h = q.stride(0), q.stride(1), q.stride(2)
    p.k_stride_b, p.k_stride_t, p.k_stride_h = k.stride(0), k.stride(1), k.stride(2)
    p.v_stride_b, p.v_stride_t, p.v_stride_h = v.stride(0), v.stride(1), v.stride(2)
    p.o_stride_b, p.o_stride_t, p.o_stride_h = out.stride(0), out.stride(1), out.stride(2)
    p.scale = scale
    # fp16 kernels (ones-column row sums) exist for the UNet's head dims; everything else takes the operand-dtype path
    p.dtype = RG_DT_F16 if (q.dtype == f16 and (bf16 is not f16 or (not causal and d in (40, 80, 160)))) else RG_DT_BF16
    p.causal = 1 if causal else 0
    e0 = _prof_begin()
    check(lib.rg_attention(C.byref(p), _stream()), "rg_attention")
    _prof_end(e0, 4.0 * B * Hh * Nq * Nk * d, "attention", f"B={B} H={Hh} Nq={Nq} Nk={Nk} d={d}")
    return out


def softmax_rows_(x: torch.Tensor):
    assert x.dtype == bf16 and x.dim() == 2 and x.stride(1) == 1
    check(_lib.load().rg_softmax_rows(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), _stream()), "rg_softmax_rows")
    return x


def groupnorm(x1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, groups: int = 32, eps: float = 1e-5,
              silu: bool = False, x2: torch.Tensor | None = None, want_raw: bool = False,
              sums: torch.Tensor | None = None):
    """GroupNorm (+SiLU) over channels-last x1 (optionally concatenated with x2 along C).
    Returns (y bf16 [N,H,W,C1+C2], raw bf16 copy of the concatenated input or None)."""
    lib = _lib.load()
    N, H, W, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[3]
    assert x1.is_contiguous() and (x2 is None or (x2.is_contiguous() and x2.dtype == x1.dtype))
    Ct = C1 + C2
    y = torch.empty((N, H, W, Ct), dtype=bf16, device=x1.device)
    raw = torch.empty((N, H, W, Ct), dtype=bf16, device=x1.device) if want_raw else None
    if sums is None:
        # RG_GN_WORKSPACE_FLOATS: counters | (mean, rstd) | per-block partials.  The arrival counters must be zero on
        # entry and reset themselves, so one zero-initialised workspace per device serves every call of a stream.
        assert N <= GN_MAX_IMAGES
        sums = _gn_workspace(x1.device, N, groups)
    p = RgGn()
    p.x1, p.C1, p.x2, p.C2 = x1.data_ptr(), C1, _ptr(x2), C2
    p.in_dtype, p.N, p.HW, p.groups, p.eps = _dt(x1), N, H * W, groups, eps
    p.gamma, p.beta, p.sums = gamma.data_ptr(), beta.data_ptr(), sums.data_ptr()
    p.y, p.raw, p.silu = y.data_ptr(), _ptr(raw), int(silu)
    e0 = _prof_begin()
    check(lib.rg_groupnorm(C.byref(p), _stream()), "rg_groupnorm")      # one-pass kernel or stats + apply
    isz = x1.element_size()
    _prof_end(e0, float(N * H * W * Ct) * (2 * isz + 2 + (2 if want_raw else 0)), "groupnorm",
              f"N={N} HW={H * W} C={Ct} in={'f32' if isz == 4 else 'bf16'} silu={int(silu)} raw={int(want_raw)}")
    return y, raw


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    assert x.is_contiguous() and x.dim() == 2
    y = torch.empty(x.shape, dtype=bf16, device=x.device)
    e0 = _prof_begin()
    check(_lib.load().rg_layernorm(x.data_ptr(), _dt(x), x.shape[0], x.shape[1], gamma.data_ptr(), beta.data_ptr(),
                                   eps, y.data_ptr(), _stream()), "rg_layernorm")
    _prof_end(e0, float(x.numel()) * (x.element_size() + 2), "layernorm", f"rows={x.shape[0]} C={x.shape[1]}")
    return y


def embed_tokens(ids: torch.Tensor, tok: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """CLIPTextEmbeddings: ids int32 [B,T] -> fp32 [B*T, C] = tok[ids] + pos[t]."""
    assert ids.dtype == torch.int32 and ids.is_contiguous() and tok.dtype == f32 and pos.dtype == f32
    B, T = ids.shape
    out = torch.empty((B * T, tok.shape[1]), dtype=f32, device=ids.device)
    check(_lib.load().rg_embed_tokens(ids.data_ptr(), tok.data_ptr(), pos.data_ptr(), B, T, tok.shape[1], tok.shape[0],
                                      out.data_ptr(), _stream()), "rg_embed_tokens")
    return out


def quick_gelu_(x: torch.Tensor) -> torch.Tensor:
    """x * sigmoid(1.702 x), bf16, in place."""
    assert x.dtype == bf16 and x.is_contiguous()
    check(_lib.load().rg_quick_gelu_bf16(x.data_ptr(), x.numel(), _stream()), "rg_quick_gelu_bf16")
    return x


def cast_bf16_f32(x: torch.Tensor) -> torch.Tensor:
    assert x.dtype == bf16 and x.is_contiguous()
    y = torch.empty(x.shape, dtype=f32, device=x.device)
    check(_lib.load().rg_cast_bf16_f32(x.data_ptr(), x.numel(), y.data_ptr(), _stream()), "rg_cast_bf16_f32")
    return y


def timestep_embedding(t: torch.Tensor, dim: int, out: torch.Tensor | None = None) -> torch.Tensor:
    assert t.dtype == f32 and t.dim() == 1
    if out is None:
        out = torch.empty((t.shape[0], dim), dtype=bf16, device=t.device)
    check(_lib.load().rg_timestep_embedding(t.data_ptr(), t.shape[0], dim, out.data_ptr(), _stream()),
          "rg_timestep_embedding")
    return out


def sched_step(eps_uc: torch.Tensor, sample: torch.Tensor, *, ets: torch.Tensor | None, cur: torch.Tensor | None,
               do_cfg: bool, guidance: float, store_slot: int, w, use_cur: bool, save_cur: bool,
               c_sample: float, c_eps: float):
    p = RgSched()
    n = sample.numel()
    assert eps_uc.dtype == f32 and sample.dtype == f32 and eps_uc.numel() == (2 * n if do_cfg else n)
    p.eps_uc, p.sample, p.ets, p.cur_sample = eps_uc.data_ptr(), sample.data_ptr(), _ptr(ets), _ptr(cur)
    p.n, p.do_cfg, p.guidance, p.store_slot = n, int(do_cfg), guidance, store_slot
    for i in range(5):
        p.w[i] = float(w[i])
    p.use_cur, p.save_cur, p.c_sample, p.c_eps = int(use_cur), int(save_cur), c_sample, c_eps
    check(_lib.load().rg_sched_step(C.byref(p), _stream()), "rg_sched_step")


def im2col_small(x: torch.Tensor, N_out: int, ksize: int, stride: int, pad: int, OH: int, OW: int,
                 Kpad: int = 64) -> torch.Tensor:
    """x f32|bf16 [n_mod,H,W,Cin] -> bf16 [N_out, OH, OW, Kpad] (image n reads x[n % n_mod])."""
    n_mod, H, W, Cin = x.shape
    assert x.is_contiguous()
    out = torch.empty((N_out, OH, OW, Kpad), dtype=bf16, device=x.device)
    check(_lib.load().rg_im2col_small(x.data_ptr(), _dt(x), N_out, n_mod, H, W, Cin, ksize, stride, pad, OH, OW, Kpad,
                                      out.data_ptr(), _stream()), "rg_im2col_small")
    return out


def upsample_nearest(x: torch.Tensor, OH: int, OW: int) -> torch.Tensor:
    N, H, W, Cc = x.shape
    assert x.dtype == bf16 and x.is_contiguous()
    y = torch.empty((N, OH, OW, Cc), dtype=bf16, device=x.device)
    check(_lib.load().rg_upsample_nearest(x.data_ptr(), N, H, W, Cc, OH, OW, y.data_ptr(), _stream()),
          "rg_upsample_nearest")
    return y


def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    N, Cc, H, W = x.shape
    assert x.dtype == f32 and x.is_contiguous()
    y = torch.empty((N, H, W, Cc), dtype=f32, device=x.device)
    check(_lib.load().rg_nchw_to_nhwc(x.data_ptr(), N, Cc, H, W, y.data_ptr(), _stream()), "rg_nchw_to_nhwc")
    return y


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    N, H, W, Cc = x.shape
    assert x.dtype == f32 and x.is_contiguous()
    y = torch.empty((N, Cc, H, W), dtype=f32, device=x.device)
    check(_lib.load().rg_nhwc_to_nchw(x.data_ptr(), N, Cc, H, W, y.data_ptr(), _stream()), "rg_nhwc_to_nchw")
    return y


def preprocess_u8(img: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
    """u8 [N,H,W,3] -> f32 [N,H,W,3] in [-1,1]; with ``mask`` f32 [N,H,W] also multiplies by (mask < 0.5)."""
    N, H, W, _ = img.shape
    assert img.dtype == torch.uint8 and img.is_contiguous()
    out = torch.empty((N, H, W, 3), dtype=f32, device=img.device)
    check(_lib.load().rg_preprocess_u8(img.data_ptr(), _ptr(mask), N, H, W, out.data_ptr(), _stream()),
          "rg_preprocess_u8")
    return out


def postprocess_u8(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    N, H, W, ldc = x.shape
    assert x.dtype == f32 and x.is_contiguous()
    if out is None:
        out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=x.device)
    check(_lib.load().rg_postprocess_u8(x.data_ptr(), N, H, W, ldc, out.data_ptr(), _stream()), "rg_postprocess_u8")
    return out


def vae_sample(moments: torch.Tensor, eps_post: torch.Tensor, noise: torch.Tensor | None, scaling: float,
               sqrt_ac: float = 1.0, sqrt_1mac: float = 0.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """moments f32 [N,h,w,8]; eps_post/noise f32 [N,h,w,4] -> latents f32 [N,h,w,4]."""
    N, h, w, ld = moments.shape
    if out is None:
        out = torch.empty((N, h, w, 4), dtype=f32, device=moments.device)
    check(_lib.load().rg_vae_sample(moments.data_ptr(), ld, eps_post.data_ptr(), _ptr(noise), N * h * w, scaling,
                                    int(noise is not None), sqrt_ac, sqrt_1mac, out.data_ptr(), _stream()),
          "rg_vae_sample")
    return out


def pack_unet_input(lat: torch.Tensor, mask: torch.Tensor, masked: torch.Tensor, out: torch.Tensor | None = None):
    N, h, w, _ = lat.shape
    if out is None:
        out = torch.empty((N, h, w, 9), dtype=f32, device=lat.device)
    check(_lib.load().rg_pack_unet_input(lat.data_ptr(), mask.data_ptr(), masked.data_ptr(), N * h * w,
                                         out.data_ptr(), _stream()), "rg_pack_unet_input")
    return out


def pointwise_small(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor | None, scale_in: float = 1.0) -> torch.Tensor:
    """x f32 [..., Cin] -> f32 [..., Cout] with W f32 [Cout, Cin]."""
    assert x.dtype == f32 and x.is_contiguous() and W.dtype == f32 and W.is_contiguous()
    Cout, Cin = W.shape
    out = torch.empty((*x.shape[:-1], Cout), dtype=f32, device=x.device)
    check(_lib.load().rg_pointwise_small(x.data_ptr(), x.numel() // Cin, Cin, Cout, W.data_ptr(), _ptr(b), scale_in,
                                         out.data_ptr(), _stream()), "rg_pointwise_small")
    return out


def mask_nearest(mask: torch.Tensor, h: int, w: int) -> torch.Tensor:
    N, H, W = mask.shape
    assert mask.dtype == f32 and mask.is_contiguous()
    out = torch.empty((N, h, w), dtype=f32, device=mask.device)
    check(_lib.load().rg_mask_nearest(mask.data_ptr(), N, H, W, h, w, out.data_ptr(), _stream()), "rg_mask_nearest")
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    assert x.dtype == f32 and x.is_contiguous()
    y = torch.empty(x.shape, dtype=bf16, device=x.device)
    check(_lib.load().rg_cast_f32_bf16(x.data_ptr(), x.numel(), y.data_ptr(), _stream()), "rg_cast_f32_bf16")
    return y


def maxpool3x3s2(x: torch.Tensor) -> torch.Tensor:
    """MaxPool2d(3, stride=2) on bf16 channels-last [N,H,W,C]."""
    N, H, W, Cc = x.shape
    assert x.dtype == bf16 and x.is_contiguous()
    y = torch.empty((N, (H - 3) // 2 + 1, (W - 3) // 2 + 1, Cc), dtype=bf16, device=x.device)
    check(_lib.load().rg_maxpool3x3s2(x.data_ptr(), N, H, W, Cc, y.data_ptr(), _stream()), "rg_maxpool3x3s2")
    return y


def lpips_layer(f0: torch.Tensor, f1: torch.Tensor, lin: torch.Tensor) -> torch.Tensor:
    """One LPIPS feature level: f0, f1 bf16 [N,H,W,C], lin f32 [C] -> f32 [N, blocks] partial sums (see restoragen.h)."""
    N, H, W, Cc = f0.shape
    assert f0.dtype == bf16 and f1.dtype == bf16 and f0.is_contiguous() and f1.is_contiguous() and f1.shape == f0.shape
    assert lin.dtype == f32 and lin.is_contiguous() and lin.numel() == Cc
    lib = _lib.load()
    out = torch.empty((N, lib.rg_lpips_layer_blocks(H * W)), dtype=f32, device=f0.device)
    check(lib.rg_lpips_layer(f0.data_ptr(), f1.data_ptr(), lin.data_ptr(), N, H * W, Cc, out.data_ptr(), _stream()),
          "rg_lpips_layer")
    return out
